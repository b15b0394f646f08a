// Device-side fast_dot (diskann/src/vector.rs:192-306), shared by the graph search / prune kernels.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace mse {

// Rust `(x * 2^32) as i64`: truncate toward zero, saturate, NaN -> 0 (vector.rs:249-250)
__device__ __forceinline__ long long fast_dot_fix(float s) {
    float v = s * 4294967296.0f;
    return (v != v) ? 0ll : __float2ll_rz(v);
}

// reduction tree of vector.rs:241-249 over the 32 per-lane partial sums; every lane returns the final value
__device__ __forceinline__ float fast_dot_reduce(float p) {
    const unsigned full = 0xffffffffu;
    // A_j (lanes 0..7) = p[j] + p[8+j];  B_j (lanes 16..23) = p[16+j] + p[24+j]
    float ab = p + __shfl_down_sync(full, p, 8);
    // pair sums: lane j (even j in 0..7 / 16..23): ab[j] + ab[j+1]
    float pr = ab + __shfl_down_sync(full, ab, 1);
    // e0 = (A0+A1)+(A4+A5), e1 = (A2+A3)+(A6+A7), e2 = (B0+B1)+(B4+B5), e3 = (B2+B3)+(B6+B7)
    float q4 = pr + __shfl_down_sync(full, pr, 4);  // valid at lanes 0,2,16,18
    float e0 = __shfl_sync(full, q4, 0), e1 = __shfl_sync(full, q4, 2), e2 = __shfl_sync(full, q4, 16), e3 = __shfl_sync(full, q4, 18);
    return ((e0 + e1) + e2) + e3;
}

// whole warp cooperates on one dot product; x, y: fp16[d], d % 32 == 0
__device__ __forceinline__ float fast_dot_warp(const __half *__restrict__ x, const __half *__restrict__ y, uint32_t d, int lane) {
    float p = 0.f;
#pragma unroll 4
    for (uint32_t c = lane; c < d; c += 32) p = fmaf(__half2float(x[c]), __half2float(y[c]), p);
    return fast_dot_reduce(p);
}

// same with the query's per-lane slice already in registers (qreg[i] = f32(query[32*i + lane]))
template <int NCH>
__device__ __forceinline__ float fast_dot_warp_q(const float (&qreg)[NCH], const __half *__restrict__ y, int lane) {
    float yv[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) yv[i] = __half2float(y[32 * i + lane]);
    float p = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; i++) p = fmaf(qreg[i], yv[i], p);
    return fast_dot_reduce(p);
}

// Two rows per pass, half a warp per row.  Every lane loads one half2 (elements 2*lane, 2*lane+1 of each 64-element chunk:
// one 128-byte request per row and chunk); lane l < 16 then owns the reference's partial sums 2l and 2l+1 of row 0 and lane
// 16 + l the same partials of row 1, so one shfl_xor(16) per chunk hands each half warp the other half of its row.  Partial i
// still accumulates elements i, i+32, i+64, ... in that order with FMA (vector.rs:212-238), and the reduction below is the
// tree of vector.rs:241-249 re-indexed for two partials per lane -- the scores are bit-identical to fast_dot.
template <int NC2>   // 64-element chunks per row (d / 64), 0 = runtime d
__device__ __forceinline__ void wq_dot2(const float *qs, const __half *__restrict__ r0, const __half *__restrict__ r1, uint32_t d, int lane,
                                        float &f0, float &f1) {
    const unsigned full = 0xffffffffu;
    const int l = lane & 15;
    const bool upper = lane >= 16;
    const uint32_t *a2 = reinterpret_cast<const uint32_t *>(r0) + lane;
    const uint32_t *b2 = reinterpret_cast<const uint32_t *>(r1) + lane;
    const float2 *q2 = reinterpret_cast<const float2 *>(qs);
    float lo = 0.f, hi = 0.f;
    auto step = [&](uint32_t va, uint32_t vb, uint32_t c) {
        const uint32_t recv = __shfl_xor_sync(full, upper ? va : vb, 16);
        const uint32_t first = upper ? recv : va, second = upper ? vb : recv;
        const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&first));
        const float2 f2 = __half22float2(*reinterpret_cast<const __half2 *>(&second));
        const float2 qa = q2[32 * c + l], qb = q2[32 * c + 16 + l];
        lo = fmaf(qa.x, f1.x, lo); hi = fmaf(qa.y, f1.y, hi);
        lo = fmaf(qb.x, f2.x, lo); hi = fmaf(qb.y, f2.y, hi);
    };
    if constexpr (NC2 > 0) {
        uint32_t va[NC2], vb[NC2];
#pragma unroll
        for (int c = 0; c < NC2; c++) { va[c] = __ldg(a2 + 32 * c); vb[c] = __ldg(b2 + 32 * c); }
#pragma unroll
        for (int c = 0; c < NC2; c++) step(va[c], vb[c], (uint32_t)c);
    } else {
        const uint32_t nc = d >> 6;
#pragma unroll 4
        for (uint32_t c = 0; c < nc; c++) step(__ldg(a2 + 32 * c), __ldg(b2 + 32 * c), c);
    }
    lo += __shfl_down_sync(full, lo, 4);             // acc1+acc2 / acc3+acc4 (:241-242): valid at l in {0..3, 8..11}
    hi += __shfl_down_sync(full, hi, 4);
    const float pr = lo + hi;                        // hadd pairs (:243)
    const float q4 = pr + __shfl_down_sync(full, pr, 2);   // lo + hi halves (:244-246): valid at l in {0, 1, 8, 9}
    const int gb = lane & 16;
    const float e0 = __shfl_sync(full, q4, gb), e1 = __shfl_sync(full, q4, gb + 1), e2 = __shfl_sync(full, q4, gb + 8), e3 = __shfl_sync(full, q4, gb + 9);
    const float r = ((e0 + e1) + e2) + e3;           // :247-249
    f0 = __shfl_sync(full, r, 0);
    f1 = __shfl_sync(full, r, 16);
}
// the same as fixed-point scores (vector.rs:249-250)
template <int NC2>
__device__ __forceinline__ void wq_score2(const float *qs, const __half *__restrict__ r0, const __half *__restrict__ r1, uint32_t d, int lane,
                                          long long &s0, long long &s1) {
    float f0, f1;
    wq_dot2<NC2>(qs, r0, r1, d, lane, f0, f1);
    s0 = fast_dot_fix(f0);
    s1 = fast_dot_fix(f1);
}

}  // namespace mse
