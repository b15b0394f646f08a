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

}  // namespace mse
