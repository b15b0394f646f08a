// Non-causal multi-head self-attention for the SigLIP towers (vision S=729, text S=64; dh=72, 16 heads).
//
// Reference op: timm Attention / nn.MultiheadAttention inside the 27 encoder blocks
// (aitemplate/model.py:30-36,42; clip_server.py:98,114): softmax(Q K^T / sqrt(dh)) V per (batch, head).
//
// Layout: qkv is the fused projection output [B*S][3*D] fp16 (q | k | v, heads contiguous inside each);
// out is [B*S][D] fp16 with heads concatenated -- exactly what the out-projection GEMM consumes.
//
// Flash-style streaming kernel: one CTA = 64 query rows of one (b, h) (4 warps x 16 rows); K/V tiles of 64 keys
// are double-buffered through shared memory with cp.async (zero-filled past S); scores never leave registers
// (online softmax in fp32, exp2 with the scale folded in).  dh = 72 is padded to 80 for the QK^T contraction
// (zero columns) and handled as 9 n-tiles of 8 for P V.  Tensor-core math is warp-level mma.sync m16n8k16
// (fp16 in, fp32 accumulate).  10 % of the tower's FLOPs live here; the GEMMs are on tcgen05.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>
#include <stdint.h>

namespace mse {
namespace attn {

static constexpr int kBM = 64;      // queries per CTA
static constexpr int kBN = 64;      // keys per tile
static constexpr int kDH = 72;      // head dim
static constexpr int kDHP = 80;     // head dim padded to a multiple of 16 for QK^T
static constexpr int kPitch = 88;   // smem row pitch in halfs (176 B: conflict-free ldmatrix)
static constexpr int kThreads = 128;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, const void *p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, const void *p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t &r0, uint32_t &r1, const void *p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}
// D(16x8, f32) += A(16x16, f16) * B(16x8, f16)
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *(uint32_t *)&h;
}

struct Smem {
    __half q[kBM][kPitch];
    __half k[2][kBN][kPitch];
    __half v[2][kBN][kPitch];
};

// One tile of 64 query rows of one (b, h).  NT threads take part in the loads and the barriers; warps 0-3 (16 rows each) do the math, so
// the function serves both the 128-thread kernel below and the 512-thread persistent text kernel (text_mega.cuh).
template <int NT>
__device__ __forceinline__ void mha_tile(const __half *__restrict__ qkv, __half *__restrict__ out, Smem &sm, int S, int H, int q0, int h, int b,
                                         float scale_log2e) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool math = warp < kBM / 16;
    const int D = H * kDH;
    const size_t ld = (size_t)3 * D;
    const __half *base = qkv + (size_t)b * S * ld;
    const __half *qg = base + (size_t)h * kDH;
    const __half *kg = base + D + (size_t)h * kDH;
    const __half *vg = base + 2 * D + (size_t)h * kDH;

    // zero the pad columns [72, 88) of every tile row once; cp.async never writes them
    for (int i = tid; i < kBM + 4 * kBN; i += NT) {
        __half *row = i < kBM ? sm.q[i] : (i < kBM + 2 * kBN ? sm.k[(i - kBM) / kBN][(i - kBM) % kBN]
                                                             : sm.v[(i - kBM - 2 * kBN) / kBN][(i - kBM - 2 * kBN) % kBN]);
        *(uint4 *)(row + 72) = make_uint4(0, 0, 0, 0);
        *(uint4 *)(row + 80) = make_uint4(0, 0, 0, 0);
    }
    // Q tile + first K/V tile
    for (int i = tid; i < kBM * 9; i += NT) {
        int r = i / 9, c = i % 9;
        bool ok = q0 + r < S;
        cp_async16(&sm.q[r][c * 8], qg + (size_t)(ok ? q0 + r : 0) * ld + c * 8, ok);
    }
    auto load_kv = [&](int buf, int k0) {
        for (int i = tid; i < kBN * 9; i += NT) {
            int r = i / 9, c = i % 9;
            bool ok = k0 + r < S;
            size_t off = (size_t)(ok ? k0 + r : 0) * ld + c * 8;
            cp_async16(&sm.k[buf][r][c * 8], kg + off, ok);
            cp_async16(&sm.v[buf][r][c * 8], vg + off, ok);
        }
    };
    load_kv(0, 0);
    cp_async_commit();

    const int ntiles = (S + kBN - 1) / kBN;
    float o[9][4];
#pragma unroll
    for (int j = 0; j < 9; j++) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // running max / sum for rows (lane/4) and (lane/4 + 8)
    uint32_t qf[5][4];

    for (int t = 0; t < ntiles; t++) {
        const int buf = t & 1;
        if (t + 1 < ntiles) {
            load_kv(buf ^ 1, (t + 1) * kBN);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (math) {
        if (t == 0) {
#pragma unroll
            for (int ks = 0; ks < 5; ks++)
                ldsm_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], &sm.q[warp * 16 + (lane & 15)][ks * 16 + (lane >> 4) * 8]);
        }
        // S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; j++) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 5; ks++) {
#pragma unroll
            for (int jp = 0; jp < 4; jp++) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4(b0, b1, b2, b3, &sm.k[buf][jp * 16 + (lane & 7) + ((lane >> 4) << 3)][ks * 16 + ((lane >> 3) & 1) * 8]);
                mma16816(s[2 * jp], qf[ks], b0, b1);
                mma16816(s[2 * jp + 1], qf[ks], b2, b3);
            }
        }
        // scale, mask keys >= S, online softmax
        const int kbase = t * kBN + (lane & 3) * 2;
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int kc = kbase + j * 8;
            const bool v0 = kc < S, v1 = kc + 1 < S;
            s[j][0] = v0 ? s[j][0] * scale_log2e : -INFINITY;
            s[j][1] = v1 ? s[j][1] * scale_log2e : -INFINITY;
            s[j][2] = v0 ? s[j][2] * scale_log2e : -INFINITY;
            s[j][3] = v1 ? s[j][3] * scale_log2e : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);  // finite: every tile holds at least one valid key
        const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
        m0 = mn0;
        m1 = mn1;
        float rs0 = 0.f, rs1 = 0.f;
        uint32_t pf[4][4];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            float p0 = exp2f(s[j][0] - mn0), p1 = exp2f(s[j][1] - mn0), p2 = exp2f(s[j][2] - mn1), p3 = exp2f(s[j][3] - mn1);
            rs0 += p0 + p1;
            rs1 += p2 + p3;
            pf[j >> 1][(j & 1) * 2] = pack_half2(p0, p1);
            pf[j >> 1][(j & 1) * 2 + 1] = pack_half2(p2, p3);
        }
        l0 = l0 * c0 + rs0;
        l1 = l1 * c1 + rs1;
#pragma unroll
        for (int j = 0; j < 9; j++) {
            o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1;
        }
        // O += P V   (k = 64 keys in 4 steps, n = 72 in 9 tiles)
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
#pragma unroll
            for (int jp = 0; jp < 4; jp++) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4_t(b0, b1, b2, b3, &sm.v[buf][ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][jp * 16 + (lane >> 4) * 8]);
                mma16816(o[2 * jp], pf[ks], b0, b1);
                mma16816(o[2 * jp + 1], pf[ks], b2, b3);
            }
            uint32_t b0, b1;
            ldsm_x2_t(b0, b1, &sm.v[buf][ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][64]);
            mma16816(o[8], pf[ks], b0, b1);
        }
        }
        __syncthreads();  // everyone is done with buf before the next iteration's prefetch overwrites it
    }
    if (!math) return;
    // finalize: rows (lane/4) and (lane/4 + 8) of this warp
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
    __half *ob = out + (size_t)b * S * D + (size_t)h * kDH + (lane & 3) * 2;
#pragma unroll
    for (int j = 0; j < 9; j++) {
        if (r0 < S) *(__half2 *)(ob + (size_t)r0 * D + j * 8) = __floats2half2_rn(o[j][0] * i0, o[j][1] * i0);
        if (r1 < S) *(__half2 *)(ob + (size_t)r1 * D + j * 8) = __floats2half2_rn(o[j][2] * i1, o[j][3] * i1);
    }
}

// grid: (ceil(S / 64), H, B)
__global__ void __launch_bounds__(kThreads) k_mha_fwd(const __half *__restrict__ qkv, __half *__restrict__ out, int S, int H,
                                                      float scale_log2e) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    pdl_trigger();
    pdl_wait();   // qkv is the previous kernel's output (launched with launch_pdl: common.cuh)
    mha_tile<kThreads>(qkv, out, sm, S, H, blockIdx.x * kBM, blockIdx.y, blockIdx.z, scale_log2e);
}

}  // namespace attn
}  // namespace mse
