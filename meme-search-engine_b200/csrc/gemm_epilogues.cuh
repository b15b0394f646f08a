// Epilogues for linear layers on the tcgen05 mainloop: each epilogue thread owns one output row and
// receives 32 consecutive fp32 columns at a time.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace mse {

enum { ACT_NONE = 0, ACT_GELU_ERF = 1, ACT_GELU_TANH = 2 };

struct GemmOut {
    float *c32;            // fp32 output [M][ldc] or NULL
    __half *c16;           // fp16 output [M][ldc] or NULL
    uint32_t ldc;
    const float *bias;     // [N] fp32 or NULL
    const __half *res;     // residual [res_rows][ldc] fp16 added after activation, or NULL
    uint32_t res_mod;      // residual row = row % res_mod (position embeddings); 0 = row
    int act;
    int debug_no_store;    // profiling only: run the epilogue math but skip the global stores
    int res_in_place;      // TMA-store path: residual == output buffer -> TMA reduce-add instead of a register add
    // skinny split-K path (gemm_skinny.cuh): per-caller scratch for the partial sums and the tile arrival counters (zeroed once;
    // the kernel leaves the counters at zero).  NULL: no split over K.
    float *splitk_ws;      // [splits][rows rounded to 64][ldws] fp32
    uint32_t *splitk_cnt;  // [n-slices x m-tiles]
    size_t splitk_ws_floats;
    // LayerNorm folded into this GEMM (text_mega.cuh explains the algebra): the A operand is the raw residual stream, the weights carry
    // gamma (k_fold_ln), and the epilogue applies  rstd[row] * (acc - mean[row] * ln_cs[col]) + bias[col]  (bias = b + W beta).
    // ln_stats: [M] (mean, rstd) from k_rowstats; both NULL = plain GEMM.  tcgen05 TMA-store path only.
    const float2 *ln_stats;
    const float *ln_cs;
    int pdl_trigger_at;    // skinny kernels: where the CTA lets the next kernel's CTAs be scheduled (common.cuh): 0 start, 1 after the main loop, 2 never
};

// erf-GELU x * Phi(x) = x / 2 + |x| * t,  t = Phi(|x|) - 1/2 = xa * P(u),  xa = min(|x|, 4.75),  u = 2 xa^2 / 4.75^2 - 1:
// degree-10 minimax fit (weighted by x^2, Lawson iteration; script in DESIGN.md 4.1) evaluated by Horner in fp32.  Max abs error of
// the GELU value against the erf form: 3.8e-6 inside the clamp, |x| (1 - Phi(4.75)) = |x| * 1e-6 beyond it -- far below the fp16
// resolution of the activations it produces.  A MUFU-based erf (rcp + ex2) was measured 35 % slower: the epilogue's eight warps
// share one 16/clk MUFU pipe.  gelu_erf_x2 is the same arithmetic, two columns per instruction with sm_100's packed fp32
// (fma.rn.f32x2 -> FFMA2): 10 issue slots per element instead of 21 -- the fc1 epilogue ran at 0.78 of its tile's MMA time.
#define MSE_GELU_COEFFS(X)                                                                                                             \
    X(-0.0037710467566411787f) X(0.0050159604463918773f) X(-0.007168797293015801f)                                                     \
    X(0.013405637279248876f) X(-0.021649570172050516f) X(0.03040285206127552f) X(-0.040568225240212849f)                               \
    X(0.053235260469440548f) X(-0.073667546748732174f) X(0.14874829418180313f)
static constexpr float kGeluC10 = 0.0012802706565601153f;   // leading coefficient, the others follow in Horner order
static constexpr float kGeluClamp = 4.75f, kGeluUScale = 2.0f / (4.75f * 4.75f);

__device__ __forceinline__ float gelu_erf(float x) {
    const float xa = fminf(fabsf(x), kGeluClamp);
    const float u = fmaf(xa * xa, kGeluUScale, -1.0f);
    float p = kGeluC10;
#define MSE_GELU_STEP(c) p = fmaf(p, u, c);
    MSE_GELU_COEFFS(MSE_GELU_STEP)
#undef MSE_GELU_STEP
    return fmaf(fabsf(x), xa * p, 0.5f * x);
}

namespace f32x2 {
__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack(unsigned long long v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long mul(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long add(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
}  // namespace f32x2

// two columns at once; bit-identical to gelu_erf per element (the same fp32 operations in the same order)
__device__ __forceinline__ void gelu_erf_x2(float &x0, float &x1) {
    const float a0 = fminf(fabsf(x0), kGeluClamp), a1 = fminf(fabsf(x1), kGeluClamp);
    const unsigned long long xa = f32x2::pack(a0, a1), x = f32x2::pack(x0, x1);
    const unsigned long long u = f32x2::fma(f32x2::mul(xa, xa), f32x2::pack(kGeluUScale, kGeluUScale), f32x2::pack(-1.0f, -1.0f));
    unsigned long long p = f32x2::pack(kGeluC10, kGeluC10);
#define MSE_GELU_STEP(c) p = f32x2::fma(p, u, f32x2::pack(c, c));
    MSE_GELU_COEFFS(MSE_GELU_STEP)
#undef MSE_GELU_STEP
    float t0, t1, h0, h1;
    f32x2::unpack(f32x2::mul(xa, p), t0, t1);
    f32x2::unpack(f32x2::mul(x, f32x2::pack(0.5f, 0.5f)), h0, h1);
    x0 = fmaf(fabsf(x0), t0, h0);
    x1 = fmaf(fabsf(x1), t1, h1);
}
__device__ __forceinline__ float gelu_tanh(float x) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    return 0.5f * x * (1.0f + tanhf(k0 * (x + k1 * x * x * x)));
}

template <bool TMA_STORE>
struct LinearEpilogueT {
    static constexpr bool kTmaStore = TMA_STORE;
    GemmOut o;
    uint32_t M, N;
    __device__ __forceinline__ void begin_tile(uint32_t, uint32_t) {}
    // TMA-store path ------------------------------------------------------------------------------------------------------
    // residual that equals the output buffer (x += linear(...)): staged through shared memory by the kernel
    __device__ __forceinline__ const __half *staged_residual() const { return o.res_in_place ? o.res : nullptr; }
    __device__ __forceinline__ uint32_t ldc() const { return o.ldc; }
    __device__ __forceinline__ uint32_t rows() const { return M; }
    __device__ __forceinline__ uint32_t cols() const { return N; }
    // bias / activation / (position-embedding style residual) applied in place on 32 fp32 accumulators
    __device__ __forceinline__ void compute(uint32_t row, uint32_t col0, uint32_t (&v)[32]) {
        const bool fullc = col0 + 32 <= N;
        if (o.ln_stats) {
            // rstd * (acc - mean * cs) + bias  =  acc * rstd + (bias - mean * rstd * cs): two FMAs per element, bias included
            const float2 st = row < M ? __ldg(o.ln_stats + row) : make_float2(0.f, 0.f);
            const float rstd = st.y, nmr = -st.x * st.y;
            if (fullc) {
                const float4 *c4 = (const float4 *)(o.ln_cs + col0), *b4 = (const float4 *)(o.bias + col0);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float4 c = __ldg(c4 + j), b = __ldg(b4 + j);
                    v[4 * j] = __float_as_uint(fmaf(__uint_as_float(v[4 * j]), rstd, fmaf(nmr, c.x, b.x)));
                    v[4 * j + 1] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 1]), rstd, fmaf(nmr, c.y, b.y)));
                    v[4 * j + 2] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 2]), rstd, fmaf(nmr, c.z, b.z)));
                    v[4 * j + 3] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 3]), rstd, fmaf(nmr, c.w, b.w)));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++)
                    if (col0 + j < N) v[j] = __float_as_uint(fmaf(__uint_as_float(v[j]), rstd, fmaf(nmr, o.ln_cs[col0 + j], o.bias[col0 + j])));
            }
        } else if (o.bias) {
            if (fullc) {
                const float4 *b4 = (const float4 *)(o.bias + col0);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float4 b = __ldg(b4 + j);
                    v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + b.x);
                    v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + b.y);
                    v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + b.z);
                    v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + b.w);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) if (col0 + j < N) v[j] = __float_as_uint(__uint_as_float(v[j]) + o.bias[col0 + j]);
            }
        }
        if (o.act == ACT_GELU_ERF) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                float x0 = __uint_as_float(v[j]), x1 = __uint_as_float(v[j + 1]);
                gelu_erf_x2(x0, x1);
                v[j] = __float_as_uint(x0);
                v[j + 1] = __float_as_uint(x1);
            }
        } else if (o.act == ACT_GELU_TANH) {
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = __float_as_uint(gelu_tanh(__uint_as_float(v[j])));
        }
        if (o.res && !o.res_in_place && row < M) {
            const uint32_t rr = o.res_mod ? row % o.res_mod : row;
            const __half *rp = o.res + (size_t)rr * o.ldc + col0;
            if (fullc && o.ldc % 8 == 0) {
                const uint4 *r4 = (const uint4 *)rp;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint4 u = __ldg(r4 + j);
                    const __half2 *h = (const __half2 *)&u;
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float2 t = __half22float2(h[e]);
                        v[8 * j + 2 * e] = __float_as_uint(__uint_as_float(v[8 * j + 2 * e]) + t.x);
                        v[8 * j + 2 * e + 1] = __float_as_uint(__uint_as_float(v[8 * j + 2 * e + 1]) + t.y);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) if (col0 + j < N) v[j] = __float_as_uint(__uint_as_float(v[j]) + __half2float(rp[j]));
            }
        }
    }
    // direct-store path ---------------------------------------------------------------------------------------------------
    __device__ __forceinline__ void columns(uint32_t row, uint32_t col0, const uint32_t (&v)[32]) {
        if (row >= M || col0 >= N) return;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; j++) f[j] = __uint_as_float(v[j]);
        const bool fullc = col0 + 32 <= N;          // all 32 columns in range
        const bool full = fullc && (o.ldc % 8 == 0);  // ... and 16-byte vector access to C / residual rows is aligned
        if (o.bias) {
            if (fullc) {
                const float4 *b4 = (const float4 *)(o.bias + col0);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    float4 b = __ldg(b4 + j);
                    f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) if (col0 + j < N) f[j] += o.bias[col0 + j];
            }
        }
        if (o.act == ACT_GELU_ERF) {
#pragma unroll
            for (int j = 0; j < 32; j++) f[j] = gelu_erf(f[j]);
        } else if (o.act == ACT_GELU_TANH) {
#pragma unroll
            for (int j = 0; j < 32; j++) f[j] = gelu_tanh(f[j]);
        }
        if (o.res) {
            const uint32_t rr = o.res_mod ? row % o.res_mod : row;
            const __half *rp = o.res + (size_t)rr * o.ldc + col0;
            if (full) {
                const uint4 *r4 = (const uint4 *)rp;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint4 u = __ldg(r4 + j);
                    const __half2 *h = (const __half2 *)&u;
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        float2 t = __half22float2(h[e]);
                        f[8 * j + 2 * e] += t.x;
                        f[8 * j + 2 * e + 1] += t.y;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) if (col0 + j < N) f[j] += __half2float(rp[j]);
            }
        }
        if (o.debug_no_store) { if (f[0] == 123456.789f) o.c16[0] = __float2half(f[1]); return; }
        if (o.c16) {
            __half *cp = o.c16 + (size_t)row * o.ldc + col0;
            if (full) {
                uint4 *c4 = (uint4 *)cp;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint4 u;
                    __half2 *h = (__half2 *)&u;
#pragma unroll
                    for (int e = 0; e < 4; e++) h[e] = __floats2half2_rn(f[8 * j + 2 * e], f[8 * j + 2 * e + 1]);
                    c4[j] = u;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) if (col0 + j < N) cp[j] = __float2half_rn(f[j]);
            }
        }
        if (o.c32) {
            float *cp = o.c32 + (size_t)row * o.ldc + col0;
            if (full) {
                float4 *c4 = (float4 *)cp;
#pragma unroll
                for (int j = 0; j < 8; j++) c4[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) if (col0 + j < N) cp[j] = f[j];
            }
        }
    }
};

using LinearEpilogue = LinearEpilogueT<false>;

int gemm_f16_tn_dev(int device, const __half *dA, const __half *dB, uint32_t M, uint32_t N, uint32_t K, uint32_t lda, uint32_t ldb,
                    const GemmOut &out, cudaStream_t st);

}  // namespace mse
