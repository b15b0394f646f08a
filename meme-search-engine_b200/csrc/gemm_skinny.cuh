// Skinny GEMM for small batches:  C[M,N] = A[M,K] * B[N,K]^T  with M <= a few hundred rows (one text query is 64 tokens:
// clip_server.py:98 at batch 1, BASELINE configs[0]).
//
// At M = 64 a linear layer is a pass over its weights: 7.96 MB (QKV), 9.9 MB (fc1 / fc2), 2.65 MB (out-proj) per block, 0.826 GB
// per text tower -- HBM-bound (0.126 ms at the copy peak), 2 FLOP per weight byte per row.  The 128 x 256 tcgen05 tiles of
// gemm_sm100.cuh give such a layer 5-17 tiles, i.e. 5-17 SMs pulling on HBM.  Here the N dimension is cut into 16- or 32-column
// slices so that every SM streams its own slice of the weights exactly once (cp.async, 16-byte requests, 6 stages = ~80 KB in
// flight per SM); the 64 x K activation panel is re-read by every CTA from L2.  Math: warp-level mma.sync m16n8k16
// (fp16 x fp16 -> fp32) -- at 2 FLOP/B the tensor pipe idles either way; tcgen05 would add TMEM round trips for no gain.
// Epilogue: GemmOut semantics of gemm_epilogues.cuh (bias, erf/tanh GELU, residual or position embedding, fp16 / fp32 stores).
#pragma once
#include "common.cuh"
#include "gemm_epilogues.cuh"
#include <cuda_fp16.h>
#include <stdint.h>

namespace mse {
namespace skinny {

static constexpr int kBM = 64, kBK = 64, kThreads = 512, kStages = 6;   // 16 warps: 4 row groups x the 4 k16 steps of a 64-wide k tile
static constexpr int kPitch = kBK + 8;   // halfs per shared-memory row (144 B): ldmatrix rows land in distinct banks

template <int BN>
constexpr uint32_t smem_bytes() { return (uint32_t)kStages * (kBM + BN) * kPitch * 2; }

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int n = valid ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void *smem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// grid (ceil(N / BN), ceil(M / 64), splits).  Warp w = 4 * kq + wr: rows 16 wr .. 16 wr + 15 of the CTA's 64 x BN tile, k16 step kq of
// every 64-wide k tile; the four partial sums of a row group are added in a fixed order through shared memory at the end.
//
// Split over K (grid.z > 1): what limits these kernels is bytes moved per CTA over load latency, and with one N slice per CTA two
// thirds of those bytes are the 64-row activation panel every CTA re-reads from L2 (ncu: profiles/r01p_skinny_gemm_ncu_full.md).
// Wider slices (BN = 64 / 128) cut the number of panel readers, and the K range is split across CTAs to keep every SM busy: a CTA
// then moves ~40-110 KB instead of 220-690 KB.  Partial tiles go to an fp32 scratch; the LAST CTA to arrive for a tile (atomic
// counter) adds the partials in split order -- a fixed order, so results do not depend on which CTA came last -- and runs the
// epilogue.
template <int BN>
__global__ void __launch_bounds__(kThreads) k_gemm_skinny(const __half *__restrict__ A, const __half *__restrict__ B, uint32_t M, uint32_t N, uint32_t K,
                                                          uint32_t lda, uint32_t ldb, GemmOut o) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ uint32_t s_last;
    __half *sA = (__half *)smem_raw;                                  // [stage][64][kPitch]
    __half *sB = sA + (size_t)kStages * kBM * kPitch;                  // [stage][BN][kPitch]
    const int tid = threadIdx.x, warp = (tid >> 5) & 3, kq = tid >> 7, lane = tid & 31;
    const uint32_t n0 = blockIdx.x * BN, m0 = blockIdx.y * kBM;
    const uint32_t nk_all = (K + kBK - 1) / kBK, splits = gridDim.z;
    const uint32_t kt0 = (uint32_t)((uint64_t)nk_all * blockIdx.z / splits), kt1 = (uint32_t)((uint64_t)nk_all * (blockIdx.z + 1) / splits);
    const uint32_t nk = kt1 - kt0;

    auto load_a = [&](uint32_t kt, int st) {                           // A: 64 rows x 8 chunks of 16 B = one chunk per thread
        const uint32_t k0 = (kt0 + kt) * kBK;
        __half *a = sA + (size_t)st * kBM * kPitch;
        const int r = tid >> 3, kc = (tid & 7) * 8;
        const bool ok = m0 + r < M && k0 + kc < K;                    // K % 8 == 0: a chunk is entirely inside or outside
        cp_async16(a + r * kPitch + kc, A + (size_t)(ok ? m0 + r : 0) * lda + (ok ? k0 + kc : 0), ok);
    };
    auto load_b = [&](uint32_t kt, int st) {                           // B: BN rows x 8 chunks
        const uint32_t k0 = (kt0 + kt) * kBK;
        __half *b = sB + (size_t)st * BN * kPitch;
#pragma unroll
        for (int i = 0; i < (BN * 8 + kThreads - 1) / kThreads; i++) {
            const int c = tid + i * kThreads, r = c >> 3, kc = (c & 7) * 8;
            if (c < BN * 8) {
                const bool ok = n0 + r < N && k0 + kc < K;
                cp_async16(b + r * kPitch + kc, B + (size_t)(ok ? n0 + r : 0) * ldb + (ok ? k0 + kc : 0), ok);
            }
        }
    };
    auto load_stage = [&](uint32_t kt, int st) { load_a(kt, st); load_b(kt, st); };

    float acc[BN / 8][4];
#pragma unroll
    for (int j = 0; j < BN / 8; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;

    // Programmatic dependent launch (common.cuh): the WEIGHT tiles of the first stages do not depend on the previous kernel, so they are
    // requested before pdl_wait() and stream in while the previous kernel drains; the activation tiles follow after it.  All of these
    // weight requests belong to the first commit group, which every later wait covers.
    if (o.pdl_trigger_at == 0) pdl_trigger();
#pragma unroll
    for (int s = 0; s < kStages - 1; s++)
        if ((uint32_t)s < nk) load_b(s, s);
    pdl_wait();
#pragma unroll
    for (int s = 0; s < kStages - 1; s++) {
        if ((uint32_t)s < nk) load_a(s, s);
        cp_async_commit();
    }
    for (uint32_t kt = 0; kt < nk; kt++) {
        cp_async_wait<kStages - 2>();
        __syncthreads();                                              // stage kt has landed; stage kt-1 is free for the next load
        const uint32_t nxt = kt + kStages - 1;
        if (nxt < nk) load_stage(nxt, nxt % kStages);
        cp_async_commit();
        const __half *a = sA + (size_t)(kt % kStages) * kBM * kPitch + (warp * 16) * kPitch;
        const __half *b = sB + (size_t)(kt % kStages) * BN * kPitch;
        uint32_t af[4];
        ldmatrix_x4(af, a + (lane & 15) * kPitch + kq * 16 + (lane >> 4) * 8);
#pragma unroll
        for (int j = 0; j < BN / 16; j++) {                            // two n8 tiles per ldmatrix.x4
            uint32_t bf[4];
            ldmatrix_x4(bf, b + (j * 16 + (lane & 7) + (lane >> 4) * 8) * kPitch + kq * 16 + ((lane >> 3) & 1) * 8);
            mma_16816(acc[2 * j], af, bf[0], bf[1]);
            mma_16816(acc[2 * j + 1], af, bf[2], bf[3]);
        }
    }
    cp_async_wait<0>();
    if (o.pdl_trigger_at == 1) pdl_trigger();
    __syncthreads();                                                  // the pipeline buffers are free: reuse them for the reduction
    float *red = (float *)smem_raw;                                   // [3][4 warps][BN / 2][32 lanes]
    if (kq > 0) {
#pragma unroll
        for (int j = 0; j < BN / 8; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) red[(((kq - 1) * 4 + warp) * (BN / 2) + j * 4 + e) * 32 + lane] = acc[j][e];   // lane fastest: conflict-free
    }
    __syncthreads();
    if (kq == 0) {
#pragma unroll
        for (int q = 0; q < 3; q++)                                   // fixed order: k16 steps 0, 1, 2, 3
#pragma unroll
            for (int j = 0; j < BN / 8; j++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[j][e] += red[((q * 4 + warp) * (BN / 2) + j * 4 + e) * 32 + lane];
    }
    // thread (kq == 0) holds rows g and g + 8 (g = lane / 4), columns 2 (lane % 4) + {0, 1} of every n8 tile
    const uint32_t g = lane >> 2, t2 = (lane & 3) * 2;
    if (splits > 1) {
        const uint32_t ldws = gridDim.x * BN, rows_ws = gridDim.y * kBM;
        if (kq == 0) {
            float *w = o.splitk_ws + ((size_t)blockIdx.z * rows_ws + m0 + warp * 16 + g) * ldws + n0 + t2;
#pragma unroll
            for (int j = 0; j < BN / 8; j++) {
                *(float2 *)(w + j * 8) = make_float2(acc[j][0], acc[j][1]);
                *(float2 *)(w + (size_t)8 * ldws + j * 8) = make_float2(acc[j][2], acc[j][3]);
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const uint32_t prev = atomicAdd(&o.splitk_cnt[blockIdx.y * gridDim.x + blockIdx.x], 1u);
            s_last = prev == splits - 1;
            if (prev == splits - 1) o.splitk_cnt[blockIdx.y * gridDim.x + blockIdx.x] = 0;   // ready for the next launch
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        if (kq == 0) {
#pragma unroll
            for (int j = 0; j < BN / 8; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
            for (uint32_t p = 0; p < splits; p++) {                   // split order, whoever arrived last
                const float *w = o.splitk_ws + ((size_t)p * rows_ws + m0 + warp * 16 + g) * ldws + n0 + t2;
#pragma unroll
                for (int j = 0; j < BN / 8; j++) {
                    const float2 u0 = __ldcg((const float2 *)(w + j * 8)), u1 = __ldcg((const float2 *)(w + (size_t)8 * ldws + j * 8));
                    acc[j][0] += u0.x; acc[j][1] += u0.y; acc[j][2] += u1.x; acc[j][3] += u1.y;
                }
            }
        }
    }
    if (kq != 0) return;
#pragma unroll
    for (int j = 0; j < BN / 8; j++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t row = m0 + warp * 16 + g + h * 8, col = n0 + j * 8 + t2;
            if (row >= M || col >= N) continue;
            float v0 = acc[j][2 * h], v1 = acc[j][2 * h + 1];
            const bool two = col + 1 < N;
            if (o.bias) { v0 += o.bias[col]; if (two) v1 += o.bias[col + 1]; }
            if (o.act == ACT_GELU_ERF) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); }
            else if (o.act == ACT_GELU_TANH) { v0 = gelu_tanh(v0); v1 = gelu_tanh(v1); }
            if (o.res) {
                const uint32_t rr = o.res_mod ? row % o.res_mod : row;
                const __half *rp = o.res + (size_t)rr * o.ldc + col;
                v0 += __half2float(rp[0]);
                if (two) v1 += __half2float(rp[1]);
            }
            if (o.debug_no_store) continue;
            if (o.c16) {
                __half *cp = o.c16 + (size_t)row * o.ldc + col;
                if (two && ((o.ldc | col) & 1) == 0) *(__half2 *)cp = __floats2half2_rn(v0, v1);
                else { cp[0] = __float2half_rn(v0); if (two) cp[1] = __float2half_rn(v1); }
            }
            if (o.c32) {
                float *cp = o.c32 + (size_t)row * o.ldc + col;
                cp[0] = v0;
                if (two) cp[1] = v1;
            }
        }
    }
}

// ------------------------------------------------------------------ panel-resident variant, M <= 64 (one text query = 64 tokens)
//
// ncu of k_gemm_skinny at 64 rows (profiles/r01p_skinny_gemm_ncu_full.md): two thirds of every pipeline stage is the 64-row activation
// panel, re-read from L2 by every CTA, so only 12-24 KB of WEIGHT bytes are in flight per SM -- far below what HBM latency needs
// (~45 KB per SM at the copy peak).  Here the panel (64 x Kc halfs, Kc <= 1152: 148 KB) is loaded into shared memory ONCE per CTA and
// the cp.async ring carries weights only: BN x 64 halfs per stage, 16-60 stages, 32-60 KB of HBM traffic in flight per SM.  A K larger
// than the panel (fc2: 4304) is processed in panel-sized chunks with the accumulators kept in registers.  N is cut into 8- or
// 32-column slices so that ~all SMs stream: 1152 -> 144 CTAs, 3456 -> 108, 4304 -> 135.
namespace pr {

static constexpr int kBM = 64, kBK = 64, kThreads = 256, kPitchB = kBK + 8;   // 8 warps: 4 row groups x 2 halves of every 64-wide k tile
static constexpr int kKcMax = 1152;                                          // panel depth (halfs)
static constexpr int kPitchA = kKcMax + 8;                                   // 2320 B rows: ldmatrix rows land in distinct banks
static constexpr uint32_t kPanelBytes = (uint32_t)kBM * kPitchA * 2;         // 148 480
static constexpr uint32_t kSmemBudget = 227u * 1024u - 1024u;

template <int BN>
constexpr int stages() {
    constexpr int per = BN * kPitchB * 2;
    constexpr int n = (int)((kSmemBudget - kPanelBytes) / per);
    return n > 48 ? 48 : n;
}
template <int BN>
constexpr uint32_t smem_bytes() { return kPanelBytes + (uint32_t)stages<BN>() * BN * kPitchB * 2; }

template <int BN>
__global__ void __launch_bounds__(kThreads) k_gemm_skinny_pr(const __half *__restrict__ A, const __half *__restrict__ B, uint32_t M, uint32_t N, uint32_t K,
                                                             uint32_t lda, uint32_t ldb, GemmOut o) {
    constexpr int kStagesB = stages<BN>();
    constexpr int NT = BN / 8;                                        // n8 tiles per CTA
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __half *sA = (__half *)smem_raw;                                  // [64][kPitchA]
    __half *sB = sA + (size_t)kBM * kPitchA;                          // [stage][BN][kPitchB]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wr = warp & 3, kh = warp >> 2;                          // row group, which two k16 steps of a k tile
    const uint32_t n0 = blockIdx.x * BN;

    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;

    for (uint32_t kc0 = 0; kc0 < K; kc0 += kKcMax) {
        const uint32_t kc = min((uint32_t)kKcMax, K - kc0);
        const uint32_t nk = (kc + kBK - 1) / kBK;
        if (kc0) __syncthreads();                                     // everyone is done with the previous panel chunk and ring
        else pdl_wait();   // TODO(perf): the first weight stages could be requested before this wait, as in k_gemm_skinny
        // panel chunk: 64 rows x kc halfs, 16-byte chunks; rows >= M and columns >= K are zero-filled
        {
            const uint32_t chunks_per_row = nk * 8;
            for (uint32_t c = tid; c < kBM * chunks_per_row; c += kThreads) {
                const uint32_t r = c / chunks_per_row, kk = (c % chunks_per_row) * 8;
                const bool ok = r < M && kc0 + kk < K;
                cp_async16(sA + r * kPitchA + kk, A + (size_t)(ok ? r : 0) * lda + (ok ? kc0 + kk : 0), ok);
            }
            cp_async_commit();
        }
        auto load_b = [&](uint32_t kt, int st) {
            const uint32_t k0 = kc0 + kt * kBK;
            __half *b = sB + (size_t)st * BN * kPitchB;
            for (int c = tid; c < BN * 8; c += kThreads) {
                const int r = c >> 3, kk = (c & 7) * 8;
                const bool ok = n0 + r < N && k0 + kk < K;
                cp_async16(b + r * kPitchB + kk, B + (size_t)(ok ? n0 + r : 0) * ldb + (ok ? k0 + kk : 0), ok);
            }
        };
#pragma unroll 1
        for (int s = 0; s < kStagesB - 1; s++) {
            if ((uint32_t)s < nk) load_b(s, s);
            cp_async_commit();
        }
        for (uint32_t kt = 0; kt < nk; kt++) {
            cp_async_wait<kStagesB - 2>();                            // the panel group is older than every weight stage: it has landed too
            __syncthreads();
            const uint32_t nxt = kt + kStagesB - 1;
            if (nxt < nk) load_b(nxt, nxt % kStagesB);
            cp_async_commit();
            const __half *b = sB + (size_t)(kt % kStagesB) * BN * kPitchB;
#pragma unroll
            for (int ks = 0; ks < 2; ks++) {
                const int kq = kh * 2 + ks;                           // k16 step inside the k tile
                uint32_t af[4];
                ldmatrix_x4(af, sA + (wr * 16 + (lane & 15)) * kPitchA + kt * kBK + kq * 16 + (lane >> 4) * 8);
                if constexpr (BN >= 16) {
#pragma unroll
                    for (int j = 0; j < BN / 16; j++) {               // two n8 tiles per ldmatrix.x4
                        uint32_t bf[4];
                        ldmatrix_x4(bf, b + (j * 16 + (lane & 7) + (lane >> 4) * 8) * kPitchB + kq * 16 + ((lane >> 3) & 1) * 8);
                        mma_16816(acc[2 * j], af, bf[0], bf[1]);
                        mma_16816(acc[2 * j + 1], af, bf[2], bf[3]);
                    }
                } else {                                              // BN = 8: one n8 tile; lanes 16-31 repeat the addresses of 0-15
                    uint32_t bf[4];
                    ldmatrix_x4(bf, b + (lane & 7) * kPitchB + kq * 16 + ((lane >> 3) & 1) * 8);
                    mma_16816(acc[0], af, bf[0], bf[1]);
                }
            }
        }
        cp_async_wait<0>();
    }
    __syncthreads();                                                  // the ring is free: reuse it for the two-way reduction over the k halves
    float *red = (float *)sB;                                         // [4 row groups][NT * 4][32 lanes]
    if (kh == 1) {
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) red[((wr * NT + j) * 4 + e) * 32 + lane] = acc[j][e];
    }
    __syncthreads();
    if (kh != 0) return;
#pragma unroll
    for (int j = 0; j < NT; j++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[j][e] += red[((wr * NT + j) * 4 + e) * 32 + lane];   // fixed order: half 0 + half 1
    // thread holds rows g and g + 8 (g = lane / 4), columns 2 (lane % 4) + {0, 1} of every n8 tile
    const uint32_t g = lane >> 2, t2 = (lane & 3) * 2;
#pragma unroll
    for (int j = 0; j < NT; j++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t row = wr * 16 + g + h * 8, col = n0 + j * 8 + t2;
            if (row >= M || col >= N) continue;
            float v0 = acc[j][2 * h], v1 = acc[j][2 * h + 1];
            const bool two = col + 1 < N;
            if (o.bias) { v0 += o.bias[col]; if (two) v1 += o.bias[col + 1]; }
            if (o.act == ACT_GELU_ERF) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); }
            else if (o.act == ACT_GELU_TANH) { v0 = gelu_tanh(v0); v1 = gelu_tanh(v1); }
            if (o.res) {
                const uint32_t rr = o.res_mod ? row % o.res_mod : row;
                const __half *rp = o.res + (size_t)rr * o.ldc + col;
                v0 += __half2float(rp[0]);
                if (two) v1 += __half2float(rp[1]);
            }
            if (o.debug_no_store) continue;
            if (o.c16) {
                __half *cp = o.c16 + (size_t)row * o.ldc + col;
                if (two && ((o.ldc | col) & 1) == 0) *(__half2 *)cp = __floats2half2_rn(v0, v1);
                else { cp[0] = __float2half_rn(v0); if (two) cp[1] = __float2half_rn(v1); }
            }
            if (o.c32) {
                float *cp = o.c32 + (size_t)row * o.ldc + col;
                cp[0] = v0;
                if (two) cp[1] = v1;
            }
        }
    }
}

}  // namespace pr

}  // namespace skinny
}  // namespace mse
