// One text query = 64 tokens (clip_server.py:98,137 at batch 1; BASELINE configs[0]): the 27 pre-LN blocks as ONE persistent kernel.
//
// Why: at 64 rows a block is four passes over 30.6 MB of weights (4.7 us at the HBM copy peak) and the forward pass was 189 kernels of
// ~9 us from launch to drain (profiles/r03c_text_latency_pdl.md) -- the kernels' fixed cost, not their work, is the latency.  Here one
// CTA per SM stays resident for the whole tower; the phases of a block are separated by a device-wide barrier (one atomic + an acquire
// spin), and every CTA requests the first WEIGHT stages of its next phase before it enters the barrier, so the weight stream never
// stops at a phase boundary.
//
// Phases of a block (aitemplate/model.py:26-55), M = 64 rows, every phase = one N slice (x one K split) per CTA:
//   1  qkv  = LN1(x) Wqkv^T + b        N 3456 in 24-column slices (144 CTAs), K 1152
//   2  att  = softmax(q k^T / sqrt(dh)) v   16 heads, one CTA each (the others already hold their out-proj weights)
//   3  x   += att Wo^T + b             N 1152 in 8-column slices (144 CTAs)
//   4  h    = gelu(LN2(x) W1^T + b)    N 4304 in 32-column slices (135 CTAs)
//   5  x   += h W2^T + b               N 1152 in 32-column slices x 4 splits of K 4304 (144 CTAs); the last CTA of a slice to arrive
//                                      adds the four partial tiles in split order (fixed order -> deterministic) and applies the epilogue
//
// LayerNorm is folded into the GEMM that consumes it:  LN(x) W^T = rstd * (x W'^T - mean * colsum(W')) + (b + W beta),  W' = W diag(gamma)
// rounded to fp16 at load time (k_fold_ln).  The GEMM runs on the raw residual stream; the row sums of x and x^2 are collected from the
// very A tiles the pipeline brings in (every thread re-reads the 16 bytes it requested), so LN costs no pass over memory at all.
// The products are exact in fp32 and the fp32 accumulation error is 2^-24 of the running sum, so the subtraction of mean * colsum is
// benign; the fp16 rounding moves from LN(x) to W', the same size of error.
//
// Math is warp-level mma.sync m16n8k16 (fp16 x fp16 -> fp32), as in gemm_skinny.cuh: at 2 FLOP per weight byte the tensor pipe idles
// either way.  16 warps = 4 row groups x the 4 k16 steps of every 64-wide k tile; the four partial sums are added in a fixed order.
#pragma once
#include "attention.cuh"
#include "common.cuh"
#include "gemm_epilogues.cuh"
#include "gemm_skinny.cuh"
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace mse {
namespace tmega {

using skinny::cp_async16;
using skinny::cp_async_commit;
using skinny::cp_async_wait;
using skinny::ldmatrix_x4;
using skinny::mma_16816;

#ifndef MSE_TMEGA_STAGES
#define MSE_TMEGA_STAGES 4
#endif
#ifndef MSE_TMEGA_ROT
#define MSE_TMEGA_ROT 1
#endif
// slice widths (columns per CTA) of the four GEMM phases; see the measurements in profiles/r03d_text_mega.md
#ifndef MSE_TMEGA_BN_QKV
#define MSE_TMEGA_BN_QKV 24
#define MSE_TMEGA_BN_PROJ 8
#define MSE_TMEGA_BN_FC1 32
#define MSE_TMEGA_BN_FC2 32
#endif
#ifndef MSE_TMEGA_THREADS
#define MSE_TMEGA_THREADS 512
#endif
#ifndef MSE_TMEGA_BK
#define MSE_TMEGA_BK 128
#endif
static constexpr int kThreads = MSE_TMEGA_THREADS, kKQ = kThreads / 128;   // k groups: the k16 steps of a k tile are dealt out to kKQ warps per row group
static_assert(kThreads == 256 || kThreads == 512, "text_mega: 256 or 512 threads");
static constexpr int kBM = 64, kBK = MSE_TMEGA_BK, kStages = MSE_TMEGA_STAGES, kPitch = kBK + 8;
static constexpr int kTPR = kBK / 8;                       // 16-byte chunks (= threads) per tile row
static constexpr int kRowsPerPass = kThreads / kTPR, kAChunks = kBM / kRowsPerPass;   // A chunks per thread and stage
static constexpr int kK16 = kBK / 16 / kKQ;                // k16 steps per warp and k tile
static_assert(kBK == 64 || kBK == 128, "text_mega: k tile");
static constexpr int kBnQkv = MSE_TMEGA_BN_QKV, kBnProj = MSE_TMEGA_BN_PROJ, kBnFc1 = MSE_TMEGA_BN_FC1, kBnFc2 = MSE_TMEGA_BN_FC2;
static constexpr int kMaxBN = (kBnQkv > kBnFc1 ? kBnQkv : kBnFc1) > (kBnProj > kBnFc2 ? kBnProj : kBnFc2) ? (kBnQkv > kBnFc1 ? kBnQkv : kBnFc1)
                                                                                                      : (kBnProj > kBnFc2 ? kBnProj : kBnFc2);
static constexpr uint32_t kOffA = 0;
static constexpr uint32_t kOffB = kOffA + (uint32_t)kStages * kBM * kPitch * 2;          // 9 216 per stage
static constexpr uint32_t kOffRed = kOffB + (uint32_t)kStages * kMaxBN * kPitch * 2;     // + 4 608 per stage
static constexpr uint32_t kOffStats = kOffRed + (uint32_t)(kKQ - 1) * 4 * (kMaxBN / 2) * 32 * 4;          // + 768 per column
// the attention tile of phase 2 lives in the A ring or in the reduction buffer when one of them is large enough (both are idle then;
// the B ring holds prefetched weights)
static constexpr uint32_t kRedBytes = kOffStats - kOffRed;
static constexpr bool kAttnInRing = (uint32_t)sizeof(attn::Smem) <= kOffB, kAttnInRed = !kAttnInRing && (uint32_t)sizeof(attn::Smem) <= kRedBytes;
static constexpr uint32_t kOffAttn = kAttnInRing ? kOffA : (kAttnInRed ? kOffRed : kOffStats + 64 * 2 * 4);
static constexpr uint32_t kSmemBytes = kOffStats + 64 * 2 * 4 + ((kAttnInRing || kAttnInRed) ? 0u : (uint32_t)sizeof(attn::Smem));
static_assert(kSmemBytes <= 227 * 1024, "text_mega: shared memory");
static constexpr int kFc2Splits = 4;

struct LayerP {
    const __half *qkv_w, *proj_w, *fc1_w, *fc2_w;       // qkv_w / fc1_w: gamma folded in (k_fold_ln)
    const float *qkv_b, *qkv_cs, *proj_b, *fc1_b, *fc1_cs, *fc2_b;   // qkv_b / fc1_b: b + W beta; *_cs: column sums of the folded weights
};

struct Params {
    const LayerP *layers;   // device array [depth]
    int depth;
    uint32_t D, F, H;       // model width, MLP width, heads
    __half *x, *qkv, *att, *h;
    float *ws;              // [kFc2Splits][64][D] fp32 partial tiles of phase 5
    uint32_t *cnt;          // [D / 32] arrival counters (left at zero)
    uint32_t *bar;          // device-wide barrier counter (zeroed before the launch)
    float scale_log2e, eps;
    int act;
    int debug;              // profiling only (MSE_TMEGA_DEBUG): 1 skip the GEMM main loops, 2 skip attention, 4 skip the device-wide barriers,
                            // 8 no A loads, 16 no ldmatrix / mma, 32 no B loads in the main loop
};

// ---- device-wide barrier: every CTA of the (cooperative, co-resident) grid calls it the same number of times
#ifndef MSE_TMEGA_BARRIER
#define MSE_TMEGA_BARRIER 0
#endif
__device__ __forceinline__ void grid_sync(uint32_t *bar, uint32_t &target, int debug = 0) {
    if (debug & 4) { __syncthreads(); return; }
#if MSE_TMEGA_BARRIER == 2
    cooperative_groups::this_grid().sync();
    return;
#endif
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        // release: acq_rel fence (cumulative over the CTA's writes ordered before it by the bar.sync above) + relaxed add, as CUTLASS's
        // GenericBarrier does; __threadfence() is the heavier fence.sc (MEMBAR.SC.GPU)
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
        uint32_t v;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if ((int32_t)(v - target) >= 0) break;
#if MSE_TMEGA_BARRIER == 1
            __nanosleep(40);
#endif
        }
    }
    __syncthreads();
}

// one phase's work item of this CTA: rows [0, 64) x columns [n0, n0 + BN) over k tiles [kt0, kt1)
struct Item {
    const __half *W;
    uint32_t N, K, n0, kt0, kt1, rot;
    bool valid;
};

// Every CTA walks the k tiles of its range from a different starting point: the 144 CTAs of a phase all read the SAME 64-row activation
// panel, and in lockstep they would all pull on the same 64 lines (a handful of L2 slices) at the same time.
__device__ __forceinline__ uint32_t rot_tile(const Item &it, uint32_t kt) {
#if MSE_TMEGA_ROT
    const uint32_t nk = it.kt1 - it.kt0, r = kt + it.rot;
    return r >= nk ? r - nk : r;
#else
    return kt;
#endif
}

template <int BN>
__device__ __forceinline__ void load_b(uint8_t *smem, const Item &it, uint32_t kt, int st) {
    const int tid = threadIdx.x;
    const uint32_t k0 = (it.kt0 + rot_tile(it, kt)) * kBK;
    __half *b = (__half *)(smem + kOffB) + (size_t)st * BN * kPitch;
#pragma unroll
    for (int c = tid; c < BN * kTPR; c += kThreads) {
        const int r = c / kTPR, kc = (c % kTPR) * 8;
        const bool ok = it.n0 + r < it.N && k0 + kc < it.K;
        cp_async16(b + r * kPitch + kc, it.W + (size_t)(ok ? it.n0 + r : 0) * it.K + (ok ? k0 + kc : 0), ok);
    }
}

// the weight tiles of the item's first stages: issued BEFORE the barrier that ends the previous phase (they join the first commit group
// of gemm_main).  Also starting the REST of the slice towards L2 here (prefetch.global.L2) measured no gain (1.34 vs 1.37 ms).
template <int BN>
__device__ __forceinline__ void prefetch_b(uint8_t *smem, const Item &it) {
    if (!it.valid) return;
    const uint32_t nk = it.kt1 - it.kt0;
#pragma unroll 1
    for (int s = 0; s < kStages - 1; s++)
        if ((uint32_t)s < nk) load_b<BN>(smem, it, s, s);
}

// the item's slice of a bias / column-sum vector, started towards L2 together with its weights (the epilogue would otherwise open with
// a round trip to HBM: every block has its own vectors)
__device__ __forceinline__ void prefetch_vec(const Item &it, const float *v, int bn) {
    if (it.valid && threadIdx.x < 2) {
        const char *p = (const char *)(v + it.n0) + threadIdx.x * 128;
        if (threadIdx.x * 128 < bn * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
}

// acc (valid in the warps with kq == 0 on return): rows 16 wr + g (+ 8), columns n0 + 8 j + 2 (lane % 4) + {0, 1}.
// STATS: row sums of a and a^2 over the K range -> smem stats[row][2] (mean, rstd) -- needs the full K range in this item.
template <int BN, bool STATS>
__device__ __forceinline__ void gemm_main(uint8_t *smem, const __half *A, uint32_t lda, const Item &it, bool b_prefetched, float (&acc)[BN / 8][4],
                                          float eps, int debug) {
    constexpr int NT = BN / 8;
    const int tid = threadIdx.x, warp = (tid >> 5) & 3, kq = tid >> 7, lane = tid & 31;
    __half *sA = (__half *)(smem + kOffA);
    __half *sB = (__half *)(smem + kOffB);
    const uint32_t nk = (debug & 1) ? 0 : it.kt1 - it.kt0;
    const int ar = tid / kTPR, akc = (tid % kTPR) * 8;                 // this thread's 16-byte chunks of every A stage: rows ar + c * kRowsPerPass
    auto load_a = [&](uint32_t kt, int st) {
        const uint32_t k0 = (it.kt0 + rot_tile(it, kt)) * kBK;
        const bool ok = k0 + akc < it.K;
        if (debug & 8) return;
#pragma unroll
        for (int c = 0; c < kAChunks; c++)
            cp_async16(sA + (size_t)st * kBM * kPitch + (ar + c * kRowsPerPass) * kPitch + akc, A + (size_t)(ar + c * kRowsPerPass) * lda + (ok ? k0 + akc : 0), ok);
    };
#pragma unroll
    for (int j = 0; j < NT; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    float s1[kAChunks], s2[kAChunks];
#pragma unroll
    for (int c = 0; c < kAChunks; c++) s1[c] = s2[c] = 0.f;
#pragma unroll
    for (int s = 0; s < kStages - 1; s++) {
        if ((uint32_t)s < nk) {
            load_a(s, s);
            if (!b_prefetched) load_b<BN>(smem, it, s, s);
        }
        cp_async_commit();
    }
    for (uint32_t kt = 0; kt < nk; kt++) {
        cp_async_wait<kStages - 2>();
        __syncthreads();                                              // stage kt has landed; stage kt-1 is free for the next load
        const uint32_t nxt = kt + kStages - 1;
        if (nxt < nk) { load_a(nxt, nxt % kStages); if (!(debug & 32)) load_b<BN>(smem, it, nxt, nxt % kStages); }
        cp_async_commit();
        const __half *a = sA + (size_t)(kt % kStages) * kBM * kPitch;
        const __half *b = sB + (size_t)(kt % kStages) * BN * kPitch;
        if (STATS) {
#pragma unroll
            for (int c = 0; c < kAChunks; c++) {
                const uint4 u = *(const uint4 *)(a + (ar + c * kRowsPerPass) * kPitch + akc);
                const __half2 *h = (const __half2 *)&u;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float2 f = __half22float2(h[e]);
                    s1[c] += f.x + f.y;
                    s2[c] = fmaf(f.x, f.x, s2[c]);
                    s2[c] = fmaf(f.y, f.y, s2[c]);
                }
            }
        }
        if (!(debug & 16))
#pragma unroll
        for (int ks = 0; ks < kK16; ks++) {
            const int k16 = kq * kK16 + ks;                            // k16 step inside the k tile
            uint32_t af[4];
            ldmatrix_x4(af, a + (warp * 16 + (lane & 15)) * kPitch + k16 * 16 + (lane >> 4) * 8);
#pragma unroll
            for (int j = 0; j < NT / 2; j++) {                         // two n8 tiles per ldmatrix.x4
                uint32_t bf[4];
                ldmatrix_x4(bf, b + (j * 16 + (lane & 7) + (lane >> 4) * 8) * kPitch + k16 * 16 + ((lane >> 3) & 1) * 8);
                mma_16816(acc[2 * j], af, bf[0], bf[1]);
                mma_16816(acc[2 * j + 1], af, bf[2], bf[3]);
            }
            if (NT & 1) {                                              // the odd last tile: lanes 16-31 repeat the addresses of 0-15
                uint32_t bf[4];
                ldmatrix_x4(bf, b + ((NT - 1) * 8 + (lane & 7)) * kPitch + k16 * 16 + ((lane >> 3) & 1) * 8);
                mma_16816(acc[NT - 1], af, bf[0], bf[1]);
            }
        }
    }
    cp_async_wait<0>();
    float *red = (float *)(smem + kOffRed);                           // [kKQ - 1][4 warps][NT * 4][32 lanes], its own region: the ring is free
    if (kq > 0) {
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) red[(((kq - 1) * 4 + warp) * (NT * 4) + j * 4 + e) * 32 + lane] = acc[j][e];
    }
    if (STATS) {
        // the kTPR threads of a row are consecutive lanes; fixed shuffle order
#pragma unroll
        for (int c = 0; c < kAChunks; c++) {
#pragma unroll
            for (int o = kTPR / 2; o; o >>= 1) {
                s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], o);
                s2[c] += __shfl_xor_sync(0xffffffffu, s2[c], o);
            }
            if (tid % kTPR == 0) {
                const float mean = s1[c] / (float)it.K;
                const float var = fmaxf(s2[c] / (float)it.K - mean * mean, 0.f);
                float *stats = (float *)(smem + kOffStats);
                stats[2 * (ar + c * kRowsPerPass)] = mean;
                stats[2 * (ar + c * kRowsPerPass) + 1] = rsqrtf(var + eps);
            }
        }
    }
    __syncthreads();                                                  // ring free, partial sums and row statistics visible
    if (kq == 0) {
#pragma unroll
        for (int q = 0; q < kKQ - 1; q++)                             // fixed order of the k groups
#pragma unroll
            for (int j = 0; j < NT; j++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[j][e] += red[((q * 4 + warp) * (NT * 4) + j * 4 + e) * 32 + lane];
    }
}

__device__ __forceinline__ float act_fn(float v, int act) {
    return act == ACT_GELU_ERF ? gelu_erf(v) : (act == ACT_GELU_TANH ? gelu_tanh(v) : v);
}

// rows g, g + 8 of row group `warp`; columns col, col + 1 (col even, N even)
template <int BN, class F>
__device__ __forceinline__ void for_each_out(const Item &it, const float (&acc)[BN / 8][4], F f) {
    const int tid = threadIdx.x, warp = (tid >> 5) & 3, lane = tid & 31;
    const uint32_t g = lane >> 2, t2 = (lane & 3) * 2;
#pragma unroll
    for (int j = 0; j < BN / 8; j++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t row = warp * 16 + g + h * 8, col = it.n0 + j * 8 + t2;
            if (col < it.N) f(row, col, acc[j][2 * h], acc[j][2 * h + 1]);
        }
}

__device__ __forceinline__ Item make_item(const __half *W, uint32_t N, uint32_t K, uint32_t BN, uint32_t splits, uint32_t idx) {
    Item it;
    const uint32_t slices = (N + BN - 1) / BN, nk = (K + kBK - 1) / kBK;
    it.W = W; it.N = N; it.K = K;
    it.valid = idx < slices * splits;
    const uint32_t sl = idx % slices, sp = idx / slices;
    it.n0 = sl * BN;
    it.kt0 = it.valid ? nk * sp / splits : 0;
    it.kt1 = it.valid ? nk * (sp + 1) / splits : 0;
    it.rot = it.valid ? (idx * 7u) % (it.kt1 - it.kt0) : 0;
    return it;
}

__global__ void __launch_bounds__(kThreads, 1) k_text_blocks(Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_last;
    const int tid = threadIdx.x, warp = (tid >> 5) & 3, kq = tid >> 7, lane = tid & 31;
    const uint32_t D = p.D, F = p.F, cta = blockIdx.x;
    const float *stats = (const float *)(smem + kOffStats);
    uint32_t bar_target = 0;
    constexpr int BN_QKV = kBnQkv, BN_PROJ = kBnProj, BN_FC1 = kBnFc1, BN_FC2 = kBnFc2;
    // every phase must fit the grid in one pass (the host checks): one item per CTA.  The items are the same in every block but for
    // the weight pointer, so their index arithmetic (integer divisions) is done once.
    const Item item_qkv = make_item(nullptr, 3 * D, D, BN_QKV, 1, cta), item_proj = make_item(nullptr, D, D, BN_PROJ, 1, cta),
               item_fc1 = make_item(nullptr, F, D, BN_FC1, 1, cta), item_fc2 = make_item(nullptr, D, F, BN_FC2, kFc2Splits, cta);
    auto with_w = [](Item it, const __half *w) { it.W = w; return it; };
    const uint32_t fc2_split = cta / ((D + BN_FC2 - 1) / BN_FC2);
    Item nxt = with_w(item_qkv, p.layers[0].qkv_w);
    prefetch_b<BN_QKV>(smem, nxt);
    prefetch_vec(nxt, p.layers[0].qkv_cs, BN_QKV);
    prefetch_vec(nxt, p.layers[0].qkv_b, BN_QKV);
    for (int l = 0; l < p.depth; l++) {
        const LayerP &L = p.layers[l];
        // ---------------------------------------------------------------- 1: qkv = LN1(x) Wqkv^T + b
        {
            const Item it = nxt;
            float acc[BN_QKV / 8][4];
            if (it.valid) gemm_main<BN_QKV, true>(smem, p.x, D, it, true, acc, p.eps, p.debug);
            nxt = with_w(item_proj, L.proj_w);
            prefetch_b<BN_PROJ>(smem, nxt);
            prefetch_vec(nxt, L.proj_b, BN_PROJ);
            if (it.valid && kq == 0)
                for_each_out<BN_QKV>(it, acc, [&](uint32_t row, uint32_t col, float v0, float v1) {
                    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
                    const float2 cs = *(const float2 *)(L.qkv_cs + col), b = *(const float2 *)(L.qkv_b + col);
                    v0 = fmaf(rstd, fmaf(-mean, cs.x, v0), b.x);
                    v1 = fmaf(rstd, fmaf(-mean, cs.y, v1), b.y);
                    *(__half2 *)(p.qkv + (size_t)row * 3 * D + col) = __floats2half2_rn(v0, v1);
                });
        }
        grid_sync(p.bar, bar_target, p.debug);
        // ---------------------------------------------------------------- 2: attention, one CTA per head
        if (cta < p.H && !(p.debug & 2)) {
            attn::Smem &sm = *reinterpret_cast<attn::Smem *>(smem + kOffAttn);
            attn::mha_tile<kThreads>(p.qkv, p.att, sm, 64, (int)p.H, 0, (int)cta, 0, p.scale_log2e);
        }
        grid_sync(p.bar, bar_target, p.debug);
        // ---------------------------------------------------------------- 3: x += att Wo^T + b
        {
            const Item it = nxt;
            float acc[BN_PROJ / 8][4];
            if (it.valid) gemm_main<BN_PROJ, false>(smem, p.att, D, it, true, acc, p.eps, p.debug);
            nxt = with_w(item_fc1, L.fc1_w);
            prefetch_b<BN_FC1>(smem, nxt);
            prefetch_vec(nxt, L.fc1_cs, BN_FC1);
            prefetch_vec(nxt, L.fc1_b, BN_FC1);
            if (it.valid && kq == 0)
                for_each_out<BN_PROJ>(it, acc, [&](uint32_t row, uint32_t col, float v0, float v1) {
                    const float2 b = *(const float2 *)(L.proj_b + col);
                    __half2 *xp = (__half2 *)(p.x + (size_t)row * D + col);
                    const float2 r = __half22float2(__ldcg(xp));
                    *xp = __floats2half2_rn(v0 + b.x + r.x, v1 + b.y + r.y);
                });
        }
        grid_sync(p.bar, bar_target, p.debug);
        // ---------------------------------------------------------------- 4: h = act(LN2(x) W1^T + b)
        {
            const Item it = nxt;
            float acc[BN_FC1 / 8][4];
            if (it.valid) gemm_main<BN_FC1, true>(smem, p.x, D, it, true, acc, p.eps, p.debug);
            nxt = with_w(item_fc2, L.fc2_w);
            prefetch_b<BN_FC2>(smem, nxt);
            prefetch_vec(nxt, L.fc2_b, BN_FC2);
            if (it.valid && kq == 0)
                for_each_out<BN_FC1>(it, acc, [&](uint32_t row, uint32_t col, float v0, float v1) {
                    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
                    const float2 cs = *(const float2 *)(L.fc1_cs + col), b = *(const float2 *)(L.fc1_b + col);
                    v0 = act_fn(fmaf(rstd, fmaf(-mean, cs.x, v0), b.x), p.act);
                    v1 = act_fn(fmaf(rstd, fmaf(-mean, cs.y, v1), b.y), p.act);
                    *(__half2 *)(p.h + (size_t)row * F + col) = __floats2half2_rn(v0, v1);
                });
        }
        grid_sync(p.bar, bar_target, p.debug);
        // ---------------------------------------------------------------- 5: x += h W2^T + b, K split four ways
        {
            const Item it = nxt;
            float acc[BN_FC2 / 8][4];
            if (it.valid) gemm_main<BN_FC2, false>(smem, p.h, F, it, true, acc, p.eps, p.debug);
            if (l + 1 < p.depth) {
                nxt = with_w(item_qkv, p.layers[l + 1].qkv_w);
                prefetch_b<BN_QKV>(smem, nxt);
                prefetch_vec(nxt, p.layers[l + 1].qkv_cs, BN_QKV);
                prefetch_vec(nxt, p.layers[l + 1].qkv_b, BN_QKV);
            }
            if (it.valid) {
                const uint32_t sl = it.n0 / BN_FC2, sp = fc2_split;
                if (kq == 0)
                    for_each_out<BN_FC2>(it, acc, [&](uint32_t row, uint32_t col, float v0, float v1) {
                        __stcg((float2 *)(p.ws + ((size_t)sp * kBM + row) * D + col), make_float2(v0, v1));
                    });
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                __syncthreads();
                if (tid == 0) {
                    const uint32_t prev = atomicAdd(&p.cnt[sl], 1u);
                    s_last = prev == kFc2Splits - 1;
                    if (prev == kFc2Splits - 1) p.cnt[sl] = 0;       // ready for the next block
                }
                __syncthreads();
                if (s_last) {
                    asm volatile("fence.acq_rel.gpu;" ::: "memory");
                    if (kq == 0)
                        for_each_out<BN_FC2>(it, acc, [&](uint32_t row, uint32_t col, float, float) {
                            float v0 = 0.f, v1 = 0.f;
#pragma unroll
                            for (int q = 0; q < kFc2Splits; q++) {     // split order, whoever arrived last
                                const float2 u = __ldcg((const float2 *)(p.ws + ((size_t)q * kBM + row) * D + col));
                                v0 += u.x;
                                v1 += u.y;
                            }
                            const float2 b = *(const float2 *)(L.fc2_b + col);
                            __half2 *xp = (__half2 *)(p.x + (size_t)row * D + col);
                            const float2 r = __half22float2(__ldcg(xp));
                            *xp = __floats2half2_rn(v0 + b.x + r.x, v1 + b.y + r.y);
                        });
                }
            }
        }
        grid_sync(p.bar, bar_target, p.debug);
    }
}

// LN folded into the weights of the GEMM behind it: wf[n][k] = fp16(w[n][k] * gamma[k]), cs[n] = sum_k wf[n][k], bf[n] = b[n] + sum_k w[n][k] * beta[k].
// One warp per output row.
__global__ void __launch_bounds__(256) k_fold_ln(const __half *__restrict__ w, const float *__restrict__ b, const float *__restrict__ gamma,
                                                 const float *__restrict__ beta, uint32_t N, uint32_t K, __half *__restrict__ wf, float *__restrict__ cs,
                                                 float *__restrict__ bf) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    float s = 0.f, t = 0.f;
    for (uint32_t k = lane; k < K; k += 32) {
        const float wv = __half2float(w[(size_t)n * K + k]);
        const __half f = __float2half_rn(wv * gamma[k]);
        wf[(size_t)n * K + k] = f;
        s += __half2float(f);
        t = fmaf(wv, beta[k], t);
    }
    for (int o = 16; o; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    if (lane == 0) {
        cs[n] = s;
        bf[n] = b[n] + t;
    }
}

}  // namespace tmega
}  // namespace mse
