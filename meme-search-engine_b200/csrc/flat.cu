// Flat (brute-force) inner-product search over an HBM-resident fp16 index.
//
// Replaces faiss IndexScalarQuantizer(QT_fp16, METRIC_INNER_PRODUCT) add/search as the reference drives it
// (src/main.rs:822,858,892,900; mse.py:72-85) and VectorList (diskann/src/vector.rs:118-186).
//
// Data flow of one search (all on one stream, no host round trip until the final status word):
//   rows are visited in chunks that double in size (4096, 4096, 8192, ...).  A scoring kernel (exact fp64
//   scan here, or the tcgen05 GEMM in flat_tc.cu) appends every (score,id) with score >= thr[q] to a
//   per-query candidate buffer; k_select folds the buffer into the query's running top-kp list and raises
//   thr[q] to the kp-th best.  Because thr[q] never exceeds the final kp-th best, no member of the final
//   top-kp is ever dropped; under exchangeable row order a chunk contributes ~kp candidates per query.
//   The tensor path then re-scores its kp survivors in fp64 and certifies the cut (see k_finalize).
#include "internal.h"
#include <math.h>
#include <stdlib.h>
#include <algorithm>

namespace mse {

static constexpr uint32_t kSortMax = 8192;    // elements k_select can sort in shared memory
static constexpr uint32_t kKpMax = 2048;      // largest running list
static constexpr uint64_t kChunk0 = 4096;     // first chunk (thr = -inf there, so it must fit the buffer)
static constexpr uint64_t kGrowth = 4;        // a chunk is kGrowth x the rows before it (see ChunkPlan)

// ------------------------------------------------------------------ index maintenance kernels

__global__ void k_f32_to_f16(const float *__restrict__ src, __half *__restrict__ dst, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = __float2half_rn(src[i]);
}

// max_i |x_i|_2 over fp16 rows, accumulated into *out (non-negative floats order like their bit patterns)
__global__ void k_row_norm_max(const __half *__restrict__ x, uint64_t n, uint32_t d, float *out) {
    uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    int lane = threadIdx.x & 31;
    float best = 0.f;
    for (uint64_t r = warp; r < n; r += nwarps) {
        const __half *row = x + r * d;
        float s = 0.f;
        for (uint32_t j = lane; j < d; j += 32) {
            float v = __half2float(row[j]);
            s = fmaf(v, v, s);
        }
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        best = fmaxf(best, s);
    }
    if (lane == 0 && best > 0.f) atomicMax((int *)out, __float_as_int(sqrtf(best) * 1.0001f));
}

// ------------------------------------------------------------------ query preparation

// one warp per query: |q|, fp16 copy for the tensor path, |q - f16(q)|, and the certificate slack
//   eps = (|q - q16| + c_acc * |q16|) * max_row_norm
// where c_acc bounds the fp32 accumulation error of the tensor-core pass (K <= 4096 sequential roundings of
// relative size 2^-23 on partial sums bounded by |q16||x|); the products q16*x are exact in fp32.
__global__ void k_prepare_queries(const float *__restrict__ q, uint32_t nq, uint32_t nq_pad, uint32_t d,
                                  __half *__restrict__ q16, float *__restrict__ qstat, const float *max_norm) {
    uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= nq_pad) return;
    if (w >= nq) {
        if (q16) for (uint32_t j = lane; j < d; j += 32) q16[(size_t)w * d + j] = __float2half_rn(0.f);
        return;
    }
    double n2 = 0.0, e2 = 0.0, h2 = 0.0;
    for (uint32_t j = lane; j < d; j += 32) {
        float v = q[(size_t)w * d + j];
        __half h = __float2half_rn(v);
        float hv = __half2float(h);
        if (q16) q16[(size_t)w * d + j] = h;
        n2 += (double)v * v;
        h2 += (double)hv * hv;
        double e = (double)v - (double)hv;
        e2 += e * e;
    }
    for (int o = 16; o; o >>= 1) {
        n2 += __shfl_xor_sync(0xffffffffu, n2, o);
        e2 += __shfl_xor_sync(0xffffffffu, e2, o);
        h2 += __shfl_xor_sync(0xffffffffu, h2, o);
    }
    if (lane == 0) {
        const double c_acc = 2.0e-4;
        double eps = (sqrt(e2) + c_acc * sqrt(h2)) * (double)(*max_norm);
        qstat[w * 4 + 0] = (float)sqrt(n2);
        qstat[w * 4 + 1] = (float)sqrt(e2);
        qstat[w * 4 + 2] = (float)(eps * 1.0001 + 1e-30);
        qstat[w * 4 + 3] = 0.f;
    }
}

__global__ void k_reset_state(uint32_t *ntop, float *thr, uint32_t *count, uint32_t *flags, const uint32_t *qsel,
                              uint32_t nsel, int clear_global) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nsel) {
        uint32_t q = qsel ? qsel[i] : i;
        ntop[q] = 0;
        thr[q] = -INFINITY;
        count[q] = 0;
        flags[4 + q] = 0;
    }
    if (clear_global && i < 4) flags[i] = 0;
}

// get_total_embedding (src/common.rs:215-274): q += weight * e, e an fp16 embedding row, q the f32 query (never renormalised)
__global__ void k_query_axpy_f16(const __half *__restrict__ e, float w, float *__restrict__ q, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) q[i] += __half2float(e[i]) * w;                 // common.rs:270 `*total += *value * weight` (f32 multiply, f32 add)
}

__global__ void k_iota(uint32_t *p, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// ------------------------------------------------------------------ exact fp64 scan
//
// One warp per row; lane l owns the 16-byte pieces l, l+32, ... of the row (coalesced 512-byte warp loads).
// The query lives in registers as doubles, so the product f32(q)*f16(x) and its accumulation are exact to
// fp64 rounding: the f32-rounded result equals the oracle's for all but ~1e-9 of scores.
// HBM-bound: 2 bytes per (row, dim) read once per QT queries.
template <int VPL, int QT>
__global__ void __launch_bounds__(256) k_flat_scan(const __half *__restrict__ x, uint64_t row0, uint64_t nrows, uint32_t d,
                                                   const float *__restrict__ q, const uint32_t *__restrict__ qsel,
                                                   uint32_t sel0, uint32_t nsel, const float *__restrict__ thr,
                                                   uint64_t *__restrict__ cand, uint32_t *__restrict__ count, uint32_t cap) {
    const int lane = threadIdx.x & 31;
    const uint32_t nv = d >> 3;  // 16-byte pieces per row
    uint32_t qi[QT];
    float th[QT];
    double qr[QT][VPL][8];
#pragma unroll
    for (int t = 0; t < QT; t++) {
        bool live = sel0 + t < nsel;
        qi[t] = live ? qsel[sel0 + t] : 0xffffffffu;
        th[t] = live ? thr[qi[t]] : INFINITY;
#pragma unroll
        for (int v = 0; v < VPL; v++) {
            uint32_t p = lane + 32 * v;
#pragma unroll
            for (int e = 0; e < 8; e++) qr[t][v][e] = (live && p < nv) ? (double)q[(size_t)qi[t] * d + p * 8 + e] : 0.0;
        }
    }
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < nrows; r += nwarps) {
        const uint4 *row = (const uint4 *)(x + (row0 + r) * d);
        uint4 xv[VPL];
#pragma unroll
        for (int v = 0; v < VPL; v++) {
            uint32_t p = lane + 32 * v;
            xv[v] = (p < nv) ? __ldg(row + p) : make_uint4(0, 0, 0, 0);
        }
        double acc[QT];
#pragma unroll
        for (int t = 0; t < QT; t++) acc[t] = 0.0;
#pragma unroll
        for (int v = 0; v < VPL; v++) {
            const __half2 *h = (const __half2 *)&xv[v];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                float2 f = __half22float2(h[e]);
                double x0 = (double)f.x, x1 = (double)f.y;
#pragma unroll
                for (int t = 0; t < QT; t++) {
                    acc[t] = fma(qr[t][v][2 * e], x0, acc[t]);
                    acc[t] = fma(qr[t][v][2 * e + 1], x1, acc[t]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < QT; t++) {
            double a = acc[t];
            for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            float s = (float)a;
            if (lane == 0 && s >= th[t]) {
                uint32_t slot = atomicAdd(&count[qi[t]], 1u);
                if (slot < cap) cand[(size_t)qi[t] * cap + slot] = rank_key(s, (uint32_t)(row0 + r));
            }
        }
    }
}

// ------------------------------------------------------------------ select: fold candidates into the running list

__device__ __forceinline__ void bitonic_sort_desc(uint64_t *s, uint32_t np2) {
    for (uint32_t k2 = 2; k2 <= np2; k2 <<= 1) {
        for (uint32_t j = k2 >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x) {
                uint32_t ixj = i ^ j;
                if (ixj > i) {
                    uint64_t a = s[i], b = s[ixj];
                    bool up = (i & k2) == 0;
                    if (up ? (a < b) : (a > b)) { s[i] = b; s[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ uint32_t next_pow2_min64(uint32_t n) {
    uint32_t p = 64;
    while (p < n) p <<= 1;
    return p;
}

// one CTA per (selected) query.  Only the SET of the kp best keys and the kp-th key (the new threshold) are needed here -- the list is
// sorted once, in k_finalize -- so this is a selection, not a sort: an 8-pass radix select over the 64-bit rank keys (byte histograms
// in shared memory, keys are distinct because they carry the row id), then a compaction.  The bitonic sort this replaces took ~300 us
// per chunk at 1024 queries (4096 keys each) and was a third of a search step at the 8-GPU shard size.
__global__ void __launch_bounds__(512) k_select(uint64_t *__restrict__ top, uint32_t top_stride, uint32_t kp,
                                                uint32_t *__restrict__ ntop, float *__restrict__ thr,
                                                const uint64_t *__restrict__ cand, uint32_t cap, uint32_t *__restrict__ count,
                                                uint32_t *__restrict__ flags, const uint32_t *__restrict__ qsel) {
    extern __shared__ uint64_t s_keys[];
    __shared__ uint32_t s_hist[256];
    __shared__ uint64_t s_prefix;
    __shared__ uint32_t s_want, s_out;
    const uint32_t q = qsel ? qsel[blockIdx.x] : blockIdx.x;
    uint32_t nt = ntop[q], c = count[q];
    const bool ovf = c > cap;
    if (ovf) c = cap;
    if (c == 0) return;  // nothing new: list and threshold stand
    const uint32_t n = nt + c;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_keys[i] = i < nt ? top[(size_t)q * top_stride + i] : cand[(size_t)q * cap + (i - nt)];
    if (threadIdx.x == 0) { s_prefix = 0ull; s_want = kp; s_out = 0; }
    __syncthreads();
    uint64_t thr_key = 0ull;        // keep everything
    uint32_t keep = n;
    if (n >= kp) {
        keep = kp;
        for (int shift = 56; shift >= 0; shift -= 8) {
            for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
            __syncthreads();
            const uint64_t prefix = s_prefix;
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                const uint64_t key = s_keys[i];
                if (shift == 56 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&s_hist[(uint32_t)(key >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                // lane l owns bins 255 - 8l .. 248 - 8l (largest digits first); find the bin holding the s_want-th largest key
                const uint32_t lane = threadIdx.x, want = s_want;
                uint32_t h[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) { h[j] = s_hist[255 - 8 * lane - j]; sum += h[j]; }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                    if ((int)lane >= o) incl += v;
                }
                const unsigned m = __ballot_sync(0xffffffffu, incl >= want);
                const int owner = __ffs(m) - 1;                        // m != 0: the candidates still in play number >= want
                if ((int)lane == owner) {
                    uint32_t run = incl - sum;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        if (run + h[j] >= want) {
                            s_prefix = prefix | ((uint64_t)(255 - 8 * lane - j) << shift);
                            s_want = want - run;
                            break;
                        }
                        run += h[j];
                    }
                }
            }
            __syncthreads();
        }
        thr_key = s_prefix;           // the kp-th largest key itself
    }
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t key = s_keys[i];
        if (key >= thr_key) top[(size_t)q * top_stride + atomicAdd(&s_out, 1u)] = key;   // exactly `keep` keys, in no particular order
    }
    if (threadIdx.x == 0) {
        ntop[q] = keep;
        thr[q] = keep == kp ? key_score(thr_key) : -INFINITY;
        count[q] = 0;
        if (ovf) {
            atomicOr(&flags[0], 1u);
            flags[4 + q] |= 1u;
        }
    }
}

// ------------------------------------------------------------------ fp64 re-score of the tensor path's survivors

// one warp per (query, survivor): exact key -> cand[q][j]
__global__ void __launch_bounds__(256) k_rerank(const __half *__restrict__ x, uint32_t d, const float *__restrict__ q,
                                                const uint64_t *__restrict__ top, uint32_t top_stride,
                                                const uint32_t *__restrict__ ntop, uint64_t *__restrict__ cand, uint32_t cap,
                                                uint32_t nq) {
    const uint32_t qi = blockIdx.y;
    const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (qi >= nq || j >= ntop[qi]) return;
    const uint32_t id = key_id(top[(size_t)qi * top_stride + j]);
    const uint4 *row = (const uint4 *)(x + (size_t)id * d);
    const float4 *qv = (const float4 *)(q + (size_t)qi * d);
    double acc = 0.0;
    for (uint32_t p = lane; p < (d >> 3); p += 32) {
        uint4 xv = __ldg(row + p);
        float4 qa = __ldg(qv + 2 * p), qb = __ldg(qv + 2 * p + 1);
        const __half2 *h = (const __half2 *)&xv;
        float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]), f2 = __half22float2(h[2]), f3 = __half22float2(h[3]);
        acc = fma((double)qa.x, (double)f0.x, acc);
        acc = fma((double)qa.y, (double)f0.y, acc);
        acc = fma((double)qa.z, (double)f1.x, acc);
        acc = fma((double)qa.w, (double)f1.y, acc);
        acc = fma((double)qb.x, (double)f2.x, acc);
        acc = fma((double)qb.y, (double)f2.y, acc);
        acc = fma((double)qb.z, (double)f3.x, acc);
        acc = fma((double)qb.w, (double)f3.y, acc);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) cand[(size_t)qi * cap + j] = rank_key((float)acc, id);
}

// ------------------------------------------------------------------ finalize: emit top-k (+ certificate on the tensor path)
//
// exact path  (rerank == 0): top[q] holds the exact keys of the k best rows (an unordered set); they are sorted here.
// tensor path (rerank == 1): cand[q][0..ntop) holds the exact keys of the approximate top-kp; sort them, emit
//   the best k and certify: every row that is NOT a survivor has approximate score <= thr[q] (the kp-th best
//   approximate score), hence exact score <= thr[q] + eps[q].  If the k-th exact score is strictly above that,
//   no excluded row can enter or tie the top-k, so ids and order equal the exhaustive exact result.
__global__ void __launch_bounds__(512) k_finalize(const uint64_t *__restrict__ top, uint32_t top_stride,
                                                  const uint64_t *__restrict__ cand, uint32_t cap,
                                                  const uint32_t *__restrict__ ntop, const float *__restrict__ thr,
                                                  const float *__restrict__ qstat, uint32_t k, uint32_t id_base, int rerank,
                                                  uint32_t *__restrict__ out_ids, float *__restrict__ out_scores,
                                                  uint64_t *__restrict__ out_keys, uint32_t out_stride, uint32_t *__restrict__ flags,
                                                  const uint32_t *__restrict__ qsel) {
    extern __shared__ uint64_t s_keys[];
    const uint32_t q = qsel ? qsel[blockIdx.x] : blockIdx.x;
    const uint32_t n = ntop[q];
    // the running list is an unordered set (k_select): sort it here, once -- the exact keys (rerank: the re-scored survivors)
    const uint64_t *src = s_keys;
    {
        const uint64_t *from = rerank ? cand + (size_t)q * cap : top + (size_t)q * top_stride;
        const uint32_t np2 = next_pow2_min64(n);
        for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x) s_keys[i] = i < n ? from[i] : 0ull;
        __syncthreads();
        bitonic_sort_desc(s_keys, np2);
    }
    if (rerank) {
        if (threadIdx.x == 0) {
            const float t = thr[q];
            bool ok = true;
            if (t > -INFINITY && n > 0) {  // the list was full: rows were excluded
                const uint32_t kk = k < n ? k : n;
                ok = (double)key_score(s_keys[kk - 1]) > (double)t + (double)qstat[q * 4 + 2];
            }
            if (!ok) {
                flags[4 + q] |= 2u;
                atomicAdd(&flags[1], 1u);
            }
        }
    }
    for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
        const bool have = i < n;
        const uint64_t key = have ? src[i] : 0ull;
        if (out_keys)   // sharded search: the rank key with the GLOBAL id, written straight into the all-gather send slot (0 = no entry)
            out_keys[(size_t)q * out_stride + i] = have ? rank_key(key_score(key), key_id(key) + id_base) : 0ull;
        if (out_ids) {
            out_ids[(size_t)q * out_stride + i] = have ? key_id(key) + id_base : MSE_ID_NONE;
            out_scores[(size_t)q * out_stride + i] = have ? key_score(key) : -INFINITY;
        }
    }
}

// compact the indices of queries whose status word is non-zero
__global__ void k_collect_flagged(const uint32_t *flags, uint32_t nq, uint32_t *sel, uint32_t *nsel) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq && flags[4 + i] != 0) sel[atomicAdd(nsel, 1u)] = i;
}

// ------------------------------------------------------------------ k-way merge of per-shard lists

__global__ void __launch_bounds__(256) k_merge_topk(const uint32_t *__restrict__ ids, const float *__restrict__ scores,
                                                    uint32_t n_shards, uint32_t nq, uint32_t k, uint32_t *__restrict__ out_ids,
                                                    float *__restrict__ out_scores) {
    extern __shared__ uint64_t s_keys[];
    const uint32_t q = blockIdx.x;
    const uint32_t n = n_shards * k;
    const uint32_t np2 = next_pow2_min64(n);
    for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x) {
        uint64_t key = 0ull;
        if (i < n) {
            uint32_t s = i / k, j = i % k;
            size_t off = ((size_t)s * nq + q) * k + j;
            uint32_t id = ids[off];
            key = id == MSE_ID_NONE ? 0ull : rank_key(scores[off], id);
        }
        s_keys[i] = key;
    }
    __syncthreads();
    bitonic_sort_desc(s_keys, np2);
    for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
        uint64_t key = s_keys[i];
        bool none = key == 0ull;
        out_ids[(size_t)q * k + i] = none ? MSE_ID_NONE : key_id(key);
        out_scores[(size_t)q * k + i] = none ? -INFINITY : key_score(key);
    }
}

// merge of all-gathered rank-key slots (sharded search epilogue): slots[r] holds [nq][k] keys of rank r at stride slot_stride
__global__ void __launch_bounds__(256) k_merge_keys(const uint64_t *__restrict__ slots, uint32_t n_shards, size_t slot_stride, uint32_t nq,
                                                    uint32_t k, uint32_t *__restrict__ out_ids, float *__restrict__ out_scores) {
    extern __shared__ uint64_t s_keys[];
    const uint32_t q = blockIdx.x;
    const uint32_t n = n_shards * k;
    const uint32_t np2 = next_pow2_min64(n);
    for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x) s_keys[i] = i < n ? slots[(size_t)(i / k) * slot_stride + (size_t)q * k + i % k] : 0ull;
    __syncthreads();
    bitonic_sort_desc(s_keys, np2);
    for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
        const uint64_t key = s_keys[i];
        out_ids[(size_t)q * k + i] = key ? key_id(key) : MSE_ID_NONE;
        out_scores[(size_t)q * k + i] = key ? key_score(key) : -INFINITY;
    }
}

// [0] = overflow seen, [1] = uncertified queries -> one status word behind the keys of the slot
__global__ void k_publish_status(const uint32_t *flags, uint64_t *word) { *word = (uint64_t)flags[0] | ((uint64_t)flags[1] << 32); }

// ================================================================== host side

static int ensure_select_smem(int device) {
    static PerDeviceOnce once;
    if (!once.first(device)) return MSE_OK;
    MSE_CUDA(cudaFuncSetAttribute(k_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSortMax * 8)));
    MSE_CUDA(cudaFuncSetAttribute(k_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSortMax * 8)));
    MSE_CUDA(cudaFuncSetAttribute(k_merge_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSortMax * 8)));
    MSE_CUDA(cudaFuncSetAttribute(k_merge_keys, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSortMax * 8)));
    return MSE_OK;
}

template <int QT>
static int launch_scan(int vpl, dim3 grid, cudaStream_t st, const __half *x, uint64_t row0, uint64_t nrows, uint32_t d,
                       const float *q, const uint32_t *qsel, uint32_t sel0, uint32_t nsel, const float *thr, uint64_t *cand,
                       uint32_t *count, uint32_t cap) {
#define MSE_SCAN_CASE(V)                                                                                        \
    case V:                                                                                                     \
        k_flat_scan<V, QT><<<grid, 256, 0, st>>>(x, row0, nrows, d, q, qsel, sel0, nsel, thr, cand, count, cap); \
        break;
    switch (vpl) {
        MSE_SCAN_CASE(1) MSE_SCAN_CASE(2) MSE_SCAN_CASE(3) MSE_SCAN_CASE(4) MSE_SCAN_CASE(5) MSE_SCAN_CASE(6)
        MSE_SCAN_CASE(7) MSE_SCAN_CASE(8)
        default:
            set_error("flat scan: d=%u unsupported (d %% 8 == 0, d <= 2048)", d);
            return MSE_ERR_UNSUPPORTED;
    }
#undef MSE_SCAN_CASE
    MSE_LAUNCH_OK();
    return MSE_OK;
}

// CUDA-event bracket around one scoring launch (only when profiling is on)
static void prof_mark(mse_index *ix, cudaStream_t st) {
    if (!ix->profile) return;
    if (ix->prof_used == ix->prof_ev.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ix->prof_ev.push_back(e);
    }
    cudaEventRecord(ix->prof_ev[ix->prof_used++], st);
}
static void prof_collect(mse_index *ix) {
    if (!ix->profile) return;
    double us = 0;
    for (size_t i = 0; i + 1 < ix->prof_used; i += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ix->prof_ev[i], ix->prof_ev[i + 1]) == cudaSuccess) us += ms * 1000.0;
    }
    ix->stats[6] = (uint64_t)(us * 1000.0);  // nanoseconds inside scoring kernels
    ix->stats[7] = ix->prof_used / 2;
}

// Chunk schedule.  After N rows the threshold is the kp-th best of those N, so under exchangeable row order a further chunk of
// g*N rows contributes ~g*kp candidates per query.  Every chunk costs a launch + a select, every candidate an append in the scoring
// epilogue (whose threshold is stale for the whole chunk): measured at 1024 queries (r02, MSE_FLAT_GROWTH): g = 1 / 2 / 4 / 12 ->
// 19.8 / 19.7 / 19.5 / 19.7 ms over 10 M rows and 2.63 / 2.57 / 2.54 / 2.58 ms over 1.25 M (the 8-GPU shard).  g = 4, capped so that
// g*kp stays below a third of the candidate buffer.  Row orders that defeat the estimate overflow the buffer, which is detected and
// repaired by the exact re-run (fixed chunks).
struct ChunkPlan {
    std::vector<std::pair<uint64_t, uint64_t>> chunks;  // (row0, nrows)
    ChunkPlan(uint64_t n, uint64_t fixed, uint32_t kp, uint32_t cap) {
        static const long g_env = getenv("MSE_FLAT_GROWTH") ? atol(getenv("MSE_FLAT_GROWTH")) : 0;   // tuning aid
        const uint64_t g_max = std::min<uint64_t>(15, std::max<uint64_t>(1, cap / (3ull * std::max(kp, 1u))));
        const uint64_t g = std::min<uint64_t>(g_max, g_env > 0 ? (uint64_t)g_env : kGrowth);
        uint64_t row0 = 0;
        while (row0 < n) {
            const uint64_t sz = fixed ? fixed : (row0 == 0 ? kChunk0 : g * row0);
            const uint64_t m = std::min(sz, n - row0);
            chunks.push_back({row0, m});
            row0 += m;
        }
    }
};

// exact scan of the queries listed in d_sel[0..nsel) (device array); results land in top/ntop of those queries
static int run_exact(mse_index *ix, const float *d_q, const uint32_t *d_sel, uint32_t nsel, uint32_t kp, uint32_t cap,
                     uint64_t fixed_chunk, cudaStream_t st) {
    FlatWork &w = ix->fw;
    const int sms = sm_count(ix->device);
    const int vpl = (int)((ix->d / 8 + 31) / 32);
    ChunkPlan plan(ix->n, fixed_chunk, kp, cap);
    k_reset_state<<<(nsel + 255) / 256, 256, 0, st>>>(w.ntop.as<uint32_t>(), w.thr.as<float>(), w.count.as<uint32_t>(),
                                                     w.flags.as<uint32_t>(), d_sel, nsel, 0);
    MSE_LAUNCH_OK();
    const uint32_t qt_max = vpl <= 5 ? 2u : 1u;  // two fp64 query copies only fit the register file up to d = 1280
    for (uint32_t s0 = 0; s0 < nsel; s0 += qt_max) {
        const uint32_t here = std::min(qt_max, nsel - s0);
        for (auto &c : plan.chunks) {
            uint64_t warps_needed = c.second;
            uint32_t blocks = (uint32_t)std::min<uint64_t>((warps_needed + 7) / 8, (uint64_t)sms * 8);
            if (blocks == 0) blocks = 1;
            prof_mark(ix, st);
            if (here == 2) {
                MSE_CHECK(launch_scan<2>(vpl, dim3(blocks), st, ix->x, c.first, c.second, ix->d, d_q, d_sel, s0, nsel,
                                         w.thr.as<float>(), w.cand.as<uint64_t>(), w.count.as<uint32_t>(), cap));
            } else {
                MSE_CHECK(launch_scan<1>(vpl, dim3(blocks), st, ix->x, c.first, c.second, ix->d, d_q, d_sel, s0, nsel,
                                         w.thr.as<float>(), w.cand.as<uint64_t>(), w.count.as<uint32_t>(), cap));
            }
            prof_mark(ix, st);
            k_select<<<here, 512, kSortMax * 8, st>>>(w.top.as<uint64_t>(), kKpMax, kp, w.ntop.as<uint32_t>(), w.thr.as<float>(),
                                                      w.cand.as<uint64_t>(), cap, w.count.as<uint32_t>(), w.flags.as<uint32_t>(),
                                                      d_sel + s0);
            MSE_LAUNCH_OK();
            ix->stats[5]++;
        }
    }
    return MSE_OK;
}

static int read_flags(mse_index *ix, cudaStream_t st, uint32_t out[4]) {
    MSE_CUDA(cudaMemcpyAsync(out, ix->fw.flags.p, 16, cudaMemcpyDeviceToHost, st));
    MSE_CUDA(cudaStreamSynchronize(st));
    return MSE_OK;
}

static int launch_finalize(mse_index *ix, uint32_t n_ctas, uint32_t k, uint32_t cap, int rerank, const FlatOut &o, const uint32_t *d_sel, cudaStream_t st) {
    FlatWork &w = ix->fw;
    k_finalize<<<n_ctas, 512, kSortMax * 8, st>>>(w.top.as<uint64_t>(), kKpMax, w.cand.as<uint64_t>(), cap, w.ntop.as<uint32_t>(), w.thr.as<float>(),
                                                  w.qstat.as<float>(), k, ix->id_base, rerank, o.ids, o.scores, o.keys, k, w.flags.as<uint32_t>(), d_sel);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

// Queues one search on `st` and returns without synchronising.  What is left on the device afterwards: the results, and the
// status word flags[0..1] (candidate-buffer overflow seen / number of queries whose cut could not be certified).
// flat_search_settle reads it and repairs the flagged queries.
int flat_search_queue(mse_index *ix, const float *d_q, uint32_t nq, uint32_t k, const FlatOut &out, cudaStream_t st) {
    MSE_CHECK(use_device(ix->device));
    MSE_CHECK(ensure_select_smem(ix->device));
    MSE_REQUIRE(k >= 1, MSE_ERR_INVALID, "search_flat: k must be >= 1");
    MSE_REQUIRE(ix->d % 8 == 0 && ix->d <= 2048, MSE_ERR_UNSUPPORTED, "search_flat: d=%u unsupported", ix->d);
    memset(ix->stats, 0, sizeof(ix->stats));
    ix->prof_used = 0;
    ix->pending = FlatPending{};
    const uint64_t launches0 = g_launches.load();
    if (nq == 0) return MSE_OK;
    FlatWork &w = ix->fw;

    // 1-2 queries over a small index: the exact fp64 scan has the lower latency.  Over a large index the scan is bound by its fp64 /
    // conversion instruction stream (5.9 ms for one query over 23 GB, 0.60 of the HBM copy rate; tools/flat_latency.py) while the
    // tensor pass runs at the HBM rate (3.6 ms, 0.97), so any batch takes the tensor pass there.
    bool tensor = ix->flat_mode == 2 || (ix->flat_mode == 0 && (nq > 2 || ix->n >= (1u << 18)));
    if (tensor && !flat_tc_supported(ix)) {
        MSE_REQUIRE(ix->flat_mode != 2, MSE_ERR_UNSUPPORTED, "search_flat: tensor path needs d %% 64 == 0 (d=%u)", ix->d);
        tensor = false;
    }
    uint32_t kp = k;
    if (tensor) {
        kp = ((k + k / 4 + 32) + 31) / 32 * 32;
        if (kp > kKpMax) tensor = false, kp = k;
    }
    MSE_REQUIRE(kp <= kKpMax, MSE_ERR_UNSUPPORTED, "search_flat: k=%u exceeds the supported maximum %u", k, kKpMax);
    const uint32_t cap = kSortMax - kKpMax;  // 6144: any chunk-0 (4096 rows) fits with the running list
    const uint32_t nq_pad = (nq + 127) / 128 * 128;

    MSE_CHECK(w.top.ensure((size_t)nq * kKpMax * 8));
    MSE_CHECK(w.ntop.ensure((size_t)nq * 4));
    MSE_CHECK(w.thr.ensure((size_t)nq * 4));
    MSE_CHECK(w.cand.ensure((size_t)nq * cap * 8));
    MSE_CHECK(w.count.ensure((size_t)nq * 4));
    MSE_CHECK(w.flags.ensure((size_t)(nq + 4) * 4));
    MSE_CHECK(w.sel.ensure((size_t)(nq + 1) * 4));
    MSE_CHECK(w.qstat.ensure((size_t)nq * 16));
    if (tensor) MSE_CHECK(w.q16.ensure((size_t)nq_pad * ix->d * 2));

    k_reset_state<<<(std::max(nq, 4u) + 255) / 256, 256, 0, st>>>(w.ntop.as<uint32_t>(), w.thr.as<float>(), w.count.as<uint32_t>(),
                                                                  w.flags.as<uint32_t>(), nullptr, nq, 1);
    MSE_LAUNCH_OK();
    k_iota<<<(nq + 255) / 256, 256, 0, st>>>(w.sel.as<uint32_t>(), nq);
    MSE_LAUNCH_OK();

    if (ix->n == 0) {
        MSE_CHECK(launch_finalize(ix, nq, k, cap, 0, out, nullptr, st));
    } else if (!tensor) {
        ix->stats[1] = nq;
        MSE_CHECK(run_exact(ix, d_q, w.sel.as<uint32_t>(), nq, kp, cap, 0, st));
        MSE_CHECK(launch_finalize(ix, nq, k, cap, 0, out, nullptr, st));
    } else {
        ix->stats[0] = nq;
        k_prepare_queries<<<(nq_pad * 32 + 255) / 256, 256, 0, st>>>(d_q, nq, nq_pad, ix->d, w.q16.as<__half>(), w.qstat.as<float>(),
                                                                     ix->max_norm);
        MSE_LAUNCH_OK();
        ChunkPlan plan(ix->n, 0, kp, cap);
        for (auto &c : plan.chunks) {
            prof_mark(ix, st);
            MSE_CHECK(flat_tc_score_chunk(ix, nq, c.first, c.second, cap, st));
            prof_mark(ix, st);
            k_select<<<nq, 512, kSortMax * 8, st>>>(w.top.as<uint64_t>(), kKpMax, kp, w.ntop.as<uint32_t>(), w.thr.as<float>(),
                                                    w.cand.as<uint64_t>(), cap, w.count.as<uint32_t>(), w.flags.as<uint32_t>(),
                                                    nullptr);
            MSE_LAUNCH_OK();
            ix->stats[5]++;
        }
        k_rerank<<<dim3((kp * 32 + 255) / 256, nq), 256, 0, st>>>(ix->x, ix->d, d_q, w.top.as<uint64_t>(), kKpMax, w.ntop.as<uint32_t>(),
                                                                 w.cand.as<uint64_t>(), cap, nq);
        MSE_LAUNCH_OK();
        MSE_CHECK(launch_finalize(ix, nq, k, cap, 1, out, nullptr, st));
    }
    ix->stats[4] = g_launches.load() - launches0;
    ix->pending = FlatPending{d_q, nq, k, out, st, ix->n != 0};
    return MSE_OK;
}

// Synchronises `st`, reads the status word of the last queued search and re-runs flagged queries on the exact scan; if that
// overflowed too (adversarial row order), once more with fixed chunks no larger than the candidate buffer, which cannot overflow.
// *repaired (optional) = number of queries that were re-run.
int flat_search_settle(mse_index *ix, uint32_t *repaired) {
    if (repaired) *repaired = 0;
    FlatPending p = ix->pending;
    if (!p.live) {
        if (p.st_valid()) MSE_CUDA(cudaStreamSynchronize(p.st));
        return MSE_OK;
    }
    MSE_CHECK(use_device(ix->device));
    FlatWork &w = ix->fw;
    cudaStream_t st = p.st;
    const uint32_t nq = p.nq, k = p.k, cap = kSortMax - kKpMax;
    const uint64_t launches0 = g_launches.load();
    uint32_t fl[4];
    MSE_CHECK(read_flags(ix, st, fl));
    int guard = 0;
    uint64_t fixed = 0;
    while (fl[0] != 0 || fl[1] != 0) {
        MSE_REQUIRE(++guard <= 2, MSE_ERR_CUDA, "search_flat: exact re-run did not converge");
        ix->stats[2] += fl[1];
        ix->stats[3] += fl[0];
        uint32_t *nsel_d = w.flags.as<uint32_t>() + 2;
        MSE_CUDA(cudaMemsetAsync(nsel_d, 0, 4, st));
        k_collect_flagged<<<(nq + 255) / 256, 256, 0, st>>>(w.flags.as<uint32_t>(), nq, w.sel.as<uint32_t>(), nsel_d);
        MSE_LAUNCH_OK();
        MSE_CHECK(read_flags(ix, st, fl));
        const uint32_t nsel = fl[2];
        MSE_CUDA(cudaMemsetAsync(w.flags.p, 0, 16, st));
        ix->stats[1] += nsel;
        if (repaired) *repaired += nsel;
        MSE_CHECK(run_exact(ix, p.d_q, w.sel.as<uint32_t>(), nsel, k, cap, fixed, st));
        MSE_CHECK(launch_finalize(ix, nsel, k, cap, 0, p.out, w.sel.as<uint32_t>(), st));
        MSE_CHECK(read_flags(ix, st, fl));
        fixed = cap;
    }
    ix->stats[4] += g_launches.load() - launches0;
    ix->pending.live = false;
    prof_collect(ix);
    return MSE_OK;
}

int flat_publish_status(mse_index *ix, uint64_t *d_word, cudaStream_t st) {
    if (!ix->fw.flags.p) { MSE_CUDA(cudaMemsetAsync(d_word, 0, 8, st)); return MSE_OK; }
    k_publish_status<<<1, 1, 0, st>>>(ix->fw.flags.as<uint32_t>(), d_word);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

int flat_merge_keys(int device, const uint64_t *d_slots, uint32_t n_shards, size_t slot_stride, uint32_t nq, uint32_t k, uint32_t *d_ids,
                    float *d_scores, cudaStream_t st) {
    MSE_CHECK(ensure_select_smem(device));
    MSE_REQUIRE((uint64_t)n_shards * k <= kSortMax, MSE_ERR_UNSUPPORTED, "sharded search: n_shards*k=%llu exceeds %u", (unsigned long long)n_shards * k, kSortMax);
    if (nq == 0) return MSE_OK;
    k_merge_keys<<<nq, 256, kSortMax * 8, st>>>(d_slots, n_shards, slot_stride, nq, k, d_ids, d_scores);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

}  // namespace mse

using namespace mse;

// ================================================================== C ABI

static int index_reserve(mse_index *ix, uint64_t rows, bool exact = false) {
    if (rows <= ix->cap) return MSE_OK;
    uint64_t ncap = exact ? rows : std::max<uint64_t>(rows, ix->cap + ix->cap / 2);
    __half *nx = nullptr;
    cudaError_t e = cudaMalloc(&nx, ncap * ix->d * sizeof(__half));
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("index: cudaMalloc(%llu rows x %u) -> %s", (unsigned long long)ncap, ix->d, cudaGetErrorString(e));
        return MSE_ERR_OOM;
    }
    if (ix->n) MSE_CUDA(cudaMemcpyAsync(nx, ix->x, ix->n * ix->d * sizeof(__half), cudaMemcpyDeviceToDevice, ix->stream));
    MSE_CUDA(cudaStreamSynchronize(ix->stream));
    if (ix->x) cudaFree(ix->x);
    ix->x = nx;
    ix->cap = ncap;
    ix->tmap_valid = false;
    return MSE_OK;
}

namespace mse {
void index_drop_side_arrays(mse_index *ix) {
    if (ix->adj) cudaFree(ix->adj);
    if (ix->deg) cudaFree(ix->deg);
    if (ix->pq_codes) cudaFree(ix->pq_codes);
    if (ix->code_scale) cudaFree(ix->code_scale);
    if (ix->desc) cudaFree(ix->desc);
    if (ix->has_url) cudaFree(ix->has_url);
    ix->adj = ix->deg = nullptr;
    ix->pq_codes = ix->desc = ix->has_url = nullptr;
    ix->code_scale = nullptr;
    ix->graph_stride = ix->code_size = ix->n_desc = 0;
    ix->side_n = 0;
}
}  // namespace mse

// VectorList::push after IndexGraph::empty(n, r) has no counterpart in the reference (the graph is sized from the final
// vector count, generate_index_shard.rs:102); here growing the row store invalidates every per-row side array
static void index_rows_grew(mse_index *ix) {
    if (ix->adj || ix->deg || ix->pq_codes || ix->code_scale || ix->desc || ix->has_url) index_drop_side_arrays(ix);
}

static int index_note_rows(mse_index *ix, uint64_t first, uint64_t n, cudaStream_t st) {
    if (n == 0) return MSE_OK;
    uint32_t blocks = (uint32_t)std::min<uint64_t>((n + 7) / 8, (uint64_t)sm_count(ix->device) * 8);
    k_row_norm_max<<<blocks, 256, 0, st>>>(ix->x + first * ix->d, n, ix->d, ix->max_norm);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

MSE_API int mse_index_create(const uint16_t *x_f16, uint64_t n, uint32_t d, int device, uint32_t id_base, mse_index **out) {
    MSE_REQUIRE(out != nullptr, MSE_ERR_INVALID, "index_create: out is NULL");
    *out = nullptr;
    MSE_REQUIRE(d > 0 && d % 8 == 0, MSE_ERR_INVALID, "index_create: d=%u must be a positive multiple of 8", d);
    MSE_REQUIRE(n == 0 || x_f16 != nullptr, MSE_ERR_INVALID, "index_create: x_f16 is NULL with n=%llu", (unsigned long long)n);
    MSE_REQUIRE(n < 0xFFFFFFFFull, MSE_ERR_INVALID, "index_create: n must be < 2^32 - 1 per shard");
    MSE_CHECK(use_device(device));
    mse_index *ix = new mse_index();
    ix->device = device;
    ix->d = d;
    ix->id_base = id_base;
    int rc = MSE_OK;
    do {
        if (cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = MSE_ERR_CUDA; set_error("index_create: stream"); break; }
        if (cudaMalloc(&ix->max_norm, 4) != cudaSuccess) { rc = MSE_ERR_OOM; set_error("index_create: alloc"); break; }
        if (cudaMemsetAsync(ix->max_norm, 0, 4, ix->stream) != cudaSuccess) { rc = MSE_ERR_CUDA; set_error("index_create: memset"); break; }
        if (n) {
            rc = mse_index_add_f16(ix, x_f16, n);
            if (rc != MSE_OK) break;
        }
    } while (0);
    if (rc != MSE_OK) {
        mse_index_destroy(ix);
        return rc;
    }
    *out = ix;
    return MSE_OK;
}

MSE_API int mse_index_add_f16(mse_index *ix, const uint16_t *x_f16, uint64_t n) {
    MSE_REQUIRE(ix != nullptr, MSE_ERR_INVALID, "index_add_f16: NULL handle");
    if (n == 0) return MSE_OK;
    MSE_REQUIRE(x_f16 != nullptr, MSE_ERR_INVALID, "index_add_f16: NULL rows");
    MSE_REQUIRE(ix->n + n < 0xFFFFFFFFull, MSE_ERR_INVALID, "index_add_f16: shard would exceed 2^32 - 2 rows");
    MSE_CHECK(use_device(ix->device));
    MSE_CHECK(index_reserve(ix, ix->n + n));
    MSE_CUDA(cudaMemcpyAsync(ix->x + ix->n * ix->d, x_f16, n * ix->d * sizeof(__half), cudaMemcpyHostToDevice, ix->stream));
    MSE_CHECK(index_note_rows(ix, ix->n, n, ix->stream));
    MSE_CUDA(cudaStreamSynchronize(ix->stream));
    ix->n += n;
    index_rows_grew(ix);
    return MSE_OK;
}

MSE_API int mse_index_add_f16_dev(mse_index *ix, const uint16_t *d_x_f16, uint64_t n, void *stream) {
    MSE_REQUIRE(ix != nullptr, MSE_ERR_INVALID, "index_add_f16_dev: NULL handle");
    if (n == 0) return MSE_OK;
    MSE_REQUIRE(d_x_f16 != nullptr, MSE_ERR_INVALID, "index_add_f16_dev: NULL rows");
    MSE_REQUIRE(ix->n + n < 0xFFFFFFFFull, MSE_ERR_INVALID, "index_add_f16_dev: shard would exceed 2^32 - 2 rows");
    MSE_CHECK(use_device(ix->device));
    cudaStream_t st = (cudaStream_t)stream;
    MSE_CUDA(cudaStreamSynchronize(st));
    MSE_CHECK(index_reserve(ix, ix->n + n));
    MSE_CUDA(cudaMemcpyAsync(ix->x + ix->n * ix->d, d_x_f16, n * ix->d * sizeof(__half), cudaMemcpyDeviceToDevice, st));
    MSE_CHECK(index_note_rows(ix, ix->n, n, st));
    MSE_CUDA(cudaStreamSynchronize(st));
    ix->n += n;
    index_rows_grew(ix);
    return MSE_OK;
}

MSE_API int mse_index_add(mse_index *ix, const float *x_f32, uint64_t n) {
    MSE_REQUIRE(ix != nullptr, MSE_ERR_INVALID, "index_add: NULL handle");
    if (n == 0) return MSE_OK;
    MSE_REQUIRE(x_f32 != nullptr, MSE_ERR_INVALID, "index_add: NULL rows");
    MSE_REQUIRE(ix->n + n < 0xFFFFFFFFull, MSE_ERR_INVALID, "index_add: shard would exceed 2^32 - 2 rows");
    MSE_CHECK(use_device(ix->device));
    MSE_CHECK(index_reserve(ix, ix->n + n));
    float *stage = nullptr;
    const uint64_t elems = n * ix->d;
    MSE_CUDA(cudaMalloc(&stage, elems * sizeof(float)));
    int rc = MSE_OK;
    do {
        if (cudaMemcpyAsync(stage, x_f32, elems * sizeof(float), cudaMemcpyHostToDevice, ix->stream) != cudaSuccess) { rc = MSE_ERR_CUDA; set_error("index_add: H2D failed"); break; }
        uint32_t blocks = (uint32_t)std::min<uint64_t>((elems + 255) / 256, (uint64_t)sm_count(ix->device) * 16);
        k_f32_to_f16<<<blocks, 256, 0, ix->stream>>>(stage, ix->x + ix->n * ix->d, elems);
        count_launch();
        rc = index_note_rows(ix, ix->n, n, ix->stream);
        if (rc != MSE_OK) break;
        if (cudaStreamSynchronize(ix->stream) != cudaSuccess) { rc = MSE_ERR_CUDA; set_error("index_add: sync failed"); break; }
    } while (0);
    cudaFree(stage);
    if (rc == MSE_OK) {
        ix->n += n;
        index_rows_grew(ix);
    }
    return rc;
}

MSE_API int mse_index_reserve(mse_index *ix, uint64_t rows) {
    MSE_REQUIRE(ix != nullptr, MSE_ERR_INVALID, "index_reserve: NULL handle");
    MSE_REQUIRE(rows < 0xFFFFFFFFull, MSE_ERR_INVALID, "index_reserve: shard would exceed 2^32 - 2 rows");
    MSE_CHECK(use_device(ix->device));
    return index_reserve(ix, rows, true);
}

MSE_API uint64_t mse_index_ntotal(const mse_index *ix) { return ix ? ix->n : 0; }
MSE_API uint32_t mse_index_dim(const mse_index *ix) { return ix ? ix->d : 0; }
MSE_API int mse_index_device(const mse_index *ix) { return ix ? ix->device : -1; }
MSE_API const uint16_t *mse_index_vectors_dev(const mse_index *ix) { return ix ? (const uint16_t *)ix->x : nullptr; }

MSE_API void mse_index_destroy(mse_index *ix) {
    if (!ix) return;
    cudaSetDevice(ix->device);
    if (ix->stream) cudaStreamSynchronize(ix->stream);
    ix->fw.release();
    ix->gw_htabs.release();
    ix->gw_status.release();
    ix->gw_vis_ids.release();
    ix->gw_vis_sc.release();
    ix->gw_vis_len.release();
    for (cudaEvent_t e : ix->prof_ev) cudaEventDestroy(e);
    if (ix->x) cudaFree(ix->x);
    if (ix->max_norm) cudaFree(ix->max_norm);
    if (ix->adj) cudaFree(ix->adj);
    if (ix->deg) cudaFree(ix->deg);
    if (ix->pq_codes) cudaFree(ix->pq_codes);
    if (ix->code_scale) cudaFree(ix->code_scale);
    if (ix->desc) cudaFree(ix->desc);
    if (ix->has_url) cudaFree(ix->has_url);
    if (ix->stream) cudaStreamDestroy(ix->stream);
    delete ix;
}

MSE_API int mse_search_flat_dev(mse_index *ix, const float *d_q, uint32_t nq, uint32_t k, uint32_t *d_ids, float *d_scores,
                                void *stream) {
    MSE_REQUIRE(ix != nullptr, MSE_ERR_INVALID, "search_flat_dev: NULL handle");
    MSE_REQUIRE(nq == 0 || (d_q && d_ids && d_scores), MSE_ERR_INVALID, "search_flat_dev: NULL buffer");
    return flat_search_queue(ix, d_q, nq, k, FlatOut{d_ids, d_scores, nullptr}, (cudaStream_t)stream);
}

MSE_API int mse_search_flat_check(mse_index *ix, uint32_t *repaired) {
    MSE_REQUIRE(ix != nullptr, MSE_ERR_INVALID, "search_flat_check: NULL handle");
    return flat_search_settle(ix, repaired);
}

MSE_API int mse_search_flat(mse_index *ix, const float *q, uint32_t nq, uint32_t k, uint32_t *ids, float *scores) {
    MSE_REQUIRE(ix != nullptr, MSE_ERR_INVALID, "search_flat: NULL handle");
    MSE_REQUIRE(nq == 0 || (q && ids && scores), MSE_ERR_INVALID, "search_flat: NULL buffer");
    MSE_REQUIRE(k >= 1, MSE_ERR_INVALID, "search_flat: k must be >= 1");
    if (nq == 0) return MSE_OK;
    MSE_CHECK(use_device(ix->device));
    FlatWork &w = ix->fw;
    MSE_CHECK(w.q.ensure((size_t)nq * ix->d * 4));
    MSE_CHECK(w.out_ids.ensure((size_t)nq * k * 4));
    MSE_CHECK(w.out_sc.ensure((size_t)nq * k * 4));
    MSE_CUDA(cudaMemcpyAsync(w.q.p, q, (size_t)nq * ix->d * 4, cudaMemcpyHostToDevice, ix->stream));
    MSE_CHECK(flat_search_queue(ix, w.q.as<float>(), nq, k, FlatOut{w.out_ids.as<uint32_t>(), w.out_sc.as<float>(), nullptr}, ix->stream));
    MSE_CHECK(flat_search_settle(ix, nullptr));
    MSE_CUDA(cudaMemcpyAsync(ids, w.out_ids.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ix->stream));
    MSE_CUDA(cudaMemcpyAsync(scores, w.out_sc.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ix->stream));
    MSE_CUDA(cudaStreamSynchronize(ix->stream));
    return MSE_OK;
}

MSE_API int mse_query_accumulate_f16_dev(int device, const uint16_t *d_e_f16, float weight, float *d_q, uint32_t nq, uint32_t d, void *stream) {
    MSE_REQUIRE(nq == 0 || (d_e_f16 && d_q), MSE_ERR_INVALID, "query_accumulate_f16_dev: NULL buffer");
    if (nq == 0 || d == 0) return MSE_OK;
    MSE_CHECK(use_device(device));
    const uint64_t n = (uint64_t)nq * d;
    k_query_axpy_f16<<<(uint32_t)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __half *)d_e_f16, weight, d_q, n);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

MSE_API int mse_search_flat_stats(const mse_index *ix, uint64_t out[8]) {
    MSE_REQUIRE(ix != nullptr && out != nullptr, MSE_ERR_INVALID, "search_flat_stats: NULL argument");
    if (ix->profile && ix->prof_used) {   // event times need the launches to have finished
        mse_index *m = const_cast<mse_index *>(ix);
        if (ix->pending.st_valid()) cudaStreamSynchronize(ix->pending.st);
        prof_collect(m);
    }
    memcpy(out, ix->stats, sizeof(ix->stats));
    return MSE_OK;
}

MSE_API int mse_search_flat_profile(mse_index *ix, int enable) {
    MSE_REQUIRE(ix != nullptr, MSE_ERR_INVALID, "search_flat_profile: NULL handle");
    ix->profile = enable ? 1 : 0;
    return MSE_OK;
}

MSE_API int mse_search_flat_set_mode(mse_index *ix, int mode) {
    MSE_REQUIRE(ix != nullptr && mode >= 0 && mode <= 2, MSE_ERR_INVALID, "search_flat_set_mode: bad argument");
    ix->flat_mode = mode;
    return MSE_OK;
}

MSE_API int mse_merge_topk_dev(int device, const uint32_t *d_ids, const float *d_scores, uint32_t n_shards, uint32_t nq,
                               uint32_t k, uint32_t *d_out_ids, float *d_out_scores, void *stream) {
    MSE_CHECK(use_device(device));
    MSE_CHECK(ensure_select_smem(device));
    MSE_REQUIRE(k >= 1 && n_shards >= 1, MSE_ERR_INVALID, "merge_topk: bad shape");
    MSE_REQUIRE((uint64_t)n_shards * k <= kSortMax, MSE_ERR_UNSUPPORTED, "merge_topk: n_shards*k=%llu exceeds %u",
                (unsigned long long)n_shards * k, kSortMax);
    if (nq == 0) return MSE_OK;
    k_merge_topk<<<nq, 256, kSortMax * 8, (cudaStream_t)stream>>>(d_ids, d_scores, n_shards, nq, k, d_out_ids, d_out_scores);
    MSE_LAUNCH_OK();
    return MSE_OK;
}
