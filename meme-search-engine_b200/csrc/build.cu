// Vamana graph construction on the GPU.
//
//   k_robust_prune     diskann/src/lib.rs:227-285  robust_prune, including the skip-one start of the inner loop (:250),
//                      the fixed-point alpha `(alpha * s) >> 16` (:266) and the saturate pass (:275-284); bit-exact
//                      against the CPU oracle for a given candidate list (fast_dot scores, stable descending order)
//   k_random_fill      diskann/src/lib.rs:376-387  uniform neighbours, duplicates rejected, self-loops allowed
//   k_centroid / medioid  diskann/src/lib.rs:54-68 running-mean centroid in f32, rounded to fp16, argmax dot (last max wins)
//   build_vamana       diskann/src/lib.rs:287-324  build_graph, restated batch-synchronously: the reference inserts points
//                      concurrently under per-node locks (rayon, order-dependent); here every batch of points is searched
//                      against the graph as it stood before the batch, pruned, and its back-edges are merged per target
//                      node (the ParlayANN scheme the reference's own comment at lib.rs:14-15 points to).  Batch sizes
//                      double from 1 up to a cap so early points see each other.  Structural parity with the reference is
//                      statistical by nature (its build is racy by design); unit parity is on robust_prune.
#include "graph.cuh"
#include "fastdot.cuh"
#include <algorithm>
#include <vector>

namespace mse {

static constexpr int kPruneThreads = 256;
static constexpr int kPruneWarps = kPruneThreads / 32;
static constexpr int kSortN = 2048;      // bitonic window: best 1024 so far + next 1024 candidates
static constexpr int kKeep = 1024;       // maxc <= kKeep
static constexpr int kMaxIn = 192;       // back-edges merged into one node per batch (beyond that they are dropped)
static constexpr long long kDead = (long long)0x8000000000000000ull;  // i64::MIN tombstone (lib.rs:241,252,269)

struct PruneCfg {
    uint32_t r, maxc;
    long long alpha, query_alpha;
    uint32_t query_breakpoint;
    int saturate;
};

// "a before b": score descending, arrival order ascending (the oracle's resolution of sort_unstable's tie freedom)
__device__ __forceinline__ bool cand_before(long long sa, uint32_t oa, long long sb, uint32_t ob) {
    return sa > sb || (sa == sb && oa < ob);
}

__device__ __forceinline__ void bitonic_sort_cands(long long *sc, uint32_t *ord, int n /* power of two */) {
    for (int k2 = 2; k2 <= n; k2 <<= 1)
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & k2) == 0;
                    const long long sa = sc[i], sb = sc[ixj];
                    const uint32_t oa = ord[i], ob = ord[ixj];
                    const bool a_first = cand_before(sa, oa, sb, ob);
                    if (up ? !a_first : a_first) { sc[i] = sb; sc[ixj] = sa; ord[i] = ob; ord[ixj] = oa; }
                }
            }
            __syncthreads();
        }
}

// One CTA prunes one point.  Candidates arrive as up to two segments (visited list, then extra ids whose scores are
// computed here with fast_dot against the point: merge_existing_neighbours lib.rs:215-221).
//   seg A: a_ids/a_scores [a_len]           seg B: b_ids [b_len] (scored here)
template <int NC2>   // 64-element chunks per row (d / 64) known at compile time, or 0 for a runtime d
__device__ void prune_point(const __half *__restrict__ x, uint32_t d, uint32_t p, const uint32_t *a_ids, const long long *a_scores, uint32_t a_len,
                            const uint32_t *b_ids, uint32_t b_len, const PruneCfg &cfg, uint32_t *out, uint32_t *out_len, uint8_t *smem) {
    long long *sc = (long long *)smem;                 // [kSortN]
    uint32_t *ord = (uint32_t *)(sc + kSortN);         // [kSortN]  arrival index
    uint32_t *cid = ord + kSortN;                      // [kKeep]   candidate ids after the sort
    float *pv = (float *)(cid + kKeep);                // [d]       current p* (or p) as fp32
    uint32_t *res = (uint32_t *)(pv + d);              // [r]       chosen neighbours
    long long *bsc = (long long *)(smem + kSortN * 12 + kKeep * 4 + (((size_t)d * 4 + 256 * 4 + 15) & ~(size_t)15));  // [512] segment-B scores
    __shared__ int s_n, s_ci, s_nout, s_pick;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t total = a_len + b_len;

    // scores of segment B against p (merge_existing_neighbours)
    if (b_len) {
        for (uint32_t i = threadIdx.x; i < d; i += blockDim.x) pv[i] = __half2float(x[(size_t)p * d + i]);
        __syncthreads();
        for (uint32_t i = 2 * warp; i < b_len; i += 2 * kPruneWarps) {          // two rows per pass (fastdot.cuh wq_score2)
            const uint32_t j = i + 1 < b_len ? i + 1 : i;
            long long s0, s1;
            wq_score2<NC2>(pv, x + (size_t)b_ids[i] * d, x + (size_t)b_ids[j] * d, d, lane, s0, s1);
            if (lane == 0) { bsc[i] = s0; bsc[j] = s1; }
        }
        __syncthreads();
    }
    // keep the best kKeep of all candidates under (score desc, arrival asc): chunked bitonic over a 2048-entry window whose
    // lower half carries the best 1024 so far.  Arrival order: segment A first, then segment B.
    for (uint32_t i = threadIdx.x; i < b_len; i += blockDim.x) { sc[i] = bsc[i]; ord[i] = a_len + i; }
    __syncthreads();
    uint32_t filled = b_len, done = 0;
    bool first = true;
    while (done < a_len || first) {
        first = false;
        const uint32_t take = min((uint32_t)kSortN - filled, a_len - done);
        uint32_t sortn = 64;                                   // the window is only as large as what it has to hold
        while (sortn < filled + take) sortn <<= 1;
        for (uint32_t i = threadIdx.x; i < take; i += blockDim.x) { sc[filled + i] = a_scores[done + i]; ord[filled + i] = done + i; }
        for (uint32_t i = filled + take + threadIdx.x; i < sortn; i += blockDim.x) { sc[i] = kDead; ord[i] = 0xFFFFFFFFu; }
        __syncthreads();
        bitonic_sort_cands(sc, ord, (int)sortn);
        done += take;
        filled = min((uint32_t)kKeep, filled + take);
    }
    // candidates.truncate(maxc) (:234)
    const uint32_t nc = min(min(total, cfg.maxc), (uint32_t)kKeep);
    for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x) {
        const uint32_t o = ord[i];
        cid[i] = o < a_len ? a_ids[o] : b_ids[o - a_len];
    }
    if (threadIdx.x == 0) { s_n = (int)nc; s_ci = 0; s_nout = 0; }
    __syncthreads();

    // main loop (:236-272)
    for (;;) {
        if (threadIdx.x == 0) {
            int ci = s_ci, pick = -1;
            while (s_nout < (int)cfg.r && ci < s_n) {
                const uint32_t ps = cid[ci];
                const long long pss = sc[ci];
                ci++;
                if (ps == p || pss == kDead) continue;
                res[s_nout++] = ps;
                pick = ci - 1;
                break;
            }
            s_ci = ci;
            s_pick = pick;
        }
        __syncthreads();
        if (s_pick < 0) break;
        const uint32_t p_star = cid[s_pick];
        const int from = s_ci + 1;  // :250 -- one past the element after p*
        if (from < s_n) {
            for (uint32_t i = threadIdx.x; i < d; i += blockDim.x) pv[i] = __half2float(x[(size_t)p_star * d + i]);
            __syncthreads();
            // fast_dot(p', p*) for every live later candidate, two per pass (operand order does not matter: each product is
            // of the same pair of values).  Each warp owns a fixed pair of positions per stride, so no two warps touch one entry.
            for (int i = from + 2 * warp; i < s_n; i += 2 * kPruneWarps) {
                const int j = i + 1;
                const long long pps0 = sc[i], pps1 = j < s_n ? sc[j] : kDead;
                if (pps0 == kDead && pps1 == kDead) continue;  // warp-uniform
                const int i0 = pps0 != kDead ? i : j, i1 = pps1 != kDead ? j : i;     // a dead slot borrows the live row (result unused)
                long long sps0, sps1;
                wq_score2<NC2>(pv, x + (size_t)cid[i0] * d, x + (size_t)cid[i1] * d, d, lane, sps0, sps1);
                if (lane == 0) {
                    if (pps0 != kDead) {
                        const long long a = cid[i] >= cfg.query_breakpoint ? cfg.query_alpha : cfg.alpha;
                        const long long scaled = (long long)((unsigned long long)a * (unsigned long long)sps0) >> 16;  // wrapping mul, arithmetic shift
                        if (scaled >= pps0) sc[i] = kDead;
                    }
                    if (pps1 != kDead) {
                        const long long a = cid[j] >= cfg.query_breakpoint ? cfg.query_alpha : cfg.alpha;
                        const long long scaled = (long long)((unsigned long long)a * (unsigned long long)sps1) >> 16;
                        if (scaled >= pps1) sc[j] = kDead;
                    }
                }
            }
        }
        __syncthreads();
    }
    // saturate (:275-284)
    if (threadIdx.x == 0) {
        int nout = s_nout;
        if (cfg.saturate || p >= cfg.query_breakpoint) {
            for (int i = 0; i < s_n && nout < (int)cfg.r; i++) {
                bool present = false;
                for (int j = 0; j < nout; j++) present |= res[j] == cid[i];
                if (!present) res[nout++] = cid[i];
            }
        }
        s_nout = nout;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < s_nout; i += blockDim.x) out[i] = res[i];
    if (threadIdx.x == 0) *out_len = (uint32_t)s_nout;
    __syncthreads();
}

static size_t prune_smem_bytes(uint32_t d, uint32_t r) { (void)r; return (size_t)kSortN * 12 + (size_t)kKeep * 4 + (((size_t)d * 4 + 256 * 4 + 15) & ~(size_t)15) + 512 * 8 + 64; }

// standalone: one point, explicit candidates with scores (unit parity against the oracle)
template <int NC2>
__global__ void __launch_bounds__(kPruneThreads) k_robust_prune_one(const __half *x, uint32_t d, uint32_t p, const uint32_t *ids, const long long *scores,
                                                                    uint32_t n, PruneCfg cfg, uint32_t *out, uint32_t *out_len) {
    extern __shared__ __align__(16) uint8_t smem[];
    prune_point<NC2>(x, d, p, ids, scores, n, nullptr, 0, cfg, out, out_len, smem);
}

// batch step 1 (lib.rs:299-309): for batch point b: candidates = visited list of its search ++ its current neighbours; prune;
// the result goes to new_adj[b] (applied after the whole batch was pruned)
template <int NC2>
__global__ void __launch_bounds__(kPruneThreads) k_prune_batch(const __half *x, uint32_t d, const uint32_t *points, uint32_t nb, const uint32_t *vl_ids,
                                                               const long long *vl_scores, const uint32_t *vl_len, uint32_t vl_cap, const uint32_t *adj,
                                                               const uint32_t *deg, uint32_t stride, PruneCfg cfg, uint32_t *new_adj, uint32_t *new_deg) {
    extern __shared__ __align__(16) uint8_t smem[];
    for (uint32_t b = blockIdx.x; b < nb; b += gridDim.x) {
        const uint32_t p = points[b];
        const uint32_t n_vl = min(vl_len[b], vl_cap);
        prune_point<NC2>(x, d, p, vl_ids + (size_t)b * vl_cap, vl_scores + (size_t)b * vl_cap, n_vl, adj + (size_t)p * stride, min(deg[p], stride), cfg,
                    new_adj + (size_t)b * stride, new_deg + b, smem);
    }
}

// batch step 2: install the new lists and emit back-edges (lib.rs:310-312): for every n in N(p): incoming[n] += p
__global__ void k_apply_and_backedges(const uint32_t *points, uint32_t nb, const uint32_t *new_adj, const uint32_t *new_deg, uint32_t stride,
                                      uint32_t *adj, uint32_t *deg, uint32_t *incoming, uint32_t *in_cnt, uint32_t *touched, uint32_t *n_touched) {
    const uint32_t b = blockIdx.x;
    if (b >= nb) return;
    const uint32_t p = points[b], dg = new_deg[b];
    for (uint32_t i = threadIdx.x; i < dg; i += blockDim.x) {
        const uint32_t nbr = new_adj[(size_t)b * stride + i];
        adj[(size_t)p * stride + i] = nbr;
        const uint32_t slot = atomicAdd(&in_cnt[nbr], 1u);
        if (slot == 0) touched[atomicAdd(n_touched, 1u)] = nbr;
        if (slot < (uint32_t)kMaxIn) incoming[(size_t)nbr * kMaxIn + slot] = p;
    }
    if (threadIdx.x == 0) deg[p] = dg;
}

// batch step 3 (lib.rs:313-321): per touched node n: append the incoming points that are absent while there is room;
// once the list is full, re-prune n over its neighbours ++ the remaining incoming points.
// (The reference re-prunes once per incoming edge under the node's lock; merging them into one prune per batch is the
// batch-synchronous restatement.)
template <int NC2>
__global__ void __launch_bounds__(kPruneThreads) k_merge_backedges(const __half *x, uint32_t d, const uint32_t *touched, uint32_t n_touched,
                                                                   uint32_t *incoming, uint32_t *in_cnt, uint32_t *adj, uint32_t *deg, uint32_t stride,
                                                                   PruneCfg cfg) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t s_list[kMaxIn + 256];
    __shared__ uint32_t s_out[256];
    __shared__ uint32_t s_cnt, s_outlen;
    for (uint32_t t = blockIdx.x; t < n_touched; t += gridDim.x) {
        const uint32_t n = touched[t];
        const uint32_t nin = min(in_cnt[n], (uint32_t)kMaxIn);
        uint32_t *nl = adj + (size_t)n * stride;
        if (threadIdx.x == 0) {
            // sequential, tiny: current neighbours, then incoming ids that are not present yet (deterministic order: sorted)
            uint32_t cnt = min(deg[n], stride);
            for (uint32_t i = 0; i < cnt; i++) s_list[i] = nl[i];
            // sort incoming ascending so the result does not depend on atomic arrival order
            uint32_t *in = incoming + (size_t)n * kMaxIn;
            for (uint32_t i = 1; i < nin; i++) { uint32_t v = in[i]; int j = (int)i - 1; while (j >= 0 && in[j] > v) { in[j + 1] = in[j]; j--; } in[j + 1] = v; }
            for (uint32_t i = 0; i < nin; i++) {
                const uint32_t v = in[i];
                bool present = false;
                for (uint32_t j = 0; j < cnt; j++) present |= s_list[j] == v;
                if (!present) s_list[cnt++] = v;
            }
            s_cnt = cnt;
            in_cnt[n] = 0;
        }
        __syncthreads();
        const uint32_t cnt = s_cnt;
        if (cnt <= cfg.r) {
            for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) nl[i] = s_list[i];
            if (threadIdx.x == 0) deg[n] = cnt;
            __syncthreads();
        } else {
            prune_point<NC2>(x, d, n, nullptr, nullptr, 0, s_list, cnt, cfg, s_out, &s_outlen, smem);
            for (uint32_t i = threadIdx.x; i < s_outlen; i += blockDim.x) nl[i] = s_out[i];
            if (threadIdx.x == 0) deg[n] = s_outlen;
            __syncthreads();
        }
    }
}

// ---- random_fill_graph (lib.rs:376-387)
__device__ __forceinline__ uint64_t splitmix64(uint64_t &s) {
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__global__ void k_random_fill(uint32_t *adj, uint32_t *deg, uint64_t n, uint32_t stride, uint32_t r, uint64_t seed) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s = seed ^ (0x5851f42d4c957f2dull * (i + 1));
    uint32_t *nb = adj + i * stride;
    uint32_t dg = min(deg[i], stride);
    const uint64_t want64 = (uint64_t)min(r, stride) < n ? (uint64_t)min(r, stride) : n;
    const uint32_t want = (uint32_t)want64;
    while (dg < want) {
        const uint32_t c = (uint32_t)(((splitmix64(s) >> 32) * n) >> 32);
        bool dup = false;
        for (uint32_t j = 0; j < dg; j++) dup |= nb[j] == c;
        if (!dup) nb[dg++] = c;
    }
    deg[i] = dg;
}

// ---- centroid (lib.rs:54-62): c += (v - c) * (1 / (i + 1)) in f32, one thread per dimension, rows in order (no FMA: the
// reference does the subtract, multiply and add as separate f32 operations)
__global__ void k_centroid(const __half *__restrict__ x, uint64_t n, uint32_t d, __half *__restrict__ out16) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d) return;
    // the running mean is a sequential recurrence per dimension (that order is the definition, lib.rs:55-58); only the loads
    // can run ahead: 32 rows are fetched before the 32 dependent updates
    float c = 0.f;
    uint64_t i = 0;
    for (; i + 32 <= n; i += 32) {
        __half v[32];
#pragma unroll
        for (int u = 0; u < 32; u++) v[u] = x[(i + u) * d + j];
#pragma unroll
        for (int u = 0; u < 32; u++) {
            const float w = __fdiv_rn(1.0f, (float)(i + u + 1));
            c = __fadd_rn(c, __fmul_rn(__fsub_rn(__half2float(v[u]), c), w));
        }
    }
    for (; i < n; i++) {
        const float w = __fdiv_rn(1.0f, (float)(i + 1));
        c = __fadd_rn(c, __fmul_rn(__fsub_rn(__half2float(x[i * d + j]), c), w));
    }
    out16[j] = __float2half_rn(c);
}

// ---- medioid (lib.rs:65-68): argmax_i trunc(2^32 * <x_i, centroid16>) (f64 dot), the LAST maximum wins
__global__ void __launch_bounds__(256) k_medioid_scores(const __half *__restrict__ x, uint64_t n, uint32_t d, const __half *__restrict__ c16,
                                                        long long *__restrict__ scores) {
    extern __shared__ float scq[];
    for (uint32_t i = threadIdx.x; i < d; i += blockDim.x) scq[i] = __half2float(c16[i]);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < n; r += nw) {
        double acc = 0.0;
        for (uint32_t c = lane; c < d; c += 32) acc = fma((double)__half2float(x[r * d + c]), (double)scq[c], acc);
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            const double v = acc * 4294967296.0;
            scores[r] = (v != v) ? 0 : __double2ll_rz(v);
        }
    }
}
__global__ void k_argmax_last(const long long *scores, uint64_t n, unsigned long long *best) {
    // key = (score biased to unsigned) then index: max over both picks the highest score and, among equals, the last index
    __shared__ unsigned long long s_hi[256];
    __shared__ unsigned long long s_lo[256];
    unsigned long long hi = 0, lo = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = (unsigned long long)scores[i] ^ 0x8000000000000000ull;
        if (k > hi || (k == hi && i >= lo)) { hi = k; lo = i; }
    }
    s_hi[threadIdx.x] = hi; s_lo[threadIdx.x] = lo;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const unsigned long long h2 = s_hi[threadIdx.x + o], l2 = s_lo[threadIdx.x + o];
            if (h2 > s_hi[threadIdx.x] || (h2 == s_hi[threadIdx.x] && l2 > s_lo[threadIdx.x])) { s_hi[threadIdx.x] = h2; s_lo[threadIdx.x] = l2; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { best[2 * blockIdx.x] = s_hi[0]; best[2 * blockIdx.x + 1] = s_lo[0]; }
}

// ------------------------------------------------------------------ robust_stitch (lib.rs:326-374)
//
// The reference drops every base -> query edge, then walks the query nodes in a shuffled order and, for each base node that
// pointed at the query, appends the query's best out-neighbours (by fast_dot to the base node) to the base node's list.
// Only base nodes' lists are written and query nodes' lists are only read, so the work factors per BASE node: its incident
// queries, in shuffled order, are applied one after the other.  One warp per base node; `rank[q - qb]` is the position of
// query q in the shuffled order.  Same result as the sequential loop for the same order.
static constexpr int kStitchWarps = 4;

__global__ void __launch_bounds__(kStitchWarps * 32) k_robust_stitch(const __half *__restrict__ x, uint32_t d, uint32_t *adj, uint32_t *deg, uint32_t stride,
                                                                     uint32_t qb, const uint32_t *__restrict__ rank, uint32_t r, uint32_t max_add) {
    extern __shared__ __align__(16) uint8_t stitch_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per warp: qs[stride] (query targets), cs[stride] scores, cid[stride] ids
    uint8_t *base = stitch_smem + (size_t)warp * ((size_t)stride * 16);
    long long *cs = (long long *)base;
    uint32_t *cid = (uint32_t *)(cs + stride);
    uint32_t *qs = cid + stride;
    const uint32_t b = blockIdx.x * kStitchWarps + warp;
    if (b >= qb) return;
    uint32_t *out = adj + (size_t)b * stride;
    const uint32_t dg = min(deg[b], stride);
    // partition the list: non-query targets stay (order kept), query targets go to qs (order kept)   :339-347
    uint32_t nkeep = 0, nqs = 0;
    for (uint32_t b0 = 0; b0 < dg; b0 += 32) {
        const uint32_t i = b0 + lane;
        const bool have = i < dg;
        const uint32_t t = have ? out[i] : 0;
        const bool isq = have && t >= qb;
        const unsigned mq = __ballot_sync(0xffffffffu, isq), mk = __ballot_sync(0xffffffffu, have && !isq);
        __syncwarp();
        if (isq) qs[nqs + __popc(mq & ((1u << lane) - 1))] = t;
        if (have && !isq) out[nkeep + __popc(mk & ((1u << lane) - 1))] = t;   // nkeep + rank <= i: never overtakes an unread entry of a later chunk
        nqs += __popc(mq);
        nkeep += __popc(mk);
        __syncwarp();
    }
    uint32_t cur = nkeep;
    if (nqs == 0) { if (lane == 0) deg[b] = cur; return; }
    // queries in shuffled order (stable for duplicates): insertion sort by lane 0, lists are <= R long
    if (lane == 0) {
        for (uint32_t i = 1; i < nqs; i++) {
            const uint32_t q = qs[i], rq = rank[q - qb];
            uint32_t j = i;
            while (j > 0 && rank[qs[j - 1] - qb] > rq) { qs[j] = qs[j - 1]; j--; }
            qs[j] = q;
        }
    }
    __syncwarp();
    const __half *xb = x + (size_t)b * d;
    for (uint32_t qi = 0; qi < nqs && cur < r; qi++) {
        const uint32_t q = qs[qi];
        const uint32_t qd = min(deg[q], stride);
        const uint32_t *qn = adj + (size_t)q * stride;
        for (uint32_t j = 0; j < qd; j++) {                                           // :355-358
            const uint32_t nid = qn[j];
            const float f = fast_dot_warp(xb, x + (size_t)nid * d, d, lane);
            if (lane == 0) { cs[j] = fast_dot_fix(f); cid[j] = nid; }
        }
        __syncwarp();
        // candidates by (score desc, position asc); walk them in that order (:359-371)
        uint32_t added = 0;
        for (uint32_t taken = 0; taken < qd; taken++) {
            if (added >= max_add || cur >= r) break;
            // next best not yet taken: tombstone taken entries with position marker in cid (kept) and score = kDead - 1? use a separate pass
            long long best = 0; uint32_t bj = 0xFFFFFFFFu;
            for (uint32_t j = lane; j < qd; j += 32) {
                const long long sj = cs[j];
                if (sj == kDead) continue;
                if (bj == 0xFFFFFFFFu || sj > best) { best = sj; bj = j; }     // ascending j per lane: first max wins
            }
            for (int o = 16; o; o >>= 1) {
                const long long ob = __shfl_xor_sync(0xffffffffu, best, o);
                const uint32_t oj = __shfl_xor_sync(0xffffffffu, bj, o);
                if (oj != 0xFFFFFFFFu && (bj == 0xFFFFFFFFu || ob > best || (ob == best && oj < bj))) { best = ob; bj = oj; }
            }
            if (bj == 0xFFFFFFFFu) break;
            const uint32_t nid = cid[bj];
            __syncwarp();
            if (lane == 0) cs[bj] = kDead;
            bool present = false;
            for (uint32_t t = lane; t < cur; t += 32) present |= out[t] == nid;
            present = __any_sync(0xffffffffu, present);
            if (!present) {
                if (lane == 0) out[cur] = nid;
                cur++;
                added++;
            }
            __syncwarp();
        }
    }
    if (lane == 0) deg[b] = cur;
}

static PruneCfg to_prune_cfg(const mse_build_config &c) {
    PruneCfg p;
    p.r = (uint32_t)c.r; p.maxc = (uint32_t)c.maxc; p.alpha = c.alpha; p.query_alpha = c.query_alpha;
    p.query_breakpoint = c.query_breakpoint; p.saturate = c.saturate_graph;
    return p;
}

static int ensure_graph_storage(mse_index *ix, uint32_t stride) {
    if (ix->adj && ix->graph_stride == stride && ix->side_n == ix->n) return MSE_OK;
    if (ix->adj) cudaFree(ix->adj);
    if (ix->deg) cudaFree(ix->deg);
    ix->adj = nullptr; ix->deg = nullptr;
    MSE_CUDA(cudaMalloc(&ix->adj, std::max<size_t>(ix->n * stride * 4, 16)));
    MSE_CUDA(cudaMalloc(&ix->deg, std::max<size_t>(ix->n * 4, 16)));
    MSE_CUDA(cudaMemset(ix->adj, 0, std::max<size_t>(ix->n * stride * 4, 16)));
    MSE_CUDA(cudaMemset(ix->deg, 0, std::max<size_t>(ix->n * 4, 16)));
    ix->graph_stride = stride;
    ix->side_n = ix->n;
    return MSE_OK;
}

}  // namespace mse

using namespace mse;

// ================================================================== C ABI

static int check_cfg(const mse_build_config *c, const char *who) {
    MSE_REQUIRE(c != nullptr, MSE_ERR_INVALID, "%s: NULL config", who);
    MSE_REQUIRE(c->r >= 1 && c->r <= 256 && c->l >= 1 && c->l <= 4096 && c->maxc >= 1 && c->maxc <= (uint64_t)kKeep, MSE_ERR_UNSUPPORTED,
                "%s: r=%llu l=%llu maxc=%llu outside the supported range (r <= 256, l <= 4096, maxc <= %d)", who, (unsigned long long)c->r,
                (unsigned long long)c->l, (unsigned long long)c->maxc, kKeep);
    return MSE_OK;
}

// robust_prune (lib.rs:227-285) of point p over an explicit candidate list; out holds <= cfg->r ids
MSE_API int mse_robust_prune(mse_index *ix, uint32_t p, const uint32_t *cand_ids, const int64_t *cand_scores, uint32_t n_cand,
                             const mse_build_config *cfg, uint32_t *out, uint32_t *out_len) {
    MSE_REQUIRE(ix && out && out_len && (n_cand == 0 || (cand_ids && cand_scores)), MSE_ERR_INVALID, "robust_prune: NULL argument");
    MSE_CHECK(check_cfg(cfg, "robust_prune"));
    MSE_REQUIRE(ix->d % 64 == 0, MSE_ERR_UNSUPPORTED, "robust_prune: d %% 64 != 0");
    MSE_CHECK(use_device(ix->device));
    const size_t smem = prune_smem_bytes(ix->d, (uint32_t)cfg->r);
    const bool fixed = ix->d == 1152;
    MSE_CUDA(cudaFuncSetAttribute(fixed ? k_robust_prune_one<18> : k_robust_prune_one<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DevBuf bi, bs, bo;
    int rc = MSE_OK;
    do {
        if ((rc = bi.ensure(std::max<size_t>(n_cand, 1) * 4)) || (rc = bs.ensure(std::max<size_t>(n_cand, 1) * 8)) || (rc = bo.ensure((cfg->r + 1) * 4))) break;
        cudaMemcpy(bi.p, cand_ids, (size_t)n_cand * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(bs.p, cand_scores, (size_t)n_cand * 8, cudaMemcpyHostToDevice);
        (fixed ? k_robust_prune_one<18> : k_robust_prune_one<0>)<<<1, kPruneThreads, smem>>>(ix->x, ix->d, p, bi.as<uint32_t>(), bs.as<long long>(), n_cand,
                                                                                            to_prune_cfg(*cfg), bo.as<uint32_t>(), bo.as<uint32_t>() + cfg->r);
        count_launch();
        std::vector<uint32_t> h(cfg->r + 1);
        cudaError_t e = cudaMemcpy(h.data(), bo.p, (cfg->r + 1) * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("robust_prune: %s", cudaGetErrorString(e)); rc = MSE_ERR_CUDA; break; }
        *out_len = h[cfg->r];
        memcpy(out, h.data(), (size_t)*out_len * 4);
    } while (0);
    bi.release(); bs.release(); bo.release();
    return rc;
}

// IndexGraph::empty + random_fill_graph (lib.rs:22-31,376-387)
MSE_API int mse_index_random_fill_graph(mse_index *ix, uint32_t r, uint64_t seed) {
    MSE_REQUIRE(ix != nullptr && r >= 1 && r <= 256, MSE_ERR_INVALID, "random_fill_graph: bad argument");
    MSE_CHECK(use_device(ix->device));
    MSE_CHECK(ensure_graph_storage(ix, ix->graph_stride && ix->graph_stride >= r ? ix->graph_stride : r));
    if (ix->n == 0) return MSE_OK;
    k_random_fill<<<(uint32_t)((ix->n + 127) / 128), 128>>>(ix->adj, ix->deg, ix->n, ix->graph_stride, r, seed);
    count_launch();
    MSE_CUDA(cudaDeviceSynchronize());
    return MSE_OK;
}

// medioid (lib.rs:65-68)
MSE_API int mse_index_medioid(mse_index *ix, uint32_t *out) {
    MSE_REQUIRE(ix != nullptr && out != nullptr && ix->n > 0, MSE_ERR_INVALID, "medioid: bad argument or empty index");
    MSE_CHECK(use_device(ix->device));
    DevBuf bc, bs, bb;
    int rc = MSE_OK;
    do {
        const uint32_t blocks = 64;
        if ((rc = bc.ensure(ix->d * 2)) || (rc = bs.ensure(ix->n * 8)) || (rc = bb.ensure(blocks * 16))) break;
        k_centroid<<<(ix->d + 127) / 128, 128>>>(ix->x, ix->n, ix->d, bc.as<__half>());
        count_launch();
        uint32_t sb = (uint32_t)std::min<uint64_t>((ix->n + 7) / 8, (uint64_t)sm_count(ix->device) * 8);
        k_medioid_scores<<<sb, 256, ix->d * 4>>>(ix->x, ix->n, ix->d, bc.as<__half>(), bs.as<long long>());
        count_launch();
        k_argmax_last<<<blocks, 256>>>(bs.as<long long>(), ix->n, bb.as<unsigned long long>());
        count_launch();
        std::vector<unsigned long long> h(blocks * 2);
        cudaError_t e = cudaMemcpy(h.data(), bb.p, blocks * 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("medioid: %s", cudaGetErrorString(e)); rc = MSE_ERR_CUDA; break; }
        unsigned long long hi = 0, lo = 0;
        for (uint32_t b = 0; b < blocks; b++)
            if (h[2 * b] > hi || (h[2 * b] == hi && h[2 * b + 1] >= lo)) { hi = h[2 * b]; lo = h[2 * b + 1]; }
        *out = (uint32_t)lo;
    } while (0);
    bc.release(); bs.release(); bb.release();
    return rc;
}

// build_graph (lib.rs:287-324), batch-synchronous.  The graph must hold a starting graph (mse_index_random_fill_graph or
// mse_index_set_graph); stats (optional, 6 values): batches, searches, back-edge merges, total distance evaluations,
// searches whose visited list was cut at its capacity, searches aborted on a visited-set overflow (the build then fails).
MSE_API int mse_index_build_vamana(mse_index *ix, uint32_t medioid, const mse_build_config *cfg, uint64_t seed, uint32_t max_batch, uint64_t *stats) {
    MSE_REQUIRE(ix != nullptr, MSE_ERR_INVALID, "build_vamana: NULL handle");
    MSE_CHECK(check_cfg(cfg, "build_vamana"));
    MSE_REQUIRE(ix->adj && ix->deg && ix->graph_stride >= cfg->r, MSE_ERR_STATE,
                "build_vamana: start from mse_index_random_fill_graph / mse_index_set_graph with stride >= r (generate_index_shard.rs:102-107)");
    MSE_REQUIRE(ix->d % 64 == 0 && medioid < ix->n, MSE_ERR_INVALID, "build_vamana: bad medioid or d %% 64 != 0");
    MSE_CHECK(use_device(ix->device));
    const uint64_t n = ix->n;
    if (max_batch == 0) max_batch = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(n / 50, 64), 16384);
    const uint32_t L = (uint32_t)cfg->l, stride = ix->graph_stride;
    // visited_list (lib.rs:205) is unbounded in the reference; a search evaluates ~25 x L rows at R = 64, so 48 x L (>= 8192) entries
    // hold it with a wide margin, and a search that still exceeds them is counted in stats[4] (its list keeps the first vl_cap entries)
    const uint32_t vl_cap = std::max<uint32_t>(8192, 48 * L);
    // sigma: host-side shuffle (lib.rs:291-292; the reference's fastrand stream need not be matched)
    std::vector<uint32_t> sigma(n);
    for (uint64_t i = 0; i < n; i++) sigma[i] = (uint32_t)i;
    uint64_t s = seed;
    auto next = [&]() { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); };
    for (uint64_t i = n; i > 1; i--) std::swap(sigma[i - 1], sigma[next() % i]);

    const uint32_t grid_s = greedy_grid(ix, max_batch);
    const uint32_t hcap = greedy_hash_capacity(L, stride);
    const size_t psmem = prune_smem_bytes(ix->d, (uint32_t)cfg->r);
    const bool fixed = ix->d == 1152;
    auto prune_batch = fixed ? k_prune_batch<18> : k_prune_batch<0>;
    auto merge_backedges = fixed ? k_merge_backedges<18> : k_merge_backedges<0>;
    MSE_CUDA(cudaFuncSetAttribute(prune_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
    MSE_CUDA(cudaFuncSetAttribute(merge_backedges, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
    DevBuf b_sigma, b_ids, b_sc, b_len, b_dist, b_st, b_h, b_vi, b_vs, b_vl, b_na, b_nd, b_in, b_ic, b_t, b_nt;
    int rc = MSE_OK;
    uint64_t st_batches = 0, st_search = 0, st_merge = 0, st_dist = 0, st_trunc = 0, st_ovf = 0;
    do {
        if ((rc = b_sigma.ensure(n * 4)) || (rc = b_ids.ensure((size_t)max_batch * L * 4)) || (rc = b_sc.ensure((size_t)max_batch * L * 8)) ||
            (rc = b_len.ensure((size_t)max_batch * 4)) || (rc = b_dist.ensure((size_t)max_batch * 8)) || (rc = b_st.ensure((size_t)max_batch * 4)) ||
            (rc = b_h.ensure((size_t)grid_s * hcap * 4)) || (rc = b_vi.ensure((size_t)max_batch * vl_cap * 4)) ||
            (rc = b_vs.ensure((size_t)max_batch * vl_cap * 8)) || (rc = b_vl.ensure((size_t)max_batch * 4)) ||
            (rc = b_na.ensure((size_t)max_batch * stride * 4)) || (rc = b_nd.ensure((size_t)max_batch * 4)) || (rc = b_in.ensure(n * kMaxIn * 4)) ||
            (rc = b_ic.ensure(n * 4)) || (rc = b_t.ensure(n * 4)) || (rc = b_nt.ensure(4)))
            break;
        cudaMemcpy(b_sigma.p, sigma.data(), n * 4, cudaMemcpyHostToDevice);
        cudaMemset(b_ic.p, 0, n * 4);
        const PruneCfg pc = to_prune_cfg(*cfg);
        const uint32_t sms = (uint32_t)sm_count(ix->device);
        uint64_t done = 0;
        uint32_t bs = 1;
        std::vector<unsigned long long> hdist(max_batch);
        std::vector<uint32_t> hst(max_batch), hvl(max_batch);
        while (done < n && rc == MSE_OK) {
            const uint32_t nb = (uint32_t)std::min<uint64_t>(bs, n - done);
            const uint32_t *pts = b_sigma.as<uint32_t>() + done;
            GreedyOut o{b_ids.as<uint32_t>(), b_sc.as<long long>(), b_len.as<uint32_t>(), b_dist.as<unsigned long long>(), b_vi.as<uint32_t>(),
                        b_vs.as<long long>(), b_vl.as<uint32_t>(), vl_cap, b_st.as<uint32_t>()};
            // greedy_search(medioid -> point) (lib.rs:299); query points search base vectors only (:298)
            if ((rc = greedy_search_launch(ix, ix->x, pts, nb, nullptr, medioid, L, cfg->query_breakpoint, cfg->query_breakpoint, b_h.as<uint32_t>(), hcap,
                                           greedy_grid(ix, nb), o, nullptr))) break;
            prune_batch<<<std::min(nb, sms * 4), kPruneThreads, psmem>>>(ix->x, ix->d, pts, nb, b_vi.as<uint32_t>(), b_vs.as<long long>(), b_vl.as<uint32_t>(),
                                                                          vl_cap, ix->adj, ix->deg, stride, pc, b_na.as<uint32_t>(), b_nd.as<uint32_t>());
            count_launch();
            cudaMemsetAsync(b_nt.p, 0, 4);
            k_apply_and_backedges<<<nb, 64>>>(pts, nb, b_na.as<uint32_t>(), b_nd.as<uint32_t>(), stride, ix->adj, ix->deg, b_in.as<uint32_t>(),
                                             b_ic.as<uint32_t>(), b_t.as<uint32_t>(), b_nt.as<uint32_t>());
            count_launch();
            uint32_t nt = 0;
            cudaError_t e = cudaMemcpy(&nt, b_nt.p, 4, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { set_error("build_vamana: %s", cudaGetErrorString(e)); rc = MSE_ERR_CUDA; break; }
            if (nt) {
                merge_backedges<<<std::min(nt, sms * 4), kPruneThreads, psmem>>>(ix->x, ix->d, b_t.as<uint32_t>(), nt, b_in.as<uint32_t>(), b_ic.as<uint32_t>(),
                                                                                  ix->adj, ix->deg, stride, pc);
                count_launch();
            }
            cudaMemcpy(hdist.data(), b_dist.p, (size_t)nb * 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(hst.data(), b_st.p, (size_t)nb * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hvl.data(), b_vl.p, (size_t)nb * 4, cudaMemcpyDeviceToHost);
            for (uint32_t i = 0; i < nb; i++) {
                st_dist += hdist[i];
                st_trunc += hvl[i] > vl_cap;
                st_ovf += hst[i] != 0;
            }
            if (st_ovf) {   // the search stopped early: the point would be pruned over a partial candidate set
                set_error("build_vamana: the visited-set table overflowed in %llu searches (l=%u, stride=%u)", (unsigned long long)st_ovf, L, stride);
                rc = MSE_ERR_UNSUPPORTED;
                break;
            }
            st_batches++; st_search += nb; st_merge += nt;
            done += nb;
            bs = std::min<uint32_t>(bs * 2, max_batch);
        }
        if (rc == MSE_OK) {
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { set_error("build_vamana: %s", cudaGetErrorString(e)); rc = MSE_ERR_CUDA; }
        }
    } while (0);
    b_sigma.release(); b_ids.release(); b_sc.release(); b_len.release(); b_dist.release(); b_st.release(); b_h.release(); b_vi.release();
    b_vs.release(); b_vl.release(); b_na.release(); b_nd.release(); b_in.release(); b_ic.release(); b_t.release(); b_nt.release();
    if (stats) { stats[0] = st_batches; stats[1] = st_search; stats[2] = st_merge; stats[3] = st_dist; stats[4] = st_trunc; stats[5] = st_ovf; }
    return rc;
}

// robust_stitch (lib.rs:326-374).  query_order: the n - query_breakpoint query node ids in the order the reference's shuffled
// loop would visit them (NULL: a seeded shuffle is drawn here).  Nodes >= cfg->query_breakpoint are query nodes.
MSE_API int mse_index_robust_stitch(mse_index *ix, const mse_build_config *cfg, const uint32_t *query_order, uint64_t seed) {
    MSE_REQUIRE(ix != nullptr && cfg != nullptr, MSE_ERR_INVALID, "robust_stitch: NULL argument");
    MSE_REQUIRE(ix->adj && ix->deg, MSE_ERR_STATE, "robust_stitch: the index has no graph");
    MSE_REQUIRE(ix->d % 64 == 0, MSE_ERR_UNSUPPORTED, "robust_stitch: fast_dot needs d %% 64 == 0");
    MSE_REQUIRE(cfg->r >= 1 && cfg->r <= ix->graph_stride, MSE_ERR_INVALID, "robust_stitch: r=%llu exceeds the adjacency stride %u",
                (unsigned long long)cfg->r, ix->graph_stride);
    const uint64_t n = ix->n;
    const uint32_t qb = cfg->query_breakpoint;
    if (qb >= n) return MSE_OK;                                  // no query nodes: nothing to stitch (generate_index_shard.rs:129)
    MSE_CHECK(use_device(ix->device));
    const uint32_t nq = (uint32_t)(n - qb);
    std::vector<uint32_t> order(nq), rank(nq);
    if (query_order) {
        std::vector<uint8_t> seen(nq, 0);
        for (uint32_t i = 0; i < nq; i++) {
            MSE_REQUIRE(query_order[i] >= qb && query_order[i] < n && !seen[query_order[i] - qb], MSE_ERR_INVALID,
                        "robust_stitch: query_order is not a permutation of [query_breakpoint, n)");
            seen[query_order[i] - qb] = 1;
            order[i] = query_order[i];
        }
    } else {
        for (uint32_t i = 0; i < nq; i++) order[i] = qb + i;
        uint64_t s = seed ? seed : 0x9e3779b97f4a7c15ull;
        auto next = [&]() { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); };
        for (uint32_t i = nq; i > 1; i--) std::swap(order[i - 1], order[next() % i]);
    }
    for (uint32_t i = 0; i < nq; i++) rank[order[i] - qb] = i;
    DevBuf b_rank;
    int rc = MSE_OK;
    do {
        if ((rc = b_rank.ensure((size_t)nq * 4))) break;
        cudaMemcpy(b_rank.p, rank.data(), (size_t)nq * 4, cudaMemcpyHostToDevice);
        if (qb > 0) {
            const size_t smem = (size_t)kStitchWarps * ix->graph_stride * 16;
            k_robust_stitch<<<(qb + kStitchWarps - 1) / kStitchWarps, kStitchWarps * 32, smem>>>(ix->x, ix->d, ix->adj, ix->deg, ix->graph_stride, qb,
                                                                                             b_rank.as<uint32_t>(), (uint32_t)cfg->r,
                                                                                             (uint32_t)cfg->max_add_per_stitch_iter);
            count_launch();
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { set_error("robust_stitch: %s", cudaGetErrorString(e)); rc = MSE_ERR_CUDA; }
    } while (0);
    b_rank.release();
    return rc;
}
