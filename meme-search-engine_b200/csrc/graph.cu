// Vamana graph search on the GPU, batched over queries, bit-exact against the reference's sequential semantics.
//
//   k_greedy_search   diskann/src/lib.rs:183-211  (in-memory greedy_search, beam 1, exact fast_dot scores)
//                     + NeighbourBuffer diskann/src/lib.rs:73-155
//   k_beam_search     src/query_disk_index.rs:83-97,135-212 (beam-W search over the packed index: PQ ADC for frontier
//                     candidates, exact fp16 dot + descriptor bias for expanded nodes; results = expanded nodes)
//   k_brute_force_i64 src/query_disk_index.rs:262-273 (exact i64 scores of every node)
//
// One persistent CTA works on one query at a time.  The vectors, adjacency, PQ codes and descriptors live in HBM (the
// reference reads 4 KiB node records from NVMe through io_uring: query_disk_index.rs:73-81); every hop gathers <= R rows of
// 2304 B (exact) or 64 B codes (ADC).  All eight warps score candidates (one warp per row, fast_dot.cuh); warp 0 then
// replays the reference's NeighbourBuffer inserts in the reference's order, so ids, i64 scores and distance counters match
// the CPU oracle bit for bit.  The reference's HashSet<u32> is an open-addressing table in global memory (exact membership).
#include "graph.cuh"
#include "fastdot.cuh"
#include <algorithm>
#include <math.h>
#include <stdlib.h>

namespace mse {

static constexpr int kGsThreads = 256;
static constexpr int kGsWarps = kGsThreads / 32;
static constexpr uint32_t kEmpty = 0xFFFFFFFFu;
static constexpr int kMaxDeg = 256;   // largest adjacency list handled per expansion (merged graphs reach 2R = 128)

struct NbView {
    uint32_t *ids;
    long long *scores;
    uint8_t *vis;
    int len, cap, nu;  // nu: next_unvisited or -1
};

// ---- NeighbourBuffer (lib.rs:73-155), executed by ONE warp; state scalars are kept uniform across the warp

__device__ __forceinline__ void nb_insert(NbView &b, uint32_t id, long long score, int lane) {
    if (b.cap == 0) return;
    if (b.len == b.cap && b.scores[b.len - 1] > score) return;                     // :118
    int g = 0, e = 0, c0 = 0;                                                      // binary_search_by on the descending list:
    if (b.len > 128 && b.len <= 1024) {
        // long lists: lane k looks at the head of chunk k first.  The chunks whose head is above `score` form a prefix and all of them
        // but the last lie entirely above it (every entry of chunk k is >= the head of chunk k+1), so the scan can start at the last
        const int nch = (b.len + 31) >> 5;
        const bool hin = lane < nch;
        const long long h = hin ? b.scores[lane << 5] : 0;
        const int G = __popc(__ballot_sync(0xffffffffu, hin && h > score));
        c0 = G > 0 ? G - 1 : 0;
        g = c0 << 5;
    }
    for (int i0 = c0 << 5; i0 < b.len; i0 += 32) {                                 // g = #entries > score, e = #entries == score
        const int i = i0 + lane;
        const bool in = i < b.len;
        const long long s = in ? b.scores[i] : 0;
        const unsigned mg = __ballot_sync(0xffffffffu, in && s > score), me = __ballot_sync(0xffffffffu, in && s == score);
        g += __popc(mg);
        e += __popc(me);
        if ((mg | me) != 0xffffffffu) break;                                       // sorted: everything further down is smaller
    }
    const int loc = e > 0 ? g + e - 1 : g;                                         // Ok(last equal) / Err(insertion point)
    if (loc < b.len && b.ids[loc] == id) return;                                   // :127
    // insert at loc, truncate to cap (:132-137): shift [loc, len) right by one, chunks from the top down
    for (int c = (b.len - 1) >> 5; b.len > 0 && c >= (loc >> 5); c--) {
        const int i = (c << 5) + lane;
        const bool mv = i >= loc && i < b.len && i + 1 < b.cap;
        uint32_t ti = 0; long long ts = 0; uint8_t tv = 0;
        if (mv) { ti = b.ids[i]; ts = b.scores[i]; tv = b.vis[i]; }
        __syncwarp();
        if (mv) { b.ids[i + 1] = ti; b.scores[i + 1] = ts; b.vis[i + 1] = tv; }
        __syncwarp();
    }
    if (lane == 0) { b.ids[loc] = id; b.scores[loc] = score; b.vis[loc] = 0; }
    __syncwarp();
    b.len = min(b.len + 1, b.cap);
    b.nu = b.nu < 0 ? loc : min(loc, b.nu);                                        // :139-146
}

// lib.rs:93-107; returns the id or kEmpty when nothing is unvisited
__device__ __forceinline__ uint32_t nb_next_unvisited(NbView &b, int lane) {
    if (b.nu < 0) return kEmpty;
    const int old = b.nu;
    if (lane == 0) b.vis[old] = 1;
    __syncwarp();
    int cur = old;
    for (;;) {  // first index >= old that is unvisited
        const int i = cur + lane;
        const bool unv = i < b.len && !b.vis[i];
        const unsigned m = __ballot_sync(0xffffffffu, unv);
        if (m) { cur += __ffs(m) - 1; break; }
        cur += 32;
        if (cur >= b.len) { cur = b.len; break; }
    }
    b.nu = cur >= b.len ? -1 : cur;
    return b.ids[old];
}

__device__ __forceinline__ long long shfl_ll_any(long long v, int src) {
    const unsigned full = 0xffffffffu;
    const int lo = __shfl_sync(full, (int)(unsigned long long)v, src), hi = __shfl_sync(full, (int)((unsigned long long)v >> 32), src);
    return (long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo);
}

// ---- NeighbourBuffer: up to 32 inserts of one pass as ONE merge (long lists: k_beam_search_wq with 256..608 entries listed; measured at
// 12.5 M rows: L = 512 14.9 -> 11.9 ms, L = 256 7.33 -> 7.03 ms, but L = 128 3.80 -> 4.38 ms, hence the threshold)
//
// A sequence of NeighbourBuffer::insert calls (lib.rs:109-147) shifts the tail of the list once per call -- at L = 512 that was two
// thirds of the beam kernel's instructions.  When no two scores involved are equal, the sequence has a closed form: the result is the
// top-`cap` of (list + candidates) in descending score order, and next_unvisited is the first unvisited entry (every insert leaves it
// there: min(loc, nu), :139-146).  Lane i holds candidate i of the pass (mask `m`: the candidates lib.rs:118 lets through against the
// tail at pass start -- the tail only grows, so later calls would reject at least those).  Each candidate finds its rank in the list
// (binary search) and among the candidates; a bit mask over the new list marks where candidates go, and the new list is written from
// the top chunk down, every destination slot pulling either its candidate or the old entry (slot - number of candidates before it).
// Sources of a destination chunk lie in that chunk or below and are not needed again, so the move is in place.
// Equal scores (a candidate against the list or another candidate -- also how an id that is already listed shows up, :127) have
// position rules that depend on the call order (:120-127): the function then changes nothing and returns false, and the caller
// replays the pass with nb_insert.  scratch: 20 words of mask (lists of up to 608 entries) + 32 x (u32, i64) candidates, per warp.
static constexpr uint32_t kNbBatchScratch = 80 + 32 * 4 + 32 * 8;   // bytes, 8-aligned layout: mask[20] | ids[32] | scores[32]

__device__ __forceinline__ bool nb_insert_batch(NbView &b, uint32_t cid, long long cs, unsigned m, int lane, uint8_t *scratch) {
    const unsigned full = 0xffffffffu;
    uint32_t *mask = (uint32_t *)scratch;
    uint32_t *c_ids = mask + 20;
    long long *c_sc = (long long *)(c_ids + 32);
    const bool mine = (m >> lane) & 1u;
    // rank in the list: r = #entries with score > cs; a tie with the list aborts
    int lo = 0, hi = b.len;
    while (__any_sync(full, mine && lo < hi)) {
        if (mine && lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (b.scores[mid] > cs) lo = mid + 1; else hi = mid;
        }
    }
    const int r = lo;
    bool tie = mine && r < b.len && b.scores[r] == cs;
    // rank among the candidates; equal candidate scores abort
    int k = 0;
    for (unsigned mm = m; mm;) {
        const int j = __ffs(mm) - 1;
        mm &= mm - 1;
        const long long sj = shfl_ll_any(cs, j);
        if (mine && j != lane) { k += sj > cs; tie |= sj == cs; }
    }
    if (__any_sync(full, tie)) return false;
    const int pos = r + k;                                    // final index of this candidate
    const int n_cand = __popc(m);
    const int new_len = min(b.cap, b.len + n_cand);
    const bool kept = mine && pos < new_len;
    // mask over the new list + candidates by rank
    if (lane < 20) mask[lane] = 0;
    __syncwarp();
    if (kept) {
        atomicOr(&mask[pos >> 5], 1u << (pos & 31));
        c_ids[k] = cid;
        c_sc[k] = cs;
    }
    __syncwarp();
    const unsigned kept_m = __ballot_sync(full, kept);
    if (kept_m == 0) return true;                             // everything fell off the end
    int min_pos = kept ? pos : 0x7fffffff;
#pragma unroll
    for (int o = 16; o; o >>= 1) min_pos = min(min_pos, __shfl_xor_sync(full, min_pos, o));
    // words of the mask: lane w holds word w and the number of candidates before it
    const int n_words = (new_len + 31) >> 5;
    const uint32_t my_word = lane < n_words ? mask[lane] : 0u;
    int before = __popc(my_word);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(full, before, o);
        if (lane >= o) before += t;
    }
    before -= __popc(my_word);                                // exclusive prefix
    // next_unvisited: the first unvisited entry of the new list
    int new_nu = min_pos;
    if (b.nu >= 0) {
        const long long snu = b.scores[b.nu];
        const int moved = b.nu + __popc(__ballot_sync(full, mine && cs > snu));
        if (moved < new_len) new_nu = min(new_nu, moved);
    }
    // new list, top chunk first, down to the chunk of the first candidate (everything below stays where it is)
    for (int c = n_words - 1; c >= (min_pos >> 5); c--) {
        const uint32_t w = __shfl_sync(full, my_word, c);
        const int bef = __shfl_sync(full, before, c);
        const int q = (c << 5) + lane;
        const int below = bef + __popc(w & ((1u << lane) - 1));
        const bool is_c = (w >> lane) & 1u;
        uint32_t ti = 0;
        long long ts = 0;
        uint8_t tv = 0;
        if (q < new_len) {
            if (is_c) { ti = c_ids[below]; ts = c_sc[below]; }
            else { const int src = q - below; ti = b.ids[src]; ts = b.scores[src]; tv = b.vis[src]; }
        }
        __syncwarp();
        if (q < new_len && q >= min_pos) { b.ids[q] = ti; b.scores[q] = ts; b.vis[q] = tv; }
        __syncwarp();
    }
    b.len = new_len;
    b.nu = new_nu;
    return true;
}

// ---- HashSet<u32>::insert: true when newly inserted.  Open addressing, linear probing, capacity a power of two.
__device__ __forceinline__ bool hs_insert(uint32_t *tab, uint32_t mask, uint32_t key, uint32_t *fill) {
    uint32_t h = (key * 2654435761u) & mask;
    for (uint32_t probes = 0; probes <= mask; probes++) {
        const uint32_t old = atomicCAS(&tab[h], kEmpty, key);
        if (old == kEmpty) { atomicAdd(fill, 1u); return true; }
        if (old == key) return false;
        h = (h + 1) & mask;
    }
    return false;  // table full: caller checks *fill against the capacity
}

// shared-memory layout helper
struct GsSmem {
    long long *nb_scores;
    uint32_t *nb_ids;
    uint8_t *nb_vis;
    float *q;               // query as fp32 (exact fp16 values)
    uint32_t *pre;          // candidate ids of the current expansion
    long long *pre_scores;
    uint32_t *raw;          // raw adjacency list of the expanded node
    int *ctl;               // [0] n_pre  [1] pt  [2] scratch
};

__device__ __forceinline__ GsSmem carve(uint8_t *base, uint32_t L, uint32_t d, uint32_t S) {   // S = adjacency stride (max degree)
    GsSmem s;
    size_t o = 0;
    s.nb_scores = (long long *)(base + o); o += (size_t)(L + 1) * 8;
    s.pre_scores = (long long *)(base + o); o += (size_t)S * 8;
    s.nb_ids = (uint32_t *)(base + o); o += (size_t)(L + 1) * 4;
    s.pre = (uint32_t *)(base + o); o += (size_t)S * 4;
    s.raw = (uint32_t *)(base + o); o += (size_t)S * 4;
    s.q = (float *)(base + o); o += (size_t)d * 4;
    s.ctl = (int *)(base + o); o += 16 * 4;
    s.nb_vis = (uint8_t *)(base + o);
    return s;
}
__host__ __device__ static size_t gs_smem_bytes(uint32_t L, uint32_t d, uint32_t S) {
    return (size_t)(L + 1) * 8 + (size_t)S * 8 + (size_t)(L + 1) * 4 + (size_t)S * 4 + (size_t)S * 4 + (size_t)d * 4 + 64 + (L + 1) + 64;
}

// exact fast_dot of the shared-memory query against one row; whole warp, bit-identical to vector.rs:192-306
__device__ __forceinline__ long long score_row(const float *q, const __half *row, uint32_t d, int lane) {
    float p = 0.f;
    for (uint32_t c = lane; c < d; c += 32) p = fmaf(q[c], __half2float(row[c]), p);
    return fast_dot_fix(fast_dot_reduce(p));
}

// expanded node's out-neighbours -> s.pre (first occurrence of every not-yet-seen id, in list order), returns the count.
// filter_from: ids >= filter_from are dropped AFTER being marked seen (lib.rs:196-199 base_vectors_only)
__device__ __forceinline__ int collect_new_neighbours(const GraphArgs &g, GsSmem &s, uint32_t pt, uint32_t *htab, uint32_t hmask,
                                                      uint32_t *hfill, uint32_t filter_from, int n_pre0) {
    const uint32_t dg = min(g.deg[pt], g.stride);
    const uint32_t *nbrs = g.adj + (size_t)pt * g.stride;
    for (uint32_t i = threadIdx.x; i < dg; i += blockDim.x) s.raw[i] = nbrs[i];
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int n_pre = n_pre0;
        for (uint32_t base = 0; base < dg; base += 32) {
            const uint32_t i = base + lane;
            bool fresh = false;
            uint32_t id = 0;
            if (i < dg) {
                id = s.raw[i];
                bool dup = false;
                for (uint32_t j = 0; j < i; j++) dup |= s.raw[j] == id;     // an earlier copy in this list wins
                if (!dup) fresh = hs_insert(htab, hmask, id, hfill) && id < filter_from;
            }
            const unsigned m = __ballot_sync(0xffffffffu, fresh);
            if (fresh) s.pre[n_pre + __popc(m & ((1u << lane) - 1))] = id;
            n_pre += __popc(m);
        }
        if (lane == 0) s.ctl[0] = n_pre;
    }
    __syncthreads();
    return s.ctl[0];
}

// ------------------------------------------------------------------ in-memory greedy search (lib.rs:183-211)

__global__ void __launch_bounds__(kGsThreads) k_greedy_search(GraphArgs g, const __half *__restrict__ queries, const uint32_t *__restrict__ q_rows, uint32_t nq,
                                                              const uint32_t *__restrict__ starts, uint32_t start_all, uint32_t L,
                                                              uint32_t filter_from_all, uint32_t filter_self_from, uint32_t *htabs, uint32_t hcap,
                                                              GreedyOut out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    GsSmem s = carve(smem_raw, L, g.d, g.stride);
    uint32_t *htab = htabs + (size_t)blockIdx.x * hcap;
    const uint32_t hmask = hcap - 1;
    __shared__ uint32_t hfill;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (uint32_t qi = blockIdx.x; qi < nq; qi += gridDim.x) {
        if ((starts ? starts[qi] : start_all) >= g.n) {                              // not a row of this shard: empty result, status bit 2
            for (uint32_t i = threadIdx.x; i < L; i += blockDim.x) { out.ids[(size_t)qi * L + i] = kEmpty; out.scores[(size_t)qi * L + i] = 0; }
            if (threadIdx.x == 0) { out.len[qi] = 0; out.distances[qi] = 0; if (out.vl_len) out.vl_len[qi] = 0; out.status[qi] = 4u; }
            continue;
        }
        for (uint32_t i = threadIdx.x; i < hcap; i += blockDim.x) htab[i] = kEmpty;
        const size_t qrow = q_rows ? q_rows[qi] : qi;
        // base_vectors_only (lib.rs:196-199) applies to the searches of query nodes only (build_graph :297-298)
        const uint32_t filter_from = (q_rows && qrow < filter_self_from) ? 0xFFFFFFFFu : filter_from_all;
        for (uint32_t i = threadIdx.x; i < g.d; i += blockDim.x) s.q[i] = __half2float(queries[qrow * g.d + i]);
        if (threadIdx.x == 0) hfill = 0;
        __syncthreads();
        NbView nb{s.nb_ids, s.nb_scores, s.nb_vis, 0, (int)L, -1};
        const uint32_t start = starts ? starts[qi] : start_all;
        unsigned long long distances = 0;
        uint32_t n_eval = 0;
        if (warp == 0) {
            const long long sc = score_row(s.q, g.x + (size_t)start * g.d, g.d, lane);
            nb_insert(nb, start, sc, lane);                                          // :188
            if (lane == 0) hs_insert(htab, hmask, start, &hfill);                    // :189
        }
        __syncthreads();
        for (;;) {
            if (warp == 0) {
                const uint32_t pt = nb_next_unvisited(nb, lane);
                if (lane == 0) s.ctl[1] = (int)pt;
            }
            __syncthreads();
            const uint32_t pt = (uint32_t)s.ctl[1];
            if (pt == kEmpty || hfill * 4 > hcap * 3) break;
            const int n_pre = collect_new_neighbours(g, s, pt, htab, hmask, &hfill, filter_from, 0);
            for (int i = warp; i < n_pre; i += kGsWarps) {                           // :201-204, one warp per row
                const long long sc = score_row(s.q, g.x + (size_t)s.pre[i] * g.d, g.d, lane);
                if (lane == 0) s.pre_scores[i] = sc;
            }
            __syncthreads();
            if (warp == 0) {
                // rows that cannot enter a full buffer are rejected exactly as :118 would (the tail score only grows)
                const bool full0 = nb.len == nb.cap;
                const long long last0 = full0 ? nb.scores[nb.len - 1] : 0;
                for (int i = 0; i < n_pre; i++) {
                    const long long sc = s.pre_scores[i];
                    if (!(full0 && last0 > sc)) nb_insert(nb, s.pre[i], sc, lane);
                }
                if (out.vl_ids) {
                    for (int i = lane; i < n_pre; i += 32) {
                        const uint32_t slot = n_eval + i;
                        if (slot < out.vl_cap) {
                            out.vl_ids[(size_t)qi * out.vl_cap + slot] = s.pre[i];
                            out.vl_scores[(size_t)qi * out.vl_cap + slot] = s.pre_scores[i];
                        }
                    }
                }
                n_eval += n_pre;
                distances += n_pre;
            }
            __syncthreads();
        }
        if (warp == 0) {
            for (uint32_t i = lane; i < L; i += 32) {
                out.ids[(size_t)qi * L + i] = (int)i < nb.len ? nb.ids[i] : kEmpty;
                out.scores[(size_t)qi * L + i] = (int)i < nb.len ? nb.scores[i] : 0;
            }
            if (lane == 0) {
                out.len[qi] = nb.len;
                out.distances[qi] = distances;
                if (out.vl_len) out.vl_len[qi] = n_eval;
                out.status[qi] = (hfill * 4 > hcap * 3) ? 1u : 0u;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ greedy search, one WARP per query (batch mode)
//
// Same algorithm, same order of every insert, same outputs as k_greedy_search; the difference is the schedule.  A hop of
// one query is a chain (pop -> adjacency -> visited set -> <= R row gathers -> inserts) with no parallelism across hops, so
// a large batch is served best by many independent chains per SM: each warp owns one query and never waits at a CTA
// barrier.  64 registers per thread -> 32 warps per SM, each with up to 36 128-byte row requests in flight (two rows per
// pass, fastdot.cuh wq_score2), which is what hides the HBM gather latency: 0.49 of the HBM copy peak at 4096 queries.

static constexpr int kWqWarps = 4;   // warps (queries in flight) per CTA
// greedy search: rows started towards L2 ahead of the scoring pass (0 = off; MSE_GREEDY_PF overrides, a tuning aid).  Measured at the
// C4 shard (12.5 M rows, 4096 queries, L = 64; profiles/r03b_greedy_prefetch_sweep.json): 0 rows 5.30 ms, 6 rows 4.36 ms, 10 rows 4.95 ms,
// 16 rows 5.22 ms -- deeper prefetch evicts rows from L2 before their pass reads them (4096 queries x 6 rows x 2304 B = 57 MB in flight).
// Also measured without effect: prefetching the next hop's adjacency list during the inserts (4.37 ms), halving the visited-set tables (4.42 ms).
static int greedy_pf_rows() {
    const char *e = getenv("MSE_GREEDY_PF");
    const int v = e ? atoi(e) : 6;
    return v < 0 ? 0 : (v > 64 ? 64 : v);
}

__host__ __device__ static size_t wq_warp_bytes(uint32_t L, uint32_t stride, uint32_t d) {
    size_t o = (size_t)(L + 1) * 8 + (size_t)stride * 8 + (size_t)(L + 1) * 4 + (size_t)stride * 4 + (size_t)stride * 4 + (size_t)d * 4 + (L + 1);
    return (o + 15) & ~(size_t)15;
}

// HashSet<u32>::insert without the fill counter (the warp counts with a ballot)
__device__ __forceinline__ bool hs_insert_nc(uint32_t *tab, uint32_t mask, uint32_t key) {
    uint32_t h = (key * 2654435761u) & mask;
    for (uint32_t probes = 0; probes <= mask; probes++) {
        const uint32_t old = atomicCAS(&tab[h], kEmpty, key);
        if (old == kEmpty) return true;
        if (old == key) return false;
        h = (h + 1) & mask;
    }
    return false;
}

template <int NC2>
__global__ void __launch_bounds__(kWqWarps * 32, 8) k_greedy_search_wq(GraphArgs g, const __half *__restrict__ queries, const uint32_t *__restrict__ q_rows,
                                                                       uint32_t nq, const uint32_t *__restrict__ starts, uint32_t start_all, uint32_t L,
                                                                       uint32_t filter_from_all, uint32_t filter_self_from, uint32_t *htabs, uint32_t hcap,
                                                                       GreedyOut out, int pf_rows) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t S = g.stride;
    uint8_t *base = smem_raw + (size_t)warp * wq_warp_bytes(L, S, g.d);
    long long *nb_scores = (long long *)base;
    long long *pre_scores = nb_scores + (L + 1);
    float *qs = (float *)(pre_scores + S);            // 8-byte aligned: read as float2
    uint32_t *nb_ids = (uint32_t *)(qs + g.d);
    uint32_t *pre = nb_ids + (L + 1);
    uint32_t *raw = pre + S;
    uint8_t *nb_vis = (uint8_t *)(raw + S);
    const uint32_t gw = blockIdx.x * kWqWarps + warp, nw = gridDim.x * kWqWarps;
    uint32_t *htab = htabs + (size_t)gw * hcap;
    const uint32_t hmask = hcap - 1;
    const unsigned full = 0xffffffffu;

    for (uint32_t qi = gw; qi < nq; qi += nw) {
        if ((starts ? starts[qi] : start_all) >= g.n) {                              // not a row of this shard: empty result, status bit 2
            for (uint32_t i = lane; i < L; i += 32) { out.ids[(size_t)qi * L + i] = kEmpty; out.scores[(size_t)qi * L + i] = 0; }
            if (lane == 0) { out.len[qi] = 0; out.distances[qi] = 0; if (out.vl_len) out.vl_len[qi] = 0; out.status[qi] = 4u; }
            continue;
        }
        {
            uint4 *t4 = (uint4 *)htab;
            const uint4 e4 = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
            for (uint32_t i = lane; i < hcap / 4; i += 32) t4[i] = e4;
        }
        const size_t qrow = q_rows ? q_rows[qi] : qi;
        const uint32_t filter_from = (q_rows && qrow < filter_self_from) ? 0xFFFFFFFFu : filter_from_all;   // lib.rs:297-298
        for (uint32_t c = lane; c < g.d; c += 32) qs[c] = __half2float(queries[qrow * g.d + c]);
        __syncwarp();
        NbView nb{nb_ids, nb_scores, nb_vis, 0, (int)L, -1};
        const uint32_t start = starts ? starts[qi] : start_all;
        unsigned long long distances = 0;
        uint32_t n_eval = 0, hfill = 1;
        {
            long long sc, sc_dup;
            const __half *r = g.x + (size_t)start * g.d;
            wq_score2<NC2>(qs, r, r, g.d, lane, sc, sc_dup);
            nb_insert(nb, start, sc, lane);                                          // :188
            if (lane == 0) hs_insert_nc(htab, hmask, start);                         // :189
            __syncwarp();
        }
        for (;;) {
            const uint32_t pt = nb_next_unvisited(nb, lane);
            if (pt == kEmpty || hfill * 4 > hcap * 3) break;
            // out-neighbours not seen before, first occurrence first (:193-200)
            const uint32_t dgl = g.deg[pt];
            const uint32_t *nbrs = g.adj + (size_t)pt * S;
            int n_pre = 0;
            for (uint32_t b0 = 0; b0 < S; b0 += 32) {
                const uint32_t i = b0 + lane;
                const uint32_t rawid = i < S ? nbrs[i] : kEmpty;                      // issued before the degree is known
                const uint32_t dg = min(dgl, S);
                if (b0 >= dg) break;
                const bool have = i < dg;
                const uint32_t id = have ? rawid : kEmpty;
                const unsigned same = __match_any_sync(full, id);
                bool ins = false;
                if (have && (same & ((1u << lane) - 1)) == 0) ins = hs_insert_nc(htab, hmask, id);   // a copy in a lower lane wins
                const bool fresh = ins && id < filter_from;
                hfill += __popc(__ballot_sync(full, ins));
                const unsigned m = __ballot_sync(full, fresh);
                if (fresh) pre[n_pre + __popc(m & ((1u << lane) - 1))] = id;
                n_pre += __popc(m);
                __syncwarp();
            }
            // exact scores, two rows per pass (:201-204).  A pass is one dependent round trip to HBM (the loads of a pass are issued together,
            // the passes follow each other), so the rows of the passes pf_rows / 2 ahead are started towards L2 while this one is scored:
            // lane l < lines asks for the l-th 128-byte line of a row.
            const int lines = (int)((g.d * 2 + 127) / 128);
            auto prefetch_row = [&](int r) {
                if (r < n_pre && lane < lines) asm volatile("prefetch.global.L2 [%0];" ::"l"((const char *)(g.x + (size_t)pre[r] * g.d) + lane * 128));
            };
            for (int r = 0; r < pf_rows; r++) prefetch_row(r);
            for (int i = 0; i < n_pre; i += 2) {
                const int j = i + 1 < n_pre ? i + 1 : i;
                if (pf_rows) { prefetch_row(i + pf_rows); prefetch_row(i + pf_rows + 1); }
                long long s0, s1;
                wq_score2<NC2>(qs, g.x + (size_t)pre[i] * g.d, g.x + (size_t)pre[j] * g.d, g.d, lane, s0, s1);
                if (lane == 0) { pre_scores[i] = s0; pre_scores[j] = s1; }
            }
            __syncwarp();
            const bool full0 = nb.len == nb.cap;
            const long long last0 = full0 ? nb.scores[nb.len - 1] : 0;
            for (int i = 0; i < n_pre; i++) {
                const long long sc = pre_scores[i];
                if (!(full0 && last0 > sc)) nb_insert(nb, pre[i], sc, lane);
            }
            if (out.vl_ids) {
                for (int i = lane; i < n_pre; i += 32) {
                    const uint32_t slot = n_eval + i;
                    if (slot < out.vl_cap) {
                        out.vl_ids[(size_t)qi * out.vl_cap + slot] = pre[i];
                        out.vl_scores[(size_t)qi * out.vl_cap + slot] = pre_scores[i];
                    }
                }
            }
            n_eval += n_pre;
            distances += n_pre;
            __syncwarp();
        }
        for (uint32_t i = lane; i < L; i += 32) {
            out.ids[(size_t)qi * L + i] = (int)i < nb.len ? nb.ids[i] : kEmpty;
            out.scores[(size_t)qi * L + i] = (int)i < nb.len ? nb.scores[i] : 0;
        }
        if (lane == 0) {
            out.len[qi] = nb.len;
            out.distances[qi] = distances;
            if (out.vl_len) out.vl_len[qi] = n_eval;
            out.status[qi] = (hfill * 4 > hcap * 3) ? 1u : 0u;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ beam search over the packed index (query_disk_index.rs:144-212)

struct BeamArgs {
    const uint8_t *codes;     // [n][M] PQ codes
    uint32_t M, C;            // chunks per code, centroids per chunk (LUT is [M][C] f32)
    const uint8_t *desc;      // [n][n_desc] or NULL
    const uint8_t *has_url;   // [n] or NULL: query_disk_index.rs:172 `node.url.len() > 0`
    uint32_t n_desc;
    int disable_pq;
    const float *code_scale;  // [n] or NULL: candidate score = f32 sum of LUT entries * code_scale[id] + code_bias[q] (RabitQ estimate)
    const float *code_bias;   // [nq] (used with code_scale)
    const float *qtm;         // [nq][rq_O + 1] or NULL: RabitQ query side (P q, <mean,q>), evaluated directly against the sign codes
    uint32_t rq_O;            // 512
    float rq_scale;           // 1 / sqrt(n_dims), rounded to f32 on the host
};

// RabitQ estimate for the traversal, straight from the 64-byte sign code (diskann/rabitq.py:42-48) in the integer form of the
// RabitQ paper (arXiv 2405.12497, the paper rabitq.py cites, section 3.3): the query side P q is quantised once per query to
// B_q = 8 bit unsigned integers  qu_i = rint((Pq_i - vmin) / delta),  delta = (vmax - vmin) / 255,  and kept as 8 bit planes of
// 512 bits; then  sum_i (+-) Pq_i  ~  vmin * (2 popc(code) - 512) + delta * (2 S1 - sum_i qu_i),  S1 = sum_b 2^b popc(code & plane_b).
// One LANE scores one candidate (16 code words x 8 planes of AND + POPC), 32 candidates per pass, no shuffles, and the sum is an
// exact integer -- independent of any summation order, so the CPU oracle restates it trivially (oracle/mse_oracle.c::rabitq_int_sum).
// The quantisation adds ~1e-4 absolute to an estimate whose own noise is ~5e-2 (tests/test_rabitq_reference.py bounds it at 1e-3
// against the script's f64 result); mse_rabitq_estimate, the codec entry point, stays the script's float formula.
static constexpr int kRqBits = 8;       // B_q
static constexpr int kRqWords = 16;     // output_dims = 512 sign bits = 16 words
static constexpr int kRqPlaneWords = kRqWords * kRqBits;   // u32 [word][plane]

struct RqQuery {
    float vmin, delta;
    int qsum;
};

// whole warp: quantise (P q)[0..512) and write the planes to shared memory (planes[w * 8 + b], bit j = bit b of qu[32 w + j])
__device__ __forceinline__ RqQuery rq_quantize_query(const float *__restrict__ qt, uint32_t *planes, int lane) {
    const unsigned full = 0xffffffffu;
    float v[kRqWords];
    float lo = INFINITY, hi = -INFINITY;
#pragma unroll
    for (int w = 0; w < kRqWords; w++) {
        v[w] = qt[32 * w + lane];
        lo = fminf(lo, v[w]);
        hi = fmaxf(hi, v[w]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(full, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(full, hi, o));
    }
    const float range = hi - lo;
    const float inv = range > 0.f ? 255.0f / range : 0.f;
    int sum = 0;
#pragma unroll
    for (int w = 0; w < kRqWords; w++) {
        int u = __float2int_rn((v[w] - lo) * inv);
        u = min(max(u, 0), 255);
        sum += u;
#pragma unroll
        for (int b = 0; b < kRqBits; b++) {
            const uint32_t word = __ballot_sync(full, (u >> b) & 1);
            if (lane == 0) planes[w * kRqBits + b] = word;
        }
    }
    sum = __reduce_add_sync(full, sum);
    __syncwarp();
    return RqQuery{lo, range / 255.0f, sum};
}

// one thread: the signed sum  sum_i (+-)(P q)_i  of one 64-byte code against the planes in shared memory
__device__ __forceinline__ float rq_signed_sum(const uint32_t *planes, const uint4 *__restrict__ code, const RqQuery &rq) {
    uint32_t acc[kRqBits];
#pragma unroll
    for (int b = 0; b < kRqBits; b++) acc[b] = 0;
    uint32_t pc = 0;
#pragma unroll
    for (int v4 = 0; v4 < kRqWords / 4; v4++) {
        const uint4 c = __ldg(code + v4);
        const uint32_t cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const uint4 p0 = *reinterpret_cast<const uint4 *>(planes + (v4 * 4 + t) * kRqBits);
            const uint4 p1 = *reinterpret_cast<const uint4 *>(planes + (v4 * 4 + t) * kRqBits + 4);
            acc[0] += __popc(cw[t] & p0.x); acc[1] += __popc(cw[t] & p0.y); acc[2] += __popc(cw[t] & p0.z); acc[3] += __popc(cw[t] & p0.w);
            acc[4] += __popc(cw[t] & p1.x); acc[5] += __popc(cw[t] & p1.y); acc[6] += __popc(cw[t] & p1.z); acc[7] += __popc(cw[t] & p1.w);
            pc += __popc(cw[t]);
        }
    }
    uint32_t s1 = 0;
#pragma unroll
    for (int b = 0; b < kRqBits; b++) s1 += acc[b] << b;
    const int isum = 2 * (int)s1 - rq.qsum, csum = 2 * (int)pc - 32 * kRqWords;
    return fmaf(rq.delta, (float)isum, rq.vmin * (float)csum);
}
__device__ __forceinline__ long long rabitq_score(float signed_sum, float rq_scale, float code_scale, float bias) {
    return fast_dot_fix(fmaf(rq_scale * signed_sum, code_scale, bias));
}
struct BeamOut {
    uint32_t *ids;            // [nq][cap] expanded-and-recorded nodes in visit order
    long long *scores;        // [nq][cap] exact scores
    uint32_t *len;            // [nq]
    uint32_t cap;
    unsigned long long *cmps, *pq_cmps;  // [nq]
    uint32_t *status;         // [nq] bit 0: visited-set overflow, bit 1: more than cap expanded nodes
    uint32_t topk;            // > 0: also emit the best topk expanded nodes, (score desc, visit order asc) = the stable sort of :303
    uint32_t *top_ids;        // [nq][topk], kEmpty past top_len
    long long *top_scores;
    uint32_t *top_len;        // [nq]
};

// descriptor_product (query_disk_index.rs:135-142): sum_j trunc(scale_j * code_j * 2^32)
__device__ __forceinline__ long long descriptor_product(const BeamArgs &b, const float *scales, uint32_t id) {
    long long r = 0;
    for (uint32_t j = 0; j < b.n_desc; j++) r += fast_dot_fix(scales[j] * (float)b.desc[(size_t)id * b.n_desc + j]);
    return r;
}

__global__ void __launch_bounds__(kGsThreads) k_beam_search(GraphArgs g, BeamArgs ba, const __half *__restrict__ queries,
                                                            const float *__restrict__ luts, const float *__restrict__ desc_scales,
                                                            uint32_t nq, const uint32_t *__restrict__ starts, uint32_t start_all, uint32_t L,
                                                            uint32_t W, uint32_t *htabs, uint32_t hcap, uint32_t vcap, BeamOut out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    GsSmem s = carve(smem_raw, L, g.d, g.stride);
    float *lut = (float *)(smem_raw + ((gs_smem_bytes(L, g.d, g.stride) + 15) & ~(size_t)15));
    // two exact sets like the reference: visited_adjacent (seen as a neighbour) and visited (expanded)
    uint32_t *hadj = htabs + (size_t)blockIdx.x * (hcap + vcap), *hvis = hadj + hcap;   // vcap slots for the expanded set (~1.5 L nodes)
    const uint32_t hmask = hcap - 1, vmask = vcap - 1;
    __shared__ uint32_t fill_adj, fill_vis;
    __shared__ uint32_t pts[64];
    __shared__ __align__(16) uint32_t rq_planes[kRqPlaneWords];
    __shared__ RqQuery rq_shared;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lut_n = ba.M * ba.C;

    for (uint32_t qi = blockIdx.x; qi < nq; qi += gridDim.x) {
        if ((starts ? starts[qi] : start_all) >= g.n) {                              // not a row of this shard: empty result, status bit 2
            for (uint32_t i = threadIdx.x; i < out.topk; i += blockDim.x) { out.top_ids[(size_t)qi * out.topk + i] = kEmpty; out.top_scores[(size_t)qi * out.topk + i] = 0; }
            if (threadIdx.x == 0) { out.len[qi] = 0; out.cmps[qi] = 0; out.pq_cmps[qi] = 0; out.status[qi] = 4u; if (out.topk) out.top_len[qi] = 0; }
            continue;
        }
        for (uint32_t i = threadIdx.x; i < hcap + vcap; i += blockDim.x) hadj[i] = kEmpty;
        for (uint32_t i = threadIdx.x; i < g.d; i += blockDim.x) s.q[i] = __half2float(queries[(size_t)qi * g.d + i]);
        if (ba.qtm) {
            if (warp == 0) {
                const RqQuery r = rq_quantize_query(ba.qtm + (size_t)qi * (ba.rq_O + 1), rq_planes, lane);
                if (lane == 0) rq_shared = r;
            }
        } else if (!ba.disable_pq) {
            for (uint32_t i = threadIdx.x; i < lut_n; i += blockDim.x) lut[i] = luts[(size_t)qi * lut_n + i];
        }
        if (threadIdx.x == 0) { fill_adj = 0; fill_vis = 0; }
        __syncthreads();
        const float *scales = desc_scales ? desc_scales + (size_t)qi * ba.n_desc : nullptr;
        const float code_bias_q = !ba.code_scale ? 0.f : ba.qtm ? ba.qtm[(size_t)qi * (ba.rq_O + 1) + ba.rq_O] : ba.code_bias[qi];
        NbView nb{s.nb_ids, s.nb_scores, s.nb_vis, 0, (int)L, -1};
        const uint32_t start = starts ? starts[qi] : start_all;
        unsigned long long cmps = 0, pq_cmps = 0;
        uint32_t n_out = 0;
        if (warp == 0) {
            nb_insert(nb, start, 0, lane);                                           // :153 seeds with score 0
            if (lane == 0) hs_insert(hadj, hmask, start, &fill_adj);
        }
        __syncthreads();
        for (;;) {
            if (warp == 0) {                                                         // next_several_unvisited :83-97
                uint32_t np = 0;
                while (np < W) {
                    const uint32_t pt = nb_next_unvisited(nb, lane);
                    if (pt == kEmpty) break;
                    if (lane == 0) pts[np] = pt;
                    np++;
                }
                if (lane == 0) s.ctl[2] = (int)np;
            }
            __syncthreads();
            const uint32_t np = (uint32_t)s.ctl[2];
            if (np == 0 || fill_adj * 4 > hcap * 3 || fill_vis * 4 > vcap * 3) break;
            for (uint32_t b = 0; b < np; b++) {
                const uint32_t id = pts[b];
                // exact score of the expanded node (+ descriptor bias) :169-170
                if (warp == 0) {
                    long long sc = score_row(s.q, g.x + (size_t)id * g.d, g.d, lane);
                    if (ba.n_desc) sc += descriptor_product(ba, scales, id);
                    if (lane == 0) {
                        cmps++;
                        if (hs_insert(hvis, vmask, id, &fill_vis) && (!ba.has_url || ba.has_url[id])) {  // :172
                            if (n_out < out.cap) {
                                out.ids[(size_t)qi * out.cap + n_out] = id;
                                out.scores[(size_t)qi * out.cap + n_out] = sc;
                            }
                            n_out++;
                        }
                    }
                    n_out = __shfl_sync(0xffffffffu, n_out, 0);
                }
                // new neighbours -> ADC (or exact) -> inserts.  The pre-buffer is cleared per expanded node; the reference
                // clears it once per beam iteration (:157), which only re-scores earlier nodes' neighbours
                const int n_pre = collect_new_neighbours(g, s, id, hadj, hmask, &fill_adj, 0xFFFFFFFFu, 0);
                if (ba.disable_pq) {
                    for (int i = warp; i < n_pre; i += kGsWarps) {
                        long long sc = score_row(s.q, g.x + (size_t)s.pre[i] * g.d, g.d, lane);
                        if (lane == 0) s.pre_scores[i] = sc;
                    }
                } else if (ba.qtm) {
                    const RqQuery rq = rq_shared;
                    for (int i = threadIdx.x; i < n_pre; i += blockDim.x) {         // one thread per candidate code
                        const uint32_t id = s.pre[i];
                        const float t = rq_signed_sum(rq_planes, (const uint4 *)(ba.codes + (size_t)id * ba.M), rq);
                        s.pre_scores[i] = rabitq_score(t, ba.rq_scale, ba.code_scale[id], code_bias_q);
                    }
                } else {
                    for (int i = threadIdx.x; i < n_pre; i += blockDim.x) {        // asymmetric_dot_product vector.rs:387-405
                        const uint32_t id = s.pre[i];
                        const uint8_t *code = ba.codes + (size_t)id * ba.M;
                        float acc = 0.f;
                        if ((ba.M & 15) == 0) {                                      // 16 codes per 128-bit load; f32 adds stay in chunk order
                            const uint4 *c4 = (const uint4 *)code;
                            for (uint32_t m = 0; m < ba.M; m += 16) {
                                const uint4 v = c4[m >> 4];
                                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                                for (int j = 0; j < 16; j++) acc += lut[(m + j) * ba.C + ((w[j >> 2] >> (8 * (j & 3))) & 255u)];
                            }
                        } else {
                            for (uint32_t m = 0; m < ba.M; m++) acc += lut[m * ba.C + code[m]];
                        }
                        if (ba.code_scale) acc = fmaf(acc, ba.code_scale[id], code_bias_q);   // RabitQ: |o| <o_bar,o> <o_bar,Pq> + <mean,q>
                        s.pre_scores[i] = fast_dot_fix(acc);
                    }
                }
                __syncthreads();
                if (ba.n_desc) {
                    for (int i = threadIdx.x; i < n_pre; i += blockDim.x) s.pre_scores[i] += descriptor_product(ba, scales, s.pre[i]);
                    __syncthreads();
                }
                if (warp == 0) {
                    const bool full0 = nb.len == nb.cap;
                    const long long last0 = full0 ? nb.scores[nb.len - 1] : 0;
                    for (int i = 0; i < n_pre; i++) {
                        const long long sc = s.pre_scores[i];
                        if (!(full0 && last0 > sc)) nb_insert(nb, s.pre[i], sc, lane);
                    }
                    if (!ba.disable_pq) pq_cmps += n_pre;
                }
                __syncthreads();
            }
        }
        if (warp == 0 && lane == 0) {
            out.len[qi] = n_out;
            out.cmps[qi] = cmps;
            out.pq_cmps[qi] = pq_cmps;
            out.status[qi] = ((fill_adj * 4 > hcap * 3 || fill_vis * 4 > vcap * 3) ? 1u : 0u) | (n_out > out.cap ? 2u : 0u);
            s.ctl[2] = (int)min(n_out, out.cap);
        }
        __syncthreads();
        if (out.topk) {                                                              // rank by counting: lists are ~L long
            const uint32_t m = (uint32_t)s.ctl[2];
            const uint32_t *vi = out.ids + (size_t)qi * out.cap;
            const long long *vs = out.scores + (size_t)qi * out.cap;
            for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
                const long long si = vs[i];
                uint32_t rank = 0;
                for (uint32_t j = 0; j < m; j++) { const long long sj = vs[j]; rank += (sj > si) || (sj == si && j < i); }
                if (rank < out.topk) { out.top_ids[(size_t)qi * out.topk + rank] = vi[i]; out.top_scores[(size_t)qi * out.topk + rank] = si; }
            }
            for (uint32_t i = m + threadIdx.x; i < out.topk; i += blockDim.x) { out.top_ids[(size_t)qi * out.topk + i] = kEmpty; out.top_scores[(size_t)qi * out.topk + i] = 0; }
            if (threadIdx.x == 0) out.top_len[qi] = min(m, out.topk);
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------ beam search over RabitQ codes, one WARP per query
//
// Same traversal as k_beam_search (query_disk_index.rs:144-212) in the warp-per-query schedule of k_greedy_search_wq.  One beam
// iteration = the W popped nodes:
//   1. their exact scores, two rows per pass (they do not depend on anything below);
//   2. node by node (the order fixes which duplicate wins): adjacency list -> visited_adjacent set -> the fresh ids are appended
//      to ONE candidate array for the whole iteration (ids of node a, then node b, ...);
//   3. the candidates are scored 32 at a time, one LANE per candidate (rq_signed_sum: 64-byte code, AND + POPC against the
//      query's bit planes), all code gathers of a pass in flight together; the score stays in the lane's registers;
//   4. a ballot drops the candidates a full list would reject (lib.rs:118) and the survivors are inserted in candidate order --
//      exactly the sequence of NeighbourBuffer::insert calls the reference makes.
// No per-query table exists in any memory, so a warp's state is ~8 KB of shared memory.

__host__ __device__ static size_t bq_warp_bytes(uint32_t L, uint32_t stride, uint32_t d, uint32_t W) {
    size_t o = (size_t)(L + 1) * 8 + (size_t)W * 8 + (size_t)d * 4 + (size_t)kRqPlaneWords * 4 + (size_t)(L + 1) * 4 + (size_t)W * stride * 4 +
               (size_t)W * 4 + (L + 1);
    o = (o + 15) & ~(size_t)15;
    return o + ((kNbBatchScratch + 15) & ~(size_t)15);   // nb_insert_batch scratch at the end
}

__device__ __forceinline__ long long shfl_ll(long long v, int src) {
    const unsigned full = 0xffffffffu;
    const int lo = __shfl_sync(full, (int)(unsigned long long)v, src), hi = __shfl_sync(full, (int)((unsigned long long)v >> 32), src);
    return (long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo);
}

template <int NC2>
__global__ void __launch_bounds__(kWqWarps * 32, 8) k_beam_search_wq(GraphArgs g, BeamArgs ba, const __half *__restrict__ queries,
                                                                     const float *__restrict__ desc_scales, uint32_t nq,
                                                                     const uint32_t *__restrict__ starts, uint32_t start_all, uint32_t L, uint32_t W,
                                                                     uint32_t *htabs, uint32_t hcap, uint32_t vcap, BeamOut out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t S = g.stride;
    uint8_t *base = smem_raw + (size_t)warp * bq_warp_bytes(L, S, g.d, W);
    uint32_t *planes = (uint32_t *)base;                       // first: read as uint4
    long long *nb_scores = (long long *)(planes + kRqPlaneWords);
    long long *pt_scores = nb_scores + (L + 1);
    float *qs = (float *)(pt_scores + W);                      // 8-byte aligned: read as float2
    uint32_t *nb_ids = (uint32_t *)(qs + g.d);
    uint32_t *pre = nb_ids + (L + 1);
    uint32_t *pts = pre + (size_t)W * S;
    uint8_t *nb_vis = (uint8_t *)(pts + W);
    uint8_t *scratch = base + bq_warp_bytes(L, S, g.d, W) - ((kNbBatchScratch + 15) & ~(size_t)15);
    const uint32_t gw = blockIdx.x * kWqWarps + warp, nw = gridDim.x * kWqWarps;
    uint32_t *hadj = htabs + (size_t)gw * (hcap + vcap), *hvis = hadj + hcap;
    const uint32_t hmask = hcap - 1, vmask = vcap - 1;
    const unsigned full = 0xffffffffu;

    for (uint32_t qi = gw; qi < nq; qi += nw) {
        if ((starts ? starts[qi] : start_all) >= g.n) {                              // not a row of this shard: empty result, status bit 2
            for (uint32_t i = lane; i < out.topk; i += 32) { out.top_ids[(size_t)qi * out.topk + i] = kEmpty; out.top_scores[(size_t)qi * out.topk + i] = 0; }
            if (lane == 0) { out.len[qi] = 0; out.cmps[qi] = 0; out.pq_cmps[qi] = 0; out.status[qi] = 4u; if (out.topk) out.top_len[qi] = 0; }
            continue;
        }
        {
            uint4 *t4 = (uint4 *)hadj;
            const uint4 e4 = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
            for (uint32_t i = lane; i < (hcap + vcap) / 4; i += 32) t4[i] = e4;   // both tables
        }
        for (uint32_t c = lane; c < g.d; c += 32) qs[c] = __half2float(queries[(size_t)qi * g.d + c]);
        const RqQuery rq = rq_quantize_query(ba.qtm + (size_t)qi * (ba.rq_O + 1), planes, lane);
        const float bias = ba.qtm[(size_t)qi * (ba.rq_O + 1) + ba.rq_O];
        const float *scales = desc_scales ? desc_scales + (size_t)qi * ba.n_desc : nullptr;
        __syncwarp();
        NbView nb{nb_ids, nb_scores, nb_vis, 0, (int)L, -1};
        const uint32_t start = starts ? starts[qi] : start_all;
        unsigned long long cmps = 0, pq_cmps = 0;
        uint32_t n_out = 0, fill_adj = 1, fill_vis = 0;
        nb_insert(nb, start, 0, lane);                                               // :153 seeds with score 0
        if (lane == 0) hs_insert_nc(hadj, hmask, start);
        __syncwarp();
        for (;;) {
            uint32_t np = 0;                                                         // next_several_unvisited :83-97
            while (np < W) {
                const uint32_t pt = nb_next_unvisited(nb, lane);
                if (pt == kEmpty) break;
                if (lane == 0) pts[np] = pt;
                np++;
            }
            __syncwarp();
            if (np == 0 || fill_adj * 4 > hcap * 3 || fill_vis * 4 > vcap * 3) break;
            // Everything below used to be a chain of ~17 dependent round trips to L2 / HBM per iteration (exact rows 2, adjacency 1, one
            // visited-set atomicCAS per popped node 4, one visited_adjacent atomicCAS per 32 neighbours 8, codes 1-2).  Independent accesses
            // are now issued together: the adjacency lists and degrees start towards L2 before the exact scores are computed, the popped
            // nodes' visited-set inserts go out in one wave, and the neighbours are first LOOKED UP (plain L2 loads, all in flight at
            // once): ~5 of 6 were seen in an earlier iteration and sit in their home slot, so only the rest -- compacted in list order --
            // go through the ordered atomicCAS passes.  Same fresh candidates in the same order as before (a key found at its home slot
            // is a key the atomicCAS would have found; everything else takes the old path).
            {
                const uint32_t adj_lines = (S * 4 + 127) / 128;
                for (uint32_t i = lane; i < np * adj_lines; i += 32)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"((const char *)(g.adj + (size_t)pts[i / adj_lines] * S) + (i % adj_lines) * 128));
                if ((uint32_t)lane < np) asm volatile("prefetch.global.L2 [%0];" ::"l"(g.deg + pts[lane]));
            }
            bool fresh_mine = false;                                                 // lane b: popped node b is expanded for the first time
            {
                const uint32_t pid = (uint32_t)lane < np ? pts[lane] : kEmpty;
                const unsigned same = __match_any_sync(full, pid);
                if ((uint32_t)lane < np && (same & ((1u << lane) - 1)) == 0) fresh_mine = hs_insert_nc(hvis, vmask, pid);
            }
            for (uint32_t i = 0; i < np; i += 2) {                                   // exact scores of the expanded nodes :169
                const uint32_t j = i + 1 < np ? i + 1 : i;
                long long s0, s1;
                wq_score2<NC2>(qs, g.x + (size_t)pts[i] * g.d, g.x + (size_t)pts[j] * g.d, g.d, lane, s0, s1);
                if (lane == 0) { pt_scores[i] = s0; pt_scores[j] = s1; }
            }
            __syncwarp();
            // the adjacency lists (and degrees) of all popped nodes go into the candidate array, which is compacted in place twice below --
            // the write index never passes the read index
            const uint32_t my_deg = (uint32_t)lane < np ? min(g.deg[pts[lane]], S) : 0u;
            for (uint32_t b = 0; b < np; b++) {
                const uint32_t *nbrs = g.adj + (size_t)pts[b] * S;
                for (uint32_t i = lane; i < S; i += 32) pre[b * S + i] = nbrs[i];
            }
            __syncwarp();
            for (uint32_t b = 0; b < np; b++) {
                const uint32_t id = pts[b];
                long long sc = pt_scores[b];
                if (ba.n_desc) sc += descriptor_product(ba, scales, id);             // :170
                cmps++;
                const bool fresh_v = __shfl_sync(full, fresh_mine, (int)b);
                fill_vis += fresh_v;
                const bool rec = fresh_v && (!ba.has_url || ba.has_url[id]);              // :172
                if (rec) {
                    if (lane == 0 && n_out < out.cap) {
                        out.ids[(size_t)qi * out.cap + n_out] = id;
                        out.scores[(size_t)qi * out.cap + n_out] = sc;
                    }
                    n_out++;
                }
            }
            // look-ups, eight 32-neighbour units in flight per lane
            const uint32_t upn = (S + 31) / 32, units = np * upn;
            int n_unk = 0;
            for (uint32_t u0 = 0; u0 < units; u0 += 8) {
                uint32_t key[8], slot[8];
#pragma unroll
                for (uint32_t t = 0; t < 8; t++) {
                    const uint32_t u = u0 + t, b = u / upn, i = (u % upn) * 32 + lane;
                    const uint32_t dg = u < units ? __shfl_sync(full, my_deg, (int)b) : 0u;
                    key[t] = i < dg ? pre[b * S + i] : kEmpty;
                    slot[t] = key[t] != kEmpty ? __ldcg(hadj + ((key[t] * 2654435761u) & hmask)) : kEmpty;
                }
                __syncwarp();                                                         // every lane holds its keys before the compaction writes
#pragma unroll
                for (uint32_t t = 0; t < 8; t++) {
                    const bool unk = key[t] != kEmpty && slot[t] != key[t];
                    const unsigned m = __ballot_sync(full, unk);
                    if (unk) pre[n_unk + __popc(m & ((1u << lane) - 1))] = key[t];
                    n_unk += __popc(m);
                }
                __syncwarp();
            }
            // out-neighbours not seen as a neighbour before, first occurrence first
            int n_pre = 0;
            for (int c0 = 0; c0 < n_unk; c0 += 32) {
                const int i = c0 + lane;
                const bool have = i < n_unk;
                const uint32_t nid = have ? pre[i] : kEmpty;
                __syncwarp();                                                         // every lane has read its slot before the compaction writes
                const unsigned same = __match_any_sync(full, nid);
                bool ins = false;
                if (have && (same & ((1u << lane) - 1)) == 0) ins = hs_insert_nc(hadj, hmask, nid);
                const unsigned m = __ballot_sync(full, ins);
                if (ins) {
                    pre[n_pre + __popc(m & ((1u << lane) - 1))] = nid;
                    // the candidate's code and scale are needed a few hundred cycles from now: start them towards L2
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(ba.codes + (size_t)nid * ba.M));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(ba.code_scale + nid));
                }
                n_pre += __popc(m);
                __syncwarp();
            }
            fill_adj += n_pre;
            pq_cmps += n_pre;
            __syncwarp();
            // candidates of the whole iteration, 32 per pass: estimate in the lane, filter, ordered inserts
            for (int c0 = 0; c0 < n_pre; c0 += 32) {
                const int i = c0 + lane;
                const bool have = i < n_pre;
                const uint32_t cid = have ? pre[i] : 0u;
                long long cs = 0;
                if (have) {
                    const float t = rq_signed_sum(planes, (const uint4 *)(ba.codes + (size_t)cid * ba.M), rq);
                    cs = rabitq_score(t, ba.rq_scale, ba.code_scale[cid], bias);
                    if (ba.n_desc) cs += descriptor_product(ba, scales, cid);        // :202
                }
                const bool full0 = nb.len == nb.cap;
                const long long last0 = full0 ? nb.scores[nb.len - 1] : 0;
                unsigned m = __ballot_sync(full, have && !(full0 && last0 > cs));    // lib.rs:118 against the tail as it is now (it only grows)
                if (nb.len >= 256 && nb.cap <= 608 && __popc(m) >= 3 && nb_insert_batch(nb, cid, cs, m, lane, scratch)) m = 0;   // one merge instead of popc(m) shifts
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    nb_insert(nb, __shfl_sync(full, cid, src), shfl_ll(cs, src), lane);
                }
            }
            __syncwarp();
        }
        const uint32_t m = min(n_out, out.cap);
        if (lane == 0) {
            out.len[qi] = n_out;
            out.cmps[qi] = cmps;
            out.pq_cmps[qi] = pq_cmps;
            out.status[qi] = ((fill_adj * 4 > hcap * 3 || fill_vis * 4 > vcap * 3) ? 1u : 0u) | (n_out > out.cap ? 2u : 0u);
        }
        __syncwarp();
        if (out.topk) {
            // best topk of the visit list under (score desc, visit order asc) -- the stable sort of :303 -- by repeated selection:
            // round r picks the largest element strictly below the previous pick
            const uint32_t *vi = out.ids + (size_t)qi * out.cap;
            const long long *vs = out.scores + (size_t)qi * out.cap;
            long long prev_s = 0;
            uint32_t prev_i = 0;
            const uint32_t rounds = min(m, out.topk);
            for (uint32_t r = 0; r < rounds; r++) {
                long long best_s = 0;
                uint32_t best_i = 0xFFFFFFFFu;
                for (uint32_t i = lane; i < m; i += 32) {
                    const long long si = vs[i];
                    const bool below = r == 0 || si < prev_s || (si == prev_s && i > prev_i);
                    if (below && (best_i == 0xFFFFFFFFu || si > best_s)) { best_s = si; best_i = i; }   // ascending i: the first maximum stays
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    const long long os = shfl_ll(best_s, lane ^ o);
                    const uint32_t oi = __shfl_xor_sync(full, best_i, o);
                    if (oi != 0xFFFFFFFFu && (best_i == 0xFFFFFFFFu || os > best_s || (os == best_s && oi < best_i))) { best_s = os; best_i = oi; }
                }
                if (lane == 0) { out.top_ids[(size_t)qi * out.topk + r] = vi[best_i]; out.top_scores[(size_t)qi * out.topk + r] = best_s; }
                prev_s = best_s;
                prev_i = best_i;
            }
            for (uint32_t i = rounds + lane; i < out.topk; i += 32) { out.top_ids[(size_t)qi * out.topk + i] = kEmpty; out.top_scores[(size_t)qi * out.topk + i] = 0; }
            if (lane == 0) out.top_len[qi] = rounds;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ runtime de-duplication of the visit list + top-k
//
// query_disk_index.rs:99,486-529: a visited node is kept unless an EARLIER KEPT node has cosine (= dot product of the unit
// fp16 rows, f32) above DUPLICATES_THRESHOLD = 0.95; the kept nodes are then sorted by score.  The reference forms the whole
// n_visited x n_visited matrix with sgemm; only the entries against kept predecessors are ever read, so one warp per query
// walks the list in visit order and scores row i against the kept rows two at a time (fast_dot summation order; sgemm's order
// is unspecified, so decisions can differ only for cosines within an ulp of the threshold).
template <int NC2>
__global__ void __launch_bounds__(kWqWarps * 32) k_dedup_topk(const __half *__restrict__ x, uint32_t d, const uint32_t *__restrict__ vis_ids,
                                                              const long long *__restrict__ vis_scores, const uint32_t *__restrict__ vis_len, uint32_t cap,
                                                              uint32_t nq, float thr, uint32_t topk, uint32_t *top_ids, long long *top_scores,
                                                              uint32_t *top_len, uint32_t *kept_count) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *base = smem_raw + (size_t)warp * (((size_t)d * 4 + (size_t)cap * 4 + 15) & ~(size_t)15);
    float *row = (float *)base;                 // row i as f32
    uint32_t *kept = (uint32_t *)(row + d);     // positions (visit order) of the kept nodes
    for (uint32_t q = blockIdx.x * kWqWarps + warp; q < nq; q += gridDim.x * kWqWarps) {
        const uint32_t m = min(vis_len[q], cap);
        const uint32_t *ids = vis_ids + (size_t)q * cap;
        const long long *scs = vis_scores + (size_t)q * cap;
        uint32_t nk = 0;
        for (uint32_t i = 0; i < m; i++) {
            const __half *ri = x + (size_t)ids[i] * d;
            for (uint32_t c = lane; c < d; c += 32) row[c] = __half2float(ri[c]);
            __syncwarp();
            bool dup = false;
            for (uint32_t t = 0; t < nk && !dup; t += 2) {
                const uint32_t j0 = kept[t], j1 = kept[t + 1 < nk ? t + 1 : t];
                float f0, f1;
                wq_dot2<NC2>(row, x + (size_t)ids[j0] * d, x + (size_t)ids[j1] * d, d, lane, f0, f1);
                dup = f0 > thr || f1 > thr;
            }
            if (!dup) { if (lane == 0) kept[nk] = i; nk++; }
            __syncwarp();
        }
        // kept nodes by (score desc, visit order asc): the stable sort of :529
        for (uint32_t t = lane; t < nk; t += 32) {
            const long long st = scs[kept[t]];
            uint32_t rank = 0;
            for (uint32_t u = 0; u < nk; u++) { const long long su = scs[kept[u]]; rank += (su > st) || (su == st && u < t); }
            if (rank < topk) { top_ids[(size_t)q * topk + rank] = ids[kept[t]]; top_scores[(size_t)q * topk + rank] = st; }
        }
        for (uint32_t t = nk + lane; t < topk; t += 32) { top_ids[(size_t)q * topk + t] = kEmpty; top_scores[(size_t)q * topk + t] = 0; }
        if (lane == 0) { top_len[q] = min(nk, topk); if (kept_count) kept_count[q] = nk; }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ evaluator brute force (query_disk_index.rs:262-273)

__global__ void __launch_bounds__(256) k_scores_i64(const __half *__restrict__ x, uint64_t n, uint32_t d, const __half *__restrict__ q,
                                                    long long *__restrict__ out) {
    extern __shared__ float sq[];
    for (uint32_t i = threadIdx.x; i < d; i += blockDim.x) sq[i] = __half2float(q[i]);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < n; r += nw) {
        const long long sc = score_row(sq, x + r * d, d, lane);
        if (lane == 0) out[r] = sc;
    }
}

// adjacency ids must address rows of this shard (local ids): flags the first offender
__global__ void k_check_adjacency(const uint32_t *__restrict__ adj, const uint32_t *__restrict__ deg, uint64_t n, uint32_t stride, uint32_t *bad) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * stride) return;
    const uint64_t node = i / stride;
    const uint32_t j = (uint32_t)(i % stride);
    if (j == 0 && deg[node] > stride) atomicMin(bad, (uint32_t)node);
    if (j < min(deg[node], stride) && adj[i] >= n) atomicMin(bad, (uint32_t)node);
}

// visited-set capacity: a search evaluates ~25 x L rows at R = 64 (1.6 k at L = 64, 4-5 k at L = 192); 4 x L x R slots keep the
// table under ~15 % full, and the kernels stop with an overflow status at 75 % rather than degrade (the tables are cleared per
// query, so their size is HBM write traffic)
uint32_t greedy_hash_capacity(uint32_t L, uint32_t stride) {
    const char *e = getenv("MSE_GREEDY_HASH_SCALE");   // tuning aid
    const long scale = e ? atol(e) : 4;
    uint64_t v = (uint64_t)(L > 64 ? L : 64) * stride * (uint64_t)(scale > 0 ? scale : 4);
    uint32_t p = 1024;
    while (p < v && p < (1u << 30)) p <<= 1;
    return p;
}

// 0 = automatic (warp-per-query for batches that fill the GPU, CTA-per-query below that), 1 = CTA-per-query, 2 = warp-per-query
static std::atomic<int> g_graph_mode{0};
static bool use_wq(const mse_index *ix, uint32_t nq) {
    const int m = g_graph_mode.load(std::memory_order_relaxed);
    if (m == 1) return false;
    if (m == 2) return true;
    return nq >= (uint32_t)sm_count(ix->device) * 4;
}
// number of workers (= visited-set tables) a launch for nq queries uses; non-decreasing in nq
uint32_t greedy_grid(const mse_index *ix, uint32_t nq) {
    const uint32_t sms = (uint32_t)sm_count(ix->device);
    if (use_wq(ix, nq)) return std::min<uint32_t>((nq + kWqWarps - 1) / kWqWarps, sms * 8) * kWqWarps;
    return std::min<uint32_t>(nq, sms * 2);
}

int greedy_search_launch(mse_index *ix, const __half *d_queries, const uint32_t *d_q_rows, uint32_t nq, const uint32_t *d_starts,
                         uint32_t start, uint32_t L, uint32_t filter_from, uint32_t filter_self_from, uint32_t *d_htabs, uint32_t hcap, uint32_t workers,
                         GreedyOut o, cudaStream_t st) {
    GraphArgs g{ix->x, ix->adj, ix->deg, ix->graph_stride, ix->d, ix->n};
    if (use_wq(ix, nq) && workers >= (uint32_t)kWqWarps) {
        const bool fixed = ix->d == 1152;
        const size_t smem = wq_warp_bytes(L, ix->graph_stride, ix->d) * kWqWarps;
        if (smem <= 200 * 1024) {
            const uint32_t grid = std::min<uint32_t>(workers / kWqWarps, (nq + kWqWarps - 1) / kWqWarps);
            if (fixed) {
                MSE_CUDA(cudaFuncSetAttribute(k_greedy_search_wq<18>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_greedy_search_wq<18><<<grid, kWqWarps * 32, smem, st>>>(g, d_queries, d_q_rows, nq, d_starts, start, L, filter_from, filter_self_from, d_htabs, hcap, o, greedy_pf_rows());
            } else {
                MSE_CUDA(cudaFuncSetAttribute(k_greedy_search_wq<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_greedy_search_wq<0><<<grid, kWqWarps * 32, smem, st>>>(g, d_queries, d_q_rows, nq, d_starts, start, L, filter_from, filter_self_from, d_htabs, hcap, o, greedy_pf_rows());
            }
            MSE_LAUNCH_OK();
            return MSE_OK;
        }
    }
    const size_t smem = gs_smem_bytes(L, ix->d, ix->graph_stride);
    MSE_REQUIRE(smem <= 200 * 1024, MSE_ERR_UNSUPPORTED, "greedy_search: L=%u does not fit shared memory", L);
    MSE_CUDA(cudaFuncSetAttribute(k_greedy_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_greedy_search<<<std::min(workers, nq), kGsThreads, smem, st>>>(g, d_queries, d_q_rows, nq, d_starts, start, L, filter_from, filter_self_from, d_htabs, hcap, o);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

}  // namespace mse

using namespace mse;

// ================================================================== C ABI

static int require_graph(const mse_index *ix, const char *who) {
    MSE_REQUIRE(ix != nullptr, MSE_ERR_INVALID, "%s: NULL handle", who);
    MSE_REQUIRE(ix->adj != nullptr && ix->deg != nullptr, MSE_ERR_STATE, "%s: the index has no graph (mse_index_set_graph / mse_index_build_vamana first)", who);
    MSE_REQUIRE(ix->d % 64 == 0, MSE_ERR_UNSUPPORTED, "%s: fast_dot needs d %% 64 == 0 (vector.rs:197), d=%u", who, ix->d);
    return MSE_OK;
}

// entry points are local row numbers of this shard (IndexHeader.shards[i].medioid is a global id: subtract the shard's id_base)
static int check_host_starts(const mse_index *ix, const uint32_t *starts, uint32_t start, uint32_t nq, const char *who) {
    if (!starts) {
        MSE_REQUIRE(start < ix->n, MSE_ERR_INVALID, "%s: start=%u is not a row of this shard (n=%llu; ids are local)", who, start, (unsigned long long)ix->n);
        return MSE_OK;
    }
    for (uint32_t i = 0; i < nq; i++)
        MSE_REQUIRE(starts[i] < ix->n, MSE_ERR_INVALID, "%s: starts[%u]=%u is not a row of this shard (n=%llu; ids are local)", who, i, starts[i],
                    (unsigned long long)ix->n);
    return MSE_OK;
}

// device-pointer variant: a scalar start is checked here; entries of a device array are checked by the search kernels, which
// skip the query and set status bit 2 (mse_search_graph_check reports it)
static int check_dev_start(const mse_index *ix, const uint32_t *d_starts, uint32_t start, const char *who) {
    MSE_REQUIRE(d_starts || start < ix->n, MSE_ERR_INVALID, "%s: start=%u is not a row of this shard (n=%llu; ids are local)", who, start,
                (unsigned long long)ix->n);
    return MSE_OK;
}

// visited_adjacent table of the beam search: 4 x stride x L slots (>= 8192).  Keys inserted per query: ~2.1-3.0 k at L = 64, 3.3-4.4 k at 128,
// 4.9-6.8 k at 256, 6.6-14.8 k at 512 (1 M - 12.5 M rows, R = 64); the kernels stop with an overflow status at 75 % fill.  The table is
// where the kernel's DRAM traffic comes from (ncu r02u, L = 512: 17.4 GB read + 4.5 GB written per launch against 1.5 GB of algorithmic
// bytes: every probe of a cold table is a sector from HBM, every query clears its table), but HALVING it measured slower, not faster
// (12.5 M rows: L = 64 2.91 -> 3.67 ms, L = 512 21.1 -> 21.6 ms; MSE_BEAM_HASH_PER_L): longer probe chains are dependent round trips.
static uint32_t beam_hash_capacity(uint32_t L, uint32_t stride) {
    const char *e = getenv("MSE_BEAM_HASH_PER_L");   // tuning aid: slots per unit of L
    const long env = e ? atol(e) : 0;
    uint64_t v = (uint64_t)std::max<uint32_t>(L, 64) * (env > 0 ? (uint64_t)env : 4ull * stride);
    uint32_t p = 8192;
    while (p < v && p < (1u << 30)) p <<= 1;
    return p;
}

static uint32_t pow2_at_least(uint64_t v) {
    uint32_t p = 1024;
    while (p < v && p < (1u << 30)) p <<= 1;
    return p;
}

MSE_API int mse_index_set_graph(mse_index *ix, const uint32_t *adj, const uint32_t *deg, uint32_t stride) {
    MSE_REQUIRE(ix != nullptr && adj != nullptr && deg != nullptr && stride >= 1, MSE_ERR_INVALID, "index_set_graph: bad argument");
    MSE_REQUIRE(stride <= (uint32_t)kMaxDeg, MSE_ERR_UNSUPPORTED, "index_set_graph: stride %u exceeds %d", stride, kMaxDeg);
    MSE_CHECK(use_device(ix->device));
    if (ix->adj) cudaFree(ix->adj);
    if (ix->deg) cudaFree(ix->deg);
    ix->adj = nullptr; ix->deg = nullptr;
    MSE_CUDA(cudaMalloc(&ix->adj, std::max<size_t>(ix->n * stride * 4, 16)));
    MSE_CUDA(cudaMalloc(&ix->deg, std::max<size_t>(ix->n * 4, 16)));
    MSE_CUDA(cudaMemcpy(ix->adj, adj, ix->n * stride * 4, cudaMemcpyHostToDevice));
    MSE_CUDA(cudaMemcpy(ix->deg, deg, ix->n * 4, cudaMemcpyHostToDevice));
    ix->graph_stride = stride;
    ix->side_n = ix->n;
    if (ix->n) {   // neighbour ids are local row numbers; a global id (e.g. from a merged index) would be an out-of-bounds gather
        DevBuf bad;
        MSE_CHECK(bad.ensure(4));
        uint32_t h = 0xFFFFFFFFu;
        cudaMemcpy(bad.p, &h, 4, cudaMemcpyHostToDevice);
        const uint64_t total = ix->n * stride;
        k_check_adjacency<<<(uint32_t)((total + 255) / 256), 256>>>(ix->adj, ix->deg, ix->n, stride, bad.as<uint32_t>());
        count_launch();
        cudaError_t e = cudaMemcpy(&h, bad.p, 4, cudaMemcpyDeviceToHost);
        bad.release();
        if (e != cudaSuccess) { set_error("index_set_graph: %s", cudaGetErrorString(e)); return MSE_ERR_CUDA; }
        if (h != 0xFFFFFFFFu) {
            index_drop_side_arrays(ix);
            set_error("index_set_graph: node %u has a degree above the stride or a neighbour id >= n=%llu (ids are local to the shard)", h,
                      (unsigned long long)ix->n);
            return MSE_ERR_INVALID;
        }
    }
    return MSE_OK;
}

MSE_API int mse_index_get_graph(const mse_index *ix, uint32_t *adj, uint32_t *deg, uint32_t *stride_out) {
    MSE_CHECK(require_graph(ix, "index_get_graph"));
    MSE_CHECK(use_device(ix->device));
    if (stride_out) *stride_out = ix->graph_stride;
    if (adj) MSE_CUDA(cudaMemcpy(adj, ix->adj, ix->n * ix->graph_stride * 4, cudaMemcpyDeviceToHost));
    if (deg) MSE_CUDA(cudaMemcpy(deg, ix->deg, ix->n * 4, cudaMemcpyDeviceToHost));
    return MSE_OK;
}

MSE_API int mse_index_set_pq_codes(mse_index *ix, const uint8_t *codes, uint32_t code_size) {
    MSE_REQUIRE(ix != nullptr && codes != nullptr && code_size >= 1, MSE_ERR_INVALID, "index_set_pq_codes: bad argument");
    MSE_CHECK(use_device(ix->device));
    if (ix->pq_codes) cudaFree(ix->pq_codes);
    ix->pq_codes = nullptr;
    MSE_CUDA(cudaMalloc(&ix->pq_codes, std::max<size_t>(ix->n * code_size, 16)));
    MSE_CUDA(cudaMemcpy(ix->pq_codes, codes, ix->n * code_size, cudaMemcpyHostToDevice));
    ix->code_size = code_size;
    return MSE_OK;
}

MSE_API int mse_index_set_descriptors(mse_index *ix, const uint8_t *desc, uint32_t n_desc, const uint8_t *has_url) {
    MSE_REQUIRE(ix != nullptr, MSE_ERR_INVALID, "index_set_descriptors: NULL handle");
    MSE_CHECK(use_device(ix->device));
    if (ix->desc) cudaFree(ix->desc);
    if (ix->has_url) cudaFree(ix->has_url);
    ix->desc = nullptr; ix->has_url = nullptr; ix->n_desc = 0;
    if (desc && n_desc) {
        MSE_CUDA(cudaMalloc(&ix->desc, std::max<size_t>(ix->n * n_desc, 16)));
        MSE_CUDA(cudaMemcpy(ix->desc, desc, ix->n * n_desc, cudaMemcpyHostToDevice));
        ix->n_desc = n_desc;
    }
    if (has_url) {
        MSE_CUDA(cudaMalloc(&ix->has_url, std::max<size_t>(ix->n, 16)));
        MSE_CUDA(cudaMemcpy(ix->has_url, has_url, ix->n, cudaMemcpyHostToDevice));
    }
    return MSE_OK;
}

// greedy_search (lib.rs:183-211) for nq queries.  Host pointers.  ids/scores: [nq][L]; len/distances/status: [nq].
// starts may be NULL (every query starts at `start`).  base_vectors_only + query_breakpoint as lib.rs:196-197.
// visited_*: optional visited_list outputs ([nq][visited_cap]) -- what build_graph feeds to robust_prune.
MSE_API int mse_search_graph(mse_index *ix, const uint16_t *q_f16, uint32_t nq, uint32_t L, const uint32_t *starts, uint32_t start,
                             int base_vectors_only, uint32_t query_breakpoint, uint32_t *ids, int64_t *scores, uint32_t *len,
                             uint64_t *distances, uint32_t *visited_ids, int64_t *visited_scores, uint32_t *visited_len,
                             uint32_t visited_cap) {
    MSE_CHECK(require_graph(ix, "search_graph"));
    MSE_REQUIRE(q_f16 && ids && scores && len && distances, MSE_ERR_INVALID, "search_graph: NULL buffer");
    MSE_REQUIRE(L >= 1 && L <= 4096, MSE_ERR_UNSUPPORTED, "search_graph: L=%u out of range [1,4096]", L);
    if (nq == 0) return MSE_OK;
    MSE_CHECK(check_host_starts(ix, starts, start, nq, "search_graph"));
    MSE_CHECK(use_device(ix->device));
    const uint32_t grid = greedy_grid(ix, nq);
    const uint32_t hcap = greedy_hash_capacity(L, ix->graph_stride);
    DevBuf b_q, b_ids, b_sc, b_len, b_dist, b_st, b_h, b_starts, b_vi, b_vs, b_vl;
    int rc = MSE_OK;
    std::vector<uint32_t> status(nq);
    do {
        if ((rc = b_q.ensure((size_t)nq * ix->d * 2)) || (rc = b_ids.ensure((size_t)nq * L * 4)) || (rc = b_sc.ensure((size_t)nq * L * 8)) ||
            (rc = b_len.ensure((size_t)nq * 4)) || (rc = b_dist.ensure((size_t)nq * 8)) || (rc = b_st.ensure((size_t)nq * 4)) ||
            (rc = b_h.ensure((size_t)grid * hcap * 4)))
            break;
        if (starts && (rc = b_starts.ensure((size_t)nq * 4))) break;
        const bool want_vl = visited_ids && visited_scores && visited_len && visited_cap;
        if (want_vl && ((rc = b_vi.ensure((size_t)nq * visited_cap * 4)) || (rc = b_vs.ensure((size_t)nq * visited_cap * 8)) ||
                        (rc = b_vl.ensure((size_t)nq * 4))))
            break;
        cudaMemcpy(b_q.p, q_f16, (size_t)nq * ix->d * 2, cudaMemcpyHostToDevice);
        if (starts) cudaMemcpy(b_starts.p, starts, (size_t)nq * 4, cudaMemcpyHostToDevice);
        GreedyOut o{b_ids.as<uint32_t>(), b_sc.as<long long>(), b_len.as<uint32_t>(), b_dist.as<unsigned long long>(),
                    want_vl ? b_vi.as<uint32_t>() : nullptr, want_vl ? b_vs.as<long long>() : nullptr, want_vl ? b_vl.as<uint32_t>() : nullptr,
                    want_vl ? visited_cap : 0, b_st.as<uint32_t>()};
        if ((rc = greedy_search_launch(ix, b_q.as<__half>(), nullptr, nq, starts ? b_starts.as<uint32_t>() : nullptr, start, L,
                                       base_vectors_only ? query_breakpoint : 0xFFFFFFFFu, 0, b_h.as<uint32_t>(), hcap, grid, o, nullptr)))
            break;
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { set_error("search_graph: %s", cudaGetErrorString(e)); rc = MSE_ERR_CUDA; break; }
        cudaMemcpy(ids, b_ids.p, (size_t)nq * L * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(scores, b_sc.p, (size_t)nq * L * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(len, b_len.p, (size_t)nq * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(distances, b_dist.p, (size_t)nq * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(status.data(), b_st.p, (size_t)nq * 4, cudaMemcpyDeviceToHost);
        if (want_vl) {
            cudaMemcpy(visited_ids, b_vi.p, (size_t)nq * visited_cap * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(visited_scores, b_vs.p, (size_t)nq * visited_cap * 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(visited_len, b_vl.p, (size_t)nq * 4, cudaMemcpyDeviceToHost);
        }
        for (uint32_t i = 0; i < nq; i++)
            if (status[i]) { set_error("search_graph: visited-set table overflowed for query %u (L=%u too large for the table)", i, L); rc = MSE_ERR_UNSUPPORTED; break; }
    } while (0);
    b_q.release(); b_ids.release(); b_sc.release(); b_len.release(); b_dist.release(); b_st.release(); b_h.release(); b_starts.release();
    b_vi.release(); b_vs.release(); b_vl.release();
    return rc;
}

MSE_API int mse_search_graph_set_mode(int mode) {
    MSE_REQUIRE(mode >= 0 && mode <= 2, MSE_ERR_INVALID, "search_graph_set_mode: mode %d (0 auto, 1 CTA per query, 2 warp per query)", mode);
    g_graph_mode.store(mode, std::memory_order_relaxed);
    return MSE_OK;
}

// greedy_search with every buffer already in HBM, asynchronous on `stream` (no allocation in steady state, no sync).
// The per-query overflow status stays on the device; mse_search_graph_check reads it back.
MSE_API int mse_search_graph_dev(mse_index *ix, const uint16_t *d_q_f16, uint32_t nq, uint32_t L, const uint32_t *d_starts, uint32_t start,
                                 int base_vectors_only, uint32_t query_breakpoint, uint32_t *d_ids, int64_t *d_scores, uint32_t *d_len,
                                 uint64_t *d_distances, void *stream) {
    MSE_CHECK(require_graph(ix, "search_graph_dev"));
    MSE_REQUIRE(d_q_f16 && d_ids && d_scores && d_len && d_distances, MSE_ERR_INVALID, "search_graph_dev: NULL buffer");
    MSE_REQUIRE(L >= 1 && L <= 4096, MSE_ERR_UNSUPPORTED, "search_graph_dev: L=%u out of range [1,4096]", L);
    if (nq == 0) return MSE_OK;
    MSE_CHECK(check_dev_start(ix, d_starts, start, "search_graph_dev"));
    MSE_CHECK(use_device(ix->device));
    const uint32_t workers = greedy_grid(ix, nq);
    const uint32_t hcap = greedy_hash_capacity(L, ix->graph_stride);
    MSE_CHECK(ix->gw_htabs.ensure((size_t)workers * hcap * 4));
    MSE_CHECK(ix->gw_status.ensure((size_t)nq * 4));
    GreedyOut o{d_ids, (long long *)d_scores, d_len, (unsigned long long *)d_distances, nullptr, nullptr, nullptr, 0, ix->gw_status.as<uint32_t>()};
    return greedy_search_launch(ix, (const __half *)d_q_f16, nullptr, nq, d_starts, start, L, base_vectors_only ? query_breakpoint : 0xFFFFFFFFu, 0,
                                ix->gw_htabs.as<uint32_t>(), hcap, workers, o, (cudaStream_t)stream);
}

// synchronises the device and reports a visited-set overflow of the last mse_search_graph_dev call on this handle
MSE_API int mse_search_graph_check(mse_index *ix, uint32_t nq) {
    MSE_REQUIRE(ix != nullptr, MSE_ERR_INVALID, "search_graph_check: NULL handle");
    MSE_CHECK(use_device(ix->device));
    MSE_CUDA(cudaDeviceSynchronize());
    if (nq == 0 || !ix->gw_status.p) return MSE_OK;
    MSE_REQUIRE((size_t)nq * 4 <= ix->gw_status.cap, MSE_ERR_INVALID, "search_graph_check: nq=%u exceeds the last search", nq);
    std::vector<uint32_t> status(nq);
    MSE_CUDA(cudaMemcpy(status.data(), ix->gw_status.p, (size_t)nq * 4, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < nq; i++) {
        MSE_REQUIRE(!(status[i] & 1u), MSE_ERR_UNSUPPORTED, "search_graph: visited-set table overflowed for query %u", i);
        MSE_REQUIRE(!(status[i] & 2u), MSE_ERR_UNSUPPORTED, "search_beam_dev: query %u expanded more nodes than the visit list holds", i);
        MSE_REQUIRE(!(status[i] & 4u), MSE_ERR_INVALID, "search_*_dev: the entry point of query %u is not a row of this shard (ids are local)", i);
    }
    return MSE_OK;
}

// Beam search over the packed index (query_disk_index.rs:144-212).  luts: [nq][M*C] f32 from mse_pq_preprocess_query;
// desc_scales: [nq][n_desc] or NULL.  out_ids/out_scores: [nq][out_cap] expanded nodes in visit order (the caller sorts by
// score as :529 does); out_len/cmps/pq_cmps: [nq].
static int search_beam_impl(mse_index *ix, const uint16_t *q_f16, const float *luts, const float *desc_scales, uint32_t nq, uint32_t L,
                            uint32_t W, const uint32_t *starts, uint32_t start, int disable_pq, uint32_t n_centroids, uint32_t *out_ids,
                            int64_t *out_scores, uint32_t *out_len, uint32_t out_cap, uint64_t *cmps, uint64_t *pq_cmps, const float *code_bias) {
    MSE_CHECK(require_graph(ix, "search_beam"));
    MSE_REQUIRE(!code_bias || (ix->code_scale && !disable_pq), MSE_ERR_STATE, "search_beam_scaled: per-vector code scales missing (mse_index_set_code_scales)");
    MSE_REQUIRE(q_f16 && out_ids && out_scores && out_len && cmps && pq_cmps && out_cap >= 1, MSE_ERR_INVALID, "search_beam: NULL buffer");
    MSE_REQUIRE(disable_pq || (ix->pq_codes && luts && n_centroids), MSE_ERR_STATE, "search_beam: PQ codes / LUTs missing (mse_index_set_pq_codes)");
    MSE_REQUIRE(L >= 1 && L <= 4096 && W >= 1 && W <= 64, MSE_ERR_UNSUPPORTED, "search_beam: L=%u W=%u out of range", L, W);
    MSE_REQUIRE(!ix->n_desc || desc_scales, MSE_ERR_INVALID, "search_beam: the index has descriptors but desc_scales is NULL");
    if (nq == 0) return MSE_OK;
    MSE_CHECK(check_host_starts(ix, starts, start, nq, "search_beam"));
    MSE_CHECK(use_device(ix->device));
    const uint32_t M = ix->code_size, C = n_centroids;
    const size_t lut_bytes = disable_pq ? 16 : (size_t)M * C * 4;
    const size_t smem = ((gs_smem_bytes(L, ix->d, ix->graph_stride) + 15) & ~(size_t)15) + lut_bytes;
    MSE_REQUIRE(smem <= 220 * 1024, MSE_ERR_UNSUPPORTED, "search_beam: L=%u with a %zu-byte LUT does not fit shared memory", L, lut_bytes);
    MSE_CUDA(cudaFuncSetAttribute(k_beam_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t grid = std::min<uint32_t>(nq, (uint32_t)sm_count(ix->device) * 2);
    const uint32_t hcap = beam_hash_capacity(L, ix->graph_stride);
    const uint32_t vcap = pow2_at_least((uint64_t)std::max<uint32_t>(L, 64) * 16);
    DevBuf b_q, b_lut, b_ds, b_ids, b_sc, b_len, b_c, b_p, b_st, b_h, b_starts, b_cb;
    int rc = MSE_OK;
    std::vector<uint32_t> status(nq);
    do {
        if (code_bias && (rc = b_cb.ensure((size_t)nq * 4))) break;
        if (code_bias) cudaMemcpy(b_cb.p, code_bias, (size_t)nq * 4, cudaMemcpyHostToDevice);
        if ((rc = b_q.ensure((size_t)nq * ix->d * 2)) || (rc = b_lut.ensure(std::max<size_t>((size_t)nq * M * C * 4, 16))) ||
            (rc = b_ids.ensure((size_t)nq * out_cap * 4)) || (rc = b_sc.ensure((size_t)nq * out_cap * 8)) || (rc = b_len.ensure((size_t)nq * 4)) ||
            (rc = b_c.ensure((size_t)nq * 8)) || (rc = b_p.ensure((size_t)nq * 8)) || (rc = b_st.ensure((size_t)nq * 4)) ||
            (rc = b_h.ensure((size_t)grid * (hcap + vcap) * 4)))
            break;
        if (starts && (rc = b_starts.ensure((size_t)nq * 4))) break;
        if (ix->n_desc && (rc = b_ds.ensure((size_t)nq * ix->n_desc * 4))) break;
        cudaMemcpy(b_q.p, q_f16, (size_t)nq * ix->d * 2, cudaMemcpyHostToDevice);
        if (!disable_pq) cudaMemcpy(b_lut.p, luts, (size_t)nq * M * C * 4, cudaMemcpyHostToDevice);
        if (starts) cudaMemcpy(b_starts.p, starts, (size_t)nq * 4, cudaMemcpyHostToDevice);
        if (ix->n_desc) cudaMemcpy(b_ds.p, desc_scales, (size_t)nq * ix->n_desc * 4, cudaMemcpyHostToDevice);
        GraphArgs g{ix->x, ix->adj, ix->deg, ix->graph_stride, ix->d, ix->n};
        BeamArgs ba{ix->pq_codes, M, C, ix->desc, ix->has_url, ix->n_desc, disable_pq, code_bias ? ix->code_scale : nullptr, b_cb.as<float>(), nullptr, 0, 0.f};
        BeamOut o{b_ids.as<uint32_t>(), b_sc.as<long long>(), b_len.as<uint32_t>(), out_cap, b_c.as<unsigned long long>(),
                  b_p.as<unsigned long long>(), b_st.as<uint32_t>(), 0, nullptr, nullptr, nullptr};
        k_beam_search<<<grid, kGsThreads, smem>>>(g, ba, b_q.as<__half>(), b_lut.as<float>(), ix->n_desc ? b_ds.as<float>() : nullptr, nq,
                                                 starts ? b_starts.as<uint32_t>() : nullptr, start, L, W, b_h.as<uint32_t>(), hcap, vcap, o);
        count_launch();
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { set_error("search_beam: %s", cudaGetErrorString(e)); rc = MSE_ERR_CUDA; break; }
        cudaMemcpy(out_ids, b_ids.p, (size_t)nq * out_cap * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(out_scores, b_sc.p, (size_t)nq * out_cap * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(out_len, b_len.p, (size_t)nq * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(cmps, b_c.p, (size_t)nq * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(pq_cmps, b_p.p, (size_t)nq * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(status.data(), b_st.p, (size_t)nq * 4, cudaMemcpyDeviceToHost);
        for (uint32_t i = 0; i < nq; i++)
            if (status[i] & 1u) { set_error("search_beam: visited-set table overflowed for query %u", i); rc = MSE_ERR_UNSUPPORTED; break; }
            else if (status[i] & 2u) { set_error("search_beam: query %u expanded more than out_cap=%u nodes (the list would be truncated)", i, out_cap); rc = MSE_ERR_UNSUPPORTED; break; }
    } while (0);
    b_q.release(); b_lut.release(); b_ds.release(); b_ids.release(); b_sc.release(); b_len.release(); b_c.release(); b_p.release();
    b_st.release(); b_h.release(); b_starts.release(); b_cb.release();
    return rc;
}

MSE_API int mse_search_beam(mse_index *ix, const uint16_t *q_f16, const float *luts, const float *desc_scales, uint32_t nq, uint32_t L,
                            uint32_t W, const uint32_t *starts, uint32_t start, int disable_pq, uint32_t n_centroids, uint32_t *out_ids,
                            int64_t *out_scores, uint32_t *out_len, uint32_t out_cap, uint64_t *cmps, uint64_t *pq_cmps) {
    return search_beam_impl(ix, q_f16, luts, desc_scales, nq, L, W, starts, start, disable_pq, n_centroids, out_ids, out_scores, out_len, out_cap,
                            cmps, pq_cmps, nullptr);
}

// The same traversal with candidates ranked by  f32(sum of LUT entries) * code_scale[id] + code_bias[q]  -- the RabitQ
// estimator of diskann/rabitq.py:42-48 when luts / code_bias come from mse_rabitq_preprocess_query, the codes from
// mse_rabitq_encode and code_scale[id] = norms[id] * dots[id] (mse_index_set_code_scales).  Expanded nodes are still scored
// exactly, as query_disk_index.rs:169-170 does.
MSE_API int mse_search_beam_scaled(mse_index *ix, const uint16_t *q_f16, const float *luts, const float *code_bias, const float *desc_scales,
                                   uint32_t nq, uint32_t L, uint32_t W, const uint32_t *starts, uint32_t start, uint32_t n_centroids,
                                   uint32_t *out_ids, int64_t *out_scores, uint32_t *out_len, uint32_t out_cap, uint64_t *cmps, uint64_t *pq_cmps) {
    MSE_REQUIRE(code_bias != nullptr, MSE_ERR_INVALID, "search_beam_scaled: code_bias is NULL");
    return search_beam_impl(ix, q_f16, luts, desc_scales, nq, L, W, starts, start, 0, n_centroids, out_ids, out_scores, out_len, out_cap, cmps,
                            pq_cmps, code_bias);
}

// Beam search with every buffer in HBM, asynchronous on `stream`, top-k selected on the device.  Candidate scores come from
// d_luts ([nq][M * n_centroids], PQ ADC), or -- when d_qtm is given -- from RabitQ byte tables the kernel builds in shared
// memory out of d_qtm ([nq][output_dims + 1] from mse_rabitq_query_dev) with the per-vector scales of mse_index_set_code_scales.
MSE_API int mse_search_beam_dev(mse_index *ix, const uint16_t *d_q_f16, const float *d_luts, const float *d_qtm, uint32_t rabitq_output_dims,
                                uint32_t rabitq_n_dims, const float *d_desc_scales, uint32_t nq, uint32_t L, uint32_t W, const uint32_t *d_starts,
                                uint32_t start, uint32_t n_centroids, uint32_t topk, uint32_t *d_top_ids, int64_t *d_top_scores,
                                uint32_t *d_top_len, uint64_t *d_cmps, uint64_t *d_pq_cmps, void *stream) {
    MSE_CHECK(require_graph(ix, "search_beam_dev"));
    MSE_REQUIRE(d_q_f16 && d_top_ids && d_top_scores && d_top_len && d_cmps && d_pq_cmps && topk >= 1, MSE_ERR_INVALID, "search_beam_dev: NULL buffer");
    MSE_REQUIRE(ix->pq_codes && (d_luts || d_qtm), MSE_ERR_STATE, "search_beam_dev: codes (mse_index_set_pq_codes) and d_luts or d_qtm are required");
    MSE_REQUIRE(!d_qtm || (ix->code_scale && rabitq_output_dims == ix->code_size * 8 && rabitq_n_dims), MSE_ERR_STATE,
                "search_beam_dev: RabitQ mode needs mse_index_set_code_scales and output_dims == 8 * code size");
    MSE_REQUIRE(L >= 1 && L <= 4096 && W >= 1 && W <= 64, MSE_ERR_UNSUPPORTED, "search_beam_dev: L=%u W=%u out of range", L, W);
    MSE_REQUIRE(!ix->n_desc || d_desc_scales, MSE_ERR_INVALID, "search_beam_dev: the index has descriptors but d_desc_scales is NULL");
    if (nq == 0) return MSE_OK;
    MSE_CHECK(check_dev_start(ix, d_starts, start, "search_beam_dev"));
    MSE_CHECK(use_device(ix->device));
    const uint32_t M = ix->code_size, C = d_qtm ? 0u : n_centroids;
    MSE_REQUIRE(d_qtm || C >= 1, MSE_ERR_INVALID, "search_beam_dev: n_centroids is 0");
    MSE_REQUIRE(!d_qtm || rabitq_output_dims == 32 * kRqWords, MSE_ERR_UNSUPPORTED, "search_beam_dev: RabitQ traversal supports output_dims = %d", 32 * kRqWords);
    const uint32_t hcap = beam_hash_capacity(L, ix->graph_stride);
    const uint32_t vcap = pow2_at_least((uint64_t)std::max<uint32_t>(L, 64) * 16);
    const uint32_t cap = std::max<uint32_t>(8 * L + 64, topk);
    MSE_CHECK(ix->gw_status.ensure((size_t)nq * 4));
    MSE_CHECK(ix->gw_vis_ids.ensure((size_t)nq * cap * 4));
    MSE_CHECK(ix->gw_vis_sc.ensure((size_t)nq * cap * 8));
    MSE_CHECK(ix->gw_vis_len.ensure((size_t)nq * 4));
    ix->gw_vis_cap = cap;
    GraphArgs g{ix->x, ix->adj, ix->deg, ix->graph_stride, ix->d, ix->n};
    BeamArgs ba{ix->pq_codes, M, C, ix->desc, ix->has_url, ix->n_desc, 0, d_qtm ? ix->code_scale : nullptr, nullptr, d_qtm, rabitq_output_dims,
                d_qtm ? (float)(1.0 / sqrt((double)rabitq_n_dims)) : 0.f};
    BeamOut o{ix->gw_vis_ids.as<uint32_t>(), ix->gw_vis_sc.as<long long>(), ix->gw_vis_len.as<uint32_t>(), cap, (unsigned long long *)d_cmps,
              (unsigned long long *)d_pq_cmps, ix->gw_status.as<uint32_t>(), topk, d_top_ids, (long long *)d_top_scores, d_top_len};
    const uint32_t sms = (uint32_t)sm_count(ix->device);
    const size_t wsmem = bq_warp_bytes(L, ix->graph_stride, ix->d, W) * kWqWarps;
    if (d_qtm && use_wq(ix, nq) && wsmem <= 200 * 1024 && W <= 32) {
        // one warp per query (lane b keeps the state of popped node b)
        const uint32_t grid = std::min<uint32_t>((nq + kWqWarps - 1) / kWqWarps, sms * 8);
        MSE_CHECK(ix->gw_htabs.ensure((size_t)grid * kWqWarps * (hcap + vcap) * 4));
        if (ix->d == 1152) {
            MSE_CUDA(cudaFuncSetAttribute(k_beam_search_wq<18>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
            k_beam_search_wq<18><<<grid, kWqWarps * 32, wsmem, (cudaStream_t)stream>>>(g, ba, (const __half *)d_q_f16, ix->n_desc ? d_desc_scales : nullptr, nq,
                                                                                       d_starts, start, L, W, ix->gw_htabs.as<uint32_t>(), hcap, vcap, o);
        } else {
            MSE_CUDA(cudaFuncSetAttribute(k_beam_search_wq<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
            k_beam_search_wq<0><<<grid, kWqWarps * 32, wsmem, (cudaStream_t)stream>>>(g, ba, (const __half *)d_q_f16, ix->n_desc ? d_desc_scales : nullptr, nq,
                                                                                      d_starts, start, L, W, ix->gw_htabs.as<uint32_t>(), hcap, vcap, o);
        }
        MSE_LAUNCH_OK();
        return MSE_OK;
    }
    // one CTA per query (small batches, or PQ tables: M x n_centroids f32 in shared memory)
    const size_t smem = ((gs_smem_bytes(L, ix->d, ix->graph_stride) + 15) & ~(size_t)15) + (size_t)M * C * 4 + 16;
    MSE_REQUIRE(smem <= 220 * 1024, MSE_ERR_UNSUPPORTED, "search_beam_dev: L=%u with a %zu-byte LUT does not fit shared memory", L, (size_t)M * C * 4);
    MSE_CUDA(cudaFuncSetAttribute(k_beam_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t per_sm = (uint32_t)std::max<size_t>(1, std::min<size_t>(4, (size_t)(227 * 1024) / (smem + 1024)));
    const uint32_t grid = std::min<uint32_t>(nq, sms * per_sm);
    MSE_CHECK(ix->gw_htabs.ensure((size_t)grid * (hcap + vcap) * 4));
    k_beam_search<<<grid, kGsThreads, smem, (cudaStream_t)stream>>>(g, ba, (const __half *)d_q_f16, d_luts, ix->n_desc ? d_desc_scales : nullptr, nq,
                                                                    d_starts, start, L, W, ix->gw_htabs.as<uint32_t>(), hcap, vcap, o);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

// De-duplicated top-k of the visit lists the last mse_search_beam_dev call on this handle left in HBM
// (query_disk_index.rs:99,486-529: drop a visited node when an earlier kept one has cosine > threshold, then sort by score).
// d_kept (optional) receives the number of nodes kept per query.  Asynchronous on `stream` (same stream as the search).
MSE_API int mse_dedup_topk_dev(mse_index *ix, uint32_t nq, float threshold, uint32_t topk, uint32_t *d_top_ids, int64_t *d_top_scores,
                               uint32_t *d_top_len, uint32_t *d_kept, void *stream) {
    MSE_REQUIRE(ix != nullptr && d_top_ids && d_top_scores && d_top_len && topk >= 1, MSE_ERR_INVALID, "dedup_topk_dev: NULL buffer");
    MSE_REQUIRE(ix->gw_vis_cap && ix->gw_vis_ids.p && (size_t)nq * 4 <= ix->gw_vis_len.cap, MSE_ERR_STATE,
                "dedup_topk_dev: no visit lists (call mse_search_beam_dev with at least nq queries first)");
    MSE_REQUIRE(ix->d % 64 == 0, MSE_ERR_UNSUPPORTED, "dedup_topk_dev: d %% 64 != 0");
    if (nq == 0) return MSE_OK;
    MSE_CHECK(use_device(ix->device));
    const uint32_t cap = ix->gw_vis_cap;
    const size_t per_warp = ((size_t)ix->d * 4 + (size_t)cap * 4 + 15) & ~(size_t)15;
    const size_t smem = per_warp * kWqWarps;
    MSE_REQUIRE(smem <= 200 * 1024, MSE_ERR_UNSUPPORTED, "dedup_topk_dev: visit lists of %u entries do not fit shared memory", cap);
    const uint32_t grid = std::min<uint32_t>((nq + kWqWarps - 1) / kWqWarps, (uint32_t)sm_count(ix->device) * 8);
    auto kern = ix->d == 1152 ? k_dedup_topk<18> : k_dedup_topk<0>;
    MSE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kWqWarps * 32, smem, (cudaStream_t)stream>>>(ix->x, ix->d, ix->gw_vis_ids.as<uint32_t>(), ix->gw_vis_sc.as<long long>(),
                                                             ix->gw_vis_len.as<uint32_t>(), cap, nq, threshold, topk, d_top_ids, (long long *)d_top_scores,
                                                             d_top_len, d_kept);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

MSE_API int mse_index_set_code_scales(mse_index *ix, const float *scales) {
    MSE_REQUIRE(ix != nullptr && scales != nullptr, MSE_ERR_INVALID, "index_set_code_scales: bad argument");
    MSE_CHECK(use_device(ix->device));
    if (ix->code_scale) cudaFree(ix->code_scale);
    ix->code_scale = nullptr;
    MSE_CUDA(cudaMalloc(&ix->code_scale, std::max<size_t>(ix->n * 4, 16)));
    MSE_CUDA(cudaMemcpy(ix->code_scale, scales, ix->n * 4, cudaMemcpyHostToDevice));
    return MSE_OK;
}

// exact i64 scores of one fp16 query against every row of the index (query_disk_index.rs:262-273 without the sort)
MSE_API int mse_scores_i64(mse_index *ix, const uint16_t *q_f16, int64_t *scores) {
    MSE_REQUIRE(ix != nullptr && q_f16 && scores, MSE_ERR_INVALID, "scores_i64: NULL argument");
    MSE_REQUIRE(ix->d % 64 == 0, MSE_ERR_UNSUPPORTED, "scores_i64: d %% 64 != 0");
    if (ix->n == 0) return MSE_OK;
    MSE_CHECK(use_device(ix->device));
    DevBuf bq, bs;
    int rc = MSE_OK;
    do {
        if ((rc = bq.ensure(ix->d * 2)) || (rc = bs.ensure(ix->n * 8))) break;
        cudaMemcpy(bq.p, q_f16, ix->d * 2, cudaMemcpyHostToDevice);
        uint32_t blocks = (uint32_t)std::min<uint64_t>((ix->n + 7) / 8, (uint64_t)sm_count(ix->device) * 8);
        k_scores_i64<<<blocks, 256, ix->d * 4>>>(ix->x, ix->n, ix->d, bq.as<__half>(), bs.as<long long>());
        count_launch();
        cudaError_t e = cudaMemcpy(scores, bs.p, ix->n * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("scores_i64: %s", cudaGetErrorString(e)); rc = MSE_ERR_CUDA; }
    } while (0);
    bq.release(); bs.release();
    return rc;
}
