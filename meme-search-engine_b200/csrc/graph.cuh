// Shared declarations of the graph search / build translation units.
#pragma once
#include "internal.h"

namespace mse {

struct GraphArgs {
    const __half *x;        // [n][d]
    const uint32_t *adj;    // fixed stride: node i -> adj[i*stride .. +deg[i])
    const uint32_t *deg;
    uint32_t stride, d;
    uint64_t n;
};

struct GreedyOut {
    uint32_t *ids;          // [nq][L]   NeighbourBuffer ids, best first; 0xFFFFFFFF past len
    long long *scores;      // [nq][L]
    uint32_t *len;          // [nq]
    unsigned long long *distances;  // [nq]  GreedySearchCounters.distances
    uint32_t *vl_ids;       // [nq][vl_cap] visited_list (id, score) in evaluation order, or NULL
    long long *vl_scores;
    uint32_t *vl_len;       // [nq] (number of evaluations; may exceed vl_cap, entries past it are dropped)
    uint32_t vl_cap;
    uint32_t *status;       // [nq] 0 ok, 1 hash table overflow
};

// greedy_search (lib.rs:183-211) for nq queries, all device pointers.  Query i is queries[q_rows ? q_rows[i] : i].
// htabs: workers * hcap u32 scratch (hcap a power of two).  Neighbours >= filter_from are not evaluated (base_vectors_only,
// lib.rs:196-199) -- for every query, or with d_q_rows only for the queries whose row is >= filter_self_from (build_graph :297-298).
int greedy_search_launch(mse_index *ix, const __half *d_queries, const uint32_t *d_q_rows, uint32_t nq, const uint32_t *d_starts,
                         uint32_t start, uint32_t L, uint32_t filter_from, uint32_t filter_self_from, uint32_t *d_htabs, uint32_t hcap, uint32_t workers,
                         GreedyOut o, cudaStream_t st);
uint32_t greedy_hash_capacity(uint32_t L, uint32_t stride);
uint32_t greedy_grid(const mse_index *ix, uint32_t nq);

}  // namespace mse
