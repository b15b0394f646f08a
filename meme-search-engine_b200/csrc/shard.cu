// Id-range sharded search across the GPUs of one box: one process per GPU, every rank holds one shard
// (an mse_index with its id_base), queries are replicated, every rank searches its shard, and the per-shard top-k lists meet in
// ONE all-gather, followed by a k-way merge on every rank.
//
// Data flow of a sharded search on one stream, without a host round trip:
//   local search  ->  its finalize kernel writes the shard's top-k as packed rank keys with GLOBAL ids straight into this
//                     rank's slot of the all-gather buffer (registered with NCCL: ncclCommRegister)
//   ncclAllGather ->  in place, slot r of every rank's buffer = rank r's list           (NVLink 5 / NVSwitch)
//   merge kernel  ->  (score desc, id asc) over n_ranks * k entries per query; ids are global and the order is total, so the
//                     result does not depend on the number of shards
// The flat search's certificate status travels in the same slot (one extra word), so every rank sees every rank's status after
// the gather and mse_search_sharded_check can decide COLLECTIVELY -- without another exchange -- whether a repair round is
// needed.
//
// NCCL is bound at run time (dlopen): libmse_b200.so keeps loading on a box without NCCL, and a one-rank group needs none.
#include "internal.h"
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <mutex>

namespace mse {

struct NcclApi {
    void *lib = nullptr;
    int version = 0;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommRegister)(const ncclComm_t, void *, size_t, void **) = nullptr;     // optional (NCCL >= 2.19)
    ncclResult_t (*CommDeregister)(const ncclComm_t, void *) = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mu;

// the NCCL already mapped into the process (e.g. the one PyTorch brought) wins, then $MSE_NCCL_LIB, then the system's libnccl.so.2
static int load_nccl() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.lib) return MSE_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    const char *env = getenv("MSE_NCCL_LIB");
    if (!h && env && *env) h = dlopen(env, RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    MSE_REQUIRE(h != nullptr, MSE_ERR_STATE, "shard group: libnccl.so.2 not found (%s); set MSE_NCCL_LIB", dlerror());
    NcclApi a;
    a.lib = h;
#define MSE_NCCL_SYM(field, name) *(void **)(&a.field) = dlsym(h, name)
    MSE_NCCL_SYM(GetVersion, "ncclGetVersion");
    MSE_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    MSE_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    MSE_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    MSE_NCCL_SYM(GetErrorString, "ncclGetErrorString");
    MSE_NCCL_SYM(AllGather, "ncclAllGather");
    MSE_NCCL_SYM(CommRegister, "ncclCommRegister");
    MSE_NCCL_SYM(CommDeregister, "ncclCommDeregister");
#undef MSE_NCCL_SYM
    MSE_REQUIRE(a.GetVersion && a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.GetErrorString && a.AllGather, MSE_ERR_STATE,
                "shard group: the NCCL library lacks a required symbol");
    a.GetVersion(&a.version);
    g_nccl = a;
    return MSE_OK;
}

#define MSE_NCCL(expr)                                                                                       \
    do {                                                                                                     \
        ncclResult_t _r = (expr);                                                                            \
        if (_r != ncclSuccess) {                                                                             \
            mse::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, mse::g_nccl.GetErrorString(_r));     \
            return MSE_ERR_CUDA;                                                                             \
        }                                                                                                    \
    } while (0)

// ---- graph lists travel as (i64 score, global id) pairs: 16 bytes per entry, {score biased to unsigned order, 1<<32 | ~id}; {0,0} = none

__global__ void k_pack_pairs(const uint32_t *__restrict__ ids, const long long *__restrict__ scores, const uint32_t *__restrict__ len,
                             uint32_t stride, uint32_t nq, uint32_t k, uint32_t id_base, ulonglong2 *__restrict__ slot) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * k) return;
    const uint32_t q = i / k, j = i % k;
    ulonglong2 e = make_ulonglong2(0ull, 0ull);
    if (j < len[q]) {
        const uint32_t id = ids[(size_t)q * stride + j];
        if (id != MSE_ID_NONE) {
            e.x = (unsigned long long)scores[(size_t)q * stride + j] ^ 0x8000000000000000ull;
            e.y = (1ull << 32) | (unsigned long long)(~(id + id_base));
        }
    }
    slot[i] = e;
}

__device__ __forceinline__ bool pair_less(const ulonglong2 &a, const ulonglong2 &b) { return a.x < b.x || (a.x == b.x && a.y < b.y); }

__global__ void __launch_bounds__(256) k_merge_pairs(const ulonglong2 *__restrict__ slots, uint32_t n_shards, size_t slot_stride, uint32_t nq,
                                                     uint32_t k, uint32_t *__restrict__ out_ids, long long *__restrict__ out_scores) {
    extern __shared__ ulonglong2 s_pairs[];
    const uint32_t q = blockIdx.x;
    const uint32_t n = n_shards * k;
    uint32_t np2 = 64;
    while (np2 < n) np2 <<= 1;
    for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x)
        s_pairs[i] = i < n ? slots[(size_t)(i / k) * slot_stride + (size_t)q * k + i % k] : make_ulonglong2(0ull, 0ull);
    __syncthreads();
    for (uint32_t k2 = 2; k2 <= np2; k2 <<= 1) {          // bitonic, descending
        for (uint32_t j = k2 >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const ulonglong2 a = s_pairs[i], b = s_pairs[ixj];
                    const bool up = (i & k2) == 0;
                    if (up ? pair_less(a, b) : pair_less(b, a)) { s_pairs[i] = b; s_pairs[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
        const ulonglong2 e = s_pairs[i];
        const bool have = e.y != 0ull;
        out_ids[(size_t)q * k + i] = have ? ~(uint32_t)e.y : MSE_ID_NONE;
        out_scores[(size_t)q * k + i] = have ? (long long)(e.x ^ 0x8000000000000000ull) : 0;
    }
}

static constexpr uint32_t kPairSortMax = 4096;   // n_shards * k entries a merge CTA sorts (64 KB of shared memory)

}  // namespace mse

using namespace mse;

struct mse_shard_group {
    int device = 0, n_ranks = 1, rank = 0;
    ncclComm_t comm = nullptr;
    void *gather = nullptr;          // all-gather buffer [n_ranks][slot words]
    size_t gather_bytes = 0;
    void *reg = nullptr;             // ncclCommRegister handle of `gather`
    DevBuf l_ids, l_sc, l_len, l_aux0, l_aux1;   // the shard's own graph-search lists before they are packed
    uint64_t n_gathers = 0;
    // the last flat search, for mse_search_sharded_check
    mse_index *p_ix = nullptr;
    uint32_t p_nq = 0, p_k = 0;
    uint32_t *p_ids = nullptr;
    float *p_scores = nullptr;
    cudaStream_t p_st = nullptr;
    size_t p_slot = 0;
    bool p_live = false;
};

static int group_buffer(mse_shard_group *g, size_t bytes) {
    if (bytes <= g->gather_bytes) return MSE_OK;
    if (g->gather) {
        MSE_CUDA(cudaDeviceSynchronize());
        if (g->reg && g_nccl.CommDeregister) g_nccl.CommDeregister(g->comm, g->reg);
        g->reg = nullptr;
        cudaFree(g->gather);
        g->gather = nullptr;
        g->gather_bytes = 0;
    }
    const size_t want = std::max<size_t>(bytes + bytes / 2, 1 << 20);
    cudaError_t e = cudaMalloc(&g->gather, want);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("shard group: cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e));
        return MSE_ERR_OOM;
    }
    g->gather_bytes = want;
    // user-buffer registration lets NCCL move the slots without staging copies (NVLS / zero-copy paths); optional
    if (g->comm && g_nccl.CommRegister && g_nccl.CommRegister(g->comm, g->gather, want, &g->reg) != ncclSuccess) g->reg = nullptr;
    return MSE_OK;
}

static int group_all_gather(mse_shard_group *g, size_t slot_words, cudaStream_t st) {
    if (g->n_ranks == 1) return MSE_OK;
    uint64_t *base = (uint64_t *)g->gather;
    MSE_NCCL(g_nccl.AllGather(base + (size_t)g->rank * slot_words, base, slot_words, ncclUint64, g->comm, st));   // in place
    g->n_gathers++;
    return MSE_OK;
}

MSE_API int mse_shard_range(uint64_t n_total, int n_ranks, int rank, uint64_t *lo, uint64_t *hi) {
    MSE_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks && lo && hi, MSE_ERR_INVALID, "shard_range: bad argument");
    *lo = (uint64_t)((unsigned __int128)n_total * (unsigned)rank / (unsigned)n_ranks);
    *hi = (uint64_t)((unsigned __int128)n_total * (unsigned)(rank + 1) / (unsigned)n_ranks);
    return MSE_OK;
}

MSE_API int mse_shard_group_unique_id(uint8_t out[MSE_SHARD_ID_BYTES]) {
    MSE_REQUIRE(out != nullptr, MSE_ERR_INVALID, "shard_group_unique_id: NULL buffer");
    static_assert(MSE_SHARD_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
    MSE_CHECK(load_nccl());
    ncclUniqueId id;
    MSE_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out, id.internal, NCCL_UNIQUE_ID_BYTES);
    return MSE_OK;
}

MSE_API int mse_shard_group_create(const uint8_t *unique_id, int n_ranks, int rank, int device, mse_shard_group **out) {
    MSE_REQUIRE(out != nullptr, MSE_ERR_INVALID, "shard_group_create: out is NULL");
    *out = nullptr;
    MSE_REQUIRE(n_ranks >= 1 && n_ranks <= 64 && rank >= 0 && rank < n_ranks, MSE_ERR_INVALID, "shard_group_create: rank %d of %d", rank, n_ranks);
    MSE_REQUIRE(n_ranks == 1 || unique_id != nullptr, MSE_ERR_INVALID, "shard_group_create: unique_id is NULL (rank 0 calls mse_shard_group_unique_id and shares the bytes)");
    MSE_CHECK(use_device(device));
    mse_shard_group *g = new mse_shard_group();
    g->device = device; g->n_ranks = n_ranks; g->rank = rank;
    if (n_ranks > 1) {
        int rc = load_nccl();
        if (rc != MSE_OK) { delete g; return rc; }
        ncclUniqueId id;
        memcpy(id.internal, unique_id, NCCL_UNIQUE_ID_BYTES);
        ncclResult_t r = g_nccl.CommInitRank(&g->comm, n_ranks, id, rank);
        if (r != ncclSuccess) {
            set_error("shard_group_create: ncclCommInitRank(rank %d of %d) -> %s", rank, n_ranks, g_nccl.GetErrorString(r));
            delete g;
            return MSE_ERR_CUDA;
        }
    }
    *out = g;
    return MSE_OK;
}

MSE_API int mse_shard_group_info(const mse_shard_group *g, int32_t out[4]) {
    MSE_REQUIRE(g && out, MSE_ERR_INVALID, "shard_group_info: NULL argument");
    out[0] = g->n_ranks; out[1] = g->rank; out[2] = g->device; out[3] = g->n_ranks > 1 ? g_nccl.version : 0;
    return MSE_OK;
}

MSE_API uint64_t mse_shard_group_gathers(const mse_shard_group *g) { return g ? g->n_gathers : 0; }

MSE_API void mse_shard_group_destroy(mse_shard_group *g) {
    if (!g) return;
    cudaSetDevice(g->device);
    cudaDeviceSynchronize();
    if (g->reg && g_nccl.CommDeregister) g_nccl.CommDeregister(g->comm, g->reg);
    if (g->gather) cudaFree(g->gather);
    g->l_ids.release(); g->l_sc.release(); g->l_len.release(); g->l_aux0.release(); g->l_aux1.release();
    if (g->comm) g_nccl.CommDestroy(g->comm);
    delete g;
}

// ---- flat

MSE_API int mse_search_flat_sharded_dev(mse_shard_group *g, mse_index *ix, const float *d_q, uint32_t nq, uint32_t k, uint32_t *d_ids,
                                        float *d_scores, void *stream) {
    MSE_REQUIRE(g && ix, MSE_ERR_INVALID, "search_flat_sharded_dev: NULL handle");
    MSE_REQUIRE(ix->device == g->device, MSE_ERR_INVALID, "search_flat_sharded_dev: the shard lives on device %d, the group on %d", ix->device, g->device);
    MSE_REQUIRE(nq == 0 || (d_q && d_ids && d_scores), MSE_ERR_INVALID, "search_flat_sharded_dev: NULL buffer");
    MSE_REQUIRE(k >= 1 && (uint64_t)g->n_ranks * k <= 8192, MSE_ERR_UNSUPPORTED, "search_flat_sharded_dev: n_ranks * k = %llu exceeds 8192",
                (unsigned long long)g->n_ranks * k);
    g->p_live = false;
    if (nq == 0) return MSE_OK;
    MSE_CHECK(use_device(g->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t slot = ((size_t)nq * k + 2) & ~(size_t)1;                 // keys + status word, 16-byte granular
    MSE_CHECK(group_buffer(g, (size_t)g->n_ranks * slot * 8));
    uint64_t *mine = (uint64_t *)g->gather + (size_t)g->rank * slot;
    MSE_CHECK(flat_search_queue(ix, d_q, nq, k, FlatOut{nullptr, nullptr, mine}, st));
    MSE_CHECK(flat_publish_status(ix, mine + (size_t)nq * k, st));
    MSE_CHECK(group_all_gather(g, slot, st));
    MSE_CHECK(flat_merge_keys(g->device, (const uint64_t *)g->gather, (uint32_t)g->n_ranks, slot, nq, k, d_ids, d_scores, st));
    g->p_ix = ix; g->p_nq = nq; g->p_k = k; g->p_ids = d_ids; g->p_scores = d_scores; g->p_st = st; g->p_slot = slot; g->p_live = true;
    return MSE_OK;
}

MSE_API int mse_search_sharded_check(mse_shard_group *g, uint32_t *repaired) {
    MSE_REQUIRE(g != nullptr, MSE_ERR_INVALID, "search_sharded_check: NULL handle");
    if (repaired) *repaired = 0;
    MSE_CHECK(use_device(g->device));
    if (!g->p_live) {
        MSE_CUDA(cudaDeviceSynchronize());
        return MSE_OK;
    }
    mse_index *ix = g->p_ix;
    cudaStream_t st = g->p_st;
    const size_t slot = g->p_slot;
    uint64_t *base = (uint64_t *)g->gather, *mine = base + (size_t)g->rank * slot;
    std::vector<uint64_t> words(g->n_ranks);
    for (int round = 0;; round++) {
        MSE_CUDA(cudaMemcpy2DAsync(words.data(), 8, base + (size_t)g->p_nq * g->p_k, slot * 8, 8, g->n_ranks, cudaMemcpyDeviceToHost, st));
        MSE_CUDA(cudaStreamSynchronize(st));
        bool any = false;
        for (uint64_t w : words) any |= w != 0;
        if (!any) break;
        MSE_REQUIRE(round < 2, MSE_ERR_CUDA, "search_sharded_check: repair did not converge");
        // every rank sees the same words, so every rank takes this branch: flagged ranks repair locally, all gather and merge again
        uint32_t rep = 0;
        if (words[g->rank] != 0) MSE_CHECK(flat_search_settle(ix, &rep));
        if (repaired) *repaired += rep;
        MSE_CHECK(flat_publish_status(ix, mine + (size_t)g->p_nq * g->p_k, st));
        MSE_CHECK(group_all_gather(g, slot, st));
        MSE_CHECK(flat_merge_keys(g->device, base, (uint32_t)g->n_ranks, slot, g->p_nq, g->p_k, g->p_ids, g->p_scores, st));
    }
    ix->pending.live = false;
    g->p_live = false;
    return MSE_OK;
}

// ---- graph (greedy_search per shard) and packed-index beam search per shard

static int gather_and_merge_pairs(mse_shard_group *g, const uint32_t *l_ids, const long long *l_sc, const uint32_t *l_len, uint32_t stride,
                                  uint32_t nq, uint32_t k, uint32_t id_base, uint32_t *d_ids, int64_t *d_scores, cudaStream_t st) {
    const size_t slot = (size_t)nq * k * 2;                                // u64 words
    uint64_t *mine = (uint64_t *)g->gather + (size_t)g->rank * slot;
    k_pack_pairs<<<(nq * k + 255) / 256, 256, 0, st>>>(l_ids, l_sc, l_len, stride, nq, k, id_base, (ulonglong2 *)mine);
    MSE_LAUNCH_OK();
    MSE_CHECK(group_all_gather(g, slot, st));
    static PerDeviceOnce once;
    if (once.first(g->device)) MSE_CUDA(cudaFuncSetAttribute(k_merge_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kPairSortMax * 16)));
    uint32_t np2 = 64;
    while (np2 < (uint32_t)g->n_ranks * k) np2 <<= 1;
    k_merge_pairs<<<nq, 256, (size_t)np2 * 16, st>>>((const ulonglong2 *)g->gather, (uint32_t)g->n_ranks, slot / 2, nq, k, d_ids, (long long *)d_scores);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

MSE_API int mse_search_graph_sharded_dev(mse_shard_group *g, mse_index *ix, const uint16_t *d_q_f16, uint32_t nq, uint32_t L, uint32_t start,
                                         uint32_t k, uint32_t *d_ids, int64_t *d_scores, uint64_t *d_distances, void *stream) {
    MSE_REQUIRE(g && ix && (nq == 0 || (d_q_f16 && d_ids && d_scores)), MSE_ERR_INVALID, "search_graph_sharded_dev: NULL argument");
    MSE_REQUIRE(ix->device == g->device, MSE_ERR_INVALID, "search_graph_sharded_dev: the shard lives on device %d, the group on %d", ix->device, g->device);
    MSE_REQUIRE(k >= 1 && k <= L && (uint64_t)g->n_ranks * k <= kPairSortMax, MSE_ERR_UNSUPPORTED, "search_graph_sharded_dev: need 1 <= k <= L and n_ranks * k <= %u",
                kPairSortMax);
    g->p_live = false;
    if (nq == 0) return MSE_OK;
    MSE_CHECK(use_device(g->device));
    cudaStream_t st = (cudaStream_t)stream;
    MSE_CHECK(g->l_ids.ensure((size_t)nq * L * 4));
    MSE_CHECK(g->l_sc.ensure((size_t)nq * L * 8));
    MSE_CHECK(g->l_len.ensure((size_t)nq * 4));
    if (!d_distances) MSE_CHECK(g->l_aux0.ensure((size_t)nq * 8));
    MSE_CHECK(group_buffer(g, (size_t)g->n_ranks * nq * k * 16));
    MSE_CHECK(mse_search_graph_dev(ix, d_q_f16, nq, L, nullptr, start, 0, 0xFFFFFFFFu, g->l_ids.as<uint32_t>(), g->l_sc.as<int64_t>(), g->l_len.as<uint32_t>(),
                                   d_distances ? d_distances : g->l_aux0.as<uint64_t>(), stream));
    return gather_and_merge_pairs(g, g->l_ids.as<uint32_t>(), g->l_sc.as<long long>(), g->l_len.as<uint32_t>(), L, nq, k, ix->id_base, d_ids, d_scores, st);
}

MSE_API int mse_search_beam_sharded_dev(mse_shard_group *g, mse_index *ix, const uint16_t *d_q_f16, const float *d_luts, const float *d_qtm,
                                        uint32_t rabitq_output_dims, uint32_t rabitq_n_dims, const float *d_desc_scales, uint32_t nq, uint32_t L, uint32_t W,
                                        uint32_t start, uint32_t n_centroids, uint32_t k, uint32_t *d_ids, int64_t *d_scores, uint64_t *d_cmps,
                                        uint64_t *d_pq_cmps, void *stream) {
    MSE_REQUIRE(g && ix && (nq == 0 || (d_q_f16 && d_ids && d_scores)), MSE_ERR_INVALID, "search_beam_sharded_dev: NULL argument");
    MSE_REQUIRE(ix->device == g->device, MSE_ERR_INVALID, "search_beam_sharded_dev: the shard lives on device %d, the group on %d", ix->device, g->device);
    MSE_REQUIRE(k >= 1 && (uint64_t)g->n_ranks * k <= kPairSortMax, MSE_ERR_UNSUPPORTED, "search_beam_sharded_dev: need n_ranks * k <= %u", kPairSortMax);
    g->p_live = false;
    if (nq == 0) return MSE_OK;
    MSE_CHECK(use_device(g->device));
    cudaStream_t st = (cudaStream_t)stream;
    MSE_CHECK(g->l_ids.ensure((size_t)nq * k * 4));
    MSE_CHECK(g->l_sc.ensure((size_t)nq * k * 8));
    MSE_CHECK(g->l_len.ensure((size_t)nq * 4));
    if (!d_cmps) MSE_CHECK(g->l_aux0.ensure((size_t)nq * 8));
    if (!d_pq_cmps) MSE_CHECK(g->l_aux1.ensure((size_t)nq * 8));
    MSE_CHECK(group_buffer(g, (size_t)g->n_ranks * nq * k * 16));
    MSE_CHECK(mse_search_beam_dev(ix, d_q_f16, d_luts, d_qtm, rabitq_output_dims, rabitq_n_dims, d_desc_scales, nq, L, W, nullptr, start, n_centroids, k,
                                  g->l_ids.as<uint32_t>(), g->l_sc.as<int64_t>(), g->l_len.as<uint32_t>(), d_cmps ? d_cmps : g->l_aux0.as<uint64_t>(),
                                  d_pq_cmps ? d_pq_cmps : g->l_aux1.as<uint64_t>(), stream));
    return gather_and_merge_pairs(g, g->l_ids.as<uint32_t>(), g->l_sc.as<long long>(), g->l_len.as<uint32_t>(), k, nq, k, ix->id_base, d_ids, d_scores, st);
}
