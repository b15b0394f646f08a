// Shard centroids and shard assignment on the GPU (between the embed service and the per-shard graph build).
//
//   mse_kmeans_assign   kmeans.py:78-95   `fitness`: every row's top-SPILL_K centroids by inner product against the L2-normalised
//                                         centroids, and the histogram of those assignments per rank
//   mse_kmeans_anneal   kmeans.py:73-131  `simulated_annealing`: random-walk the centroids towards equal shard sizes (the loop is the
//                                         script's, statement for statement; the fitness evaluations run on the device)
//   mse_shard_assign    dump_processor.rs:438-456  the indexer's use of those centroids: per record, shards ordered by
//                                         dot(centroid, embedding) - balance_fudge * shard_count / records_so_far, the first SHARD_SPILL
//                                         take the record.  The dots come from the device in batches; the count feedback is sequential
//                                         by definition and runs on the host exactly as written.
//
// One kernel does the work: k_centroid_scores.  A warp holds four fp16 rows in registers (lane l owns elements 64 i + 2 l + {0, 1}) and
// walks the centroids, which are staged through shared memory as fp32 in chunks of 16 (73.7 KB at d = 1152); a centroid costs one
// conflict-free LDS.64 per 64 elements for all four rows and a five-step butterfly per row.  HBM traffic is the rows, once (2304 B each);
// the arithmetic (k x d FMA per row) is what bounds it at k = 42: ~17 ms per 12.5 M rows against 4.4 ms of row traffic.  Scores are
// fp32 sums in a fixed order (per lane in element order, then the butterfly), so results do not depend on the launch shape.
#include "internal.h"
#include <algorithm>
#include <cmath>
#include <numeric>
#include <random>
#include <vector>

namespace mse {

static constexpr int kKmThreads = 256, kKmRPW = 4, kKmChunk = 16, kKmMaxSpill = 4;

// mode bit 0: write dots [rows][k] f32; bit 1: top-`spill` ids per row -> assign [rows][spill] (+ histogram hist [spill][k])
__global__ void __launch_bounds__(kKmThreads) k_centroid_scores(const __half *__restrict__ x, uint64_t row0, uint64_t rows, uint32_t d,
                                                                const float *__restrict__ cent, uint32_t k, uint32_t spill, int mode,
                                                                float *__restrict__ dots, uint32_t *__restrict__ assign, uint32_t *__restrict__ hist) {
    extern __shared__ __align__(16) float s_cent[];   // [kKmChunk][d]
    __shared__ uint32_t s_hist[kKmMaxSpill * 256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nv = d / 64;                        // full 64-element groups; the tail (d % 64, a multiple of 8) is handled per element
    const uint32_t tail0 = nv * 64;
    if (mode & 2)
        for (uint32_t i = threadIdx.x; i < spill * k; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const uint64_t groups = (rows + kKmRPW - 1) / kKmRPW;
    const uint64_t wpg = (uint64_t)gridDim.x * (kKmThreads / 32);
    // every warp of a CTA must take part in the staging barriers: iterate CTA-uniformly
    for (uint64_t g0 = (uint64_t)blockIdx.x * (kKmThreads / 32); g0 < groups; g0 += wpg) {
        const uint64_t g = g0 + warp;
        const bool active = g < groups;
        uint32_t v[kKmRPW][18];
        float tailv[kKmRPW][2];
        uint64_t rix[kKmRPW];
#pragma unroll
        for (int r = 0; r < kKmRPW; r++) {
            rix[r] = g * kKmRPW + r;
            const bool ok = active && rix[r] < rows;
            const __half *xr = x + (row0 + (ok ? rix[r] : 0)) * d;
#pragma unroll
            for (int i = 0; i < 18; i++) v[r][i] = (ok && (uint32_t)i < nv) ? *(const uint32_t *)(xr + 64 * i + 2 * lane) : 0u;
            tailv[r][0] = (ok && tail0 + 2 * lane < d) ? __half2float(xr[tail0 + 2 * lane]) : 0.f;
            tailv[r][1] = (ok && tail0 + 2 * lane + 1 < d) ? __half2float(xr[tail0 + 2 * lane + 1]) : 0.f;
        }
        float best[kKmRPW][kKmMaxSpill];
        uint32_t bid[kKmRPW][kKmMaxSpill];
#pragma unroll
        for (int r = 0; r < kKmRPW; r++)
#pragma unroll
            for (int j = 0; j < kKmMaxSpill; j++) { best[r][j] = -INFINITY; bid[r][j] = 0xFFFFFFFFu; }
        for (uint32_t c0 = 0; c0 < k; c0 += kKmChunk) {
            const uint32_t nc = min((uint32_t)kKmChunk, k - c0);
            __syncthreads();                           // the previous chunk is no longer read
            for (uint32_t i = threadIdx.x; i < nc * d / 4; i += blockDim.x) ((float4 *)s_cent)[i] = ((const float4 *)(cent + (size_t)c0 * d))[i];
            __syncthreads();
            for (uint32_t c = 0; c < nc; c++) {
                const float *cr = s_cent + (size_t)c * d;
                float acc[kKmRPW];
#pragma unroll
                for (int r = 0; r < kKmRPW; r++) acc[r] = 0.f;
#pragma unroll
                for (int i = 0; i < 18; i++) {
                    if ((uint32_t)i < nv) {
                        const float2 cv = *(const float2 *)(cr + 64 * i + 2 * lane);
#pragma unroll
                        for (int r = 0; r < kKmRPW; r++) {
                            const float2 xv = __half22float2(*(const __half2 *)&v[r][i]);
                            acc[r] = fmaf(xv.x, cv.x, acc[r]);
                            acc[r] = fmaf(xv.y, cv.y, acc[r]);
                        }
                    }
                }
                if (tail0 + 2 * lane < d) {
                    const float2 cv = *(const float2 *)(cr + tail0 + 2 * lane);
#pragma unroll
                    for (int r = 0; r < kKmRPW; r++) acc[r] = fmaf(tailv[r][1], cv.y, fmaf(tailv[r][0], cv.x, acc[r]));
                }
#pragma unroll
                for (int r = 0; r < kKmRPW; r++) {
#pragma unroll
                    for (int o = 16; o; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
                    const uint32_t cid = c0 + c;
                    if ((mode & 1) && lane == 0 && active && rix[r] < rows) dots[rix[r] * k + cid] = acc[r];
                    if (mode & 2) {
                        // insert into the descending list; on equal scores the lower centroid index stays ahead
                        float s = acc[r];
                        uint32_t id = cid;
                        bool carried = false;                          // once inserted, everything below moves down one place
#pragma unroll
                        for (int j = 0; j < kKmMaxSpill; j++) {
                            if ((uint32_t)j < spill && (carried || s > best[r][j])) {
                                carried = true;
                                const float ts = best[r][j];
                                const uint32_t ti = bid[r][j];
                                best[r][j] = s;
                                bid[r][j] = id;
                                s = ts;
                                id = ti;
                            }
                        }
                    }
                }
            }
        }
        if ((mode & 2) && lane == 0 && active) {
#pragma unroll
            for (int r = 0; r < kKmRPW; r++) {
                if (rix[r] < rows) {
#pragma unroll
                    for (int j = 0; j < kKmMaxSpill; j++) {
                        if ((uint32_t)j < spill) {
                            if (assign) assign[rix[r] * spill + j] = bid[r][j];
                            if (bid[r][j] < k) atomicAdd(&s_hist[j * k + bid[r][j]], 1u);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    if (mode & 2)
        for (uint32_t i = threadIdx.x; i < spill * k; i += blockDim.x)
            if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

static int km_launch(mse_index *ix, uint64_t row0, uint64_t rows, const float *d_cent, uint32_t k, uint32_t spill, int mode, float *d_dots,
                     uint32_t *d_assign, uint32_t *d_hist, cudaStream_t st) {
    const size_t smem = (size_t)kKmChunk * ix->d * 4;
    static PerDeviceOnce once;
    if (once.first(ix->device)) MSE_CUDA(cudaFuncSetAttribute(k_centroid_scores, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const uint64_t groups = (rows + kKmRPW - 1) / kKmRPW;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((groups + kKmThreads / 32 - 1) / (kKmThreads / 32), (uint64_t)sm_count(ix->device) * 2);
    k_centroid_scores<<<grid, kKmThreads, smem, st>>>(ix->x, row0, rows, ix->d, d_cent, k, spill, mode, d_dots, d_assign, d_hist);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

static int km_check_args(const mse_index *ix, const float *centroids, uint32_t k, uint32_t spill, const char *who) {
    MSE_REQUIRE(ix != nullptr && centroids != nullptr, MSE_ERR_INVALID, "%s: NULL argument", who);
    MSE_REQUIRE(ix->n > 0, MSE_ERR_STATE, "%s: the index holds no rows", who);
    MSE_REQUIRE(k >= 1 && k <= 256 && spill >= 1 && spill <= (uint32_t)kKmMaxSpill && spill <= k, MSE_ERR_UNSUPPORTED,
                "%s: k=%u (1..256), spill=%u (1..%d, <= k)", who, k, spill, kKmMaxSpill);
    MSE_REQUIRE(ix->d % 8 == 0 && ix->d <= 1152 + 63 && (size_t)kKmChunk * ix->d * 4 <= 200 * 1024, MSE_ERR_UNSUPPORTED, "%s: d=%u unsupported", who, ix->d);
    return MSE_OK;
}

// torch.nn.functional.normalize(c) (kmeans.py:82): c / max(|c|_2, 1e-12), in f32
static void km_normalize(std::vector<float> &c, uint32_t k, uint32_t d) {
    for (uint32_t i = 0; i < k; i++) {
        double s = 0;
        for (uint32_t j = 0; j < d; j++) s += (double)c[(size_t)i * d + j] * c[(size_t)i * d + j];
        const float inv = 1.0f / std::max((float)std::sqrt(s), 1e-12f);
        for (uint32_t j = 0; j < d; j++) c[(size_t)i * d + j] *= inv;
    }
}

struct KmWork {
    DevBuf cent, hist;
    ~KmWork() { cent.release(); hist.release(); }
};

// counts [spill][k] for one set of centroids (host f32, already normalised if wanted); assign_dev optional
static int km_counts(mse_index *ix, KmWork &w, const float *cent, uint32_t k, uint32_t spill, uint32_t *counts, uint32_t *d_assign) {
    MSE_CHECK(w.cent.ensure((size_t)k * ix->d * 4));
    MSE_CHECK(w.hist.ensure((size_t)spill * k * 4));
    MSE_CUDA(cudaMemcpyAsync(w.cent.p, cent, (size_t)k * ix->d * 4, cudaMemcpyHostToDevice, ix->stream));
    MSE_CUDA(cudaMemsetAsync(w.hist.p, 0, (size_t)spill * k * 4, ix->stream));
    MSE_CHECK(km_launch(ix, 0, ix->n, w.cent.as<float>(), k, spill, 2, nullptr, d_assign, w.hist.as<uint32_t>(), ix->stream));
    MSE_CUDA(cudaMemcpyAsync(counts, w.hist.p, (size_t)spill * k * 4, cudaMemcpyDeviceToHost, ix->stream));
    MSE_CUDA(cudaStreamSynchronize(ix->stream));
    return MSE_OK;
}

// kmeans.py:92-94: distances_from_ideal_cluster_size = |cluster_sizes - desired_size|; (max over everything, argmax per spill rank)
static float km_fitness(const uint32_t *counts, uint32_t k, uint32_t spill, double desired, uint32_t *worst) {
    float mx = -1.f;
    for (uint32_t j = 0; j < spill; j++) {
        float mj = -1.f;
        for (uint32_t c = 0; c < k; c++) {
            const float dist = std::fabs((float)counts[j * k + c] - (float)desired);
            if (dist > mj) { mj = dist; worst[j] = c; }     // torch.argmax returns the first maximum
            mx = std::max(mx, dist);
        }
    }
    return mx;
}

}  // namespace mse

using namespace mse;

MSE_API int mse_kmeans_assign(mse_index *ix, const float *centroids, uint32_t k, uint32_t spill, int normalize, uint32_t *counts, uint32_t *assign) {
    MSE_CHECK(km_check_args(ix, centroids, k, spill, "kmeans_assign"));
    MSE_REQUIRE(counts != nullptr, MSE_ERR_INVALID, "kmeans_assign: NULL counts");
    MSE_CHECK(use_device(ix->device));
    std::vector<float> c(centroids, centroids + (size_t)k * ix->d);
    if (normalize) km_normalize(c, k, ix->d);
    KmWork w;
    DevBuf a;
    int rc = MSE_OK;
    do {
        if (assign && (rc = a.ensure((size_t)ix->n * spill * 4))) break;
        if ((rc = km_counts(ix, w, c.data(), k, spill, counts, assign ? a.as<uint32_t>() : nullptr))) break;
        if (assign && cudaMemcpy(assign, a.p, (size_t)ix->n * spill * 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
            set_error("kmeans_assign: D2H failed");
            rc = MSE_ERR_CUDA;
        }
    } while (0);
    a.release();
    return rc;
}

MSE_API int mse_kmeans_anneal(mse_index *ix, uint32_t k, uint32_t spill, uint32_t max_iter, uint64_t seed, float *centroids_out, float *fitness_out,
                              uint32_t *iters_out) {
    MSE_REQUIRE(centroids_out != nullptr, MSE_ERR_INVALID, "kmeans_anneal: NULL output");
    MSE_CHECK(km_check_args(ix, centroids_out, k, spill, "kmeans_anneal"));
    MSE_CHECK(use_device(ix->device));
    const uint32_t d = ix->d;
    const size_t nel = (size_t)k * d;
    std::mt19937_64 rng(seed);
    std::normal_distribution<float> gauss(0.f, 1.f);
    std::vector<float> cent(nel), cand(nel), normed(nel);
    std::vector<uint32_t> counts((size_t)spill * k), worst(spill), worst_new(spill);
    KmWork w;
    const double desired = (double)ix->n / k;                                     // :76
    auto fitness = [&](const std::vector<float> &c, std::vector<uint32_t> &wc, float &f) -> int {
        normed = c;
        km_normalize(normed, k, d);                                               // :82
        MSE_CHECK(km_counts(ix, w, normed.data(), k, spill, counts.data(), nullptr));
        f = km_fitness(counts.data(), k, spill, desired, wc.data());
        return MSE_OK;
    };
    for (auto &v : cent) v = gauss(rng);                                          // :75
    float temperature = 1.0f, last_fitness = 0.f, new_fitness = 0.f;
    MSE_CHECK(fitness(cent, worst, last_fitness));                                // :101
    uint32_t last_improvement = 0, it = 0;
    for (; it < max_iter; it++) {                                                 // :103
        for (size_t i = 0; i < nel; i++) cand[i] = cent[i] + gauss(rng) * temperature;   // :104
        MSE_CHECK(fitness(cand, worst_new, new_fitness));
        if (new_fitness < last_fitness) {                                         // :107-111
            cent = cand;
            temperature *= 0.999f;
            last_fitness = new_fitness;
            last_improvement = 0;
        } else {                                                                  // :112-114
            temperature *= 0.9995f;
            last_improvement++;
        }
        if (last_improvement > 100) {                                             // :115-120 "rerolling": the worst centroid of every rank
            for (uint32_t j = 0; j < spill; j++)
                for (uint32_t e = 0; e < d; e++) cent[(size_t)worst_new[j] * d + e] = gauss(rng);
            last_improvement = 0;
            temperature *= 1.1f;
            last_fitness = new_fitness;
        }
        if (last_fitness < desired * 0.1) { it++; break; }                        // :121-122
        temperature = std::min(1.5f, temperature);                                // :123
    }
    km_normalize(cent, k, d);                                                     // :131
    std::copy(cent.begin(), cent.end(), centroids_out);
    if (fitness_out) *fitness_out = last_fitness;
    if (iters_out) *iters_out = it;
    return MSE_OK;
}

MSE_API int mse_shard_assign(mse_index *ix, const float *centroids, uint32_t k, uint32_t spill, double balance_fudge, uint64_t *shard_counts,
                             uint64_t *bal_count, uint32_t *assign) {
    MSE_CHECK(km_check_args(ix, centroids, k, spill, "shard_assign"));
    MSE_REQUIRE(shard_counts && bal_count && assign, MSE_ERR_INVALID, "shard_assign: NULL buffer");
    MSE_REQUIRE(*bal_count >= 1, MSE_ERR_INVALID, "shard_assign: bal_count starts at 1 (dump_processor.rs:426), got 0");
    MSE_CHECK(use_device(ix->device));
    const uint64_t batch = 1u << 18;
    DevBuf cent, dots;
    std::vector<float> h((size_t)std::min<uint64_t>(batch, ix->n) * k);
    std::vector<uint32_t> order(k);
    std::iota(order.begin(), order.end(), 0u);   // `shards` is sorted in place record after record (:441): ties keep the previous record's order
    std::vector<int64_t> key(k);
    int rc = MSE_OK;
    do {
        if ((rc = cent.ensure((size_t)k * ix->d * 4)) || (rc = dots.ensure(h.size() * 4))) break;
        if (cudaMemcpy(cent.p, centroids, (size_t)k * ix->d * 4, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("shard_assign: H2D failed"); rc = MSE_ERR_CUDA; break; }
        for (uint64_t r0 = 0; r0 < ix->n && rc == MSE_OK; r0 += batch) {
            const uint64_t rows = std::min<uint64_t>(batch, ix->n - r0);
            if ((rc = km_launch(ix, r0, rows, cent.as<float>(), k, spill, 1, dots.as<float>(), nullptr, nullptr, ix->stream))) break;
            if (cudaMemcpyAsync(h.data(), dots.p, (size_t)rows * k * 4, cudaMemcpyDeviceToHost, ix->stream) != cudaSuccess ||
                cudaStreamSynchronize(ix->stream) != cudaSuccess) {
                set_error("shard_assign: D2H failed");
                rc = MSE_ERR_CUDA;
                break;
            }
            for (uint64_t i = 0; i < rows; i++) {
                // dump_processor.rs:441-445: sort_by_cached_key(-scale_dot_result_f64(dot - fudge * count / bal_count)), stable
                for (uint32_t c = 0; c < k; c++) {
                    double dot = (double)h[i * k + c];
                    dot -= balance_fudge * ((double)shard_counts[c] / (double)*bal_count);
                    key[c] = -(int64_t)(dot * 4294967296.0);
                }
                std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
                for (uint32_t j = 0; j < spill; j++) {                          // :452-455
                    assign[(r0 + i) * spill + j] = order[j];
                    shard_counts[order[j]]++;
                }
                (*bal_count)++;                                                   // :457
            }
        }
    } while (0);
    cent.release();
    dots.release();
    return rc;
}
