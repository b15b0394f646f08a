// Tensor-core scoring pass of the flat search: S = Q16 * X^T tile by tile on tcgen05, with the score
// matrix never leaving TMEM -- the epilogue compares each score against the query's running threshold
// and appends the few survivors (rank key = ordered score | ~row id) to the candidate buffer.
//
// A operand = queries (fp16-rounded copy, padded to 128 rows), B operand = index rows straight from the
// HBM-resident fp16 index (no copy, TMA reads [256 rows x 64 dims] boxes).  Tiles are numbered m-fastest so
// the CTAs that run together score the same 256 index rows against different query tiles: each index row
// is fetched from HBM once and re-read from L2 by the other query tiles.
// Algorithmic work: 2*nq*nrows*d flop; bytes: nrows*d*2 from HBM.
#include "internal.h"
#include "gemm_sm100.cuh"
#include <algorithm>

namespace mse {

static constexpr int kFlatBN = 256;

struct FlatEpilogue {
    static constexpr bool kTmaStore = false;
    __device__ __forceinline__ void compute(uint32_t, uint32_t, uint32_t (&)[32]) {}
    __device__ __forceinline__ const __half *staged_residual() const { return nullptr; }
    __device__ __forceinline__ uint32_t ldc() const { return 0; }
    __device__ __forceinline__ uint32_t rows() const { return 0; }
    __device__ __forceinline__ uint32_t cols() const { return 0; }
    const float *thr;
    uint64_t *cand;
    uint32_t *count;
    uint32_t cap, nq;
    uint32_t row_begin, row_end;  // absolute index rows scored by this launch: [row_begin, row_end)
    float th;
    __device__ __forceinline__ void begin_tile(uint32_t qrow, uint32_t) { th = qrow < nq ? thr[qrow] : INFINITY; }
    __device__ __forceinline__ void columns(uint32_t qrow, uint32_t col0, const uint32_t (&v)[32]) {
        const uint32_t r0 = row_begin + col0;  // absolute index row of v[0]
        uint32_t mask = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) mask |= (__uint_as_float(v[j]) >= th ? 1u : 0u) << j;
        if (r0 >= row_end) mask = 0;
        else if (row_end - r0 < 32) mask &= (1u << (row_end - r0)) - 1u;
        if (mask) {
            const uint32_t base = atomicAdd(&count[qrow], (uint32_t)__popc(mask));
            uint32_t slot = base;
            uint64_t *dst = cand + (size_t)qrow * cap;
#pragma unroll
            for (int j = 0; j < 32; j++) {
                if ((mask >> j) & 1u) {
                    if (slot < cap) dst[slot] = rank_key(__uint_as_float(v[j]), r0 + j);
                    slot++;
                }
            }
        }
    }
};

int flat_tc_supported(const mse_index *ix) { return ix->d % 8 == 0 && ix->d >= 64; }

int flat_tc_score_chunk(mse_index *ix, uint32_t nq, uint64_t row0, uint64_t nrows, uint32_t cap, cudaStream_t st) {
    using Cfg = GemmCfg<kFlatBN>;
    static PerDeviceOnce once;
    auto kern = k_gemm_tn<kFlatBN, 0, FlatEpilogue>;
    if (once.first(ix->device)) MSE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes));
    FlatWork &w = ix->fw;
    const uint32_t nq_pad = (nq + 127) / 128 * 128;
    if (!ix->tmap_valid) {
        MSE_CHECK(encode_tmap_2d((CUtensorMap *)ix->tmap_x, ix->x, ix->n, ix->d, ix->d, kFlatBN));
        ix->tmap_valid = true;
    }
    CUtensorMap tmq;
    MSE_CHECK(encode_tmap_2d(&tmq, w.q16.p, nq_pad, ix->d, ix->d, kGemmBM));
    GemmShape shp;
    shp.M = nq_pad;
    shp.N = (uint32_t)nrows;
    shp.K = ix->d;
    shp.tiles_m = nq_pad / kGemmBM;
    shp.tiles_n = (uint32_t)((nrows + kFlatBN - 1) / kFlatBN);
    shp.m_fastest = 1;
    shp.a_row0 = 0;
    shp.b_row0 = (int32_t)row0;
    FlatEpilogue epi;
    epi.thr = w.thr.as<float>();
    epi.cand = w.cand.as<uint64_t>();
    epi.count = w.count.as<uint32_t>();
    epi.cap = cap;
    epi.nq = nq;
    epi.row_begin = (uint32_t)row0;
    epi.row_end = (uint32_t)(row0 + nrows);
    epi.th = 0.f;
    const uint32_t ntiles = shp.tiles_m * shp.tiles_n;
    const uint32_t grid = std::min<uint32_t>(ntiles, (uint32_t)sm_count(ix->device));
    kern<<<grid, kGemmThreads, Cfg::kSmemBytes, st>>>(tmq, *(const CUtensorMap *)ix->tmap_x, tmq, shp, epi);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

}  // namespace mse
