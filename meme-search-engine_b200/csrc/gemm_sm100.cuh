// Persistent, warp-specialised tcgen05 GEMM mainloop for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T
// (both operands K-major fp16/bf16 in global memory, fp32 accumulation in TMEM).
//
//   warp 0      TMA producer: one elected lane streams 128x64 A tiles and BNx64 B tiles (SWIZZLE_128B)
//               into a kStages-deep shared-memory ring, signalling full[stage] with complete_tx bytes
//   warp 1      MMA issuer: one lane issues tcgen05.mma (M=128, N=BN, K=16) x4 per stage into one of two
//               TMEM accumulators (2 x BN columns), frees the stage with tcgen05.commit -> empty[stage],
//               and publishes the finished tile with tcgen05.commit -> tmem_full[acc]
//   warps 2..9  epilogue: warp w may touch TMEM lanes 32*(w%4)..+31 (two warps per quadrant split the columns); each thread owns one output row,
//               pulls 32 fp32 columns at a time with tcgen05.ld.32x32b.x32 and hands them to the Epilogue
//               functor; the accumulator is released with tmem_empty[acc] so the next tile's MMAs overlap
//
// The Epilogue functor decides what a tile is used for (threshold filter for flat search, bias/GELU/residual
// stores for the encoder).  Tiles are numbered n-fastest or m-fastest so that CTAs running at the same time
// share one operand through L2.
#pragma once
#include "ptx.cuh"

namespace mse {

static constexpr int kGemmBM = 128;
static constexpr int kGemmBK = 64;
static constexpr int kGemmEpiWarps = 8;   // two warps per TMEM lane quadrant, each draining half of the tile's columns
static constexpr int kGemmThreads = 64 + 32 * kGemmEpiWarps;

static constexpr uint32_t kGemmCStageBytes = kGemmEpiWarps * 2 * 32 * 128;  // TMA-store epilogue: per warp two [32 rows][64 halfs] SW128 buffers

// CG = 1: one CTA per 128 x BN tile.  CG = 2 (cta_group::2): a CTA pair per 256 x BN tile; each CTA stages its own 128 rows of A
// and HALF of the B tile, so shared-memory and L2 traffic per flop drop by a third.
template <int BN, bool TMA_STORE = false, int CG = 1>
struct GemmCfg {
    static constexpr uint32_t kABytes = kGemmBM * kGemmBK * 2;
    static constexpr uint32_t kBBytes = (BN / CG) * kGemmBK * 2;
    // stage ring sized to what is left of the 227 KB after the optional 64 KB of C staging
    static constexpr int kStages = ((TMA_STORE ? 160 : 224) * 1024) / (kABytes + kBBytes) > 6 ? 6 : ((TMA_STORE ? 160 : 224) * 1024) / (kABytes + kBBytes);
    static constexpr uint32_t kStageBytes = kABytes + kBBytes;
    static constexpr uint32_t kBarBytes = 256;
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + (TMA_STORE ? kGemmCStageBytes : 0) + kBarBytes + 1024;  // + slack for 1024-B alignment
    static constexpr uint32_t kTmemCols = 512;
};

struct GemmShape {
    uint32_t M, N, K;
    uint32_t tiles_m, tiles_n;
    uint32_t m_fastest;   // 1: consecutive tile ids walk M (share the B tile); 0: walk N (share the A tile)
    int32_t a_row0, b_row0;  // row offsets added to the TMA coordinates (chunking without re-encoding maps)
};

__device__ __forceinline__ void gemm_tile_coords(const GemmShape &s, uint32_t tile, uint32_t &mt, uint32_t &nt) {
    if (s.m_fastest) { mt = tile % s.tiles_m; nt = tile / s.tiles_m; }
    else { nt = tile % s.tiles_n; mt = tile / s.tiles_n; }
}

template <int BN, class Epilogue, int CG = 1>
constexpr uint32_t gemm_smem_bytes() { return GemmCfg<BN, Epilogue::kTmaStore, CG>::kSmemBytes; }

template <int BN, int BF16, class Epilogue, int CG = 1>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_gemm_tn(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
          const GemmShape shp, Epilogue epi) {
    using Cfg = GemmCfg<BN, Epilogue::kTmaStore, CG>;
    constexpr int kTileM = kGemmBM * CG;  // rows of C per tile (shp.tiles_m counts these)
    const uint32_t cta_rank = CG == 2 ? ptx::cluster_ctarank() : 0;  // 0 = leader (issues the MMAs)
    const uint32_t unit = blockIdx.x / CG, n_units = gridDim.x / CG;   // a unit = the CTA or CTA pair that owns a tile
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    // [stage ring][C staging (TMA-store epilogue only; 1024-aligned for SWIZZLE_128B)][barriers]
    uint8_t *cstage = smem + Cfg::kStages * Cfg::kStageBytes;
    uint64_t *bars = (uint64_t *)(cstage + (Epilogue::kTmaStore ? kGemmCStageBytes : 0));
    uint64_t *full = bars;
    uint64_t *empty = bars + Cfg::kStages;
    uint64_t *tfull = bars + 2 * Cfg::kStages;
    uint64_t *tempty = tfull + 2;
    uint32_t *tmem_slot = (uint32_t *)(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t ntiles = shp.tiles_m * shp.tiles_n;
    const uint32_t nkb = (shp.K + kGemmBK - 1) / kGemmBK;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
        for (int s = 0; s < Cfg::kStages; s++) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            ptx::mbar_init(&tfull[a], 1);
            ptx::mbar_init(&tempty[a], kGemmEpiWarps * CG);  // one arrive per epilogue warp (of both CTAs of a pair, on the leader)
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (CG == 2) { ptx::tmem_alloc_pair(tmem_slot, Cfg::kTmemCols); ptx::tmem_relinquish_pair(); }
        else { ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols); ptx::tmem_relinquish(); }
    }
    ptx::tc_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync();  // the peer's barriers must exist before remote arrives / TMA signals reach them
    else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t tile = unit; tile < ntiles; tile += n_units) {
                uint32_t mt, nt;
                gemm_tile_coords(shp, tile, mt, nt);
                const int32_t a_row = shp.a_row0 + (int32_t)(mt * kTileM + cta_rank * kGemmBM);
                const int32_t b_row = shp.b_row0 + (int32_t)(nt * BN + cta_rank * (BN / CG));
                for (uint32_t kb = 0; kb < nkb; kb++) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t *sa = smem + stage * Cfg::kStageBytes;
                    if constexpr (CG == 2) {
                        // both CTAs load their halves; all bytes are accounted on the leader's barrier, which only the leader arms
                        if (cta_rank == 0) ptx::mbar_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
                        ptx::tma_load_2d_pair(sa, &tmA, &full[stage], (int32_t)(kb * kGemmBK), a_row);
                        ptx::tma_load_2d_pair(sa + Cfg::kABytes, &tmB, &full[stage], (int32_t)(kb * kGemmBK), b_row);
                    } else {
                        ptx::mbar_expect_tx(&full[stage], Cfg::kStageBytes);
                        ptx::tma_load_2d(sa, &tmA, &full[stage], (int32_t)(kb * kGemmBK), a_row);
                        ptx::tma_load_2d(sa + Cfg::kABytes, &tmB, &full[stage], (int32_t)(kb * kGemmBK), b_row);
                    }
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = ptx::umma_idesc_f16(kTileM, BN, BF16);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (uint32_t tile = unit; tile < ntiles; tile += n_units) {
                ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (uint32_t kb = 0; kb < nkb; kb++) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint64_t a_desc = ptx::smem_desc_sw128(sa);
                    const uint64_t b_desc = ptx::smem_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
                    for (uint32_t k = 0; k < kGemmBK / 16; k++) {
                        if constexpr (CG == 2) ptx::umma_f16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        else ptx::umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    if constexpr (CG == 2) ptx::umma_commit_pair(&empty[stage], 3);  // frees the stage in both CTAs
                    else ptx::umma_commit(&empty[stage]);
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
                if constexpr (CG == 2) ptx::umma_commit_pair(&tfull[acc], 3);
                else ptx::umma_commit(&tfull[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        const uint32_t quad = warp & 3;
        const uint32_t part = (uint32_t)(warp - 2) >> 2;                     // which slice of the columns this warp drains
        constexpr uint32_t kChunks = BN / 32, kParts = kGemmEpiWarps / 4;
        // direct-store epilogues split the columns evenly; the TMA-store epilogue works in 64-column units (128 columns per half)
        const uint32_t c_begin = Epilogue::kTmaStore ? min(part * 4u, kChunks) : part * kChunks / kParts;
        const uint32_t c_end = Epilogue::kTmaStore ? min((part + 1) * 4u, kChunks) : (part + 1) * kChunks / kParts;
        uint32_t acc = 0, acc_phase = 0, store_n = 0;
        auto release_acc = [&](uint32_t a) {  // hand the accumulator back to the (leader's) MMA issuer
            if constexpr (CG == 2) ptx::mbar_arrive_remote(&tempty[a], 0);
            else ptx::mbar_arrive(&tempty[a]);
        };
        // in-place residual (x += linear(...)): this warp's 32 rows x 64 columns of x are fetched one 64-column step AHEAD
        // (also across tiles) with coalesced 128-byte row segments, so the HBM latency hides behind a whole step
        uint4 resv[8];
        auto prefetch_residual = [&](uint32_t t, uint32_t c) {
            if constexpr (Epilogue::kTmaStore) {
                const __half *res = epi.staged_residual();
                if (res == nullptr || t >= ntiles) return;
                uint32_t pm, pn;
                gemm_tile_coords(shp, t, pm, pn);
                const uint32_t ldc = epi.ldc(), Mrows = epi.rows(), Ncols = epi.cols();
                const uint32_t prow = pm * kTileM + cta_rank * kGemmBM + quad * 32 + (lane >> 3), pcol = pn * BN + c * 32 + (lane & 7) * 8;
                const __half *src = res + (size_t)prow * ldc + pcol;
#pragma unroll
                for (uint32_t it = 0; it < 8; it++) {
                    resv[it] = make_uint4(0, 0, 0, 0);
                    if (prow + it * 4 < Mrows && pcol + 8 <= Ncols) resv[it] = *(const uint4 *)(src + (size_t)(it * 4) * ldc);
                }
            }
        };
        if (c_begin < c_end) prefetch_residual(unit, c_begin);
        for (uint32_t tile = unit; tile < ntiles; tile += n_units) {
            uint32_t mt, nt;
            gemm_tile_coords(shp, tile, mt, nt);
            ptx::mbar_wait(&tfull[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t row0 = mt * kTileM + cta_rank * kGemmBM;   // first C row of this CTA's half of the tile
            const uint32_t row = row0 + quad * 32 + lane;
            const uint32_t taddr = tmem_base + ((quad * 32) << 16) + acc * BN;
            epi.begin_tile(row, nt * BN);
            if constexpr (Epilogue::kTmaStore) {
                // fp16 tile -> per-warp SW128 staging slab (32 rows x 64 columns) -> TMA store.  The 32 lanes of a warp own 32
                // different rows, so direct stores would touch 32 cache lines per instruction; each warp instead runs its own
                // double-buffered store pipeline (no cross-warp barriers): a slab is reused only after the store issued two
                // chunks earlier has finished reading it.
                static_assert(!Epilogue::kTmaStore || (BN % 64 == 0 && BN <= 256 && kGemmEpiWarps == 8), "TMA-store epilogue: 64-column units, two column halves");
                uint8_t *cw = cstage + (warp - 2) * (2 * 32 * 128);
                if (c_begin >= c_end) {  // this half has no columns in a narrow tile: just release the accumulator
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) release_acc(acc);
                }
#pragma unroll 1
                for (uint32_t c = c_begin; c < c_end; c += 2) {
                    uint32_t v0[32], v1[32];
                    ptx::tmem_ld_32x32(taddr + c * 32, v0);
                    ptx::tmem_ld_32x32(taddr + c * 32 + 32, v1);
                    ptx::tmem_ld_wait();
                    if (c + 2 >= c_end) {  // last TMEM read of this tile: hand the accumulator back before the stores
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) release_acc(acc);
                    }
                    const uint32_t col0 = nt * BN + c * 32;
                    epi.compute(row, col0, v0);        // bias / activation / small residuals, in place (fp32 bit patterns)
                    epi.compute(row, col0 + 32, v1);
                    uint8_t *buf = cw + (store_n & 1) * (32 * 128);
                    store_n++;
                    if (lane == 0) ptx::tma_store_wait_read1();  // the store that used this slab two chunks ago is done reading
                    __syncwarp();
                    uint8_t *myrow = buf + lane * 128;
                    if (epi.staged_residual() != nullptr) {
                        // the prefetched rows of x go through the slab, then every lane picks up its own row
#pragma unroll
                        for (uint32_t it = 0; it < 8; it++) {
                            const uint32_t rr = it * 4 + (lane >> 3), ch = lane & 7;
                            *(uint4 *)(buf + rr * 128 + ((ch ^ (rr & 7)) << 4)) = resv[it];
                        }
                        if (c + 2 < c_end) prefetch_residual(tile, c + 2);
                        else prefetch_residual(tile + n_units, c_begin);
                        __syncwarp();
#pragma unroll
                        for (uint32_t q = 0; q < 8; q++) {
                            const uint4 u = *(const uint4 *)(myrow + ((q ^ (lane & 7)) << 4));
                            const __half2 *h = (const __half2 *)&u;
                            uint32_t *dst = q < 4 ? &v0[q * 8] : &v1[(q - 4) * 8];
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const float2 t = __half22float2(h[e]);
                                dst[2 * e] = __float_as_uint(__uint_as_float(dst[2 * e]) + t.x);
                                dst[2 * e + 1] = __float_as_uint(__uint_as_float(dst[2 * e + 1]) + t.y);
                            }
                        }
                        __syncwarp();
                    }
#pragma unroll
                    for (uint32_t q = 0; q < 8; q++) {
                        const uint32_t *src = q < 4 ? &v0[q * 8] : &v1[(q - 4) * 8];
                        uint4 u;
                        __half2 *h = (__half2 *)&u;
#pragma unroll
                        for (int e = 0; e < 4; e++) h[e] = __floats2half2_rn(__uint_as_float(src[2 * e]), __uint_as_float(src[2 * e + 1]));
                        *(uint4 *)(myrow + ((q ^ (lane & 7)) << 4)) = u;
                    }
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        ptx::tma_store_2d(&tmC, buf, (int32_t)col0, (int32_t)(row0 + quad * 32));
                        ptx::tma_store_commit();
                    }
                }
            } else {
#pragma unroll 1
                for (uint32_t c = c_begin; c < c_end; c++) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(taddr + c * 32, v);
                    ptx::tmem_ld_wait();
                    epi.columns(row, nt * BN + c * 32, v);
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) release_acc(acc);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    if constexpr (Epilogue::kTmaStore) {
        if (warp >= 2 && lane == 0) ptx::tma_store_wait_all();
    }
    ptx::tc_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync();  // neither CTA may exit (or free TMEM) while its partner can still signal it
    else __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        if constexpr (CG == 2) ptx::tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
        else ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

// host: encode a 2D K-major tensor map (rows x K elements of 2 bytes), box = 64 x box_rows, SWIZZLE_128B
int encode_tmap_2d(CUtensorMap *out, const void *base, uint64_t rows, uint64_t k, uint64_t row_stride_elems, uint32_t box_rows);

int encode_tmap_3d(CUtensorMap *out, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                   uint64_t stride2_bytes, uint32_t b0, uint32_t b2, int swizzle);

}  // namespace mse
