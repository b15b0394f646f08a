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

template <int BN>
struct GemmCfg {
    static constexpr int kStages = BN >= 256 ? 4 : (BN >= 192 ? 5 : 6);
    static constexpr uint32_t kABytes = kGemmBM * kGemmBK * 2;
    static constexpr uint32_t kBBytes = BN * kGemmBK * 2;
    static constexpr uint32_t kStageBytes = kABytes + kBBytes;
    static constexpr uint32_t kBarBytes = 256;
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // + slack for 1024-B alignment
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

template <int BN, int BF16, class Epilogue>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_gemm_tn(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmShape shp, Epilogue epi) {
    using Cfg = GemmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t *)(smem + Cfg::kStages * Cfg::kStageBytes);
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
            ptx::mbar_init(&tempty[a], kGemmEpiWarps);  // one arrive per epilogue warp
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                uint32_t mt, nt;
                gemm_tile_coords(shp, tile, mt, nt);
                for (uint32_t kb = 0; kb < nkb; kb++) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    ptx::mbar_expect_tx(&full[stage], Cfg::kStageBytes);
                    uint8_t *sa = smem + stage * Cfg::kStageBytes;
                    ptx::tma_load_2d(sa, &tmA, &full[stage], (int32_t)(kb * kGemmBK), shp.a_row0 + (int32_t)(mt * kGemmBM));
                    ptx::tma_load_2d(sa + Cfg::kABytes, &tmB, &full[stage], (int32_t)(kb * kGemmBK), shp.b_row0 + (int32_t)(nt * BN));
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::umma_idesc_f16(kGemmBM, BN, BF16);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
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
                    for (uint32_t k = 0; k < kGemmBK / 16; k++)
                        ptx::umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(&empty[stage]);
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tfull[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        const uint32_t quad = warp & 3;
        const uint32_t part = (uint32_t)(warp - 2) >> 2;                     // which slice of the columns this warp drains
        constexpr uint32_t kChunks = BN / 32, kParts = kGemmEpiWarps / 4;
        const uint32_t c_begin = part * kChunks / kParts, c_end = (part + 1) * kChunks / kParts;
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            uint32_t mt, nt;
            gemm_tile_coords(shp, tile, mt, nt);
            ptx::mbar_wait(&tfull[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t row = mt * kGemmBM + quad * 32 + lane;
            const uint32_t taddr = tmem_base + ((quad * 32) << 16) + acc * BN;
            epi.begin_tile(row, nt * BN);
#pragma unroll 1
            for (uint32_t c = c_begin; c < c_end; c++) {
                uint32_t v[32];
                ptx::tmem_ld_32x32(taddr + c * 32, v);
                ptx::tmem_ld_wait();
                epi.columns(row, nt * BN + c * 32, v);
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

// host: encode a 2D K-major tensor map (rows x K elements of 2 bytes), box = 64 x box_rows, SWIZZLE_128B
int encode_tmap_2d(CUtensorMap *out, const void *base, uint64_t rows, uint64_t k, uint64_t row_stride_elems, uint32_t box_rows);

int encode_tmap_3d(CUtensorMap *out, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                   uint64_t stride2_bytes, uint32_t b0, uint32_t b2, int swizzle);

}  // namespace mse
