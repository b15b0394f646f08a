// Dense GEMM entry points built on the tcgen05 mainloop (gemm_sm100.cuh):  C[M,N] = A[M,K] * B[N,K]^T.
// Used by the encoder towers (linear layers: activations x weight[out,in]^T) and the OPQ rotation
// (diskann/src/vector.rs:320-329), and exposed on the C ABI so the tensor path can be validated on its own.
#include "internal.h"
#include "gemm_sm100.cuh"
#include "gemm_epilogues.cuh"
#include "gemm_skinny.cuh"
#include <algorithm>
#include <stdlib.h>

namespace mse {

template <int BN, class Epi, int CG = 1>
static int launch_gemm(int device, const void *dA, const void *dB, uint32_t M, uint32_t N, uint32_t K, uint32_t lda, uint32_t ldb,
                       Epi epi, cudaStream_t st) {
    constexpr uint32_t kSmem = gemm_smem_bytes<BN, Epi, CG>();
    auto kern = k_gemm_tn<BN, 0, Epi, CG>;
    static PerDeviceOnce once;
    if (once.first(device)) MSE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    CUtensorMap tmA, tmB, tmC;
    MSE_CHECK(encode_tmap_2d(&tmA, dA, M, K, lda, kGemmBM));
    MSE_CHECK(encode_tmap_2d(&tmB, dB, N, K, ldb, BN / CG));
    if (Epi::kTmaStore) MSE_CHECK(encode_tmap_2d(&tmC, epi.o.c16, M, N, epi.o.ldc, 32));  // per-warp 32-row slabs
    else tmC = tmA;
    GemmShape shp;
    shp.M = M; shp.N = N; shp.K = K;
    shp.tiles_m = (M + kGemmBM * CG - 1) / (kGemmBM * CG);
    shp.tiles_n = (N + BN - 1) / BN;
    shp.m_fastest = 0;
    shp.a_row0 = 0; shp.b_row0 = 0;
    const uint32_t ntiles = shp.tiles_m * shp.tiles_n;
    const uint32_t units = std::min<uint32_t>(ntiles, (uint32_t)sm_count(device) / CG);
    if (CG == 1) {
        kern<<<units, kGemmThreads, kSmem, st>>>(tmA, tmB, tmC, shp, epi);
    } else {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(units * CG);
        cfg.blockDim = dim3(kGemmThreads);
        cfg.dynamicSmemBytes = kSmem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MSE_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, shp, epi));
    }
    MSE_LAUNCH_OK();
    return MSE_OK;
}

static int force_bn = 0;  // profiling only (mse_debug_gemm)

// small M (text tower at small batches): stream the weights once through every SM (gemm_skinny.cuh)
static constexpr uint32_t kSkinnyMaxM = 128;   // measured (tools/text_latency.py): 2.80 -> 2.14 ms at 64 rows, 2.85 -> 2.25 ms at 128; slower than the tcgen05 tiles at 512
static constexpr uint32_t kSkinnyMaxTiles = 256;   // arrival counters a caller provides (GemmOut::splitk_cnt)
template <int BN>
static int launch_skinny(int out_device, const __half *dA, const __half *dB, uint32_t M, uint32_t N, uint32_t K, uint32_t lda, uint32_t ldb, const GemmOut &out,
                         cudaStream_t st) {
    auto kern = skinny::k_gemm_skinny<BN>;
    static PerDeviceOnce once;
    if (once.first(out_device)) MSE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)skinny::smem_bytes<BN>()));
    const uint32_t slices = (N + BN - 1) / BN, mt = (M + skinny::kBM - 1) / skinny::kBM, nk = (K + skinny::kBK - 1) / skinny::kBK;
    uint32_t splits = 1;
    if (out.splitk_ws && out.splitk_cnt && slices * mt <= kSkinnyMaxTiles) {
        splits = std::max<uint32_t>(1, std::min<uint32_t>(nk, (uint32_t)sm_count(out_device) / (slices * mt)));
        while (splits > 1 && (size_t)splits * mt * skinny::kBM * slices * BN > out.splitk_ws_floats) splits--;
    }
    launch_pdl(kern, dim3(slices, mt, splits), skinny::kThreads, skinny::smem_bytes<BN>(), st, dA, dB, M, N, K, lda, ldb, out);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

// M <= 64: activation panel resident in shared memory, the pipeline streams weights only (gemm_skinny.cuh, namespace pr)
template <int BN>
static int launch_skinny_pr(int device, const __half *dA, const __half *dB, uint32_t M, uint32_t N, uint32_t K, uint32_t lda, uint32_t ldb, const GemmOut &out,
                            cudaStream_t st) {
    auto kern = skinny::pr::k_gemm_skinny_pr<BN>;
    static PerDeviceOnce once;
    if (once.first(device)) MSE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)skinny::pr::smem_bytes<BN>()));
    launch_pdl(kern, (N + BN - 1) / BN, skinny::pr::kThreads, skinny::pr::smem_bytes<BN>(), st, dA, dB, M, N, K, lda, ldb, out);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

int gemm_f16_tn_dev(int device, const __half *dA, const __half *dB, uint32_t M, uint32_t N, uint32_t K, uint32_t lda, uint32_t ldb,
                    const GemmOut &out, cudaStream_t st) {
    MSE_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, MSE_ERR_UNSUPPORTED, "gemm: K, lda, ldb must be multiples of 8");
    MSE_REQUIRE(M > 0 && N > 0 && K > 0, MSE_ERR_INVALID, "gemm: empty shape");
    // Panel-resident variant: measured NO faster than the ring-buffer kernel below (tools/gemm_probe.py, 64 rows: QKV 12.4 vs 10.9 us, out-proj
    // 10.3 vs 10.3, fc1 12.4 vs 12.4, fc2 30.8 vs 22.6) -- at these sizes a layer is ~4 us of SM-active time inside a ~10 us launch-to-drain
    // envelope (ncu: profiles/r01p_skinny_gemm_ncu_full.md), so what is left for batch-1 latency is the NUMBER of kernels (194 per forward),
    // not their insides.  Kept behind MSE_GEMM_PANEL=1 for experiments.
    // a folded LayerNorm lives in LinearEpilogueT<true>::compute only: refuse every other route rather than drop it silently
    const bool tma_store_route = out.c16 && !out.c32 && out.ldc % 8 == 0 && ((uintptr_t)out.c16 & 15) == 0 && !out.debug_no_store && force_bn != 128 &&
                                 force_bn != 192 && getenv("MSE_GEMM_DIRECT_STORE") == nullptr;
    MSE_REQUIRE(!out.ln_stats || (out.ln_cs && out.bias && M > kSkinnyMaxM && tma_store_route), MSE_ERR_UNSUPPORTED,
                "gemm: a folded LayerNorm needs the tcgen05 TMA-store path (M > %u, fp16 output with 16-byte rows)", kSkinnyMaxM);
    static const bool use_pr = getenv("MSE_GEMM_PANEL") != nullptr;
    if (use_pr && M <= (uint32_t)skinny::pr::kBM && force_bn == 0 && ((uintptr_t)dA & 15) == 0 && ((uintptr_t)dB & 15) == 0 && getenv("MSE_GEMM_NO_SKINNY") == nullptr) {
        GemmOut o = out;
        o.res_in_place = 0;
        o.splitk_ws = nullptr;
        // slice width: ~one CTA per SM (1152 columns -> 144 slices of 8; 3456 -> 108 and 4304 -> 135 slices of 32)
        return N <= 2048 ? launch_skinny_pr<8>(device, dA, dB, M, N, K, lda, ldb, o, st) : launch_skinny_pr<32>(device, dA, dB, M, N, K, lda, ldb, o, st);
    }
    if (M <= kSkinnyMaxM && force_bn == 0 && ((uintptr_t)dA & 15) == 0 && ((uintptr_t)dB & 15) == 0 && getenv("MSE_GEMM_NO_SKINNY") == nullptr) {
        GemmOut o = out;
        o.res_in_place = 0;
        { const char *t = getenv("MSE_PDL_TRIGGER"); o.pdl_trigger_at = t ? atoi(t) : 2; }   // tuning aid; measured at batch 1: 0 -> 2.57 ms, 1 -> 1.93 ms, 2 -> 1.77 ms (no PDL: 1.84 ms)
        // Wide slices + split over K (few readers of the activation panel, every SM busy) -- measured SLOWER than one narrow slice per
        // CTA (text tower at batch 1: 4.0 ms vs 2.12 ms, profiles/r01p_skinny_gemm_ncu_full.md): the last CTA of a tile serialises the
        // fixed-order reduction of up to 16 partial tiles behind a device-wide fence.  Kept for experiments only.
        if (o.splitk_ws && o.splitk_cnt && getenv("MSE_GEMM_SPLITK") != nullptr) {
            return (N >= 2048 || K > 2048) ? launch_skinny<128>(device, dA, dB, M, N, K, lda, ldb, o, st)
                                           : launch_skinny<64>(device, dA, dB, M, N, K, lda, ldb, o, st);
        }
        o.splitk_ws = nullptr;
        // no scratch from the caller: 16-column slices when N is small, so that at least ~70 SMs pull on the weights
        return N <= 2048 ? launch_skinny<16>(device, dA, dB, M, N, K, lda, ldb, o, st) : launch_skinny<32>(device, dA, dB, M, N, K, lda, ldb, o, st);
    }
    // fp16 output with 16-byte aligned rows: 256-wide tiles, staged through shared memory and written by TMA
    // (in-place residual -> TMA reduce-add).  Measured (tools/gemm_ablation.py): direct per-thread stores cost 25-50 %.
    if (out.c16 && !out.c32 && out.ldc % 8 == 0 && ((uintptr_t)out.c16 & 15) == 0 && !out.debug_no_store && force_bn != 128 && force_bn != 192 &&
        getenv("MSE_GEMM_DIRECT_STORE") == nullptr) {
        LinearEpilogueT<true> e2;
        e2.o = out;
        e2.o.res_in_place = (out.res == out.c16 && out.res_mod == 0) ? 1 : 0;
        e2.M = M;
        e2.N = N;
        // 192-wide tiles when they divide N with less padding (1152 = 6 x 192, 3456 = 18 x 192): padded columns are wasted
        // MMAs, and the towers run against the board's power cap
        const uint32_t w256 = (N + 255) / 256 * 256 - N, w192 = (N + 191) / 192 * 192 - N;
        const char *bn_env = getenv("MSE_GEMM_BN");  // profiling only
        const bool allow192 = !(bn_env && atoi(bn_env) == 256);
        const char *cg_env = getenv("MSE_GEMM_CG");  // profiling only: 1 forces single-CTA tiles
        // mid-size M (text tower at batch 3-32, MAP head): the 256 x 256 pair tiles leave most of the chip idle (M = 512, N = 1152 is 10
        // pairs on 74); 128 x 128 single-CTA tiles give 4x the units
        {
            const uint32_t sms = (uint32_t)sm_count(device);
            const uint32_t pairs = ((M + 255) / 256) * ((N + 255) / 256), t128 = ((M + 127) / 128) * ((N + 127) / 128);
            if (pairs * 2 * 10 < sms * 6 && t128 > pairs * 2 && !(bn_env || cg_env) && force_bn == 0)
                return launch_gemm<128, LinearEpilogueT<true>, 1>(device, dA, dB, M, N, K, lda, ldb, e2, st);
        }
        if (!(cg_env && atoi(cg_env) == 1)) {
            // 256 x 256 pair tiles even where 192 would divide N exactly: measured 146.5 ms vs 157.4 ms of GEMM time per tower step
            if (bn_env && atoi(bn_env) == 192 && w192 < w256) return launch_gemm<192, LinearEpilogueT<true>, 2>(device, dA, dB, M, N, K, lda, ldb, e2, st);
            return launch_gemm<256, LinearEpilogueT<true>, 2>(device, dA, dB, M, N, K, lda, ldb, e2, st);
        }
        if (w192 < w256 && force_bn != 256 && allow192) return launch_gemm<192>(device, dA, dB, M, N, K, lda, ldb, e2, st);
        return launch_gemm<256>(device, dA, dB, M, N, K, lda, ldb, e2, st);
    }
    LinearEpilogue epi;
    epi.o = out;
    epi.o.res_in_place = 0;
    epi.M = M;
    epi.N = N;
    if (force_bn == 256) return launch_gemm<256>(device, dA, dB, M, N, K, lda, ldb, epi, st);
    if (force_bn == 192) return launch_gemm<192>(device, dA, dB, M, N, K, lda, ldb, epi, st);
    if (force_bn == 128) return launch_gemm<128>(device, dA, dB, M, N, K, lda, ldb, epi, st);
    // tile width: least padded columns wins, ties go to the wider tile (fewer A re-reads per flop)
    auto waste = [&](uint32_t bn) { return (N + bn - 1) / bn * bn - N; };
    uint32_t best = 256;
    if (waste(192) < waste(best)) best = 192;
    if (waste(128) < waste(best)) best = 128;
    if (best == 256) return launch_gemm<256>(device, dA, dB, M, N, K, lda, ldb, epi, st);
    if (best == 192) return launch_gemm<192>(device, dA, dB, M, N, K, lda, ldb, epi, st);
    return launch_gemm<128>(device, dA, dB, M, N, K, lda, ldb, epi, st);
}

}  // namespace mse

using namespace mse;

// Test/diagnostic entry: host pointers, fp16 in, fp32 out.  C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) with optional GELU.
MSE_API int mse_gemm_f16_tn(int device, const uint16_t *a, const uint16_t *b, uint32_t M, uint32_t N, uint32_t K, const float *bias,
                            int act, float *c) {
    MSE_CHECK(use_device(device));
    MSE_REQUIRE(a && b && c, MSE_ERR_INVALID, "gemm: NULL buffer");
    __half *dA = nullptr, *dB = nullptr;
    float *dC = nullptr, *dbias = nullptr;
    int rc = MSE_OK;
    do {
        if (cudaMalloc(&dA, (size_t)M * K * 2) != cudaSuccess || cudaMalloc(&dB, (size_t)N * K * 2) != cudaSuccess ||
            cudaMalloc(&dC, (size_t)M * N * 4) != cudaSuccess || (bias && cudaMalloc(&dbias, (size_t)N * 4) != cudaSuccess)) {
            (void)cudaGetLastError();
            set_error("gemm: device allocation failed");
            rc = MSE_ERR_OOM;
            break;
        }
        cudaMemcpy(dA, a, (size_t)M * K * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, b, (size_t)N * K * 2, cudaMemcpyHostToDevice);
        if (bias) cudaMemcpy(dbias, bias, (size_t)N * 4, cudaMemcpyHostToDevice);
        cudaMemset(dC, 0xff, (size_t)M * N * 4);
        GemmOut o{};
        o.c32 = dC;
        o.ldc = N;
        o.bias = dbias;
        o.act = act;
        rc = gemm_f16_tn_dev(device, dA, dB, M, N, K, K, K, o, nullptr);
        if (rc != MSE_OK) break;
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            set_error("gemm: kernel failed: %s", cudaGetErrorString(e));
            rc = MSE_ERR_CUDA;
            break;
        }
        cudaMemcpy(c, dC, (size_t)M * N * 4, cudaMemcpyDeviceToHost);
    } while (0);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dbias);
    return rc;
}

// Profiling aid (not part of the reference surface): average time (ms) of one fp16-output linear-layer GEMM on random data.
// bn: 0 auto / 128 / 192 / 256; mode bit 0: skip the global stores, bit 1: + bias, bit 2: + erf GELU, bit 3: + residual
MSE_API int mse_debug_gemm(int device, uint32_t M, uint32_t N, uint32_t K, int bn, int mode, int iters, float *ms_out) {
    MSE_CHECK(use_device(device));
    __half *dA = nullptr, *dB = nullptr, *dC = nullptr, *dR = nullptr;
    float *dbias = nullptr;
    MSE_CUDA(cudaMalloc(&dA, (size_t)M * K * 2));
    MSE_CUDA(cudaMalloc(&dB, (size_t)N * K * 2));
    MSE_CUDA(cudaMalloc(&dC, (size_t)M * N * 2));
    MSE_CUDA(cudaMalloc(&dR, (size_t)M * N * 2));
    MSE_CUDA(cudaMalloc(&dbias, (size_t)N * 4));
    cudaMemset(dA, 0x11, (size_t)M * K * 2);
    cudaMemset(dB, 0x11, (size_t)N * K * 2);
    cudaMemset(dR, 0, (size_t)M * N * 2);
    cudaMemset(dbias, 0, (size_t)N * 4);
    GemmOut o{};
    o.c16 = dC; o.ldc = N;
    o.debug_no_store = mode & 1;
    if (mode & 2) o.bias = dbias;
    if (mode & 4) o.act = ACT_GELU_ERF;
    if (mode & 8) o.res = dR;
    force_bn = bn;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int rc = MSE_OK;
    for (int i = 0; i < 2 && rc == MSE_OK; i++) rc = gemm_f16_tn_dev(device, dA, dB, M, N, K, K, K, o, nullptr);
    cudaEventRecord(e0);
    for (int i = 0; i < iters && rc == MSE_OK; i++) rc = gemm_f16_tn_dev(device, dA, dB, M, N, K, K, K, o, nullptr);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    force_bn = 0;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dR); cudaFree(dbias);
    MSE_CHECK(rc);
    MSE_REQUIRE(err == cudaSuccess, MSE_ERR_CUDA, "debug_gemm: %s", cudaGetErrorString(err));
    *ms_out = ms / iters;
    return MSE_OK;
}
