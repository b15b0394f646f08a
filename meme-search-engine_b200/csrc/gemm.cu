// Dense GEMM entry points built on the tcgen05 mainloop (gemm_sm100.cuh):  C[M,N] = A[M,K] * B[N,K]^T.
// Used by the encoder towers (linear layers: activations x weight[out,in]^T) and the OPQ rotation
// (diskann/src/vector.rs:320-329), and exposed on the C ABI so the tensor path can be validated on its own.
#include "internal.h"
#include "gemm_sm100.cuh"
#include "gemm_epilogues.cuh"
#include <algorithm>

namespace mse {

template <int BN, class Epi>
static int launch_gemm(int device, const void *dA, const void *dB, uint32_t M, uint32_t N, uint32_t K, uint32_t lda, uint32_t ldb,
                       Epi epi, cudaStream_t st) {
    using Cfg = GemmCfg<BN>;
    auto kern = k_gemm_tn<BN, 0, Epi>;
    static bool attr_done = false;
    if (!attr_done) {
        MSE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes));
        attr_done = true;
    }
    CUtensorMap tmA, tmB;
    MSE_CHECK(encode_tmap_2d(&tmA, dA, M, K, lda, kGemmBM));
    MSE_CHECK(encode_tmap_2d(&tmB, dB, N, K, ldb, BN));
    GemmShape shp;
    shp.M = M; shp.N = N; shp.K = K;
    shp.tiles_m = (M + kGemmBM - 1) / kGemmBM;
    shp.tiles_n = (N + BN - 1) / BN;
    shp.m_fastest = 0;
    shp.a_row0 = 0; shp.b_row0 = 0;
    const uint32_t ntiles = shp.tiles_m * shp.tiles_n;
    const uint32_t grid = std::min<uint32_t>(ntiles, (uint32_t)sm_count(device));
    kern<<<grid, kGemmThreads, Cfg::kSmemBytes, st>>>(tmA, tmB, shp, epi);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

int gemm_f16_tn_dev(int device, const __half *dA, const __half *dB, uint32_t M, uint32_t N, uint32_t K, uint32_t lda, uint32_t ldb,
                    const GemmOut &out, cudaStream_t st) {
    MSE_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, MSE_ERR_UNSUPPORTED, "gemm: K, lda, ldb must be multiples of 8");
    MSE_REQUIRE(M > 0 && N > 0 && K > 0, MSE_ERR_INVALID, "gemm: empty shape");
    LinearEpilogue epi;
    epi.o = out;
    epi.M = M;
    epi.N = N;
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
