// fast_dot on the GPU, bit-identical to diskann/src/vector.rs:192-306.
//
// The reference keeps 4 AVX2 accumulators x 8 lanes = 32 fp32 partial sums; partial sum p owns the elements
// d with d % 32 == p and folds them in increasing d with one FMA each.  A warp is exactly 32 lanes, so lane p
// owns partial sum p: each lane walks the row with stride 32 (2-byte loads, 64 contiguous bytes per warp
// request) and the reference's reduction tree is replayed with shuffles:
//   A_j = p[j] + p[8+j],  B_j = p[16+j] + p[24+j]            (acc1+acc2, acc3+acc4)
//   hadd -> [A0+A1, A2+A3, B0+B1, B2+B3 | A4+A5, A6+A7, B4+B5, B6+B7];  lo + hi;  ((e0+e1)+e2)+e3
// fmaf / + on the GPU are IEEE-754 binary32 like x86 FMA / ADDPS, so the i64 result matches bit for bit.
#include "internal.h"
#include "fastdot.cuh"
#include <algorithm>

namespace mse {

__global__ void __launch_bounds__(256) k_fast_dot_batch(const __half *__restrict__ query, const __half *__restrict__ rows,
                                                        uint32_t d, const uint32_t *__restrict__ row_ids, uint64_t n_ids,
                                                        long long *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = warp; i < n_ids; i += nwarps) {
        const uint64_t r = row_ids ? row_ids[i] : i;
        float s = fast_dot_warp(query, rows + r * d, d, lane);
        if (lane == 0) out[i] = fast_dot_fix(s);
    }
}

}  // namespace mse

using namespace mse;

MSE_API int mse_fast_dot_batch(int device, const uint16_t *query_f16, const uint16_t *rows_f16, uint64_t n_rows, uint32_t d,
                               const uint32_t *row_ids, uint64_t n_ids, int64_t *scores) {
    MSE_CHECK(use_device(device));
    MSE_REQUIRE(d > 0 && d % 64 == 0, MSE_ERR_INVALID, "fast_dot_batch: d=%u must be a multiple of 64 (vector.rs:197)", d);
    MSE_REQUIRE(query_f16 && rows_f16 && scores, MSE_ERR_INVALID, "fast_dot_batch: NULL buffer");
    if (n_ids == 0) return MSE_OK;
    __half *dq = nullptr, *dr = nullptr;
    uint32_t *di = nullptr;
    long long *ds = nullptr;
    int rc = MSE_OK;
    do {
        if (cudaMalloc(&dq, d * 2) != cudaSuccess || cudaMalloc(&dr, n_rows * d * 2) != cudaSuccess ||
            cudaMalloc(&ds, n_ids * 8) != cudaSuccess || (row_ids && cudaMalloc(&di, n_ids * 4) != cudaSuccess)) {
            (void)cudaGetLastError();
            set_error("fast_dot_batch: device allocation failed");
            rc = MSE_ERR_OOM;
            break;
        }
        cudaMemcpy(dq, query_f16, d * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dr, rows_f16, n_rows * d * 2, cudaMemcpyHostToDevice);
        if (row_ids) cudaMemcpy(di, row_ids, n_ids * 4, cudaMemcpyHostToDevice);
        uint32_t blocks = (uint32_t)std::min<uint64_t>((n_ids + 7) / 8, (uint64_t)sm_count(device) * 8);
        k_fast_dot_batch<<<blocks, 256>>>(dq, dr, d, di, n_ids, ds);
        count_launch();
        cudaError_t e = cudaMemcpy(scores, ds, n_ids * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            set_error("fast_dot_batch: %s", cudaGetErrorString(e));
            rc = MSE_ERR_CUDA;
        }
    } while (0);
    cudaFree(dq); cudaFree(dr); cudaFree(di); cudaFree(ds);
    return rc;
}
