// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links only cudart).
#include "common.cuh"
#include "gemm_sm100.cuh"

namespace mse {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

int encode_tmap_2d(CUtensorMap *out, const void *base, uint64_t rows, uint64_t k, uint64_t row_stride_elems, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode();
    MSE_REQUIRE(enc != nullptr, MSE_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    MSE_REQUIRE(((uintptr_t)base & 15) == 0 && (row_stride_elems * 2) % 16 == 0, MSE_ERR_INVALID,
                "tensor map: base/stride must be 16-byte aligned (stride %llu elems)", (unsigned long long)row_stride_elems);
    MSE_REQUIRE(box_rows >= 1 && box_rows <= 256, MSE_ERR_INVALID, "tensor map: box_rows=%u", box_rows);
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_stride_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t)kGemmBK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSE_REQUIRE(r == CUDA_SUCCESS, MSE_ERR_CUDA, "cuTensorMapEncodeTiled failed: CUresult %d (rows=%llu k=%llu)", (int)r,
                (unsigned long long)rows, (unsigned long long)k);
    return MSE_OK;
}

// 3-D map over a [d2][d1][d0] fp16 tensor (d0 contiguous); box {b0, 1, b2}; swizzle: 128 -> SWIZZLE_128B, 32 -> SWIZZLE_32B
int encode_tmap_3d(CUtensorMap *out, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                   uint64_t stride2_bytes, uint32_t b0, uint32_t b2, int swizzle) {
    PFN_encodeTiled enc = get_encode();
    MSE_REQUIRE(enc != nullptr, MSE_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    MSE_REQUIRE(((uintptr_t)base & 15) == 0 && stride1_bytes % 16 == 0 && stride2_bytes % 16 == 0, MSE_ERR_INVALID,
                "tensor map: base/strides must be 16-byte aligned");
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)stride1_bytes, (cuuint64_t)stride2_bytes};
    cuuint32_t box[3] = {b0, 1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSE_REQUIRE(r == CUDA_SUCCESS, MSE_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed: CUresult %d", (int)r);
    return MSE_OK;
}

}  // namespace mse
