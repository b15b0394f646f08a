// Shared host/device helpers of libmse_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <vector>
#include "../../include/mse_b200.h"

#define MSE_API extern "C" __attribute__((visibility("default")))

namespace mse {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define MSE_CUDA(expr)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            mse::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (_e == cudaErrorMemoryAllocation) ? MSE_ERR_OOM : MSE_ERR_CUDA;                 \
        }                                                                                          \
    } while (0)

#define MSE_CHECK(rc)            \
    do {                         \
        int _rc = (rc);          \
        if (_rc != MSE_OK) return _rc; \
    } while (0)

#define MSE_REQUIRE(cond, code, ...)    \
    do {                                \
        if (!(cond)) {                  \
            mse::set_error(__VA_ARGS__); \
            return (code);              \
        }                               \
    } while (0)

// launch check: cheap (no sync); errors surface as sticky errors at the next sync as well
#define MSE_LAUNCH_OK()                                                               \
    do {                                                                              \
        mse::count_launch();                                                          \
        cudaError_t _e = cudaGetLastError();                                          \
        if (_e != cudaSuccess) {                                                      \
            mse::set_error("%s:%d: launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return MSE_ERR_CUDA;                                                      \
        }                                                                             \
    } while (0)

int use_device(int device);   // cudaSetDevice + verifies compute capability 10.x

// cudaFuncSetAttribute applies to the current device only: a call site remembers which devices it has configured
// (handles are independent across devices, so one process may drive several)
struct PerDeviceOnce {
    std::atomic<uint64_t> mask{0};
    bool first(int device) {
        const uint64_t bit = 1ull << (device & 63);
        return !(mask.fetch_or(bit, std::memory_order_relaxed) & bit);
    }
};
int sm_count(int device);

// growable device buffer owned by a handle
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return MSE_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
            (void)cudaGetLastError();
            return MSE_ERR_OOM;
        }
        cap = bytes;
        return MSE_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return (T *)p; }
};

// ------------------------------------------------------------------ device helpers

// order-preserving map f32 -> u32 (ascending); -0.0 is folded onto +0.0 first
__host__ __device__ __forceinline__ uint32_t f32_ordered(float f) {
    f = f + 0.0f;
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float f32_from_ordered(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
// rank key: larger key ranks first under (score desc, id asc)
__host__ __device__ __forceinline__ uint64_t rank_key(float score, uint32_t id) {
    return ((uint64_t)f32_ordered(score) << 32) | (uint64_t)(~id);
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return f32_from_ordered((uint32_t)(k >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_id(uint64_t k) { return ~(uint32_t)k; }

// Rust `(x * 2^32) as i64` (vector.rs:408-411): truncate toward zero, saturate, NaN -> 0
__device__ __forceinline__ long long scale_dot_result(float x) {
    float v = x * 4294967296.0f;
    return __float2ll_rz(v);  // cvt.rzi.s64.f32 saturates and maps NaN to 0... see below
}

}  // namespace mse
