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
bool pdl_enabled();            // MSE_NO_PDL=1 switches programmatic dependent launch off (A/B measurements)

// Programmatic dependent launch (PDL).  A kernel launched with launch_pdl may start while its predecessor on the stream is still
// running: everything it does before pdl_wait() (index math, shared-memory setup, prefetch of WEIGHTS -- nothing the predecessor
// writes) overlaps the predecessor's tail; pdl_wait() returns once the predecessor has completed and its writes are visible.
// pdl_trigger() lets the NEXT kernel's CTAs be scheduled as soon as every CTA of this one has called it.  The batch-1 text tower is
// 194 dependent kernels of a few microseconds: the launch-to-launch gap and the cold start of each weight stream are what PDL hides.
// Measured (profiles/r03c_text_latency_pdl.md): letting the next kernel's CTAs in EARLY (pdl_trigger at kernel start) is a loss -- the
// co-resident CTAs of up to three kernels fight over shared memory and HBM (batch 1: 2.57 ms against 1.84 ms without PDL); with the
// implicit trigger at kernel exit the pre-staged launch still saves 0.45 ms of an eager forward pass (2.24 -> 1.79 ms) and 0.07 ms of
// the graph replay (1.84 -> 1.77 ms).
// Rule: a kernel is launched with launch_pdl only if it calls pdl_wait() before its first dependent access.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif


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
