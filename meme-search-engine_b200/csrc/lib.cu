// Library-wide state: error string, device checks, launch counter.
#include "common.cuh"
#include <mutex>
#include <stdlib.h>

namespace mse {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

bool pdl_enabled() {
    const char *e = getenv("MSE_NO_PDL");   // read per call: a tuning switch, not a hot path
    return !(e && atoi(e) != 0);
}

static std::mutex g_dev_mu;
static int g_dev_ok[64];   // 0 unknown, 1 ok, -1 unusable
static int g_dev_sms[64];

int use_device(int device) {
    if (device < 0 || device >= 64) {
        set_error("device %d out of range", device);
        return MSE_ERR_INVALID;
    }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("cudaSetDevice(%d) -> %s (this library has no CPU fallback)", device, cudaGetErrorString(e));
        return MSE_ERR_CUDA;
    }
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (g_dev_ok[device] == 0) {
        cudaDeviceProp p;
        e = cudaGetDeviceProperties(&p, device);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            set_error("cudaGetDeviceProperties(%d) -> %s", device, cudaGetErrorString(e));
            return MSE_ERR_CUDA;
        }
        g_dev_sms[device] = p.multiProcessorCount;
        g_dev_ok[device] = (p.major == 10) ? 1 : -1;
        if (g_dev_ok[device] < 0) set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
    }
    if (g_dev_ok[device] < 0) {
        set_error("device %d is not an sm_100 part; this library is built for sm_100a only", device);
        return MSE_ERR_CUDA;
    }
    return MSE_OK;
}

int sm_count(int device) { return g_dev_sms[device] > 0 ? g_dev_sms[device] : 148; }

}  // namespace mse

MSE_API const char *mse_last_error(void) { return mse::t_err; }
MSE_API uint64_t mse_launch_count(void) { return mse::g_launches.load(); }

MSE_API int mse_device_info(int device, char *json_out, size_t cap) {
    MSE_CHECK(mse::use_device(device));
    cudaDeviceProp p;
    MSE_CUDA(cudaGetDeviceProperties(&p, device));
    size_t free_b = 0, total_b = 0;
    MSE_CUDA(cudaMemGetInfo(&free_b, &total_b));
    snprintf(json_out, cap,
             "{\"name\":\"%s\",\"sm\":%d,\"cc\":\"%d.%d\",\"hbm_total\":%zu,\"hbm_free\":%zu,\"smem_optin\":%zu,\"l2\":%d}",
             p.name, p.multiProcessorCount, p.major, p.minor, total_b, free_b, (size_t)p.sharedMemPerBlockOptin, p.l2CacheSize);
    return MSE_OK;
}
