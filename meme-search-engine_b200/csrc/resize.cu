// resize_for_embed_sync on the GPU (src/common.rs:31-54): an RGB8 image of any size -> image_size x image_size RGB8 by a separable
// convolution, Hamming window when both dimensions shrink, Lanczos3 otherwise (:43-44) -- the step the ingest client runs on the CPU
// before it ships a BMP to the embed service.  Here the decoded pixels go to the device once and the tower reads the result in place.
//
// The reference calls the fast_image_resize crate (v5, U8x3 convolution), which is not available offline; like Pillow's
// ImagingResample -- the code it descends from -- it is a horizontal pass followed by a vertical pass, each with per-output-pixel
// windows of normalised filter weights in fixed point and a u8 intermediate image.  This file follows Pillow's definition exactly
// (Resample.c: precompute_coeffs, normalize_coeffs_8bpc with PRECISION_BITS = 22, ImagingResampleHorizontal_8bpc / Vertical_8bpc), because
// Pillow is what can be executed here as the oracle (tests/test_resize.py: bit-exact against PIL.Image.resize); against the crate the
// result may differ by one code value where its i16 coefficients round differently -- unpinned.
// Coefficients are computed on the host in double precision (they depend on the two sizes only); the passes are gather kernels,
// one thread per output byte triple.
#include "internal.h"
#include <math.h>
#include <vector>

namespace mse {

static constexpr int kPrecisionBits = 32 - 8 - 2;

static double hamming_filter(double x) {
    if (x < 0.0) x = -x;
    if (x == 0.0) return 1.0;
    if (x >= 1.0) return 0.0;
    x = x * M_PI;
    return sin(x) / x * ((double)0.54f + (double)0.46f * cos(x));   // Pillow writes the constants as floats
}
static double sinc_filter(double x) {
    if (x == 0.0) return 1.0;
    x = x * M_PI;
    return sin(x) / x;
}
static double lanczos_filter(double x) {
    if (-3.0 <= x && x < 3.0) return sinc_filter(x) * sinc_filter(x / 3);
    return 0.0;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the whole input range [0, in_size)
static int precompute(uint32_t in_size, uint32_t out_size, bool lanczos, std::vector<int32_t> &bounds, std::vector<int32_t> &kk) {
    const double support_f = lanczos ? 3.0 : 1.0;
    double (*filter)(double) = lanczos ? lanczos_filter : hamming_filter;
    const double scale = (double)in_size / out_size;
    double filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = support_f * filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    bounds.assign((size_t)out_size * 2, 0);
    kk.assign((size_t)out_size * ksize, 0);
    std::vector<double> k(ksize);
    for (uint32_t xx = 0; xx < out_size; xx++) {
        const double center = (xx + 0.5) * scale;
        double ww = 0.0;
        const double ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > (int)in_size) xmax = (int)in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; x++) {
            const double w = filter((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; x++)
            if (ww != 0.0) k[x] /= ww;
        for (int x = 0; x < xmax; x++)
            kk[(size_t)xx * ksize + x] = k[x] < 0 ? (int32_t)(-0.5 + k[x] * (1 << kPrecisionBits)) : (int32_t)(0.5 + k[x] * (1 << kPrecisionBits));
        bounds[2 * xx] = xmin;
        bounds[2 * xx + 1] = xmax;
    }
    return ksize;
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= kPrecisionBits;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// out[y][xx][c] = clip8(round + sum_x in[y][xmin + x][c] * k[xx][x]); one thread per output pixel (3 channels)
__global__ void __launch_bounds__(256) k_resample_h(const uint8_t *__restrict__ in, uint32_t in_w, uint32_t rows, uint8_t *__restrict__ out, uint32_t out_w,
                                                    const int32_t *__restrict__ bounds, const int32_t *__restrict__ kk, int ksize) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * out_w) return;
    const uint32_t y = i / out_w, xx = i % out_w;
    const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
    const int32_t *k = kk + (size_t)xx * ksize;
    const uint8_t *row = in + ((size_t)y * in_w + xmin) * 3;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < xmax; x++) {
        const int w = k[x];
        s0 += row[3 * x] * w;
        s1 += row[3 * x + 1] * w;
        s2 += row[3 * x + 2] * w;
    }
    uint8_t *o = out + (size_t)i * 3;
    o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
}

// out[yy][x][c] = clip8(round + sum_y in[ymin + y][x][c] * k[yy][y])
__global__ void __launch_bounds__(256) k_resample_v(const uint8_t *__restrict__ in, uint32_t w, uint8_t *__restrict__ out, uint32_t out_h,
                                                    const int32_t *__restrict__ bounds, const int32_t *__restrict__ kk, int ksize) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= out_h * w) return;
    const uint32_t yy = i / w, x = i % w;
    const int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
    const int32_t *k = kk + (size_t)yy * ksize;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int y = 0; y < ymax; y++) {
        const uint8_t *p = in + ((size_t)(ymin + y) * w + x) * 3;
        const int wgt = k[y];
        s0 += p[0] * wgt;
        s1 += p[1] * wgt;
        s2 += p[2] * wgt;
    }
    uint8_t *o = out + (size_t)i * 3;
    o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
}

// host RGB8 [h][w][3] -> device RGB8 [out_h][out_w][3] at d_out; filter: 0 = the reference's rule, 1 = Hamming, 2 = Lanczos3
int resize_rgb_to_device(ResizeWork &wk, const uint8_t *rgb, uint32_t w, uint32_t h, uint32_t out_w, uint32_t out_h, int filter, uint8_t *d_out,
                         cudaStream_t st) {
    MSE_REQUIRE(rgb && w >= 1 && h >= 1 && out_w >= 1 && out_h >= 1 && w <= 32768 && h <= 32768, MSE_ERR_INVALID, "resize: bad size %ux%u -> %ux%u", w, h, out_w, out_h);
    const bool lanczos = filter == 2 || (filter == 0 && !(w > out_w && h > out_h));     // common.rs:43-44
    MSE_CHECK(wk.src.ensure((size_t)w * h * 3));
    MSE_CUDA(cudaMemcpyAsync(wk.src.p, rgb, (size_t)w * h * 3, cudaMemcpyHostToDevice, st));
    const uint8_t *cur = wk.src.as<uint8_t>();
    std::vector<int32_t> bounds, kk;
    const bool need_h = out_w != w, need_v = out_h != h;
    if (!need_h && !need_v) {
        MSE_CUDA(cudaMemcpyAsync(d_out, cur, (size_t)w * h * 3, cudaMemcpyDeviceToDevice, st));
        return MSE_OK;
    }
    if (need_h) {
        const int ksize = precompute(w, out_w, lanczos, bounds, kk);
        MSE_CHECK(wk.bh.ensure(bounds.size() * 4));
        MSE_CHECK(wk.kh.ensure(kk.size() * 4));
        MSE_CUDA(cudaMemcpyAsync(wk.bh.p, bounds.data(), bounds.size() * 4, cudaMemcpyHostToDevice, st));
        MSE_CUDA(cudaMemcpyAsync(wk.kh.p, kk.data(), kk.size() * 4, cudaMemcpyHostToDevice, st));
        MSE_CUDA(cudaStreamSynchronize(st));                     // the host vectors are reused below
        uint8_t *dst = need_v ? nullptr : d_out;
        if (need_v) {
            MSE_CHECK(wk.tmp.ensure((size_t)out_w * h * 3));
            dst = wk.tmp.as<uint8_t>();
        }
        k_resample_h<<<(h * out_w + 255) / 256, 256, 0, st>>>(cur, w, h, dst, out_w, wk.bh.as<int32_t>(), wk.kh.as<int32_t>(), ksize);
        MSE_LAUNCH_OK();
        cur = dst;
    }
    if (need_v) {
        const int ksize = precompute(h, out_h, lanczos, bounds, kk);
        MSE_CHECK(wk.bv.ensure(bounds.size() * 4));
        MSE_CHECK(wk.kv.ensure(kk.size() * 4));
        MSE_CUDA(cudaMemcpyAsync(wk.bv.p, bounds.data(), bounds.size() * 4, cudaMemcpyHostToDevice, st));
        MSE_CUDA(cudaMemcpyAsync(wk.kv.p, kk.data(), kk.size() * 4, cudaMemcpyHostToDevice, st));
        MSE_CUDA(cudaStreamSynchronize(st));
        k_resample_v<<<(out_h * out_w + 255) / 256, 256, 0, st>>>(cur, out_w, d_out, out_h, wk.bv.as<int32_t>(), wk.kv.as<int32_t>(), ksize);
        MSE_LAUNCH_OK();
    }
    return MSE_OK;
}

}  // namespace mse

using namespace mse;

MSE_API int mse_resize_rgb_u8(int device, const uint8_t *rgb, uint32_t w, uint32_t h, uint32_t out_w, uint32_t out_h, int filter, uint8_t *out) {
    MSE_REQUIRE(out != nullptr && filter >= 0 && filter <= 2, MSE_ERR_INVALID, "resize_rgb_u8: bad argument");
    MSE_CHECK(use_device(device));
    ResizeWork wk;
    DevBuf d;
    int rc = d.ensure((size_t)out_w * out_h * 3);
    if (rc == MSE_OK) rc = resize_rgb_to_device(wk, rgb, w, h, out_w, out_h, filter, d.as<uint8_t>(), nullptr);
    if (rc == MSE_OK && cudaMemcpy(out, d.p, (size_t)out_w * out_h * 3, cudaMemcpyDeviceToHost) != cudaSuccess) {
        set_error("resize_rgb_u8: D2H failed");
        rc = MSE_ERR_CUDA;
    }
    wk.release();
    d.release();
    return rc;
}
