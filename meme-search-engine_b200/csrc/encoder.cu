// SigLIP ViT-SO400M-14/384 image and text towers behind the clip_server boundary.
//
// Reference: clip_server.py:23 (OpenCLIP `ViT-SO400M-14-SigLIP-384`, precision fp16), :98 encode_text, :114 encode_image,
// :99/:115 L2 normalisation, :140 preprocessing; the layer graph is the one the author restates in
// aitemplate/model.py:13-122 (pre-LN blocks, erf-GELU MLP, learned position embedding, final LN, MAP head with one
// probe) with weight names as clip_server.py:46-62 maps them.  Text tower: OpenCLIP TextTransformer (token + position
// embedding, 27 non-causal blocks, final LN, LAST token, Linear 1152->1152 with bias).
//
// Every dense contraction runs on the tcgen05 GEMM (gemm_sm100.cuh) with bias / GELU / residual / position-embedding
// fused into its epilogue; LayerNorm, the im2col + u8 normalisation, attention (attention.cuh), the MAP pooling
// attention and the final L2 normalisation are the hand-written kernels in this file.  Activations are token-major
// fp16 [B*S][D] in HBM; statistics, softmax and accumulation are fp32.
#include "internal.h"
#include "gemm_epilogues.cuh"
#include "attention.cuh"
#include "attention_tc.cuh"
#include "text_mega.cuh"
#include "gemm_sm100.cuh"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <map>
#include <string>

namespace mse {

// ------------------------------------------------------------------ kernels

// u8 HWC image -> im2col rows for the 14x14/14 VALID patch conv: A[b*P*P + py*P + px][c*196 + ky*14 + kx] =
// x/127.5 - 1 (torchvision ToTensor + Normalize(0.5, 0.5), clip_server.py:140), K padded 588 -> kpad with zeros.
__global__ void __launch_bounds__(256) k_im2col_patch14(const uint8_t *__restrict__ img, __half *__restrict__ A, int B, int img_size,
                                                        int grid_p, int kpad) {
    const int patch = blockIdx.x;  // b*P*P + py*P + px
    const int b = patch / (grid_p * grid_p), pp = patch % (grid_p * grid_p), py = pp / grid_p, px = pp % grid_p;
    const uint8_t *src = img + (size_t)b * img_size * img_size * 3;
    __half *dst = A + (size_t)patch * kpad;
    for (int i = threadIdx.x; i < kpad; i += blockDim.x) {
        float v = 0.f;
        if (i < 588) {
            // iterate in source order (ky, kx, c) for coalesced reads; scatter into (c, ky, kx)
            int ky = i / 42, r = i % 42, kx = r / 3, c = r % 3;
            uint8_t u = src[((size_t)(py * 14 + ky) * img_size + (px * 14 + kx)) * 3 + c];
            v = (float)u * (1.0f / 127.5f) - 1.0f;
            dst[c * 196 + ky * 14 + kx] = __float2half_rn(v);
        } else {
            dst[i] = __float2half_rn(0.f);
        }
    }
}

// 24-bit BI_RGB BMP pixel arrays (BGR, rows padded to 4 bytes, bottom-up unless the header's height is negative) -> RGB top-down HWC:
// the container the reference's clients send (src/common.rs:42-53), unpacked on the device instead of by PIL (clip_server.py:140).
// meta[b] = {byte offset of the pixel array in `files`, row stride, bottom_up, 0}.  grid (img rows, batch)
__global__ void __launch_bounds__(256) k_bmp24_to_rgb(const uint8_t *__restrict__ files, const uint32_t *__restrict__ meta, uint8_t *__restrict__ out,
                                                      int img_size) {
    const int y = blockIdx.x, b = blockIdx.y;
    const uint32_t off = meta[4 * b], stride = meta[4 * b + 1], bottom_up = meta[4 * b + 2];
    const uint8_t *src = files + off + (size_t)(bottom_up ? img_size - 1 - y : y) * stride;
    uint8_t *dst = out + ((size_t)b * img_size + y) * img_size * 3;
    for (int i = threadIdx.x; i < img_size * 3; i += blockDim.x) {
        const int x = i / 3, c = i % 3;
        dst[i] = src[x * 3 + (2 - c)];
    }
}

// LayerNorm over the last dim (D % 8 == 0, D <= 256 * NV), fp32 statistics, one warp per row; NV = 16-byte pieces per lane
template <int NV>
__global__ void __launch_bounds__(256) k_layernorm(const __half *__restrict__ x, __half *__restrict__ y, const float *__restrict__ g,
                                                   const float *__restrict__ bta, uint32_t rows, uint32_t D, float eps,
                                                   uint32_t in_stride, uint32_t in_offset) {
    const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    pdl_trigger();
    pdl_wait();
    if (row >= rows) return;
    const uint4 *xr = (const uint4 *)(x + (size_t)row * in_stride + in_offset);
    const uint32_t nv = D >> 3;
    float v[NV][8];
    float s = 0.f;
    uint4 raw[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) {
        uint32_t p = lane + 32 * i;
        raw[i] = p < nv ? __ldg(xr + p) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < NV; i++) {
        uint32_t p = lane + 32 * i;
        if (p < nv) {
            uint4 u = raw[i];
            const __half2 *h = (const __half2 *)&u;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                float2 f = __half22float2(h[e]);
                v[i][2 * e] = f.x; v[i][2 * e + 1] = f.y;
                s += f.x + f.y;
            }
        }
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        uint32_t p = lane + 32 * i;
        if (p < nv) {
#pragma unroll
            for (int e = 0; e < 8; e++) { float d = v[i][e] - mean; q = fmaf(d, d, q); }
        }
    }
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)D + eps);
    uint4 *yr = (uint4 *)(y + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < NV; i++) {
        uint32_t p = lane + 32 * i;
        if (p < nv) {
            const float4 g0 = __ldg((const float4 *)(g + p * 8)), g1 = __ldg((const float4 *)(g + p * 8 + 4));
            const float4 b0 = __ldg((const float4 *)(bta + p * 8)), b1 = __ldg((const float4 *)(bta + p * 8 + 4));
            uint4 u;
            __half2 *h = (__half2 *)&u;
            h[0] = __floats2half2_rn((v[i][0] - mean) * rstd * g0.x + b0.x, (v[i][1] - mean) * rstd * g0.y + b0.y);
            h[1] = __floats2half2_rn((v[i][2] - mean) * rstd * g0.z + b0.z, (v[i][3] - mean) * rstd * g0.w + b0.w);
            h[2] = __floats2half2_rn((v[i][4] - mean) * rstd * g1.x + b1.x, (v[i][5] - mean) * rstd * g1.y + b1.y);
            h[3] = __floats2half2_rn((v[i][6] - mean) * rstd * g1.z + b1.z, (v[i][7] - mean) * rstd * g1.w + b1.w);
            yr[p] = u;
        }
    }
}

// Row statistics for a LayerNorm that is folded into the GEMM behind it (gemm_epilogues.cuh GemmOut::ln_stats): (mean, rstd) per row,
// the same two-pass fp32 arithmetic as k_layernorm; reads x once and writes 8 bytes per row instead of a normalised copy.
template <int NV>
__global__ void __launch_bounds__(256) k_rowstats(const __half *__restrict__ x, float2 *__restrict__ stats, uint32_t rows, uint32_t D, float eps) {
    const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    pdl_trigger();
    pdl_wait();
    if (row >= rows) return;
    const uint4 *xr = (const uint4 *)(x + (size_t)row * D);
    const uint32_t nv = D >> 3;
    float v[NV][8];
    float s = 0.f;
    uint4 raw[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) {
        uint32_t p = lane + 32 * i;
        raw[i] = p < nv ? __ldg(xr + p) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < NV; i++) {
        const __half2 *h = (const __half2 *)&raw[i];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            float2 f = __half22float2(h[e]);
            v[i][2 * e] = f.x; v[i][2 * e + 1] = f.y;
            s += f.x + f.y;
        }
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        uint32_t p = lane + 32 * i;
        if (p < nv) {
#pragma unroll
            for (int e = 0; e < 8; e++) { float d = v[i][e] - mean; q = fmaf(d, d, q); }
        }
    }
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) stats[row] = make_float2(mean, rsqrtf(q / (float)D + eps));
}

// MAP head pooling (aitemplate/model.py:98-111): one learned probe query per head against all S tokens.
// kv: [B*S][2*D] (k | v).  qp: [D] fp32 = Wq*probe + bq, precomputed at load.  out: [B][D] fp16.  grid (H, B), 128 threads
__global__ void __launch_bounds__(128) k_map_pool(const __half *__restrict__ kv, const float *__restrict__ qp, __half *__restrict__ out,
                                                  int S, int H, float scale) {
    extern __shared__ float s_sc[];  // S scores + 72 query values + scratch
    float *s_q = s_sc + S;
    __shared__ float s_red[4];
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, D = H * attn::kDH;
    const __half *kb = kv + (size_t)b * S * 2 * D + (size_t)h * attn::kDH;
    const __half *vb = kb + D;
    if (tid < attn::kDH) s_q[tid] = qp[h * attn::kDH + tid] * scale;
    __syncthreads();
    float mx = -INFINITY;
    for (int t = tid; t < S; t += blockDim.x) {
        const uint4 *kr = (const uint4 *)(kb + (size_t)t * 2 * D);
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < 9; c++) {
            uint4 u = __ldg(kr + c);
            const __half2 *hh = (const __half2 *)&u;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                float2 f = __half22float2(hh[e]);
                a = fmaf(f.x, s_q[c * 8 + 2 * e], a);
                a = fmaf(f.y, s_q[c * 8 + 2 * e + 1], a);
            }
        }
        s_sc[t] = a;
        mx = fmaxf(mx, a);
    }
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
    __syncthreads();
    float sum = 0.f;
    for (int t = tid; t < S; t += blockDim.x) {
        float p = __expf(s_sc[t] - mx);
        s_sc[t] = p;
        sum += p;
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = sum;
    __syncthreads();
    sum = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
    if (tid < attn::kDH) {
        float acc = 0.f;
        for (int t = 0; t < S; t++) acc = fmaf(s_sc[t], __half2float(vb[(size_t)t * 2 * D + tid]), acc);
        out[(size_t)b * D + h * attn::kDH + tid] = __float2half_rn(acc / sum);
    }
}

// out[b] = x[b] / |x[b]|_2 as fp16 (clip_server.py:99,115,166); one warp per row
__global__ void __launch_bounds__(256) k_l2norm_f16(const __half *__restrict__ x, __half *__restrict__ out, uint32_t rows, uint32_t D) {
    const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    pdl_trigger();
    pdl_wait();
    if (row >= rows) return;
    float s = 0.f;
    for (uint32_t j = lane; j < D; j += 32) {
        float v = __half2float(x[(size_t)row * D + j]);
        s = fmaf(v, v, s);
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = rsqrtf(s);
    for (uint32_t j = lane; j < D; j += 32) out[(size_t)row * D + j] = __float2half_rn(__half2float(x[(size_t)row * D + j]) * inv);
}

// text embedding: x[b*S + t] = tok_emb[ids[b*S + t]] + pos_emb[t]
__global__ void __launch_bounds__(256) k_text_embed(const int32_t *__restrict__ ids, const __half *__restrict__ tok, const __half *__restrict__ pos,
                                                    __half *__restrict__ x, uint32_t rows, uint32_t S, uint32_t D, uint32_t vocab) {
    const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    pdl_trigger();
    pdl_wait();   // the previous forward pass may still be reading x
    if (row >= rows) return;
    uint32_t id = (uint32_t)ids[row];
    if (id >= vocab) id = 0;
    const __half2 *tr = (const __half2 *)(tok + (size_t)id * D), *pr = (const __half2 *)(pos + (size_t)(row % S) * D);
    __half2 *xr = (__half2 *)(x + (size_t)row * D);
    for (uint32_t j = lane; j < D / 2; j += 32) {
        float2 a = __half22float2(tr[j]), p = __half22float2(pr[j]);
        xr[j] = __floats2half2_rn(a.x + p.x, a.y + p.y);
    }
}

// gather one token per sequence: out[b] = x[b*S + idx]
__global__ void k_gather_token(const __half *__restrict__ x, __half *__restrict__ out, uint32_t B, uint32_t S, uint32_t idx, uint32_t D) {
    const uint32_t b = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    for (uint32_t j = threadIdx.x; j < D / 8; j += blockDim.x)
        ((uint4 *)(out + (size_t)b * D))[j] = ((const uint4 *)(x + ((size_t)b * S + idx) * D))[j];
}

// ------------------------------------------------------------------ weights

struct Tensor {
    int dtype = 0;  // 0 f32, 1 f16, 2 i32
    std::vector<uint32_t> dims;
    const uint8_t *data = nullptr;
    size_t nbytes = 0;
    size_t numel() const { size_t n = 1; for (auto d : dims) n *= d; return n; }
};

struct WeightFile {
    std::vector<uint8_t> blob;
    std::map<std::string, Tensor> t;
    int load(const char *path) {
        FILE *f = fopen(path, "rb");
        MSE_REQUIRE(f != nullptr, MSE_ERR_INVALID, "encoder: cannot open weights file %s", path);
        fseek(f, 0, SEEK_END);
        long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        blob.resize((size_t)sz);
        size_t got = fread(blob.data(), 1, (size_t)sz, f);
        fclose(f);
        MSE_REQUIRE(got == (size_t)sz && sz >= 12 && memcmp(blob.data(), "MSEW0001", 8) == 0, MSE_ERR_INVALID,
                    "encoder: %s is not an MSEW0001 weights file", path);
        uint32_t n;
        memcpy(&n, blob.data() + 8, 4);
        size_t off = 12;
        for (uint32_t i = 0; i < n; i++) {
            MSE_REQUIRE(off + 2 <= blob.size(), MSE_ERR_INVALID, "encoder: truncated weights file");
            uint16_t nl;
            memcpy(&nl, blob.data() + off, 2);
            off += 2;
            std::string name((const char *)blob.data() + off, nl);
            off += nl;
            Tensor T;
            T.dtype = blob[off++];
            uint8_t nd = blob[off++];
            for (uint8_t k = 0; k < nd; k++) {
                uint32_t dd;
                memcpy(&dd, blob.data() + off, 4);
                off += 4;
                T.dims.push_back(dd);
            }
            uint64_t nb;
            memcpy(&nb, blob.data() + off, 8);
            off += 8;
            off = (off + 15) & ~(size_t)15;
            MSE_REQUIRE(off + nb <= blob.size(), MSE_ERR_INVALID, "encoder: tensor %s overruns the file", name.c_str());
            T.data = blob.data() + off;
            T.nbytes = nb;
            off += nb;
            t[name] = T;
        }
        return MSE_OK;
    }
    const Tensor *find(const std::string &n) const {
        auto it = t.find(n);
        return it == t.end() ? nullptr : &it->second;
    }
};

static inline float h2f_host(uint16_t h) { return __half2float(*reinterpret_cast<const __half *>(&h)); }

}  // namespace mse

using namespace mse;

struct LayerW {
    float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *qkv_b, *proj_b, *fc1_b, *fc2_b;
    __half *qkv_w, *proj_w, *fc1_w, *fc2_w;
};

struct TowerW {
    std::vector<LayerW> layers;
    float *lnf_g = nullptr, *lnf_b = nullptr;
};

struct mse_encoder {
    int device = 0;
    int32_t cfg[16] = {0};  // img, patch, dim, depth_v, heads, mlp, vocab, ctx, act, has_vision, has_text, depth_t, kpad
    int max_batch = 0;
    std::vector<void *> allocs;
    // vision
    TowerW vis;
    __half *patch_w = nullptr, *pos_v = nullptr, *kv_w = nullptr, *pproj_w = nullptr, *pfc1_w = nullptr, *pfc2_w = nullptr;
    float *patch_b = nullptr, *q_pool = nullptr, *kv_b = nullptr, *pproj_b = nullptr, *pln_g = nullptr, *pln_b = nullptr, *pfc1_b = nullptr,
          *pfc2_b = nullptr;
    // text
    TowerW txt;
    __half *tok_emb = nullptr, *pos_t = nullptr, *tproj_w = nullptr;
    float *tproj_b = nullptr;
    // workspace
    __half *x = nullptr, *xn = nullptr, *qkv = nullptr, *att = nullptr, *hbuf = nullptr, *pool = nullptr, *y = nullptr, *yn = nullptr,
           *hh = nullptr, *z = nullptr, *outb = nullptr;
    uint8_t *img_dev = nullptr;
    uint8_t *bmp_dev = nullptr;        // staging for mse_encode_images_bmp (allocated on first use)
    size_t bmp_cap = 0;
    uint32_t *bmp_meta = nullptr;
    int32_t *ids_dev = nullptr;
    float *splitk_ws = nullptr;       // skinny GEMM scratch (gemm_skinny.cuh): partial tiles + arrival counters
    uint32_t *splitk_cnt = nullptr;
    static constexpr size_t kSplitkFloats = (size_t)2 << 20;
    cudaStream_t stream = nullptr;
    size_t max_tokens = 0;
    // profiling (bench.py roofline): CUDA-event brackets around GEMM (class 0) and attention (class 1) launches
    int profile = 0;
    std::vector<cudaEvent_t> ev;
    std::vector<int> ev_class;
    size_t ev_used = 0;
    uint64_t stats[8] = {0};
    double gemm_flops = 0;
    // CUDA graphs of the text tower's forward pass, one per small batch size: at batch 1 the ~200 launches of a forward are
    // launch-latency bound (query traffic is batch 1: src/main.rs:899-934 embeds one text per search)
    static constexpr int kGraphMaxBatch = 16;
    cudaGraphExec_t text_graph[kGraphMaxBatch + 1] = {nullptr};
    uint64_t text_graph_launches[kGraphMaxBatch + 1] = {0};
    cudaStream_t cap_stream = nullptr;
    // vision tower: LN1 / LN2 folded into the qkv / fc1 GEMMs (k_fold_ln weights, k_rowstats statistics)
    struct FoldW { __half *qkv_w = nullptr, *fc1_w = nullptr; float *qkv_b = nullptr, *qkv_cs = nullptr, *fc1_b = nullptr, *fc1_cs = nullptr; };
    std::vector<FoldW> vis_fold;
    float2 *ln_stats = nullptr;             // [max_tokens]
    // one text query (64 tokens): the blocks as one persistent kernel (text_mega.cuh)
    tmega::LayerP *mega_layers = nullptr;   // device array [depth_t], LN folded into the qkv / fc1 weights
    uint32_t *mega_bar = nullptr;
    int mega_grid = 0;                      // 0: not available for this tower shape / device
};

namespace {

int dev_alloc(mse_encoder *e, void **p, size_t bytes) {
    cudaError_t err = cudaMalloc(p, bytes ? bytes : 16);
    if (err != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("encoder: cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(err));
        return MSE_ERR_OOM;
    }
    e->allocs.push_back(*p);
    return MSE_OK;
}

// upload a tensor as fp16 (matrices) or fp32 (vectors), converting on the host; optional column padding
int upload(mse_encoder *e, const WeightFile &wf, const std::string &name, bool as_half, void **out, size_t expect_numel,
           uint32_t rows = 0, uint32_t cols = 0, uint32_t cols_pad = 0) {
    const Tensor *T = wf.find(name);
    MSE_REQUIRE(T != nullptr, MSE_ERR_INVALID, "encoder: weights file lacks tensor '%s'", name.c_str());
    MSE_REQUIRE(T->numel() == expect_numel, MSE_ERR_INVALID, "encoder: tensor '%s' has %zu elements, expected %zu", name.c_str(),
                T->numel(), expect_numel);
    MSE_REQUIRE(T->dtype == 0 || T->dtype == 1, MSE_ERR_INVALID, "encoder: tensor '%s' has unsupported dtype", name.c_str());
    const size_t n = T->numel();
    auto get = [&](size_t i) -> float {
        return T->dtype == 0 ? ((const float *)T->data)[i] : h2f_host(((const uint16_t *)T->data)[i]);
    };
    if (as_half) {
        const size_t on = cols_pad ? (size_t)rows * cols_pad : n;
        std::vector<__half> h(on, __float2half(0.f));
        if (cols_pad) {
            for (uint32_t r = 0; r < rows; r++)
                for (uint32_t c = 0; c < cols; c++) h[(size_t)r * cols_pad + c] = __float2half_rn(get((size_t)r * cols + c));
        } else if (T->dtype == 1) {
            memcpy(h.data(), T->data, n * 2);
        } else {
            for (size_t i = 0; i < n; i++) h[i] = __float2half_rn(get(i));
        }
        MSE_CHECK(dev_alloc(e, out, on * 2));
        MSE_CUDA(cudaMemcpy(*out, h.data(), on * 2, cudaMemcpyHostToDevice));
    } else {
        std::vector<float> f(n);
        for (size_t i = 0; i < n; i++) f[i] = get(i);
        MSE_CHECK(dev_alloc(e, out, n * 4));
        MSE_CUDA(cudaMemcpy(*out, f.data(), n * 4, cudaMemcpyHostToDevice));
    }
    return MSE_OK;
}

int load_blocks(mse_encoder *e, const WeightFile &wf, TowerW &tw, int depth, bool vision) {
    const size_t D = e->cfg[2], F = e->cfg[5];
    tw.layers.resize(depth);
    char nm[256];
    for (int l = 0; l < depth; l++) {
        LayerW &L = tw.layers[l];
        auto N = [&](const char *vis_fmt, const char *txt_fmt) {
            snprintf(nm, sizeof(nm), vision ? vis_fmt : txt_fmt, l);
            return std::string(nm);
        };
        MSE_CHECK(upload(e, wf, N("visual.trunk.blocks.%d.norm1.weight", "text.transformer.resblocks.%d.ln_1.weight"), false, (void **)&L.ln1_g, D));
        MSE_CHECK(upload(e, wf, N("visual.trunk.blocks.%d.norm1.bias", "text.transformer.resblocks.%d.ln_1.bias"), false, (void **)&L.ln1_b, D));
        MSE_CHECK(upload(e, wf, N("visual.trunk.blocks.%d.attn.qkv.weight", "text.transformer.resblocks.%d.attn.in_proj_weight"), true, (void **)&L.qkv_w, 3 * D * D));
        MSE_CHECK(upload(e, wf, N("visual.trunk.blocks.%d.attn.qkv.bias", "text.transformer.resblocks.%d.attn.in_proj_bias"), false, (void **)&L.qkv_b, 3 * D));
        MSE_CHECK(upload(e, wf, N("visual.trunk.blocks.%d.attn.proj.weight", "text.transformer.resblocks.%d.attn.out_proj.weight"), true, (void **)&L.proj_w, D * D));
        MSE_CHECK(upload(e, wf, N("visual.trunk.blocks.%d.attn.proj.bias", "text.transformer.resblocks.%d.attn.out_proj.bias"), false, (void **)&L.proj_b, D));
        MSE_CHECK(upload(e, wf, N("visual.trunk.blocks.%d.norm2.weight", "text.transformer.resblocks.%d.ln_2.weight"), false, (void **)&L.ln2_g, D));
        MSE_CHECK(upload(e, wf, N("visual.trunk.blocks.%d.norm2.bias", "text.transformer.resblocks.%d.ln_2.bias"), false, (void **)&L.ln2_b, D));
        MSE_CHECK(upload(e, wf, N("visual.trunk.blocks.%d.mlp.fc1.weight", "text.transformer.resblocks.%d.mlp.c_fc.weight"), true, (void **)&L.fc1_w, F * D));
        MSE_CHECK(upload(e, wf, N("visual.trunk.blocks.%d.mlp.fc1.bias", "text.transformer.resblocks.%d.mlp.c_fc.bias"), false, (void **)&L.fc1_b, F));
        MSE_CHECK(upload(e, wf, N("visual.trunk.blocks.%d.mlp.fc2.weight", "text.transformer.resblocks.%d.mlp.c_proj.weight"), true, (void **)&L.fc2_w, D * F));
        MSE_CHECK(upload(e, wf, N("visual.trunk.blocks.%d.mlp.fc2.bias", "text.transformer.resblocks.%d.mlp.c_proj.bias"), false, (void **)&L.fc2_b, D));
    }
    return MSE_OK;
}

void prof_mark(mse_encoder *e, int cls, cudaStream_t st) {
    if (!e->profile) return;
    if (e->ev_used == e->ev.size()) {
        cudaEvent_t x;
        cudaEventCreate(&x);
        e->ev.push_back(x);
        e->ev_class.push_back(0);
    }
    e->ev_class[e->ev_used] = cls;
    cudaEventRecord(e->ev[e->ev_used++], st);
}

void prof_begin(mse_encoder *e) {
    e->ev_used = 0;
    e->gemm_flops = 0;
    memset(e->stats, 0, sizeof(e->stats));
}

void prof_collect(mse_encoder *e, uint64_t launches) {
    e->stats[4] = launches;
    if (!e->profile) return;
    double ns[2] = {0, 0};
    uint64_t cnt[2] = {0, 0};
    for (size_t i = 0; i + 1 < e->ev_used; i += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, e->ev[i], e->ev[i + 1]) == cudaSuccess) {
            ns[e->ev_class[i]] += (double)ms * 1e6;
            cnt[e->ev_class[i]]++;
        }
    }
    e->stats[0] = (uint64_t)ns[0]; e->stats[1] = cnt[0]; e->stats[2] = (uint64_t)ns[1]; e->stats[3] = cnt[1];
    e->stats[5] = (uint64_t)(e->gemm_flops / 1e6);  // MFLOP, algorithmic (2*M*N*K, unpadded)
}

int gemm(mse_encoder *e, const __half *A, const __half *W, uint32_t M, uint32_t N, uint32_t K, __half *C, const float *bias, int act,
         const __half *res, uint32_t res_mod, cudaStream_t st, const float2 *ln_stats = nullptr, const float *ln_cs = nullptr) {
    e->gemm_flops += 2.0 * M * N * (double)K;
    prof_mark(e, 0, st);
    struct Done { mse_encoder *e; cudaStream_t st; ~Done() { prof_mark(e, 0, st); } } done{e, st};
    GemmOut o{};
    o.c16 = C;
    o.ldc = N;
    o.bias = bias;
    o.act = act;
    o.res = res;
    o.res_mod = res_mod;
    o.splitk_ws = e->splitk_ws;
    o.splitk_cnt = e->splitk_cnt;
    o.splitk_ws_floats = mse_encoder::kSplitkFloats;
    o.ln_stats = ln_stats;
    o.ln_cs = ln_cs;
    return gemm_f16_tn_dev(e->device, A, W, M, N, K, K, K, o, st);
}

int layernorm(const __half *x, __half *y, const float *g, const float *b, uint32_t rows, uint32_t D, cudaStream_t st) {
    if (D <= 5 * 256) launch_pdl(k_layernorm<5>, (rows * 32 + 255) / 256, 256, 0, st, x, y, g, b, rows, D, 1e-6f, D, 0u);
    else launch_pdl(k_layernorm<8>, (rows * 32 + 255) / 256, 256, 0, st, x, y, g, b, rows, D, 1e-6f, D, 0u);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

int rowstats(const __half *x, float2 *stats, uint32_t rows, uint32_t D, cudaStream_t st) {
    if (D <= 5 * 256) launch_pdl(k_rowstats<5>, (rows * 32 + 255) / 256, 256, 0, st, x, stats, rows, D, 1e-6f);
    else launch_pdl(k_rowstats<8>, (rows * 32 + 255) / 256, 256, 0, st, x, stats, rows, D, 1e-6f);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

// vision tower: LN1 / LN2 folded into the weights of the GEMM behind them (the algebra is in text_mega.cuh)
int prepare_vision_fold(mse_encoder *e) {
    const uint32_t D = e->cfg[2], F = e->cfg[5];
    const int depth = e->cfg[3];
    if (getenv("MSE_NO_LN_FOLD") || D % 8 != 0 || D > 8 * 256) return MSE_OK;
    std::vector<mse_encoder::FoldW> fw(depth);
    for (int l = 0; l < depth; l++) {
        const LayerW &L = e->vis.layers[l];
        mse_encoder::FoldW &f = fw[l];
        MSE_CHECK(dev_alloc(e, (void **)&f.qkv_w, (size_t)3 * D * D * 2));
        MSE_CHECK(dev_alloc(e, (void **)&f.fc1_w, (size_t)F * D * 2));
        MSE_CHECK(dev_alloc(e, (void **)&f.qkv_cs, (size_t)3 * D * 4));
        MSE_CHECK(dev_alloc(e, (void **)&f.qkv_b, (size_t)3 * D * 4));
        MSE_CHECK(dev_alloc(e, (void **)&f.fc1_cs, (size_t)F * 4));
        MSE_CHECK(dev_alloc(e, (void **)&f.fc1_b, (size_t)F * 4));
        tmega::k_fold_ln<<<(3 * D * 32 + 255) / 256, 256>>>(L.qkv_w, L.qkv_b, L.ln1_g, L.ln1_b, 3 * D, D, f.qkv_w, f.qkv_cs, f.qkv_b);
        tmega::k_fold_ln<<<(F * 32 + 255) / 256, 256>>>(L.fc1_w, L.fc1_b, L.ln2_g, L.ln2_b, F, D, f.fc1_w, f.fc1_cs, f.fc1_b);
        MSE_LAUNCH_OK();
    }
    MSE_CUDA(cudaDeviceSynchronize());
    e->vis_fold = std::move(fw);
    return MSE_OK;
}

// text_mega.cuh: fold LN1 / LN2 into the qkv / fc1 weights of every text block and size the persistent grid
int prepare_text_mega(mse_encoder *e) {
    const uint32_t D = e->cfg[2], F = e->cfg[5], H = e->cfg[4], S = e->cfg[7];
    const int depth = e->cfg[11];
    e->mega_grid = 0;
    if (getenv("MSE_NO_TEXT_MEGA")) return MSE_OK;
    if (S != 64 || D % 64 != 0 || D != H * attn::kDH || F % 8 != 0 || depth <= 0) return MSE_OK;
    auto slices = [](uint32_t n, int bn) { return (n + (uint32_t)bn - 1) / (uint32_t)bn; };
    const uint32_t items = std::max(std::max(slices(3 * D, tmega::kBnQkv), slices(D, tmega::kBnProj)),
                                    std::max(std::max(slices(F, tmega::kBnFc1), slices(D, tmega::kBnFc2) * tmega::kFc2Splits), H));
    if (items > (uint32_t)sm_count(e->device) || slices(D, tmega::kBnFc2) > 256 ||
        (size_t)tmega::kFc2Splits * 64 * D > mse_encoder::kSplitkFloats)
        return MSE_OK;
    int coop = 0;
    MSE_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->device));
    if (!coop) return MSE_OK;                  // no co-residency guarantee: keep the multi-kernel path
    MSE_CUDA(cudaFuncSetAttribute(tmega::k_text_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tmega::kSmemBytes));
    int per_sm = 0;
    MSE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tmega::k_text_blocks, tmega::kThreads, tmega::kSmemBytes));
    if (per_sm < 1) return MSE_OK;
    std::vector<tmega::LayerP> lp(depth);
    for (int l = 0; l < depth; l++) {
        const LayerW &L = e->txt.layers[l];
        __half *qw = nullptr, *fw = nullptr;
        float *qcs = nullptr, *qb = nullptr, *fcs = nullptr, *fb = nullptr;
        MSE_CHECK(dev_alloc(e, (void **)&qw, (size_t)3 * D * D * 2));
        MSE_CHECK(dev_alloc(e, (void **)&fw, (size_t)F * D * 2));
        MSE_CHECK(dev_alloc(e, (void **)&qcs, (size_t)3 * D * 4));
        MSE_CHECK(dev_alloc(e, (void **)&qb, (size_t)3 * D * 4));
        MSE_CHECK(dev_alloc(e, (void **)&fcs, (size_t)F * 4));
        MSE_CHECK(dev_alloc(e, (void **)&fb, (size_t)F * 4));
        tmega::k_fold_ln<<<(3 * D * 32 + 255) / 256, 256>>>(L.qkv_w, L.qkv_b, L.ln1_g, L.ln1_b, 3 * D, D, qw, qcs, qb);
        tmega::k_fold_ln<<<(F * 32 + 255) / 256, 256>>>(L.fc1_w, L.fc1_b, L.ln2_g, L.ln2_b, F, D, fw, fcs, fb);
        MSE_LAUNCH_OK();
        lp[l] = tmega::LayerP{qw, L.proj_w, fw, L.fc2_w, qb, qcs, L.proj_b, fb, fcs, L.fc2_b};
    }
    MSE_CHECK(dev_alloc(e, (void **)&e->mega_layers, sizeof(tmega::LayerP) * depth));
    MSE_CHECK(dev_alloc(e, (void **)&e->mega_bar, 64));
    MSE_CUDA(cudaMemcpy(e->mega_layers, lp.data(), sizeof(tmega::LayerP) * depth, cudaMemcpyHostToDevice));
    MSE_CUDA(cudaDeviceSynchronize());
    e->mega_grid = (int)items;
    return MSE_OK;
}

int run_text_mega(mse_encoder *e, int depth, cudaStream_t st) {
    tmega::Params p{};
    p.layers = e->mega_layers;
    p.depth = depth;
    p.D = e->cfg[2]; p.F = e->cfg[5]; p.H = e->cfg[4];
    p.x = e->x; p.qkv = e->qkv; p.att = e->att; p.h = e->hbuf;
    p.ws = e->splitk_ws; p.cnt = e->splitk_cnt; p.bar = e->mega_bar;
    p.scale_log2e = (1.0f / sqrtf((float)attn::kDH)) * 1.4426950408889634f;
    p.eps = 1e-6f;
    p.act = e->cfg[8];
    { const char *d = getenv("MSE_TMEGA_DEBUG"); p.debug = d ? atoi(d) : 0; }
    MSE_CUDA(cudaMemsetAsync(e->mega_bar, 0, 4, st));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)e->mega_grid);
    cfg.blockDim = dim3(tmega::kThreads);
    cfg.dynamicSmemBytes = tmega::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the kernel synchronises the grid itself
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MSE_CUDA(cudaLaunchKernelEx(&cfg, tmega::k_text_blocks, p));
    MSE_LAUNCH_OK();
    return MSE_OK;
}

// the 27 pre-LN blocks shared by both towers (aitemplate/model.py:26-55)
int run_blocks(mse_encoder *e, const TowerW &tw, int depth, uint32_t B, uint32_t S, cudaStream_t st) {
    const uint32_t D = e->cfg[2], F = e->cfg[5], H = e->cfg[4], T = B * S;
    const int act = e->cfg[8];
    if (&tw == &e->txt && B == 1 && S == 64 && e->mega_grid > 0 && !e->profile && depth > 0 && getenv("MSE_NO_TEXT_MEGA") == nullptr)
        return run_text_mega(e, depth, st);
    static PerDeviceOnce attr_once;
    if (attr_once.first(e->device)) MSE_CUDA(cudaFuncSetAttribute(attn::k_mha_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(attn::Smem)));
    const float scale_log2e = (1.0f / sqrtf((float)attn::kDH)) * 1.4426950408889634f;
    // S >= 128: tcgen05 kernel (attention_tc.cuh); the 64-token text tower keeps the warp-level kernel (a 128-row tile would be half empty)
    const bool use_tc = S >= 128 && getenv("MSE_ATTN_MMA_SYNC") == nullptr;
    CUtensorMap tm64, tm16;
    attn_tc::Params ap{};
    if (use_tc) {
        static PerDeviceOnce tc_once;
        if (tc_once.first(e->device)) MSE_CUDA(cudaFuncSetAttribute(attn_tc::k_mha_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_tc::kSmemBytes));
        MSE_CHECK(encode_tmap_3d(&tm64, e->qkv, attn::kDH, 3 * H, (uint64_t)T, attn::kDH * 2, (uint64_t)3 * D * 2, 64, attn_tc::kBM, 128));
        MSE_CHECK(encode_tmap_3d(&tm16, e->qkv, attn::kDH, 3 * H, (uint64_t)T, attn::kDH * 2, (uint64_t)3 * D * 2, 16, attn_tc::kBM, 32));
        ap.S = (int)S; ap.H = (int)H; ap.B = (int)B;
        ap.q_items = (int)((S + 2 * attn_tc::kBM - 1) / (2 * attn_tc::kBM));
        ap.n_blocks = (int)((S + attn_tc::kBN - 1) / attn_tc::kBN);
        ap.n_items = (int)(B * H) * ap.q_items;
        ap.scale_log2e = scale_log2e;
    }
    // vision tower: no LayerNorm pass -- row statistics (half the traffic) + the LN-folded GEMM on the raw residual stream
    const bool fold = &tw == &e->vis && (int)e->vis_fold.size() >= depth && e->ln_stats && T > 128 && getenv("MSE_NO_LN_FOLD") == nullptr;
    for (int l = 0; l < depth; l++) {
        const LayerW &L = tw.layers[l];
        if (fold) {
            const mse_encoder::FoldW &f = e->vis_fold[l];
            MSE_CHECK(rowstats(e->x, e->ln_stats, T, D, st));
            MSE_CHECK(gemm(e, e->x, f.qkv_w, T, 3 * D, D, e->qkv, f.qkv_b, ACT_NONE, nullptr, 0, st, e->ln_stats, f.qkv_cs));
        } else {
        MSE_CHECK(layernorm(e->x, e->xn, L.ln1_g, L.ln1_b, T, D, st));
        MSE_CHECK(gemm(e, e->xn, L.qkv_w, T, 3 * D, D, e->qkv, L.qkv_b, ACT_NONE, nullptr, 0, st));
        }
        prof_mark(e, 1, st);
        if (use_tc)
            attn_tc::k_mha_tc<<<std::min(ap.n_items, sm_count(e->device)), attn_tc::kThreads, attn_tc::kSmemBytes, st>>>(tm64, tm16, e->att, ap);
        else
            launch_pdl(attn::k_mha_fwd, dim3((S + attn::kBM - 1) / attn::kBM, H, B), attn::kThreads, sizeof(attn::Smem), st, (const __half *)e->qkv, e->att,
                       (int)S, (int)H, scale_log2e);
        prof_mark(e, 1, st);
        MSE_LAUNCH_OK();
        MSE_CHECK(gemm(e, e->att, L.proj_w, T, D, D, e->x, L.proj_b, ACT_NONE, e->x, 0, st));
        if (fold) {
            const mse_encoder::FoldW &f = e->vis_fold[l];
            MSE_CHECK(rowstats(e->x, e->ln_stats, T, D, st));
            MSE_CHECK(gemm(e, e->x, f.fc1_w, T, F, D, e->hbuf, f.fc1_b, act, nullptr, 0, st, e->ln_stats, f.fc1_cs));
        } else {
        MSE_CHECK(layernorm(e->x, e->xn, L.ln2_g, L.ln2_b, T, D, st));
        MSE_CHECK(gemm(e, e->xn, L.fc1_w, T, F, D, e->hbuf, L.fc1_b, act, nullptr, 0, st));
        }
        MSE_CHECK(gemm(e, e->hbuf, L.fc2_w, T, D, F, e->x, L.fc2_b, ACT_NONE, e->x, 0, st));
    }
    return MSE_OK;
}

// images already in e->img_dev (u8 HWC); result fp16 [B][D] in e->outb.  layer_stop >= 0: stop after that many blocks
// and leave the token activations in e->x (debug / per-layer parity)
int vision_forward(mse_encoder *e, uint32_t B, int layer_stop, cudaStream_t st) {
    const uint32_t D = e->cfg[2], F = e->cfg[5], H = e->cfg[4], img = e->cfg[0], P = img / e->cfg[1], S = P * P, T = B * S, kpad = e->cfg[12];
    const int act = e->cfg[8];
    k_im2col_patch14<<<T, 256, 0, st>>>(e->img_dev, e->hbuf, (int)B, (int)img, (int)P, (int)kpad);
    MSE_LAUNCH_OK();
    MSE_CHECK(gemm(e, e->hbuf, e->patch_w, T, D, kpad, e->x, e->patch_b, ACT_NONE, e->pos_v, S, st));
    const int depth = layer_stop >= 0 ? std::min(layer_stop, (int)e->cfg[3]) : (int)e->cfg[3];
    MSE_CHECK(run_blocks(e, e->vis, depth, B, S, st));
    if (layer_stop >= 0) return MSE_OK;
    MSE_CHECK(layernorm(e->x, e->xn, e->vis.lnf_g, e->vis.lnf_b, T, D, st));
    // MAP head (aitemplate/model.py:82-111)
    MSE_CHECK(gemm(e, e->xn, e->kv_w, T, 2 * D, D, e->qkv, e->kv_b, ACT_NONE, nullptr, 0, st));
    k_map_pool<<<dim3(H, B), 128, (S + attn::kDH + 8) * sizeof(float), st>>>(e->qkv, e->q_pool, e->pool, (int)S, (int)H,
                                                                             1.0f / sqrtf((float)attn::kDH));
    MSE_LAUNCH_OK();
    MSE_CHECK(gemm(e, e->pool, e->pproj_w, B, D, D, e->y, e->pproj_b, ACT_NONE, nullptr, 0, st));
    MSE_CHECK(layernorm(e->y, e->yn, e->pln_g, e->pln_b, B, D, st));
    MSE_CHECK(gemm(e, e->yn, e->pfc1_w, B, F, D, e->hh, e->pfc1_b, act, nullptr, 0, st));
    MSE_CHECK(gemm(e, e->hh, e->pfc2_w, B, D, F, e->z, e->pfc2_b, ACT_NONE, e->y, 0, st));
    launch_pdl(k_l2norm_f16, (B * 32 + 255) / 256, 256, 0, st, e->z, e->outb, B, D);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

int text_forward(mse_encoder *e, uint32_t B, int layer_stop, cudaStream_t st) {
    const uint32_t D = e->cfg[2], S = e->cfg[7], T = B * S;
    launch_pdl(k_text_embed, (T * 32 + 255) / 256, 256, 0, st, e->ids_dev, e->tok_emb, e->pos_t, e->x, T, S, D, (uint32_t)e->cfg[6]);
    MSE_LAUNCH_OK();
    const int depth = layer_stop >= 0 ? std::min(layer_stop, (int)e->cfg[11]) : (int)e->cfg[11];
    MSE_CHECK(run_blocks(e, e->txt, depth, B, S, st));
    if (layer_stop >= 0) return MSE_OK;
    MSE_CHECK(layernorm(e->x, e->xn, e->txt.lnf_g, e->txt.lnf_b, T, D, st));
    launch_pdl(k_gather_token, B, 128, 0, st, e->xn, e->pool, B, S, S - 1, D);
    MSE_LAUNCH_OK();
    MSE_CHECK(gemm(e, e->pool, e->tproj_w, B, D, D, e->z, e->tproj_b, ACT_NONE, nullptr, 0, st));
    launch_pdl(k_l2norm_f16, (B * 32 + 255) / 256, 256, 0, st, e->z, e->outb, B, D);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

}  // namespace

// ================================================================== C ABI

MSE_API void mse_encoder_destroy(mse_encoder *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (void *p : e->allocs) cudaFree(p);
    for (cudaEvent_t x : e->ev) cudaEventDestroy(x);
    if (e->bmp_dev) cudaFree(e->bmp_dev);
    if (e->bmp_meta) cudaFree(e->bmp_meta);
    for (cudaGraphExec_t gx : e->text_graph)
        if (gx) cudaGraphExecDestroy(gx);
    if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

MSE_API int mse_encoder_create(const char *weights_path, int device, int max_batch, mse_encoder **out) {
    MSE_REQUIRE(out != nullptr && weights_path != nullptr, MSE_ERR_INVALID, "encoder_create: NULL argument");
    *out = nullptr;
    MSE_REQUIRE(max_batch >= 1 && max_batch <= 1024, MSE_ERR_INVALID, "encoder_create: max_batch=%d out of range [1,1024]", max_batch);
    MSE_CHECK(use_device(device));
    WeightFile wf;
    MSE_CHECK(wf.load(weights_path));
    const Tensor *cfgT = wf.find("config");
    MSE_REQUIRE(cfgT && cfgT->dtype == 2 && cfgT->numel() >= 12, MSE_ERR_INVALID, "encoder_create: weights file lacks the i32 'config' tensor");
    mse_encoder *e = new mse_encoder();
    e->device = device;
    e->max_batch = max_batch;
    memcpy(e->cfg, cfgT->data, std::min<size_t>(cfgT->nbytes, sizeof(e->cfg)));
    int rc = MSE_OK;
    auto fail = [&](int code) { mse_encoder_destroy(e); return code; };
    const int img = e->cfg[0], patch = e->cfg[1], D = e->cfg[2], depth_v = e->cfg[3], H = e->cfg[4], F = e->cfg[5], vocab = e->cfg[6],
              ctx = e->cfg[7], act = e->cfg[8], has_v = e->cfg[9], has_t = e->cfg[10], depth_t = e->cfg[11];
    if (!(patch == 14 && D == H * attn::kDH && D % 64 == 0 && D <= 2048 && F % 8 == 0 && (act == 1 || act == 2) && img % patch <= patch &&
          img >= patch && ctx >= 1 && (has_v || has_t))) {
        set_error("encoder_create: unsupported architecture (img %d patch %d dim %d heads %d mlp %d ctx %d act %d)", img, patch, D, H, F, ctx, act);
        return fail(MSE_ERR_UNSUPPORTED);
    }
    const int kpad = 640;
    e->cfg[12] = kpad;
    const size_t P = img / patch, Sv = P * P;
    do {
        if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = MSE_ERR_CUDA; set_error("encoder_create: stream"); break; }
        if (has_v) {
            if ((rc = upload(e, wf, "visual.trunk.patch_embed.proj.weight", true, (void **)&e->patch_w, (size_t)D * 588, D, 588, kpad))) break;
            if ((rc = upload(e, wf, "visual.trunk.patch_embed.proj.bias", false, (void **)&e->patch_b, D))) break;
            if ((rc = upload(e, wf, "visual.trunk.pos_embed", true, (void **)&e->pos_v, Sv * D))) break;
            if ((rc = load_blocks(e, wf, e->vis, depth_v, true))) break;
            if ((rc = upload(e, wf, "visual.trunk.norm.weight", false, (void **)&e->vis.lnf_g, D))) break;
            if ((rc = upload(e, wf, "visual.trunk.norm.bias", false, (void **)&e->vis.lnf_b, D))) break;
            if ((rc = upload(e, wf, "visual.trunk.attn_pool.kv.weight", true, (void **)&e->kv_w, (size_t)2 * D * D))) break;
            if ((rc = upload(e, wf, "visual.trunk.attn_pool.kv.bias", false, (void **)&e->kv_b, 2 * D))) break;
            if ((rc = upload(e, wf, "visual.trunk.attn_pool.proj.weight", true, (void **)&e->pproj_w, (size_t)D * D))) break;
            if ((rc = upload(e, wf, "visual.trunk.attn_pool.proj.bias", false, (void **)&e->pproj_b, D))) break;
            if ((rc = upload(e, wf, "visual.trunk.attn_pool.norm.weight", false, (void **)&e->pln_g, D))) break;
            if ((rc = upload(e, wf, "visual.trunk.attn_pool.norm.bias", false, (void **)&e->pln_b, D))) break;
            if ((rc = upload(e, wf, "visual.trunk.attn_pool.mlp.fc1.weight", true, (void **)&e->pfc1_w, (size_t)F * D))) break;
            if ((rc = upload(e, wf, "visual.trunk.attn_pool.mlp.fc1.bias", false, (void **)&e->pfc1_b, F))) break;
            if ((rc = upload(e, wf, "visual.trunk.attn_pool.mlp.fc2.weight", true, (void **)&e->pfc2_w, (size_t)D * F))) break;
            if ((rc = upload(e, wf, "visual.trunk.attn_pool.mlp.fc2.bias", false, (void **)&e->pfc2_b, D))) break;
            // probe query q = Wq * latent + bq (aitemplate/model.py:99-100), once per weight load, fp32 on the host from the
            // fp16-rounded weights the GPU uses
            const Tensor *lat = wf.find("visual.trunk.attn_pool.latent"), *qw = wf.find("visual.trunk.attn_pool.q.weight"),
                         *qb = wf.find("visual.trunk.attn_pool.q.bias");
            if (!lat || !qw || !qb || lat->numel() != (size_t)D || qw->numel() != (size_t)D * D || qb->numel() != (size_t)D) {
                set_error("encoder_create: attn_pool.latent / q.weight / q.bias missing or mis-shaped");
                rc = MSE_ERR_INVALID;
                break;
            }
            auto val = [&](const Tensor *T, size_t i) { return T->dtype == 0 ? ((const float *)T->data)[i] : h2f_host(((const uint16_t *)T->data)[i]); };
            auto r16 = [&](float v) { return __half2float(__float2half_rn(v)); };
            std::vector<float> qp(D);
            for (int o = 0; o < D; o++) {
                double acc = 0.0;
                for (int i = 0; i < D; i++) acc += (double)r16(val(qw, (size_t)o * D + i)) * (double)r16(val(lat, i));
                qp[o] = (float)acc + val(qb, o);
            }
            if ((rc = dev_alloc(e, (void **)&e->q_pool, (size_t)D * 4))) break;
            if (cudaMemcpy(e->q_pool, qp.data(), (size_t)D * 4, cudaMemcpyHostToDevice) != cudaSuccess) { rc = MSE_ERR_CUDA; set_error("encoder_create: H2D"); break; }
        }
        if (has_t) {
            if ((rc = upload(e, wf, "text.token_embedding.weight", true, (void **)&e->tok_emb, (size_t)vocab * D))) break;
            if ((rc = upload(e, wf, "text.positional_embedding", true, (void **)&e->pos_t, (size_t)ctx * D))) break;
            if ((rc = load_blocks(e, wf, e->txt, depth_t, false))) break;
            if ((rc = upload(e, wf, "text.ln_final.weight", false, (void **)&e->txt.lnf_g, D))) break;
            if ((rc = upload(e, wf, "text.ln_final.bias", false, (void **)&e->txt.lnf_b, D))) break;
            if ((rc = upload(e, wf, "text.text_projection.weight", true, (void **)&e->tproj_w, (size_t)D * D))) break;
            if ((rc = upload(e, wf, "text.text_projection.bias", false, (void **)&e->tproj_b, D))) break;
            if ((rc = prepare_text_mega(e))) break;
        }
        const size_t S_max = std::max<size_t>(has_v ? Sv : 0, has_t ? (size_t)ctx : 0);
        const size_t T = (size_t)max_batch * S_max;
        e->max_tokens = T;
        if ((rc = dev_alloc(e, (void **)&e->x, T * D * 2))) break;
        if ((rc = dev_alloc(e, (void **)&e->xn, T * D * 2))) break;
        if ((rc = dev_alloc(e, (void **)&e->qkv, T * 3 * D * 2))) break;
        if ((rc = dev_alloc(e, (void **)&e->att, T * D * 2))) break;
        if ((rc = dev_alloc(e, (void **)&e->hbuf, T * std::max<size_t>(F, kpad) * 2))) break;
        const size_t Bm = max_batch;
        if ((rc = dev_alloc(e, (void **)&e->pool, Bm * D * 2))) break;
        if ((rc = dev_alloc(e, (void **)&e->y, Bm * D * 2))) break;
        if ((rc = dev_alloc(e, (void **)&e->yn, Bm * D * 2))) break;
        if ((rc = dev_alloc(e, (void **)&e->hh, Bm * F * 2))) break;
        if ((rc = dev_alloc(e, (void **)&e->z, Bm * D * 2))) break;
        if ((rc = dev_alloc(e, (void **)&e->outb, Bm * D * 2))) break;
        if (has_v && (rc = dev_alloc(e, (void **)&e->img_dev, Bm * img * img * 3))) break;
        if (has_t && (rc = dev_alloc(e, (void **)&e->ids_dev, Bm * ctx * 4))) break;
        if ((rc = dev_alloc(e, (void **)&e->splitk_ws, mse_encoder::kSplitkFloats * 4))) break;
        if ((rc = dev_alloc(e, (void **)&e->ln_stats, T * sizeof(float2)))) break;
        if (has_v && (rc = prepare_vision_fold(e))) break;
        if ((rc = dev_alloc(e, (void **)&e->splitk_cnt, 256 * 4))) break;
        if (cudaMemset(e->splitk_cnt, 0, 256 * 4) != cudaSuccess) { set_error("encoder_create: cudaMemset failed"); rc = MSE_ERR_CUDA; break; }
    } while (0);
    if (rc != MSE_OK) return fail(rc);
    *out = e;
    return MSE_OK;
}

MSE_API int mse_encoder_profile(mse_encoder *e, int enable) {
    MSE_REQUIRE(e != nullptr, MSE_ERR_INVALID, "encoder_profile: NULL handle");
    e->profile = enable ? 1 : 0;
    return MSE_OK;
}

// synchronises the device, then reports the last encode call
MSE_API int mse_encoder_stats(mse_encoder *e, uint64_t out[8]) {
    MSE_REQUIRE(e != nullptr && out != nullptr, MSE_ERR_INVALID, "encoder_stats: NULL argument");
    MSE_CHECK(use_device(e->device));
    MSE_CUDA(cudaDeviceSynchronize());
    prof_collect(e, e->stats[4]);
    memcpy(out, e->stats, sizeof(e->stats));
    return MSE_OK;
}

MSE_API int mse_encoder_config(const mse_encoder *e, int32_t out[16]) {
    MSE_REQUIRE(e != nullptr && out != nullptr, MSE_ERR_INVALID, "encoder_config: NULL argument");
    memcpy(out, e->cfg, sizeof(e->cfg));
    return MSE_OK;
}

static int encode_images_impl(mse_encoder *e, const uint8_t *img, bool img_on_device, int batch, uint16_t *out, bool out_on_device,
                              int layer_stop, cudaStream_t st) {
    MSE_REQUIRE(e != nullptr, MSE_ERR_INVALID, "encode_images: NULL handle");
    MSE_REQUIRE(e->cfg[9], MSE_ERR_STATE, "encode_images: this encoder was loaded without a vision tower");
    MSE_REQUIRE(batch >= 0 && (batch == 0 || (img && out)), MSE_ERR_INVALID, "encode_images: bad argument");
    MSE_REQUIRE(batch <= e->max_batch, MSE_ERR_INVALID, "encode_images: max batch size is %d", e->max_batch);  // clip_server.py:139
    if (batch == 0) return MSE_OK;
    MSE_CHECK(use_device(e->device));
    const size_t D = e->cfg[2], img_px = (size_t)e->cfg[0] * e->cfg[0] * 3, P = e->cfg[0] / e->cfg[1];
    prof_begin(e);
    const uint64_t l0 = g_launches.load();
    MSE_CUDA(cudaMemcpyAsync(e->img_dev, img, (size_t)batch * img_px, img_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    MSE_CHECK(vision_forward(e, (uint32_t)batch, layer_stop, st));
    e->stats[4] = g_launches.load() - l0;
    if (layer_stop >= 0)
        MSE_CUDA(cudaMemcpyAsync(out, e->x, (size_t)batch * P * P * D * 2, out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    else
        MSE_CUDA(cudaMemcpyAsync(out, e->outb, (size_t)batch * D * 2, out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (!out_on_device) MSE_CUDA(cudaStreamSynchronize(st));
    return MSE_OK;
}

static int encode_text_impl(mse_encoder *e, const int32_t *ids, bool ids_on_device, int batch, uint16_t *out, bool out_on_device,
                            int layer_stop, cudaStream_t st) {
    MSE_REQUIRE(e != nullptr, MSE_ERR_INVALID, "encode_text: NULL handle");
    MSE_REQUIRE(e->cfg[10], MSE_ERR_STATE, "encode_text: this encoder was loaded without a text tower");
    MSE_REQUIRE(batch >= 0 && (batch == 0 || (ids && out)), MSE_ERR_INVALID, "encode_text: bad argument");
    MSE_REQUIRE(batch <= e->max_batch, MSE_ERR_INVALID, "encode_text: max batch size is %d", e->max_batch);  // clip_server.py:136
    if (batch == 0) return MSE_OK;
    MSE_CHECK(use_device(e->device));
    const size_t D = e->cfg[2], S = e->cfg[7];
    prof_begin(e);
    const uint64_t l0 = g_launches.load();
    MSE_CUDA(cudaMemcpyAsync(e->ids_dev, ids, (size_t)batch * S * 4, ids_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    static const bool no_graph = getenv("MSE_NO_GRAPH") != nullptr;
    if (layer_stop < 0 && !e->profile && batch <= mse_encoder::kGraphMaxBatch && !no_graph) {
        // small batches: replay the whole forward pass as one CUDA graph.  The first call at a batch size runs eagerly (that also sets
        // every kernel's attributes) and records the graph; buffers, shapes and tensor maps of a batch size never change.
        if (!e->text_graph[batch]) {
            MSE_CHECK(text_forward(e, (uint32_t)batch, -1, st));
            MSE_CUDA(cudaStreamSynchronize(st));
            if (!e->cap_stream) MSE_CUDA(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
            cudaGraph_t graph = nullptr;
            const uint64_t c0 = g_launches.load();
            MSE_CUDA(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
            const int rc = text_forward(e, (uint32_t)batch, -1, e->cap_stream);
            const cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
            if (rc != MSE_OK || ce != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                (void)cudaGetLastError();
                if (rc == MSE_OK) set_error("encode_text: graph capture failed: %s", cudaGetErrorString(ce));
                return rc != MSE_OK ? rc : MSE_ERR_CUDA;
            }
            e->text_graph_launches[batch] = g_launches.load() - c0;
            const cudaError_t ie = cudaGraphInstantiate(&e->text_graph[batch], graph, 0);
            cudaGraphDestroy(graph);
            MSE_REQUIRE(ie == cudaSuccess, MSE_ERR_CUDA, "encode_text: cudaGraphInstantiate -> %s", cudaGetErrorString(ie));
        } else {
            MSE_CUDA(cudaGraphLaunch(e->text_graph[batch], st));
            count_launch(e->text_graph_launches[batch]);      // the kernels inside the graph
        }
    } else {
        MSE_CHECK(text_forward(e, (uint32_t)batch, layer_stop, st));
    }
    e->stats[4] = g_launches.load() - l0;
    if (layer_stop >= 0)
        MSE_CUDA(cudaMemcpyAsync(out, e->x, (size_t)batch * S * D * 2, out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    else
        MSE_CUDA(cudaMemcpyAsync(out, e->outb, (size_t)batch * D * 2, out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (!out_on_device) MSE_CUDA(cudaStreamSynchronize(st));
    return MSE_OK;
}

MSE_API int mse_encode_images_u8(mse_encoder *e, const uint8_t *rgb_hwc, int batch, uint16_t *out_f16) {
    return encode_images_impl(e, rgb_hwc, false, batch, out_f16, false, -1, e ? e->stream : nullptr);
}
MSE_API int mse_encode_images_u8_dev(mse_encoder *e, const uint8_t *d_rgb_hwc, int batch, uint16_t *d_out_f16, void *stream) {
    return encode_images_impl(e, d_rgb_hwc, true, batch, d_out_f16, true, -1, (cudaStream_t)stream);
}
MSE_API int mse_encode_text_ids(mse_encoder *e, const int32_t *ids, int batch, uint16_t *out_f16) {
    return encode_text_impl(e, ids, false, batch, out_f16, false, -1, e ? e->stream : nullptr);
}
MSE_API int mse_encode_text_ids_dev(mse_encoder *e, const int32_t *d_ids, int batch, uint16_t *d_out_f16, void *stream) {
    return encode_text_impl(e, d_ids, true, batch, d_out_f16, true, -1, (cudaStream_t)stream);
}
// The files the reference's clients send (src/common.rs:42-53: image_size x image_size 24-bit BI_RGB BMPs from the image crate's BmpEncoder):
// headers are read on the host, the pixel arrays are unpacked on the device.  Anything else is MSE_ERR_UNSUPPORTED -- decode it on the host
// and call mse_encode_images_u8.
MSE_API int mse_encode_images_bmp(mse_encoder *e, const uint8_t *const *bmps, const size_t *lens, int batch, uint16_t *out_f16) {
    MSE_REQUIRE(e != nullptr, MSE_ERR_INVALID, "encode_images_bmp: NULL handle");
    MSE_REQUIRE(e->cfg[9], MSE_ERR_STATE, "encode_images_bmp: this encoder was loaded without a vision tower");
    MSE_REQUIRE(batch >= 0 && (batch == 0 || (bmps && lens && out_f16)), MSE_ERR_INVALID, "encode_images_bmp: bad argument");
    MSE_REQUIRE(batch <= e->max_batch, MSE_ERR_INVALID, "encode_images_bmp: max batch size is %d", e->max_batch);
    if (batch == 0) return MSE_OK;
    MSE_CHECK(use_device(e->device));
    const int S = e->cfg[0];
    auto rd32 = [](const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); };
    std::vector<uint32_t> meta((size_t)batch * 4);
    size_t total = 0;
    for (int i = 0; i < batch; i++) {
        const uint8_t *f = bmps[i];
        MSE_REQUIRE(f && lens[i] >= 54 && f[0] == 'B' && f[1] == 'M', MSE_ERR_UNSUPPORTED, "encode_images_bmp: image %d is not a BMP file", i);
        const uint32_t data_off = rd32(f + 10), dib = rd32(f + 14);
        const int32_t w = (int32_t)rd32(f + 18), h = (int32_t)rd32(f + 22);
        const uint32_t bpp = f[28] | (f[29] << 8), comp = rd32(f + 30);
        const uint32_t stride = ((uint32_t)S * 3 + 3) & ~3u;
        MSE_REQUIRE(dib >= 40 && bpp == 24 && comp == 0 && w == S && (h == S || h == -S), MSE_ERR_UNSUPPORTED,
                    "encode_images_bmp: image %d is %dx%d, %u bpp, compression %u; only %dx%d 24-bit BI_RGB is unpacked on the device", i, w, h, bpp, comp, S, S);
        MSE_REQUIRE((size_t)data_off + (size_t)stride * S <= lens[i], MSE_ERR_INVALID, "encode_images_bmp: image %d is truncated", i);
        meta[4 * i] = (uint32_t)total + data_off;
        meta[4 * i + 1] = stride;
        meta[4 * i + 2] = h > 0 ? 1u : 0u;
        meta[4 * i + 3] = 0;
        total += (lens[i] + 15) & ~(size_t)15;
        MSE_REQUIRE(total < (1ull << 32), MSE_ERR_UNSUPPORTED, "encode_images_bmp: batch too large");
    }
    if (total > e->bmp_cap) {
        if (e->bmp_dev) cudaFree(e->bmp_dev);
        e->bmp_dev = nullptr; e->bmp_cap = 0;
        MSE_CUDA(cudaMalloc(&e->bmp_dev, total + total / 4));
        e->bmp_cap = total + total / 4;
    }
    if (!e->bmp_meta) MSE_CUDA(cudaMalloc(&e->bmp_meta, (size_t)e->max_batch * 16));
    cudaStream_t st = e->stream;
    size_t at = 0;
    for (int i = 0; i < batch; i++) {
        MSE_CUDA(cudaMemcpyAsync(e->bmp_dev + at, bmps[i], lens[i], cudaMemcpyHostToDevice, st));
        at += (lens[i] + 15) & ~(size_t)15;
    }
    MSE_CUDA(cudaMemcpyAsync(e->bmp_meta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice, st));
    MSE_CUDA(cudaStreamSynchronize(st));                         // `meta` lives on this stack frame
    prof_begin(e);
    const uint64_t l0 = g_launches.load();
    k_bmp24_to_rgb<<<dim3(S, batch), 256, 0, st>>>(e->bmp_dev, e->bmp_meta, e->img_dev, S);
    MSE_LAUNCH_OK();
    MSE_CHECK(vision_forward(e, (uint32_t)batch, -1, st));
    e->stats[4] = g_launches.load() - l0;
    MSE_CUDA(cudaMemcpyAsync(out_f16, e->outb, (size_t)batch * e->cfg[2] * 2, cudaMemcpyDeviceToHost, st));
    MSE_CUDA(cudaStreamSynchronize(st));
    return MSE_OK;
}

// Decoded images of any size (what the ingest client holds before resize_for_embed_sync, src/common.rs:31-54): resized on the device
// straight into the tower's input buffer, no BMP round trip.
MSE_API int mse_encode_images_resized(mse_encoder *e, const uint8_t *const *rgb, const uint32_t *widths, const uint32_t *heights, int batch,
                                      uint16_t *out_f16) {
    MSE_REQUIRE(e != nullptr, MSE_ERR_INVALID, "encode_images_resized: NULL handle");
    MSE_REQUIRE(e->cfg[9], MSE_ERR_STATE, "encode_images_resized: this encoder was loaded without a vision tower");
    MSE_REQUIRE(batch >= 0 && (batch == 0 || (rgb && widths && heights && out_f16)), MSE_ERR_INVALID, "encode_images_resized: bad argument");
    MSE_REQUIRE(batch <= e->max_batch, MSE_ERR_INVALID, "encode_images_resized: max batch size is %d", e->max_batch);
    if (batch == 0) return MSE_OK;
    MSE_CHECK(use_device(e->device));
    const uint32_t S = e->cfg[0];
    cudaStream_t st = e->stream;
    ResizeWork wk;
    int rc = MSE_OK;
    for (int i = 0; i < batch && rc == MSE_OK; i++)
        rc = resize_rgb_to_device(wk, rgb[i], widths[i], heights[i], S, S, 0, e->img_dev + (size_t)i * S * S * 3, st);
    if (rc == MSE_OK) {
        prof_begin(e);
        const uint64_t l0 = g_launches.load();
        rc = vision_forward(e, (uint32_t)batch, -1, st);
        e->stats[4] = g_launches.load() - l0;
    }
    if (rc == MSE_OK && cudaMemcpyAsync(out_f16, e->outb, (size_t)batch * e->cfg[2] * 2, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
        set_error("encode_images_resized: D2H failed");
        rc = MSE_ERR_CUDA;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess && rc == MSE_OK) { set_error("encode_images_resized: stream failed"); rc = MSE_ERR_CUDA; }
    wk.release();
    return rc;
}

MSE_API int mse_encode_images_hidden(mse_encoder *e, const uint8_t *rgb_hwc, int batch, int n_blocks, uint16_t *out_tokens_f16) {
    MSE_REQUIRE(n_blocks >= 0, MSE_ERR_INVALID, "encode_images_hidden: n_blocks must be >= 0");
    return encode_images_impl(e, rgb_hwc, false, batch, out_tokens_f16, false, n_blocks, e ? e->stream : nullptr);
}
MSE_API int mse_encode_text_hidden(mse_encoder *e, const int32_t *ids, int batch, int n_blocks, uint16_t *out_tokens_f16) {
    MSE_REQUIRE(n_blocks >= 0, MSE_ERR_INVALID, "encode_text_hidden: n_blocks must be >= 0");
    return encode_text_impl(e, ids, false, batch, out_tokens_f16, false, n_blocks, e ? e->stream : nullptr);
}

// Profiling aid (not part of the reference surface): times the tcgen05 attention kernel alone on random data.
// mode: attn_tc::Params::debug bits.  Returns the average kernel time in milliseconds.
MSE_API int mse_debug_attention(int device, int B, int S, int mode, int iters, float *ms_out) {
    MSE_CHECK(use_device(device));
    const int H = 16, D = H * attn::kDH;
    const size_t T = (size_t)B * S;
    __half *qkv = nullptr, *out = nullptr;
    MSE_CUDA(cudaMalloc(&qkv, T * 3 * D * 2));
    MSE_CUDA(cudaMalloc(&out, T * D * 2));
    std::vector<__half> h(T * 3 * D);
    uint32_t x = 12345;
    for (auto &v : h) { x = x * 1664525u + 1013904223u; v = __float2half(((int)(x >> 16) % 2001 - 1000) * 1e-3f); }
    cudaMemcpy(qkv, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    MSE_CUDA(cudaFuncSetAttribute(attn_tc::k_mha_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_tc::kSmemBytes));
    CUtensorMap tm64, tm16;
    MSE_CHECK(encode_tmap_3d(&tm64, qkv, attn::kDH, 3 * H, (uint64_t)T, attn::kDH * 2, (uint64_t)3 * D * 2, 64, attn_tc::kBM, 128));
    MSE_CHECK(encode_tmap_3d(&tm16, qkv, attn::kDH, 3 * H, (uint64_t)T, attn::kDH * 2, (uint64_t)3 * D * 2, 16, attn_tc::kBM, 32));
    attn_tc::Params ap{};
    ap.S = S; ap.H = H; ap.B = B;
    ap.q_items = (S + 2 * attn_tc::kBM - 1) / (2 * attn_tc::kBM);
    ap.n_blocks = (S + attn_tc::kBN - 1) / attn_tc::kBN;
    ap.n_items = B * H * ap.q_items;
    ap.scale_log2e = (1.0f / sqrtf((float)attn::kDH)) * 1.4426950408889634f;
    ap.debug = mode & 255;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = std::min(ap.n_items, sm_count(device));
    for (int i = 0; i < 2; i++) attn_tc::k_mha_tc<<<grid, attn_tc::kThreads, attn_tc::kSmemBytes>>>(tm64, tm16, out, ap);
    if (mode & 256) {
        // phase timeline of CTA 0 (clock64 at the kernel's trace points), printed as cycles since the first stamp
        long long *tr = nullptr;
        MSE_CUDA(cudaMalloc(&tr, 5 * 256 * 8));
        cudaMemset(tr, 0, 5 * 256 * 8);
        ap.trace = tr;
        attn_tc::k_mha_tc<<<grid, attn_tc::kThreads, attn_tc::kSmemBytes>>>(tm64, tm16, out, ap);
        std::vector<long long> h(5 * 256);
        cudaMemcpy(h.data(), tr, h.size() * 8, cudaMemcpyDeviceToHost);
        cudaFree(tr);
        ap.trace = nullptr;
        long long t0 = 0;
        for (long long v : h) if (v && (!t0 || v < t0)) t0 = v;
        const char *names[4] = {"softmax_A", "softmax_B", "issuer_A", "issuer_B"};
        for (int r = 0; r < 4; r++) {
            printf("trace %s:", names[r]);
            for (int i = 0; i < 100 && h[r * 256 + i]; i++) printf(" %lld", h[r * 256 + i] - t0);
            printf("\n");
        }
    }
    cudaEventRecord(e0);
    for (int i = 0; i < iters; i++) attn_tc::k_mha_tc<<<grid, attn_tc::kThreads, attn_tc::kSmemBytes>>>(tm64, tm16, out, ap);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(qkv); cudaFree(out);
    MSE_REQUIRE(err == cudaSuccess, MSE_ERR_CUDA, "debug_attention: %s", cudaGetErrorString(err));
    *ms_out = ms / iters;
    return MSE_OK;
}
