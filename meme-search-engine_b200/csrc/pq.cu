// Product quantizer (OPQ) and RabitQ codecs on the GPU.
//
//   ProductQuantizer          diskann/src/vector.rs:308-406
//     apply_transform         :319-329   y = T x            (T row-major D x D)
//     quantize_batch          :331-364   per subspace argmax <y_m, c_j,m>, first maximum wins
//     preprocess_query        :367-384   LUT[m][j] = <(T q)_m, c_j,m>
//     asymmetric_dot_product  :387-405   f32 sum of LUT entries in chunk order, then trunc(x * 2^32)
//   opq.msgpack               diskann/aopq_train.py:87-93 (map: centroids, transform, n_dims_per_code, n_dims)
//   RabitQ                    diskann/rabitq.py:8-48 (centre, normalise, 512-row orthogonal projection, sign bits,
//                             estimate = |o| * <o_bar, P q> * dots + <mean, q>)
//
// Summation orders are the CPU oracle's (sequential fused multiply-adds over k), so LUTs, codes and ADC scores are
// bit-identical to it; the ADC itself is bit-identical to the reference by construction (f32 adds in chunk order).
#include "internal.h"
#include "fastdot.cuh"
#include <math.h>
#include <algorithm>
#include <string>

struct mse_pq {
    int device = 0;
    uint32_t D = 0, S = 0, M = 0, C = 0;   // dims, dims per code, chunks, centroids
    float *Tt = nullptr;        // transform transposed: Tt[k][i] = T[i][k]  (coalesced over outputs i)
    float *cent_t = nullptr;    // centroids regrouped: cent_t[(m*S + k)*C + j] = centroids[j][m*S + k]
};

struct mse_rabitq {
    int device = 0;
    uint32_t D = 0, O = 0;      // input dims, output dims (bits)
    float *mean = nullptr;      // [D]
    float *Pt = nullptr;        // projection transposed: Pt[k][o] = P[o][k]
};

namespace mse {

// y[v][i] = sum_k T[i][k] x[v][k], sequential fmaf over k (the oracle's order).  grid (ceil(D/256), n_vec)
__global__ void __launch_bounds__(256) k_pq_transform(const float *__restrict__ Tt, const float *__restrict__ x, float *__restrict__ y, uint32_t D) {
    extern __shared__ float sx[];
    const uint32_t v = blockIdx.y;
    for (uint32_t k = threadIdx.x; k < D; k += blockDim.x) sx[k] = x[(size_t)v * D + k];
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D) return;
    float acc = 0.f;
    for (uint32_t k = 0; k < D; k++) acc = fmaf(Tt[(size_t)k * D + i], sx[k], acc);
    y[(size_t)v * D + i] = acc;
}

// sims[m][j] = <y_m, c_j,m>; mode 0: write the LUT row (preprocess_query); mode 1: argmax -> code (quantize_batch).
// grid (M, n_vec), block C threads (C <= 256)
__global__ void __launch_bounds__(256) k_pq_subspace(const float *__restrict__ cent_t, const float *__restrict__ y, uint32_t D, uint32_t S,
                                                     uint32_t C, float *__restrict__ lut, uint8_t *__restrict__ codes, int mode) {
    __shared__ float sy[64];
    __shared__ float sbest[256];
    __shared__ int sidx[256];
    const uint32_t m = blockIdx.x, v = blockIdx.y, M = gridDim.x, j = threadIdx.x;
    if (j < S) sy[j] = y[(size_t)v * D + m * S + j];
    __syncthreads();
    float acc = -INFINITY;
    if (j < C) {
        acc = 0.f;
        for (uint32_t k = 0; k < S; k++) acc = fmaf(sy[k], cent_t[((size_t)m * S + k) * C + j], acc);
    }
    if (mode == 0) {
        if (j < C) lut[((size_t)v * M + m) * C + j] = acc;
        return;
    }
    // first maximum wins; a row of NaN / -inf keeps code 0 (vector.rs:352-360: `score > best` from -inf)
    sbest[j] = (j < C && acc > -INFINITY) ? acc : -INFINITY;
    sidx[j] = (j < C && acc > -INFINITY) ? (int)j : 0x7fffffff;
    __syncthreads();
    for (uint32_t o = 128; o; o >>= 1) {
        if (j < o) {
            const float a = sbest[j], b = sbest[j + o];
            const int ia = sidx[j], ib = sidx[j + o];
            if (b > a || (b == a && ib < ia)) { sbest[j] = b; sidx[j] = ib; }
        }
        __syncthreads();
    }
    if (j == 0) codes[(size_t)v * M + m] = (uint8_t)(sidx[0] == 0x7fffffff ? 0 : sidx[0]);
}

// scores[v] = trunc(2^32 * sum_m LUT[m][code[v][m]]), f32 adds in chunk order (vector.rs:393-404)
__global__ void __launch_bounds__(256) k_pq_adc(const float *__restrict__ lut, uint32_t M, uint32_t C, const uint8_t *__restrict__ codes,
                                                uint64_t n, long long *__restrict__ out) {
    extern __shared__ float sl[];
    for (uint32_t i = threadIdx.x; i < M * C; i += blockDim.x) sl[i] = lut[i];
    __syncthreads();
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t *c = codes + v * M;
        float acc = 0.f;
        for (uint32_t m = 0; m < M; m++) acc += sl[m * C + c[m]];
        out[v] = fast_dot_fix(acc);
    }
}

// ---- RabitQ (rabitq.py:14-36): one CTA per vector.  c = x - mean; norm = |c|; xs = P (c / norm); bit o = xs_o > 0;
// dots = sum_o (1/sqrt(D)) |xs_o|
__global__ void __launch_bounds__(256) k_rabitq_encode(const float *__restrict__ mean, const float *__restrict__ Pt, const __half *__restrict__ x,
                                                       uint32_t D, uint32_t O, uint8_t *__restrict__ codes, float *__restrict__ norms,
                                                       float *__restrict__ dots) {
    extern __shared__ float sc[];  // centred, normalised vector
    __shared__ float red[8];
    __shared__ float s_norm;
    const uint64_t v = blockIdx.x;
    float part = 0.f;
    for (uint32_t k = threadIdx.x; k < D; k += blockDim.x) {
        const float c = __half2float(x[v * D + k]) - mean[k];
        sc[k] = c;
        part = fmaf(c, c, part);
    }
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; i++) t += red[i];
        s_norm = sqrtf(t);
    }
    __syncthreads();
    const float inv = 1.0f / s_norm;
    const float scale = rsqrtf((float)D);
    float dpart = 0.f;
    for (uint32_t o = threadIdx.x; o < O; o += blockDim.x) {
        float acc = 0.f;
        for (uint32_t k = 0; k < D; k++) acc = fmaf(Pt[(size_t)k * O + o], sc[k] * inv, acc);
        const bool bit = acc > 0.f;
        dpart += scale * fabsf(acc);
        const unsigned m = __ballot_sync(0xffffffffu, bit);   // O % 32 == 0: every lane of the warp is active here
        if ((threadIdx.x & 31) == 0) ((uint32_t *)codes)[v * (O / 32) + o / 32] = m;  // bit i of byte b = output 8b + i
    }
    __syncthreads();
    for (int o = 16; o; o >>= 1) dpart += __shfl_xor_sync(0xffffffffu, dpart, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dpart;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; i++) t += red[i];
        dots[v] = t;
        norms[v] = s_norm;
    }
}

// query side (rabitq.py:42-46): qt = P q, mq = <mean, q>.  One CTA per query.  out: [nq][O + 1] (qt..., mq)
__global__ void __launch_bounds__(256) k_rabitq_query(const float *__restrict__ mean, const float *__restrict__ Pt, const float *__restrict__ q,
                                                      uint32_t D, uint32_t O, float *__restrict__ out) {
    extern __shared__ float sq[];
    __shared__ float red[8];
    const uint32_t v = blockIdx.x;
    float part = 0.f;
    for (uint32_t k = threadIdx.x; k < D; k += blockDim.x) {
        const float a = q[(size_t)v * D + k];
        sq[k] = a;
        part = fmaf(a, mean[k], part);
    }
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    for (uint32_t o = threadIdx.x; o < O; o += blockDim.x) {
        float acc = 0.f;
        for (uint32_t k = 0; k < D; k++) acc = fmaf(Pt[(size_t)k * O + o], sq[k], acc);
        out[(size_t)v * (O + 1) + o] = acc;
    }
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; i++) t += red[i];
        out[(size_t)v * (O + 1) + O] = t;
    }
}

// estimate[v] = norm_v * (scale * sum_o sign_o qt_o) * dots_v + mq   (rabitq.py:47-48), one query against n codes.
// The signed sum runs over bytes with a 256-entry table per byte -- the same gather as the PQ ADC.
__global__ void __launch_bounds__(256) k_rabitq_estimate(const float *__restrict__ qtm, uint32_t O, uint32_t D, const uint8_t *__restrict__ codes,
                                                         const float *__restrict__ norms, const float *__restrict__ dots, uint64_t n,
                                                         float *__restrict__ out) {
    extern __shared__ float tab[];  // [O/8][256]
    const uint32_t nb = O / 8;
    for (uint32_t i = threadIdx.x; i < nb * 256; i += blockDim.x) {
        const uint32_t b = i >> 8, val = i & 255;
        float s = 0.f;
        for (int j = 0; j < 8; j++) s += ((val >> j) & 1) ? qtm[b * 8 + j] : -qtm[b * 8 + j];
        tab[i] = s;
    }
    __syncthreads();
    const float scale = rsqrtf((float)D), mq = qtm[O];
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t *c = codes + v * nb;
        float acc = 0.f;
        for (uint32_t b = 0; b < nb; b++) acc += tab[b * 256 + c[b]];
        out[v] = norms[v] * (scale * acc) * dots[v] + mq;
    }
}

// per-query byte tables of the RabitQ estimator for the graph traversal (graph.cu k_beam_search): lut[b][v] =
// (1/sqrt(D)) * sum_{j<8} (+-) qt[8b+j], bias = <mean, q>.  One CTA per query.  qtm: [nq][O + 1] from k_rabitq_query.
__global__ void __launch_bounds__(256) k_rabitq_lut(const float *__restrict__ qtm, uint32_t O, uint32_t D, float *__restrict__ lut,
                                                    float *__restrict__ bias) {
    const uint32_t nb = O / 8, v = blockIdx.x;
    const float *qt = qtm + (size_t)v * (O + 1);
    const float scale = rsqrtf((float)D);
    for (uint32_t i = threadIdx.x; i < nb * 256; i += blockDim.x) {
        const uint32_t b = i >> 8, val = i & 255;
        float s = 0.f;
        for (int j = 0; j < 8; j++) s += ((val >> j) & 1) ? qt[b * 8 + j] : -qt[b * 8 + j];
        lut[(size_t)v * nb * 256 + i] = scale * s;
    }
    if (threadIdx.x == 0) bias[v] = qt[O];
}

// scales[v] = norms[v] * dots[v] (diskann/rabitq.py:48 as written) or norms[v] / dots[v] (the RabitQ paper's estimator)
__global__ void k_rabitq_scales(const float *__restrict__ norms, const float *__restrict__ dots, uint64_t n, int divide, float *__restrict__ out) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) out[v] = divide ? norms[v] / dots[v] : norms[v] * dots[v];
}

// ---- RabitQ "training" (diskann/rabitq.py:11-28): dataset mean + the first output_dims rows of a random orthogonal matrix

// column means of fp16 rows [n][D], f64 accumulation: block = 32 columns x 8 row lanes
__global__ void __launch_bounds__(256) k_col_mean(const __half *__restrict__ x, uint64_t n, uint32_t D, float *__restrict__ mean) {
    __shared__ double part[8][33];
    const uint32_t col = blockIdx.x * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5;
    double acc = 0.0;
    if (col < D)
        for (uint64_t r = ry; r < n; r += 8) acc += (double)__half2float(x[r * D + col]);
    part[ry][threadIdx.x & 31] = acc;
    __syncthreads();
    if (ry == 0 && col < D) {
        double t = 0.0;
        for (int i = 0; i < 8; i++) t += part[i][threadIdx.x & 31];
        mean[col] = (float)(t / (double)n);
    }
}

// standard normal entries from a counter-based generator (splitmix64 of seed and index, Box-Muller)
__global__ void k_gauss_fill(float *__restrict__ g, uint64_t n, uint64_t seed) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto mix = [](uint64_t z) { z += 0x9e3779b97f4a7c15ull; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); };
    const uint64_t a = mix(seed * 0x2545f4914f6cdd1dull + 2 * i), b = mix(seed * 0x2545f4914f6cdd1dull + 2 * i + 1);
    const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740993.0), u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
    g[i] = (float)(sqrt(-2.0 * log(u1)) * cospi(2.0 * u2));
}

// one Gram-Schmidt step: rows j > i lose their component along row i.  One block per row j.
__global__ void __launch_bounds__(256) k_mgs_step(float *__restrict__ g, uint32_t D, uint32_t i) {
    __shared__ double red[2][8];
    const uint32_t j = i + 1 + blockIdx.x;
    const float *ri = g + (size_t)i * D;
    float *rj = g + (size_t)j * D;
    double dot = 0.0, nn = 0.0;
    for (uint32_t k = threadIdx.x; k < D; k += blockDim.x) { const double a = ri[k]; dot += a * (double)rj[k]; nn += a * a; }
    for (int o = 16; o; o >>= 1) { dot += __shfl_xor_sync(0xffffffffu, dot, o); nn += __shfl_xor_sync(0xffffffffu, nn, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = dot; red[1][threadIdx.x >> 5] = nn; }
    __syncthreads();
    dot = 0.0; nn = 0.0;
    for (int w = 0; w < 8; w++) { dot += red[0][w]; nn += red[1][w]; }
    const double c = nn > 0.0 ? dot / nn : 0.0;
    for (uint32_t k = threadIdx.x; k < D; k += blockDim.x) rj[k] = (float)((double)rj[k] - c * (double)ri[k]);
}

// unit rows, written transposed: Pt[k][o] = P[o][k]
__global__ void __launch_bounds__(256) k_normalise_transpose(const float *__restrict__ g, uint32_t O, uint32_t D, float *__restrict__ Pt) {
    __shared__ double red[8];
    const uint32_t o = blockIdx.x;
    const float *r = g + (size_t)o * D;
    double nn = 0.0;
    for (uint32_t k = threadIdx.x; k < D; k += blockDim.x) nn += (double)r[k] * r[k];
    for (int s = 16; s; s >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = nn;
    __syncthreads();
    nn = 0.0;
    for (int w = 0; w < 8; w++) nn += red[w];
    const double inv = 1.0 / sqrt(nn);
    for (uint32_t k = threadIdx.x; k < D; k += blockDim.x) Pt[(size_t)k * O + o] = (float)((double)r[k] * inv);
}

__global__ void k_untranspose(const float *__restrict__ Pt, uint32_t O, uint32_t D, float *__restrict__ P) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (uint64_t)O * D) { const uint32_t o = (uint32_t)(i / D), k = (uint32_t)(i % D); P[i] = Pt[(size_t)k * O + o]; }
}

// ---- minimal msgpack reader for opq.msgpack / rabitq.msgpack (maps of str -> int | float array)
struct MpReader {
    const uint8_t *p, *end;
    bool ok = true;
    uint8_t u8() { if (p >= end) { ok = false; return 0; } return *p++; }
    uint64_t be(int n) { uint64_t v = 0; for (int i = 0; i < n; i++) v = (v << 8) | u8(); return v; }
    bool read_len(uint8_t tag, uint8_t fix_lo, uint8_t fix_hi, uint8_t t16, uint8_t t32, uint64_t &len) {
        if (tag >= fix_lo && tag <= fix_hi) { len = tag - fix_lo; return true; }
        if (tag == t16) { len = be(2); return true; }
        if (tag == t32) { len = be(4); return true; }
        return false;
    }
    bool number(double &out) {
        const uint8_t t = u8();
        if (t <= 0x7f) { out = t; return true; }
        if (t >= 0xe0) { out = (int8_t)t; return true; }
        switch (t) {
            case 0xca: { uint32_t b = (uint32_t)be(4); float f; memcpy(&f, &b, 4); out = f; return true; }
            case 0xcb: { uint64_t b = be(8); double d; memcpy(&d, &b, 8); out = d; return true; }
            case 0xcc: out = (double)be(1); return true;
            case 0xcd: out = (double)be(2); return true;
            case 0xce: out = (double)be(4); return true;
            case 0xcf: out = (double)be(8); return true;
            case 0xd0: out = (double)(int8_t)be(1); return true;
            case 0xd1: out = (double)(int16_t)be(2); return true;
            case 0xd2: out = (double)(int32_t)be(4); return true;
            case 0xd3: out = (double)(int64_t)be(8); return true;
        }
        return false;
    }
};

static int parse_codec_msgpack(const uint8_t *buf, size_t len, std::vector<std::pair<std::string, std::vector<float>>> &arrays,
                               std::vector<std::pair<std::string, double>> &scalars) {
    MpReader r{buf, buf + len};
    uint64_t n;
    MSE_REQUIRE(r.read_len(r.u8(), 0x80, 0x8f, 0xde, 0xdf, n), MSE_ERR_INVALID, "codec msgpack: top level is not a map");
    for (uint64_t i = 0; i < n && r.ok; i++) {
        uint64_t kl;
        const uint8_t kt = r.u8();
        if (!(r.read_len(kt, 0xa0, 0xbf, 0xda, 0xdb, kl) || (kt == 0xd9 && ((kl = r.be(1)), true)))) { set_error("codec msgpack: key is not a string"); return MSE_ERR_INVALID; }
        MSE_REQUIRE(r.ok && kl <= (uint64_t)(r.end - r.p), MSE_ERR_INVALID, "codec msgpack: truncated key");
        std::string key((const char *)r.p, kl);
        r.p += kl;
        MSE_REQUIRE(r.p < r.end, MSE_ERR_INVALID, "codec msgpack: no value after key '%s'", key.c_str());
        const uint8_t vt = *r.p;
        uint64_t al;
        if ((vt >= 0x90 && vt <= 0x9f) || vt == 0xdc || vt == 0xdd) {
            r.u8();
            r.read_len(vt, 0x90, 0x9f, 0xdc, 0xdd, al);
            // every element takes at least one byte: a length beyond the remaining bytes is a corrupt (or hostile) header
            MSE_REQUIRE(r.ok && al <= (uint64_t)(r.end - r.p), MSE_ERR_INVALID, "codec msgpack: array '%s' claims %llu elements, %llu bytes remain",
                        key.c_str(), (unsigned long long)al, (unsigned long long)(r.end - r.p));
            std::vector<float> v;
            v.reserve(al);
            for (uint64_t j = 0; j < al; j++) {
                double d;
                MSE_REQUIRE(r.number(d) && r.ok, MSE_ERR_INVALID, "codec msgpack: array '%s' holds a non-number", key.c_str());
                v.push_back((float)d);
            }
            arrays.emplace_back(key, std::move(v));
        } else {
            double d;
            MSE_REQUIRE(r.number(d), MSE_ERR_INVALID, "codec msgpack: value of '%s' is neither a number nor an array", key.c_str());
            scalars.emplace_back(key, d);
        }
    }
    MSE_REQUIRE(r.ok, MSE_ERR_INVALID, "codec msgpack: truncated");
    return MSE_OK;
}

}  // namespace mse

using namespace mse;

// ================================================================== C ABI: ProductQuantizer

MSE_API void mse_pq_destroy(mse_pq *pq) {
    if (!pq) return;
    cudaSetDevice(pq->device);
    cudaFree(pq->Tt);
    cudaFree(pq->cent_t);
    delete pq;
}

MSE_API int mse_pq_create(const float *centroids, const float *transform, uint32_t n_dims, uint32_t n_dims_per_code, uint32_t n_centroids,
                          int device, mse_pq **out) {
    MSE_REQUIRE(out && centroids && transform, MSE_ERR_INVALID, "pq_create: NULL argument");
    *out = nullptr;
    MSE_REQUIRE(n_dims >= 1 && n_dims_per_code >= 1 && n_dims_per_code <= 64 && n_dims % n_dims_per_code == 0 && n_centroids >= 1 && n_centroids <= 256,
                MSE_ERR_INVALID, "pq_create: n_dims=%u n_dims_per_code=%u n_centroids=%u unsupported (vector.rs:337 asserts <= 256 centroids)", n_dims,
                n_dims_per_code, n_centroids);
    MSE_CHECK(use_device(device));
    mse_pq *pq = new mse_pq();
    pq->device = device; pq->D = n_dims; pq->S = n_dims_per_code; pq->M = n_dims / n_dims_per_code; pq->C = n_centroids;
    const size_t D = n_dims, C = n_centroids;
    std::vector<float> tt(D * D), ct(D * C);
    for (size_t i = 0; i < D; i++)
        for (size_t k = 0; k < D; k++) tt[k * D + i] = transform[i * D + k];
    for (size_t j = 0; j < C; j++)
        for (size_t k = 0; k < D; k++) ct[k * C + j] = centroids[j * D + k];
    if (cudaMalloc(&pq->Tt, D * D * 4) != cudaSuccess || cudaMalloc(&pq->cent_t, D * C * 4) != cudaSuccess) {
        (void)cudaGetLastError();
        mse_pq_destroy(pq);
        set_error("pq_create: device allocation failed");
        return MSE_ERR_OOM;
    }
    cudaMemcpy(pq->Tt, tt.data(), D * D * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(pq->cent_t, ct.data(), D * C * 4, cudaMemcpyHostToDevice);
    *out = pq;
    return MSE_OK;
}

// rmp_serde::from_slice::<ProductQuantizer> on opq.msgpack (vector.rs:456; aopq_train.py:87-93)
MSE_API int mse_pq_load(const uint8_t *msgpack, size_t len, int device, mse_pq **out) {
    MSE_REQUIRE(msgpack && out, MSE_ERR_INVALID, "pq_load: NULL argument");
    std::vector<std::pair<std::string, std::vector<float>>> arrays;
    std::vector<std::pair<std::string, double>> scalars;
    MSE_CHECK(parse_codec_msgpack(msgpack, len, arrays, scalars));
    const std::vector<float> *cent = nullptr, *tr = nullptr;
    uint32_t nd = 0, ndpc = 0;
    for (auto &a : arrays) { if (a.first == "centroids") cent = &a.second; if (a.first == "transform") tr = &a.second; }
    for (auto &s : scalars) { if (s.first == "n_dims") nd = (uint32_t)s.second; if (s.first == "n_dims_per_code") ndpc = (uint32_t)s.second; }
    MSE_REQUIRE(cent && tr && nd && ndpc && tr->size() == (size_t)nd * nd && cent->size() % nd == 0, MSE_ERR_INVALID,
                "pq_load: msgpack lacks centroids / transform / n_dims / n_dims_per_code or their sizes disagree");
    return mse_pq_create(cent->data(), tr->data(), nd, ndpc, (uint32_t)(cent->size() / nd), device, out);
}

MSE_API int mse_pq_info(const mse_pq *pq, uint32_t out[4]) {
    MSE_REQUIRE(pq && out, MSE_ERR_INVALID, "pq_info: NULL argument");
    out[0] = pq->D; out[1] = pq->S; out[2] = pq->M; out[3] = pq->C;
    return MSE_OK;
}

static int pq_run(mse_pq *pq, const float *x, uint64_t n, float *y_out, float *lut_out, uint8_t *codes_out) {
    MSE_CHECK(use_device(pq->device));
    const size_t D = pq->D;
    DevBuf bx, by, bl, bc;
    int rc = MSE_OK;
    const uint64_t step = 16384;
    do {
        if ((rc = bx.ensure(step * D * 4)) || (rc = by.ensure(step * D * 4))) break;
        if (lut_out && (rc = bl.ensure(step * pq->M * pq->C * 4))) break;
        if (codes_out && (rc = bc.ensure(step * pq->M))) break;
        for (uint64_t v0 = 0; v0 < n && rc == MSE_OK; v0 += step) {
            const uint32_t m = (uint32_t)std::min<uint64_t>(step, n - v0);
            cudaMemcpy(bx.p, x + v0 * D, (size_t)m * D * 4, cudaMemcpyHostToDevice);
            k_pq_transform<<<dim3((uint32_t)(D + 255) / 256, m), 256, D * 4>>>(pq->Tt, bx.as<float>(), by.as<float>(), (uint32_t)D);
            count_launch();
            if (y_out) cudaMemcpy(y_out + v0 * D, by.p, (size_t)m * D * 4, cudaMemcpyDeviceToHost);
            if (lut_out) {
                k_pq_subspace<<<dim3(pq->M, m), 256>>>(pq->cent_t, by.as<float>(), (uint32_t)D, pq->S, pq->C, bl.as<float>(), nullptr, 0);
                count_launch();
                cudaMemcpy(lut_out + v0 * pq->M * pq->C, bl.p, (size_t)m * pq->M * pq->C * 4, cudaMemcpyDeviceToHost);
            }
            if (codes_out) {
                k_pq_subspace<<<dim3(pq->M, m), 256>>>(pq->cent_t, by.as<float>(), (uint32_t)D, pq->S, pq->C, nullptr, bc.as<uint8_t>(), 1);
                count_launch();
                cudaMemcpy(codes_out + v0 * pq->M, bc.p, (size_t)m * pq->M, cudaMemcpyDeviceToHost);
            }
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { set_error("pq: %s", cudaGetErrorString(e)); rc = MSE_ERR_CUDA; }
        }
    } while (0);
    bx.release(); by.release(); bl.release(); bc.release();
    return rc;
}

MSE_API int mse_pq_apply_transform(mse_pq *pq, const float *x, uint64_t n, float *y) {
    MSE_REQUIRE(pq && x && y, MSE_ERR_INVALID, "pq_apply_transform: NULL argument");
    return pq_run(pq, x, n, y, nullptr, nullptr);
}
MSE_API int mse_pq_encode(mse_pq *pq, const float *x, uint64_t n, uint8_t *codes) {
    MSE_REQUIRE(pq && x && codes, MSE_ERR_INVALID, "pq_encode: NULL argument");
    return pq_run(pq, x, n, nullptr, nullptr, codes);
}
MSE_API int mse_pq_preprocess_query(mse_pq *pq, const float *q, uint32_t nq, float *lut) {
    MSE_REQUIRE(pq && q && lut, MSE_ERR_INVALID, "pq_preprocess_query: NULL argument");
    return pq_run(pq, q, nq, nullptr, lut, nullptr);
}

MSE_API int mse_pq_adc(mse_pq *pq, const float *lut, const uint8_t *codes, uint64_t n, int64_t *scores) {
    MSE_REQUIRE(pq && lut && codes && scores, MSE_ERR_INVALID, "pq_adc: NULL argument");
    if (n == 0) return MSE_OK;
    MSE_CHECK(use_device(pq->device));
    const size_t lb = (size_t)pq->M * pq->C * 4;
    MSE_CUDA(cudaFuncSetAttribute(k_pq_adc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lb));
    DevBuf bl, bc, bs;
    int rc = MSE_OK;
    do {
        if ((rc = bl.ensure(lb)) || (rc = bc.ensure(n * pq->M)) || (rc = bs.ensure(n * 8))) break;
        cudaMemcpy(bl.p, lut, lb, cudaMemcpyHostToDevice);
        cudaMemcpy(bc.p, codes, n * pq->M, cudaMemcpyHostToDevice);
        uint32_t blocks = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)sm_count(pq->device) * 4);
        k_pq_adc<<<blocks, 256, lb>>>(bl.as<float>(), pq->M, pq->C, bc.as<uint8_t>(), n, bs.as<long long>());
        count_launch();
        cudaError_t e = cudaMemcpy(scores, bs.p, n * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("pq_adc: %s", cudaGetErrorString(e)); rc = MSE_ERR_CUDA; }
    } while (0);
    bl.release(); bc.release(); bs.release();
    return rc;
}

// ================================================================== C ABI: RabitQ

MSE_API void mse_rabitq_destroy(mse_rabitq *r) {
    if (!r) return;
    cudaSetDevice(r->device);
    cudaFree(r->mean);
    cudaFree(r->Pt);
    delete r;
}

// rabitq.msgpack content (rabitq.py:62-68): mean [n_dims], transform [output_dims][n_dims]
MSE_API int mse_rabitq_create(const float *mean, const float *transform, uint32_t n_dims, uint32_t output_dims, int device, mse_rabitq **out) {
    MSE_REQUIRE(out && mean && transform, MSE_ERR_INVALID, "rabitq_create: NULL argument");
    *out = nullptr;
    MSE_REQUIRE(n_dims >= 1 && output_dims >= 32 && output_dims % 32 == 0 && output_dims <= 4096, MSE_ERR_INVALID,
                "rabitq_create: n_dims=%u output_dims=%u unsupported (output_dims %% 32 == 0)", n_dims, output_dims);
    MSE_CHECK(use_device(device));
    mse_rabitq *r = new mse_rabitq();
    r->device = device; r->D = n_dims; r->O = output_dims;
    const size_t D = n_dims, O = output_dims;
    std::vector<float> pt(D * O);
    for (size_t o = 0; o < O; o++)
        for (size_t k = 0; k < D; k++) pt[k * O + o] = transform[o * D + k];
    if (cudaMalloc(&r->mean, D * 4) != cudaSuccess || cudaMalloc(&r->Pt, D * O * 4) != cudaSuccess) {
        (void)cudaGetLastError();
        mse_rabitq_destroy(r);
        set_error("rabitq_create: device allocation failed");
        return MSE_ERR_OOM;
    }
    cudaMemcpy(r->mean, mean, D * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(r->Pt, pt.data(), D * O * 4, cudaMemcpyHostToDevice);
    *out = r;
    return MSE_OK;
}

// rabitq.py:11-28 on the GPU, over the rows already in HBM: mean of the first sample_rows rows (the script takes 100 000), and P =
// output_dims orthonormal rows (seeded Gaussian rows, two Gram-Schmidt sweeps -- Haar distributed like the Q of the script's QR)
MSE_API int mse_rabitq_train(mse_index *ix, uint64_t sample_rows, uint32_t output_dims, uint64_t seed, mse_rabitq **out) {
    MSE_REQUIRE(ix && out, MSE_ERR_INVALID, "rabitq_train: NULL argument");
    *out = nullptr;
    MSE_REQUIRE(ix->n > 0, MSE_ERR_STATE, "rabitq_train: the index is empty");
    MSE_REQUIRE(output_dims >= 32 && output_dims % 32 == 0 && output_dims <= ix->d && output_dims <= 4096, MSE_ERR_INVALID,
                "rabitq_train: output_dims=%u must be a multiple of 32 and <= n_dims=%u", output_dims, ix->d);
    MSE_CHECK(use_device(ix->device));
    const uint64_t ns = sample_rows == 0 ? ix->n : std::min<uint64_t>(sample_rows, ix->n);
    const uint32_t D = ix->d, O = output_dims;
    mse_rabitq *r = new mse_rabitq();
    r->device = ix->device; r->D = D; r->O = O;
    float *g = nullptr;
    if (cudaMalloc(&r->mean, (size_t)D * 4) != cudaSuccess || cudaMalloc(&r->Pt, (size_t)D * O * 4) != cudaSuccess ||
        cudaMalloc(&g, (size_t)D * O * 4) != cudaSuccess) {
        (void)cudaGetLastError();
        if (g) cudaFree(g);
        mse_rabitq_destroy(r);
        set_error("rabitq_train: device allocation failed");
        return MSE_ERR_OOM;
    }
    k_col_mean<<<(D + 31) / 32, 256>>>(ix->x, ns, D, r->mean);
    count_launch();
    k_gauss_fill<<<(uint32_t)(((size_t)D * O + 255) / 256), 256>>>(g, (uint64_t)D * O, seed);
    count_launch();
    for (int sweep = 0; sweep < 2; sweep++)
        for (uint32_t i = 0; i + 1 < O; i++) {
            k_mgs_step<<<O - 1 - i, 256>>>(g, D, i);
            count_launch();
        }
    k_normalise_transpose<<<O, 256>>>(g, O, D, r->Pt);
    count_launch();
    const cudaError_t e = cudaDeviceSynchronize();
    cudaFree(g);
    if (e != cudaSuccess) {
        set_error("rabitq_train: %s", cudaGetErrorString(e));
        mse_rabitq_destroy(r);
        return MSE_ERR_CUDA;
    }
    *out = r;
    return MSE_OK;
}

MSE_API int mse_rabitq_info(const mse_rabitq *r, uint32_t out[2]) {
    MSE_REQUIRE(r && out, MSE_ERR_INVALID, "rabitq_info: NULL argument");
    out[0] = r->D; out[1] = r->O;
    return MSE_OK;
}

// mean [n_dims] and transform [output_dims][n_dims] row-major: the two arrays rabitq.msgpack stores (rabitq.py:62-68)
MSE_API int mse_rabitq_export(const mse_rabitq *r, float *mean, float *transform) {
    MSE_REQUIRE(r && mean && transform, MSE_ERR_INVALID, "rabitq_export: NULL argument");
    MSE_CHECK(use_device(r->device));
    DevBuf p;
    MSE_CHECK(p.ensure((size_t)r->D * r->O * 4));
    k_untranspose<<<(uint32_t)(((size_t)r->D * r->O + 255) / 256), 256>>>(r->Pt, r->O, r->D, p.as<float>());
    count_launch();
    cudaError_t e = cudaMemcpy(transform, p.p, (size_t)r->D * r->O * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(mean, r->mean, (size_t)r->D * 4, cudaMemcpyDeviceToHost);
    p.release();
    MSE_REQUIRE(e == cudaSuccess, MSE_ERR_CUDA, "rabitq_export: %s", cudaGetErrorString(e));
    return MSE_OK;
}

MSE_API int mse_rabitq_load(const uint8_t *msgpack, size_t len, int device, mse_rabitq **out) {
    MSE_REQUIRE(msgpack && out, MSE_ERR_INVALID, "rabitq_load: NULL argument");
    std::vector<std::pair<std::string, std::vector<float>>> arrays;
    std::vector<std::pair<std::string, double>> scalars;
    MSE_CHECK(parse_codec_msgpack(msgpack, len, arrays, scalars));
    const std::vector<float> *mean = nullptr, *tr = nullptr;
    uint32_t nd = 0, od = 0;
    for (auto &a : arrays) { if (a.first == "mean") mean = &a.second; if (a.first == "transform") tr = &a.second; }
    for (auto &s : scalars) { if (s.first == "n_dims") nd = (uint32_t)s.second; if (s.first == "output_dims") od = (uint32_t)s.second; }
    MSE_REQUIRE(mean && tr && nd && od && mean->size() == nd && tr->size() == (size_t)nd * od, MSE_ERR_INVALID,
                "rabitq_load: msgpack lacks mean / transform / n_dims / output_dims or their sizes disagree");
    return mse_rabitq_create(mean->data(), tr->data(), nd, od, device, out);
}

// codes: [n][output_dims/8] (bit i of byte b = sign of output 8b+i), norms/dots: [n]  (rabitq.py:14-36)
MSE_API int mse_rabitq_encode(mse_rabitq *r, const uint16_t *x_f16, uint64_t n, uint8_t *codes, float *norms, float *dots) {
    MSE_REQUIRE(r && x_f16 && codes && norms && dots, MSE_ERR_INVALID, "rabitq_encode: NULL argument");
    if (n == 0) return MSE_OK;
    MSE_CHECK(use_device(r->device));
    DevBuf bx, bc, bn, bd;
    int rc = MSE_OK;
    const uint64_t step = 1 << 16;
    do {
        if ((rc = bx.ensure(step * r->D * 2)) || (rc = bc.ensure(step * r->O / 8)) || (rc = bn.ensure(step * 4)) || (rc = bd.ensure(step * 4))) break;
        for (uint64_t v0 = 0; v0 < n && rc == MSE_OK; v0 += step) {
            const uint32_t m = (uint32_t)std::min<uint64_t>(step, n - v0);
            cudaMemcpy(bx.p, x_f16 + v0 * r->D, (size_t)m * r->D * 2, cudaMemcpyHostToDevice);
            k_rabitq_encode<<<m, 256, r->D * 4>>>(r->mean, r->Pt, bx.as<__half>(), r->D, r->O, bc.as<uint8_t>(), bn.as<float>(), bd.as<float>());
            count_launch();
            cudaMemcpy(codes + v0 * r->O / 8, bc.p, (size_t)m * r->O / 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(norms + v0, bn.p, (size_t)m * 4, cudaMemcpyDeviceToHost);
            cudaError_t e = cudaMemcpy(dots + v0, bd.p, (size_t)m * 4, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { set_error("rabitq_encode: %s", cudaGetErrorString(e)); rc = MSE_ERR_CUDA; }
        }
    } while (0);
    bx.release(); bc.release(); bn.release(); bd.release();
    return rc;
}

// approx_dot (rabitq.py:42-48) of ONE f32 query against n encoded vectors
MSE_API int mse_rabitq_estimate(mse_rabitq *r, const float *q, const uint8_t *codes, const float *norms, const float *dots, uint64_t n,
                                float *estimates) {
    MSE_REQUIRE(r && q && codes && norms && dots && estimates, MSE_ERR_INVALID, "rabitq_estimate: NULL argument");
    if (n == 0) return MSE_OK;
    MSE_CHECK(use_device(r->device));
    const size_t tab = (size_t)(r->O / 8) * 256 * 4;
    MSE_CUDA(cudaFuncSetAttribute(k_rabitq_estimate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab));
    DevBuf bq, bt, bc, bn, bd, bo;
    int rc = MSE_OK;
    do {
        if ((rc = bq.ensure(r->D * 4)) || (rc = bt.ensure((r->O + 1) * 4)) || (rc = bc.ensure(n * r->O / 8)) || (rc = bn.ensure(n * 4)) ||
            (rc = bd.ensure(n * 4)) || (rc = bo.ensure(n * 4)))
            break;
        cudaMemcpy(bq.p, q, r->D * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(bc.p, codes, n * r->O / 8, cudaMemcpyHostToDevice);
        cudaMemcpy(bn.p, norms, n * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(bd.p, dots, n * 4, cudaMemcpyHostToDevice);
        k_rabitq_query<<<1, 256, r->D * 4>>>(r->mean, r->Pt, bq.as<float>(), r->D, r->O, bt.as<float>());
        count_launch();
        uint32_t blocks = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)sm_count(r->device) * 4);
        k_rabitq_estimate<<<blocks, 256, tab>>>(bt.as<float>(), r->O, r->D, bc.as<uint8_t>(), bn.as<float>(), bd.as<float>(), n, bo.as<float>());
        count_launch();
        cudaError_t e = cudaMemcpy(estimates, bo.p, n * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("rabitq_estimate: %s", cudaGetErrorString(e)); rc = MSE_ERR_CUDA; }
    } while (0);
    bq.release(); bt.release(); bc.release(); bn.release(); bd.release(); bo.release();
    return rc;
}

// query side of the estimator as byte tables (for mse_search_beam_scaled): luts [nq][output_dims/8][256], bias [nq]
MSE_API int mse_rabitq_preprocess_query(mse_rabitq *r, const float *q, uint32_t nq, float *luts, float *bias) {
    MSE_REQUIRE(r && q && luts && bias, MSE_ERR_INVALID, "rabitq_preprocess_query: NULL argument");
    if (nq == 0) return MSE_OK;
    MSE_CHECK(use_device(r->device));
    const size_t per = (size_t)(r->O / 8) * 256;
    DevBuf bq, bt, bl, bb;
    int rc = MSE_OK;
    do {
        if ((rc = bq.ensure((size_t)nq * r->D * 4)) || (rc = bt.ensure((size_t)nq * (r->O + 1) * 4)) || (rc = bl.ensure((size_t)nq * per * 4)) ||
            (rc = bb.ensure((size_t)nq * 4)))
            break;
        cudaMemcpy(bq.p, q, (size_t)nq * r->D * 4, cudaMemcpyHostToDevice);
        k_rabitq_query<<<nq, 256, r->D * 4>>>(r->mean, r->Pt, bq.as<float>(), r->D, r->O, bt.as<float>());
        count_launch();
        k_rabitq_lut<<<nq, 256>>>(bt.as<float>(), r->O, r->D, bl.as<float>(), bb.as<float>());
        count_launch();
        cudaMemcpy(luts, bl.p, (size_t)nq * per * 4, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaMemcpy(bias, bb.p, (size_t)nq * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("rabitq_preprocess_query: %s", cudaGetErrorString(e)); rc = MSE_ERR_CUDA; }
    } while (0);
    bq.release(); bt.release(); bl.release(); bb.release();
    return rc;
}

// query side of the estimator in HBM, asynchronous: d_qtm [nq][output_dims + 1] = (P q, <mean, q>) -- what
// mse_search_beam_dev turns into byte tables in shared memory.  d_q: [nq][n_dims] f32.
MSE_API int mse_rabitq_query_dev(mse_rabitq *r, const float *d_q, uint32_t nq, float *d_qtm, void *stream) {
    MSE_REQUIRE(r && d_q && d_qtm, MSE_ERR_INVALID, "rabitq_query_dev: NULL argument");
    if (nq == 0) return MSE_OK;
    MSE_CHECK(use_device(r->device));
    k_rabitq_query<<<nq, 256, r->D * 4, (cudaStream_t)stream>>>(r->mean, r->Pt, d_q, r->D, r->O, d_qtm);
    MSE_LAUNCH_OK();
    return MSE_OK;
}

// Encodes the index's own rows where they lie in HBM (no host round trip): codes go to the index's code store
// (mse_index_set_pq_codes), the per-vector factor norms * dots (estimator 0, diskann/rabitq.py:48) or norms / dots
// (estimator 1, the RabitQ paper) to its code scales (mse_index_set_code_scales).
MSE_API int mse_index_encode_rabitq(mse_index *ix, mse_rabitq *r, int estimator) {
    MSE_REQUIRE(ix && r, MSE_ERR_INVALID, "index_encode_rabitq: NULL argument");
    MSE_REQUIRE(ix->d == r->D && ix->device == r->device, MSE_ERR_INVALID, "index_encode_rabitq: codec (d=%u, device %d) does not match the index (d=%u, device %d)",
                r->D, r->device, ix->d, ix->device);
    MSE_REQUIRE(estimator == 0 || estimator == 1, MSE_ERR_INVALID, "index_encode_rabitq: estimator %d (0 = norms*dots, 1 = norms/dots)", estimator);
    MSE_CHECK(use_device(ix->device));
    const uint64_t n = ix->n;
    const uint32_t cs = r->O / 8;
    if (ix->pq_codes) cudaFree(ix->pq_codes);
    if (ix->code_scale) cudaFree(ix->code_scale);
    ix->pq_codes = nullptr; ix->code_scale = nullptr; ix->code_size = 0;
    MSE_CUDA(cudaMalloc(&ix->pq_codes, std::max<size_t>(n * cs, 16)));
    MSE_CUDA(cudaMalloc(&ix->code_scale, std::max<size_t>(n * 4, 16)));
    ix->code_size = cs;
    if (n == 0) return MSE_OK;
    DevBuf bn, bd;
    int rc = MSE_OK;
    do {
        if ((rc = bn.ensure(n * 4)) || (rc = bd.ensure(n * 4))) break;
        const uint64_t step = 1u << 20;   // grid.x per launch
        for (uint64_t v0 = 0; v0 < n; v0 += step) {
            const uint32_t m = (uint32_t)std::min<uint64_t>(step, n - v0);
            k_rabitq_encode<<<m, 256, r->D * 4>>>(r->mean, r->Pt, ix->x + v0 * ix->d, r->D, r->O, ix->pq_codes + v0 * cs, bn.as<float>() + v0, bd.as<float>() + v0);
            count_launch();
        }
        k_rabitq_scales<<<(uint32_t)((n + 255) / 256), 256>>>(bn.as<float>(), bd.as<float>(), n, estimator, ix->code_scale);
        count_launch();
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { set_error("index_encode_rabitq: %s", cudaGetErrorString(e)); rc = MSE_ERR_CUDA; }
    } while (0);
    bn.release(); bd.release();
    return rc;
}
