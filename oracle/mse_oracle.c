/*
 * mse_oracle.c -- CPU restatement of the meme-search-engine search/build hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  The product (libmse_b200.so)
 * never links, loads or calls anything in here.
 *
 * PARITY UNPINNED: the reference ships no golden vectors, known-answer tests or
 * fixtures for this path, and its own implementation (Rust
 * nightly + faiss + simsimd) cannot be built in this image.  Every function below
 * is a from-scratch restatement of the cited reference lines; nothing is copied.
 *
 * Citations are relative to /root/reference.
 *
 *   orc_fast_dot*            diskann/src/vector.rs:192-306   (accumulator layout + reduction tree, bit-identical)
 *   orc_scale_dot_*          diskann/src/vector.rs:408-416   (Rust `as i64`: truncate, saturate, NaN -> 0)
 *   orc_dot_f64              diskann/src/vector.rs:49-52     (simsimd f16 dot; restated as an f64 sum)
 *   orc_flat_search          src/main.rs:822,900             (FAISS IndexScalarQuantizer QT_fp16 / IP; restated per SURVEY 8c)
 *   orc_nb_*                 diskann/src/lib.rs:73-155       (NeighbourBuffer)
 *   orc_greedy_search        diskann/src/lib.rs:183-211
 *   orc_robust_prune         diskann/src/lib.rs:227-285      (incl. the skip-one quirk at :250)
 *   orc_build_graph          diskann/src/lib.rs:287-324
 *   orc_robust_stitch        diskann/src/lib.rs:326-374
 *   orc_random_fill_graph    diskann/src/lib.rs:376-387
 *   orc_medioid              diskann/src/lib.rs:54-68
 *   orc_pq_*                 diskann/src/vector.rs:319-406
 *   orc_beam_search          src/query_disk_index.rs:83-97,135-212
 *   orc_brute_force_i64      src/query_disk_index.rs:262-273
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(__AVX2__) && defined(__F16C__) && defined(__FMA__)
#include <immintrin.h>
#define ORC_HAVE_AVX2 1
#else
#define ORC_HAVE_AVX2 0
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ fp16 <-> fp32 */

static inline float h2f(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal: renormalise */
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            man &= 0x3ffu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

/* round-to-nearest-even f32 -> f16 (matches half::f16::from_f32 and F16C) */
static inline uint16_t f2h(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t absx = x & 0x7fffffffu;
    if (absx >= 0x7f800000u) { /* inf / nan */
        return (uint16_t)(sign | 0x7c00u | ((absx > 0x7f800000u) ? 0x200u | ((absx >> 13) & 0x3ffu) : 0));
    }
    if (absx >= 0x477ff000u) { /* overflows to inf after rounding */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (absx < 0x33000001u) { /* underflows to zero (<= 2^-25) */
        return (uint16_t)sign;
    }
    int32_t e = (int32_t)(absx >> 23) - 127;
    uint32_t m = (absx & 0x7fffffu) | 0x800000u;
    if (e < -14) { /* subnormal result */
        int shift = -14 - e + 13; /* bits to drop */
        uint32_t r = m >> shift;
        uint32_t rem = m & ((1u << shift) - 1);
        uint32_t halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (r & 1u))) r++;
        return (uint16_t)(sign | r);
    }
    uint32_t r = ((uint32_t)(e + 15) << 10) | ((m >> 13) & 0x3ffu);
    uint32_t rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) r++;
    return (uint16_t)(sign | r);
}

ORC_API float orc_h2f(uint16_t h) { return h2f(h); }
ORC_API uint16_t orc_f2h(float f) { return f2h(f); }

/* ------------------------------------------------------------------ fixed-point scores */

/* Rust `(x * 2^32) as i64`: truncates toward zero, saturates, NaN -> 0 (vector.rs:46,408-416) */
static inline int64_t sat_trunc_f32(float v) {
    if (v != v) return 0;
    if (v >= 9223372036854775808.0f) return INT64_MAX;
    if (v <= -9223372036854775808.0f) return INT64_MIN;
    return (int64_t)v;
}
static inline int64_t sat_trunc_f64(double v) {
    if (v != v) return 0;
    if (v >= 9223372036854775808.0) return INT64_MAX;
    if (v <= -9223372036854775808.0) return INT64_MIN;
    return (int64_t)v;
}
ORC_API int64_t orc_scale_dot_result(float x) { return sat_trunc_f32(x * 4294967296.0f); }
ORC_API int64_t orc_scale_dot_result_f64(double x) { return sat_trunc_f64(x * 4294967296.0); }

/* ------------------------------------------------------------------ fast_dot */

/*
 * Scalar model of vector.rs:255-306.  The reference keeps four 8-lane f32
 * accumulators; per 32-half chunk, accumulator a (0..3) lane j (0..7) gets
 * fma(x[32c+8a+j], y[32c+8a+j], acc).  So partial sum p = 8a+j (0..31) owns the
 * elements d with d%32 == p, accumulated in increasing d with one fused
 * multiply-add each.  Reduction (:291-301): A=acc1+acc2, B=acc3+acc4,
 * hadd(A,B) -> [A0+A1,A2+A3,B0+B1,B2+B3 | A4+A5,A6+A7,B4+B5,B6+B7], lo+hi,
 * then e0+e1+e2+e3 left to right.
 */
static float fast_dot_f32_scalar(const uint16_t *x, const uint16_t *y, size_t n) {
    float p[32];
    for (int i = 0; i < 32; i++) p[i] = 0.0f;
    for (size_t c = 0; c + 32 <= n; c += 32)
        for (int i = 0; i < 32; i++)
            p[i] = fmaf(h2f(x[c + i]), h2f(y[c + i]), p[i]);
    float A[8], B[8];
    for (int j = 0; j < 8; j++) {
        A[j] = p[j] + p[8 + j];
        B[j] = p[16 + j] + p[24 + j];
    }
    float e0 = (A[0] + A[1]) + (A[4] + A[5]);
    float e1 = (A[2] + A[3]) + (A[6] + A[7]);
    float e2 = (B[0] + B[1]) + (B[4] + B[5]);
    float e3 = (B[2] + B[3]) + (B[6] + B[7]);
    return ((e0 + e1) + e2) + e3;
}

ORC_API float orc_fast_dot_f32_scalar(const uint16_t *x, const uint16_t *y, size_t n) {
    return fast_dot_f32_scalar(x, y, n);
}
ORC_API int64_t orc_fast_dot_scalar(const uint16_t *x, const uint16_t *y, size_t n) {
    return sat_trunc_f32(fast_dot_f32_scalar(x, y, n) * 4294967296.0f);
}

#if ORC_HAVE_AVX2
static inline float fast_dot_f32_avx2(const uint16_t *x, const uint16_t *y, size_t n) {
    __m256 a1 = _mm256_setzero_ps(), a2 = a1, a3 = a1, a4 = a1;
    for (size_t c = 0; c + 32 <= n; c += 32) {
        __m256i xv1 = _mm256_loadu_si256((const __m256i *)(x + c));
        __m256i yv1 = _mm256_loadu_si256((const __m256i *)(y + c));
        __m256i xv2 = _mm256_loadu_si256((const __m256i *)(x + c + 16));
        __m256i yv2 = _mm256_loadu_si256((const __m256i *)(y + c + 16));
        a1 = _mm256_fmadd_ps(_mm256_cvtph_ps(_mm256_castsi256_si128(xv1)), _mm256_cvtph_ps(_mm256_castsi256_si128(yv1)), a1);
        a2 = _mm256_fmadd_ps(_mm256_cvtph_ps(_mm256_extracti128_si256(xv1, 1)), _mm256_cvtph_ps(_mm256_extracti128_si256(yv1, 1)), a2);
        a3 = _mm256_fmadd_ps(_mm256_cvtph_ps(_mm256_castsi256_si128(xv2)), _mm256_cvtph_ps(_mm256_castsi256_si128(yv2)), a3);
        a4 = _mm256_fmadd_ps(_mm256_cvtph_ps(_mm256_extracti128_si256(xv2, 1)), _mm256_cvtph_ps(_mm256_extracti128_si256(yv2, 1)), a4);
    }
    __m256 A = _mm256_add_ps(a1, a2);
    __m256 B = _mm256_add_ps(a3, a4);
    __m256 h = _mm256_hadd_ps(A, B);
    __m128 s = _mm_add_ps(_mm256_castps256_ps128(h), _mm256_extractf128_ps(h, 1));
    float e[4];
    _mm_storeu_ps(e, s);
    return ((e[0] + e[1]) + e[2]) + e[3];
}
#endif

static inline float fast_dot_f32(const uint16_t *x, const uint16_t *y, size_t n) {
#if ORC_HAVE_AVX2
    return fast_dot_f32_avx2(x, y, n);
#else
    return fast_dot_f32_scalar(x, y, n);
#endif
}
static inline int64_t fast_dot(const uint16_t *x, const uint16_t *y, size_t n) {
    return sat_trunc_f32(fast_dot_f32(x, y, n) * 4294967296.0f);
}
ORC_API int orc_have_avx2(void) { return ORC_HAVE_AVX2; }
ORC_API int64_t orc_fast_dot(const uint16_t *x, const uint16_t *y, size_t n) { return fast_dot(x, y, n); }
ORC_API void orc_fast_dot_batch(const uint16_t *q, const uint16_t *rows, size_t n_rows, size_t d, int64_t *out) {
    for (size_t i = 0; i < n_rows; i++) out[i] = fast_dot(q, rows + i * d, d);
}

/* vector.rs:49-52: simsimd f16 dot widened to f64 then scaled.  simsimd's own
 * accumulation width is backend-dependent; restated as a plain f64 sum. */
static double dot_f64(const uint16_t *x, const uint16_t *y, size_t n) {
    double s = 0.0;
    for (size_t i = 0; i < n; i++) s += (double)h2f(x[i]) * (double)h2f(y[i]);
    return s;
}
ORC_API int64_t orc_dot(const uint16_t *x, const uint16_t *y, size_t n) {
    return sat_trunc_f64(dot_f64(x, y, n) * 4294967296.0);
}

/* ------------------------------------------------------------------ flat (FAISS QT_fp16 / IP) search */

typedef struct { float s; uint32_t id; } hit_t;

/* "a ranks before b": score desc, id asc */
static inline int hit_before(hit_t a, hit_t b) { return a.s > b.s || (a.s == b.s && a.id < b.id); }

/* fixed-size worst-at-root heap: root = the hit that ranks LAST among the kept k */
static void heap_sift_down(hit_t *h, size_t n, size_t i) {
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, w = i;
        if (l < n && hit_before(h[w], h[l])) w = l;
        if (r < n && hit_before(h[w], h[r])) w = r;
        if (w == i) return;
        hit_t t = h[i]; h[i] = h[w]; h[w] = t;
        i = w;
    }
}
static void heap_push(hit_t *h, size_t *n, size_t k, hit_t v) {
    if (*n < k) {
        size_t i = (*n)++;
        h[i] = v;
        while (i > 0) {
            size_t p = (i - 1) / 2;
            if (hit_before(h[p], h[i])) { hit_t t = h[i]; h[i] = h[p]; h[p] = t; i = p; } else break;
        }
    } else if (hit_before(v, h[0])) {
        h[0] = v;
        heap_sift_down(h, k, 0);
    }
}
static int hit_cmp(const void *a, const void *b) {
    hit_t x = *(const hit_t *)a, y = *(const hit_t *)b;
    return hit_before(x, y) ? -1 : (hit_before(y, x) ? 1 : 0);
}

#if ORC_HAVE_AVX2
/* f32 query x fp16 row, 4 x 8-lane f32 accumulators -- how a CPU scan (FAISS-style) does it; baseline timing only */
static inline float qdot_f32_avx2(const float *q, const uint16_t *x, size_t d) {
    __m256 a0 = _mm256_setzero_ps(), a1 = a0, a2 = a0, a3 = a0;
    size_t i = 0;
    for (; i + 32 <= d; i += 32) {
        a0 = _mm256_fmadd_ps(_mm256_loadu_ps(q + i), _mm256_cvtph_ps(_mm_loadu_si128((const __m128i *)(x + i))), a0);
        a1 = _mm256_fmadd_ps(_mm256_loadu_ps(q + i + 8), _mm256_cvtph_ps(_mm_loadu_si128((const __m128i *)(x + i + 8))), a1);
        a2 = _mm256_fmadd_ps(_mm256_loadu_ps(q + i + 16), _mm256_cvtph_ps(_mm_loadu_si128((const __m128i *)(x + i + 16))), a2);
        a3 = _mm256_fmadd_ps(_mm256_loadu_ps(q + i + 24), _mm256_cvtph_ps(_mm_loadu_si128((const __m128i *)(x + i + 24))), a3);
    }
    __m256 s = _mm256_add_ps(_mm256_add_ps(a0, a1), _mm256_add_ps(a2, a3));
    float e[8];
    _mm256_storeu_ps(e, s);
    float r = ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
    for (; i < d; i++) r += q[i] * h2f(x[i]);
    return r;
}
#endif

/*
 * Exact flat inner-product top-k (SURVEY 8c definition): score_i = f32( sum_d f64(q_d) * f64(f32(x_id)) ),
 * ranked by (score desc, id asc); slots past min(k,n) get id UINT32_MAX, score -inf
 * (FAISS pads labels with -1: src/main.rs:908).
 * mode 0: f64 accumulation (the oracle).  mode 1: f32 AVX2 accumulation (CPU baseline timing).
 */
ORC_API void orc_flat_search(const float *q, size_t nq, const uint16_t *x, size_t n, size_t d, size_t k,
                             uint32_t *ids, float *scores, int mode) {
    if (k == 0) return;
    for (size_t qi = 0; qi < nq; qi++) {
        const float *qv = q + qi * d;
        int nt = 1;
#ifdef _OPENMP
        nt = omp_get_max_threads();
#endif
        if ((size_t)nt > n / 1024 + 1) nt = (int)(n / 1024 + 1);
        hit_t *heaps = (hit_t *)malloc(sizeof(hit_t) * k * (size_t)nt);
        size_t *lens = (size_t *)calloc((size_t)nt, sizeof(size_t));
#pragma omp parallel num_threads(nt)
        {
            int t = 0;
#ifdef _OPENMP
            t = omp_get_thread_num();
#endif
            hit_t *h = heaps + (size_t)t * k;
            size_t len = 0;
            size_t lo = n * (size_t)t / (size_t)nt, hi = n * (size_t)(t + 1) / (size_t)nt;
            for (size_t i = lo; i < hi; i++) {
                const uint16_t *row = x + i * d;
                float s;
                if (mode == 0) {
                    double acc = 0.0;
                    for (size_t j = 0; j < d; j++) acc += (double)qv[j] * (double)h2f(row[j]);
                    s = (float)acc;
                } else {
#if ORC_HAVE_AVX2
                    s = qdot_f32_avx2(qv, row, d);
#else
                    float acc = 0.0f;
                    for (size_t j = 0; j < d; j++) acc += qv[j] * h2f(row[j]);
                    s = acc;
#endif
                }
                hit_t v = { s, (uint32_t)i };
                heap_push(h, &len, k, v);
            }
            lens[t] = len;
        }
        size_t tot = 0;
        for (int t = 0; t < nt; t++) {
            memmove(heaps + tot, heaps + (size_t)t * k, lens[t] * sizeof(hit_t));
            tot += lens[t];
        }
        qsort(heaps, tot, sizeof(hit_t), hit_cmp);
        for (size_t j = 0; j < k; j++) {
            if (j < tot) { ids[qi * k + j] = heaps[j].id; scores[qi * k + j] = heaps[j].s; }
            else { ids[qi * k + j] = UINT32_MAX; scores[qi * k + j] = -INFINITY; }
        }
        free(heaps);
        free(lens);
    }
}

/* f64 scores of one query against listed rows (used to measure rank gaps in parity tests) */
ORC_API void orc_flat_scores_f64(const float *q, const uint16_t *x, size_t d, const uint32_t *row_ids, size_t m, double *out) {
    for (size_t r = 0; r < m; r++) {
        const uint16_t *row = x + (size_t)row_ids[r] * d;
        double acc = 0.0;
        for (size_t j = 0; j < d; j++) acc += (double)q[j] * (double)h2f(row[j]);
        out[r] = acc;
    }
}

/* query_disk_index.rs:262-273: exact i64 scores of every node, descending.  Stable order for equal
 * scores (the reference's sort_unstable leaves tie order unspecified). */
typedef struct { int64_t s; uint32_t id; } ihit_t;
static int ihit_cmp(const void *a, const void *b) {
    const ihit_t *x = (const ihit_t *)a, *y = (const ihit_t *)b;
    if (x->s != y->s) return x->s > y->s ? -1 : 1;
    return x->id < y->id ? -1 : (x->id > y->id ? 1 : 0);
}
ORC_API void orc_brute_force_i64(const uint16_t *q, const uint16_t *x, size_t n, size_t d, size_t k,
                                 uint32_t *ids, int64_t *scores) {
    ihit_t *all = (ihit_t *)malloc(sizeof(ihit_t) * n);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) { all[i].s = fast_dot(q, x + i * d, d); all[i].id = (uint32_t)i; }
    qsort(all, n, sizeof(ihit_t), ihit_cmp);
    for (size_t j = 0; j < k && j < n; j++) { ids[j] = all[j].id; scores[j] = all[j].s; }
    free(all);
}

/* ------------------------------------------------------------------ NeighbourBuffer (lib.rs:73-155) */

typedef struct {
    uint32_t *ids;
    int64_t *scores;
    uint8_t *visited;
    size_t len, cap;
    int64_t next_unvisited; /* -1 = None */
} orc_nb;

ORC_API orc_nb *orc_nb_new(size_t cap) {
    orc_nb *b = (orc_nb *)calloc(1, sizeof(orc_nb));
    b->ids = (uint32_t *)malloc(sizeof(uint32_t) * (cap + 1));
    b->scores = (int64_t *)malloc(sizeof(int64_t) * (cap + 1));
    b->visited = (uint8_t *)malloc(cap + 1);
    b->cap = cap;
    b->next_unvisited = -1;
    return b;
}
ORC_API void orc_nb_free(orc_nb *b) {
    if (!b) return;
    free(b->ids); free(b->scores); free(b->visited); free(b);
}
ORC_API void orc_nb_clear(orc_nb *b) { b->len = 0; b->next_unvisited = -1; }
ORC_API size_t orc_nb_len(const orc_nb *b) { return b->len; }
ORC_API size_t orc_nb_cap(const orc_nb *b) { return b->cap; }
ORC_API const uint32_t *orc_nb_ids(const orc_nb *b) { return b->ids; }
ORC_API const int64_t *orc_nb_scores(const orc_nb *b) { return b->scores; }

/* lib.rs:93-107. returns 1 and writes *out, or 0 when nothing is unvisited */
ORC_API int orc_nb_next_unvisited(orc_nb *b, uint32_t *out) {
    if (b->next_unvisited < 0) return 0;
    size_t cur = (size_t)b->next_unvisited, old = cur;
    b->visited[cur] = 1;
    while (cur < b->len && b->visited[cur]) cur++;
    b->next_unvisited = (cur == b->len) ? -1 : (int64_t)cur;
    *out = b->ids[old];
    return 1;
}

/*
 * lib.rs:117-147.  The position comes from slice::binary_search_by on the descending list with
 * comparator score.cmp(x).  With the (branch-free) std implementation that converges on the last
 * element >= score: G = #(x > score), E = #(x == score); E > 0 -> Ok(G+E-1), else Err(G).
 * Restated as that closed form so a data-parallel implementation can match it exactly.
 */
ORC_API void orc_nb_insert(orc_nb *b, uint32_t id, int64_t score) {
    if (b->len == b->cap && b->len > 0 && b->scores[b->len - 1] > score) return;
    if (b->cap == 0) return;
    size_t lo = 0, hi = b->len; /* first index with x < score */
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (b->scores[mid] >= score) lo = mid + 1; else hi = mid;
    }
    size_t loc = lo; /* = G + E */
    if (loc > 0 && b->scores[loc - 1] == score) loc -= 1;
    if (loc < b->len && b->ids[loc] == id) return;
    size_t move = b->len - loc;
    memmove(b->ids + loc + 1, b->ids + loc, move * sizeof(uint32_t));
    memmove(b->scores + loc + 1, b->scores + loc, move * sizeof(int64_t));
    memmove(b->visited + loc + 1, b->visited + loc, move);
    b->ids[loc] = id; b->scores[loc] = score; b->visited[loc] = 0;
    b->len++;
    if (b->len > b->cap) b->len = b->cap;
    if (b->next_unvisited < 0 || (int64_t)loc < b->next_unvisited) b->next_unvisited = (int64_t)loc;
}

/* ------------------------------------------------------------------ graph + config */

/* fixed-stride adjacency: node i owns adj[i*stride .. i*stride+deg[i]) */
typedef struct {
    uint32_t *adj;
    uint32_t *deg;
    size_t n, stride;
} orc_graph;

/* lib.rs:41-51 */
typedef struct {
    uint64_t r, l, maxc;
    int64_t alpha;
    int32_t saturate_graph;
    uint32_t query_breakpoint;
    uint64_t max_add_per_stitch_iter;
    int64_t query_alpha;
} orc_build_config;

ORC_API orc_graph *orc_graph_new(size_t n, size_t stride) {
    orc_graph *g = (orc_graph *)calloc(1, sizeof(orc_graph));
    g->adj = (uint32_t *)calloc(n * stride + 1, sizeof(uint32_t));
    g->deg = (uint32_t *)calloc(n + 1, sizeof(uint32_t));
    g->n = n; g->stride = stride;
    return g;
}
ORC_API void orc_graph_free(orc_graph *g) { if (g) { free(g->adj); free(g->deg); free(g); } }
ORC_API uint32_t *orc_graph_adj(orc_graph *g) { return g->adj; }
ORC_API uint32_t *orc_graph_deg(orc_graph *g) { return g->deg; }

/* small deterministic generator (the reference uses fastrand; its stream need not be matched: SURVEY 8c) */
static inline uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline uint32_t rand_below(uint64_t *s, uint32_t n) { return (uint32_t)(((splitmix64(s) >> 32) * (uint64_t)n) >> 32); }

/* lib.rs:376-387: uniform neighbours, duplicates rejected, self-loops allowed */
ORC_API void orc_random_fill_graph(orc_graph *g, size_t r, uint64_t seed) {
    if (r > g->stride) r = g->stride;
    for (size_t i = 0; i < g->n; i++) {
        uint64_t s = seed ^ (0x5851f42d4c957f2dull * (uint64_t)(i + 1));
        uint32_t *nb = g->adj + i * g->stride;
        uint32_t dg = g->deg[i];
        size_t distinct_possible = g->n;
        while (dg < r && dg < distinct_possible) {
            uint32_t c = rand_below(&s, (uint32_t)g->n);
            int dup = 0;
            for (uint32_t j = 0; j < dg; j++) if (nb[j] == c) { dup = 1; break; }
            if (!dup) nb[dg++] = c;
        }
        g->deg[i] = dg;
    }
}

/* lib.rs:54-68: running-mean centroid in f32, rounded to fp16, argmax of the f64-scaled dot.
 * Iterator::max_by returns the LAST maximum on ties. */
ORC_API uint32_t orc_medioid(const uint16_t *x, size_t n, size_t d) {
    float *c = (float *)calloc(d, sizeof(float));
    for (size_t i = 0; i < n; i++) {
        float w = 1.0f / (float)(i + 1);
        const uint16_t *row = x + i * d;
        for (size_t j = 0; j < d; j++) c[j] = c[j] + (h2f(row[j]) - c[j]) * w;
    }
    uint16_t *ch = (uint16_t *)malloc(d * sizeof(uint16_t));
    for (size_t j = 0; j < d; j++) ch[j] = f2h(c[j]);
    int64_t best = INT64_MIN; uint32_t besti = 0;
    for (size_t i = 0; i < n; i++) {
        int64_t s = sat_trunc_f64(dot_f64(x + i * d, ch, d) * 4294967296.0);
        if (s >= best) { best = s; besti = (uint32_t)i; }
    }
    free(c); free(ch);
    return besti;
}

/* ------------------------------------------------------------------ Scratch + greedy_search (lib.rs:157-211) */

typedef struct { uint32_t id; int64_t score; } cand_t;

typedef struct {
    uint8_t *visited;       /* membership over [0,n): same semantics as the reference's HashSet<u32> */
    uint32_t *touched; size_t n_touched, touched_cap;
    orc_nb *nb;
    uint32_t *pre; size_t pre_cap;
    cand_t *vlist; size_t vlen, vcap;
    size_t n;
} orc_scratch;

ORC_API orc_scratch *orc_scratch_new(size_t n, size_t l, size_t r) {
    orc_scratch *s = (orc_scratch *)calloc(1, sizeof(orc_scratch));
    s->visited = (uint8_t *)calloc(n + 1, 1);
    s->touched_cap = 1024; s->touched = (uint32_t *)malloc(sizeof(uint32_t) * s->touched_cap);
    s->nb = orc_nb_new(l);
    s->pre_cap = r + 256; s->pre = (uint32_t *)malloc(sizeof(uint32_t) * s->pre_cap);
    s->vcap = l * 8 + 64; s->vlist = (cand_t *)malloc(sizeof(cand_t) * s->vcap);
    s->n = n;
    return s;
}
ORC_API void orc_scratch_free(orc_scratch *s) {
    if (!s) return;
    free(s->visited); free(s->touched); orc_nb_free(s->nb); free(s->pre); free(s->vlist); free(s);
}
ORC_API orc_nb *orc_scratch_nb(orc_scratch *s) { return s->nb; }
ORC_API size_t orc_scratch_visited_len(const orc_scratch *s) { return s->vlen; }
ORC_API void orc_scratch_visited_copy(const orc_scratch *s, uint32_t *ids, int64_t *scores) {
    for (size_t i = 0; i < s->vlen; i++) { ids[i] = s->vlist[i].id; scores[i] = s->vlist[i].score; }
}

static inline int visit_insert(orc_scratch *s, uint32_t id) {
    if (s->visited[id]) return 0;
    s->visited[id] = 1;
    if (s->n_touched == s->touched_cap) {
        s->touched_cap *= 2;
        s->touched = (uint32_t *)realloc(s->touched, sizeof(uint32_t) * s->touched_cap);
    }
    s->touched[s->n_touched++] = id;
    return 1;
}
static inline void visit_clear(orc_scratch *s) {
    for (size_t i = 0; i < s->n_touched; i++) s->visited[s->touched[i]] = 0;
    s->n_touched = 0;
}
static inline void vlist_push(orc_scratch *s, uint32_t id, int64_t score) {
    if (s->vlen == s->vcap) { s->vcap *= 2; s->vlist = (cand_t *)realloc(s->vlist, sizeof(cand_t) * s->vcap); }
    s->vlist[s->vlen].id = id; s->vlist[s->vlen].score = score; s->vlen++;
}

/* lib.rs:183-211.  Returns GreedySearchCounters.distances.  Results: scratch nb (candidate list) and vlist. */
ORC_API uint64_t orc_greedy_search(orc_scratch *s, uint32_t start, int base_vectors_only, const uint16_t *query,
                                   const uint16_t *x, size_t d, const orc_graph *g, const orc_build_config *cfg) {
    visit_clear(s);
    orc_nb_clear(s->nb);
    s->vlen = 0;
    orc_nb_insert(s->nb, start, fast_dot(query, x + (size_t)start * d, d));
    visit_insert(s, start);
    uint64_t distances = 0;
    uint32_t pt;
    while (orc_nb_next_unvisited(s->nb, &pt)) {
        size_t npre = 0;
        const uint32_t *nbrs = g->adj + (size_t)pt * g->stride;
        uint32_t dg = g->deg[pt];
        if (dg > s->pre_cap) { s->pre_cap = dg; s->pre = (uint32_t *)realloc(s->pre, sizeof(uint32_t) * dg); }
        for (uint32_t j = 0; j < dg; j++) {
            uint32_t nbh = nbrs[j];
            int is_query = nbh >= cfg->query_breakpoint;
            if (visit_insert(s, nbh) && !(base_vectors_only && is_query)) s->pre[npre++] = nbh;
        }
        for (size_t j = 0; j < npre; j++) {
            int64_t sc = fast_dot(query, x + (size_t)s->pre[j] * d, d);
            distances++;
            orc_nb_insert(s->nb, s->pre[j], sc);
            vlist_push(s, s->pre[j], sc);
        }
    }
    return distances;
}

/* greedy_search over a batch of queries, OpenMP over queries with one Scratch per thread (the way diskann/src/main.rs:117-125
 * drives it through rayon); prefetches the next neighbour's row like lib.rs:202-203.  Used by bench.py's cpu_baseline leg. */
ORC_API void orc_greedy_search_batch(const uint16_t *x, size_t n, size_t d, const orc_graph *g, const orc_build_config *cfg, uint32_t start,
                                     const uint16_t *queries, size_t nq, uint32_t *out_ids, int64_t *out_scores, uint32_t *out_len,
                                     uint64_t *out_distances) {
    size_t L = cfg->l;
#pragma omp parallel
    {
        orc_scratch *s = orc_scratch_new(n, cfg->l, cfg->r);
#pragma omp for schedule(dynamic, 4)
        for (long qi = 0; qi < (long)nq; qi++) {
            const uint16_t *query = queries + (size_t)qi * d;
            visit_clear(s);
            orc_nb_clear(s->nb);
            s->vlen = 0;
            orc_nb_insert(s->nb, start, fast_dot(query, x + (size_t)start * d, d));
            visit_insert(s, start);
            uint64_t distances = 0;
            uint32_t pt;
            while (orc_nb_next_unvisited(s->nb, &pt)) {
                size_t npre = 0;
                const uint32_t *nbrs = g->adj + (size_t)pt * g->stride;
                uint32_t dg = g->deg[pt];
                if (dg > s->pre_cap) { s->pre_cap = dg; s->pre = (uint32_t *)realloc(s->pre, sizeof(uint32_t) * dg); }
                for (uint32_t j = 0; j < dg; j++)
                    if (visit_insert(s, nbrs[j])) s->pre[npre++] = nbrs[j];
                for (size_t j = 0; j < npre; j++) {
                    if (j + 1 < npre) {
                        const char *nx = (const char *)(x + (size_t)s->pre[j + 1] * d);
                        for (size_t b = 0; b < d * 2; b += 64) __builtin_prefetch(nx + b, 0, 3);
                    }
                    int64_t sc = fast_dot(query, x + (size_t)s->pre[j] * d, d);
                    distances++;
                    orc_nb_insert(s->nb, s->pre[j], sc);
                }
            }
            size_t len = orc_nb_len(s->nb);
            for (size_t i = 0; i < L; i++) {
                out_ids[(size_t)qi * L + i] = i < len ? s->nb->ids[i] : 0xFFFFFFFFu;
                out_scores[(size_t)qi * L + i] = i < len ? s->nb->scores[i] : 0;
            }
            out_len[qi] = (uint32_t)len;
            out_distances[qi] = distances;
        }
        orc_scratch_free(s);
    }
}

/* ------------------------------------------------------------------ robust_prune (lib.rs:215-285) */

static void merge_existing(orc_scratch *s, uint32_t point, const uint32_t *neigh, size_t n_neigh, const uint16_t *x, size_t d) {
    const uint16_t *pv = x + (size_t)point * d;
    for (size_t i = 0; i < n_neigh; i++) vlist_push(s, neigh[i], fast_dot(pv, x + (size_t)neigh[i] * d, d));
}

/* score descending; equal scores keep arrival order (the reference's sort_unstable_by_key leaves
 * the order of equal keys unspecified -- this restatement fixes it to "stable") */
typedef struct { cand_t c; uint32_t ord; } scand_t;
static int scand_cmp(const void *a, const void *b) {
    const scand_t *x = (const scand_t *)a, *y = (const scand_t *)b;
    if (x->c.score != y->c.score) return x->c.score > y->c.score ? -1 : 1;
    return x->ord < y->ord ? -1 : (x->ord > y->ord ? 1 : 0);
}
static void sort_candidates(cand_t *c, size_t n) {
    scand_t *t = (scand_t *)malloc(sizeof(scand_t) * (n + 1));
    for (size_t i = 0; i < n; i++) { t[i].c = c[i]; t[i].ord = (uint32_t)i; }
    qsort(t, n, sizeof(scand_t), scand_cmp);
    for (size_t i = 0; i < n; i++) c[i] = t[i].c;
    free(t);
}

/* candidates = s->vlist (consumed).  Writes the new neighbour list to out (<= r entries), returns its length. */
static size_t robust_prune(orc_scratch *s, uint32_t p, uint32_t *out, const uint16_t *x, size_t d, const orc_build_config *cfg) {
    cand_t *c = s->vlist;
    sort_candidates(c, s->vlen);
    if (s->vlen > cfg->maxc) s->vlen = cfg->maxc;
    size_t nc = s->vlen, nout = 0, ci = 0;
    while (nout < cfg->r && ci < nc) {
        uint32_t p_star = c[ci].id;
        int64_t p_star_score = c[ci].score;
        ci++;
        if (p_star == p || p_star_score == INT64_MIN) continue;
        out[nout++] = p_star;
        const uint16_t *sv = x + (size_t)p_star * d;
        /* :250 -- starts one PAST the element after p_star */
        for (size_t i = ci + 1; i < nc; i++) {
            if (c[i].score == INT64_MIN) continue;
            uint32_t p_prime = c[i].id;
            int64_t star_prime = fast_dot(x + (size_t)p_prime * d, sv, d);
            int64_t a = p_prime >= cfg->query_breakpoint ? cfg->query_alpha : cfg->alpha;
            /* wrapping multiply then arithmetic shift, as i64 `*` and `>>` do in a release build */
            int64_t scaled = (int64_t)((uint64_t)a * (uint64_t)star_prime) >> 16;
            if (scaled >= c[i].score) c[i].score = INT64_MIN;
        }
    }
    if (cfg->saturate_graph || p >= cfg->query_breakpoint) {
        for (size_t i = 0; i < nc && nout < cfg->r; i++) {
            int present = 0;
            for (size_t j = 0; j < nout; j++) if (out[j] == c[i].id) { present = 1; break; }
            if (!present) out[nout++] = c[i].id;
        }
    }
    return nout;
}

/* standalone entry: prune an explicit candidate list for point p */
ORC_API size_t orc_robust_prune(uint32_t p, const uint32_t *cand_ids, const int64_t *cand_scores, size_t n_cand,
                                const uint16_t *x, size_t n, size_t d, const orc_build_config *cfg, uint32_t *out) {
    orc_scratch *s = orc_scratch_new(n, cfg->l, cfg->r);
    s->vlen = 0;
    for (size_t i = 0; i < n_cand; i++) vlist_push(s, cand_ids[i], cand_scores[i]);
    size_t r = robust_prune(s, p, out, x, d, cfg);
    orc_scratch_free(s);
    return r;
}

/* ------------------------------------------------------------------ build_graph (lib.rs:287-324) */

static void shuffle_u32(uint32_t *a, size_t n, uint64_t *seed) {
    for (size_t i = n; i > 1; i--) {
        size_t j = (size_t)(splitmix64(seed) % i);
        uint32_t t = a[i - 1]; a[i - 1] = a[j]; a[j] = t;
    }
}

static void build_one(orc_scratch *s, uint32_t sigma, uint32_t medioid, orc_graph *g, const uint16_t *x, size_t d,
                      const orc_build_config *cfg, uint32_t *tmp
#ifdef _OPENMP
                      , omp_lock_t *locks
#endif
) {
    int is_query = sigma >= cfg->query_breakpoint;
    /* the reference takes a read lock per expanded node; a racing writer can only ever hand the
     * search a slightly newer list, so the parallel variant here reads without the lock */
    orc_greedy_search(s, medioid, is_query, x + (size_t)sigma * d, x, d, g, cfg);
    uint32_t *my = g->adj + (size_t)sigma * g->stride;
#ifdef _OPENMP
    if (locks) omp_set_lock(&locks[sigma]);
#endif
    merge_existing(s, sigma, my, g->deg[sigma], x, d);
    size_t nn = robust_prune(s, sigma, tmp, x, d, cfg);
    memcpy(my, tmp, nn * sizeof(uint32_t));
    g->deg[sigma] = (uint32_t)nn;
#ifdef _OPENMP
    if (locks) omp_unset_lock(&locks[sigma]);
#endif
    /* tmp[0..nn) is the owned copy of the new neighbour list (`to_owned()` at :310) */
    uint32_t *tmp2 = tmp + cfg->r + 1;
    for (size_t k = 0; k < nn; k++) {
        uint32_t nb = tmp[k];
#ifdef _OPENMP
        if (locks) omp_set_lock(&locks[nb]);
#endif
        uint32_t *nbl = g->adj + (size_t)nb * g->stride;
        uint32_t nd = g->deg[nb];
        if (nd == cfg->r) {
            s->vlen = 0;
            merge_existing(s, nb, nbl, nd, x, d);
            merge_existing(s, nb, &sigma, 1, x, d);
            size_t m = robust_prune(s, nb, tmp2, x, d, cfg);
            memcpy(nbl, tmp2, m * sizeof(uint32_t));
            g->deg[nb] = (uint32_t)m;
        } else {
            int present = 0;
            for (uint32_t j = 0; j < nd; j++) if (nbl[j] == sigma) { present = 1; break; }
            if (!present && nd < cfg->r) { nbl[nd] = sigma; g->deg[nb] = nd + 1; }
        }
#ifdef _OPENMP
        if (locks) omp_unset_lock(&locks[nb]);
#endif
    }
}

/* parallel != 0: OpenMP over points with one lock per node, like the reference's rayon + RwLock build
 * (order-dependent, only statistically reproducible).  parallel == 0: sequential, deterministic. */
ORC_API void orc_build_graph(orc_graph *g, uint32_t medioid, const uint16_t *x, size_t d, const orc_build_config *cfg,
                             uint64_t seed, int parallel) {
    size_t n = g->n;
    uint32_t *sigmas = (uint32_t *)malloc(sizeof(uint32_t) * (n + 1));
    for (size_t i = 0; i < n; i++) sigmas[i] = (uint32_t)i;
    shuffle_u32(sigmas, n, &seed);
#ifdef _OPENMP
    if (parallel) {
        omp_lock_t *locks = (omp_lock_t *)malloc(sizeof(omp_lock_t) * n);
        for (size_t i = 0; i < n; i++) omp_init_lock(&locks[i]);
#pragma omp parallel
        {
            orc_scratch *s = orc_scratch_new(n, cfg->l, cfg->r);
            uint32_t *tmp = (uint32_t *)malloc(sizeof(uint32_t) * 2 * (cfg->r + 2));
#pragma omp for schedule(dynamic, 16)
            for (size_t i = 0; i < n; i++) build_one(s, sigmas[i], medioid, g, x, d, cfg, tmp, locks);
            free(tmp);
            orc_scratch_free(s);
        }
        for (size_t i = 0; i < n; i++) omp_destroy_lock(&locks[i]);
        free(locks);
        free(sigmas);
        return;
    }
#else
    (void)parallel;
#endif
    orc_scratch *s = orc_scratch_new(n, cfg->l, cfg->r);
    uint32_t *tmp = (uint32_t *)malloc(sizeof(uint32_t) * 2 * (cfg->r + 2));
    for (size_t i = 0; i < n; i++) build_one(s, sigmas[i], medioid, g, x, d, cfg, tmp
#ifdef _OPENMP
        , NULL
#endif
    );
    free(tmp);
    orc_scratch_free(s);
    free(sigmas);
}

/* ------------------------------------------------------------------ robust_stitch (lib.rs:326-374), sequential */

static int cand_desc_cmp(const void *a, const void *b) {
    const scand_t *x = (const scand_t *)a, *y = (const scand_t *)b;
    if (x->c.score != y->c.score) return x->c.score > y->c.score ? -1 : 1;
    return x->ord < y->ord ? -1 : (x->ord > y->ord ? 1 : 0);
}
static void robust_stitch_impl(orc_graph *g, const uint16_t *x, size_t d, const orc_build_config *cfg, uint64_t seed, const uint32_t *given_order) {
    uint32_t qb = cfg->query_breakpoint;
    size_t n = g->n;
    if (qb >= n) return;
    size_t nq = n - qb;
    /* in-edges to each query node from base nodes, dropping those edges from the base nodes */
    uint32_t **in = (uint32_t **)calloc(nq, sizeof(uint32_t *));
    size_t *in_n = (size_t *)calloc(nq, sizeof(size_t)), *in_c = (size_t *)calloc(nq, sizeof(size_t));
    for (uint32_t b = 0; b < qb; b++) {
        uint32_t *nb = g->adj + (size_t)b * g->stride;
        uint32_t w = 0;
        for (uint32_t j = 0; j < g->deg[b]; j++) {
            uint32_t t = nb[j];
            if (t >= qb) {
                size_t qi = t - qb;
                if (in_n[qi] == in_c[qi]) { in_c[qi] = in_c[qi] ? in_c[qi] * 2 : 8; in[qi] = (uint32_t *)realloc(in[qi], in_c[qi] * sizeof(uint32_t)); }
                in[qi][in_n[qi]++] = b;
            } else nb[w++] = t;
        }
        g->deg[b] = w;
    }
    uint32_t *order = (uint32_t *)malloc(sizeof(uint32_t) * nq);
    if (given_order) memcpy(order, given_order, sizeof(uint32_t) * nq);
    else {
        for (size_t i = 0; i < nq; i++) order[i] = qb + (uint32_t)i;
        shuffle_u32(order, nq, &seed);
    }
    scand_t *cs = (scand_t *)malloc(sizeof(scand_t) * (g->stride + 1));
    for (size_t oi = 0; oi < nq; oi++) {
        uint32_t q = order[oi];
        const uint32_t *qn = g->adj + (size_t)q * g->stride;
        uint32_t qd = g->deg[q];
        for (size_t e = 0; e < in_n[q - qb]; e++) {
            uint32_t inb = in[q - qb][e];
            for (uint32_t j = 0; j < qd; j++) {
                cs[j].c.id = qn[j];
                cs[j].c.score = fast_dot(x + (size_t)inb * d, x + (size_t)qn[j] * d, d);
                cs[j].ord = j;
            }
            qsort(cs, qd, sizeof(scand_t), cand_desc_cmp);
            uint32_t *out = g->adj + (size_t)inb * g->stride;
            size_t added = 0;
            for (uint32_t j = 0; j < qd; j++) {
                if (added >= cfg->max_add_per_stitch_iter || g->deg[inb] >= cfg->r) break;
                int present = 0;
                for (uint32_t t = 0; t < g->deg[inb]; t++) if (out[t] == cs[j].c.id) { present = 1; break; }
                if (present) continue;
                out[g->deg[inb]++] = cs[j].c.id;
                added++;
            }
        }
    }
    free(cs); free(order);
    for (size_t i = 0; i < nq; i++) free(in[i]);
    free(in); free(in_n); free(in_c);
}

ORC_API void orc_robust_stitch(orc_graph *g, const uint16_t *x, size_t d, const orc_build_config *cfg, uint64_t seed) {
    robust_stitch_impl(g, x, d, cfg, seed, NULL);
}
/* the same with the shuffled visiting order of the query nodes (lib.rs:333-334) given by the caller */
ORC_API void orc_robust_stitch_order(orc_graph *g, const uint16_t *x, size_t d, const orc_build_config *cfg, const uint32_t *order) {
    robust_stitch_impl(g, x, d, cfg, 0, order);
}

/* ------------------------------------------------------------------ ProductQuantizer (vector.rs:308-406) */

typedef struct {
    const float *centroids; /* [n_centroids][n_dims] */
    const float *transform; /* [n_dims][n_dims] row-major; y = T x */
    size_t n_dims_per_code, n_dims, n_centroids;
} orc_pq;

/* vector.rs:320-329: y = T.x (matrixmultiply's summation order is not pinned; sequential fmaf here) */
ORC_API void orc_pq_apply_transform(const orc_pq *pq, const float *x, size_t n_vec, float *y) {
    size_t D = pq->n_dims;
#pragma omp parallel for schedule(static) if (n_vec > 4)
    for (size_t v = 0; v < n_vec; v++)
        for (size_t i = 0; i < D; i++) {
            const float *t = pq->transform + i * D;
            float acc = 0.0f;
            for (size_t k = 0; k < D; k++) acc = fmaf(t[k], x[v * D + k], acc);
            y[v * D + i] = acc;
        }
}

/* vector.rs:331-364: per subspace, argmax inner product, first maximum wins; a NaN/-inf row keeps code 0 */
ORC_API void orc_pq_quantize_batch(const orc_pq *pq, const float *x, size_t n_vec, uint8_t *codes) {
    size_t D = pq->n_dims, S = pq->n_dims_per_code, M = D / S, C = pq->n_centroids;
    float *y = (float *)malloc(sizeof(float) * n_vec * D);
    orc_pq_apply_transform(pq, x, n_vec, y);
    memset(codes, 0, n_vec * M);
#pragma omp parallel for schedule(static) if (n_vec > 4)
    for (size_t v = 0; v < n_vec; v++)
        for (size_t m = 0; m < M; m++) {
            float best = -INFINITY;
            for (size_t c = 0; c < C; c++) {
                const float *cen = pq->centroids + c * D + m * S;
                const float *yy = y + v * D + m * S;
                float acc = 0.0f;
                for (size_t k = 0; k < S; k++) acc = fmaf(yy[k], cen[k], acc);
                if (acc > best) { best = acc; codes[v * M + m] = (uint8_t)c; }
            }
        }
    free(y);
}

/* vector.rs:367-384: LUT[m*C + c] = <(T q)[mS..mS+S], centroid_c[mS..mS+S]> */
ORC_API void orc_pq_preprocess_query(const orc_pq *pq, const float *q, float *lut) {
    size_t D = pq->n_dims, S = pq->n_dims_per_code, M = D / S, C = pq->n_centroids;
    float *y = (float *)malloc(sizeof(float) * D);
    orc_pq_apply_transform(pq, q, 1, y);
    for (size_t m = 0; m < M; m++)
        for (size_t c = 0; c < C; c++) {
            const float *cen = pq->centroids + c * D + m * S;
            float acc = 0.0f;
            for (size_t k = 0; k < S; k++) acc = fmaf(y[m * S + k], cen[k], acc);
            lut[m * C + c] = acc;
        }
    free(y);
}

/* vector.rs:387-405: f32 accumulation over chunks in chunk order, then scale to i64 */
ORC_API void orc_pq_adc(const float *lut, size_t n_chunks, size_t n_centroids, const uint8_t *codes, size_t n_vec, int64_t *out) {
    for (size_t j = 0; j < n_vec; j++) {
        float s = 0.0f;
        for (size_t i = 0; i < n_chunks; i++) s += lut[i * n_centroids + codes[j * n_chunks + i]];
        out[j] = sat_trunc_f32(s * 4294967296.0f);
    }
}

/* ------------------------------------------------------------------ beam search over the packed index
 * (query_disk_index.rs:83-97,135-212) with the node records held in memory instead of read via io_uring */

typedef struct {
    const uint16_t *vectors;    /* [n][d] fp16 */
    const uint32_t *adj;        /* CSR or fixed stride: node i -> adj[offsets[i] .. offsets[i+1]) */
    const uint64_t *offsets;    /* n+1 */
    const uint8_t *pq_codes;    /* [n][code_size] */
    const uint8_t *descriptors; /* [n][n_desc] or NULL */
    const uint8_t *has_url;     /* [n] or NULL (= all 1): query_disk_index.rs:172 `node.url.len() > 0` */
    size_t n, d, code_size, n_desc, n_centroids;
} orc_disk_index;

static int64_t descriptor_product(const orc_disk_index *ix, const float *scales, uint32_t id) {
    int64_t r = 0;
    for (size_t j = 0; j < ix->n_desc; j++)
        r += sat_trunc_f32((scales[j] * (float)ix->descriptors[(size_t)id * ix->n_desc + j]) * 4294967296.0f);
    return r;
}

/*
 * Returns the number of expanded-and-recorded nodes written to out_ids/out_scores (in visit order, NOT sorted;
 * the caller sorts by score as :529 does).  counts[0] = cmps, counts[1] = pq_cmps.
 * faithful_prebuffer != 0 reproduces :157 (the pre-buffer is cleared once per beam iteration, so later nodes
 * of a beam re-score earlier nodes' neighbours); 0 clears it per expanded node.
 */
/* RabitQ signed sum  sum_i (+-)(P q)_i  straight from a 512-bit sign code (bit i of byte b = output 8b + i), in the integer form the
 * GPU traversal uses (csrc/graph.cu::rq_quantize_query / rq_signed_sum, after the RabitQ paper's query quantisation, arXiv
 * 2405.12497 section 3.3): P q is quantised once per query to 8-bit unsigned integers qu_i = rint((Pq_i - vmin) / delta),
 * delta = (vmax - vmin) / 255; then the sum is  vmin * (2 popc(code) - 512) + delta * (2 S1 - sum_i qu_i),  S1 = sum over set bits
 * of qu_i.  S1 is an exact integer, so no summation order is involved; the three f32 operations below are the definition. */
typedef struct {
    float vmin, delta;
    int qsum;
    uint8_t qu[512];
} rq_query;

static void rabitq_quantize_query(const float *qt, rq_query *o) {
    float lo = qt[0], hi = qt[0];
    for (int i = 1; i < 512; i++) {
        lo = fminf(lo, qt[i]);
        hi = fmaxf(hi, qt[i]);
    }
    const float range = hi - lo;
    const float inv = range > 0.0f ? 255.0f / range : 0.0f;
    o->vmin = lo;
    o->delta = range / 255.0f;
    o->qsum = 0;
    for (int i = 0; i < 512; i++) {
        const float scaled = (qt[i] - lo) * inv;
        long u = lrintf(scaled); /* round to nearest even, as cvt.rni */
        if (u < 0) u = 0;
        if (u > 255) u = 255;
        o->qu[i] = (uint8_t)u;
        o->qsum += (int)u;
    }
}

static float rabitq_int_sum(const rq_query *q, const uint8_t *code) {
    int s1 = 0, pc = 0;
    for (int i = 0; i < 512; i++)
        if ((code[i >> 3] >> (i & 7)) & 1) {
            s1 += q->qu[i];
            pc++;
        }
    const int isum = 2 * s1 - q->qsum, csum = 2 * pc - 512;
    const float vc = q->vmin * (float)csum;
    return fmaf(q->delta, (float)isum, vc);
}

static size_t beam_search_impl(const orc_disk_index *ix, uint32_t start, const uint16_t *query, const float *lut,
                               const float *desc_scales, size_t L, size_t beamwidth, int disable_pq, int faithful_prebuffer,
                               uint32_t *out_ids, int64_t *out_scores, size_t out_cap, uint64_t *counts,
                               const float *code_scale, float code_bias, const float *rq_qt, float rq_scale) {
    size_t n = ix->n, d = ix->d;
    uint8_t *vadj = (uint8_t *)calloc(n + 1, 1), *vis = (uint8_t *)calloc(n + 1, 1);
    orc_nb *nb = orc_nb_new(L);
    size_t pre_cap = 1024, npre = 0;
    uint32_t *pre = (uint32_t *)malloc(sizeof(uint32_t) * pre_cap);
    uint32_t *pts = (uint32_t *)malloc(sizeof(uint32_t) * (beamwidth + 1));
    int64_t *approx = (int64_t *)malloc(sizeof(int64_t) * pre_cap);
    uint64_t cmps = 0, pq_cmps = 0;
    size_t nout = 0;
    size_t M = ix->code_size;
    rq_query rqq;
    if (rq_qt) rabitq_quantize_query(rq_qt, &rqq);

    orc_nb_insert(nb, start, 0); /* :153 seeds with score 0 */
    vadj[start] = 1;
    for (;;) {
        size_t np = 0;
        uint32_t p;
        while (np < beamwidth && orc_nb_next_unvisited(nb, &p)) pts[np++] = p;
        if (np == 0) break;
        npre = 0;
        for (size_t b = 0; b < np; b++) {
            uint32_t id = pts[b];
            if (!faithful_prebuffer) npre = 0;
            int64_t score = fast_dot(query, ix->vectors + (size_t)id * d, d);
            if (ix->n_desc) score += descriptor_product(ix, desc_scales, id);
            cmps++;
            if (!vis[id]) {
                vis[id] = 1;
                if (!ix->has_url || ix->has_url[id]) {
                    if (nout < out_cap) { out_ids[nout] = id; out_scores[nout] = score; }
                    nout++;
                }
            }
            for (uint64_t e = ix->offsets[id]; e < ix->offsets[id + 1]; e++) {
                uint32_t t = ix->adj[e];
                if (!vadj[t]) {
                    vadj[t] = 1;
                    if (npre == pre_cap) {
                        pre_cap *= 2;
                        pre = (uint32_t *)realloc(pre, sizeof(uint32_t) * pre_cap);
                        approx = (int64_t *)realloc(approx, sizeof(int64_t) * pre_cap);
                    }
                    pre[npre++] = t;
                }
            }
            for (size_t i = 0; i < npre; i++) {
                uint32_t t = pre[i];
                int64_t sc;
                if (disable_pq) {
                    sc = fast_dot(query, ix->vectors + (size_t)t * d, d);
                } else {
                    float s = 0.0f;
                    const uint8_t *code = ix->pq_codes + (size_t)t * M;
                    if (rq_qt) s = rq_scale * rabitq_int_sum(&rqq, code);
                    else for (size_t m = 0; m < M; m++) s += lut[m * ix->n_centroids + code[m]];
                    if (code_scale) s = fmaf(s, code_scale[t], code_bias); /* RabitQ estimate, rabitq.py:47-48 */
                    sc = sat_trunc_f32(s * 4294967296.0f);
                    pq_cmps++;
                }
                if (ix->n_desc) sc += descriptor_product(ix, desc_scales, t);
                orc_nb_insert(nb, t, sc);
            }
        }
    }
    if (counts) { counts[0] = cmps; counts[1] = pq_cmps; }
    free(vadj); free(vis); orc_nb_free(nb); free(pre); free(pts); free(approx);
    return nout;
}

ORC_API size_t orc_beam_search(const orc_disk_index *ix, uint32_t start, const uint16_t *query, const float *lut,
                               const float *desc_scales, size_t L, size_t beamwidth, int disable_pq, int faithful_prebuffer,
                               uint32_t *out_ids, int64_t *out_scores, size_t out_cap, uint64_t *counts) {
    return beam_search_impl(ix, start, query, lut, desc_scales, L, beamwidth, disable_pq, faithful_prebuffer, out_ids, out_scores, out_cap,
                            counts, NULL, 0.0f, NULL, 0.0f);
}

/* candidates ranked by (sum of LUT entries) * code_scale[id] + code_bias: the traversal over RabitQ codes (diskann/rabitq.py:42-48
 * as byte tables); everything else as orc_beam_search */
ORC_API size_t orc_beam_search_scaled(const orc_disk_index *ix, uint32_t start, const uint16_t *query, const float *lut,
                                      const float *code_scale, float code_bias, const float *desc_scales, size_t L, size_t beamwidth,
                                      uint32_t *out_ids, int64_t *out_scores, size_t out_cap, uint64_t *counts) {
    return beam_search_impl(ix, start, query, lut, desc_scales, L, beamwidth, 0, 0, out_ids, out_scores, out_cap, counts, code_scale, code_bias,
                            NULL, 0.0f);
}

/* candidates ranked by the RabitQ estimate computed straight from the sign codes: qtm = (P q)[512] followed by <mean, q>;
 * estimate = fma(rq_scale * signed_sum, code_scale[id], <mean, q>) with the integer signed sum above (the device-pointer entry
 * mse_search_beam_dev) */
ORC_API size_t orc_beam_search_rabitq(const orc_disk_index *ix, uint32_t start, const uint16_t *query, const float *qtm, float rq_scale,
                                      const float *code_scale, const float *desc_scales, size_t L, size_t beamwidth,
                                      uint32_t *out_ids, int64_t *out_scores, size_t out_cap, uint64_t *counts) {
    return beam_search_impl(ix, start, query, NULL, desc_scales, L, beamwidth, 0, 0, out_ids, out_scores, out_cap, counts, code_scale, qtm[512],
                            qtm, rq_scale);
}

/* the traversal's RabitQ estimate for n codes (test hook: lets the CPU suite bound the query-quantised value against the numpy
 * restatement of rabitq.py:42-48) */
ORC_API void orc_rabitq_direct_estimates(const float *qtm, float rq_scale, const uint8_t *codes, size_t n, const float *code_scale, float *out) {
    rq_query rqq;
    rabitq_quantize_query(qtm, &rqq);
    for (size_t i = 0; i < n; i++) out[i] = fmaf(rq_scale * rabitq_int_sum(&rqq, codes + i * 64), code_scale[i], qtm[512]);
}

/* ------------------------------------------------------------------ runtime de-duplication (src/query_disk_index.rs:99,486-529)
 * keep[i] = 1 unless an earlier KEPT visited node has dot product (f32, fast_dot summation order; the reference uses sgemm, whose
 * order is unspecified) above `threshold`.  ids: visited nodes in visit order.  Returns the number kept. */
ORC_API size_t orc_dedup_visited(const uint16_t *x, size_t d, const uint32_t *ids, size_t m, float threshold, uint8_t *keep) {
    size_t nk = 0;
    for (size_t i = 0; i < m; i++) {
        int dup = 0;
        for (size_t j = 0; j < i && !dup; j++)
            if (keep[j] && fast_dot_f32(x + (size_t)ids[i] * d, x + (size_t)ids[j] * d, d) > threshold) dup = 1;
        keep[i] = (uint8_t)!dup;
        nk += !dup;
    }
    return nk;
}
