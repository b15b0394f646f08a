// Internal structures shared by the translation units of libmse_b200.so.
#pragma once
#include "common.cuh"

namespace mse {

// Flat-search workspace, sized per (nq, kp); lives in the index handle so steady-state searches allocate nothing.
struct FlatWork {
    DevBuf q;         // f32 [nq][d]        (host API only: staged queries)
    DevBuf q16;       // f16 [nq_pad][d]    tensor path A operand (rows >= nq are zero)
    DevBuf qstat;     // f32 [nq][4]        {|q|, |q - f16(q)|, eps, unused}
    DevBuf top;       // u64 [nq][kp]       running top list (rank keys, best first)
    DevBuf ntop;      // u32 [nq]
    DevBuf thr;       // f32 [nq]           score of the kp-th best so far, or -inf
    DevBuf cand;      // u64 [nq][cap]      candidates that passed the threshold in the current chunk
    DevBuf count;     // u32 [nq]
    DevBuf flags;     // u32 [4 + nq]       [0]=overflow seen, [1]=#uncertified, [4+q]=per-query status bits
    DevBuf out_ids;   // u32 [nq][k]        (host API only)
    DevBuf out_sc;    // f32 [nq][k]
    DevBuf sel;       // u32 [nq]           query subset for the exact re-run
    DevBuf tmp_ids;   // u32 [nq][k]        exact re-run results
    DevBuf tmp_sc;    // f32 [nq][k]
    void release() {
        q.release(); q16.release(); qstat.release(); top.release(); ntop.release(); thr.release(); cand.release();
        count.release(); flags.release(); out_ids.release(); out_sc.release(); sel.release(); tmp_ids.release(); tmp_sc.release();
    }
};

// where a flat search leaves its results: ids + scores ([nq][k]), and/or rank keys with global ids ([nq][k] u64, 0 = no entry)
struct FlatOut {
    uint32_t *ids = nullptr;
    float *scores = nullptr;
    uint64_t *keys = nullptr;
};
// the last queued flat search of a handle, kept for flat_search_settle
struct FlatPending {
    const float *d_q = nullptr;
    uint32_t nq = 0, k = 0;
    FlatOut out;
    cudaStream_t st = nullptr;
    bool live = false;       // results still need their status word read
    bool queued = false;
    bool st_valid() const { return queued; }
    FlatPending() {}
    FlatPending(const float *q, uint32_t nq_, uint32_t k_, const FlatOut &o, cudaStream_t s, bool live_) : d_q(q), nq(nq_), k(k_), out(o), st(s), live(live_), queued(true) {}
};

// resize.cu: device buffers of one resize (source image, intermediate image, bounds + fixed-point coefficients of the two passes)
struct ResizeWork {
    DevBuf src, tmp, bh, kh, bv, kv;
    void release() { src.release(); tmp.release(); bh.release(); kh.release(); bv.release(); kv.release(); }
};
// host RGB8 [h][w][3] -> device RGB8 [out_h][out_w][3]; filter 0 = resize_for_embed_sync's rule (common.rs:43-44), 1 Hamming, 2 Lanczos3
int resize_rgb_to_device(ResizeWork &wk, const uint8_t *rgb, uint32_t w, uint32_t h, uint32_t out_w, uint32_t out_h, int filter, uint8_t *d_out,
                         cudaStream_t st);

}  // namespace mse

struct mse_index {
    int device = 0;
    uint32_t d = 0;
    uint32_t id_base = 0;
    uint64_t n = 0;      // rows held
    uint64_t cap = 0;    // rows allocated
    __half *x = nullptr; // [cap][d] fp16 rows in HBM
    float *max_norm = nullptr;  // device scalar: max_i |x_i|_2 (upper bound; feeds the certificate)
    // graph + packed-index side arrays (IndexGraph lib.rs:16-39; index.pq-codes.bin / index.descriptor-codes.bin)
    uint32_t *adj = nullptr, *deg = nullptr;   // fixed stride adjacency [n][graph_stride], degrees [n]
    uint32_t graph_stride = 0;
    uint64_t side_n = 0;                       // row count the side arrays below (adj, deg, codes, scales, descriptors) were sized for
    uint8_t *pq_codes = nullptr;               // [n][code_size]
    uint32_t code_size = 0;
    float *code_scale = nullptr;               // [n] per-vector factor of scaled codes (RabitQ: |o| * <o_bar, o>)
    uint8_t *desc = nullptr, *has_url = nullptr;  // [n][n_desc], [n]
    uint32_t n_desc = 0;
    int flat_mode = 0;
    int profile = 0;                 // time the scoring kernels with CUDA events (bench.py roofline)
    std::vector<cudaEvent_t> prof_ev; // start/stop pairs, reused
    size_t prof_used = 0;
    uint64_t stats[8] = {0};
    mse::FlatWork fw;
    mse::FlatPending pending;
    mse::DevBuf gw_htabs, gw_status, gw_vis_ids, gw_vis_sc, gw_vis_len;
    uint32_t gw_vis_cap = 0;          // entries per query of the visit lists the last mse_search_beam_dev call wrote  // graph search workspace of the device-pointer API (visited-set tables, per-query status)
    cudaStream_t stream = nullptr;  // handle-owned stream for the host-pointer API
    // tensor-map cache for the tensor path (encoded lazily, invalidated on growth)
    bool tmap_valid = false;
    alignas(64) unsigned char tmap_x[128];
};

namespace mse {
// flat.cu: rows were appended -> the per-row side arrays (graph, codes, scales, descriptors) no longer cover the index;
// they are released so that later graph / beam calls fail with MSE_ERR_STATE instead of reading past them
void index_drop_side_arrays(mse_index *ix);
// flat_tc.cu: tensor-core scoring pass over rows [row0, row0+nrows) for queries [0,nq) (q16 padded to 128 rows)
int flat_tc_score_chunk(mse_index *ix, uint32_t nq, uint64_t row0, uint64_t nrows, uint32_t cap, cudaStream_t st);
int flat_tc_supported(const mse_index *ix);
// flat.cu: queue a search without synchronising / read its status word and repair flagged queries (synchronises)
int flat_search_queue(mse_index *ix, const float *d_q, uint32_t nq, uint32_t k, const FlatOut &out, cudaStream_t st);
int flat_search_settle(mse_index *ix, uint32_t *repaired);
int flat_publish_status(mse_index *ix, uint64_t *d_word, cudaStream_t st);
int flat_merge_keys(int device, const uint64_t *d_slots, uint32_t n_shards, size_t slot_stride, uint32_t nq, uint32_t k, uint32_t *d_ids,
                    float *d_scores, cudaStream_t st);
}  // namespace mse
