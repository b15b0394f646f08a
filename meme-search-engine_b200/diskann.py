"""Host-side mirror of the ``diskann`` crate's public names (diskann/src/lib.rs, diskann/src/vector.rs) on the C ABI.

    reference (Rust)                                   here
    VectorList::from_f16s / push / len                 VectorList (an mse_index handle; rows live in HBM)
    IndexGraph::empty + random_fill_graph              random_fill_graph(vecs, r, seed)
    medioid(&vecs)                                     medioid(vecs)
    build_graph(rng, graph, medioid, vecs, config)     build_graph(vecs, medioid, config, seed)
    robust_stitch(rng, graph, vecs, config)            robust_stitch(vecs, config, seed, order)
    greedy_search(scratch, start, ..., query, ...)     greedy_search(vecs, queries, start, config) -> SearchResult (batched)
    ProductQuantizer::{quantize_batch, preprocess_query, asymmetric_dot_product}   ProductQuantizer
    query_disk_index.rs greedy_search (beam, PQ)       beam_search(vecs, ...)
    SCALE, scale_dot_result                            SCALE

Scores are the reference's fixed point: i64 = trunc(f32 * 2^32).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import MseError, check, lib
from .flat import FlatIndex

SCALE = 4294967296.0  # diskann/src/vector.rs:46


class IndexBuildConfig(C.Structure):
    """diskann/src/lib.rs:41-51."""
    _fields_ = [("r", C.c_uint64), ("l", C.c_uint64), ("maxc", C.c_uint64), ("alpha", C.c_int64), ("saturate_graph", C.c_int32),
                ("query_breakpoint", C.c_uint32), ("max_add_per_stitch_iter", C.c_uint64), ("query_alpha", C.c_int64)]

    def __init__(self, r=64, l=192, maxc=750, alpha=65536, saturate_graph=False, query_breakpoint=0xFFFFFFFF,
                 max_add_per_stitch_iter=16, query_alpha=65536):
        super().__init__(r, l, maxc, alpha, int(saturate_graph), query_breakpoint, max_add_per_stitch_iter, query_alpha)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class VectorList(FlatIndex):
    """diskann/src/vector.rs:118-186: flat row-major fp16 store (+ the graph and packed-index arrays attached to it)."""

    @classmethod
    def from_f16s(cls, x16: np.ndarray, device: int = 0, id_base: int = 0) -> "VectorList":
        return cls.from_f16(x16, device, id_base)

    def __len__(self):
        return self.ntotal

    def push(self, vec16: np.ndarray):
        self.add_f16(np.asarray(vec16).reshape(1, -1))

    # -- graph ------------------------------------------------------------------------------
    def set_graph(self, adj: np.ndarray, deg: np.ndarray):
        adj = np.ascontiguousarray(adj, np.uint32)
        deg = np.ascontiguousarray(deg, np.uint32)
        check(lib().mse_index_set_graph(self._h, _p(adj), _p(deg), adj.shape[1]), "mse_index_set_graph")

    def get_graph(self):
        stride = C.c_uint32()
        check(lib().mse_index_get_graph(self._h, None, None, C.byref(stride)), "mse_index_get_graph")
        n = self.ntotal
        adj = np.empty((n, stride.value), np.uint32)
        deg = np.empty(n, np.uint32)
        check(lib().mse_index_get_graph(self._h, _p(adj), _p(deg), None), "mse_index_get_graph")
        return adj, deg

    def set_pq_codes(self, codes: np.ndarray):
        codes = np.ascontiguousarray(codes, np.uint8)
        check(lib().mse_index_set_pq_codes(self._h, _p(codes), codes.shape[1]), "mse_index_set_pq_codes")

    def set_descriptors(self, desc=None, has_url=None):
        desc = np.ascontiguousarray(desc, np.uint8) if desc is not None else None
        hu = np.ascontiguousarray(has_url, np.uint8) if has_url is not None else None
        check(lib().mse_index_set_descriptors(self._h, _p(desc), desc.shape[1] if desc is not None else 0, _p(hu)), "mse_index_set_descriptors")

    def set_code_scales(self, scales: np.ndarray):
        """Per-vector factor of scaled codes; RabitQ (rabitq.py:47-48): norms * dots."""
        scales = np.ascontiguousarray(scales, np.float32)
        assert scales.size == self.ntotal
        check(lib().mse_index_set_code_scales(self._h, _p(scales)), "mse_index_set_code_scales")

    def scores_i64(self, q16: np.ndarray) -> np.ndarray:
        """fast_dot_noprefetch of one query against every row (query_disk_index.rs:262-273)."""
        q16 = np.ascontiguousarray(q16)
        out = np.empty(self.ntotal, np.int64)
        check(lib().mse_scores_i64(self._h, _p(q16), _p(out)), "mse_scores_i64")
        return out


def random_fill_graph(vecs: VectorList, r: int, seed: int = 0):
    check(lib().mse_index_random_fill_graph(vecs._h, r, seed), "mse_index_random_fill_graph")


def medioid(vecs: VectorList) -> int:
    out = C.c_uint32()
    check(lib().mse_index_medioid(vecs._h, C.byref(out)), "mse_index_medioid")
    return int(out.value)


def build_graph(vecs: VectorList, medioid_: int, config: IndexBuildConfig, seed: int = 0, max_batch: int = 0) -> dict:
    stats = (C.c_uint64 * 6)()
    check(lib().mse_index_build_vamana(vecs._h, medioid_, C.byref(config), seed, max_batch, stats), "mse_index_build_vamana")
    return {"batches": int(stats[0]), "searches": int(stats[1]), "backedge_merges": int(stats[2]), "distances": int(stats[3]),
            "truncated_visit_lists": int(stats[4])}


def robust_stitch(vecs: VectorList, config: IndexBuildConfig, seed: int = 0, order=None):
    """diskann/src/lib.rs:326-374 on the graph attached to `vecs`; `order` fixes the shuffled visiting order of the query nodes."""
    o = np.ascontiguousarray(order, np.uint32) if order is not None else None
    check(lib().mse_index_robust_stitch(vecs._h, C.byref(config), _p(o), seed), "mse_index_robust_stitch")


def robust_prune(vecs: VectorList, p: int, cand_ids, cand_scores, config: IndexBuildConfig) -> np.ndarray:
    ci = np.ascontiguousarray(cand_ids, np.uint32)
    cs = np.ascontiguousarray(cand_scores, np.int64)
    out = np.empty(config.r + 1, np.uint32)
    n = C.c_uint32()
    check(lib().mse_robust_prune(vecs._h, p, _p(ci), _p(cs), ci.size, C.byref(config), _p(out), C.byref(n)), "mse_robust_prune")
    return out[: n.value].copy()


@dataclass
class SearchResult:
    ids: np.ndarray        # [nq, L] scratch.neighbour_buffer.ids, best first (0xFFFFFFFF past len)
    scores: np.ndarray     # [nq, L] i64
    len: np.ndarray        # [nq]
    distances: np.ndarray  # [nq] GreedySearchCounters.distances
    visited: list | None = None  # per query (ids, scores) of scratch.visited_list


def greedy_search(vecs: VectorList, queries16: np.ndarray, start, config: IndexBuildConfig, base_vectors_only: bool = False,
                  visited_cap: int = 0) -> SearchResult:
    """diskann/src/lib.rs:183-211, batched.  `start` is one node id or an array of per-query ids."""
    q = np.ascontiguousarray(queries16).reshape(-1, vecs.d)
    nq, L = q.shape[0], int(config.l)
    ids = np.empty((nq, L), np.uint32)
    sc = np.empty((nq, L), np.int64)
    ln = np.empty(nq, np.uint32)
    dist = np.empty(nq, np.uint64)
    starts = None if np.isscalar(start) else np.ascontiguousarray(start, np.uint32)
    vi = vs = vl = None
    if visited_cap:
        vi, vs, vl = np.empty((nq, visited_cap), np.uint32), np.empty((nq, visited_cap), np.int64), np.empty(nq, np.uint32)
    check(lib().mse_search_graph(vecs._h, _p(q), nq, L, _p(starts), int(start) if starts is None else 0, int(base_vectors_only),
                                 config.query_breakpoint, _p(ids), _p(sc), _p(ln), _p(dist), _p(vi), _p(vs), _p(vl), visited_cap), "mse_search_graph")
    visited = None
    if visited_cap:
        visited = [(vi[i, : min(vl[i], visited_cap)].copy(), vs[i, : min(vl[i], visited_cap)].copy()) for i in range(nq)]
    return SearchResult(ids, sc, ln, dist, visited)


GRAPH_MODE_AUTO, GRAPH_MODE_CTA, GRAPH_MODE_WARP = 0, 1, 2


def set_graph_mode(mode: int):
    """Schedule of greedy_search on the GPU (one CTA or one warp per query); results are identical in every mode."""
    check(lib().mse_search_graph_set_mode(int(mode)), "mse_search_graph_set_mode")


def greedy_search_dev(vecs: VectorList, d_queries16: int, nq: int, L: int, start: int, d_ids: int, d_scores: int, d_len: int,
                      d_distances: int, stream: int = 0, d_starts: int = 0, base_vectors_only: bool = False,
                      query_breakpoint: int = 0xFFFFFFFF):
    """greedy_search with device pointers (ints), asynchronous on `stream`; call greedy_search_check afterwards."""
    check(lib().mse_search_graph_dev(vecs._h, d_queries16, nq, L, d_starts or None, start, int(base_vectors_only), query_breakpoint,
                                     d_ids, d_scores, d_len, d_distances, stream or None), "mse_search_graph_dev")


def greedy_search_check(vecs: VectorList, nq: int):
    check(lib().mse_search_graph_check(vecs._h, nq), "mse_search_graph_check")


class ProductQuantizer:
    """diskann/src/vector.rs:308-406."""

    def __init__(self, centroids=None, transform=None, n_dims_per_code: int = 18, device: int = 0, _handle=None):
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = _handle
        else:
            cen = np.ascontiguousarray(centroids, np.float32)
            tr = np.ascontiguousarray(transform, np.float32)
            check(lib().mse_pq_create(_p(cen), _p(tr), tr.shape[0], n_dims_per_code, cen.shape[0], device, C.byref(self._h)), "mse_pq_create")
        info = (C.c_uint32 * 4)()
        check(lib().mse_pq_info(self._h, info), "mse_pq_info")
        self.n_dims, self.n_dims_per_code, self.n_chunks, self.n_centroids = [int(v) for v in info]

    @classmethod
    def from_msgpack(cls, data: bytes, device: int = 0) -> "ProductQuantizer":
        """rmp_serde::from_slice on opq.msgpack (aopq_train.py:87-93)."""
        h = C.c_void_p()
        buf = np.frombuffer(data, np.uint8)
        check(lib().mse_pq_load(_p(buf), buf.size, device, C.byref(h)), "mse_pq_load")
        return cls(_handle=h)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            try:
                lib().mse_pq_destroy(self._h)
            except Exception:  # interpreter shutdown
                pass
            self._h = C.c_void_p()

    __del__ = close

    def apply_transform(self, x) -> np.ndarray:
        x = np.ascontiguousarray(np.atleast_2d(x), np.float32)
        y = np.empty_like(x)
        check(lib().mse_pq_apply_transform(self._h, _p(x), x.shape[0], _p(y)), "mse_pq_apply_transform")
        return y

    def quantize_batch(self, x) -> np.ndarray:
        x = np.ascontiguousarray(np.atleast_2d(x), np.float32)
        codes = np.empty((x.shape[0], self.n_chunks), np.uint8)
        check(lib().mse_pq_encode(self._h, _p(x), x.shape[0], _p(codes)), "mse_pq_encode")
        return codes

    def preprocess_query(self, q) -> np.ndarray:
        q = np.ascontiguousarray(np.atleast_2d(q), np.float32)
        lut = np.empty((q.shape[0], self.n_chunks, self.n_centroids), np.float32)
        check(lib().mse_pq_preprocess_query(self._h, _p(q), q.shape[0], _p(lut)), "mse_pq_preprocess_query")
        return lut

    def asymmetric_dot_product(self, lut, codes) -> np.ndarray:
        lut = np.ascontiguousarray(lut, np.float32)
        codes = np.ascontiguousarray(np.atleast_2d(codes), np.uint8)
        out = np.empty(codes.shape[0], np.int64)
        check(lib().mse_pq_adc(self._h, _p(lut), _p(codes), codes.shape[0], _p(out)), "mse_pq_adc")
        return out


def beam_search(vecs: VectorList, queries16, luts, start, L: int, beamwidth: int, desc_scales=None, disable_pq=False, n_centroids=256,
                out_cap: int = 4096, code_bias=None):
    """src/query_disk_index.rs:144-212, batched. -> per query (ids, exact scores) of expanded nodes in visit order, cmps, pq_cmps.
    With code_bias (one f32 per query) candidates are ranked by lut_sum * code_scale[id] + code_bias[q] (RabitQ codes:
    RabitQ.preprocess_query + VectorList.set_code_scales)."""
    q = np.ascontiguousarray(queries16).reshape(-1, vecs.d)
    nq = q.shape[0]
    luts = np.ascontiguousarray(luts, np.float32) if luts is not None else None
    ds = np.ascontiguousarray(desc_scales, np.float32) if desc_scales is not None else None
    starts = None if np.isscalar(start) else np.ascontiguousarray(start, np.uint32)
    ids = np.empty((nq, out_cap), np.uint32)
    sc = np.empty((nq, out_cap), np.int64)
    ln = np.empty(nq, np.uint32)
    cmps, pq = np.empty(nq, np.uint64), np.empty(nq, np.uint64)
    if code_bias is not None:
        cb = np.ascontiguousarray(code_bias, np.float32)
        check(lib().mse_search_beam_scaled(vecs._h, _p(q), _p(luts), _p(cb), _p(ds), nq, L, beamwidth, _p(starts),
                                           int(start) if starts is None else 0, n_centroids, _p(ids), _p(sc), _p(ln), out_cap, _p(cmps), _p(pq)),
              "mse_search_beam_scaled")
    else:
        check(lib().mse_search_beam(vecs._h, _p(q), _p(luts), _p(ds), nq, L, beamwidth, _p(starts), int(start) if starts is None else 0,
                                    int(disable_pq), n_centroids, _p(ids), _p(sc), _p(ln), out_cap, _p(cmps), _p(pq)), "mse_search_beam")
    res = [(ids[i, : min(ln[i], out_cap)].copy(), sc[i, : min(ln[i], out_cap)].copy()) for i in range(nq)]
    return res, cmps, pq


def beam_search_dev(vecs: VectorList, d_queries16: int, nq: int, L: int, beamwidth: int, start: int, topk: int, d_top_ids: int,
                    d_top_scores: int, d_top_len: int, d_cmps: int, d_pq_cmps: int, stream: int = 0, d_luts: int = 0, n_centroids: int = 256,
                    d_qtm: int = 0, rabitq: "RabitQ | None" = None, d_desc_scales: int = 0, d_starts: int = 0):
    """Beam search with device pointers (ints), top-k on the device; PQ tables (d_luts) or RabitQ (d_qtm + rabitq codec)."""
    check(lib().mse_search_beam_dev(vecs._h, d_queries16, d_luts or None, d_qtm or None, rabitq.output_dims if rabitq else 0,
                                    rabitq.n_dims if rabitq else 0, d_desc_scales or None, nq, L, beamwidth, d_starts or None, start, n_centroids,
                                    topk, d_top_ids, d_top_scores, d_top_len, d_cmps, d_pq_cmps, stream or None), "mse_search_beam_dev")


DUPLICATES_THRESHOLD = 0.95  # src/query_disk_index.rs:99


def dedup_topk_dev(vecs: VectorList, nq: int, topk: int, d_top_ids: int, d_top_scores: int, d_top_len: int, d_kept: int = 0,
                   threshold: float = DUPLICATES_THRESHOLD, stream: int = 0):
    """query_disk_index.rs:486-529 over the visit lists of the last beam_search_dev call: drop near-duplicates, rank the rest."""
    check(lib().mse_dedup_topk_dev(vecs._h, nq, threshold, topk, d_top_ids, d_top_scores, d_top_len, d_kept or None, stream or None),
          "mse_dedup_topk_dev")


class RabitQ:
    """diskann/rabitq.py:8-48 (codec from rabitq.msgpack: mean + truncated orthogonal transform)."""

    def __init__(self, mean, transform, device: int = 0):
        self._h = C.c_void_p()
        mean = np.ascontiguousarray(mean, np.float32)
        tr = np.ascontiguousarray(transform, np.float32)
        self.n_dims, self.output_dims = tr.shape[1], tr.shape[0]
        check(lib().mse_rabitq_create(_p(mean), _p(tr), self.n_dims, self.output_dims, device, C.byref(self._h)), "mse_rabitq_create")

    @classmethod
    def from_msgpack(cls, data: bytes, device: int = 0) -> "RabitQ":
        import msgpack
        d = msgpack.unpackb(data)
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        self.n_dims, self.output_dims = int(d["n_dims"]), int(d["output_dims"])
        buf = np.frombuffer(data, np.uint8)
        check(lib().mse_rabitq_load(_p(buf), buf.size, device, C.byref(self._h)), "mse_rabitq_load")
        return self

    @classmethod
    def train(cls, vecs: "VectorList", sample_rows: int = 100_000, output_dims: int = 512, seed: int = 0) -> "RabitQ":
        """rabitq.py:11-28 on the GPU over the rows of `vecs`: dataset mean + truncated random orthogonal transform."""
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        check(lib().mse_rabitq_train(vecs._h, sample_rows, output_dims, seed, C.byref(self._h)), "mse_rabitq_train")
        info = (C.c_uint32 * 2)()
        check(lib().mse_rabitq_info(self._h, info), "mse_rabitq_info")
        self.n_dims, self.output_dims = int(info[0]), int(info[1])
        return self

    def export(self):
        """-> (mean f32[n_dims], transform f32[output_dims, n_dims])"""
        mean = np.empty(self.n_dims, np.float32)
        tr = np.empty((self.output_dims, self.n_dims), np.float32)
        check(lib().mse_rabitq_export(self._h, _p(mean), _p(tr)), "mse_rabitq_export")
        return mean, tr

    def to_msgpack(self) -> bytes:
        """rabitq.msgpack as rabitq.py:62-68 writes it"""
        import msgpack
        mean, tr = self.export()
        return msgpack.packb({"mean": mean.tolist(), "transform": tr.flatten().tolist(), "output_dims": int(self.output_dims), "n_dims": int(self.n_dims)})

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            try:
                lib().mse_rabitq_destroy(self._h)
                self._h = C.c_void_p()
            except Exception:  # interpreter shutdown
                pass

    __del__ = close

    def quantize(self, x16: np.ndarray):
        x16 = np.ascontiguousarray(x16).reshape(-1, self.n_dims)
        n = x16.shape[0]
        codes = np.empty((n, self.output_dims // 8), np.uint8)
        norms, dots = np.empty(n, np.float32), np.empty(n, np.float32)
        check(lib().mse_rabitq_encode(self._h, _p(x16), n, _p(codes), _p(norms), _p(dots)), "mse_rabitq_encode")
        return codes, norms, dots

    def preprocess_query(self, q):
        """Query side of approx_dot as byte tables: (luts [nq, output_dims/8 * 256], bias [nq]) for beam_search(code_bias=...)."""
        q = np.ascontiguousarray(q, np.float32).reshape(-1, self.n_dims)
        luts = np.empty((q.shape[0], self.output_dims // 8 * 256), np.float32)
        bias = np.empty(q.shape[0], np.float32)
        check(lib().mse_rabitq_preprocess_query(self._h, _p(q), q.shape[0], _p(luts), _p(bias)), "mse_rabitq_preprocess_query")
        return luts, bias

    def encode_index(self, vecs: "VectorList", estimator: int = 0):
        """Encode the rows of `vecs` where they lie in HBM and attach codes + scales to it (0: norms*dots as rabitq.py:48, 1: norms/dots)."""
        check(lib().mse_index_encode_rabitq(vecs._h, self._h, int(estimator)), "mse_index_encode_rabitq")

    def query_dev(self, d_q_f32: int, nq: int, d_qtm: int, stream: int = 0):
        """(P q, <mean, q>) rows [nq][output_dims + 1] in HBM for beam_search_dev(d_qtm=...)."""
        check(lib().mse_rabitq_query_dev(self._h, d_q_f32, nq, d_qtm, stream or None), "mse_rabitq_query_dev")

    def approx_dot(self, codes, norms, dots, q) -> np.ndarray:
        q = np.ascontiguousarray(q, np.float32)
        codes = np.ascontiguousarray(codes, np.uint8)
        out = np.empty(codes.shape[0], np.float32)
        check(lib().mse_rabitq_estimate(self._h, _p(q), _p(codes), _p(np.ascontiguousarray(norms, np.float32)),
                                        _p(np.ascontiguousarray(dots, np.float32)), codes.shape[0], _p(out)), "mse_rabitq_estimate")
        return out
