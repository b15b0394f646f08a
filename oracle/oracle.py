"""ctypes front-end of oracle/mse_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
PARITY UNPINNED: the reference holds no golden vectors for this path and cannot be built here.

The Python names mirror the reference's (diskann crate: diskann/src/lib.rs, diskann/src/vector.rs).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SCALE = 4294967296.0  # diskann/src/vector.rs:46


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libmse_oracle.so")
    src = os.path.join(_HERE, "mse_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    return so


class BuildConfig(C.Structure):
    """diskann/src/lib.rs:41-51 IndexBuildConfig (alpha fixed-point x 2^16)."""
    _fields_ = [
        ("r", C.c_uint64), ("l", C.c_uint64), ("maxc", C.c_uint64), ("alpha", C.c_int64),
        ("saturate_graph", C.c_int32), ("query_breakpoint", C.c_uint32),
        ("max_add_per_stitch_iter", C.c_uint64), ("query_alpha", C.c_int64),
    ]


def make_config(r=64, l=192, maxc=750, alpha=65536, saturate_graph=False, query_breakpoint=0xFFFFFFFF,
                max_add_per_stitch_iter=16, query_alpha=65536) -> BuildConfig:
    return BuildConfig(r, l, maxc, alpha, int(saturate_graph), query_breakpoint, max_add_per_stitch_iter, query_alpha)


class _PQ(C.Structure):
    _fields_ = [("centroids", C.c_void_p), ("transform", C.c_void_p), ("n_dims_per_code", C.c_size_t),
                ("n_dims", C.c_size_t), ("n_centroids", C.c_size_t)]


class _DiskIndex(C.Structure):
    _fields_ = [("vectors", C.c_void_p), ("adj", C.c_void_p), ("offsets", C.c_void_p), ("pq_codes", C.c_void_p),
                ("descriptors", C.c_void_p), ("has_url", C.c_void_p), ("n", C.c_size_t), ("d", C.c_size_t),
                ("code_size", C.c_size_t), ("n_desc", C.c_size_t), ("n_centroids", C.c_size_t)]


def lib(scalar: bool = False):
    global _LIB
    build()
    if scalar:
        return _bind(C.CDLL(os.path.join(_HERE, "libmse_oracle_scalar.so")))
    if _LIB is None:
        _LIB = _bind(C.CDLL(os.path.join(_HERE, "libmse_oracle.so")))
    return _LIB


def _bind(l):
    vp, sz, u32, i64, u64 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int64, C.c_uint64
    sig = {
        "orc_h2f": (C.c_float, [C.c_uint16]), "orc_f2h": (C.c_uint16, [C.c_float]),
        "orc_scale_dot_result": (i64, [C.c_float]), "orc_scale_dot_result_f64": (i64, [C.c_double]),
        "orc_have_avx2": (C.c_int, []),
        "orc_fast_dot": (i64, [vp, vp, sz]), "orc_fast_dot_scalar": (i64, [vp, vp, sz]),
        "orc_fast_dot_f32_scalar": (C.c_float, [vp, vp, sz]),
        "orc_fast_dot_batch": (None, [vp, vp, sz, sz, vp]),
        "orc_dot": (i64, [vp, vp, sz]),
        "orc_flat_search": (None, [vp, sz, vp, sz, sz, sz, vp, vp, C.c_int]),
        "orc_flat_scores_f64": (None, [vp, vp, sz, vp, sz, vp]),
        "orc_brute_force_i64": (None, [vp, vp, sz, sz, sz, vp, vp]),
        "orc_nb_new": (vp, [sz]), "orc_nb_free": (None, [vp]), "orc_nb_clear": (None, [vp]),
        "orc_nb_len": (sz, [vp]), "orc_nb_cap": (sz, [vp]), "orc_nb_ids": (vp, [vp]), "orc_nb_scores": (vp, [vp]),
        "orc_nb_next_unvisited": (C.c_int, [vp, vp]), "orc_nb_insert": (None, [vp, u32, i64]),
        "orc_graph_new": (vp, [sz, sz]), "orc_graph_free": (None, [vp]), "orc_graph_adj": (vp, [vp]),
        "orc_graph_deg": (vp, [vp]),
        "orc_random_fill_graph": (None, [vp, sz, u64]), "orc_medioid": (u32, [vp, sz, sz]),
        "orc_scratch_new": (vp, [sz, sz, sz]), "orc_scratch_free": (None, [vp]), "orc_scratch_nb": (vp, [vp]),
        "orc_scratch_visited_len": (sz, [vp]), "orc_scratch_visited_copy": (None, [vp, vp, vp]),
        "orc_greedy_search": (u64, [vp, u32, C.c_int, vp, vp, sz, vp, vp]),
        "orc_greedy_search_batch": (None, [vp, sz, sz, vp, vp, u32, vp, sz, vp, vp, vp, vp]),
        "orc_robust_prune": (sz, [u32, vp, vp, sz, vp, sz, sz, vp, vp]),
        "orc_build_graph": (None, [vp, u32, vp, sz, vp, u64, C.c_int]),
        "orc_robust_stitch": (None, [vp, vp, sz, vp, u64]),
        "orc_robust_stitch_order": (None, [vp, vp, sz, vp, vp]),
        "orc_pq_apply_transform": (None, [vp, vp, sz, vp]), "orc_pq_quantize_batch": (None, [vp, vp, sz, vp]),
        "orc_pq_preprocess_query": (None, [vp, vp, vp]), "orc_pq_adc": (None, [vp, sz, sz, vp, sz, vp]),
        "orc_beam_search": (sz, [vp, u32, vp, vp, vp, sz, sz, C.c_int, C.c_int, vp, vp, sz, vp]),
        "orc_beam_search_scaled": (sz, [vp, u32, vp, vp, vp, C.c_float, vp, sz, sz, vp, vp, sz, vp]),
        "orc_dedup_visited": (sz, [vp, sz, vp, sz, C.c_float, vp]),
        "orc_rabitq_direct_estimates": (None, [vp, C.c_float, vp, sz, vp, vp]),
        "orc_beam_search_rabitq": (sz, [vp, u32, vp, vp, C.c_float, vp, vp, sz, sz, vp, vp, sz, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(l, name)
        f.restype, f.argtypes = res, args
    return l


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def as_u16(x: np.ndarray) -> np.ndarray:
    """fp16 array -> its u16 bit pattern view (contiguous)."""
    x = np.ascontiguousarray(x)
    return x.view(np.uint16) if x.dtype == np.float16 else _c(x, np.uint16)


# ---------------------------------------------------------------- distance kernels

def fast_dot(x, y, scalar=False) -> int:
    x, y = as_u16(x), as_u16(y)
    fn = lib().orc_fast_dot_scalar if scalar else lib().orc_fast_dot
    return int(fn(_p(x), _p(y), x.size))


def fast_dot_batch(q, rows) -> np.ndarray:
    q, rows = as_u16(q), as_u16(rows)
    out = np.empty(rows.shape[0], np.int64)
    lib().orc_fast_dot_batch(_p(q), _p(rows), rows.shape[0], rows.shape[1], _p(out))
    return out


def dot(x, y) -> int:
    x, y = as_u16(x), as_u16(y)
    return int(lib().orc_dot(_p(x), _p(y), x.size))


def scale_dot_result(x: float) -> int:
    return int(lib().orc_scale_dot_result(x))


def flat_search(q, x, k, mode=0):
    """FAISS IndexScalarQuantizer(QT_fp16, IP).search restated (src/main.rs:822,900). -> (ids u32 [nq,k], scores f32)."""
    q = _c(np.atleast_2d(q), np.float32)
    x16 = as_u16(x)
    n, d = x16.shape
    ids = np.empty((q.shape[0], k), np.uint32)
    sc = np.empty((q.shape[0], k), np.float32)
    lib().orc_flat_search(_p(q), q.shape[0], _p(x16), n, d, k, _p(ids), _p(sc), mode)
    return ids, sc


def flat_scores_f64(q, x, row_ids) -> np.ndarray:
    q = _c(q, np.float32)
    x16 = as_u16(x)
    rid = _c(row_ids, np.uint32)
    out = np.empty(rid.size, np.float64)
    lib().orc_flat_scores_f64(_p(q), _p(x16), x16.shape[1], _p(rid), rid.size, _p(out))
    return out


def brute_force_i64(q16, x, k):
    q16, x16 = as_u16(q16), as_u16(x)
    n, d = x16.shape
    k = min(k, n)
    ids = np.empty(k, np.uint32)
    sc = np.empty(k, np.int64)
    lib().orc_brute_force_i64(_p(q16), _p(x16), n, d, k, _p(ids), _p(sc))
    return ids, sc


# ---------------------------------------------------------------- NeighbourBuffer

class NeighbourBuffer:
    """diskann/src/lib.rs:73-155."""

    def __init__(self, size: int):
        self._h = lib().orc_nb_new(size)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_nb_free(self._h)
            self._h = None

    def insert(self, id: int, score: int):
        lib().orc_nb_insert(self._h, id, score)

    def next_unvisited(self):
        out = C.c_uint32()
        return int(out.value) if lib().orc_nb_next_unvisited(self._h, C.byref(out)) else None

    def clear(self):
        lib().orc_nb_clear(self._h)

    def __len__(self):
        return int(lib().orc_nb_len(self._h))

    def cap(self):
        return int(lib().orc_nb_cap(self._h))

    @property
    def ids(self) -> np.ndarray:
        n = len(self)
        return np.ctypeslib.as_array(C.cast(lib().orc_nb_ids(self._h), C.POINTER(C.c_uint32)), (n,)).copy() if n else np.empty(0, np.uint32)

    @property
    def scores(self) -> np.ndarray:
        n = len(self)
        return np.ctypeslib.as_array(C.cast(lib().orc_nb_scores(self._h), C.POINTER(C.c_int64)), (n,)).copy() if n else np.empty(0, np.int64)


# ---------------------------------------------------------------- graph

class IndexGraph:
    """diskann/src/lib.rs:16-39 with fixed-stride adjacency storage."""

    def __init__(self, n: int, capacity: int):
        self.n, self.stride = n, capacity
        self._h = lib().orc_graph_new(n, capacity)

    @classmethod
    def empty(cls, n, capacity):
        return cls(n, capacity)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_graph_free(self._h)
            self._h = None

    @property
    def adj(self) -> np.ndarray:  # live view [n, stride]
        return np.ctypeslib.as_array(C.cast(lib().orc_graph_adj(self._h), C.POINTER(C.c_uint32)), (self.n, self.stride))

    @property
    def deg(self) -> np.ndarray:  # live view [n]
        return np.ctypeslib.as_array(C.cast(lib().orc_graph_deg(self._h), C.POINTER(C.c_uint32)), (self.n,))

    def set(self, adj: np.ndarray, deg: np.ndarray):
        self.adj[:, :] = adj
        self.deg[:] = deg

    def out_neighbours(self, i) -> np.ndarray:
        return self.adj[i, : self.deg[i]].copy()

    def to_csr(self):
        deg = self.deg.astype(np.uint64)
        offsets = np.zeros(self.n + 1, np.uint64)
        np.cumsum(deg, out=offsets[1:])
        mask = np.arange(self.stride)[None, :] < self.deg[:, None]
        return self.adj[mask].astype(np.uint32), offsets


def random_fill_graph(graph: IndexGraph, r: int, seed: int = 0):
    lib().orc_random_fill_graph(graph._h, r, seed)


def medioid(x) -> int:
    x16 = as_u16(x)
    return int(lib().orc_medioid(_p(x16), x16.shape[0], x16.shape[1]))


class Scratch:
    """diskann/src/lib.rs:157-175."""

    def __init__(self, n: int, config: BuildConfig):
        self._h = lib().orc_scratch_new(n, config.l, config.r)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_scratch_free(self._h)
            self._h = None

    @property
    def neighbour_ids(self) -> np.ndarray:
        nb = lib().orc_scratch_nb(self._h)
        n = int(lib().orc_nb_len(nb))
        return np.ctypeslib.as_array(C.cast(lib().orc_nb_ids(nb), C.POINTER(C.c_uint32)), (n,)).copy() if n else np.empty(0, np.uint32)

    @property
    def neighbour_scores(self) -> np.ndarray:
        nb = lib().orc_scratch_nb(self._h)
        n = int(lib().orc_nb_len(nb))
        return np.ctypeslib.as_array(C.cast(lib().orc_nb_scores(nb), C.POINTER(C.c_int64)), (n,)).copy() if n else np.empty(0, np.int64)

    def visited_list(self):
        n = int(lib().orc_scratch_visited_len(self._h))
        ids, sc = np.empty(n, np.uint32), np.empty(n, np.int64)
        if n:
            lib().orc_scratch_visited_copy(self._h, _p(ids), _p(sc))
        return ids, sc


def greedy_search(scratch: Scratch, start: int, base_vectors_only: bool, query, x, graph: IndexGraph, config: BuildConfig) -> int:
    """diskann/src/lib.rs:183-211 -> GreedySearchCounters.distances; results in scratch.neighbour_ids."""
    q16, x16 = as_u16(query), as_u16(x)
    return int(lib().orc_greedy_search(scratch._h, start, int(base_vectors_only), _p(q16), _p(x16), x16.shape[1],
                                       graph._h, C.byref(config)))


def greedy_search_batch(start: int, queries, x, graph: IndexGraph, config: BuildConfig):
    """greedy_search for a batch of queries on all host cores (OpenMP over queries, one Scratch per thread).
    -> ids [nq, L] (0xFFFFFFFF past len), scores [nq, L], len [nq], distances [nq]."""
    q16, x16 = as_u16(queries), as_u16(x)
    q16 = q16.reshape(-1, x16.shape[1])
    nq, L = q16.shape[0], int(config.l)
    ids, sc = np.empty((nq, L), np.uint32), np.empty((nq, L), np.int64)
    ln, dist = np.empty(nq, np.uint32), np.empty(nq, np.uint64)
    lib().orc_greedy_search_batch(_p(x16), x16.shape[0], x16.shape[1], graph._h, C.byref(config), start, _p(q16), nq, _p(ids), _p(sc), _p(ln),
                                  _p(dist))
    return ids, sc, ln, dist


def robust_prune(p: int, cand_ids, cand_scores, x, config: BuildConfig) -> np.ndarray:
    x16 = as_u16(x)
    ci, cs = _c(cand_ids, np.uint32), _c(cand_scores, np.int64)
    out = np.empty(config.r + 1, np.uint32)
    n = lib().orc_robust_prune(p, _p(ci), _p(cs), ci.size, _p(x16), x16.shape[0], x16.shape[1], C.byref(config), _p(out))
    return out[:n].copy()


def build_graph(graph: IndexGraph, medioid_: int, x, config: BuildConfig, seed: int = 0, parallel: bool = False):
    x16 = as_u16(x)
    lib().orc_build_graph(graph._h, medioid_, _p(x16), x16.shape[1], C.byref(config), seed, int(parallel))


def robust_stitch(graph: IndexGraph, x, config: BuildConfig, seed: int = 0, order=None):
    """diskann/src/lib.rs:326-374; `order` (query node ids in visiting order) replaces the seeded shuffle of :333-334."""
    x16 = as_u16(x)
    if order is not None:
        o = _c(order, np.uint32)
        assert o.size == graph.n - config.query_breakpoint
        lib().orc_robust_stitch_order(graph._h, _p(x16), x16.shape[1], C.byref(config), _p(o))
    else:
        lib().orc_robust_stitch(graph._h, _p(x16), x16.shape[1], C.byref(config), seed)


# ---------------------------------------------------------------- ProductQuantizer

class ProductQuantizer:
    """diskann/src/vector.rs:308-406.  centroids [C, D] f32, transform [D, D] f32 row-major (y = T x)."""

    def __init__(self, centroids, transform, n_dims_per_code: int):
        self.centroids = _c(centroids, np.float32)
        self.transform = _c(transform, np.float32)
        self.n_dims = self.transform.shape[0]
        self.n_dims_per_code = n_dims_per_code
        self.n_centroids = self.centroids.shape[0]
        self.n_chunks = self.n_dims // n_dims_per_code
        self._s = _PQ(self.centroids.ctypes.data, self.transform.ctypes.data, n_dims_per_code, self.n_dims, self.n_centroids)

    def apply_transform(self, x) -> np.ndarray:
        x = _c(np.atleast_2d(x), np.float32)
        y = np.empty_like(x)
        lib().orc_pq_apply_transform(C.byref(self._s), _p(x), x.shape[0], _p(y))
        return y

    def quantize_batch(self, x) -> np.ndarray:
        x = _c(np.atleast_2d(x), np.float32)
        codes = np.empty((x.shape[0], self.n_chunks), np.uint8)
        lib().orc_pq_quantize_batch(C.byref(self._s), _p(x), x.shape[0], _p(codes))
        return codes

    def preprocess_query(self, q) -> np.ndarray:
        q = _c(q, np.float32)
        lut = np.empty((self.n_chunks, self.n_centroids), np.float32)
        lib().orc_pq_preprocess_query(C.byref(self._s), _p(q), _p(lut))
        return lut

    def asymmetric_dot_product(self, lut, codes) -> np.ndarray:
        lut = _c(lut, np.float32)
        codes = _c(np.atleast_2d(codes), np.uint8)
        out = np.empty(codes.shape[0], np.int64)
        lib().orc_pq_adc(_p(lut), self.n_chunks, self.n_centroids, _p(codes), codes.shape[0], _p(out))
        return out


# ---------------------------------------------------------------- packed-index beam search

def beam_search(vectors, adj, offsets, pq_codes, lut, start, query, L, beamwidth, descriptors=None, desc_scales=None,
                has_url=None, disable_pq=False, faithful_prebuffer=False, n_centroids=256, code_scale=None, code_bias=0.0,
                rabitq_qtm=None, rabitq_scale=0.0):
    """src/query_disk_index.rs:144-212 over in-memory node records.
    -> (ids, scores) of expanded nodes in visit order, (cmps, pq_cmps)."""
    v16 = as_u16(vectors)
    n, d = v16.shape
    adj = _c(adj, np.uint32)
    offsets = _c(offsets, np.uint64)
    codes = _c(pq_codes, np.uint8)
    lut = _c(lut, np.float32) if lut is not None else np.zeros(1, np.float32)
    q16 = as_u16(query)
    desc = _c(descriptors, np.uint8) if descriptors is not None else None
    scales = _c(desc_scales, np.float32) if desc_scales is not None else np.zeros(1, np.float32)
    hu = _c(has_url, np.uint8) if has_url is not None else None
    ix = _DiskIndex(v16.ctypes.data, adj.ctypes.data, offsets.ctypes.data, codes.ctypes.data,
                    desc.ctypes.data if desc is not None else None, hu.ctypes.data if hu is not None else None,
                    n, d, codes.shape[1], desc.shape[1] if desc is not None else 0, n_centroids)
    cap = n
    ids, sc = np.empty(cap, np.uint32), np.empty(cap, np.int64)
    counts = np.zeros(2, np.uint64)
    if rabitq_qtm is not None:
        cs, qtm = _c(code_scale, np.float32), _c(rabitq_qtm, np.float32)
        assert qtm.size == 513 and codes.shape[1] == 64
        m = lib().orc_beam_search_rabitq(C.byref(ix), start, _p(q16), _p(qtm), C.c_float(rabitq_scale), _p(cs), _p(scales), L, beamwidth,
                                         _p(ids), _p(sc), cap, _p(counts))
    elif code_scale is not None:
        cs = _c(code_scale, np.float32)
        m = lib().orc_beam_search_scaled(C.byref(ix), start, _p(q16), _p(lut), _p(cs), C.c_float(code_bias), _p(scales), L, beamwidth,
                                         _p(ids), _p(sc), cap, _p(counts))
    else:
        m = lib().orc_beam_search(C.byref(ix), start, _p(q16), _p(lut), _p(scales), L, beamwidth, int(disable_pq),
                                  int(faithful_prebuffer), _p(ids), _p(sc), cap, _p(counts))
    return ids[:m].copy(), sc[:m].copy(), (int(counts[0]), int(counts[1]))


def rabitq_direct_estimates(qtm, rq_scale, codes, code_scale) -> np.ndarray:
    """f32 RabitQ estimates straight from 64-byte sign codes in the GPU's summation order (mse_search_beam_dev)."""
    qtm, codes, cs = _c(qtm, np.float32), _c(codes, np.uint8), _c(code_scale, np.float32)
    out = np.empty(codes.shape[0], np.float32)
    lib().orc_rabitq_direct_estimates(_p(qtm), C.c_float(rq_scale), _p(codes), codes.shape[0], _p(cs), _p(out))
    return out


def dedup_visited(x, ids, threshold: float = 0.95) -> np.ndarray:
    """src/query_disk_index.rs:486-527: boolean keep mask over the visited nodes (visit order)."""
    x16, ids = as_u16(x), _c(ids, np.uint32)
    keep = np.zeros(ids.size, np.uint8)
    lib().orc_dedup_visited(_p(x16), x16.shape[1], _p(ids), ids.size, C.c_float(threshold), _p(keep))
    return keep.astype(bool)
