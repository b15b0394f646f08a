"""Host-side mirror of the FAISS index the reference's small-scale backend uses.

Reference: src/main.rs:822 (``ScalarQuantizerIndexImpl::new(d, QT_fp16, InnerProduct)``), :858/:892 (``add``),
:900 (``search`` -> ``distances``, ``labels``), :1015/:1053 (``ntotal``); legacy mse.py:72-85.
Same names, argument meaning and padding behaviour (labels past ntotal are -1).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import MseError, check, lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class FlatIndex:
    """IndexScalarQuantizer(QT_fp16, METRIC_INNER_PRODUCT) resident in one B200's HBM."""

    def __init__(self, d: int, device: int = 0, id_base: int = 0):
        self.d = int(d)
        self.device = int(device)
        self._h = C.c_void_p()
        check(lib().mse_index_create(None, 0, self.d, self.device, id_base, C.byref(self._h)), "mse_index_create")

    @classmethod
    def from_f16(cls, x16: np.ndarray, device: int = 0, id_base: int = 0) -> "FlatIndex":
        ix = cls(x16.shape[1], device, id_base)
        ix.add_f16(x16)
        return ix

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            try:
                lib().mse_index_destroy(self._h)
            except Exception:  # interpreter shutdown
                pass
            self._h = C.c_void_p()

    __del__ = close

    # -- faiss Index trait ------------------------------------------------------------------
    def add(self, x: np.ndarray):
        """faiss ``add(&[f32])``: rows are stored as fp16 (round to nearest even), as QT_fp16 does."""
        x = np.ascontiguousarray(x, np.float32).reshape(-1, self.d)
        check(lib().mse_index_add(self._h, _ptr(x), x.shape[0]), "mse_index_add")

    def add_f16(self, x16: np.ndarray):
        x16 = np.ascontiguousarray(x16)
        if x16.dtype != np.float16 and x16.dtype != np.uint16:
            raise MseError("add_f16 expects float16 (or its uint16 bit pattern)")
        x16 = x16.reshape(-1, self.d)
        check(lib().mse_index_add_f16(self._h, _ptr(x16), x16.shape[0]), "mse_index_add_f16")

    def add_f16_dev(self, dev_ptr: int, n: int, stream: int = 0):
        check(lib().mse_index_add_f16_dev(self._h, C.c_void_p(dev_ptr), n, C.c_void_p(stream)), "mse_index_add_f16_dev")

    def reserve(self, rows: int):
        check(lib().mse_index_reserve(self._h, rows), "mse_index_reserve")

    @property
    def ntotal(self) -> int:
        return int(lib().mse_index_ntotal(self._h))

    def search(self, q: np.ndarray, k: int):
        """-> (distances f32 [nq,k], labels i64 [nq,k]); labels are -1 past ntotal, like faiss."""
        q = np.ascontiguousarray(q, np.float32).reshape(-1, self.d)
        nq = q.shape[0]
        ids = np.empty((nq, k), np.uint32)
        sc = np.empty((nq, k), np.float32)
        check(lib().mse_search_flat(self._h, _ptr(q), nq, k, _ptr(ids), _ptr(sc)), "mse_search_flat")
        labels = ids.astype(np.int64)
        labels[ids == 0xFFFFFFFF] = -1
        return sc, labels

    def search_dev(self, q_ptr: int, nq: int, k: int, ids_ptr: int, scores_ptr: int, stream: int = 0):
        """Device-pointer variant (queries and results stay in HBM): queues the search on `stream` and returns; the results are
        final after check()."""
        check(lib().mse_search_flat_dev(self._h, C.c_void_p(q_ptr), nq, k, C.c_void_p(ids_ptr), C.c_void_p(scores_ptr),
                                        C.c_void_p(stream)), "mse_search_flat_dev")

    def check(self) -> int:
        """Synchronises the last search_dev, re-runs the (rare) queries whose cut could not be certified; returns how many."""
        rep = C.c_uint32()
        check(lib().mse_search_flat_check(self._h, C.byref(rep)), "mse_search_flat_check")
        return int(rep.value)

    # -- diagnostics -------------------------------------------------------------------------
    def set_mode(self, mode: int):
        """0 auto, 1 exact fp64 scan only, 2 tensor-core pass + certified fp64 rerank only."""
        check(lib().mse_search_flat_set_mode(self._h, mode), "mse_search_flat_set_mode")

    def profile(self, enable: bool = True):
        check(lib().mse_search_flat_profile(self._h, int(enable)), "mse_search_flat_profile")

    def stats(self) -> dict:
        out = (C.c_uint64 * 8)()
        check(lib().mse_search_flat_stats(self._h, out), "mse_search_flat_stats")
        keys = ["tensor_queries", "exact_queries", "uncertified", "overflows", "launches", "chunks", "scoring_ns", "scoring_launches"]
        return {k: int(out[i]) for i, k in enumerate(keys)}

    @property
    def vectors_dev(self) -> int:
        return int(lib().mse_index_vectors_dev(self._h) or 0)


def merge_topk(device: int, ids_ptr: int, scores_ptr: int, n_shards: int, nq: int, k: int, out_ids_ptr: int,
               out_scores_ptr: int, stream: int = 0):
    """k-way merge of all-gathered per-shard top-k lists (device pointers)."""
    check(lib().mse_merge_topk_dev(device, C.c_void_p(ids_ptr), C.c_void_p(scores_ptr), n_shards, nq, k,
                                   C.c_void_p(out_ids_ptr), C.c_void_p(out_scores_ptr), C.c_void_p(stream)), "mse_merge_topk_dev")
