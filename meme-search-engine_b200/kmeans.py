"""Shard centroids and shard assignment on the GPU: the host-side mirror of the reference's kmeans.py (same function names and
argument meaning) and of the indexer's shard assignment (src/dump_processor.rs:426-457), on top of the C ABI
(mse_kmeans_assign / mse_kmeans_anneal / mse_shard_assign -> csrc/kmeans.cu).  The rows live in a diskann.VectorList / FlatIndex
handle (fp16 in HBM, as everywhere in this package); there is no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib

SPILL_K = 2          # kmeans.py:72
SHARD_SPILL = 2      # dump_processor.rs:134


def _handle(vectors):
    h = getattr(vectors, "_h", None)
    if h is None:
        raise TypeError("vectors must be a VectorList / FlatIndex holding the rows on the GPU")
    return h


def cluster_sizes(vectors, centroids: np.ndarray, spill: int = SPILL_K, normalize: bool = True, want_assignment: bool = False):
    """kmeans.py:78-91: `cluster_sizes` [spill][k] (and, optionally, every row's top-`spill` centroid indices [n][spill])."""
    c = np.ascontiguousarray(centroids, np.float32)
    k, d = c.shape
    if d != vectors.d:
        raise ValueError(f"centroids have {d} dims, the rows {vectors.d}")
    counts = np.zeros((spill, k), np.uint32)
    assign = np.empty((len(vectors), spill), np.uint32) if want_assignment else None
    check(lib().mse_kmeans_assign(_handle(vectors), c.ctypes.data, k, spill, 1 if normalize else 0, counts.ctypes.data,
                                  assign.ctypes.data if want_assignment else None), "mse_kmeans_assign")
    return (counts, assign) if want_assignment else counts


def fitness(vectors, centroids: np.ndarray, spill: int = SPILL_K):
    """kmeans.py:78-95: (max |cluster size - n / k| over every rank and cluster, the worst cluster of every rank)."""
    counts = cluster_sizes(vectors, centroids, spill, True)
    dist = np.abs(counts.astype(np.float32) - np.float32(len(vectors) / centroids.shape[0]))
    return float(dist.max()), dist.argmax(axis=1)


def simulated_annealing(vectors, k: int, max_iter: int = 100, spill: int = SPILL_K, seed: int = 0):
    """kmeans.py:73-131.  Returns (L2-normalised centroids [k][d], last accepted fitness, iterations run)."""
    out = np.empty((k, vectors.d), np.float32)
    fit, its = C.c_float(), C.c_uint32()
    check(lib().mse_kmeans_anneal(_handle(vectors), k, spill, max_iter, seed, out.ctypes.data, C.byref(fit), C.byref(its)), "mse_kmeans_anneal")
    return out, float(fit.value), int(its.value)


class ShardAssigner:
    """dump_processor.rs:426-457: the running state (records per shard, records seen) and the per-record rule."""

    def __init__(self, centroids: np.ndarray, balance_fudge: float = 0.2, spill: int = SHARD_SPILL):
        self.centroids = np.ascontiguousarray(centroids, np.float32)
        self.balance_fudge, self.spill = float(balance_fudge), int(spill)
        self.shard_counts = np.zeros(self.centroids.shape[0], np.uint64)    # :207 `shards.push((centroid, file, 0, i))`
        self.bal_count = np.ones(1, np.uint64)                              # :426

    def assign(self, vectors) -> np.ndarray:
        """shard indices [n][spill] for the rows of `vectors`, in row order; updates the running counts."""
        out = np.empty((len(vectors), self.spill), np.uint32)
        check(lib().mse_shard_assign(_handle(vectors), self.centroids.ctypes.data, self.centroids.shape[0], self.spill, self.balance_fudge,
                                     self.shard_counts.ctypes.data, self.bal_count.ctypes.data, out.ctypes.data), "mse_shard_assign")
        return out
