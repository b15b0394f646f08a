"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's shard k-means fitness (kmeans.py:78-95) and of the indexer's
shard assignment (src/dump_processor.rs:426-457).  Pinned to an execution of the reference's own `simulated_annealing`
(tests/golden/make_kmeans_golden.py -> tests/golden/kmeans_reference.json).  Never imported by the product."""
import numpy as np


def normalize(c: np.ndarray) -> np.ndarray:
    """torch.nn.functional.normalize (kmeans.py:82): rows / max(|row|_2, 1e-12)."""
    c = np.asarray(c, np.float32)
    n = np.sqrt((c.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)
    return c / np.maximum(n, np.float32(1e-12))[:, None]


def cluster_sizes(vectors: np.ndarray, centroids: np.ndarray, spill: int = 2, norm: bool = True):
    """kmeans.py:80-91: top-`spill` centroids of every row by inner product, bincount per rank.  Returns (counts [spill][k], top [n][spill])."""
    c = normalize(centroids) if norm else np.asarray(centroids, np.float32)
    sims = np.asarray(vectors, np.float32).astype(np.float64) @ c.astype(np.float64).T
    top = np.argsort(-sims, axis=1, kind="stable")[:, :spill]               # equal scores: lower index first
    k = c.shape[0]
    return np.stack([np.bincount(top[:, j], minlength=k) for j in range(spill)]).astype(np.int64), top


def fitness(vectors: np.ndarray, centroids: np.ndarray, spill: int = 2):
    """kmeans.py:92-95."""
    counts, _ = cluster_sizes(vectors, centroids, spill)
    dist = np.abs(counts.astype(np.float32) - np.float32(len(vectors) / centroids.shape[0]))
    return float(dist.max()), dist.argmax(axis=1)


def shard_assign(vectors: np.ndarray, centroids: np.ndarray, spill: int = 2, balance_fudge: float = 0.2, shard_counts=None, bal_count: int = 1):
    """dump_processor.rs:438-457, record by record: `shards` is stable-sorted IN PLACE by
    -scale_dot_result_f64(dot - fudge * shard_count / bal_count); the first `spill` take the record."""
    c = np.asarray(centroids, np.float32)
    k = c.shape[0]
    counts = np.zeros(k, np.int64) if shard_counts is None else np.asarray(shard_counts, np.int64).copy()
    order = list(range(k))
    dots = np.asarray(vectors, np.float32).astype(np.float64) @ c.astype(np.float64).T
    out = np.empty((len(vectors), spill), np.int64)
    for i in range(len(vectors)):
        key = {s: -int((float(np.float32(dots[i, s])) - balance_fudge * (counts[s] / bal_count)) * 4294967296.0) for s in order}
        order.sort(key=lambda s: key[s])                                    # list.sort is stable
        for j in range(spill):
            out[i, j] = order[j]
            counts[order[j]] += 1
        bal_count += 1
    return out, counts, bal_count
