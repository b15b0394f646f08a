"""Shared synthetic-data builders (SURVEY 8d seeds) and pure-numpy models used by several tests."""
import numpy as np

D = 1152


def unit_rows(seed: int, n: int, d: int = D) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def index_f16(seed: int, n: int, d: int = D) -> np.ndarray:
    return unit_rows(seed, n, d).astype(np.float16)


def clustered_f16(seed: int, n: int, n_clusters: int = 64, sigma: float = 0.3, d: int = D) -> np.ndarray:
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((n_clusters, d)).astype(np.float32)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    a = rng.integers(0, n_clusters, n)
    x = c[a] + sigma * rng.standard_normal((n, d)).astype(np.float32) / np.sqrt(d)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float16)


def np_fast_dot_f32(x16: np.ndarray, y16: np.ndarray) -> np.float32:
    """Independent numpy model of diskann/src/vector.rs:255-306.  fp16 products are exact in f64 and the
    f64 sum product+acc is exact for these magnitudes, so float32(float64 fma) reproduces a fused multiply-add."""
    x = x16.astype(np.float64).reshape(-1, 32)
    y = y16.astype(np.float64).reshape(-1, 32)
    p = np.zeros(32, np.float32)
    for c in range(x.shape[0]):
        p = (x[c] * y[c] + p.astype(np.float64)).astype(np.float32)
    A = (p[0:8] + p[8:16]).astype(np.float32)
    B = (p[16:24] + p[24:32]).astype(np.float32)
    f = np.float32
    e0 = f(f(A[0] + A[1]) + f(A[4] + A[5]))
    e1 = f(f(A[2] + A[3]) + f(A[6] + A[7]))
    e2 = f(f(B[0] + B[1]) + f(B[4] + B[5]))
    e3 = f(f(B[2] + B[3]) + f(B[6] + B[7]))
    return f(f(f(e0 + e1) + e2) + e3)


def np_fast_dot(x16, y16) -> int:
    v = np.float32(np_fast_dot_f32(x16, y16) * np.float32(4294967296.0))
    return int(np.trunc(np.float64(v)))


def np_flat_topk(q: np.ndarray, x16: np.ndarray, k: int):
    """(score desc, id asc) on f32-rounded f64 scores -- SURVEY 8c's definition of the flat oracle."""
    s = (q.astype(np.float64) @ x16.astype(np.float64).T).astype(np.float32)
    ids = np.empty((q.shape[0], k), np.int64)
    sc = np.empty((q.shape[0], k), np.float32)
    n = x16.shape[0]
    for i in range(q.shape[0]):
        order = np.lexsort((np.arange(n), -s[i].astype(np.float64)))[:k]
        m = len(order)
        ids[i, :m] = order
        sc[i, :m] = s[i][order]
        ids[i, m:] = -1
        sc[i, m:] = -np.inf
    return ids, sc


class PyNeighbourBuffer:
    """Line-by-line Python model of diskann/src/lib.rs:73-155 (binary_search_by modelled on the std algorithm)."""

    def __init__(self, size):
        self.ids, self.scores, self.visited, self.nu, self.size = [], [], [], None, size

    @staticmethod
    def _bsearch(scores, score):
        # core::slice::binary_search_by with f(x) = score.cmp(x)
        size = len(scores)
        if size == 0:
            return 0
        base = 0
        while size > 1:
            half = size // 2
            mid = base + half
            if not (score > scores[mid]):  # cmp != Greater
                base = mid
            size -= half
        x = scores[base]
        if score == x:
            return base
        return base + (1 if score < x else 0)

    def insert(self, id, score):
        if len(self.ids) == self.size and self.size > 0 and self.scores[-1] > score:
            return
        if self.size == 0:
            return
        loc = self._bsearch(self.scores, score)
        if loc < len(self.ids) and self.ids[loc] == id:
            return
        self.ids.insert(loc, id); self.scores.insert(loc, score); self.visited.insert(loc, False)
        del self.ids[self.size:], self.scores[self.size:], self.visited[self.size:]
        self.nu = loc if self.nu is None else min(loc, self.nu)

    def next_unvisited(self):
        if self.nu is None:
            return None
        cur = old = self.nu
        self.visited[cur] = True
        while cur < len(self.ids) and self.visited[cur]:
            cur += 1
        self.nu = None if cur == len(self.ids) else cur
        return self.ids[old]
