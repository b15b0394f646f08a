"""numpy restatement of diskann/rabitq.py:8-48 -- TEST INFRASTRUCTURE ONLY (tests/ and bench.py's cpu_baseline).

PINNED TO AN EXECUTION OF THE REFERENCE: tests/golden/make_rabitq_golden.py runs the unmodified script (runpy) on seeded inputs and
tests/test_rabitq_reference.py holds this restatement to what it printed (sign bits exact, floats to 2e-6).  It is restated
line for line: centre by the dataset mean (:14-16), normalise and keep the norm (:17), P = first `output_dims` rows of a
random orthogonal matrix (:22-28), code = sign(P o_hat) (:30-33), dots = <o_bar, P o_hat> with o_bar = +-1/sqrt(n_dims)
(:34-35), estimate = |o| * <o_bar, P q> * dots + <mean, q> (:42-48).
"""
from __future__ import annotations

import numpy as np


def random_ortho(dim: int, seed: int) -> np.ndarray:
    h = np.random.default_rng(seed).standard_normal((dim, dim))
    q, _ = np.linalg.qr(h)
    return q


class RabitQ:
    def __init__(self, mean: np.ndarray, transform: np.ndarray):
        self.mean = np.asarray(mean, np.float32)
        self.p = np.asarray(transform, np.float32)        # [output_dims, n_dims]
        self.n_dims = self.p.shape[1]
        self.output_dims = self.p.shape[0]
        self.scale = 1.0 / np.sqrt(self.n_dims)

    @classmethod
    def train(cls, dataset: np.ndarray, output_dims: int = 512, seed: int = 0) -> "RabitQ":
        d = dataset.shape[1]
        return cls(np.mean(dataset.astype(np.float32), axis=0), random_ortho(d, seed)[:output_dims, :])

    def quantize(self, x: np.ndarray):
        c = x.astype(np.float32) - self.mean
        norms = np.linalg.norm(c, axis=1)
        c = c / norms[:, None]
        xs = (self.p.astype(np.float64) @ c.astype(np.float64).T).T
        bits = xs > 0
        dots = np.sum(self.scale * (2.0 * bits - 1.0) * xs, axis=1)
        return bits, norms.astype(np.float32), dots.astype(np.float32), xs

    @staticmethod
    def pack(bits: np.ndarray) -> np.ndarray:
        """bit i of byte b = sign of output 8b+i (little-endian bit order, as the CUDA ballot packs it)."""
        return np.packbits(bits, axis=1, bitorder="little")

    def approx_dot(self, bits, norms, dots, q: np.ndarray) -> np.ndarray:
        q = q.astype(np.float32)
        mean_to_query = float(np.dot(self.mean.astype(np.float64), q.astype(np.float64)))
        qt = self.p.astype(np.float64) @ q.astype(np.float64)
        o_bar_dot_q = np.sum(self.scale * (2.0 * bits - 1.0) * qt, axis=1)
        return norms * o_bar_dot_q * dots + mean_to_query

    def to_msgpack(self) -> bytes:
        import msgpack
        return msgpack.packb({"mean": self.mean.flatten().tolist(), "transform": self.p.flatten().tolist(),
                              "output_dims": int(self.output_dims), "n_dims": int(self.n_dims)})
