"""Generates the committed fixtures in this directory.

The reference holds no golden vectors for this path and cannot be built or imported here, so
these fixtures are produced by INDEPENDENT pure-numpy / pure-Python models of the cited reference lines
(tests/helpers.py), not by the C oracle: they pin the oracle (and through it the CUDA path) from a second side.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import PyNeighbourBuffer, index_f16, np_fast_dot, np_flat_topk, unit_rows  # noqa: E402


def main():
    # 1. fast_dot known answers: 64 pairs of 1152-d fp16 vectors (incl. large-magnitude and denormal lanes)
    rng = np.random.default_rng(20240501)
    a = rng.standard_normal((64, 1152)).astype(np.float16)
    b = rng.standard_normal((64, 1152)).astype(np.float16)
    a[3] *= np.float16(30.0); b[3] *= np.float16(30.0)
    a[4, ::7] = np.float16(6e-8); b[5, ::3] = np.float16(-6e-8)
    a[6] = 0
    out = np.array([np_fast_dot(a[i], b[i]) for i in range(64)], np.int64)
    np.savez_compressed(os.path.join(HERE, "fast_dot_kat.npz"), a=a.view(np.uint16), b=b.view(np.uint16), out=out)

    # 2. NeighbourBuffer trace: random inserts (with duplicate ids and tied scores) interleaved with next_unvisited
    rng = np.random.default_rng(7)
    ops, results = [], []
    nb = PyNeighbourBuffer(24)
    for step in range(600):
        if rng.random() < 0.7:
            id_ = int(rng.integers(0, 80)); sc = int(rng.integers(-40, 40)) * 1000
            nb.insert(id_, sc); ops.append((0, id_, sc)); results.append(-2)
        else:
            r = nb.next_unvisited(); ops.append((1, 0, 0)); results.append(-1 if r is None else r)
    np.savez_compressed(os.path.join(HERE, "neighbour_buffer_trace.npz"), ops=np.array(ops, np.int64),
                        results=np.array(results, np.int64), final_ids=np.array(nb.ids, np.int64),
                        final_scores=np.array(nb.scores, np.int64))

    # 3. flat top-10 for config C1 (1 query over 1k x 1152, SURVEY 8d seeds) + a 3-query / 2k-row case
    x = index_f16(0, 1000)
    q = unit_rows(3, 1)
    ids, sc = np_flat_topk(q, x, 10)
    x2 = index_f16(2, 2000)
    q2 = unit_rows(5, 3) * np.float32(1.7)  # un-normalised queries (common.rs:215-274)
    ids2, sc2 = np_flat_topk(q2, x2, 25)
    np.savez_compressed(os.path.join(HERE, "flat_topk.npz"), c1_ids=ids, c1_scores=sc, m_ids=ids2, m_scores=sc2)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
