"""Generates tests/golden/kmeans_reference.json by EXECUTING the reference's own `simulated_annealing` (kmeans.py:73-131, with the
`fitness` closure :78-95 inside it).  kmeans.py runs its whole experiment at import time (reads 500k_vecs.bin, needs a CUDA device),
so the function's source is taken from the file with `ast` and executed UNCHANGED on the CPU with torch; the script prints
`last_fitness, new_fitness, temperature` every iteration (:106) and those lines are the golden values.  Run in the build container
only (the GPU box has no /root/reference).

  rows     helpers.clustered_f16(seed 71, 3000 rows, 16 clusters) as f32 -- what `np.fromfile(...).astype(np.float32)` gives (:9)
  k        8 centroids, SPILL_K = 2 (the script's), max_iter = 6, batch_size = 1024 (so the batch loop of :84 runs three times)
  RNG      torch.manual_seed(20261017); the test replays the same torch.randn / randn_like calls to obtain the candidate centroids of
           every iteration (no reroll happens within 6 iterations) and holds oracle/kmeans_np.py and the CUDA kernel to the printed values
"""
import ast
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import clustered_f16  # noqa: E402

SCRIPT = "/root/reference/kmeans.py"
SEED, ROWS, K, ITERS, BATCH = 20261017, 3000, 8, 6, 1024


def reference_function():
    tree = ast.parse(open(SCRIPT).read())
    keep = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name == "simulated_annealing") or
            (isinstance(n, ast.Assign) and any(getattr(t, "id", "") in ("SPILL_K", "n_dims") for t in n.targets))]
    ns = {"torch": torch, "np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), SCRIPT, "exec"), ns)
    return ns


def main():
    ns = reference_function()
    assert ns["n_dims"] == 1152 and ns["SPILL_K"] == 2
    data = torch.tensor(clustered_f16(71, ROWS, n_clusters=16).astype(np.float32))
    torch.manual_seed(SEED)
    buf = io.StringIO()
    # the script's cluster_sizes never leave `fitness`; torch.bincount (kmeans.py:88) is observed while the unchanged function runs
    seen = []
    real_bincount = torch.bincount

    def spy(*a, **kw):
        r = real_bincount(*a, **kw)
        seen.append(r.tolist())
        return r
    torch.bincount = spy
    try:
        with contextlib.redirect_stdout(buf):
            out = ns["simulated_annealing"](data, K, max_iter=ITERS, batch_size=BATCH)
    finally:
        torch.bincount = real_bincount
    n_batches = -(-ROWS // BATCH)
    per_call = n_batches * ns["SPILL_K"]                      # bincount calls per fitness evaluation: batches x ranks (:84-89)
    assert len(seen) == per_call * (ITERS + 1)
    cluster_sizes = []
    for c in range(ITERS + 1):
        calls = seen[c * per_call:(c + 1) * per_call]
        cluster_sizes.append([[sum(calls[b * ns["SPILL_K"] + j][i] for b in range(n_batches)) for i in range(K)] for j in range(ns["SPILL_K"])])
    lines = [ln.split() for ln in buf.getvalue().strip().splitlines()]
    assert len(lines) == ITERS and all(len(ln) == 3 for ln in lines), buf.getvalue()
    golden = {"script": "kmeans.py:73-131 executed unchanged (ast-extracted simulated_annealing)", "seed": SEED, "rows": ROWS, "k": K, "iters": ITERS,
              "batch_size": BATCH, "data": "helpers.clustered_f16(71, 3000, n_clusters=16)",
              "printed": [[float(a), float(b), float(c)] for a, b, c in lines],
              "cluster_sizes": cluster_sizes,   # [evaluation][rank][cluster], evaluation 0 = the initial centroids, i >= 1 = iteration i's candidate
              "result_head": [float(v) for v in out[0, :8]], "result_norms": [float(v) for v in out.norm(dim=1)]}
    with open(os.path.join(HERE, "kmeans_reference.json"), "w") as f:
        json.dump(golden, f, indent=1)
    print(json.dumps(golden["printed"]))


if __name__ == "__main__":
    main()
