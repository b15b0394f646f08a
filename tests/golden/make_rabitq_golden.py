"""Generates tests/golden/rabitq_reference.npz by EXECUTING the reference's own script, /root/reference/diskann/rabitq.py, unmodified
(runpy, in a scratch directory holding the two input files it reads).  Run in the build container only (the GPU box has no
/root/reference); the fixture it writes is what pins oracle/rabitq_np.py and the CUDA codec (mse_rabitq_*) to the reference.

  inputs  embeddings.bin: 2000 clustered unit rows, fp16 (seed 91); query.bin: 8 unit rows, fp16 (seed 92)
  seed    np.random.seed(20261017) before the script runs (its random_ortho draws from the global numpy RNG, rabitq.py:22-25)
  stored  mean, p (first 512 rows of the orthogonal matrix, as f32), the 64 sample rows, query 0, and the script's own outputs:
          qsample (sign bits), dots, norms[:64], approx_results, exact_results, plus the head / length / key order of the
          rabitq.msgpack it writes (the 5 MB file itself is not committed)
"""
import contextlib
import io
import os
import runpy
import sys
import tempfile

import msgpack
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import clustered_f16, unit_rows  # noqa: E402

SCRIPT = "/root/reference/diskann/rabitq.py"


def main():
    data = clustered_f16(91, 2000, n_clusters=16)
    queries = unit_rows(92, 8).astype(np.float16)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        data.tofile(os.path.join(tmp, "embeddings.bin"))
        queries.tofile(os.path.join(tmp, "query.bin"))
        os.chdir(tmp)
        try:
            np.random.seed(20261017)
            with contextlib.redirect_stdout(io.StringIO()):
                g = runpy.run_path(SCRIPT, run_name="__main__")
            raw = open("rabitq.msgpack", "rb").read()
        finally:
            os.chdir(cwd)
    packed = msgpack.unpackb(raw)
    assert list(packed.keys()) == ["mean", "transform", "output_dims", "n_dims"]
    p32 = g["p"].astype(np.float32)
    # the f32 copy of P must not flip any of the script's sign bits on the sample (it does not: |P o| >> f32 rounding)
    xs32 = (p32.astype(np.float64) @ g["sample"].astype(np.float64).T).T
    assert np.array_equal(xs32 > 0, g["qsample"])
    np.savez_compressed(
        os.path.join(HERE, "rabitq_reference.npz"),
        mean=g["mean"].astype(np.float32), p=p32, sample_rows_f16=data[:64], sample_centered=g["sample"].astype(np.float32),
        query0_f16=queries[0], qsample=np.packbits(g["qsample"], axis=1, bitorder="little"), dots=g["dots"].astype(np.float64),
        norms=g["norms"][:64].astype(np.float64), approx_results=g["approx_results"].astype(np.float64),
        exact_results=g["exact_results"].astype(np.float64), msgpack_head=np.frombuffer(raw[:64], np.uint8),
        msgpack_len=np.int64(len(raw)), msgpack_keys=np.array(list(packed.keys())), output_dims=np.int64(packed["output_dims"]),
        n_dims=np.int64(packed["n_dims"]), mean_of_msgpack=np.asarray(packed["mean"], np.float64)[:8],
        transform_of_msgpack=np.asarray(packed["transform"], np.float64)[:8])
    print("wrote rabitq_reference.npz:", {k: (v.shape, v.dtype) for k, v in np.load(os.path.join(HERE, "rabitq_reference.npz")).items()})


if __name__ == "__main__":
    main()
