"""a18 pinned to an EXECUTION OF THE REFERENCE: tests/golden/rabitq_reference.npz holds what /root/reference/diskann/rabitq.py itself
computed (tests/golden/make_rabitq_golden.py runs the unmodified script under runpy) -- its mean, its P, its sign codes, dots,
norms, approx_dot results and the head of the rabitq.msgpack it wrote.  The numpy restatement (oracle/rabitq_np.py), the C
oracle's direct estimate and -- on the GPU -- the CUDA codec (mse_rabitq_create / _load / _encode / _estimate / _query_dev) must
reproduce them: sign bits exactly, floating point within the stated tolerance."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "rabitq_reference.npz"))
BITS = np.unpackbits(G["qsample"], axis=1, bitorder="little").astype(bool)


def test_numpy_restatement_matches_the_script():
    from oracle.rabitq_np import RabitQ
    rq = RabitQ(G["mean"], G["p"])
    assert (rq.n_dims, rq.output_dims) == (int(G["n_dims"]), int(G["output_dims"])) == (1152, 512)
    bits, norms, dots, xs = rq.quantize(G["sample_rows_f16"])
    assert np.array_equal(bits, BITS)                                         # rabitq.py:30-33
    assert np.allclose(norms, G["norms"], rtol=2e-6)                          # :16
    assert np.allclose(dots, G["dots"], rtol=1e-5)                            # :34-35
    q = G["query0_f16"].astype(np.float32)
    assert np.abs(rq.approx_dot(bits, norms, dots, q) - G["approx_results"]).max() < 2e-6   # :42-48
    # the script's centred / normalised sample is what quantize() forms internally
    c = G["sample_rows_f16"].astype(np.float32) - G["mean"]
    assert np.allclose(c / np.linalg.norm(c, axis=1, keepdims=True), G["sample_centered"], atol=1e-6)


def test_msgpack_layout_matches_the_script():
    """rabitq.py:62-68 writes {"mean": [...], "transform": [...], "output_dims": 512, "n_dims": 1152} with f64 floats."""
    import msgpack
    from oracle.rabitq_np import RabitQ
    raw = RabitQ(G["mean"], G["p"]).to_msgpack()
    assert len(raw) == int(G["msgpack_len"])
    assert raw[:16] == G["msgpack_head"].tobytes()[:16]                       # map header, "mean" key, array header, first f64
    d = msgpack.unpackb(raw)
    assert list(d.keys()) == [str(k) for k in G["msgpack_keys"]]
    assert np.allclose(d["mean"][:8], G["mean_of_msgpack"], rtol=0, atol=0)   # f32 mean -> f64: exact
    assert np.allclose(d["transform"][:8], G["transform_of_msgpack"], rtol=1e-6)


def test_c_oracle_direct_estimate_matches_the_script(oracle):
    """the estimate the GPU traversal ranks candidates by (oracle/mse_oracle.c::rabitq_int_sum: query side quantised to 8 bits, exact
    integer popcount sum) against the script's f64 result: the quantisation costs < 5e-4 absolute"""
    p64, q = G["p"].astype(np.float64), G["query0_f16"].astype(np.float64)
    qtm = np.concatenate([p64 @ q, [G["mean"].astype(np.float64) @ q]]).astype(np.float32)
    scale = (G["norms"] * G["dots"]).astype(np.float32)
    got = oracle.rabitq_direct_estimates(qtm, np.float32(1.0 / np.sqrt(1152.0)), G["qsample"], scale)
    assert np.abs(got - G["approx_results"]).max() < 5e-4


@pytest.mark.gpu
def test_cuda_codec_matches_the_script(mse):
    from oracle.rabitq_np import RabitQ as NpRabitQ
    rows, q = G["sample_rows_f16"], G["query0_f16"].astype(np.float32)
    for make in (lambda: mse.diskann.RabitQ(G["mean"], G["p"]),
                 lambda: mse.diskann.RabitQ.from_msgpack(NpRabitQ(G["mean"], G["p"]).to_msgpack())):      # mse_rabitq_create / mse_rabitq_load
        rq = make()
        codes, norms, dots = rq.quantize(rows)
        assert np.array_equal(codes, G["qsample"])                            # every one of the 64 x 512 sign bits
        assert np.allclose(norms, G["norms"], rtol=1e-5) and np.allclose(dots, G["dots"], rtol=1e-4, atol=1e-6)
        assert np.abs(rq.approx_dot(codes, norms, dots, q) - G["approx_results"]).max() < 2e-4
        # fed the script's own norms / dots, the estimator alone
        assert np.abs(rq.approx_dot(G["qsample"], G["norms"].astype(np.float32), G["dots"].astype(np.float32), q) - G["approx_results"]).max() < 2e-4
        rq.close()


@pytest.mark.gpu
def test_cuda_query_side_and_traversal_estimate_match_the_script(mse, oracle):
    """mse_rabitq_query_dev (P q, <mean, q>) and the estimate the beam kernel forms from it (restated by the C oracle)"""
    import torch
    rq = mse.diskann.RabitQ(G["mean"], G["p"])
    q = torch.from_numpy(G["query0_f16"].astype(np.float32)[None]).cuda()
    qtm = torch.empty((1, 513), dtype=torch.float32, device="cuda")
    rq.query_dev(q.data_ptr(), 1, qtm.data_ptr())
    torch.cuda.synchronize()
    qtm = qtm.cpu().numpy()[0]
    want = np.concatenate([G["p"].astype(np.float64) @ G["query0_f16"].astype(np.float64), [G["mean"].astype(np.float64) @ G["query0_f16"].astype(np.float64)]])
    assert np.abs(qtm - want).max() < 2e-5
    est = oracle.rabitq_direct_estimates(qtm, np.float32(1.0 / np.sqrt(1152.0)), G["qsample"], (G["norms"] * G["dots"]).astype(np.float32))
    assert np.abs(est - G["approx_results"]).max() < 5e-4
    rq.close()


@pytest.mark.gpu
def test_train_on_device(mse):
    """mse_rabitq_train (rabitq.py:11-28 on the GPU): the mean is the script's mean of the sample; the transform has orthonormal rows;
    codes / estimates made with the trained codec agree with the numpy restatement given the same (mean, P); msgpack round trip."""
    from helpers import clustered_f16, unit_rows
    from oracle.rabitq_np import RabitQ as NpRabitQ
    x = clustered_f16(91, 2000, n_clusters=16)
    vl = mse.diskann.VectorList.from_f16s(x)
    rq = mse.diskann.RabitQ.train(vl, sample_rows=1500, output_dims=512, seed=7)
    mean, P = rq.export()
    assert np.allclose(mean, x[:1500].astype(np.float32).mean(axis=0), atol=2e-6)      # :14
    assert P.shape == (512, 1152) and np.abs(P.astype(np.float64) @ P.astype(np.float64).T - np.eye(512)).max() < 2e-6   # :22-28
    assert abs(float(P.mean())) < 5e-3 and 0.9 < float(P.std() * np.sqrt(1152.0)) < 1.1   # looks Haar: entries ~ N(0, 1/1152)
    other = mse.diskann.RabitQ.train(vl, sample_rows=1500, output_dims=512, seed=8)
    assert np.abs(other.export()[1] - P).max() > 1e-2                                   # the seed matters
    ref = NpRabitQ(mean, P)
    bits, norms, dots, xs = ref.quantize(x[:300])
    codes, gn, gd = rq.quantize(x[:300])
    diff = np.unpackbits(codes ^ NpRabitQ.pack(bits), axis=1, bitorder="little").astype(bool)
    assert (np.abs(xs[diff]) < 1e-6).all()
    q = unit_rows(92, 1)[0]
    assert np.abs(rq.approx_dot(codes, gn, gd, q) - ref.approx_dot(bits, norms, dots, q)).max() < 2e-4
    again = mse.diskann.RabitQ.from_msgpack(rq.to_msgpack())
    assert np.array_equal(again.quantize(x[:50])[0], codes[:50])
    for h in (rq, other, again):
        h.close()
    vl.close()
