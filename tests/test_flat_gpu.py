"""GPU parity: flat (FAISS QT_fp16 / IP) search through the C ABI against the oracle.  Integer outputs
(ids) must match bit for bit; f32 scores are f32-rounded f64 sums on both sides and must be identical."""
import numpy as np
import pytest

from helpers import clustered_f16, index_f16, np_flat_topk, unit_rows

pytestmark = pytest.mark.gpu


def _check(mse, oracle, x, q, k, mode=0, id_base=0):
    ix = mse.FlatIndex.from_f16(x, id_base=id_base)
    ix.set_mode(mode)
    sc, lab = ix.search(q, k)
    oi, os_ = oracle.flat_search(q, x, k)
    want = oi.astype(np.int64)
    want[oi == 0xFFFFFFFF] = -1
    want[want >= 0] += id_base
    assert np.array_equal(lab, want), (np.argwhere(lab != want)[:5], ix.stats())
    assert np.array_equal(sc, os_)
    st = ix.stats()
    ix.close()
    return st


def test_config_c1_golden(mse):
    """BASELINE config[0]: one query, top-10 over 1k x 1152 (fixture from tests/golden/make_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "flat_topk.npz"))
    ix = mse.FlatIndex.from_f16(index_f16(0, 1000))
    sc, lab = ix.search(unit_rows(3, 1), 10)
    assert lab.tolist() == g["c1_ids"].tolist() and np.array_equal(sc, g["c1_scores"])
    assert ix.ntotal == 1000


@pytest.mark.parametrize("n", [1, 7, 255, 4096, 4097, 12345])
@pytest.mark.parametrize("nq", [1, 2, 3])
def test_exact_scan_sizes(mse, oracle, n, nq):
    _check(mse, oracle, index_f16(n, n), unit_rows(100 + n, nq) * np.float32(1.3), 10, mode=1)


@pytest.mark.parametrize("n,nq,k", [(300, 5, 10), (4096, 128, 100), (20000, 130, 100), (50001, 300, 10), (9000, 17, 1000)])
def test_tensor_path_sizes(mse, oracle, n, nq, k):
    st = _check(mse, oracle, index_f16(n + 1, n), unit_rows(200 + n, nq), k, mode=2)
    assert st["tensor_queries"] == nq


def test_auto_mode_and_id_base(mse, oracle):
    x = index_f16(77, 30000)
    st = _check(mse, oracle, x, unit_rows(78, 64), 100, mode=0, id_base=1_000_000)
    assert st["tensor_queries"] == 64
    st = _check(mse, oracle, x, unit_rows(79, 1), 100, mode=0, id_base=5)
    assert st["exact_queries"] == 1


def test_clustered_data(mse, oracle):
    x = clustered_f16(5, 40000, n_clusters=32, sigma=0.3)
    q = clustered_f16(6, 96, n_clusters=32, sigma=0.3).astype(np.float32)
    _check(mse, oracle, x, q, 100, mode=2)


def test_duplicates_force_exact_rerun(mse, oracle):
    """Many exact duplicates: ties at the cut cannot be certified, the affected queries are re-run exactly."""
    base = index_f16(9, 500)
    x = np.concatenate([base] * 40)  # every row 40 times
    q = unit_rows(10, 9)
    st = _check(mse, oracle, x, q, 100, mode=0)
    assert st["tensor_queries"] == 9


def test_adversarial_order_overflow(mse, oracle):
    """Rows sorted by ascending score for query 0: every row beats the running threshold -> buffer overflow path."""
    x = index_f16(12, 30000)
    q = unit_rows(13, 4)
    order = np.argsort(x.astype(np.float32) @ q[0])
    xs = np.ascontiguousarray(x[order])
    st = _check(mse, oracle, xs, q, 10, mode=0)
    assert st["overflows"] >= 1
    _check(mse, oracle, xs, q[:1], 10, mode=1)


def test_add_f32_rounds_like_qt_fp16(mse, oracle):
    x32 = unit_rows(14, 3000)
    ix = mse.FlatIndex(1152)
    for i in range(0, 3000, 1024):  # src/main.rs:815 adds 1024 rows at a time
        ix.add(x32[i:i + 1024])
    assert ix.ntotal == 3000
    q = unit_rows(15, 2)
    sc, lab = ix.search(q, 10)
    oi, os_ = oracle.flat_search(q, x32.astype(np.float16), 10)
    assert np.array_equal(lab, oi.astype(np.int64)) and np.array_equal(sc, os_)


def test_empty_index_and_k_gt_n(mse):
    ix = mse.FlatIndex(1152)
    sc, lab = ix.search(unit_rows(1, 2), 5)
    assert (lab == -1).all() and np.isneginf(sc).all()
    ix.add_f16(index_f16(2, 3))
    sc, lab = ix.search(unit_rows(1, 2), 5)
    assert (lab[:, 3:] == -1).all() and (lab[:, :3] >= 0).all()


def test_merge_topk(mse, oracle):
    import torch
    x = index_f16(20, 8000)
    q = unit_rows(21, 33)
    k, shards = 10, 4
    per = 2000
    ids_l, sc_l = [], []
    for s in range(shards):
        ix = mse.FlatIndex.from_f16(x[s * per:(s + 1) * per], id_base=s * per)
        sc, lab = ix.search(q, k)
        ids_l.append(lab.astype(np.uint32)); sc_l.append(sc)
    ids = torch.from_numpy(np.stack(ids_l).astype(np.int64)).to(torch.int32).cuda()  # [shards, nq, k] bit pattern of u32
    scs = torch.from_numpy(np.stack(sc_l)).cuda()
    out_i = torch.empty((33, k), dtype=torch.int32, device="cuda")
    out_s = torch.empty((33, k), dtype=torch.float32, device="cuda")
    mse.merge_topk(0, ids.data_ptr(), scs.data_ptr(), shards, 33, k, out_i.data_ptr(), out_s.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    oi, os_ = oracle.flat_search(q, x, k)
    assert np.array_equal(out_i.cpu().numpy().astype(np.uint32), oi) and np.array_equal(out_s.cpu().numpy(), os_)
