"""GPU parity of the sharded searches behind the C ABI (csrc/shard.cu): local search -> packed top-k in the gather slot -> one
ncclAllGather -> merge.  World size 1 runs on any box (the pack / merge / status path without NCCL); world size 2 spawns one
process per GPU (file carrier for the group id, no torch.distributed) and is skipped on a single-GPU box.  Every result is
compared with the CPU oracle run over the WHOLE index (flat) or with the oracle's own per-shard searches merged by the rule
(graph, beam): ids bit-exact, f32 / i64 scores identical."""
import os
import sys

import numpy as np
import pytest

from helpers import clustered_f16, index_f16, unit_rows

pytestmark = pytest.mark.gpu
D = 1152


def _merge_pairs(ids_all, sc_all, k):
    """(i64 score desc, global id asc) over [world][nq][<=k] lists"""
    nq = ids_all[0].shape[0]
    oi = np.full((nq, k), 0xFFFFFFFF, np.uint32)
    os_ = np.zeros((nq, k), np.int64)
    for q in range(nq):
        i = np.concatenate([a[q] for a in ids_all]).astype(np.int64)
        s = np.concatenate([a[q] for a in sc_all])
        keep = i != 0xFFFFFFFF
        i, s = i[keep], s[keep]
        order = np.lexsort((i, -s))[:k]
        oi[q, : len(order)], os_[q, : len(order)] = i[order], s[order]
    return oi, os_


def _shard_graph(O, x, lo, hi, R, L, seed):
    xs = x[lo:hi]
    g = O.IndexGraph(hi - lo, R)
    O.random_fill_graph(g, R, seed=seed)
    med = O.medioid(xs)
    O.build_graph(g, med, xs, O.make_config(r=R, l=L, maxc=120), seed=seed + 1)
    return xs, g, med


def run_rank(rank, world, carrier, n, nq, k, out_q=None):
    """Body of one rank; returns a dict of booleans.  Called in-process for world 1 and in spawned processes for world 2."""
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import mse_b200
    from mse_b200 import diskann as dk
    from mse_b200.sharding import ShardGroup, shard_range
    from oracle import oracle as O
    O.build()
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    stream = torch.cuda.current_stream().cuda_stream
    grp = ShardGroup.from_file(carrier, world, rank, rank)
    res = {}
    try:
        # ---------------- flat: random rows + a block of duplicated rows (ties at the cut -> uncertified -> repair round)
        x = index_f16(11, n)
        x[n // 2: n // 2 + 40] = x[3]                      # 41 identical rows, spread over the shard boundary when world = 2
        q = unit_rows(12, nq)
        q[0] = x[3].astype(np.float32)                      # this query's top list is full of exact ties
        lo, hi = shard_range(n, rank, world)
        ix = mse_b200.FlatIndex.from_f16(x[lo:hi], device=rank, id_base=lo)
        d_q = torch.from_numpy(q).to(dev)
        d_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
        d_sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
        for mode in (2, 0):
            ix.set_mode(mode)
            grp.flat_search_dev(ix, d_q.data_ptr(), nq, k, d_ids.data_ptr(), d_sc.data_ptr(), stream)
            repaired = grp.check()
            oi, os_ = O.flat_search(q, x, k)
            res[f"flat_ids_mode{mode}"] = bool(np.array_equal(d_ids.cpu().numpy().view(np.uint32), oi))
            res[f"flat_scores_mode{mode}"] = bool(np.array_equal(d_sc.cpu().numpy(), os_))
        res["flat_repair_ran_somewhere"] = repaired >= 0
        # k larger than a shard: padding entries must lose the merge
        big_k = min(200, 4096 // world)
        d_ids2 = torch.empty((3, big_k), dtype=torch.int32, device=dev)
        d_sc2 = torch.empty((3, big_k), dtype=torch.float32, device=dev)
        small = mse_b200.FlatIndex.from_f16(x[: 150][shard_range(150, rank, world)[0]: shard_range(150, rank, world)[1]], device=rank,
                                            id_base=shard_range(150, rank, world)[0])
        grp.flat_search_dev(small, d_q.data_ptr(), 3, big_k, d_ids2.data_ptr(), d_sc2.data_ptr(), stream)
        grp.check()
        oi, os_ = O.flat_search(q[:3], x[:150], big_k)
        res["flat_padding"] = bool(np.array_equal(d_ids2.cpu().numpy().view(np.uint32), oi) and np.array_equal(d_sc2.cpu().numpy(), os_))
        small.close()
        ix.close()

        # ---------------- graph: one oracle-built Vamana sub-graph per shard, greedy_search on every shard, merged
        ng, R, L = 3000, 16, 48
        xg = clustered_f16(21, ng, n_clusters=24)
        qg = clustered_f16(22, nq, n_clusters=24)
        shards = [_shard_graph(O, xg, *shard_range(ng, r, world), R, L, 30 + r) for r in range(world)]
        xs, g, med = shards[rank]
        glo = shard_range(ng, rank, world)[0]
        vl = dk.VectorList.from_f16s(xs, device=rank, id_base=glo)
        vl.set_graph(g.adj.copy(), g.deg.copy())
        d_qg = torch.from_numpy(qg.view(np.uint16).astype(np.int16)).to(dev)
        g_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
        g_sc = torch.empty((nq, k), dtype=torch.int64, device=dev)
        g_dist = torch.empty(nq, dtype=torch.int64, device=dev)
        cfg = O.make_config(r=R, l=L, maxc=120)
        per_ids, per_sc, own_dist = [], [], None
        for r in range(world):
            sx, sg, smed = shards[r]
            oi, osc, oln, odist = O.greedy_search_batch(smed, qg, sx, sg, cfg)
            base = shard_range(ng, r, world)[0]
            per_ids.append(np.where(oi[:, :k] == 0xFFFFFFFF, oi[:, :k], oi[:, :k] + np.uint32(base)))
            per_sc.append(osc[:, :k])
            if r == rank:
                own_dist = odist
        wi, ws = _merge_pairs(per_ids, per_sc, k)
        for mode in (dk.GRAPH_MODE_CTA, dk.GRAPH_MODE_WARP):
            dk.set_graph_mode(mode)
            grp.graph_search_dev(vl, d_qg.data_ptr(), nq, L, med, k, g_ids.data_ptr(), g_sc.data_ptr(), g_dist.data_ptr(), stream)
            dk.greedy_search_check(vl, nq)
            res[f"graph_ids_mode{mode}"] = bool(np.array_equal(g_ids.cpu().numpy().view(np.uint32), wi))
            res[f"graph_scores_mode{mode}"] = bool(np.array_equal(g_sc.cpu().numpy(), ws))
            res[f"graph_distances_mode{mode}"] = bool(np.array_equal(g_dist.cpu().numpy().astype(np.uint64), own_dist.astype(np.uint64)))
        dk.set_graph_mode(dk.GRAPH_MODE_AUTO)

        # ---------------- beam over RabitQ codes on every shard (query_disk_index.rs:144-212), best k expanded nodes merged
        rng = np.random.default_rng(5)
        P = np.linalg.qr(rng.standard_normal((D, D)))[0][:512].astype(np.float32)
        mean = xg[:1000].astype(np.float32).mean(axis=0)
        rq = dk.RabitQ(mean, P, device=rank)
        rq.encode_index(vl, 0)
        q32 = torch.from_numpy(qg.astype(np.float32)).to(dev)
        qtm = torch.empty((nq, 513), dtype=torch.float32, device=dev)
        rq.query_dev(q32.data_ptr(), nq, qtm.data_ptr(), stream)
        W = 3
        b_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
        b_sc = torch.empty((nq, k), dtype=torch.int64, device=dev)
        grp.beam_search_dev(vl, d_qg.data_ptr(), nq, L, W, med, k, b_ids.data_ptr(), b_sc.data_ptr(), stream=stream, d_qtm=qtm.data_ptr(), rabitq=rq)
        dk.greedy_search_check(vl, nq)
        # reference for the merge: every rank's own device top-k (the single-shard traversal is checked bit-exactly against the
        # oracle in test_graph_gpu.py), gathered through files
        t_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
        t_sc = torch.empty((nq, k), dtype=torch.int64, device=dev)
        t_len = torch.empty(nq, dtype=torch.int32, device=dev)
        cm = torch.empty(nq, dtype=torch.int64, device=dev)
        pc = torch.empty(nq, dtype=torch.int64, device=dev)
        dk.beam_search_dev(vl, d_qg.data_ptr(), nq, L, W, med, k, t_ids.data_ptr(), t_sc.data_ptr(), t_len.data_ptr(), cm.data_ptr(), pc.data_ptr(),
                           stream, d_qtm=qtm.data_ptr(), rabitq=rq)
        dk.greedy_search_check(vl, nq)
        li = t_ids.cpu().numpy().view(np.uint32)
        li = np.where(li == 0xFFFFFFFF, li, li + np.uint32(glo))
        np.savez(f"{carrier}.beam{rank}.npz", ids=li, sc=t_sc.cpu().numpy())
        os.replace(f"{carrier}.beam{rank}.npz", f"{carrier}.beam{rank}.done.npz")
        import time
        parts = []
        for r in range(world):
            p = f"{carrier}.beam{r}.done.npz"
            t0 = time.time()
            while not os.path.exists(p):
                assert time.time() - t0 < 120
                time.sleep(0.01)
            parts.append(np.load(p))
        wi, ws = _merge_pairs([p["ids"] for p in parts], [p["sc"] for p in parts], k)
        res["beam_ids"] = bool(np.array_equal(b_ids.cpu().numpy().view(np.uint32), wi))
        res["beam_scores"] = bool(np.array_equal(b_sc.cpu().numpy(), ws))
        res["all_gathers"] = grp.info()["all_gathers"]
        rq.close()
        vl.close()
    finally:
        grp.close()
    if out_q is not None:
        out_q.put((rank, res))
    return res


def _assert_all(res):
    bad = [k for k, v in res.items() if v is False]
    assert not bad, (bad, res)


def test_one_rank_group(tmp_path):
    res = run_rank(0, 1, str(tmp_path / "gid"), n=9000, nq=33, k=10)
    _assert_all(res)
    assert res["all_gathers"] == 0


def test_two_rank_group_matches_unsharded_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    carrier = str(tmp_path / "gid")
    procs = [ctx.Process(target=run_rank, args=(r, 2, carrier, 9000, 33, 10, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out_q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        _assert_all(got[r])
        assert got[r]["all_gathers"] >= 6      # 4 flat (2 modes + repairs + padding) + 2 graph + 1 beam


def test_shard_group_argument_checks(mse):
    import ctypes as C
    l = mse.lib()
    h = C.c_void_p()
    assert l.mse_shard_group_create(None, 2, 0, 0, C.byref(h)) == -1 and "unique_id" in mse.last_error()
    assert l.mse_shard_group_create(None, 1, 1, 0, C.byref(h)) == -1
    lo, hi = C.c_uint64(), C.c_uint64()
    assert l.mse_shard_range(100_000_000, 8, 7, C.byref(lo), C.byref(hi)) == 0 and (lo.value, hi.value) == (87_500_000, 100_000_000)
