"""world_size-2 gloo test (CPU) of the N > 1 host logic: id-range shards, global ids via id_base, all-gather layout, and the
merge rule.  The per-shard searches and the merge are done by the CPU checker here; on GPUs they are mse_search_flat_dev and
mse_merge_topk_dev (tests/test_flat_gpu.py::test_merge_topk checks the kernel against the same rule)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import index_f16, unit_rows


def merge_rule(ids, scores, k):
    """checker: k-way merge of [world, nq, k] lists by (score desc, id asc); padding id 0xFFFFFFFF sorts last"""
    w, nq, _ = ids.shape
    out_i = np.empty((nq, k), np.uint32)
    out_s = np.empty((nq, k), np.float32)
    for q in range(nq):
        i = ids[:, q].reshape(-1)
        s = scores[:, q].reshape(-1)
        valid = i != 0xFFFFFFFF
        order = np.lexsort((i, -s.astype(np.float64), ~valid))[:k]
        out_i[q], out_s[q] = i[order], s[order]
    return out_i, out_s


def _worker(rank, world, port, n, nq, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import mse_b200
    from mse_b200.sharding import all_gather_topk, shard_range
    from oracle import oracle as O
    x = index_f16(5, n)
    q = unit_rows(6, nq)
    lo, hi = shard_range(n, rank, world)
    ids, sc = O.flat_search(q, x[lo:hi], k)                      # this rank's shard (CPU checker stands in for the kernel)
    ids = np.where(ids == 0xFFFFFFFF, ids, ids + np.uint32(lo))  # id_base
    ids_all, sc_all = all_gather_topk(dist, torch.from_numpy(ids.astype(np.int64)), torch.from_numpy(sc), world)
    mi, ms = merge_rule(ids_all.numpy().astype(np.uint32), sc_all.numpy(), k)
    gi, gs = O.flat_search(q, x, k)
    ok = bool(np.array_equal(mi, gi) and np.array_equal(ms, gs))
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(3001, 10), (37, 25)])
def test_two_rank_sharded_flat_search(n, k):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, 5, k, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


def _graph_worker(rank, world, port, n, nq, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import mse_b200
    from mse_b200.sharding import all_gather_topk, shard_range
    from oracle import oracle as O
    from helpers import clustered_f16
    R, L = 16, 40
    x = clustered_f16(7, n, n_clusters=12)
    q = clustered_f16(8, nq, n_clusters=12)
    cfg = O.make_config(r=R, l=L, maxc=200)

    def shard_search(g_rank):
        """independent sub-graph over the shard's id range, greedy search, global ids"""
        lo, hi = shard_range(n, g_rank, world)
        xs = x[lo:hi]
        g = O.IndexGraph(hi - lo, R)
        O.random_fill_graph(g, R, seed=3 + g_rank)
        med = O.medioid(xs)
        O.build_graph(g, med, xs, cfg, seed=5 + g_rank)
        ids, sc, ln, _ = O.greedy_search_batch(med, q, xs, g, cfg)
        return (ids[:, :k].astype(np.int64) + lo), (sc[:, :k].astype(np.float64) / 4294967296.0).astype(np.float32)

    ids, sc = shard_search(rank)
    ids_all, sc_all = all_gather_topk(dist, torch.from_numpy(ids), torch.from_numpy(sc), world)
    mi, ms = merge_rule(ids_all.numpy().astype(np.uint32), sc_all.numpy(), k)
    ok = True
    if rank == 0:
        # the gathered layout is [shard][query][k] in rank order, whatever rank computes it
        every = [shard_search(r) for r in range(world)]
        wi, ws = merge_rule(np.stack([e[0] for e in every]).astype(np.uint32), np.stack([e[1] for e in every]), k)
        ok = bool(np.array_equal(mi, wi) and np.array_equal(ms, ws))
        gi, _ = O.flat_search(q.astype(np.float32), x, k)
        rec = np.mean([len(set(mi[i].tolist()) & set(gi[i].tolist())) / k for i in range(nq)])
        ok = ok and rec >= 0.9
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()


def test_two_rank_sharded_graph_search():
    """Graph path at N > 1: one independent Vamana sub-graph per id-range shard, every query searched on every shard, one
    all-gather of per-shard top-k and the same merge as the flat path; recall@10 of the merged result vs brute force."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_graph_worker, args=(r, 2, port, 1500, 24, 10, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


def test_shard_ranges_cover_everything():
    import mse_b200
    from mse_b200.sharding import shard_range
    for n in (0, 1, 7, 10_000_000, 100_000_001):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, g, w) for g in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
