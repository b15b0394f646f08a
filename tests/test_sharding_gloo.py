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


def test_shard_ranges_cover_everything():
    import mse_b200
    from mse_b200.sharding import shard_range
    for n in (0, 1, 7, 10_000_000, 100_000_001):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, g, w) for g in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
