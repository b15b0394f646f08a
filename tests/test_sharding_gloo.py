"""world_size-2 gloo tests (CPU) of the N > 1 path's host logic and wire format: id-range shards (mse_shard_range), global ids
via id_base, the packed entries that travel in the single all-gather (u64 rank keys for the flat search, (i64 score, id) pairs
for the graph searches), the [shard][query][k] slot layout, the merge rule, and the carrier of the group id.  The per-shard
searches are done by the CPU checker here; on GPUs the same flow runs inside libmse_b200.so (csrc/shard.cu:
mse_search_flat_sharded_dev / mse_search_graph_sharded_dev), checked on hardware by tests/test_sharded_gpu.py."""
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


def all_gather_slots(slot: torch.Tensor, world: int) -> torch.Tensor:
    """one all-gather of every rank's slot -> [world, *slot.shape] in rank order (what ncclAllGather does in place on the GPUs)"""
    out = torch.empty((world,) + tuple(slot.shape), dtype=slot.dtype)
    dist.all_gather_into_tensor(out.view(-1, *slot.shape[1:]), slot.contiguous())
    return out


def rank_keys(scores_f32: np.ndarray, ids_u32: np.ndarray) -> np.ndarray:
    """csrc/common.cuh rank_key: (order-preserving map of the f32 score) << 32 | ~id; 0 = no entry"""
    u = (scores_f32.astype(np.float32) + np.float32(0)).view(np.uint32).astype(np.uint64)
    o = np.where(u & 0x80000000, ~u & 0xFFFFFFFF, u | 0x80000000)
    key = (o << np.uint64(32)) | ((~ids_u32.astype(np.uint64)) & np.uint64(0xFFFFFFFF))
    return np.where(ids_u32 == 0xFFFFFFFF, np.uint64(0), key)


def _merge_keys(slots: np.ndarray, k: int):
    w, nq, _ = slots.shape
    flat = slots.transpose(1, 0, 2).reshape(nq, -1).astype(np.uint64)
    flat = np.sort(flat, axis=1)[:, ::-1][:, :k]
    ids = (~flat & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    o = (flat >> np.uint64(32)).astype(np.uint32)
    u = np.where(o & 0x80000000, o & 0x7FFFFFFF, ~o)
    sc = u.astype(np.uint32).view(np.float32)
    ids = np.where(flat == 0, np.uint32(0xFFFFFFFF), ids)
    sc = np.where(flat == 0, np.float32(-np.inf), sc)
    return ids, sc


def _worker(rank, world, port, n, nq, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import mse_b200
    from mse_b200.sharding import shard_range
    from oracle import oracle as O
    x = index_f16(5, n)
    q = unit_rows(6, nq)
    lo, hi = shard_range(n, rank, world)
    ids, sc = O.flat_search(q, x[lo:hi], k)                      # this rank's shard (CPU checker stands in for the kernel)
    ids = np.where(ids == 0xFFFFFFFF, ids, ids + np.uint32(lo))  # id_base
    # the slot this rank contributes: packed rank keys with global ids, exactly what k_finalize writes into the gather buffer
    slot = torch.from_numpy(rank_keys(sc, ids.astype(np.uint32)).view(np.int64))
    slots = all_gather_slots(slot, world).numpy().view(np.uint64)
    mi, ms = _merge_keys(slots, k)
    gi, gs = O.flat_search(q, x, k)
    ok = bool(np.array_equal(mi, gi) and np.array_equal(ms, gs))
    # and the same through the (ids, scores) form of mse_merge_topk_dev
    ids_all = all_gather_slots(torch.from_numpy(ids.astype(np.int64)), world)
    sc_all = all_gather_slots(torch.from_numpy(sc), world)
    mi2, ms2 = merge_rule(ids_all.numpy().astype(np.uint32), sc_all.numpy(), k)
    ok = ok and bool(np.array_equal(mi2, gi) and np.array_equal(ms2, gs))
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
    from mse_b200.sharding import shard_range
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
        return (ids[:, :k].astype(np.int64) + lo), sc[:, :k].astype(np.int64)

    def merge_pairs(ids_all, sc_all):
        """(i64 score desc, id asc) over [world, nq, k] -- k_merge_pairs; scores stay fixed-point i64, no float round trip"""
        w, nq_, _ = ids_all.shape
        oi, os_ = np.empty((nq_, k), np.int64), np.empty((nq_, k), np.int64)
        for qi in range(nq_):
            i, s_ = ids_all[:, qi].reshape(-1), sc_all[:, qi].reshape(-1)
            order = np.lexsort((i, -s_))[:k]
            oi[qi], os_[qi] = i[order], s_[order]
        return oi, os_

    ids, sc = shard_search(rank)
    # one gather of (score, id) pairs: [nq, k, 2] i64 per rank
    pairs = all_gather_slots(torch.from_numpy(np.stack([sc, ids], axis=2)), world).numpy()
    mi, ms = merge_pairs(pairs[..., 1], pairs[..., 0])
    ok = True
    if rank == 0:
        # the gathered layout is [shard][query][k] in rank order, whatever rank computes it
        every = [shard_search(r) for r in range(world)]
        wi, ws = merge_pairs(np.stack([e[0] for e in every]), np.stack([e[1] for e in every]))
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


def _id_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import mse_b200
    from mse_b200 import sharding
    made = {}

    def fake_init(self, group_id, n_ranks, rank_, device):      # no GPU here: record what the C ABI would be handed
        made.update(id=group_id, n=n_ranks, r=rank_, d=device)
    sharding.ShardGroup.__init__ = fake_init
    sharding.ShardGroup.close = lambda self: None
    sharding.new_group_id = lambda: bytes(range(128))           # rank 0's ncclGetUniqueId bytes
    sharding.ShardGroup.from_torch_distributed(dist, device=rank)
    ok = made == {"id": bytes(range(128)), "n": world, "r": rank, "d": rank}
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()


def test_group_id_reaches_every_rank():
    """rank 0 draws the 128-byte group id, every rank receives the same bytes with its own rank / device"""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_id_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


def test_group_id_file_carrier(tmp_path, monkeypatch):
    import mse_b200
    from mse_b200 import sharding
    made = []
    monkeypatch.setattr(sharding.ShardGroup, "__init__", lambda self, gid, n, r, d: made.append((gid, n, r, d)))
    monkeypatch.setattr(sharding.ShardGroup, "close", lambda self: None)
    monkeypatch.setattr(sharding, "new_group_id", lambda: b"\x07" * 128)
    path = str(tmp_path / "group.id")
    sharding.ShardGroup.from_file(path, 2, 0, 0)
    sharding.ShardGroup.from_file(path, 2, 1, 1)
    assert made == [(b"\x07" * 128, 2, 0, 0), (b"\x07" * 128, 2, 1, 1)]
    with pytest.raises(TimeoutError):
        sharding.ShardGroup.from_file(str(tmp_path / "absent.id"), 2, 1, 1, timeout_s=0.05)
