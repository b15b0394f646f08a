"""Id-range sharding of the index across the GPUs of one box: rank g owns the vector ids
[n*g/N, n*(g+1)/N), queries are replicated, every rank searches its shard, one all-gather of the per-shard top-k
([world, nq, k] ids + scores) and a k-way merge by (score desc, id asc) on every rank.  Because ids are global
(the shard's id_base) and the order is total, the merged result does not depend on the number of shards."""
from __future__ import annotations


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    return n_total * rank // world, n_total * (rank + 1) // world


def all_gather_topk(dist, ids_local, scores_local, world: int):
    """ids_local/scores_local: torch tensors [nq, k] on this rank -> ([world, nq, k], [world, nq, k]) in rank order."""
    import torch
    ids_all = torch.empty((world,) + tuple(ids_local.shape), dtype=ids_local.dtype, device=ids_local.device)
    sc_all = torch.empty((world,) + tuple(scores_local.shape), dtype=scores_local.dtype, device=scores_local.device)
    if world == 1:
        ids_all[0], sc_all[0] = ids_local, scores_local
    else:
        # concatenated-along-dim-0 output form: accepted by both the NCCL and the gloo backends
        dist.all_gather_into_tensor(ids_all.view(-1, *ids_local.shape[1:]), ids_local.contiguous())
        dist.all_gather_into_tensor(sc_all.view(-1, *scores_local.shape[1:]), scores_local.contiguous())
    return ids_all, sc_all
