"""Host-side handle on the C ABI's shard group (mse_shard_group_*, csrc/shard.cu): id-range sharding of the index over the
GPUs of one box.  Rank g owns the vector ids [n*g/N, n*(g+1)/N); queries are replicated; every rank searches its
shard; the per-shard top-k lists meet in one NCCL all-gather inside the library and are merged by (score desc, id asc) on
every rank.  Ids are global (the shard's id_base) and the order is total, so the result does not depend on the shard count.

The only thing the host has to do is hand rank 0's 128-byte group id to the other ranks.  Two carriers are provided:
torch.distributed (any backend; what bench.py uses under torchrun) and a plain file (for hosts without torch)."""
from __future__ import annotations

import ctypes as C
import os
import time

from ._lib import check, lib

ID_BYTES = 128


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    lo, hi = C.c_uint64(), C.c_uint64()
    check(lib().mse_shard_range(n_total, world, rank, C.byref(lo), C.byref(hi)), "mse_shard_range")
    return int(lo.value), int(hi.value)


def new_group_id() -> bytes:
    buf = (C.c_uint8 * ID_BYTES)()
    check(lib().mse_shard_group_unique_id(buf), "mse_shard_group_unique_id")
    return bytes(buf)


class ShardGroup:
    def __init__(self, group_id: bytes | None, n_ranks: int, rank: int, device: int):
        self._h = C.c_void_p()
        idbuf = (C.c_uint8 * ID_BYTES).from_buffer_copy(group_id) if group_id is not None else None
        check(lib().mse_shard_group_create(idbuf, n_ranks, rank, device, C.byref(self._h)), "mse_shard_group_create")
        self.n_ranks, self.rank, self.device = n_ranks, rank, device

    # -- carriers of the group id ----------------------------------------------------------------------------------
    @classmethod
    def from_torch_distributed(cls, dist, device: int) -> "ShardGroup":
        """Every rank of an initialised torch.distributed job (NCCL or gloo) calls this."""
        world, rank = (dist.get_world_size(), dist.get_rank()) if dist is not None and dist.is_initialized() else (1, 0)
        if world == 1:
            return cls(None, 1, 0, device)
        box = [new_group_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(box[0], world, rank, device)

    @classmethod
    def from_file(cls, path: str, n_ranks: int, rank: int, device: int, timeout_s: float = 120.0) -> "ShardGroup":
        """Rank 0 writes the id to `path` (atomically); the others wait for it."""
        if n_ranks == 1:
            return cls(None, 1, 0, device)
        if rank == 0:
            tmp = f"{path}.{os.getpid()}.tmp"
            with open(tmp, "wb") as f:
                f.write(new_group_id())
            os.replace(tmp, path)
        deadline = time.monotonic() + timeout_s
        while not (os.path.exists(path) and os.path.getsize(path) == ID_BYTES):
            if time.monotonic() > deadline:
                raise TimeoutError(f"shard group id did not appear at {path}")
            time.sleep(0.01)
        with open(path, "rb") as f:
            return cls(f.read(), n_ranks, rank, device)

    # -- lifecycle -------------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            try:
                lib().mse_shard_group_destroy(self._h)
            except Exception:  # interpreter shutdown
                pass
            self._h = C.c_void_p()

    __del__ = close

    def info(self) -> dict:
        out = (C.c_int32 * 4)()
        check(lib().mse_shard_group_info(self._h, out), "mse_shard_group_info")
        v = int(out[3])
        return {"n_ranks": int(out[0]), "rank": int(out[1]), "device": int(out[2]),
                "nccl_version": f"{v // 10000}.{v // 100 % 100}.{v % 100}" if v else None, "all_gathers": int(lib().mse_shard_group_gathers(self._h))}

    # -- searches (device pointers as ints, stream as int) -----------------------------------------------------------
    def flat_search_dev(self, shard, q_ptr: int, nq: int, k: int, ids_ptr: int, scores_ptr: int, stream: int = 0):
        check(lib().mse_search_flat_sharded_dev(self._h, shard._h, C.c_void_p(q_ptr), nq, k, C.c_void_p(ids_ptr), C.c_void_p(scores_ptr),
                                                C.c_void_p(stream)), "mse_search_flat_sharded_dev")

    def check(self) -> int:
        """Collective: finalises the last flat search (synchronises); returns the number of queries this rank re-ran."""
        rep = C.c_uint32()
        check(lib().mse_search_sharded_check(self._h, C.byref(rep)), "mse_search_sharded_check")
        return int(rep.value)

    def graph_search_dev(self, shard, q16_ptr: int, nq: int, L: int, start: int, k: int, ids_ptr: int, scores_ptr: int, distances_ptr: int = 0,
                         stream: int = 0):
        check(lib().mse_search_graph_sharded_dev(self._h, shard._h, C.c_void_p(q16_ptr), nq, L, start, k, C.c_void_p(ids_ptr), C.c_void_p(scores_ptr),
                                                 C.c_void_p(distances_ptr or None), C.c_void_p(stream)), "mse_search_graph_sharded_dev")

    def beam_search_dev(self, shard, q16_ptr: int, nq: int, L: int, W: int, start: int, k: int, ids_ptr: int, scores_ptr: int, cmps_ptr: int = 0,
                        pq_cmps_ptr: int = 0, stream: int = 0, d_qtm: int = 0, rabitq=None, d_luts: int = 0, n_centroids: int = 0, d_desc_scales: int = 0):
        od, nd = (rabitq.output_dims, rabitq.n_dims) if rabitq is not None else (0, 0)
        check(lib().mse_search_beam_sharded_dev(self._h, shard._h, C.c_void_p(q16_ptr), C.c_void_p(d_luts or None), C.c_void_p(d_qtm or None), od, nd,
                                                C.c_void_p(d_desc_scales or None), nq, L, W, start, n_centroids, k, C.c_void_p(ids_ptr),
                                                C.c_void_p(scores_ptr), C.c_void_p(cmps_ptr or None), C.c_void_p(pq_cmps_ptr or None), C.c_void_p(stream)),
              "mse_search_beam_sharded_dev")
