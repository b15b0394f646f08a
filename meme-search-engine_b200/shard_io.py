"""Shard files either side of the graph build -- the reference's `generate-index-shard` binary (src/generate_index_shard.rs).

    input  `N.shard-input` stream   rmp_serde compact (array) encoding: ShardInputHeader [id, centroid f32[d]] followed by
                                    ShardedRecord [id, bin(2*d) fp16-LE] until EOF         (dump_processor.rs:206,447-451; common.rs:131-142)
    output `N.shard.bin`            the out-neighbour lists of the base nodes, u32-LE, concatenated (generate_index_shard.rs:143-153)
           `N.shard-header.msgpack` ShardHeader [id, max, centroid, medioid, offsets u64 (BYTE offsets, n+1 entries), mapping]
                                                                                           (generate_index_shard.rs:155-164; common.rs:144-152)

`generate_index_shard` is the body of that binary's main() on the GPU: read shard -> (append query vectors) ->
random_fill_graph -> medioid -> build_graph (-> second pass) (-> robust_stitch) -> write shard.  rmp_serde accepts structs as
arrays or as maps, so the readers here accept both; the writers emit the compact array form the reference emits, byte for byte
(smallest integer encodings, f32 floats, bin for serde_bytes).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import msgpack
import numpy as np

D_EMB = 1152  # generate_index_shard.rs:40


@dataclass
class ShardInput:
    id: int
    centroid: np.ndarray      # f32 [d]
    original_ids: np.ndarray  # u32 [n]
    vectors: np.ndarray       # f16 [n, d]


@dataclass
class ShardHeader:            # common.rs:144-152
    id: int
    max: int
    centroid: np.ndarray
    medioid: int
    offsets: np.ndarray       # u64 [n + 1], byte offsets into N.shard.bin
    mapping: np.ndarray       # u32 [n] original ids


def _fields(obj, names):
    if isinstance(obj, dict):
        return [obj[k] if k in obj else obj[k.encode()] for k in names]
    if len(obj) != len(names):
        raise ValueError(f"expected {len(names)} fields {names}, got {len(obj)}")
    return list(obj)


def write_shard_input(path: str, shard_id: int, centroid, original_ids, vectors16) -> None:
    """dump_processor.rs:206 + :447-451 (what feeds generate-index-shard)."""
    v = np.ascontiguousarray(vectors16).view(np.uint16)
    with open(path, "wb") as f:
        f.write(msgpack.packb([int(shard_id), [float(c) for c in np.asarray(centroid, np.float32)]], use_single_float=True))
        for i, oid in enumerate(np.asarray(original_ids)):
            f.write(msgpack.packb([int(oid), v[i].astype("<u2").tobytes()], use_bin_type=True))


def read_shard_input(path: str, d: int = D_EMB) -> ShardInput:
    """generate_index_shard.rs:50-70: header, then records until EOF."""
    with open(path, "rb") as f:
        up = msgpack.Unpacker(f, raw=False, max_buffer_size=1 << 31)
        sid, centroid = _fields(next(up), ["id", "centroid"])
        ids, rows = [], []
        for rec in up:
            rid, vec = _fields(rec, ["id", "vector"])
            if len(vec) != 2 * d:
                raise ValueError(f"record {rid}: vector has {len(vec)} bytes, expected {2 * d}")
            ids.append(rid)
            rows.append(vec)
    vectors = np.frombuffer(b"".join(rows), dtype="<f2").reshape(-1, d) if rows else np.empty((0, d), np.float16)
    return ShardInput(int(sid), np.asarray(centroid, np.float32), np.asarray(ids, np.uint32), vectors)


def write_shard(out_dir: str, shard_id: int, centroid, medioid: int, adj: np.ndarray, deg: np.ndarray, original_ids) -> tuple[str, str]:
    """generate_index_shard.rs:139-164.  adj/deg cover base nodes first; only the first len(original_ids) lists are written."""
    n = len(original_ids)
    deg = np.asarray(deg[:n], np.int64)
    adj = np.asarray(adj[:n], "<u4")
    mask = np.arange(adj.shape[1])[None, :] < deg[:, None]
    flat = adj[mask]                                        # row-major: list 0, list 1, ...
    offsets = np.zeros(n + 1, np.uint64)
    offsets[1:] = np.cumsum(deg) * 4
    bin_path = os.path.join(out_dir, f"{shard_id}.shard.bin")
    hdr_path = os.path.join(out_dir, f"{shard_id}.shard-header.msgpack")
    with open(bin_path, "wb") as f:
        f.write(flat.tobytes())
    with open(hdr_path, "wb") as f:
        f.write(msgpack.packb([int(shard_id), int(np.max(original_ids)), [float(c) for c in np.asarray(centroid, np.float32)], int(medioid),
                               [int(o) for o in offsets], [int(i) for i in original_ids]], use_single_float=True))
    return bin_path, hdr_path


def read_shard(out_dir: str, shard_id: int):
    """-> (ShardHeader, adj [n, stride] u32, deg [n]) as dump_processor.rs:239-299 consumes them."""
    with open(os.path.join(out_dir, f"{shard_id}.shard-header.msgpack"), "rb") as f:
        sid, mx, centroid, medioid, offsets, mapping = _fields(msgpack.unpackb(f.read(), raw=False, strict_map_key=False),
                                                               ["id", "max", "centroid", "medioid", "offsets", "mapping"])
    hdr = ShardHeader(int(sid), int(mx), np.asarray(centroid, np.float32), int(medioid), np.asarray(offsets, np.uint64), np.asarray(mapping, np.uint32))
    data = np.fromfile(os.path.join(out_dir, f"{shard_id}.shard.bin"), dtype="<u4")
    off = (hdr.offsets // 4).astype(np.int64)
    deg = np.diff(off).astype(np.uint32)
    stride = int(deg.max()) if deg.size else 1
    adj = np.zeros((deg.size, max(stride, 1)), np.uint32)
    for i in range(deg.size):                               # test / tooling path, not a hot loop
        adj[i, : deg[i]] = data[off[i]: off[i + 1]]
    return hdr, adj, deg


def generate_index_shard(input_file: str, out_dir: str, queries_bin: str | None = None, l: int = 192, r: int = 64, maxc: int = 750,
                         alpha: int = 65536, query_alpha: int = 65536, alpha_2: int = 65536, second_pass: bool = False, seed: int = 0,
                         device: int = 0) -> dict:
    """main() of src/generate_index_shard.rs with the graph work on the GPU (CLI flags keep their names: -L -R -C -A -Q -B -s)."""
    from . import diskann as dk
    si = read_shard_input(input_file)
    vectors = si.vectors
    query_breakpoint = len(si.original_ids)                               # :72
    if queries_bin:
        q = np.fromfile(queries_bin, dtype="<f2")
        vectors = np.concatenate([vectors, q[: q.size // D_EMB * D_EMB].reshape(-1, D_EMB)])   # :74-84
    cfg = dk.IndexBuildConfig(r=r, l=l, maxc=maxc, alpha=alpha, query_alpha=query_alpha, saturate_graph=False,
                              query_breakpoint=query_breakpoint, max_add_per_stitch_iter=16)      # :85-94
    vl = dk.VectorList.from_f16s(np.ascontiguousarray(vectors), device)
    dk.random_fill_graph(vl, r, seed)                                     # :102-105
    med = dk.medioid(vl)                                                  # :109
    stats = dk.build_graph(vl, med, cfg, seed)                            # :111-114
    if second_pass:                                                       # :118-125
        cfg.alpha = alpha_2
        dk.build_graph(vl, med, cfg, seed + 1)
    if query_breakpoint < len(vl):                                        # :127-131
        dk.robust_stitch(vl, cfg, seed)
    adj, deg = vl.get_graph()
    vl.close()
    write_shard(out_dir, si.id, si.centroid, med, adj, deg, si.original_ids)
    return {"id": si.id, "vectors": int(query_breakpoint), "queries": int(len(vectors) - query_breakpoint), "medioid": med, "build": stats,
            "mean_degree": float(deg[:query_breakpoint].mean()) if query_breakpoint else 0.0}


def main(argv=None) -> int:
    """CLI with the reference binary's arguments (generate_index_shard.rs:13-37): input_file out_dir [queries_bin] -L -R -C -A -Q -B -s."""
    import argparse
    import json
    ap = argparse.ArgumentParser(description="Generate indices from shard files (GPU)")
    ap.add_argument("input_file")
    ap.add_argument("out_dir")
    ap.add_argument("queries_bin", nargs="?")
    ap.add_argument("-L", dest="l", type=int, default=192, help="search list size (higher is better but slower)")
    ap.add_argument("-R", dest="r", type=int, default=64, help="graph degree")
    ap.add_argument("-C", dest="maxc", type=int, default=750, help="max candidate list size")
    ap.add_argument("-A", dest="alpha", type=int, default=65536, help="first pass relaxation factor (times 2^16)")
    ap.add_argument("-Q", dest="query_alpha", type=int, default=65536, help="query set special relaxation factor (times 2^16)")
    ap.add_argument("-B", dest="alpha_2", type=int, default=65536, help="second pass relaxation factor (times 2^16)")
    ap.add_argument("-s", dest="second_pass", action="store_true", help="do second pass")
    ap.add_argument("-N", dest="n", type=int, default=None, help="number of vectors to allocate for (accepted for compatibility; unused)")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args(argv)
    info = generate_index_shard(a.input_file, a.out_dir, a.queries_bin, l=a.l, r=a.r, maxc=a.maxc, alpha=a.alpha, query_alpha=a.query_alpha,
                                alpha_2=a.alpha_2, second_pass=a.second_pass, seed=a.seed, device=a.device)
    print(json.dumps(info))
    print(f"{info['vectors']} vectors")                                     # generate_index_shard.rs:167
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
