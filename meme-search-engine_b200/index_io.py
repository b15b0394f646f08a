"""Packed-index files the serving side loads at start-up (src/query_disk_index.rs:664-705, written by dump_processor.rs:560-569).

    index.msgpack                 IndexHeader, rmp_serde::to_vec_named (MAP form, field names as keys):
                                  shards [[centroid f32[d], medioid u32], ...], count, dead_count, record_pad_size,
                                  quantizer {centroids f32[256*d], transform f32[d*d], n_dims_per_code, n_dims}, descriptor_cdfs [[f32]]
                                  (common.rs:166-174; diskann/src/vector.rs:308-314)
    index.pq-codes.bin            count x (n_dims / n_dims_per_code) bytes, node-major (query_disk_index.rs:101-104,688)
    index.descriptor-codes.bin    count x len(descriptor_cdfs) bytes (query_disk_index.rs:695)

Not covered: index.bin (4096-byte records holding a u16-LE length and a `bitcode`-encoded PackedIndexEntry, dump_processor.rs:500-522):
the `bitcode` crate's wire format is not described anywhere in the reference tree, so vectors and adjacency lists reach the GPU
through mse_index_create / mse_index_set_graph (or shard_io.py) instead.

load_packed_index() attaches the codes / descriptors to a VectorList and returns the codec handle + header, i.e. everything
`Index` (query_disk_index.rs:650-662) holds except the node records.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import msgpack
import numpy as np


@dataclass
class IndexHeader:                      # common.rs:166-174
    shards: list                        # [(centroid f32[d], medioid int)]
    count: int
    dead_count: int
    record_pad_size: int
    quantizer: dict                     # {"centroids", "transform", "n_dims_per_code", "n_dims"}
    descriptor_cdfs: list               # [f32[...]] per descriptor

    @property
    def pq_code_size(self) -> int:      # query_disk_index.rs:677
        return int(self.quantizer["n_dims"]) // int(self.quantizer["n_dims_per_code"])

    @property
    def n_descriptors(self) -> int:     # query_disk_index.rs:678
        return len(self.descriptor_cdfs)


_HEADER_FIELDS = ["shards", "count", "dead_count", "record_pad_size", "quantizer", "descriptor_cdfs"]
_PQ_FIELDS = ["centroids", "transform", "n_dims_per_code", "n_dims"]


def _struct(obj, names):
    """rmp_serde reads a struct from a map (to_vec_named) or from an array (to_vec)."""
    if isinstance(obj, dict):
        return {k: obj[k] for k in names}
    if len(obj) != len(names):
        raise ValueError(f"expected {names}, got {len(obj)} fields")
    return dict(zip(names, obj))


def read_index_header(path: str) -> IndexHeader:
    with open(path, "rb") as f:
        h = _struct(msgpack.unpackb(f.read(), raw=False, strict_map_key=False), _HEADER_FIELDS)
    q = _struct(h["quantizer"], _PQ_FIELDS)
    q = {"centroids": np.asarray(q["centroids"], np.float32), "transform": np.asarray(q["transform"], np.float32),
         "n_dims_per_code": int(q["n_dims_per_code"]), "n_dims": int(q["n_dims"])}
    shards = [(np.asarray(c, np.float32), int(m)) for c, m in h["shards"]]
    return IndexHeader(shards, int(h["count"]), int(h["dead_count"]), int(h["record_pad_size"]), q,
                       [np.asarray(c, np.float32) for c in h["descriptor_cdfs"]])


def write_index_header(path: str, hdr: IndexHeader) -> None:
    """dump_processor.rs:560-568 (`rmp_serde::to_vec_named`): maps keyed by field name, fields in declaration order."""
    f32s = lambda a: [float(v) for v in np.asarray(a, np.float32).reshape(-1)]
    obj = {"shards": [[f32s(c), int(m)] for c, m in hdr.shards], "count": int(hdr.count), "dead_count": int(hdr.dead_count),
           "record_pad_size": int(hdr.record_pad_size),
           "quantizer": {"centroids": f32s(hdr.quantizer["centroids"]), "transform": f32s(hdr.quantizer["transform"]),
                         "n_dims_per_code": int(hdr.quantizer["n_dims_per_code"]), "n_dims": int(hdr.quantizer["n_dims"])},
           "descriptor_cdfs": [f32s(c) for c in hdr.descriptor_cdfs]}
    with open(path, "wb") as f:
        f.write(msgpack.packb(obj, use_single_float=True))


def read_codes(index_dir: str, hdr: IndexHeader):
    """-> (pq_codes u8 [count, code_size], descriptors u8 [count, n_descriptors] or None)"""
    pq = np.fromfile(os.path.join(index_dir, "index.pq-codes.bin"), np.uint8)
    if pq.size != hdr.count * hdr.pq_code_size:
        raise ValueError(f"index.pq-codes.bin holds {pq.size} bytes, header says {hdr.count} x {hdr.pq_code_size}")
    desc = None
    if hdr.n_descriptors:
        desc = np.fromfile(os.path.join(index_dir, "index.descriptor-codes.bin"), np.uint8)
        if desc.size != hdr.count * hdr.n_descriptors:
            raise ValueError(f"index.descriptor-codes.bin holds {desc.size} bytes, header says {hdr.count} x {hdr.n_descriptors}")
        desc = desc.reshape(hdr.count, hdr.n_descriptors)
    return pq.reshape(hdr.count, hdr.pq_code_size), desc


def load_packed_index(index_dir: str, vecs, has_url=None, device: int = 0):
    """Attach index.pq-codes.bin / index.descriptor-codes.bin to `vecs` (a diskann.VectorList that already holds the node vectors
    and adjacency lists) and build the ProductQuantizer from the header.  -> (IndexHeader, ProductQuantizer)"""
    from . import diskann as dk
    hdr = read_index_header(os.path.join(index_dir, "index.msgpack"))
    if len(vecs) != hdr.count:
        raise ValueError(f"the vector list holds {len(vecs)} rows, the index header says {hdr.count}")
    pq_codes, desc = read_codes(index_dir, hdr)
    vecs.set_pq_codes(pq_codes)
    vecs.set_descriptors(desc, has_url)
    q = hdr.quantizer
    d = q["n_dims"]
    pq = dk.ProductQuantizer(q["centroids"].reshape(-1, d), q["transform"].reshape(d, d), q["n_dims_per_code"], device)
    return hdr, pq
