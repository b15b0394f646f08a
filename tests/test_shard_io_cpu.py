"""Shard file formats of generate-index-shard (src/generate_index_shard.rs:50-70,139-164; common.rs:131-152): the writers emit
rmp_serde's compact array encoding byte for byte (checked against bytes assembled by hand from the MessagePack spec), and
the readers accept both the array and the map (to_vec_named) form like rmp_serde's Deserialize does."""
import struct

import msgpack
import numpy as np

import mse_b200  # noqa: F401
from mse_b200 import shard_io
from helpers import index_f16


def _mp_array_hdr(n):
    if n < 16:
        return bytes([0x90 | n])
    if n < 65536:
        return b"\xdc" + struct.pack(">H", n)
    return b"\xdd" + struct.pack(">I", n)


def _mp_uint(v):
    if v < 128:
        return bytes([v])
    if v < 256:
        return b"\xcc" + bytes([v])
    if v < 65536:
        return b"\xcd" + struct.pack(">H", v)
    if v < 1 << 32:
        return b"\xce" + struct.pack(">I", v)
    return b"\xcf" + struct.pack(">Q", v)


def _mp_f32_array(a):
    return _mp_array_hdr(len(a)) + b"".join(b"\xca" + struct.pack(">f", float(v)) for v in a)


def _mp_bin(b):
    n = len(b)
    return (b"\xc4" + bytes([n]) if n < 256 else b"\xc5" + struct.pack(">H", n) if n < 65536 else b"\xc6" + struct.pack(">I", n)) + b


def test_shard_input_bytes_and_round_trip(tmp_path):
    d = shard_io.D_EMB
    x = index_f16(3, 5)
    ids = np.array([7, 300, 70000, 5, 4_000_000_000], np.uint32)
    cent = np.linspace(-1, 1, d).astype(np.float32)
    p = str(tmp_path / "3.shard-input")
    shard_io.write_shard_input(p, 3, cent, ids, x)
    want = _mp_array_hdr(2) + _mp_uint(3) + _mp_f32_array(cent)
    for i in range(5):
        want += _mp_array_hdr(2) + _mp_uint(int(ids[i])) + _mp_bin(x[i].view(np.uint16).astype("<u2").tobytes())
    assert open(p, "rb").read() == want
    si = shard_io.read_shard_input(p)
    assert si.id == 3 and np.array_equal(si.centroid, cent) and np.array_equal(si.original_ids, ids)
    assert np.array_equal(si.vectors.view(np.uint16), x.view(np.uint16))
    # the named-map form (rmp_serde accepts it for the same struct)
    with open(p, "wb") as f:
        f.write(msgpack.packb({"id": 9, "centroid": cent.tolist()}, use_single_float=True))
        f.write(msgpack.packb({"id": 11, "vector": x[0].tobytes()}, use_bin_type=True))
    si = shard_io.read_shard_input(p)
    assert si.id == 9 and si.original_ids.tolist() == [11] and np.array_equal(si.vectors[0].view(np.uint16), x[0].view(np.uint16))


def test_shard_output_bytes_and_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    n_base, n_all, R = 6, 9, 4                                  # 3 trailing query nodes are not written (:146)
    adj = rng.integers(0, n_all, (n_all, R)).astype(np.uint32)
    deg = np.array([4, 0, 2, 4, 1, 3, 4, 4, 4], np.uint32)
    mapping = np.array([10, 11, 500, 13, 70000, 15], np.uint32)
    cent = rng.standard_normal(8).astype(np.float32)
    bp, hp = shard_io.write_shard(str(tmp_path), 12, cent, 5, adj, deg, mapping)
    lists = [adj[i, : deg[i]] for i in range(n_base)]
    assert open(bp, "rb").read() == b"".join(l.astype("<u4").tobytes() for l in lists)
    offs = np.concatenate([[0], np.cumsum([4 * len(l) for l in lists])])
    want = (_mp_array_hdr(6) + _mp_uint(12) + _mp_uint(70000) + _mp_f32_array(cent) + _mp_uint(5)
            + _mp_array_hdr(n_base + 1) + b"".join(_mp_uint(int(o)) for o in offs)
            + _mp_array_hdr(n_base) + b"".join(_mp_uint(int(m)) for m in mapping))
    assert open(hp, "rb").read() == want
    hdr, a2, d2 = shard_io.read_shard(str(tmp_path), 12)
    assert (hdr.id, hdr.max, hdr.medioid) == (12, 70000, 5) and np.array_equal(hdr.mapping, mapping) and np.array_equal(hdr.centroid, cent)
    assert np.array_equal(d2, deg[:n_base])
    for i in range(n_base):
        assert np.array_equal(a2[i, : d2[i]], lists[i])
