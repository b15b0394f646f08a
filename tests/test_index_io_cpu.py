"""index.msgpack / code files of the packed index (common.rs:166-174, dump_processor.rs:560-568, query_disk_index.rs:664-705):
the header writer emits rmp_serde::to_vec_named's map form byte for byte (checked against bytes assembled by hand), and the
reader accepts the map and the array form."""
import struct

import msgpack
import numpy as np
import pytest

import mse_b200  # noqa: F401
from mse_b200 import index_io


def _str(s):
    b = s.encode()
    return (bytes([0xa0 | len(b)]) if len(b) < 32 else b"\xd9" + bytes([len(b)])) + b


def _arr(n):
    return bytes([0x90 | n]) if n < 16 else (b"\xdc" + struct.pack(">H", n) if n < 65536 else b"\xdd" + struct.pack(">I", n))


def _uint(v):
    return bytes([v]) if v < 128 else (b"\xcc" + bytes([v]) if v < 256 else (b"\xcd" + struct.pack(">H", v) if v < 65536 else b"\xce" + struct.pack(">I", v)))


def _f32s(a):
    return _arr(len(a)) + b"".join(b"\xca" + struct.pack(">f", float(v)) for v in a)


def test_header_bytes_and_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    d, dpc = 16, 4
    hdr = index_io.IndexHeader(shards=[(rng.standard_normal(d).astype(np.float32), 7), (rng.standard_normal(d).astype(np.float32), 70000)],
                               count=1234, dead_count=5, record_pad_size=4096,
                               quantizer={"centroids": rng.standard_normal(3 * d).astype(np.float32), "transform": np.eye(d, dtype=np.float32).reshape(-1),
                                          "n_dims_per_code": dpc, "n_dims": d},
                               descriptor_cdfs=[np.linspace(0, 1, 5).astype(np.float32), np.linspace(0, 1, 3).astype(np.float32)])
    p = str(tmp_path / "index.msgpack")
    index_io.write_index_header(p, hdr)
    want = (bytes([0x86])
            + _str("shards") + _arr(2) + b"".join(_arr(2) + _f32s(c) + _uint(m) for c, m in hdr.shards)
            + _str("count") + _uint(1234) + _str("dead_count") + _uint(5) + _str("record_pad_size") + _uint(4096)
            + _str("quantizer") + bytes([0x84]) + _str("centroids") + _f32s(hdr.quantizer["centroids"]) + _str("transform") + _f32s(hdr.quantizer["transform"])
            + _str("n_dims_per_code") + _uint(dpc) + _str("n_dims") + _uint(d)
            + _str("descriptor_cdfs") + _arr(2) + b"".join(_f32s(c) for c in hdr.descriptor_cdfs))
    assert open(p, "rb").read() == want
    h2 = index_io.read_index_header(p)
    assert (h2.count, h2.dead_count, h2.record_pad_size, h2.pq_code_size, h2.n_descriptors) == (1234, 5, 4096, d // dpc, 2)
    assert h2.shards[1][1] == 70000 and np.array_equal(h2.shards[0][0], hdr.shards[0][0])
    assert np.array_equal(h2.quantizer["centroids"], hdr.quantizer["centroids"]) and np.array_equal(h2.descriptor_cdfs[1], hdr.descriptor_cdfs[1])
    # compact (array) form of the same structs
    arr = [[[c.tolist(), m] for c, m in hdr.shards], 1234, 5, 4096,
           [hdr.quantizer["centroids"].tolist(), hdr.quantizer["transform"].tolist(), dpc, d], [c.tolist() for c in hdr.descriptor_cdfs]]
    open(p, "wb").write(msgpack.packb(arr, use_single_float=True))
    h3 = index_io.read_index_header(p)
    assert h3.count == 1234 and np.array_equal(h3.quantizer["transform"], hdr.quantizer["transform"])


def test_code_files(tmp_path):
    rng = np.random.default_rng(1)
    hdr = index_io.IndexHeader([], 50, 0, 4096, {"centroids": np.zeros(1, np.float32), "transform": np.zeros(1, np.float32), "n_dims_per_code": 18, "n_dims": 1152},
                               [np.zeros(2, np.float32)] * 4)
    pq = rng.integers(0, 256, (50, 64)).astype(np.uint8)
    desc = rng.integers(0, 256, (50, 4)).astype(np.uint8)
    pq.tofile(str(tmp_path / "index.pq-codes.bin"))
    desc.tofile(str(tmp_path / "index.descriptor-codes.bin"))
    a, b = index_io.read_codes(str(tmp_path), hdr)
    assert np.array_equal(a, pq) and np.array_equal(b, desc)
    pq[:49].tofile(str(tmp_path / "index.pq-codes.bin"))
    with pytest.raises(ValueError):
        index_io.read_codes(str(tmp_path), hdr)
