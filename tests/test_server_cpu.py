"""CPU tests of the clip_server boundary (routes, msgpack shapes, error behaviour, metrics names) with a test double in place
of the GPU towers, and of the query-assembly helpers."""
import asyncio
import base64
import io

import msgpack
import numpy as np
import pytest
from prometheus_client import CollectorRegistry


class FakeEncoder:
    """Stands in for mse_b200.Encoder in CPU tests only: deterministic unit vectors, no model."""
    image_size, dim, ctx = 384, 1152, 64

    def __init__(self, delay=0.0):
        self.calls, self.delay = [], delay

    def encode_image(self, images):
        import time
        time.sleep(self.delay)
        self.calls.append(("image", images.shape[0]))
        assert images.dtype == np.uint8 and images.shape[1:] == (384, 384, 3)
        f = np.stack([np.full(1152, float(im.mean()) + 1.0, np.float32) for im in images])
        return (f / np.linalg.norm(f, axis=1, keepdims=True)).astype(np.float16)

    def encode_text(self, ids):
        import time
        time.sleep(self.delay)
        self.calls.append(("text", ids.shape[0]))
        assert ids.shape[1] == 64
        f = np.stack([np.arange(1152, dtype=np.float32) + float(t.sum()) for t in ids])
        return (f / np.linalg.norm(f, axis=1, keepdims=True)).astype(np.float16)


class FakeTokenizer:
    def __call__(self, texts):
        out = np.ones((len(texts), 64), np.int32)
        for i, t in enumerate(texts):
            out[i, : len(t)] = [ord(c) % 100 + 2 for c in t][:64]
        return out


def _bmp(seed):
    from PIL import Image
    rng = np.random.default_rng(seed)
    buf = io.BytesIO()
    Image.fromarray(rng.integers(0, 256, (384, 384, 3), dtype=np.uint8)).save(buf, format="BMP")
    return buf.getvalue()


@pytest.fixture
def server(mse):
    from mse_b200.clip_server import ClipServer
    cfg = {"device": "cuda:0", "model": "ViT-SO400M-14-SigLIP-384", "model_name": "siglip-so400m-14-384", "max_batch_size": 4, "port": 0}
    return ClipServer(cfg, encoder=FakeEncoder(), tokenizer=FakeTokenizer(), registry=CollectorRegistry())


def _run(server, coro_fn):
    from aiohttp.test_utils import TestClient, TestServer

    async def go():
        async with TestClient(TestServer(server.app)) as client:
            return await coro_fn(client)
    return asyncio.new_event_loop().run_until_complete(go())


def test_routes_and_wire_format(server):
    async def scenario(c):
        r = await c.get("/")
        assert r.status == 204
        r = await c.get("/config")
        cfg = msgpack.loads(await r.read())
        assert cfg == {"model": "ViT-SO400M-14-SigLIP-384", "batch": 4, "image_size": [384, 384], "embedding_size": 1152}
        r = await c.post("/", data=msgpack.dumps({"images": [_bmp(1), _bmp(2)]}))
        assert r.status == 200 and r.content_type == "application/msgpack"
        out = msgpack.loads(await r.read())
        assert len(out) == 2 and all(isinstance(x, bytes) and len(x) == 2304 for x in out)
        v = np.frombuffer(out[0], "<f2").astype(np.float32)
        assert abs(np.linalg.norm(v) - 1) < 2e-3
        r = await c.post("/", data=msgpack.dumps({"text": ["a meme", "another"], "images": [_bmp(3)]}))  # text wins
        out = msgpack.loads(await r.read())
        assert r.status == 200 and len(out) == 2
        r = await c.post("/", data=msgpack.dumps({"text": "bare string"}))  # src/get_embedding.py:21 sends a bare str
        assert r.status == 200 and len(msgpack.loads(await r.read())) == 1
        r = await c.get("/metrics")
        m = (await r.read()).decode()
        assert 'modelserver_total_items_total{modality="image",model="siglip-so400m-14-384"} 2.0' in m
        assert "modelserver_inftime" in m and "modelserver_batchcount_total" in m
    _run(server, scenario)


def test_error_behaviour(server):
    async def scenario(c):
        r = await c.post("/", data=msgpack.dumps({"images": [_bmp(i) for i in range(5)]}))
        assert r.status == 500 and msgpack.loads(await r.read()) == "max batch size is 4"       # clip_server.py:139
        r = await c.post("/", data=msgpack.dumps({"text": ["x"] * 5}))
        assert r.status == 500 and msgpack.loads(await r.read()) == "max batch size is 4"       # :136
        r = await c.post("/", data=msgpack.dumps({"images": []}))
        assert r.status == 500 and msgpack.loads(await r.read()) == "images or text required"   # :142
        r = await c.post("/", data=msgpack.dumps({"images": [b"not an image"]}))
        assert r.status == 500
    _run(server, scenario)


def test_concurrent_requests_share_tower_calls(mse):
    """Cross-request batching: rows of concurrent requests are packed into tower calls of <= max_batch_size rows, results go back
    to the right request in the right order, and a request that does not fit opens the next call."""
    from mse_b200.clip_server import ClipServer
    enc = FakeEncoder(delay=0.05)
    cfg = {"device": "cuda:0", "model": "m", "model_name": "m", "max_batch_size": 8, "port": 0, "batch_window_ms": 30, "queue_depth": 64}
    srv = ClipServer(cfg, encoder=enc, tokenizer=FakeTokenizer(), registry=CollectorRegistry())
    texts = [[f"query {i} {j}" for j in range(1 + i % 3)] for i in range(12)]       # 12 requests of 1-3 rows: 24 rows

    async def scenario(c):
        async def one(t):
            r = await c.post("/", data=msgpack.dumps({"text": t}))
            assert r.status == 200
            return msgpack.loads(await r.read())
        outs = await asyncio.gather(*[one(t) for t in texts])
        single = FakeEncoder()
        for t, out in zip(texts, outs):
            want = single.encode_text(FakeTokenizer()(t))
            assert len(out) == len(t)
            for row, w in zip(out, want):
                assert row == w.tobytes()
        r = await c.get("/metrics")
        assert "modelserver_requests_per_batch" in (await r.read()).decode()
    _run(srv, scenario)
    sizes = [n for kind, n in enc.calls]
    assert sum(sizes) == 24 and max(sizes) <= 8
    assert len(sizes) < 12, f"requests were not coalesced: {sizes}"


def test_mixed_modalities_are_not_mixed_in_a_call(mse):
    from mse_b200.clip_server import ClipServer
    enc = FakeEncoder(delay=0.02)
    cfg = {"device": "cuda:0", "model": "m", "model_name": "m", "max_batch_size": 8, "port": 0, "batch_window_ms": 20, "queue_depth": 64}
    srv = ClipServer(cfg, encoder=enc, tokenizer=FakeTokenizer(), registry=CollectorRegistry())

    async def scenario(c):
        reqs = [{"text": ["a"]}, {"images": [_bmp(1)]}, {"text": ["b", "c"]}, {"images": [_bmp(2), _bmp(3)]}]
        rs = await asyncio.gather(*[c.post("/", data=msgpack.dumps(r)) for r in reqs])
        assert [r.status for r in rs] == [200] * 4
        assert [len(msgpack.loads(await r.read())) for r in rs] == [1, 1, 2, 2]
    _run(srv, scenario)
    assert sum(n for k, n in enc.calls if k == "text") == 3 and sum(n for k, n in enc.calls if k == "image") == 3


def test_tower_failure_reaches_every_request_of_the_call(mse):
    from mse_b200.clip_server import ClipServer

    class Broken(FakeEncoder):
        def encode_text(self, ids):
            raise RuntimeError("mse_encode_text_ids failed (-2): no device")
    srv = ClipServer({"device": "cuda:0", "model": "m", "model_name": "m", "max_batch_size": 8, "port": 0}, encoder=Broken(),
                     tokenizer=FakeTokenizer(), registry=CollectorRegistry())

    async def scenario(c):
        rs = await asyncio.gather(*[c.post("/", data=msgpack.dumps({"text": ["x"]})) for _ in range(3)])
        for r in rs:
            assert r.status == 500 and "no device" in msgpack.loads(await r.read())
    _run(srv, scenario)


def test_canonicalize_and_token_framing(mse, tmp_path):
    """SigLIP text preprocessing (clip_server.py:25,137; misc/clip_accursed.py:55: max_len 64, eos sticky, pad 1).  The c4_en
    SentencePiece model is not available offline, so the framing is exercised with a model trained here."""
    from mse_b200.clip_server import SiglipTokenizer
    canon = SiglipTokenizer.canonicalize
    assert canon("Foo_Bar") == "foo bar"                      # underscores become spaces before punctuation is dropped
    assert canon("  Hello,   WORLD!!  ") == "hello world"
    assert canon("it's a_meme: (2024)") == "its a meme 2024"
    assert canon("tabs\tand\nnewlines") == "tabs and newlines"
    assert canon("") == ""
    import sentencepiece as spm
    corpus = tmp_path / "corpus.txt"
    corpus.write_text("\n".join(f"the quick brown fox number {i} jumps over the lazy dog meme cat picture" for i in range(200)))
    spm.SentencePieceTrainer.train(input=str(corpus), model_prefix=str(tmp_path / "sp"), vocab_size=60, model_type="unigram",
                                   pad_id=0, eos_id=1, unk_id=2, bos_id=-1, minloglevel=2)
    tok = SiglipTokenizer(str(tmp_path / "sp.model"))
    ids = tok(["The quick_brown FOX!", "", "cat " * 200])
    assert ids.shape == (3, 64) and ids.dtype == np.int32
    n0 = len(tok.sp.encode("the quick brown fox"))
    assert list(ids[0, :n0]) == list(tok.sp.encode("the quick brown fox")) and ids[0, n0] == 1 and (ids[0, n0:] == 1).all()
    assert (ids[1] == 1).all()                                # empty text: EOS then padding, all id 1
    assert ids[2, 63] == 1 and (ids[2, :63] != 1).all()       # truncated to 63 pieces + sticky EOS
    assert (tok("bare string") == tok(["bare string"])).all()


def test_cpu_device_is_refused(mse):
    from mse_b200.clip_server import ClipServer
    with pytest.raises(RuntimeError) as e:
        ClipServer({"device": "cpu", "model": "m", "model_name": "m", "max_batch_size": 1, "port": 0, "model_path": "/nonexistent"},
                   registry=CollectorRegistry())
    assert "no CPU path" in str(e.value)


def test_get_total_embedding_and_select_shard(mse):
    from mse_b200.query import get_total_embedding, select_shard
    rng = np.random.default_rng(0)
    e_img = rng.standard_normal(1152).astype(np.float16)
    e_txt = rng.standard_normal(1152).astype(np.float16)
    raw = rng.standard_normal(1152).astype(np.float32)
    pre = {"nsfw": rng.standard_normal(1152).astype(np.float32)}

    def server(batch):
        return [e_img.tobytes()] * len(batch["images"]) if "images" in batch else [e_txt.tobytes()] * len(batch["text"])

    terms = [{"image": base64.b64encode(b"img").decode(), "weight": 0.5}, {"text": "cat"}, {"embedding": raw.tolist(), "weight": -2.0},
             {"predefined_embedding": "nsfw", "weight": 0.25}, {"predefined_embedding": "missing"}]
    got = get_total_embedding(terms, 1152, server, predefined_embeddings=pre)
    want = raw * np.float32(-2.0) + pre["nsfw"] * np.float32(0.25) + e_img.astype(np.float32) * np.float32(0.5) + e_txt.astype(np.float32)
    assert np.allclose(got, want, atol=1e-6)
    assert abs(np.linalg.norm(got) - 1) > 0.1  # not renormalised (common.rs:215-274)
    cents = [(rng.standard_normal(1152).astype(np.float32), 10 * i) for i in range(5)]
    q = cents[3][0] * 2 + 0.01
    assert select_shard(cents, q) == 3
    assert select_shard([(cents[0][0], 1), (cents[0][0], 2)], q) == 1  # ties keep the last maximum


def test_client_bmps_take_the_device_path(mse):
    """size x size 24-bit BMPs (what src/common.rs:42-53 sends) reach the encoder as files; anything else is decoded on the host"""
    from mse_b200.clip_server import ClipServer, is_client_bmp
    from PIL import Image

    class BmpEncoder(FakeEncoder):
        def encode_image_bmp(self, files):
            self.calls.append(("bmp", len(files)))
            assert all(is_client_bmp(bytes(f), 384) for f in files)
            return np.tile((np.ones(1152, np.float32) / np.sqrt(1152)).astype(np.float16), (len(files), 1))
    enc = BmpEncoder()
    srv = ClipServer({"device": "cuda:0", "model": "m", "model_name": "m", "max_batch_size": 4, "port": 0}, encoder=enc, tokenizer=FakeTokenizer(),
                     registry=CollectorRegistry())
    png = io.BytesIO()
    Image.fromarray(np.zeros((384, 384, 3), np.uint8)).save(png, format="PNG")
    small = io.BytesIO()
    Image.fromarray(np.zeros((100, 100, 3), np.uint8)).save(small, format="BMP")
    assert is_client_bmp(_bmp(1), 384) and not is_client_bmp(png.getvalue(), 384) and not is_client_bmp(small.getvalue(), 384)

    async def scenario(c):
        r = await c.post("/", data=msgpack.dumps({"images": [_bmp(1), _bmp(2)]}))
        assert r.status == 200 and len(msgpack.loads(await r.read())) == 2
        r = await c.post("/", data=msgpack.dumps({"images": [_bmp(1), png.getvalue()]}))
        assert r.status == 200 and len(msgpack.loads(await r.read())) == 2
        m = (await (await c.get("/metrics")).read()).decode()
        assert 'modelserver_total_items_total{modality="image",model="m"} 4.0' in m
    _run(srv, scenario)
    assert ("bmp", 2) in enc.calls and ("image", 2) in enc.calls
