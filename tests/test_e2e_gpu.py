"""BASELINE configs[4] at test scale, on one GPU: encode synthetic images with the towers, index the embeddings, build the
Vamana graph on the GPU, serve mixed text / image / weighted queries assembled as src/common.rs:215-274 does, and check every
stage against the oracle (towers: cosine >= 1 - 1e-3; graph search over the embeddings: bit-exact; flat: bit-exact)."""
import base64

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_encode_build_serve(tmp_path_factory, mse, oracle):
    from oracle import towers as T
    from mse_b200 import diskann as dk
    from mse_b200.query import get_total_embedding
    v, t = T.build_vision(depth=2, seed=42), T.build_text(depth=2, seed=43)
    sd = T.export_openclip(v, t)
    path = str(tmp_path_factory.mktemp("w") / "towers2.msew")
    mse.weights.save_weights(path, sd, mse.weights.config_for(sd))
    enc = mse.Encoder(path, max_batch=16)
    # 1. encode: 96 synthetic images in batches of 16 (the ingest loop of src/main.rs:680-694)
    imgs = T.synthetic_images(11, 96)
    emb = np.concatenate([enc.encode_image(imgs[i:i + 16]) for i in range(0, 96, 16)])
    ref = T.encode_image(v, imgs[:8])
    cos = (emb[:8].astype(np.float64) * ref).sum(1) / (np.linalg.norm(emb[:8].astype(np.float64), axis=1) * np.linalg.norm(ref, axis=1))
    assert cos.min() >= 1 - 1e-3
    # 2. index + graph (R small: 96 nodes)
    vl = dk.VectorList.from_f16s(emb)
    cfg = dk.IndexBuildConfig(r=8, l=24, maxc=64)
    dk.random_fill_graph(vl, 8, seed=1)
    med = dk.medioid(vl)
    assert med == oracle.medioid(emb)
    dk.build_graph(vl, med, cfg, seed=2)
    adj, deg = vl.get_graph()
    # 3. queries through the clip_server-shaped boundary: image term, text term, weighted sum of both (not renormalised)
    ids = T.synthetic_token_ids(5, 2)

    def query_server(batch):
        if "images" in batch:
            arr = np.stack([np.frombuffer(b, np.uint8).reshape(384, 384, 3) for b in batch["images"]])
            return [r.tobytes() for r in enc.encode_image(arr)]
        return [r.tobytes() for r in enc.encode_text(np.stack([ids[int(s)] for s in batch["text"]]))]

    img_b64 = base64.standard_b64encode(imgs[3].tobytes()).decode()
    queries = [get_total_embedding([{"image": img_b64}], 1152, query_server),
               get_total_embedding([{"text": "0"}], 1152, query_server),
               get_total_embedding([{"image": img_b64, "weight": 0.7}, {"text": "1", "weight": -0.3}], 1152, query_server)]
    q = np.stack(queries).astype(np.float32)
    assert abs(np.linalg.norm(q[2]) - 1.0) > 1e-3                      # weighted sums are not renormalised (common.rs:215-274)
    # 4a. flat (src/main.rs path): ids and scores equal the oracle's; the image query finds its own row first
    sc, lab = vl.search(q, 10)
    oi, os_ = oracle.flat_search(q, emb, 10)
    assert np.array_equal(lab, oi.astype(np.int64)) and np.array_equal(sc, os_)
    assert 3 in lab[0, :3]
    # 4b. graph (diskann path): bit-exact against the oracle's greedy_search over the GPU-built graph
    q16 = q.astype(np.float16)
    res = dk.greedy_search(vl, q16, med, cfg)
    og = oracle.IndexGraph(96, adj.shape[1])
    og.set(adj, deg)
    ocfg = oracle.make_config(r=8, l=24, maxc=64)
    s = oracle.Scratch(96, ocfg)
    for i in range(3):
        d = oracle.greedy_search(s, med, False, q16[i], emb, og, ocfg)
        m = int(res.len[i])
        assert np.array_equal(res.ids[i, :m], s.neighbour_ids) and np.array_equal(res.scores[i, :m], s.neighbour_scores) and int(res.distances[i]) == d
    assert 3 in res.ids[0, :3]
    vl.close()
    enc.close()


def test_http_embed_service_on_gpu(tmp_path_factory, mse):
    """clip_server boundary end to end on the GPU: concurrent HTTP/msgpack requests -> coalescer -> the real towers behind the C ABI ->
    fp16 byte strings.  Every returned vector matches the fp32 oracle (cosine >= 1 - 1e-3) and is unit norm; concurrent requests
    share tower calls."""
    import asyncio
    import io

    import msgpack
    from aiohttp.test_utils import TestClient, TestServer
    from PIL import Image
    from prometheus_client import CollectorRegistry

    from mse_b200.clip_server import ClipServer
    from oracle import towers as T
    v, t = T.build_vision(depth=2, seed=42), T.build_text(depth=2, seed=43)
    sd = T.export_openclip(v, t)
    path = str(tmp_path_factory.mktemp("w") / "towers2.msew")
    mse.weights.save_weights(path, sd, mse.weights.config_for(sd))
    token_rows = T.synthetic_token_ids(5, 6)

    class IdTokenizer:                      # no SentencePiece model offline: "text" i stands for the i-th synthetic token row
        def __call__(self, texts):
            return np.stack([token_rows[int(s)] for s in texts]).astype(np.int32)

    cfg = {"device": "cuda:0", "model": "ViT-SO400M-14-SigLIP-384", "model_name": "siglip-test", "max_batch_size": 8, "port": 0,
           "model_path": path, "batch_window_ms": 30}
    srv = ClipServer(cfg, tokenizer=IdTokenizer(), registry=CollectorRegistry())
    imgs = T.synthetic_images(21, 6)

    def bmp(a):
        buf = io.BytesIO()
        Image.fromarray(a).save(buf, format="BMP")
        return buf.getvalue()

    async def go():
        async with TestClient(TestServer(srv.app)) as c:
            r = await c.get("/config")
            assert msgpack.loads(await r.read()) == {"model": "ViT-SO400M-14-SigLIP-384", "batch": 8, "image_size": [384, 384], "embedding_size": 1152}

            async def post(body):
                r = await c.post("/", data=msgpack.dumps(body))
                assert r.status == 200, await r.read()
                return [np.frombuffer(b, "<f2").astype(np.float32) for b in msgpack.loads(await r.read())]
            reqs = [{"images": [bmp(imgs[0]), bmp(imgs[1])]}, {"images": [bmp(imgs[2])]}, {"text": ["0", "1"]}, {"images": [bmp(imgs[3]), bmp(imgs[4]), bmp(imgs[5])]},
                    {"text": ["2"]}, {"text": ["3", "4", "5"]}]
            return await asyncio.gather(*[post(r) for r in reqs])
    outs = asyncio.new_event_loop().run_until_complete(go())
    got_img = np.stack(outs[0] + outs[1] + outs[3])
    got_txt = np.stack(outs[2] + outs[4] + outs[5])
    ref_img, ref_txt = T.encode_image(v, imgs), T.encode_text(t, token_rows)

    def cos(a, b):
        return ((a.astype(np.float64) * b).sum(1) / (np.linalg.norm(a.astype(np.float64), axis=1) * np.linalg.norm(b, axis=1))).min()
    assert cos(got_img, ref_img) >= 1 - 1e-3 and cos(got_txt, ref_txt) >= 1 - 1e-3
    assert np.allclose(np.linalg.norm(got_img, axis=1), 1.0, atol=2e-3) and np.allclose(np.linalg.norm(got_txt, axis=1), 1.0, atol=2e-3)
    batches = srv.registry.get_sample_value("modelserver_batchcount_total", {"model": "siglip-test"})
    assert batches is not None and batches < 6                       # six requests, fewer tower calls
    srv.encoder.close()
