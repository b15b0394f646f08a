"""The `query-disk-index` work-alike (meme-search-engine_b200/query_server.py): JSON boundary on the CPU with a stand-in batcher, and
on the GPU the whole path HTTP -> micro-batch -> mse_search_beam_dev -> mse_dedup_topk_dev -> JSON against the CPU oracle's beam
search + de-duplication + stable sort (src/query_disk_index.rs:420-551): ids and scores of every match identical."""
import asyncio
import json

import numpy as np
import pytest
from prometheus_client import CollectorRegistry

from helpers import clustered_f16


def _run(server, coro_fn):
    from aiohttp.test_utils import TestClient, TestServer

    async def go():
        async with TestClient(TestServer(server.app)) as client:
            return await coro_fn(client)
    return asyncio.new_event_loop().run_until_complete(go())


class StubBatcher:
    """CPU stand-in for SearchBatcher: returns the nodes 3, 1, 2 with fixed scores"""
    def __init__(self):
        self.seen = []

    def start(self):
        pass

    async def stop(self):
        pass

    async def search(self, q, desc):
        self.seen.append((q.copy(), desc.copy()))
        return np.array([3, 1, 2], np.uint32), np.array([3 << 30, 1 << 31, -(1 << 29)], np.int64), 3, 40


def _stub_server(mse, embed=None):
    from mse_b200.query_server import PackedIndex, QueryServer
    ix = PackedIndex(vecs=None, shards=[], n_dims=8, urls=["a", "b", "c", "d"], dimensions=np.array([[1, 2], [3, 4], [5, 6], [7, 8]], np.uint32),
                     count=4, dead_count=1, timestamps=np.array([10, 11, 12, 13]), node_scores=[[0.5]] * 4, node_shards=[[0, 1]] * 4)
    b = StubBatcher()
    cfg = {"descriptor_names": ["nsfw", "meme"], "search_list": 64, "beam_width": 4}
    return QueryServer(cfg, ix, embed=embed, registry=CollectorRegistry(), batcher=b), b


def test_json_boundary(mse):
    srv, b = _stub_server(mse, embed=lambda batch: [np.full(8, 0.5, np.float16).tobytes()] * len(batch.get("text", batch.get("images"))))

    async def scenario(c):
        r = await c.get("/")
        assert r.status == 200 and r.headers["Access-Control-Allow-Origin"] == "*" and r.content_type == "application/json"
        assert await r.text() == '{"n_total":3,"predefined_embedding_names":["nsfw","meme"],"d_emb":8}'      # FrontendInit, field order of common.rs:177-181
        body = {"terms": [{"embedding": [1, 0, 0, 0, 0, 0, 0, 0], "weight": 2.0}, {"text": "cat"}, {"predefined_embedding": "meme", "weight": 0.5},
                          {"predefined_embedding": "unknown"}], "debug_enabled": True, "k": 2}
        r = await c.post("/", data=json.dumps(body))
        assert r.status == 200
        out = json.loads(await r.text())
        assert out["formats"] == [] and out["extensions"] == {} and len(out["matches"]) == 2
        assert out["matches"][0] == [0.75, "d", "", 0, [7, 8], [[0.5], [0, 1], 13]]                          # (score / 2^32) as f32, url, "", 0, dims, debug
        assert out["matches"][1][:5] == [0.5, "b", "", 0, [3, 4]]
        q, desc = b.seen[-1]
        assert np.allclose(q, [2.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5]) and np.allclose(desc, [0.0, 0.5 / 512.0])   # common.rs:215-274, :463-471
        r = await c.post("/", data=json.dumps({"terms": [{"embedding": [0.0] * 8}]}))
        assert json.loads(await r.text())["matches"][2] == [-0.125, "c", "", 0, [5, 6], None]
        assert (await c.options("/")).status == 204
        r = await c.get("/nothing")
        assert r.status == 404 and await r.text() == "Not Found"
        r = await c.post("/", data=b"x" * ((1 << 23) + 10))
        assert r.status == 413 and await r.text() == "Body too big"
        m = await (await c.get("/metrics")).text()
        assert "mse_queries_total 2.0" in m and 'mse_terms_total{type="text"} 1.0' in m and 'mse_terms_total{type="embedding"} 2.0' in m
    _run(srv, scenario)


def test_f32_json_and_entry_points(mse):
    from mse_b200.query_server import PackedIndex, SearchBatcher, f32_json
    assert json.dumps(f32_json(np.float32(0.3))) == "0.3" and json.dumps(f32_json(0.1 + 0.2)) == "0.3"
    assert json.dumps(f32_json(np.float32(1) / np.float32(3))) == "0.33333334"
    rng = np.random.default_rng(0)
    cents = [(rng.standard_normal(16).astype(np.float32), 100 + i) for i in range(5)]
    cents.append((cents[2][0].copy(), 999))                                         # a tie: position_max_by_key keeps the LAST maximum
    b = SearchBatcher.__new__(SearchBatcher)
    b.centroids64 = np.stack([c for c, _ in cents]).astype(np.float64)
    b.medioids = np.asarray([m for _, m in cents], np.uint32)
    q = np.stack([cents[2][0] * 3, cents[4][0], -cents[0][0]]).astype(np.float32)
    got = b.entry_points(q)
    from mse_b200.query import select_shard
    want = [cents[select_shard(cents, qi)][1] for qi in q]
    assert got.tolist() == want and got[0] == 999 and got[1] == 104


@pytest.mark.gpu
def test_http_to_gpu_search_matches_oracle(mse, oracle):
    """concurrent POSTs -> one or more GPU batches -> every response equals the oracle's beam search (PQ ADC + descriptor bias, entry
    point by nearest shard centroid) + 0.95-cosine de-duplication + stable sort"""
    from mse_b200 import diskann as dk
    from mse_b200.query_server import PackedIndex, QueryServer
    n, R, L, W = 3000, 16, 40, 3
    x = clustered_f16(71, n, n_clusters=20)
    x[50:60] = x[49]                                                               # near-duplicates for the runtime de-duplication
    g = oracle.IndexGraph(n, R)
    oracle.random_fill_graph(g, R, seed=3)
    med = oracle.medioid(x)
    oracle.build_graph(g, med, x, oracle.make_config(r=R, l=L, maxc=120), seed=5)
    rng = np.random.default_rng(9)
    d = 1152
    cents = rng.standard_normal((256, d)).astype(np.float32) * 0.05
    T = np.linalg.qr(rng.standard_normal((d, d)))[0].astype(np.float32)
    po = oracle.ProductQuantizer(cents, T, 18)
    pg = dk.ProductQuantizer(cents, T, 18)
    codes = pg.quantize_batch(x.astype(np.float32))
    desc = rng.integers(0, 256, (n, 2)).astype(np.uint8)
    has_url = (rng.random(n) > 0.05).astype(np.uint8)
    vl = dk.VectorList.from_f16s(x)
    vl.set_graph(g.adj.copy(), g.deg.copy())
    vl.set_pq_codes(codes)
    vl.set_descriptors(desc, has_url)
    shards = [(x[100:600].astype(np.float32).mean(axis=0), med), (x[2000:2400].astype(np.float32).mean(axis=0), 2100)]
    ix = PackedIndex(vecs=vl, shards=shards, n_dims=d, urls=[f"u{i}" if has_url[i] else "" for i in range(n)],
                     dimensions=np.stack([np.arange(n), np.arange(n) + 1], axis=1).astype(np.uint32), count=n, pq=pg)
    cfg = {"descriptor_names": ["nsfw", "meme"], "search_list": L, "beam_width": W, "max_batch": 16, "batch_window_ms": 20, "max_results": 400}
    srv = QueryServer(cfg, ix, registry=CollectorRegistry())
    queries = np.concatenate([x[700:712], clustered_f16(72, 12, n_clusters=20)]).astype(np.float32) * np.float32(1.5)
    weights = [None, 0.5, -1.0]

    async def scenario(c):
        async def one(i):
            terms = [{"embedding": queries[i].tolist()}]
            if weights[i % 3] is not None:
                terms.append({"predefined_embedding": "meme", "weight": weights[i % 3]})
            r = await c.post("/", data=json.dumps({"terms": terms}))
            assert r.status == 200
            return json.loads(await r.text())["matches"]
        return await asyncio.gather(*[one(i) for i in range(len(queries))])
    outs = _run(srv, scenario)
    adj, off = g.to_csr()
    from mse_b200.query import select_shard
    for i, matches in enumerate(outs):
        q32 = queries[i]
        scales = np.zeros(2, np.float32)
        if weights[i % 3] is not None:
            scales[1] = np.float32(weights[i % 3]) * np.float32(1.0 / 512.0)
        start = shards[select_shard(shards, q32)][1]
        lut = po.preprocess_query(q32)
        ids, sc, _ = oracle.beam_search(x, adj, off, codes, lut, start, q32.astype(np.float16), L, W, descriptors=desc, desc_scales=scales, has_url=has_url)
        keep = oracle.dedup_visited(x, ids)
        ids, sc = ids[keep], sc[keep]
        o = np.argsort(-sc, kind="stable")
        want = [[float(str(np.float32(s / 4294967296.0))), f"u{j}", "", 0, [int(j), int(j) + 1], None] for j, s in zip(ids[o].tolist(), sc[o].tolist())]
        assert matches == want, i
    m = srv.registry.get_sample_value("mse_search_batches_total")
    assert m is not None and m < len(queries)                                       # requests were served in batches
    vl.close()
