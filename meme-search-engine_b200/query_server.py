"""Work-alike of the reference's large-scale query server (src/query_disk_index.rs:395-610, the `query-disk-index` binary) with
the packed index resident in HBM and request micro-batching in front of the batched search kernels.

Boundary (reference lines in brackets; JSON unless noted, CORS headers as :603-606):
  GET  /          FrontendInit {"n_total": count - dead_count, "predefined_embedding_names": [...], "d_emb": n_dims}  [:415-419]
  POST /          QueryRequest {"terms": [{embedding | image (base64) | text | predefined_embedding, weight}], "k", "include_video",
                  "debug_enabled"} (common.rs:185-209) -> QueryResult {"matches": [[score, url, "", 0, [w, h], debug | null], ...],
                  "formats": [], "extensions": {}} (common.rs:176-183); bodies over 2^23 bytes -> 413 "Body too big"  [:421-426]
  GET  /metrics   Prometheus text: mse_queries, mse_terms{type}, mse_node_reads, mse_pq_comparisons  [:67-70,553-560]
  OPTIONS /       204; anything else 404 "Not Found"  [:584-593]

Per query the reference does: get_total_embedding (common.rs:215-274) -> entry point = medioid of the shard whose centroid has
the largest dot product with the query [:447-450] -> descriptor scales weight/512 for named descriptor terms [:463-471] ->
PQ query tables [:475] -> beam search over the packed index, compressed scores for the frontier, exact scores + descriptor bias
for expanded nodes [:479, :144-212] -> runtime de-duplication (drop a visited node when an earlier kept one has cosine > 0.95)
[:486-527] -> stable sort by score [:529] -> JSON.  One query at a time per connection, NVMe reads through io_uring.

Here the same steps run for a BATCH of concurrent requests: a `Coalescer`-style queue collects queries for up to
batch_window_ms (or max_batch), and one pass of mse_search_beam_dev + mse_dedup_topk_dev serves them all (csrc/graph.cu), every
query from its own entry point.  Candidate scores come from the index's own PQ codes (index.pq-codes.bin + the header's
quantizer) or, for an index encoded with mse_index_encode_rabitq, from the RabitQ sign codes.
"""
from __future__ import annotations

import asyncio
import io
import json
import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field

import numpy as np
from aiohttp import web
from prometheus_client import REGISTRY, CollectorRegistry, Counter, generate_latest

from .query import get_total_embedding

CORS = {"Access-Control-Allow-Origin": "*", "Access-Control-Allow-Methods": "GET, POST, OPTIONS", "Access-Control-Allow-Headers": "Content-Type"}
SCALE_F64 = 4294967296.0           # diskann/src/vector.rs:46-47
DESCRIPTOR_WEIGHT = 1.0 / 512.0    # query_disk_index.rs:468


def f32_json(x) -> float:
    """serde_json writes an f32 with the shortest decimal that round-trips as f32; json.dumps of a Python float would print the f64
    expansion of the same value"""
    return float(str(np.float32(x)))


@dataclass
class PackedIndex:
    """What query_disk_index.rs::Index (:650-662) holds, with the node records unpacked: vectors, adjacency, codes and descriptors in
    HBM behind `vecs`; the per-node metadata the response needs on the host."""
    vecs: object                                   # diskann.VectorList with graph + codes (+ descriptors, has_url)
    shards: list                                   # [(centroid f32[d], medioid LOCAL row)]  IndexHeader.shards
    n_dims: int
    urls: list                                     # PackedIndexEntry.url per node ("" = not returned, :172)
    dimensions: np.ndarray                         # [n, 2] u32
    count: int = 0
    dead_count: int = 0
    pq: object = None                              # diskann.ProductQuantizer (PQ ADC traversal) ...
    rabitq: object = None                          # ... or diskann.RabitQ (sign-code traversal)
    n_centroids: int = 256
    timestamps: np.ndarray | None = None           # debug payload (:533): node.scores, node.shards, node.timestamp
    node_scores: list | None = None
    node_shards: list | None = None


class SearchBatcher:
    """Collects the (query, descriptor scales, entry point) triples of concurrent requests and runs them as one GPU batch."""

    def __init__(self, index: PackedIndex, search_list: int, beam_width: int, max_batch: int, window_s: float, max_results: int, on_batch=None):
        import torch
        self.torch = torch
        self.ix, self.L, self.W, self.max_batch, self.window, self.topk = index, search_list, beam_width, max_batch, window_s, max_results
        self.dev = torch.device("cuda", index.vecs.device)
        self.pending: asyncio.Queue = asyncio.Queue()
        self.lane = ThreadPoolExecutor(max_workers=1, thread_name_prefix="mse-search-lane")
        self.task = None
        self.on_batch = on_batch
        cents = np.stack([np.asarray(c, np.float32) for c, _ in index.shards]) if index.shards else np.zeros((0, index.n_dims), np.float32)
        self.centroids64 = cents.astype(np.float64)
        self.medioids = np.asarray([m for _, m in index.shards], np.uint32)

    def start(self):
        self.task = asyncio.get_running_loop().create_task(self._consume())

    async def stop(self):
        if self.task:
            self.task.cancel()
            try:
                await self.task
            except asyncio.CancelledError:
                pass
        self.lane.shutdown(wait=False)

    async def search(self, query: np.ndarray, desc_scales: np.ndarray):
        fut = asyncio.get_running_loop().create_future()
        self.pending.put_nowait((query, desc_scales, fut, time.monotonic()))
        return await fut

    async def _consume(self):
        loop = asyncio.get_running_loop()
        while True:
            first = await self.pending.get()
            group, close_at = [first], first[3] + self.window
            while len(group) < self.max_batch:
                try:
                    group.append(self.pending.get_nowait())
                except asyncio.QueueEmpty:
                    wait = close_at - time.monotonic()
                    if wait <= 0:
                        break
                    try:
                        group.append(await asyncio.wait_for(self.pending.get(), wait))
                    except asyncio.TimeoutError:
                        break
            try:
                results = await loop.run_in_executor(self.lane, self._run, np.stack([g[0] for g in group]), np.stack([g[1] for g in group]))
                for g, r in zip(group, results):
                    if not g[2].done():
                        g[2].set_result(r)
            except asyncio.CancelledError:
                raise
            except Exception as e:
                for g in group:
                    if not g[2].done():
                        g[2].set_exception(e)

    def entry_points(self, q: np.ndarray) -> np.ndarray:
        """:447-450: position_max_by_key over trunc(2^32 * <centroid, q>) (simsimd f64 dot); the LAST maximum wins"""
        if self.medioids.size == 0:
            return np.zeros(q.shape[0], np.uint32)
        keys = np.trunc(np.clip((self.centroids64 @ q.astype(np.float64).T) * SCALE_F64, -9.223372036854775808e18, 9.223372036854775807e18))
        keys = np.nan_to_num(keys, nan=0.0)
        last_max = keys.shape[0] - 1 - np.argmax(keys[::-1], axis=0)
        return self.medioids[last_max]

    def _run(self, q32: np.ndarray, desc: np.ndarray):
        """one GPU batch: -> per query (ids, scores i64, n_expanded, n_pq) of the kept nodes, best first"""
        from . import diskann as dk
        torch, ix, nq = self.torch, self.ix, q32.shape[0]
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        starts = torch.from_numpy(self.entry_points(q32).astype(np.int64)).to(self.dev, torch.int32)
        d_q32 = torch.from_numpy(np.ascontiguousarray(q32, np.float32)).to(self.dev)
        d_q16 = d_q32.to(torch.float16).contiguous()                     # half::f16::from_f32 (:477): round to nearest even
        d_desc = torch.from_numpy(np.ascontiguousarray(desc, np.float32)).to(self.dev) if desc.shape[1] else None
        topk = self.topk
        t_ids = torch.empty((nq, topk), dtype=torch.int32, device=self.dev)
        t_sc = torch.empty((nq, topk), dtype=torch.int64, device=self.dev)
        t_len = torch.empty(nq, dtype=torch.int32, device=self.dev)
        cm = torch.empty(nq, dtype=torch.int64, device=self.dev)
        pc = torch.empty(nq, dtype=torch.int64, device=self.dev)
        kw = {"d_desc_scales": d_desc.data_ptr() if d_desc is not None else 0, "d_starts": starts.data_ptr()}
        if ix.rabitq is not None:
            qtm = torch.empty((nq, ix.rabitq.output_dims + 1), dtype=torch.float32, device=self.dev)
            ix.rabitq.query_dev(d_q32.data_ptr(), nq, qtm.data_ptr(), stream)
            kw.update(d_qtm=qtm.data_ptr(), rabitq=ix.rabitq)
        else:
            luts = torch.from_numpy(ix.pq.preprocess_query(q32)).to(self.dev)   # ProductQuantizer::preprocess_query (:475)
            kw.update(d_luts=luts.data_ptr(), n_centroids=ix.n_centroids)
        dk.beam_search_dev(ix.vecs, d_q16.data_ptr(), nq, self.L, self.W, 0, topk, t_ids.data_ptr(), t_sc.data_ptr(), t_len.data_ptr(), cm.data_ptr(),
                           pc.data_ptr(), stream, **kw)
        dk.dedup_topk_dev(ix.vecs, nq, topk, t_ids.data_ptr(), t_sc.data_ptr(), t_len.data_ptr(), stream=stream)   # :486-529
        dk.greedy_search_check(ix.vecs, nq)
        ids, sc, ln = t_ids.cpu().numpy().view(np.uint32), t_sc.cpu().numpy(), t_len.cpu().numpy()
        cmh, pch = cm.cpu().numpy(), pc.cpu().numpy()
        if self.on_batch:
            self.on_batch(nq, int(cmh.sum()), int(pch.sum()))
        return [(ids[i, : ln[i]].copy(), sc[i, : ln[i]].copy(), int(cmh[i]), int(pch[i])) for i in range(nq)]


class QueryServer:
    def __init__(self, config: dict, index: PackedIndex, embed=None, resize_image=None, registry: CollectorRegistry | None = None, batcher=None):
        """config: ServerConfig (:57-64) keys descriptor_names, search_list, beam_width (+ listen_address, clip_server) and the batching
        knobs max_batch (64), batch_window_ms (1), max_results (1000).  embed(batch) -> list of fp16-LE byte strings is the clip_server
        boundary (an HTTP client, or clip_server.ClipServer's encoder in process); resize_image(bytes) -> bytes as common.rs:31-54."""
        self.config, self.index, self.embed, self.resize_image = config, index, embed, resize_image
        self.descriptor_names = list(config.get("descriptor_names", []))
        self.registry = registry if registry is not None else REGISTRY
        self.queries = Counter("mse_queries", "queries executed", registry=self.registry)
        self.terms = Counter("mse_terms", "terms used in queries, by type", ["type"], registry=self.registry)
        self.node_reads = Counter("mse_node_reads", "graph nodes read", registry=self.registry)
        self.pq_cmps = Counter("mse_pq_comparisons", "product quantization comparisons", registry=self.registry)
        self.batches = Counter("mse_search_batches", "GPU search batches run", registry=self.registry)
        self.batcher = batcher or SearchBatcher(index, int(config["search_list"]), int(config["beam_width"]), int(config.get("max_batch", 64)),
                                                float(config.get("batch_window_ms", 1.0)) * 1e-3, int(config.get("max_results", 1000)), self._account)
        self.embed_pool = ThreadPoolExecutor(max_workers=int(config.get("embed_threads", 2)), thread_name_prefix="mse-embed")
        self.app = web.Application(client_max_size=1 << 23)
        self.app.add_routes([web.get("/", self.frontend_init), web.post("/", self.query), web.get("/metrics", self.metrics),
                             web.options("/", self.preflight), web.route("*", "/{tail:.*}", self.not_found)])
        self.app.cleanup_ctx.append(self._lifecycle)

    async def _lifecycle(self, app):
        self.batcher.start()
        yield
        await self.batcher.stop()
        self.embed_pool.shutdown(wait=False)

    def _account(self, nq, node_reads, pq_cmps):
        self.batches.inc()
        self.node_reads.inc(node_reads)
        self.pq_cmps.inc(pq_cmps)

    @staticmethod
    def _json(obj, status=200):
        return web.Response(body=json.dumps(obj, separators=(",", ":")).encode(), status=status, content_type="application/json", headers=CORS)

    async def frontend_init(self, request):
        ix = self.index
        return self._json({"n_total": int(ix.count - ix.dead_count), "predefined_embedding_names": self.descriptor_names, "d_emb": int(ix.n_dims)})

    async def preflight(self, request):
        return web.Response(status=204, body=b"", content_type="application/json", headers=CORS)

    async def not_found(self, request):
        return web.Response(status=404, body=b"Not Found", content_type="application/json", headers=CORS)

    async def metrics(self, request):
        return web.Response(body=generate_latest(self.registry), headers={**CORS, "Content-Type": "text/plain; version=0.0.4"})

    def _assemble(self, terms):
        for t in terms:
            for kind in ("image", "text", "embedding"):
                if t.get(kind) is not None:
                    self.terms.labels(kind).inc()
        q = get_total_embedding(terms, self.index.n_dims, self.embed, self.resize_image, {})      # no predefined embeddings here (:459)
        desc = np.zeros(len(self.descriptor_names), np.float32)
        for t in terms:                                                                           # :463-471
            name = t.get("predefined_embedding")
            if name is not None and name in self.descriptor_names:
                w = t.get("weight")
                desc[self.descriptor_names.index(name)] = np.float32(1.0 if w is None else w) * np.float32(DESCRIPTOR_WEIGHT)
        return q.astype(np.float32), desc

    async def query(self, request):
        try:
            raw = await request.read()
        except web.HTTPRequestEntityTooLarge:
            return web.Response(status=413, body=b"Body too big", content_type="application/json", headers=CORS)
        body = json.loads(raw)
        terms = body["terms"]
        needs_embed = any(t.get("image") is not None or t.get("text") is not None for t in terms)
        if needs_embed:
            q, desc = await asyncio.get_running_loop().run_in_executor(self.embed_pool, self._assemble, terms)
        else:
            q, desc = self._assemble(terms)
        ids, scores, _, _ = await self.batcher.search(q, desc)
        self.queries.inc()
        k = body.get("k")
        if k is not None:
            ids, scores = ids[: int(k)], scores[: int(k)]
        ix, debug = self.index, bool(body.get("debug_enabled", False))
        matches = []
        for i, s in zip(ids.tolist(), scores.tolist()):
            dbg = None
            if debug:
                dbg = [[f32_json(v) for v in (ix.node_scores[i] if ix.node_scores is not None else [])],
                       [int(v) for v in (ix.node_shards[i] if ix.node_shards is not None else [])],
                       int(ix.timestamps[i]) if ix.timestamps is not None else 0]
            matches.append([f32_json(np.float32(s / SCALE_F64)), ix.urls[i], "", 0, [int(ix.dimensions[i][0]), int(ix.dimensions[i][1])], dbg])
        return self._json({"matches": matches, "formats": [], "extensions": {}})

    async def serve(self):
        host, _, port = str(self.config.get("listen_address", "0.0.0.0:5601")).rpartition(":")
        runner = web.AppRunner(self.app)
        await runner.setup()
        await web.TCPSite(runner, host or "0.0.0.0", int(port)).start()


def resize_for_embed(image_size: int):
    """common.rs:31-54 resize_for_embed_sync: decode, squash to image_size x image_size (the reference uses Hamming down / Lanczos3 up;
    PIL's LANCZOS both ways here), encode as a 24-bit BMP for the embedding service"""
    def fn(data: bytes) -> bytes:
        from PIL import Image
        im = Image.open(io.BytesIO(data)).convert("RGB").resize((image_size, image_size), Image.LANCZOS)
        out = io.BytesIO()
        im.save(out, format="BMP")
        return out.getvalue()
    return fn
