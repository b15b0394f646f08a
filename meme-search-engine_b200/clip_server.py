"""Embedding service with the wire format of the reference's clip_server.py, built around a request-coalescing batcher.

Boundary kept byte-compatible with the reference (its line numbers in brackets):
  POST /          body = msgpack map, key "text" (list of str, or one bare str as src/get_embedding.py:21 sends) or "images"
                  (list of encoded image bytes); "text" is looked at first [135-139]
                  200 -> msgpack array with one fp16-LE byte string (unit-norm vector) per item, request order [166,170]
                  500 -> msgpack str carrying the message [167-170]: "max batch size is N" for a request above
                         max_batch_size [136,139], "images or text required" when both are missing/empty [142]
  GET /config     msgpack {"model", "batch", "image_size", "embedding_size"} [176-183]; src/common.rs:24-29 reads it
  GET /           204 [185-187]
  GET /metrics    Prometheus text; modelserver_total_items{model,modality}, modelserver_inftime{model,batch_size},
                  modelserver_batchcount{model} [86-88,189-191]
  body limit      2**26 bytes [148]
  config file     argv[1], JSON: device, model, model_name, max_batch_size, port [19-20,27-28,197].  `model_path` names the
                  MSEW0001 weights container (mse_b200.weights) and `tokenizer_model` a SentencePiece file (the reference
                  gets both from open_clip [23-25]).  Optional: batch_window_ms (2), queue_depth (10), decode_threads (4).

What differs from the reference is the schedule behind that boundary.  The reference runs one request = one model batch
through two single-threaded stages [91-146], so the towers only ever see the batch one client happened to send (the Rust
ingest client compensates by sending `batch`-sized chunks, src/main.rs:680-694; query traffic is batch 1).  Here every
request is decoded/tokenised on a small worker pool and handed to a `Coalescer`: rows of concurrent requests of the same
modality are packed into one tower call of up to max_batch_size rows (waiting at most batch_window_ms for company while
the GPU lane is idle; while it is busy, arrivals simply pile up for the next call), and the result rows are dealt back to
their requests.  One tower call is in flight at a time -- the C ABI's rule of one call per encoder handle.

The towers run in libmse_b200.so (csrc/encoder.cu); nothing in this file computes an embedding.
"""
from __future__ import annotations

import asyncio
import io
import json
import string
import sys
import time
import traceback
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field

import msgpack
import numpy as np
from aiohttp import web
from prometheus_client import REGISTRY, CollectorRegistry, Counter, Histogram, generate_latest

MSGPACK = "application/msgpack"


class RequestError(Exception):
    """A request the reference answers with 500 + msgpack str."""


class SiglipTokenizer:
    """The text side of open_clip's SigLIP pipeline (clip_server.py:25,137; parameters as misc/clip_accursed.py:50-55 records
    them for the same checkpoints: c4_en SentencePiece vocabulary of 32 000, max_len 64, eos "sticky", pad_value 1):
    canonicalise, SentencePiece-encode, cut to 63 pieces, append EOS (id 1), pad with id 1.

    UNVERIFIED against the real tokenizer: no SentencePiece model for c4_en ships with this image or the reference, so only the
    canonicaliser and the framing (EOS / pad / truncation) are tested (with a SentencePiece model trained in the test)."""

    _strip = str.maketrans("", "", string.punctuation)

    def __init__(self, model_file: str, context_length: int = 64, pad_id: int = 1, eos_id: int = 1):
        import sentencepiece as spm
        self.sp = spm.SentencePieceProcessor(model_file=model_file)
        self.context_length, self.pad_id, self.eos_id = context_length, pad_id, eos_id

    @classmethod
    def canonicalize(cls, text: str) -> str:
        # open_clip canonicalize_text: underscores become spaces BEFORE punctuation is dropped ("foo_bar" -> "foo bar"),
        # then lower-case and collapse runs of whitespace
        return " ".join(text.replace("_", " ").translate(cls._strip).lower().split())

    def __call__(self, texts) -> np.ndarray:
        if isinstance(texts, str):
            texts = [texts]
        rows = np.full((len(texts), self.context_length), self.pad_id, np.int32)
        for row, text in zip(rows, texts):
            pieces = list(self.sp.encode(self.canonicalize(text)))[: self.context_length - 1]
            pieces.append(self.eos_id)
            row[: len(pieces)] = pieces
        return rows


def is_client_bmp(data: bytes, size: int) -> bool:
    """True for exactly what src/common.rs:42-53 produces: a size x size 24-bit uncompressed BMP (54-byte header + BGR rows)."""
    if len(data) < 54 or data[:2] != b"BM":
        return False
    w, h = int.from_bytes(data[18:22], "little", signed=True), int.from_bytes(data[22:26], "little", signed=True)
    bpp, comp, off = int.from_bytes(data[28:30], "little"), int.from_bytes(data[30:34], "little"), int.from_bytes(data[10:14], "little")
    stride = (size * 3 + 3) & ~3
    return w == size and abs(h) == size and bpp == 24 and comp == 0 and off + stride * size <= len(data)


def decode_image(data: bytes, size: int) -> np.ndarray:
    """Encoded image bytes -> [size, size, 3] u8 RGB.  The reference's clients send size x size 24-bit BMPs
    (src/common.rs:31-54), for which open_clip's Resize is the identity; anything else is squashed bicubically."""
    from PIL import Image
    im = Image.open(io.BytesIO(data)).convert("RGB")
    if im.size != (size, size):
        im = im.resize((size, size), Image.BICUBIC)
    return np.asarray(im, dtype=np.uint8)


@dataclass
class _Ticket:
    """Rows of one request waiting for a tower call."""
    modality: str                      # "text" | "image" (decoded pixels) | "bmp" (whole client BMP files, unpacked on the device)
    rows: np.ndarray                   # [m, 64] int32 token ids, [m, S, S, 3] u8 pixels, or [m, file length] u8
    done: asyncio.Future = field(repr=False)
    arrived: float = field(default_factory=time.monotonic)


class Coalescer:
    """Packs tickets of one modality into tower calls of at most `limit` rows.  Single consumer coroutine; the tower call runs
    on a one-thread executor (the GPU lane), so the event loop keeps accepting requests while the GPU works."""

    def __init__(self, run_batch, limit: int, window_s: float):
        self._run_batch, self._limit, self._window = run_batch, limit, window_s
        self._waiting: asyncio.Queue[_Ticket] = asyncio.Queue()
        self._held: _Ticket | None = None          # a ticket that did not fit the previous call
        self._lane = ThreadPoolExecutor(max_workers=1, thread_name_prefix="mse-gpu-lane")
        self._task: asyncio.Task | None = None

    def start(self):
        self._task = asyncio.get_running_loop().create_task(self._consume())

    async def stop(self):
        if self._task:
            self._task.cancel()
            try:
                await self._task
            except asyncio.CancelledError:
                pass
        self._lane.shutdown(wait=False)

    def depth(self) -> int:
        return self._waiting.qsize() + (self._held is not None)

    async def submit(self, modality: str, rows: np.ndarray) -> np.ndarray:
        t = _Ticket(modality, rows, asyncio.get_running_loop().create_future())
        self._waiting.put_nowait(t)
        return await t.done

    async def _next_group(self) -> list[_Ticket]:
        head = self._held or await self._waiting.get()
        self._held = None
        group, rows = [head], head.rows.shape[0]
        close_at = head.arrived + self._window
        while rows < self._limit:
            try:
                t = self._waiting.get_nowait()
            except asyncio.QueueEmpty:
                wait = close_at - time.monotonic()
                if wait <= 0:
                    break
                try:
                    t = await asyncio.wait_for(self._waiting.get(), wait)
                except asyncio.TimeoutError:
                    break
            if t.modality != head.modality or t.rows.shape[1:] != head.rows.shape[1:] or rows + t.rows.shape[0] > self._limit:
                self._held = t                      # opens the next call; arrival order is kept
                break
            group.append(t)
            rows += t.rows.shape[0]
        return group

    async def _consume(self):
        loop = asyncio.get_running_loop()
        while True:
            group = await self._next_group()
            stacked = group[0].rows if len(group) == 1 else np.concatenate([t.rows for t in group])
            try:
                feats = await loop.run_in_executor(self._lane, self._run_batch, group[0].modality, stacked, len(group))
            except asyncio.CancelledError:
                raise
            except Exception as e:  # the tower call failed: every request of the call gets the message (reference: 500 + str)
                traceback.print_exc()
                for t in group:
                    if not t.done.done():
                        t.done.set_exception(RequestError(str(e)))
                continue
            at = 0
            for t in group:
                m = t.rows.shape[0]
                if not t.done.done():
                    t.done.set_result(feats[at:at + m])
                at += m


class ClipServer:
    def __init__(self, config: dict, encoder=None, tokenizer=None, registry: CollectorRegistry | None = None):
        self.config = config
        self.max_batch = int(config["max_batch_size"])
        self.model_label = config["model_name"]
        if encoder is None:
            dev = str(config.get("device", "cuda:0"))
            if not dev.startswith("cuda"):
                raise RuntimeError(f'device "{dev}": this server only runs its towers on a B200 (no CPU path)')
            from .encoder import Encoder
            encoder = Encoder(config["model_path"], device=int(dev.split(":")[1]) if ":" in dev else 0, max_batch=self.max_batch)
        self.encoder = encoder
        if tokenizer is None and config.get("tokenizer_model"):
            tokenizer = SiglipTokenizer(config["tokenizer_model"], context_length=getattr(encoder, "ctx", 64))
        self.tokenizer = tokenizer
        self.image_size = int(getattr(encoder, "image_size", 384))
        self.queue_depth = int(config.get("queue_depth", 10))
        self.bmp_on_device = bool(config.get("bmp_on_device", True)) and hasattr(encoder, "encode_image_bmp")
        self.registry = registry if registry is not None else REGISTRY
        self.items_ctr = Counter("modelserver_total_items", "Items run through model server", ["model", "modality"], registry=self.registry)
        self.inference_time_hist = Histogram("modelserver_inftime", "Time running inference", ["model", "batch_size"], registry=self.registry)
        self.batch_count_ctr = Counter("modelserver_batchcount", "Inference batches run", ["model"], registry=self.registry)
        self.coalesced_hist = Histogram("modelserver_requests_per_batch", "Requests packed into one tower call", ["model"],
                                        buckets=(1, 2, 4, 8, 16, 32, 64, 128, 256), registry=self.registry)
        self._decoders = ThreadPoolExecutor(max_workers=int(config.get("decode_threads", 4)), thread_name_prefix="mse-decode")
        self.coalescer = Coalescer(self._tower_call, self.max_batch, float(config.get("batch_window_ms", 2.0)) * 1e-3)
        self._in_prep = 0
        self.app = web.Application(client_max_size=2 ** 26)
        self.app.add_routes([web.post("/", self.embed), web.get("/config", self.describe), web.get("/", self.alive),
                             web.get("/metrics", self.metrics)])
        self.app.cleanup_ctx.append(self._lifecycle)

    async def _lifecycle(self, app):
        self.coalescer.start()
        yield
        await self.coalescer.stop()
        self._decoders.shutdown(wait=False)

    # -- GPU lane ------------------------------------------------------------------------------------------------
    def _tower_call(self, modality: str, rows: np.ndarray, n_requests: int) -> np.ndarray:
        n = rows.shape[0]
        self.items_ctr.labels(self.model_label, "image" if modality == "bmp" else modality).inc(n)
        label = "image" if modality == "bmp" else modality
        with self.inference_time_hist.labels(f"{self.model_label}-{label}", n).time():
            if modality == "text":
                feats = self.encoder.encode_text(rows)
            elif modality == "bmp":                       # whole BMP files, unpacked on the device (mse_encode_images_bmp)
                feats = self.encoder.encode_image_bmp(list(rows))
            else:
                feats = self.encoder.encode_image(rows)
        self.batch_count_ctr.labels(self.model_label).inc()
        self.coalesced_hist.labels(self.model_label).observe(n_requests)
        return feats

    # -- request preparation (worker pool) -------------------------------------------------------------------------
    def _prepare(self, payload) -> tuple[str, np.ndarray]:
        if not isinstance(payload, dict):
            raise RequestError("images or text required")
        text, images = payload.get("text"), payload.get("images")
        if text:
            texts = [text] if isinstance(text, str) else list(text)
            if len(texts) > self.max_batch:
                raise RequestError(f"max batch size is {self.max_batch}")
            if self.tokenizer is None:
                raise RequestError("text requests need `tokenizer_model` (a SentencePiece model file) in the config")
            return "text", np.ascontiguousarray(self.tokenizer(texts), np.int32)
        if images:
            if len(images) > self.max_batch:
                raise RequestError(f"max batch size is {self.max_batch}")
            # the reference's clients send image_size x image_size 24-bit BMPs (src/common.rs:42-53): those go to the GPU as they are
            if self.bmp_on_device and all(is_client_bmp(b, self.image_size) for b in images) and len({len(b) for b in images}) == 1:
                return "bmp", np.stack([np.frombuffer(b, np.uint8) for b in images])
            return "image", np.stack([decode_image(b, self.image_size) for b in images])
        raise RequestError("images or text required")

    # -- routes ----------------------------------------------------------------------------------------------------
    @staticmethod
    def _reply(status: int, obj) -> web.Response:
        return web.Response(body=msgpack.dumps(obj), status=status, content_type=MSGPACK)

    async def embed(self, request: web.Request) -> web.Response:
        # the reference's request queue holds 10 entries and put_nowait raises beyond that [125,161]
        if self._in_prep + self.coalescer.depth() >= 2 * self.queue_depth:
            return self._reply(500, "queue full")
        raw = await request.read()
        self._in_prep += 1
        try:
            payload = msgpack.loads(raw)
            modality, rows = await asyncio.get_running_loop().run_in_executor(self._decoders, self._prepare, payload)
        except RequestError as e:
            return self._reply(500, str(e))
        except Exception as e:  # undecodable body / image
            traceback.print_exc()
            return self._reply(500, str(e))
        finally:
            self._in_prep -= 1
        try:
            feats = await self.coalescer.submit(modality, rows)
        except RequestError as e:
            print(e)
            return self._reply(500, str(e))
        return self._reply(200, [np.asarray(v, np.float16).tobytes() for v in feats])

    async def describe(self, request: web.Request) -> web.Response:
        return self._reply(200, {"model": self.config["model"], "batch": self.max_batch, "image_size": (self.image_size, self.image_size),
                                 "embedding_size": int(getattr(self.encoder, "dim", 1152))})

    async def alive(self, request: web.Request) -> web.Response:
        return web.Response(status=204)

    async def metrics(self, request: web.Request) -> web.Response:
        return web.Response(body=generate_latest(self.registry))

    async def serve(self):
        runner = web.AppRunner(self.app)
        await runner.setup()
        await web.TCPSite(runner, "", self.config["port"]).start()
        print("Ready")


def main(argv=None):
    argv = sys.argv if argv is None else argv
    with open(argv[1]) as f:
        server = ClipServer(json.load(f))
    print("Model loaded")
    loop = asyncio.new_event_loop()
    asyncio.set_event_loop(loop)
    try:
        loop.run_until_complete(server.serve())
        loop.run_forever()
    except KeyboardInterrupt:
        sys.exit(0)


if __name__ == "__main__":
    main()
