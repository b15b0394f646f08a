"""Work-alike of the reference's embedding service (clip_server.py) on top of the C-ABI towers.

Same boundary, byte for byte (reference lines in brackets):
  POST /          msgpack map {"text": [str, ...]} or {"images": [bytes, ...]}; text wins if both [135-139]
                  200 -> msgpack array of fp16-LE byte strings, one unit-norm vector per item [166,170]
                  500 -> msgpack str with the error [167-170]; batches over max_batch_size fail the assertion [136,139]
  GET /config     msgpack {"model", "batch", "image_size", "embedding_size"} [176-183]
  GET /           204 [185-187]
  GET /metrics    Prometheus text with modelserver_total_items / modelserver_inftime / modelserver_batchcount [86-88,189-191]
  threading       event loop + one preprocessing thread + one inference thread, queues of 10, put_nowait on the
                  request queue [125-146,161]; client_max_size 2**26 [148]
  config file     argv[1] JSON with device, model, model_name, max_batch_size, port [19-20,27-28,197]; here `model_path`
                  names the MSEW0001 weights container (mse_b200.weights) and the optional `tokenizer_model` a SentencePiece
                  model file (the reference obtains both through open_clip, clip_server.py:23-25).

The towers themselves run in libmse_b200.so (csrc/encoder.cu); nothing here computes an embedding on the CPU.
"""
from __future__ import annotations

import asyncio
import collections
import io
import json
import queue
import re
import string
import sys
import threading
import traceback

import msgpack
import numpy as np
from aiohttp import web
from prometheus_client import REGISTRY, CollectorRegistry, Counter, Histogram, generate_latest

InferenceParameters = collections.namedtuple("InferenceParameters", ["text", "images", "callback"])


class SiglipTokenizer:
    """open_clip's SigLIP text pipeline as misc/clip_accursed.py:55 records it: canonicalise (lower-case, strip punctuation,
    collapse whitespace), SentencePiece (c4_en, 32k), append EOS, pad/truncate to 64 with id 1 (eos = "sticky")."""

    def __init__(self, model_file: str, context_length: int = 64, pad_id: int = 1, eos_id: int = 1):
        import sentencepiece as spm
        self.sp = spm.SentencePieceProcessor(model_file=model_file)
        self.context_length, self.pad_id, self.eos_id = context_length, pad_id, eos_id
        self._punct = str.maketrans("", "", string.punctuation)

    def canonicalize(self, text: str) -> str:
        text = text.translate(self._punct).lower()
        return re.sub(r"\s+", " ", text).strip()

    def __call__(self, texts) -> np.ndarray:
        if isinstance(texts, str):
            texts = [texts]
        out = np.full((len(texts), self.context_length), self.pad_id, np.int32)
        for i, t in enumerate(texts):
            ids = list(self.sp.encode(self.canonicalize(t)))[: self.context_length - 1] + [self.eos_id]
            out[i, : len(ids)] = ids
        return out


def decode_image(data: bytes, size: int) -> np.ndarray:
    """PIL open -> RGB -> (size, size) u8 HWC.  The reference's clients already send size x size 24-bit BMPs
    (src/common.rs:31-54), for which open_clip's Resize is the identity; other sizes are squashed bicubically."""
    from PIL import Image
    im = Image.open(io.BytesIO(data)).convert("RGB")
    if im.size != (size, size):
        im = im.resize((size, size), Image.BICUBIC)
    return np.asarray(im, dtype=np.uint8)


class ClipServer:
    def __init__(self, config: dict, encoder=None, tokenizer=None, registry: CollectorRegistry | None = None):
        self.config = config
        self.BS = config["max_batch_size"]
        self.MODELNAME = config["model_name"]
        if encoder is None:
            from .encoder import Encoder
            dev = config.get("device", "cuda:0")
            if not str(dev).startswith("cuda"):
                raise RuntimeError(f'device "{dev}": this server only runs its towers on a B200 (no CPU path)')
            device_index = int(str(dev).split(":")[1]) if ":" in str(dev) else 0
            encoder = Encoder(config["model_path"], device=device_index, max_batch=self.BS)
        self.encoder = encoder
        if tokenizer is None and config.get("tokenizer_model"):
            tokenizer = SiglipTokenizer(config["tokenizer_model"], context_length=getattr(encoder, "ctx", 64))
        self.tokenizer = tokenizer
        reg = registry if registry is not None else REGISTRY
        self.registry = reg
        self.items_ctr = Counter("modelserver_total_items", "Items run through model server", ["model", "modality"], registry=reg)
        self.inference_time_hist = Histogram("modelserver_inftime", "Time running inference", ["model", "batch_size"], registry=reg)
        self.batch_count_ctr = Counter("modelserver_batchcount", "Inference batches run", ["model"], registry=reg)
        self.iq: queue.Queue = queue.Queue(10)
        self.pq: queue.Queue = queue.Queue(10)
        self.app = web.Application(client_max_size=2 ** 26)
        self.app.router.add_post("/", self.run_inference)
        self.app.router.add_get("/config", self.config_route)
        self.app.router.add_get("/", self.health)
        self.app.router.add_get("/metrics", self.metrics)
        self._threads = []

    # -- worker threads (clip_server.py:91-146) ------------------------------------------------------------------
    def do_inference(self, params: InferenceParameters):
        try:
            text, images, callback = params
            if text is not None:
                self.items_ctr.labels(self.MODELNAME, "text").inc(text.shape[0])
                with self.inference_time_hist.labels(self.MODELNAME + "-text", text.shape[0]).time():
                    features = self.encoder.encode_text(text)
            elif images is not None:
                with self.inference_time_hist.labels(self.MODELNAME + "-image", images.shape[0]).time():
                    self.items_ctr.labels(self.MODELNAME, "image").inc(images.shape[0])
                    features = self.encoder.encode_image(images)
            self.batch_count_ctr.labels(self.MODELNAME).inc()
            callback(True, features)
        except Exception as e:
            traceback.print_exc()
            callback(False, str(e))

    def infer_thread(self):
        while True:
            item = self.iq.get()
            if item is None:
                return
            self.do_inference(item)

    def preprocessing_thread(self):
        while True:
            item = self.pq.get()
            if item is None:
                self.iq.put(None)
                return
            text, images, callback = item
            try:
                if text:
                    if isinstance(text, str):
                        text = [text]
                    assert len(text) <= self.BS, f"max batch size is {self.BS}"
                    if self.tokenizer is None:
                        raise RuntimeError("text requests need `tokenizer_model` (a SentencePiece model file) in the config")
                    text = self.tokenizer(text)
                    images = None
                elif images:
                    assert len(images) <= self.BS, f"max batch size is {self.BS}"
                    size = getattr(self.encoder, "image_size", 384)
                    images = np.stack([decode_image(im, size) for im in images])
                    text = None
                else:
                    assert False, "images or text required"
                self.iq.put(InferenceParameters(text, images, callback))
            except Exception as e:
                traceback.print_exc()
                callback(False, str(e))

    # -- routes (clip_server.py:151-191) -------------------------------------------------------------------------
    async def run_inference(self, request):
        loop = asyncio.get_event_loop()
        data = msgpack.loads(await request.read())
        event = asyncio.Event()
        results = None

        def callback(*argv):
            nonlocal results
            results = argv
            loop.call_soon_threadsafe(lambda: event.set())

        self.pq.put_nowait(InferenceParameters(data.get("text"), data.get("images"), callback))
        await event.wait()
        body_data = results[1]
        if results[0]:
            status = 200
            body_data = [np.asarray(x).astype("float16").tobytes() for x in body_data]
        else:
            status = 500
            print(results[1])
        return web.Response(body=msgpack.dumps(body_data), status=status, content_type="application/msgpack")

    async def config_route(self, request):
        size = getattr(self.encoder, "image_size", 384)
        return web.Response(body=msgpack.dumps({
            "model": self.config["model"],
            "batch": self.BS,
            "image_size": (size, size),
            "embedding_size": getattr(self.encoder, "dim", 1152),
        }), status=200, content_type="application/msgpack")

    async def health(self, request):
        return web.Response(status=204)

    async def metrics(self, request):
        return web.Response(body=generate_latest(self.registry))

    # -- lifecycle -----------------------------------------------------------------------------------------------
    def start_threads(self):
        for fn in (self.infer_thread, self.preprocessing_thread):
            th = threading.Thread(target=fn, daemon=True)
            th.start()
            self._threads.append(th)

    def stop_threads(self):
        self.pq.put(None)
        for th in self._threads:
            th.join(timeout=5)

    async def run_webserver(self):
        runner = web.AppRunner(self.app)
        await runner.setup()
        site = web.TCPSite(runner, "", self.config["port"])
        print("Ready")
        await site.start()


def main(argv=None):
    argv = argv if argv is not None else sys.argv
    with open(argv[1], "r") as config_file:
        config = json.load(config_file)
    server = ClipServer(config)
    print("Model loaded")
    try:
        server.start_threads()
        loop = asyncio.new_event_loop()
        asyncio.set_event_loop(loop)
        loop.run_until_complete(server.run_webserver())
        loop.run_forever()
    except KeyboardInterrupt:
        sys.exit(0)


if __name__ == "__main__":
    main()
