"""Query assembly around the search kernels, restated from the reference's front-ends.

  get_total_embedding   src/common.rs:215-274          q = sum_t weight_t * embedding_t over text / image / raw / predefined
                                                       terms, f32, NOT renormalised
  select_shard          src/query_disk_index.rs:254-256,447-450   argmax_s trunc(2^32 * <centroid_s, q>) -> that shard's medioid
  decode_fp16_buffer    src/common.rs:98-102
"""
from __future__ import annotations

import base64

import numpy as np


def decode_fp16_buffer(buf: bytes) -> np.ndarray:
    return np.frombuffer(buf, dtype="<f2").astype(np.float32)


def get_total_embedding(terms, embedding_size: int, query_server, resize_image=None, predefined_embeddings=None) -> np.ndarray:
    """terms: iterable of dicts with optional keys image (base64 str), text, embedding (list of f32), predefined_embedding,
    weight.  query_server({"images": [...]}) / query_server({"text": [...]}) returns a list of fp16-LE byte strings, like
    the clip_server boundary.  resize_image(bytes) -> bytes of the 384x384 BMP the server expects (common.rs:31-54)."""
    total = np.zeros(embedding_size, np.float32)
    image_batch, image_weights, text_batch, text_weights = [], [], [], []
    predefined_embeddings = predefined_embeddings or {}
    for term in terms:
        w = np.float32(term.get("weight") if term.get("weight") is not None else 1.0)
        if term.get("image") is not None:
            raw = base64.standard_b64decode(term["image"])
            image_batch.append(resize_image(raw) if resize_image else raw)
            image_weights.append(w)
        if term.get("text") is not None:
            text_batch.append(term["text"])
            text_weights.append(w)
        if term.get("embedding") is not None:
            e = np.asarray(term["embedding"], np.float32)
            total[: e.size] += e * w
        if term.get("predefined_embedding") is not None and term["predefined_embedding"] in predefined_embeddings:
            total = total + np.asarray(predefined_embeddings[term["predefined_embedding"]], np.float32) * w
    batches = []
    if image_batch:
        batches.append(({"images": image_batch}, image_weights))
    if text_batch:
        batches.append(({"text": text_batch}, text_weights))
    for batch, weights in batches:
        for emb, w in zip(query_server(batch), weights):
            total += decode_fp16_buffer(emb) * w
    return total


def select_shard(shards, query: np.ndarray) -> int:
    """shards: list of (centroid f32[d], medioid id) as in IndexHeader.shards (common.rs:167-174).  Returns the index of the
    shard whose centroid has the largest scaled dot product with the query; Iterator::position_max_by_key keeps the LAST
    maximum."""
    best, best_i = None, 0
    q = np.asarray(query, np.float32)
    for i, (centroid, _medioid) in enumerate(shards):
        v = float(np.dot(np.asarray(centroid, np.float32).astype(np.float64), q.astype(np.float64))) * 4294967296.0
        k = 0 if v != v else int(max(min(v, 9.223372036854775807e18), -9.223372036854775808e18))
        if best is None or k >= best:
            best, best_i = k, i
    return best_i
