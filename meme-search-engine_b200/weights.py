"""MSEW0001 weights container: the OpenCLIP / timm state_dict of ViT-SO400M-14-SigLIP-384 under its own tensor
names (the names clip_server.py:46-62 iterates over) plus an i32 ``config`` tensor.

    from mse_b200.weights import save_weights, config_for
    sd = open_clip_model.state_dict()            # where open_clip is available
    save_weights("siglip.msew", {k: v.numpy() for k, v in sd.items()}, config_for(sd))

Matrices are stored as fp16 (the reference runs precision="fp16", clip_server.py:23); vectors as fp32.
"""
from __future__ import annotations

import struct

import numpy as np

ACT_GELU_ERF, ACT_GELU_TANH = 1, 2
_DT = {np.dtype(np.float32): 0, np.dtype(np.float16): 1, np.dtype(np.int32): 2}


def make_config(image_size=384, patch=14, dim=1152, depth_v=27, heads=16, mlp=4304, vocab=32000, ctx=64, act=ACT_GELU_ERF,
                has_vision=True, has_text=True, depth_t=27) -> np.ndarray:
    c = np.zeros(16, np.int32)
    c[:12] = [image_size, patch, dim, depth_v, heads, mlp, vocab, ctx, act, int(has_vision), int(has_text), depth_t]
    return c


def config_for(sd: dict, act=ACT_GELU_ERF) -> np.ndarray:
    """Derive the config tensor from an OpenCLIP state_dict."""
    has_v = "visual.trunk.pos_embed" in sd
    has_t = "text.token_embedding.weight" in sd
    depth_v = len({k.split(".")[3] for k in sd if k.startswith("visual.trunk.blocks.")})
    depth_t = len({k.split(".")[3] for k in sd if k.startswith("text.transformer.resblocks.")})
    dim = int(np.shape(sd["visual.trunk.pos_embed"])[-1]) if has_v else int(np.shape(sd["text.token_embedding.weight"])[1])
    n_tok = int(np.shape(sd["visual.trunk.pos_embed"])[-2]) if has_v else 729
    grid = int(round(n_tok ** 0.5))
    mlp_key = "visual.trunk.blocks.0.mlp.fc1.weight" if has_v else "text.transformer.resblocks.0.mlp.c_fc.weight"
    return make_config(image_size=grid * 14 + 6 if grid == 27 else grid * 14, dim=dim, depth_v=depth_v, heads=dim // 72,
                       mlp=int(np.shape(sd[mlp_key])[0]),
                       vocab=int(np.shape(sd["text.token_embedding.weight"])[0]) if has_t else 32000,
                       ctx=int(np.shape(sd["text.positional_embedding"])[0]) if has_t else 64, act=act, has_vision=has_v,
                       has_text=has_t, depth_t=depth_t)


def save_weights(path: str, tensors: dict, config: np.ndarray):
    items = {"config": np.ascontiguousarray(config, np.int32)}
    for name, a in tensors.items():
        a = np.asarray(a)
        if a.dtype not in (np.float32, np.float16):
            a = a.astype(np.float32)
        if a.ndim >= 2 and a.dtype == np.float32:
            a = a.astype(np.float16)  # matrices / embeddings / conv kernels: fp16 like the reference's precision="fp16"
        items[name] = np.ascontiguousarray(a)
    with open(path, "wb") as f:
        f.write(b"MSEW0001")
        f.write(struct.pack("<I", len(items)))
        off = 12
        for name, a in items.items():
            nb = name.encode()
            hdr = struct.pack("<H", len(nb)) + nb + struct.pack("<BB", _DT[a.dtype], a.ndim) + struct.pack(f"<{a.ndim}I", *a.shape) + struct.pack("<Q", a.nbytes)
            f.write(hdr)
            off += len(hdr)
            pad = (-off) % 16
            f.write(b"\0" * pad)
            off += pad
            f.write(a.tobytes())
            off += a.nbytes
