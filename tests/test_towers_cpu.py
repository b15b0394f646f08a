"""CPU tests of the tower oracle and the weights container (no GPU)."""
import os
import struct

import numpy as np
import pytest

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def towers():
    from oracle import towers as T
    return T


@pytest.fixture(scope="module")
def models(towers):
    return towers.build_vision(depth=2, seed=42), towers.build_text(depth=2, seed=43)


def test_oracle_reproduces_golden(towers, models):
    g = np.load(os.path.join(G, "towers_depth2.npz"))
    v, t = models
    fi, hi = towers.encode_image(v, towers.synthetic_images(1, 2), hidden_states=True)
    ft, ht = towers.encode_text(t, towers.synthetic_token_ids(1, 3), hidden_states=True)
    assert np.allclose(fi, g["image_features"], atol=2e-5) and np.allclose(ft, g["text_features"], atol=2e-5)
    assert np.allclose(hi[2][0, 0], g["image_token0_block2"], atol=1e-3)
    assert np.allclose(ht[2][0, -1], g["text_last_block2"], atol=1e-3)
    assert np.allclose(np.linalg.norm(fi, axis=1), 1, atol=1e-5)


def test_oracle_matches_plain_numpy_block(towers, models):
    """The stand-in follows aitemplate/model.py:26-55: x + proj(MHA(LN1 x)); x + fc2(gelu_erf(fc1(LN2 x)))."""
    import torch
    from math import erf
    v, _ = models
    sd = towers.export_openclip(vision=v)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 1152)).astype(np.float64)

    def ln(a, g, b):
        m, s = a.mean(-1, keepdims=True), a.var(-1, keepdims=True)
        return (a - m) / np.sqrt(s + 1e-6) * g + b

    p = "visual.trunk.blocks.0."
    h = ln(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    qkv = h @ sd[p + "attn.qkv.weight"].astype(np.float64).T + sd[p + "attn.qkv.bias"]
    q, k, vv = [a.reshape(5, 16, 72).transpose(1, 0, 2) for a in np.split(qkv, 3, axis=1)]
    s = q @ k.transpose(0, 2, 1) / np.sqrt(72)
    s = np.exp(s - s.max(-1, keepdims=True)); s /= s.sum(-1, keepdims=True)
    a = (s @ vv).transpose(1, 0, 2).reshape(5, 1152)
    x1 = x + a @ sd[p + "attn.proj.weight"].astype(np.float64).T + sd[p + "attn.proj.bias"]
    h = ln(x1, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    f = h @ sd[p + "mlp.fc1.weight"].astype(np.float64).T + sd[p + "mlp.fc1.bias"]
    f = 0.5 * f * (1 + np.vectorize(erf)(f / np.sqrt(2)))
    x2 = x1 + f @ sd[p + "mlp.fc2.weight"].astype(np.float64).T + sd[p + "mlp.fc2.bias"]
    with torch.no_grad():
        ref = v.vision_model.encoder.layers[0](torch.from_numpy(x[None]).float(), attention_mask=None)
        ref = (ref[0] if isinstance(ref, tuple) else ref)[0].numpy()
    assert np.abs(ref - x2).max() < 2e-4


def test_export_names_and_container_roundtrip(towers, models, tmp_path, mse):
    v, t = models
    sd = towers.export_openclip(v, t)
    for k in ("visual.trunk.patch_embed.proj.weight", "visual.trunk.pos_embed", "visual.trunk.blocks.1.attn.qkv.weight",
              "visual.trunk.attn_pool.latent", "visual.trunk.attn_pool.kv.weight", "visual.trunk.norm.bias",
              "text.token_embedding.weight", "text.transformer.resblocks.1.attn.in_proj_weight", "text.text_projection.bias"):
        assert k in sd, k
    assert sd["visual.trunk.blocks.0.attn.qkv.weight"].shape == (3456, 1152)
    assert sd["visual.trunk.pos_embed"].shape == (1, 729, 1152)
    cfg = mse.weights.config_for(sd)
    assert cfg[:12].tolist() == [384, 14, 1152, 2, 16, 4304, 32000, 64, 1, 1, 1, 2]
    path = str(tmp_path / "w.msew")
    mse.weights.save_weights(path, sd, cfg)
    blob = open(path, "rb").read()
    assert blob[:8] == b"MSEW0001"
    n = struct.unpack_from("<I", blob, 8)[0]
    assert n == len(sd) + 1
    off, seen = 12, {}
    for _ in range(n):
        nl = struct.unpack_from("<H", blob, off)[0]; off += 2
        name = blob[off:off + nl].decode(); off += nl
        dt, nd = struct.unpack_from("<BB", blob, off); off += 2
        dims = struct.unpack_from(f"<{nd}I", blob, off); off += 4 * nd
        nb = struct.unpack_from("<Q", blob, off)[0]; off += 8
        off = (off + 15) // 16 * 16
        seen[name] = (dt, dims, off)
        off += nb
    assert off == len(blob)
    dt, dims, o = seen["visual.trunk.blocks.1.mlp.fc2.weight"]
    assert dt == 1 and dims == (1152, 4304)
    w = np.frombuffer(blob, np.float16, 1152 * 4304, o).reshape(1152, 4304)
    assert np.array_equal(w.astype(np.float32), sd["visual.trunk.blocks.1.mlp.fc2.weight"])  # matrices were fp16-exact already
    assert seen["config"][0] == 2 and seen["visual.trunk.norm.bias"][0] == 0
