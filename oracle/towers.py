"""fp32 CPU oracle of the SigLIP towers -- TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline).

PARITY UNPINNED: the reference computes the towers inside open-clip-torch (>=2.31,<2.32; requirements.txt:4) -> timm
-> torch, none of which (nor the checkpoint, nor the tokenizer model) exist in this image, and it ships no golden
vectors for them.  The stand-in is ``transformers.SiglipVisionModel`` / ``SiglipTextModel`` at SO400M dimensions with
``hidden_act="gelu"`` (erf, as aitemplate/model.py:18 fixes it) and SEEDED RANDOM weights; the same math as
timm/OpenCLIP for this architecture up to a re-layout of the pooling head.  Matrices are rounded to
fp16 before use on both sides (the reference runs precision="fp16", clip_server.py:23), so what is compared is the
arithmetic, not the weight rounding.

export_openclip() renames the HF tensors to the OpenCLIP/timm state_dict names clip_server.py:46-62 iterates over --
the names the product's weights container uses.
"""
from __future__ import annotations

import numpy as np
import torch

D, HEADS, MLP, IMG, PATCH, VOCAB, CTX = 1152, 16, 4304, 384, 14, 32000, 64


def _seeded_init(model: torch.nn.Module, seed: int):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.ndim >= 2:
                if "position_embedding" in name or "probe" in name:
                    p.copy_(torch.randn(p.shape, generator=g) * (1.0 / np.sqrt(D)))
                elif "token_embedding" in name:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.05)
                elif "patch_embedding" in name:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.03)
                else:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.02)
                p.copy_(p.to(torch.float16).to(torch.float32))  # matrices live in fp16 on both sides
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.05 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g))


def build_vision(depth: int = 27, seed: int = 42, act: str = "gelu"):
    from transformers import SiglipVisionConfig, SiglipVisionModel
    cfg = SiglipVisionConfig(hidden_size=D, intermediate_size=MLP, num_hidden_layers=depth, num_attention_heads=HEADS,
                             image_size=IMG, patch_size=PATCH, hidden_act=act, layer_norm_eps=1e-6)
    cfg._attn_implementation = "eager"
    m = SiglipVisionModel(cfg).eval()
    _seeded_init(m, seed)
    return m


def build_text(depth: int = 27, seed: int = 43, act: str = "gelu"):
    from transformers import SiglipTextConfig, SiglipTextModel
    cfg = SiglipTextConfig(vocab_size=VOCAB, hidden_size=D, intermediate_size=MLP, num_hidden_layers=depth,
                           num_attention_heads=HEADS, max_position_embeddings=CTX, hidden_act=act, layer_norm_eps=1e-6,
                           projection_size=D)
    cfg._attn_implementation = "eager"
    m = SiglipTextModel(cfg).eval()
    _seeded_init(m, seed)
    return m


def preprocess_u8(images_u8: np.ndarray) -> torch.Tensor:
    """clip_server.py:140: ToTensor, Normalize(mean=std=0.5) -> [-1,1], .half(); resize is an identity for 384x384 input."""
    x = torch.from_numpy(np.ascontiguousarray(images_u8)).to(torch.float32)
    x = (x / 255.0 - 0.5) / 0.5
    x = x.to(torch.float16).to(torch.float32)
    return x.permute(0, 3, 1, 2).contiguous()


@torch.no_grad()
def encode_image(model, images_u8: np.ndarray, hidden_states: bool = False):
    """-> unit-norm fp32 [B, D] (clip_server.py:114-116), optionally with the per-block token activations."""
    out = model(pixel_values=preprocess_u8(images_u8), output_hidden_states=hidden_states)
    f = out.pooler_output
    f = f / f.norm(dim=-1, keepdim=True)
    if hidden_states:
        return f.numpy(), [h.numpy() for h in out.hidden_states]
    return f.numpy()


@torch.no_grad()
def encode_text(model, ids: np.ndarray, hidden_states: bool = False):
    """-> unit-norm fp32 [B, D] (clip_server.py:98-100).  No padding mask: SigLIP attends to pad tokens."""
    out = model(input_ids=torch.from_numpy(np.ascontiguousarray(ids)).long(), output_hidden_states=hidden_states)
    f = out.pooler_output
    f = f / f.norm(dim=-1, keepdim=True)
    if hidden_states:
        return f.numpy(), [h.numpy() for h in out.hidden_states]
    return f.numpy()


def export_openclip(vision=None, text=None) -> dict:
    """HF state_dict -> OpenCLIP/timm names (clip_server.py:46-62)."""
    sd = {}
    if vision is not None:
        v = {k.replace("vision_model.", ""): t.detach().numpy() for k, t in vision.state_dict().items()}
        sd["visual.trunk.patch_embed.proj.weight"] = v["embeddings.patch_embedding.weight"]
        sd["visual.trunk.patch_embed.proj.bias"] = v["embeddings.patch_embedding.bias"]
        sd["visual.trunk.pos_embed"] = v["embeddings.position_embedding.weight"][None]
        n = len({k.split(".")[2] for k in v if k.startswith("encoder.layers.")})
        for i in range(n):
            s, d = f"encoder.layers.{i}.", f"visual.trunk.blocks.{i}."
            sd[d + "norm1.weight"], sd[d + "norm1.bias"] = v[s + "layer_norm1.weight"], v[s + "layer_norm1.bias"]
            sd[d + "attn.qkv.weight"] = np.concatenate([v[s + f"self_attn.{x}_proj.weight"] for x in "qkv"], 0)
            sd[d + "attn.qkv.bias"] = np.concatenate([v[s + f"self_attn.{x}_proj.bias"] for x in "qkv"], 0)
            sd[d + "attn.proj.weight"], sd[d + "attn.proj.bias"] = v[s + "self_attn.out_proj.weight"], v[s + "self_attn.out_proj.bias"]
            sd[d + "norm2.weight"], sd[d + "norm2.bias"] = v[s + "layer_norm2.weight"], v[s + "layer_norm2.bias"]
            for fc in ("fc1", "fc2"):
                sd[d + f"mlp.{fc}.weight"], sd[d + f"mlp.{fc}.bias"] = v[s + f"mlp.{fc}.weight"], v[s + f"mlp.{fc}.bias"]
        sd["visual.trunk.norm.weight"], sd["visual.trunk.norm.bias"] = v["post_layernorm.weight"], v["post_layernorm.bias"]
        p = "visual.trunk.attn_pool."
        sd[p + "latent"] = v["head.probe"]
        w, b = v["head.attention.in_proj_weight"], v["head.attention.in_proj_bias"]
        sd[p + "q.weight"], sd[p + "q.bias"] = w[:D], b[:D]
        sd[p + "kv.weight"], sd[p + "kv.bias"] = w[D:], b[D:]
        sd[p + "proj.weight"], sd[p + "proj.bias"] = v["head.attention.out_proj.weight"], v["head.attention.out_proj.bias"]
        sd[p + "norm.weight"], sd[p + "norm.bias"] = v["head.layernorm.weight"], v["head.layernorm.bias"]
        for fc in ("fc1", "fc2"):
            sd[p + f"mlp.{fc}.weight"], sd[p + f"mlp.{fc}.bias"] = v[f"head.mlp.{fc}.weight"], v[f"head.mlp.{fc}.bias"]
    if text is not None:
        t = {k.replace("text_model.", ""): x.detach().numpy() for k, x in text.state_dict().items()}
        sd["text.token_embedding.weight"] = t["embeddings.token_embedding.weight"]
        sd["text.positional_embedding"] = t["embeddings.position_embedding.weight"]
        n = len({k.split(".")[2] for k in t if k.startswith("encoder.layers.")})
        for i in range(n):
            s, d = f"encoder.layers.{i}.", f"text.transformer.resblocks.{i}."
            sd[d + "ln_1.weight"], sd[d + "ln_1.bias"] = t[s + "layer_norm1.weight"], t[s + "layer_norm1.bias"]
            sd[d + "attn.in_proj_weight"] = np.concatenate([t[s + f"self_attn.{x}_proj.weight"] for x in "qkv"], 0)
            sd[d + "attn.in_proj_bias"] = np.concatenate([t[s + f"self_attn.{x}_proj.bias"] for x in "qkv"], 0)
            sd[d + "attn.out_proj.weight"], sd[d + "attn.out_proj.bias"] = t[s + "self_attn.out_proj.weight"], t[s + "self_attn.out_proj.bias"]
            sd[d + "ln_2.weight"], sd[d + "ln_2.bias"] = t[s + "layer_norm2.weight"], t[s + "layer_norm2.bias"]
            sd[d + "mlp.c_fc.weight"], sd[d + "mlp.c_fc.bias"] = t[s + "mlp.fc1.weight"], t[s + "mlp.fc1.bias"]
            sd[d + "mlp.c_proj.weight"], sd[d + "mlp.c_proj.bias"] = t[s + "mlp.fc2.weight"], t[s + "mlp.fc2.bias"]
        sd["text.ln_final.weight"], sd["text.ln_final.bias"] = t["final_layer_norm.weight"], t["final_layer_norm.bias"]
        sd["text.text_projection.weight"], sd["text.text_projection.bias"] = t["head.weight"], t["head.bias"]
    return sd


def synthetic_images(seed: int, n: int) -> np.ndarray:
    """uniform u8 noise with a smooth component so LayerNorm inputs are not degenerate."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (n, IMG // 16, IMG // 16, 3)).astype(np.float32)
    up = np.repeat(np.repeat(base, 16, axis=1), 16, axis=2)
    noise = rng.integers(-40, 41, (n, IMG, IMG, 3)).astype(np.float32)
    return np.clip(up + noise, 0, 255).astype(np.uint8)


def synthetic_token_ids(seed: int, n: int) -> np.ndarray:
    """random ids in [2, VOCAB) of random length 3..16, then EOS/pad id 1 up to 64."""
    rng = np.random.default_rng(seed)
    ids = np.ones((n, CTX), np.int32)
    for i in range(n):
        L = int(rng.integers(3, 17))
        ids[i, :L] = rng.integers(2, VOCAB, L)
    return ids
