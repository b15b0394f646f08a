#!/usr/bin/env python
"""bench.py -- one JSON line per run (contract: the task prompt's bench section).

Headline workload (config.workload = "siglip_image_tower_b256"): BASELINE.json configs[1] -- the SigLIP
ViT-SO400M-14/384 image tower, batch 256, fp16 storage / fp32 accumulate, one B200; a step = one batch of 256
synthetic 384x384 u8 images through im2col+normalise, patch-embed GEMM, 27 blocks (LN, QKV GEMM, attention,
out-proj GEMM, LN, MLP GEMMs), final LN, MAP head and L2 normalisation.  Random-init weights of that architecture.
For --gpus N each rank encodes its own batch of 256 (data parallel, weights replicated, no collective): weak scaling.

  value  images/s with the u8 images resident in HBM (device pointers into the C ABI)
  e2e    images/s through the host-pointer C-ABI call: pinned host u8 images -> H2D -> towers -> D2H fp16 features

The same run also measures BASELINE.json configs[2] (1024 f32 queries, top-100 over a 10M x 1152 fp16 index,
id-range sharded over the ranks, strong scaling, one all-gather + merge) and reports it under "search".

--impl reference times the CPU stand-ins of the reference's path on the host cores (bounded samples): transformers
SiglipVisionModel fp32 for clip_server.py device=cpu, and the C restatement of the flat scan (oracle/, AVX2+OpenMP).
The reference itself (open_clip + Rust nightly + faiss) cannot be installed or built in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

D = 1152
METRIC = "SigLIP ViT-SO400M-14/384 images/sec (image tower, batch 256 fp16)"
GRAPH_METRIC = "queries/sec@recall10 (Vamana graph search, 4096 batched queries, L=64)"
SEARCH_METRIC = "queries/sec (flat top-100, 1024 queries x 10M x 1152 fp16 index)"
FLOP_PER_IMAGE = 670.35e9  # 27 layers 665.46 + patch-embed 0.99 + MAP head 3.90 GFLOP
# DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernels, taken from the committed `ncu --set full`
# captures -- a number measured under a profiler on an earlier run of the same configuration, not by this run.
NCU_TRAFFIC = {
    # profiles/r01b_tower_ncu_full.md: one block at batch 256 = fc1 2.065 + fc2 3.454 + QKV 1.682 + out-proj 1.262 GB; x 27 blocks
    "tower_gemm_bytes_per_step_b256": 27 * (2.065 + 3.454 + 1.682 + 1.262) * 1e9,
    # profiles/r01_flat_gemm_ncu_full.md: 3.745 GB read + 5.6 MB written for the 3.74 GB of rows the launch scored
    "flat_gemm_bytes_per_row_byte": (3.745176 + 0.005563) / 3.74,
    # profiles/r01i_greedy_1m_ncu_full.md: 14.376 GB read + 0.652 GB written for 15.00 GB of gathered rows (1 M rows, 4096 queries, L = 64)
    "greedy_bytes_per_row_byte": (14.376188 + 0.651920) / (1589.4404296875 * 4096 * 2304 / 1e9),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm": j["hbm_gbs"], "tf_burst": j["bf16_tflops"], "tf_sustained": j.get("bf16_tflops_sustained", j["bf16_tflops"]),
                "src": "measured"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self._stop, self._t = gpu_index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, pw, reasons = [], 0.0, 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1])); pw = max(pw, float(r[2]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "power_w_max": pw or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------- CPU stand-ins of the reference (bounded samples)

def cpu_reference_tower(batch: int, steps: int, warmup: int):
    import numpy as np
    import torch
    from oracle import towers as T
    torch.set_num_threads(os.cpu_count() or 1)
    m = T.build_vision(depth=27, seed=42)
    imgs = T.synthetic_images(1, batch)
    for _ in range(warmup):
        T.encode_image(m, imgs[:1])
    t0 = time.perf_counter()
    for _ in range(steps):
        T.encode_image(m, imgs)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return batch / dt, dt, os.cpu_count()


def cpu_reference_flat(rows_total: int, k: int, steps: int, warmup: int, sample_rows: int, sample_q: int):
    """fp16 rows x f32 query, f32 accumulate (AVX2), top-k heap, OpenMP over rows (oracle mode 1); a scan is linear in rows,
    so throughput on a sample_rows slice is scaled by sample_rows / rows_total."""
    import numpy as np
    from oracle import oracle as O
    O.build()
    rng = np.random.default_rng(2)
    x = rng.standard_normal((sample_rows, D), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    x16 = x.astype(np.float16)
    del x
    q = np.random.default_rng(3).standard_normal((sample_q, D), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    for _ in range(warmup):
        O.flat_search(q[:2], x16, k, mode=1)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.flat_search(q, x16, k, mode=1)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return (sample_q / dt) * sample_rows / rows_total, dt, os.cpu_count()


# ---------------------------------------------------------------- synthetic weights (random init of the named architecture)

def random_openclip_state_dict(dev, depth_v=27, seed=42):
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    F = 4304

    def mat(*shape, std=0.02):
        return (torch.randn(shape, generator=g, device=dev) * std).to(torch.float16).cpu().numpy()

    def vec(n, base=0.0, std=0.02):
        return (base + std * torch.randn((n,), generator=g, device=dev)).float().cpu().numpy()

    sd = {"visual.trunk.patch_embed.proj.weight": mat(D, 3, 14, 14, std=0.03), "visual.trunk.patch_embed.proj.bias": vec(D),
          "visual.trunk.pos_embed": mat(1, 729, D, std=D ** -0.5)}
    for i in range(depth_v):
        p = f"visual.trunk.blocks.{i}."
        sd[p + "norm1.weight"], sd[p + "norm1.bias"] = vec(D, 1.0, 0.05), vec(D)
        sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"] = mat(3 * D, D), vec(3 * D)
        sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"] = mat(D, D), vec(D)
        sd[p + "norm2.weight"], sd[p + "norm2.bias"] = vec(D, 1.0, 0.05), vec(D)
        sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"] = mat(F, D), vec(F)
        sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"] = mat(D, F), vec(D)
    sd["visual.trunk.norm.weight"], sd["visual.trunk.norm.bias"] = vec(D, 1.0, 0.05), vec(D)
    p = "visual.trunk.attn_pool."
    sd[p + "latent"] = mat(1, 1, D, std=D ** -0.5)
    sd[p + "q.weight"], sd[p + "q.bias"] = mat(D, D), vec(D)
    sd[p + "kv.weight"], sd[p + "kv.bias"] = mat(2 * D, D), vec(2 * D)
    sd[p + "proj.weight"], sd[p + "proj.bias"] = mat(D, D), vec(D)
    sd[p + "norm.weight"], sd[p + "norm.bias"] = vec(D, 1.0, 0.05), vec(D)
    sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"] = mat(F, D), vec(F)
    sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"] = mat(D, F), vec(D)
    return sd


def random_openclip_text_state_dict(dev, depth_t=27, seed=43, vocab=32000, ctx=64):
    """Random init of the text tower (OpenCLIP TextTransformer names, clip_server.py:98)."""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    F = 4304

    def mat(*shape, std=0.02):
        return (torch.randn(shape, generator=g, device=dev) * std).to(torch.float16).cpu().numpy()

    def vec(n, base=0.0, std=0.02):
        return (base + std * torch.randn((n,), generator=g, device=dev)).float().cpu().numpy()

    sd = {"text.token_embedding.weight": mat(vocab, D), "text.positional_embedding": mat(ctx, D, std=D ** -0.5)}
    for i in range(depth_t):
        p = f"text.transformer.resblocks.{i}."
        sd[p + "ln_1.weight"], sd[p + "ln_1.bias"] = vec(D, 1.0, 0.05), vec(D)
        sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"] = mat(3 * D, D), vec(3 * D)
        sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"] = mat(D, D), vec(D)
        sd[p + "ln_2.weight"], sd[p + "ln_2.bias"] = vec(D, 1.0, 0.05), vec(D)
        sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"] = mat(F, D), vec(F)
        sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"] = mat(D, F), vec(D)
    sd["text.ln_final.weight"], sd["text.ln_final.bias"] = vec(D, 1.0, 0.05), vec(D)
    sd["text.text_projection.weight"], sd["text.text_projection.bias"] = mat(D, D), vec(D)
    return sd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workloads", default="tower,flat,graph", help="comma list of tower, flat, graph, e2e (e2e = BASELINE configs[4] at reduced scale; not in the default set)")
    ap.add_argument("--graph-rows", type=int, default=1_000_000, help="graph: total index rows (all shards); C4's 1e8 needs 8 GPUs x 12.5M and a long build")
    ap.add_argument("--graph-queries", type=int, default=4096)
    ap.add_argument("--graph-L", type=int, default=64, help="search list size (C4 'beam 64')")
    ap.add_argument("--graph-W", type=int, default=4, help="beam width of the compressed traversal")
    ap.add_argument("--cpu-sample-graph-queries", type=int, default=2048)
    ap.add_argument("--e2e-images", type=int, default=32768, help="e2e workload (not in the default set): images per rank to encode and index")
    ap.add_argument("--graph-data", default="mixture", choices=["mixture", "latent"],
                    help="mixture: SURVEY 8d C4 (4096 Gaussians, sigma 0.3); latent: unit rows on a 48-dimensional latent subspace + 5 %% noise "
                         "(neighbourhoods a 64-byte code can resolve; used to judge the compressed traversal)")
    ap.add_argument("--batch", type=int, default=256, help="images per rank per step")
    ap.add_argument("--rows", type=int, default=10_000_000, help="flat: total index rows (all shards)")
    ap.add_argument("--queries", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--cpu-sample-rows", type=int, default=500_000)
    ap.add_argument("--cpu-sample-queries", type=int, default=32)
    ap.add_argument("--cpu-sample-images", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    steps, warmup = args.steps, max(args.warmup, 0)
    wl = [w for w in args.workloads.split(",") if w]
    tower_cfg = {"workload": "siglip_image_tower_b256", "tower": "ViT-SO400M-14-SigLIP-384 image tower (random init)", "batch_per_gpu": args.batch,
                 "image": "384x384x3 u8", "storage": "fp16", "accumulate": "fp32", "parallelism": f"dp{world}",
                 "l2_policy": "inputs larger than L2 (per-layer activations 430 MB - 1.6 GB >> 126 MB)"}
    flat_cfg = {"workload": "flat_top100", "queries": args.queries, "k": args.k, "index_rows": args.rows, "dim": D, "index_dtype": "fp16",
                "query_dtype": "f32", "sharding": f"id-range x{world}", "l2_policy": "inputs larger than L2 (index shard >> 126 MB)"}

    if args.impl == "reference":
        if rank != 0:
            return
        ips, dt, cores = cpu_reference_tower(args.cpu_sample_images, max(1, min(steps, 2)), min(warmup, 1))
        sample = f"{args.cpu_sample_images} images per step (full 27-block tower, transformers SiglipVisionModel fp32, torch threads = all cores)"
        out = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
               "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32 (CPU)",
               "data": "synthetic", "config": tower_cfg,
               "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
               "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        if "flat" in wl:
            qps, fdt, _ = cpu_reference_flat(args.rows, args.k, 1, 1, args.cpu_sample_rows, args.cpu_sample_queries)
            out["search"] = {"metric": SEARCH_METRIC, "value": qps, "unit": "queries/s", "ms_per_step": fdt * 1e3, "config": flat_cfg,
                             "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                                              "sample": f"{args.cpu_sample_queries} queries x {args.cpu_sample_rows}-row slice, scaled by rows to {args.rows}"}}
        print(json.dumps(out))
        return

    import numpy as np
    import torch
    import mse_b200

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        if "MSE_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["MSE_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)   # NCCL_DEBUG=WARN/VERSION prints a version banner on stdout; stdout carries the one JSON line
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream().cuda_stream
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n_warm, n_steps):
        for _ in range(n_warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = mse_b200.launch_count()
        e0.record()
        for _ in range(n_steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = mse_b200.launch_count() - launches0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / n_steps, launches

    result = {}

    # ============================================================ tower (headline)
    if "tower" in wl:
        B = args.batch
        wpath = os.path.join(tempfile.gettempdir(), f"mse_bench_vision27_rank{rank}.msew")
        sd = random_openclip_state_dict(dev)
        mse_b200.weights.save_weights(wpath, sd, mse_b200.weights.config_for(sd))
        del sd
        enc = mse_b200.Encoder(wpath, device=local_rank, max_batch=B)
        os.remove(wpath)
        g = torch.Generator(device=dev).manual_seed(1 + rank)
        imgs_dev = torch.randint(0, 256, (B, 384, 384, 3), generator=g, device=dev, dtype=torch.uint8)
        imgs_host = imgs_dev.cpu().pin_memory()
        feat_dev = torch.empty((B, D), dtype=torch.float16, device=dev)
        feat_host = torch.empty((B, D), dtype=torch.float16).pin_memory()

        def tower_resident():
            enc.encode_image_dev(imgs_dev.data_ptr(), B, feat_dev.data_ptr(), stream)

        def tower_e2e():
            mse_b200.check(mse_b200.lib().mse_encode_images_u8(enc._h, imgs_host.data_ptr(), B, feat_host.data_ptr()), "mse_encode_images_u8")

        with ClockSampler(local_rank) as cs:
            ms_step, launches = timed(tower_resident, warmup, steps)
        clocks = cs.summary()
        ms_e2e, _ = timed(tower_e2e, min(warmup, 2), steps)
        enc.profile(True)
        tower_resident()
        st = enc.stats()
        enc.profile(False)
        torch.cuda.synchronize()
        f = feat_dev.float()
        assert torch.isfinite(f).all() and (f.norm(dim=1) - 1).abs().max() < 5e-3, "tower output is not unit-norm finite"
        gemm_tf = st["gemm_mflop"] * 1e6 / (st["gemm_ns"] * 1e-9) / 1e12 if st["gemm_ns"] else None
        attn_flop = 27 * 4.0 * 729 * 729 * 72 * 16 * B
        roof = {"kernel": "k_gemm_tn<BN,LinearEpilogue> (all GEMM launches of one step)", "bound": "tensor", "achieved": gemm_tf,
                "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": gemm_tf / pk["tf_sustained"] if gemm_tf else None,
                "peak_source": pk["src"] + " (sustained bf16 cuBLAS)", "frac_of_burst": gemm_tf / pk["tf_burst"] if gemm_tf else None,
                "launches_per_step": st["gemm_launches"], "kernel_ms_per_step": st["gemm_ns"] * 1e-6,
                "kernel_share_of_step": st["gemm_ns"] * 1e-6 / ms_step,
                "traffic": NCU_TRAFFIC["tower_gemm_bytes_per_step_b256"] * B / 256 / max(st["gemm_launches"], 1),
                "traffic_note": "bytes per launch (mean over the step's GEMM launches), from profiles/r01b_tower_ncu_full.md x 27 blocks; algorithmic A+B+C "
                                "(+ residual) bytes are 7.55 GB per block vs 8.46 GB measured",
                "attention": {"kernel": "k_mha_tc (tcgen05 flash attention)", "ms_per_step": st["attn_ns"] * 1e-6, "launches": st["attn_launches"],
                              "achieved_tflops": attn_flop / (st["attn_ns"] * 1e-9) / 1e12 if st["attn_ns"] else None,
                              "share_of_step": st["attn_ns"] * 1e-6 / ms_step},
                "whole_step": {"achieved_tflops": FLOP_PER_IMAGE * B / (ms_step * 1e-3) / 1e12,
                               "frac": FLOP_PER_IMAGE * B / (ms_step * 1e-3) / 1e12 / pk["tf_sustained"]}}
        result = {"metric": METRIC, "value": world * B / (ms_step * 1e-3), "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                  "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                  "dtype": "fp16 storage, fp32 accumulate (tcgen05 kind::f16)", "data": "synthetic", "config": tower_cfg, "clocks": clocks,
                  "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": B * 384 * 384 * 3, "d2h_bytes_per_step": B * D * 2,
                          "ms_per_step": ms_e2e},
                  "gpu_launches": launches, "roofline": roof}
        enc.close()
        del imgs_dev, feat_dev
        torch.cuda.empty_cache()

    # ============================================================ flat search
    if "flat" in wl:
        rows_total, nq, k = args.rows, args.queries, args.k
        row_lo, row_hi = rows_total * rank // world, rows_total * (rank + 1) // world
        n_local = row_hi - row_lo
        ix = mse_b200.FlatIndex(D, device=local_rank, id_base=row_lo)
        ix.reserve(n_local)
        gen = torch.Generator(device=dev)
        chunk = 1 << 19
        for c0 in range(row_lo, row_hi, chunk):
            m = min(chunk, row_hi - c0)
            gen.manual_seed(2 * 1_000_003 + c0)
            xb = torch.randn((m, D), generator=gen, device=dev, dtype=torch.float32)
            xb = (xb / xb.norm(dim=1, keepdim=True)).to(torch.float16).contiguous()
            ix.add_f16_dev(xb.data_ptr(), m, stream)
            del xb
        gq = torch.Generator(device="cpu").manual_seed(3)
        q_host = torch.randn((nq, D), generator=gq, dtype=torch.float32)
        q_host = (q_host / q_host.norm(dim=1, keepdim=True)).contiguous().pin_memory()
        q_dev = q_host.to(dev)
        ids_dev = torch.empty((nq, k), dtype=torch.int32, device=dev)
        sc_dev = torch.empty((nq, k), dtype=torch.float32, device=dev)
        if world > 1:
            all_ids = torch.empty((world, nq, k), dtype=torch.int32, device=dev)
            all_sc = torch.empty((world, nq, k), dtype=torch.float32, device=dev)
            out_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
            out_sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
        ids_host = torch.empty((nq, k), dtype=torch.int32).pin_memory()
        sc_host = torch.empty((nq, k), dtype=torch.float32).pin_memory()

        def flat_resident():
            ix.search_dev(q_dev.data_ptr(), nq, k, ids_dev.data_ptr(), sc_dev.data_ptr(), stream)
            if world > 1:
                dist.all_gather_into_tensor(all_ids.view(-1, k), ids_dev)
                dist.all_gather_into_tensor(all_sc.view(-1, k), sc_dev)
                mse_b200.merge_topk(local_rank, all_ids.data_ptr(), all_sc.data_ptr(), world, nq, k, out_ids.data_ptr(), out_sc.data_ptr(), stream)

        def flat_e2e():
            if world == 1:
                mse_b200.check(mse_b200.lib().mse_search_flat(ix._h, q_host.data_ptr(), nq, k, ids_host.data_ptr(), sc_host.data_ptr()), "mse_search_flat")
            else:
                q_dev.copy_(q_host, non_blocking=True)
                flat_resident()
                ids_host.copy_(out_ids, non_blocking=True)
                sc_host.copy_(out_sc, non_blocking=True)
                torch.cuda.current_stream().synchronize()

        with ClockSampler(local_rank) as cs:
            ms_step, launches = timed(flat_resident, warmup, steps)
        fclocks = cs.summary()
        ms_e2e, _ = timed(flat_e2e, min(warmup, 2), steps)
        ix.profile(True)
        for _ in range(2):
            ix.search_dev(q_dev.data_ptr(), nq, k, ids_dev.data_ptr(), sc_dev.data_ptr(), stream)
        st = ix.stats()
        ix.profile(False)
        nq_pad = (nq + 127) // 128 * 128
        tf = 2.0 * nq_pad * n_local * D / (st["scoring_ns"] * 1e-9) / 1e12 if st["scoring_ns"] else None
        search = {"metric": SEARCH_METRIC, "value": nq / (ms_step * 1e-3), "unit": "queries/s", "n_gpus": world, "ms_per_step": ms_step,
                  "higher_is_better": True, "scaling": "strong", "dtype": "fp16 x fp16 -> fp32 (tcgen05) + fp64 rerank", "config": flat_cfg,
                  "clocks": fclocks,
                  "e2e": {"value": nq / (ms_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": nq * D * 4, "d2h_bytes_per_step": nq * k * 8,
                          "ms_per_step": ms_e2e},
                  "gpu_launches": launches,
                  "roofline": {"kernel": "k_gemm_tn<256,FlatEpilogue> (all chunk launches of one step)", "bound": "tensor", "achieved": tf,
                               "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": tf / pk["tf_sustained"] if tf else None,
                               "peak_source": pk["src"] + " (sustained bf16 cuBLAS)", "frac_of_burst": tf / pk["tf_burst"] if tf else None,
                               "hbm_gbs": n_local * D * 2 / (st["scoring_ns"] * 1e-9) / 1e9 if st["scoring_ns"] else None,
                               "launches_per_step": st["scoring_launches"], "kernel_ms_per_step": st["scoring_ns"] * 1e-6,
                               "kernel_share_of_step": st["scoring_ns"] * 1e-6 / ms_step,
                               "traffic": NCU_TRAFFIC["flat_gemm_bytes_per_row_byte"] * n_local * D * 2 / max(st["scoring_launches"], 1),
                               "traffic_note": "bytes per launch (mean over the step's chunk launches), ratio from profiles/r01_flat_gemm_ncu_full.md: "
                                               "each index row leaves HBM once"},
                  "search_stats": st}
        if result:
            result["search"] = search
        else:
            result = dict(search, steps=steps, warmup=warmup, vs_baseline=None, data="synthetic")
        ix.close()

    # ============================================================ graph search (C4 at a single-GPU size)
    if "graph" in wl:
        from mse_b200 import diskann as dk
        n_total, nq, L, W, R, k = args.graph_rows, args.graph_queries, args.graph_L, args.graph_W, 64, 10
        lo, hi = n_total * rank // world, n_total * (rank + 1) // world
        n_local = hi - lo
        gcfg = {"workload": "vamana_graph_search", "index_rows": n_total, "dim": D, "R": R, "L_build": 192, "maxc": 750, "alpha": 1.0,
                "queries": nq, "L": L, "k": k,
                "data": "mixture of 4096 Gaussians (sigma 0.3), unit rows, fp16 (SURVEY 8d C4)" if args.graph_data == "mixture" else
                        "unit rows on a 48-dimensional latent subspace + 5 % isotropic noise, fp16",
                "sharding": f"id-range x{world}, one independent sub-graph per GPU, all-gather + merge of per-shard top-{k}",
                "l2_policy": "index larger than L2 (rows x 2304 B >> 126 MB)",
                "note": "BASELINE configs[3] names 1e8 rows on 8 GPUs; the default run holds index_rows on this GPU count"}
        g4 = torch.Generator(device=dev).manual_seed(4)
        cent = torch.randn((4096, D), generator=g4, device=dev)
        cent /= cent.norm(dim=1, keepdim=True)

        basis = torch.linalg.qr(torch.randn((D, 48), generator=g4, device=dev))[0].T.contiguous()   # 48 orthonormal directions

        def draw(m, seed):
            gg = torch.Generator(device=dev).manual_seed(seed)
            if args.graph_data == "latent":
                z = torch.randn((m, 48), generator=gg, device=dev)
                xx = (z / z.norm(dim=1, keepdim=True)) @ basis + 0.05 * torch.randn((m, D), generator=gg, device=dev) / D ** 0.5
            else:
                a = torch.randint(0, 4096, (m,), generator=gg, device=dev)
                xx = cent[a] + 0.3 * torch.randn((m, D), generator=gg, device=dev) / D ** 0.5
            return (xx / xx.norm(dim=1, keepdim=True)).to(torch.float16).contiguous()

        vl = dk.VectorList(D, device=local_rank)
        vl.reserve(n_local)
        want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline and n_local <= 4_000_000
        x_host = np.empty((n_local, D), np.float16) if want_cpu else None   # only the CPU oracle leg needs the rows on the host
        x_head = None
        chunk = 1 << 18
        for c0 in range(0, n_local, chunk):
            m = min(chunk, n_local - c0)
            xb = draw(m, 4_000_003 + lo + c0)
            vl.add_f16_dev(xb.data_ptr(), m, stream)
            if want_cpu:
                x_host[c0:c0 + m] = xb.cpu().numpy()
            if c0 == 0:
                x_head = xb[: min(m, 100_000)].float().mean(dim=0).cpu().numpy()      # RabitQ "training": the dataset mean (rabitq.py:14)
            del xb
        q16 = draw(nq, 5)
        q32 = q16.float().contiguous()
        q16_host = q16.cpu().pin_memory()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dk.random_fill_graph(vl, R, seed=1 + rank)
        med = dk.medioid(vl)
        bst = dk.build_graph(vl, med, dk.IndexBuildConfig(r=R, l=192, maxc=750), seed=7 + rank)
        build_s = time.perf_counter() - t0
        # ground truth: exact top-k of the whole index (flat search per shard + merge)
        gt_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
        gt_sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
        vl.search_dev(q32.data_ptr(), nq, k, gt_ids.data_ptr(), gt_sc.data_ptr(), stream)
        gt_ids += lo

        def merge(ids_local, sc_local):
            if world == 1:
                return ids_local, sc_local
            a_ids = torch.empty((world, nq, k), dtype=torch.int32, device=dev)
            a_sc = torch.empty((world, nq, k), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(a_ids.view(-1, k), ids_local.contiguous())
            dist.all_gather_into_tensor(a_sc.view(-1, k), sc_local.contiguous())
            o_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
            o_sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
            mse_b200.merge_topk(local_rank, a_ids.data_ptr(), a_sc.data_ptr(), world, nq, k, o_ids.data_ptr(), o_sc.data_ptr(), stream)
            return o_ids, o_sc

        gt_ids, gt_sc = merge(gt_ids, gt_sc)

        def recall(ids):
            a = ids.long().unsqueeze(2) == gt_ids.long().unsqueeze(1)
            return float(a.any(dim=2).float().sum().item() / (nq * k))

        ids_d = torch.empty((nq, L), dtype=torch.int32, device=dev)
        sc_d = torch.empty((nq, L), dtype=torch.int64, device=dev)
        len_d = torch.empty(nq, dtype=torch.int32, device=dev)
        dist_d = torch.empty(nq, dtype=torch.int64, device=dev)
        res_host = torch.empty((nq, k), dtype=torch.int32).pin_memory()
        state = {}

        def greedy_resident():
            dk.greedy_search_dev(vl, q16.data_ptr(), nq, L, med, ids_d.data_ptr(), sc_d.data_ptr(), len_d.data_ptr(), dist_d.data_ptr(), stream)
            state["top"] = merge((ids_d[:, :k] + lo), sc_d[:, :k].float() * (1.0 / 4294967296.0))

        def greedy_e2e():
            q16.copy_(q16_host, non_blocking=True)
            greedy_resident()
            res_host.copy_(state["top"][0], non_blocking=True)
            torch.cuda.current_stream().synchronize()

        with ClockSampler(local_rank) as cs:
            ms_g, launches_g = timed(greedy_resident, warmup, steps)
        gclocks = cs.summary()
        dk.greedy_search_check(vl, nq)
        rec_g = recall(state["top"][0])
        ms_g_e2e, _ = timed(greedy_e2e, min(warmup, 2), steps)
        # the search kernel alone (CUDA events on its stream), for the roofline
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(steps):
            dk.greedy_search_dev(vl, q16.data_ptr(), nq, L, med, ids_d.data_ptr(), sc_d.data_ptr(), len_d.data_ptr(), dist_d.data_ptr(), stream)
        ev[1].record()
        torch.cuda.synchronize()
        ms_kernel = ev[0].elapsed_time(ev[1]) / steps
        n_dist = float(dist_d.double().sum().item())
        row_bytes = n_dist * D * 2
        gbs = row_bytes / (ms_kernel * 1e-3) / 1e9
        graph = {"metric": GRAPH_METRIC, "value": nq / (ms_g * 1e-3), "unit": "queries/s", "recall_at_10": rec_g, "n_gpus": world, "ms_per_step": ms_g,
                 "higher_is_better": True, "scaling": "strong", "dtype": "fp16 rows, f32 FMA, i64 fixed-point scores (bit-exact fast_dot)",
                 "config": gcfg, "clocks": gclocks,
                 "e2e": {"value": nq / (ms_g_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": nq * D * 2, "d2h_bytes_per_step": nq * k * 4,
                         "ms_per_step": ms_g_e2e},
                 "gpu_launches": launches_g,
                 "distances_per_query": n_dist / nq,
                 "roofline": {"kernel": "k_greedy_search_wq<18> (one warp per query, half a warp per gathered row, exact fp16 rows)", "bound": "hbm", "achieved": gbs, "peak": pk["hbm"],
                              "unit": "GB/s", "frac": gbs / pk["hbm"], "peak_source": pk["src"] + " (HBM copy)",
                              "algorithmic_bytes": "distances x 2304 B (gathered rows; adjacency lists and the query are < 2 %)",
                              "launches_per_step": 1, "kernel_ms_per_step": ms_kernel, "kernel_share_of_step": ms_kernel / ms_g,
                              "traffic": NCU_TRAFFIC["greedy_bytes_per_row_byte"] * row_bytes,
                              "traffic_note": "ratio measured at 1 M rows (profiles/r01i_greedy_1m_ncu_full.md): every gathered row leaves HBM once"},
                 "build": {"seconds": build_s, "points_per_s": n_local / build_s, "stats": bst, "mean_degree": None}}
        # compressed traversal (C4: RabitQ codes, beam W): candidates by the RabitQ estimate, expanded nodes exact
        try:
            gm = torch.Generator(device=dev).manual_seed(11)
            P = torch.linalg.qr(torch.randn((D, D), generator=gm, device=dev))[0][:512].contiguous()
            rq = dk.RabitQ(x_head, P.cpu().numpy(), device=local_rank)
            t0 = time.perf_counter()
            rq.encode_index(vl, 0)                                   # rows are encoded where they lie in HBM
            enc_s = time.perf_counter() - t0
            qtm = torch.empty((nq, 513), dtype=torch.float32, device=dev)
            top_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
            top_sc = torch.empty((nq, k), dtype=torch.int64, device=dev)
            top_len = torch.empty(nq, dtype=torch.int32, device=dev)
            cm_d = torch.empty(nq, dtype=torch.int64, device=dev)
            pc_d = torch.empty(nq, dtype=torch.int64, device=dev)

            def make_beam(Lb):
                def beam_resident():
                    rq.query_dev(q32.data_ptr(), nq, qtm.data_ptr(), stream)
                    dk.beam_search_dev(vl, q16.data_ptr(), nq, Lb, W, med, k, top_ids.data_ptr(), top_sc.data_ptr(), top_len.data_ptr(), cm_d.data_ptr(),
                                       pc_d.data_ptr(), stream, d_qtm=qtm.data_ptr(), rabitq=rq)
                    state["btop"] = merge(top_ids + lo, top_sc.float() * (1.0 / 4294967296.0))
                return beam_resident

            variants = {}
            est_now = 0
            for name, est, Lb in (("script (norms * dots, rabitq.py:48) L=64", 0, L), ("script L=128", 0, 2 * L), ("script L=256", 0, 4 * L),
                                  ("paper (norms / dots) L=64", 1, L)):
                if est != est_now:
                    rq.encode_index(vl, est)
                    est_now = est
                ms_b, launches_b = timed(make_beam(Lb), min(warmup, 2), steps)
                dk.greedy_search_check(vl, nq)
                exact_b = float(cm_d.double().sum().item()) * D * 2
                code_b = float(pc_d.double().sum().item()) * (64 + 4)
                variants[name] = {"value": nq / (ms_b * 1e-3), "unit": "queries/s", "recall_at_10": recall(state["btop"][0]), "L": Lb, "ms_per_step": ms_b,
                                  "exact_rows_per_query": float(cm_d.double().mean().item()), "code_cmps_per_query": float(pc_d.double().mean().item()),
                                  "algorithmic_gbs": (exact_b + code_b) / (ms_b * 1e-3) / 1e9, "gpu_launches": launches_b}
            graph["rabitq_beam"] = {"config": {"W": W, "code_bytes": 64, "output_dims": 512,
                                               "kernel": "k_beam_search_wq<18> (one warp per query, estimate straight from the sign codes, exact rows for expanded nodes)",
                                               "note": "results are the expanded nodes only (query_disk_index.rs:172-183); on this mixture every cluster is a near-isotropic "
                                                       "ball of ~n/4096 rows, so recall tracks expanded rows / cluster size and grows with L"},
                                    "encode_seconds": enc_s, "variants": variants}
            rq.close()
        except Exception as e:  # the exact path above is the graph headline; report, do not hide
            graph["rabitq_beam"] = {"error": repr(e)}
        if want_cpu:
            from oracle import oracle as O
            O.build()
            adj, deg = vl.get_graph()
            graph["build"]["mean_degree"] = float(deg.mean())
            og = O.IndexGraph(n_local, adj.shape[1])
            og.set(adj, deg)
            ocfg = O.make_config(r=R, l=L, maxc=750)
            qs = q16_host.numpy()[: args.cpu_sample_graph_queries]
            O.greedy_search_batch(med, qs[:64], x_host, og, ocfg)
            t0 = time.perf_counter()
            o_ids, o_sc, o_len, o_dist = O.greedy_search_batch(med, qs, x_host, og, ocfg)
            dt = time.perf_counter() - t0
            same = bool(np.array_equal(o_ids, ids_d.cpu().numpy().view(np.uint32)[: qs.shape[0]]) and np.array_equal(o_sc, sc_d.cpu().numpy()[: qs.shape[0]]))
            graph["cpu_baseline"] = {"value": qs.shape[0] / dt, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                                     "sample": f"{qs.shape[0]} of the {nq} queries, same graph (built on the GPU) and rows, oracle greedy_search L={L}, OpenMP over queries",
                                     "distances_per_query": float(o_dist.mean()), "gpu_results_bit_identical": same}
        if result:
            result["graph"] = graph
        else:
            result = dict(graph, steps=steps, warmup=warmup, vs_baseline=None, data="synthetic")
        vl.close()

    # ============================================================ end to end (C5 at reduced scale): encode -> index -> build -> serve
    if "e2e" in wl:
        from mse_b200 import diskann as dk
        B, n_img, nq_t, nq_i, Ls, k = 256, args.e2e_images // 256 * 256, 500, 500, 64, 10
        wpath = os.path.join(tempfile.gettempdir(), f"mse_bench_e2e_rank{rank}.msew")
        sd = random_openclip_state_dict(dev)
        sd.update(random_openclip_text_state_dict(dev))
        mse_b200.weights.save_weights(wpath, sd, mse_b200.weights.config_for(sd))
        del sd
        enc = mse_b200.Encoder(wpath, device=local_rank, max_batch=B)
        os.remove(wpath)
        vl = dk.VectorList(D, device=local_rank)
        vl.reserve(n_img)
        feat = torch.empty((B, D), dtype=torch.float16, device=dev)
        g6 = torch.Generator(device=dev).manual_seed(6 + rank)
        base_imgs = torch.randint(0, 256, (B, 24, 24, 3), generator=g6, device=dev, dtype=torch.uint8)

        def make_batch():
            # blocky pattern + per-pixel noise so that images (and their embeddings) differ from each other
            up = base_imgs[torch.randperm(B, generator=g6, device=dev)].repeat_interleave(16, 1).repeat_interleave(16, 2).to(torch.int16)
            noise = torch.randint(-40, 41, (B, 384, 384, 3), generator=g6, device=dev, dtype=torch.int16)
            return (up + noise).clamp_(0, 255).to(torch.uint8).contiguous()

        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        imgs = make_batch()
        enc.encode_image_dev(imgs.data_ptr(), B, feat.data_ptr(), stream)      # warm-up
        barrier()
        t_wall0 = time.perf_counter()
        enc_ms = 0.0
        for b0 in range(0, n_img, B):
            imgs = make_batch()
            ev[0].record()
            enc.encode_image_dev(imgs.data_ptr(), B, feat.data_ptr(), stream)
            vl.add_f16_dev(feat.data_ptr(), B, stream)
            ev[1].record()
            ev[1].synchronize()
            enc_ms += ev[0].elapsed_time(ev[1])
        t0 = time.perf_counter()
        dk.random_fill_graph(vl, 64, seed=1 + rank)
        med = dk.medioid(vl)
        bst = dk.build_graph(vl, med, dk.IndexBuildConfig(r=64, l=192, maxc=750), seed=7 + rank)
        build_s = time.perf_counter() - t0
        # serve: 500 text queries (token ids: no tokenizer model offline) + 500 image queries, one search batch each
        gid = torch.Generator(device="cpu").manual_seed(8)
        ids = torch.ones((nq_t, 64), dtype=torch.int32)
        for i in range(nq_t):
            Lt = int(torch.randint(3, 17, (1,), generator=gid))
            ids[i, :Lt] = torch.randint(2, 32000, (Lt,), generator=gid, dtype=torch.int32)
        ids_dev = ids.to(dev)
        q_feat = torch.empty((nq_t + nq_i, D), dtype=torch.float16, device=dev)
        q_imgs = [make_batch() for _ in range((nq_i + B - 1) // B)]
        out_ids = torch.empty((nq_t + nq_i, Ls), dtype=torch.int32, device=dev)
        out_sc = torch.empty((nq_t + nq_i, Ls), dtype=torch.int64, device=dev)
        out_len = torch.empty(nq_t + nq_i, dtype=torch.int32, device=dev)
        out_dist = torch.empty(nq_t + nq_i, dtype=torch.int64, device=dev)
        res_host = torch.empty((nq_t + nq_i, k), dtype=torch.int32).pin_memory()

        def serve():
            for b0 in range(0, nq_t, B):
                m = min(B, nq_t - b0)
                enc.encode_text_dev(ids_dev[b0:b0 + m].data_ptr(), m, q_feat[b0:b0 + m].data_ptr(), stream)
            for bi, b0 in enumerate(range(0, nq_i, B)):
                m = min(B, nq_i - b0)
                enc.encode_image_dev(q_imgs[bi].data_ptr(), m, q_feat[nq_t + b0:nq_t + b0 + m].data_ptr(), stream)
            dk.greedy_search_dev(vl, q_feat.data_ptr(), nq_t + nq_i, Ls, med, out_ids.data_ptr(), out_sc.data_ptr(), out_len.data_ptr(),
                                 out_dist.data_ptr(), stream)
            res_host.copy_(out_ids[:, :k], non_blocking=True)
            torch.cuda.current_stream().synchronize()

        serve()
        ms_serve, _ = timed(serve, 1, 3)
        dk.greedy_search_check(vl, nq_t + nq_i)
        wall = time.perf_counter() - t_wall0
        e2e = {"metric": "end to end: encode + index + build + serve (BASELINE configs[4] at reduced scale)", "n_gpus": world,
               "config": {"workload": "e2e_encode_build_serve", "images_per_gpu": n_img, "graph": "R 64, L 192, C 750", "queries": "500 text (token ids) + 500 image, L = 64, top-10",
                          "note": "configs[4] names 1M images on 8 GPUs (>= 61 s of encoder work at the tensor roofline); this run encodes images_per_gpu per rank"},
               "encode": {"images_per_s": world * n_img / (enc_ms * 1e-3), "seconds": enc_ms * 1e-3},
               "build": {"seconds": build_s, "points_per_s": n_img / build_s, "stats": bst},
               "serve": {"queries_per_s": (nq_t + nq_i) / (ms_serve * 1e-3), "ms_per_batch_of_1000": ms_serve,
                         "includes": "text tower (500) + image tower (500) + graph search + D2H of top-10 ids"},
               "wall_seconds_total": wall}
        if result:
            result["e2e_pipeline"] = e2e
        else:
            result = dict(e2e, steps=steps, warmup=warmup, vs_baseline=None, data="synthetic")
        vl.close()
        enc.close()

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count()
            if "tower" in wl:
                ips, dt, cores = cpu_reference_tower(args.cpu_sample_images, 1, 1)
                result["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                                          "sample": f"{args.cpu_sample_images} images, full 27-block tower, transformers SiglipVisionModel fp32 "
                                                    f"(stand-in for clip_server.py device=cpu; open_clip is not installable here)"}
            if "flat" in wl:
                qps, dt, cores = cpu_reference_flat(args.rows, args.k, 1, 1, args.cpu_sample_rows, args.cpu_sample_queries)
                cb = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                      "sample": f"{args.cpu_sample_queries} queries x {args.cpu_sample_rows}-row slice, scaled by rows to {args.rows}"}
                if "search" in result:
                    result["search"]["cpu_baseline"] = cb
                else:
                    result["cpu_baseline"] = cb
        print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
