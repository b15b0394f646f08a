#!/usr/bin/env python
"""bench.py -- ONE compact JSON line per run on stdout (contract: the task prompt's bench section); the long form of every
workload goes to gpurun_out/bench_full_n<N>.json.

BASELINE.json's metric has two halves, and the run measures both plus the two remaining single-node configs:

  headline  configs[1]  SigLIP ViT-SO400M-14/384 image tower, batch 256 per GPU, fp16 storage / fp32 accumulate (data parallel, weak)
  search    configs[2]  flat top-100: 1024 f32 queries over a 10M x 1152 fp16 index, id-range sharded over the ranks (strong)
  graph     configs[3]  Vamana graph over 12.5M x 1152 rows PER GPU (1e8 rows at N = 8), RabitQ-compressed beam traversal with exact
                        scores for expanded nodes, 4096 batched queries, recall@10 of the merged result against the exact (flat)
                        ground truth; the exact-row greedy_search is reported beside it (weak: the index grows with N)
  c1        configs[0]  one text query: text tower at batch 1 + top-10 over a 1k x 1152 index

Every search goes through the C ABI's shard group (csrc/shard.cu): local search -> one ncclAllGather of packed top-k -> merge.
`value` = inputs resident in HBM; `e2e` = the same step through host buffers (H2D of the step's inputs, D2H of its results).

--impl reference times the CPU stand-ins of the reference's path on the host cores, EXACTLY --steps steps after --warmup warm-ups
of a bounded sample (stated in cpu_baseline.sample): transformers SiglipVisionModel fp32 for clip_server.py device=cpu, the C
restatement of the flat scan and of greedy_search (oracle/, AVX2 + OpenMP).  The reference itself (open_clip + Rust nightly +
faiss) cannot be installed or built in this image.
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
GRAPH_METRIC = "queries/sec@recall10 (Vamana graph, RabitQ-compressed beam search, 4096 batched queries)"
SEARCH_METRIC = "queries/sec (flat top-100, 1024 queries x 10M x 1152 fp16 index)"
C1_METRIC = "queries/sec (1 text query: text tower batch 1 + flat top-10 over 1k x 1152)"
FLOP_PER_IMAGE = 670.35e9  # 27 layers 665.46 + patch-embed 0.99 + MAP head 3.90 GFLOP
TEXT_WEIGHT_BYTES = 0.826e9  # 412.9 M parameters touched x 2 B
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernels from the committed `ncu --set full` captures (per launch
# or per algorithmic byte).  Measured under a profiler on the named capture, not by this run.
NCU_TRAFFIC = {
    # profiles/r03k_tower_block_final_ncu.md: one block at batch 256 = QKV 1.759 + out-proj 1.263 + fc1 2.072 + fc2 3.606 GB (read + write); x 27 blocks
    "tower_gemm_bytes_per_step_b256": (27 * (1.759 + 1.263 + 2.072 + 3.606) * 1e9, "profiles/r03k_tower_block_final_ncu.md"),
    # profiles/r02t_ncu_full.md: 17.155 GB read + 10.8 MB written for the 17.150 GB of rows the launch scored
    "flat_gemm_bytes_per_row_byte": ((17.154500 + 0.010778) / 17.149723, "profiles/r02t_ncu_full.md"),
    # profiles/r02t_ncu_full.md (12.5 M rows, 4096 queries): greedy L = 64: 14.872 GB read + 0.594 GB written for 15.17 GB of gathered rows;
    # RabitQ beam L = 512: 17.377 + 4.516 GB per launch (visited-set tables), 1.53 GB algorithmic
    "greedy_bytes_per_row_byte": ((15.818 + 0.595) / 15.17, "profiles/r03e_ncu_full.md"),   # with the L2 row prefetch (r02t: 14.872 + 0.594)
    "text_blocks_bytes_per_launch": ((824.36 + 4.99) * 1e6, "profiles/r03e_ncu_full.md"),     # k_text_blocks: the 0.826 GB of weights, once
    "beam_l512_bytes_per_launch": ((20.838 + 3.397) * 1e9, "profiles/r03h_beam_l512_lines.txt"),   # r02t: 17.377 + 4.516
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

def cpu_reference_tower(images_per_step: int, steps: int, warmup: int):
    import torch
    from oracle import towers as T
    torch.set_num_threads(os.cpu_count() or 1)
    m = T.build_vision(depth=27, seed=42)
    imgs = T.synthetic_images(1, images_per_step)
    for _ in range(warmup):
        T.encode_image(m, imgs)
    t0 = time.perf_counter()
    for _ in range(steps):
        T.encode_image(m, imgs)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return images_per_step / dt, dt, os.cpu_count()


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


def c1_inputs():
    """token ids from rng(seed=1) (no tokenizer model offline), index rows from default_rng(0)."""
    import numpy as np
    rng = np.random.default_rng(1)
    ids = np.ones((1, 64), np.int32)
    n = int(rng.integers(3, 17))
    ids[0, :n] = rng.integers(2, 32000, n)
    x = np.random.default_rng(0).standard_normal((1000, D)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return ids, x.astype(np.float16)


def cpu_reference_c1(steps: int, warmup: int):
    """clip_server.py device=cpu stand-in (transformers SiglipTextModel fp32, all cores) + the C brute force over 1k rows."""
    import numpy as np
    import torch
    from oracle import oracle as O
    from oracle import towers as T
    O.build()
    torch.set_num_threads(os.cpu_count() or 1)
    m = T.build_text(depth=27, seed=43)
    ids, x16 = c1_inputs()

    def step():
        e = np.asarray(T.encode_text(m, ids), np.float32)
        return O.flat_search(e, x16, 10, mode=1)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return 1.0 / dt, dt, os.cpu_count()


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


# ---------------------------------------------------------------- C4 data

GRAPH_DATA_NOTE = {
    "families": "SURVEY 8d C4's clustered mixture with a second level: 4096 cluster centres (unit), family centres at sigma 0.5 around them, "
                "rows in families of 32 at sigma 0.3 around their family centre, unit-normalised fp16; queries = held-out members of "
                "existing families.  (One isotropic level leaves the ~3000 rows of a cluster equidistant to +-0.003 at 12.5 M rows -- no "
                "neighbourhood structure inside a cluster, which is SURVEY's own argument against isotropic data, one level down.)",
    "mixture": "mixture of 4096 Gaussians (sigma 0.3), unit rows, fp16 (SURVEY 8d C4 as written)",
    "latent": "unit rows on a 48-dimensional latent subspace + 5 % isotropic noise, fp16",
}
FAMILY = 32


class GraphData:
    """Seeded generators on the device; row i of the global index is a pure function of (seed, i's chunk), so any shard of any
    world size holds the same rows."""

    def __init__(self, kind: str, dev):
        import torch
        self.kind, self.dev, self.torch = kind, dev, torch
        g4 = torch.Generator(device=dev).manual_seed(4)
        self.cent = torch.randn((4096, D), generator=g4, device=dev)
        self.cent /= self.cent.norm(dim=1, keepdim=True)
        self.basis = torch.linalg.qr(torch.randn((D, 48), generator=g4, device=dev))[0].T.contiguous()
        self.chunk = 1 << 18                                       # rows per generation chunk (a multiple of FAMILY)

    def _family_centres(self, fam0: int, n_fam: int):
        """centres of families fam0 .. fam0+n_fam (global family ids), generated in blocks of 8192 families"""
        torch, out, blk = self.torch, [], 8192
        b = fam0 // blk
        while b * blk < fam0 + n_fam:
            gg = torch.Generator(device=self.dev).manual_seed(9_000_001 + b)
            a = torch.randint(0, 4096, (blk,), generator=gg, device=self.dev)
            f = self.cent[a] + 0.5 * torch.randn((blk, D), generator=gg, device=self.dev) / D ** 0.5
            lo, hi = max(fam0, b * blk), min(fam0 + n_fam, (b + 1) * blk)
            out.append(f[lo - b * blk: hi - b * blk])
            b += 1
        return torch.cat(out)

    def rows(self, row0: int, m: int):
        """global rows [row0, row0+m); row0 must be a multiple of the chunk size"""
        torch = self.torch
        gg = torch.Generator(device=self.dev).manual_seed(4_000_003 + row0)
        if self.kind == "latent":
            z = torch.randn((m, 48), generator=gg, device=self.dev)
            xx = (z / z.norm(dim=1, keepdim=True)) @ self.basis + 0.05 * torch.randn((m, D), generator=gg, device=self.dev) / D ** 0.5
        elif self.kind == "mixture":
            a = torch.randint(0, 4096, (m,), generator=gg, device=self.dev)
            xx = self.cent[a] + 0.3 * torch.randn((m, D), generator=gg, device=self.dev) / D ** 0.5
        else:
            n_fam = (m + FAMILY - 1) // FAMILY
            f = self._family_centres(row0 // FAMILY, n_fam)
            fam_of_row = torch.arange(m, device=self.dev) // FAMILY
            perm = torch.randperm(m, generator=gg, device=self.dev)        # families are not contiguous in memory
            xx = f[fam_of_row[perm]] + 0.3 * torch.randn((m, D), generator=gg, device=self.dev) / D ** 0.5
        return (xx / xx.norm(dim=1, keepdim=True)).to(torch.float16).contiguous()

    def queries(self, nq: int, n_total: int):
        torch = self.torch
        gg = torch.Generator(device=self.dev).manual_seed(5)
        if self.kind == "latent":
            z = torch.randn((nq, 48), generator=gg, device=self.dev)
            xx = (z / z.norm(dim=1, keepdim=True)) @ self.basis + 0.05 * torch.randn((nq, D), generator=gg, device=self.dev) / D ** 0.5
        elif self.kind == "mixture":
            a = torch.randint(0, 4096, (nq,), generator=gg, device=self.dev)
            xx = self.cent[a] + 0.3 * torch.randn((nq, D), generator=gg, device=self.dev) / D ** 0.5
        else:
            fam = torch.randint(0, max(n_total // FAMILY, 1), (nq,), generator=gg, device=self.dev)
            f = torch.cat([self._family_centres(int(i), 1) for i in fam.tolist()])
            xx = f + 0.3 * torch.randn((nq, D), generator=gg, device=self.dev) / D ** 0.5
        return (xx / xx.norm(dim=1, keepdim=True)).to(torch.float16).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workloads", default="tower,flat,graph,c1", help="comma list of tower, flat, graph, c1, e2e (e2e = BASELINE configs[4] at reduced scale; not in the default set)")
    ap.add_argument("--graph-rows-per-gpu", type=int, default=12_500_000, help="graph: index rows per rank (C4: 12.5M x 8 GPUs = 1e8)")
    ap.add_argument("--graph-queries", type=int, default=4096)
    ap.add_argument("--graph-L", type=int, default=64, help="search list size of the exact greedy_search (C4 'beam 64')")
    ap.add_argument("--graph-sweep", default="64:4,128:4,256:4,512:4", help="L:W variants of the RabitQ beam traversal, cheapest first; the headline is the first with recall@10 >= --recall-target")
    ap.add_argument("--recall-target", type=float, default=0.9)
    ap.add_argument("--graph-data", default="families", choices=sorted(GRAPH_DATA_NOTE))
    ap.add_argument("--cpu-sample-graph-rows", type=int, default=1_000_000, help="rows of the separate index the CPU graph baseline runs on")
    ap.add_argument("--cpu-sample-graph-queries", type=int, default=512)
    ap.add_argument("--e2e-images", type=int, default=32768, help="e2e workload: images per rank to encode and index (when --e2e-images-total is 0)")
    ap.add_argument("--e2e-images-total", type=int, default=0, help="e2e workload: images over all ranks (BASELINE configs[4]: 1 000 000 on 8 GPUs)")
    ap.add_argument("--batch", type=int, default=256, help="images per rank per step")
    ap.add_argument("--rows", type=int, default=10_000_000, help="flat: total index rows (all shards)")
    ap.add_argument("--queries", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--cpu-sample-rows", type=int, default=500_000)
    ap.add_argument("--cpu-sample-queries", type=int, default=32)
    ap.add_argument("--cpu-sample-images", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    steps, warmup = args.steps, max(args.warmup, 0)
    wl = [w for w in args.workloads.split(",") if w]
    if world == 8 and args.workloads == ap.get_default("workloads") and args.impl == "ours":
        # BASELINE configs[4] is an 8-GPU configuration: at N = 8 the default run also does the end-to-end pipeline at its stated size
        wl.append("e2e")
        if not args.e2e_images_total:
            args.e2e_images_total = 1_000_000
    tower_cfg = {"workload": "siglip_image_tower_b256", "tower": "ViT-SO400M-14-SigLIP-384 image tower (random init)", "batch_per_gpu": args.batch,
                 "image": "384x384x3 u8", "storage": "fp16", "accumulate": "fp32", "parallelism": f"dp{world}",
                 "l2_policy": "inputs larger than L2 (per-layer activations 430 MB - 1.6 GB >> 126 MB)"}
    flat_cfg = {"workload": "flat_top100", "queries": args.queries, "k": args.k, "index_rows": args.rows, "dim": D, "index_dtype": "fp16",
                "query_dtype": "f32", "sharding": f"id-range x{world}", "l2_policy": "inputs larger than L2 (index shard >> 126 MB)"}

    if args.impl == "reference":
        if rank != 0:
            return
        n_img = args.cpu_sample_images
        ips, dt, cores = cpu_reference_tower(n_img, steps, warmup)
        sample = (f"{n_img} images per step, {steps} timed steps after {warmup} warm-ups (full 27-block tower, transformers SiglipVisionModel fp32, "
                  f"torch threads = all cores); the GPU arm's step is {args.batch} images")
        out = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
               "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32 (CPU)",
               "data": "synthetic", "config": tower_cfg, "sample": {"images_per_step": n_img},
               "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
               "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        if "flat" in wl:
            qps, fdt, _ = cpu_reference_flat(args.rows, args.k, 1, 1, args.cpu_sample_rows, args.cpu_sample_queries)
            out["search"] = {"value": qps, "unit": "queries/s", "ms_per_step": fdt * 1e3,
                             "sample": f"{args.cpu_sample_queries} queries x {args.cpu_sample_rows}-row slice, 1 pass, scaled by rows to {args.rows}"}
        if "c1" in wl:
            qps, cdt, _ = cpu_reference_c1(2, 1)
            out["c1"] = {"value": qps, "unit": "queries/s", "ms_per_step": cdt * 1e3, "sample": "2 queries (SiglipTextModel fp32 + brute force over 1k rows)"}
        print(json.dumps(out), flush=True)
        return

    import numpy as np
    import torch
    import mse_b200
    from mse_b200 import diskann as dk
    from mse_b200.sharding import ShardGroup, shard_range

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    nccl_log = None
    if world > 1:
        if "NCCL_DEBUG" not in os.environ:
            # keep NCCL's own account of the communicators (rank / nranks lines) without putting it on stdout, which carries the JSON line
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            nccl_log = os.path.join(ROOT, "gpurun_out", f"nccl_n{world}_rank{rank}.log")
            os.environ.update(NCCL_DEBUG="INFO", NCCL_DEBUG_SUBSYS="INIT", NCCL_DEBUG_FILE=nccl_log)
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    grp = ShardGroup.from_torch_distributed(dist, local_rank)      # the library's own communicator: the data plane of every search below
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

    def kernel_ms(fn, n):
        """CUDA events around n back-to-back calls of one launch on the launching stream"""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        fn()
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(n):
            fn()
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / n

    def all_true(flag: bool) -> bool:
        if world == 1:
            return bool(flag)
        t = torch.tensor([1 if flag else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    result, full = {}, {}

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
        tr, tr_src = NCU_TRAFFIC["tower_gemm_bytes_per_step_b256"]
        roof = {"kernel": "k_gemm_tn<BN,LinearEpilogue> (all GEMM launches of one step)", "bound": "tensor", "achieved": gemm_tf,
                "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": gemm_tf / pk["tf_sustained"] if gemm_tf else None,
                "peak_source": pk["src"] + " (sustained bf16 cuBLAS)", "frac_of_burst": gemm_tf / pk["tf_burst"] if gemm_tf else None,
                "launches_per_step": st["gemm_launches"], "kernel_ms_per_step": st["gemm_ns"] * 1e-6,
                "kernel_share_of_step": st["gemm_ns"] * 1e-6 / ms_step,
                "traffic": tr * B / 256 / max(st["gemm_launches"], 1),
                "traffic_note": f"bytes per launch (mean over the step's GEMM launches) from the ncu capture {tr_src}, not measured by this run",
                "attention": {"kernel": "k_mha_tc", "ms_per_step": st["attn_ns"] * 1e-6, "launches": st["attn_launches"],
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

    # ============================================================ flat search (C3)
    if "flat" in wl:
        rows_total, nq, k = args.rows, args.queries, args.k
        row_lo, row_hi = shard_range(rows_total, rank, world)
        n_local = row_hi - row_lo
        chunk = 1 << 19

        def flat_rows(c0, m):
            gen = torch.Generator(device=dev).manual_seed(2 * 1_000_003 + c0)
            xb = torch.randn((m, D), generator=gen, device=dev, dtype=torch.float32)
            return (xb / xb.norm(dim=1, keepdim=True)).to(torch.float16).contiguous()

        def fill(ix, lo, hi):
            """global rows [lo, hi): generated in the same 2^19-row chunks whatever the shard boundaries are"""
            ix.reserve(hi - lo)
            for c0 in range(lo // chunk * chunk, hi, chunk):
                xb = flat_rows(c0, min(chunk, rows_total - c0))
                a, b = max(lo, c0) - c0, min(hi, c0 + chunk) - c0
                ix.add_f16_dev(xb[a:b].contiguous().data_ptr(), b - a, stream)
                del xb

        ix = mse_b200.FlatIndex(D, device=local_rank, id_base=row_lo)
        fill(ix, row_lo, row_hi)
        gq = torch.Generator(device="cpu").manual_seed(3)
        q_host = torch.randn((nq, D), generator=gq, dtype=torch.float32)
        q_host = (q_host / q_host.norm(dim=1, keepdim=True)).contiguous().pin_memory()
        q_dev = q_host.to(dev)
        ids_dev = torch.empty((nq, k), dtype=torch.int32, device=dev)
        sc_dev = torch.empty((nq, k), dtype=torch.float32, device=dev)
        ids_host = torch.empty((nq, k), dtype=torch.int32).pin_memory()
        sc_host = torch.empty((nq, k), dtype=torch.float32).pin_memory()

        def flat_resident():
            grp.flat_search_dev(ix, q_dev.data_ptr(), nq, k, ids_dev.data_ptr(), sc_dev.data_ptr(), stream)

        def flat_e2e():
            if world == 1:   # the reference-facing host-pointer call (faiss Index::search)
                mse_b200.check(mse_b200.lib().mse_search_flat(ix._h, q_host.data_ptr(), nq, k, ids_host.data_ptr(), sc_host.data_ptr()), "mse_search_flat")
            else:
                q_dev.copy_(q_host, non_blocking=True)
                flat_resident()
                ids_host.copy_(ids_dev, non_blocking=True)
                sc_host.copy_(sc_dev, non_blocking=True)
                grp.check()

        gathers0 = grp.info()["all_gathers"]
        with ClockSampler(local_rank) as cs:
            ms_step, launches = timed(flat_resident, warmup, steps)
        fclocks = cs.summary()
        repaired = grp.check()                       # status word of the (identical) last step, read once outside the timed loop
        gathers_per_step = (grp.info()["all_gathers"] - gathers0) / max(warmup + steps, 1)
        ms_e2e, _ = timed(flat_e2e, min(warmup, 2), steps)
        ix.profile(True)
        for _ in range(2):
            ix.search_dev(q_dev.data_ptr(), nq, k, ids_dev.data_ptr(), sc_dev.data_ptr(), stream)
            ix.check()
        st = ix.stats()
        ix.profile(False)
        # N > 1: the merged result of a 32-query sample must equal the result of ONE unsharded index holding all the rows
        match = None
        if world > 1:
            ns = min(32, nq)
            flat_resident()
            grp.check()
            sh_ids, sh_sc = ids_dev[:ns].clone(), sc_dev[:ns].clone()
            ok = True
            if rank == 0:
                whole = mse_b200.FlatIndex(D, device=local_rank, id_base=0)
                fill(whole, 0, rows_total)
                w_ids = torch.empty((ns, k), dtype=torch.int32, device=dev)
                w_sc = torch.empty((ns, k), dtype=torch.float32, device=dev)
                whole.search_dev(q_dev.data_ptr(), ns, k, w_ids.data_ptr(), w_sc.data_ptr(), stream)
                whole.check()
                ok = bool(torch.equal(w_ids, sh_ids) and torch.equal(w_sc, sh_sc))
                whole.close()
                del whole
                torch.cuda.empty_cache()
            match = all_true(ok)                                     # reported, not asserted: a mismatch must reach the record
        nq_pad = (nq + 127) // 128 * 128
        tf = 2.0 * nq_pad * n_local * D / (st["scoring_ns"] * 1e-9) / 1e12 if st["scoring_ns"] else None
        trf, trf_src = NCU_TRAFFIC["flat_gemm_bytes_per_row_byte"]
        full["search"] = {"metric": SEARCH_METRIC, "value": nq / (ms_step * 1e-3), "unit": "queries/s", "n_gpus": world, "ms_per_step": ms_step,
                          "higher_is_better": True, "scaling": "strong", "dtype": "fp16 x fp16 -> fp32 (tcgen05) + fp64 rerank", "config": flat_cfg,
                          "clocks": fclocks,
                          "e2e": {"value": nq / (ms_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": nq * D * 4, "d2h_bytes_per_step": nq * k * 8,
                                  "ms_per_step": ms_e2e},
                          "gpu_launches": launches, "all_gathers_per_step": gathers_per_step, "uncertified_queries_repaired": repaired,
                          "sharded_ids_match": match,
                          "roofline": {"kernel": "k_gemm_tn<256,FlatEpilogue> (all chunk launches of one step)", "bound": "tensor", "achieved": tf,
                                       "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": tf / pk["tf_sustained"] if tf else None,
                                       "peak_source": pk["src"] + " (sustained bf16 cuBLAS)", "frac_of_burst": tf / pk["tf_burst"] if tf else None,
                                       "hbm_gbs": n_local * D * 2 / (st["scoring_ns"] * 1e-9) / 1e9 if st["scoring_ns"] else None,
                                       "launches_per_step": st["scoring_launches"], "kernel_ms_per_step": st["scoring_ns"] * 1e-6,
                                       "kernel_share_of_step": st["scoring_ns"] * 1e-6 / ms_step,
                                       "traffic": trf * n_local * D * 2 / max(st["scoring_launches"], 1),
                                       "traffic_note": f"bytes per launch (mean over the step's chunk launches), ratio from the ncu capture {trf_src}"},
                          "search_stats": st}
        ix.close()
        del ix
        torch.cuda.empty_cache()

    # ============================================================ graph search (C4: 12.5 M rows per GPU, 1e8 at N = 8)
    if "graph" in wl:
        per_gpu, nq, L, R, k = args.graph_rows_per_gpu, args.graph_queries, args.graph_L, 64, 10
        n_total = per_gpu * world
        lo, hi = shard_range(n_total, rank, world)
        n_local = hi - lo
        data = GraphData(args.graph_data, dev)
        gcfg = {"workload": "vamana_rabitq_beam_search", "index_rows": n_total, "rows_per_gpu": per_gpu, "dim": D, "R": R, "L_build": 192, "maxc": 750,
                "alpha": 1.0, "queries": nq, "k": k, "code": "RabitQ 512 sign bits (64 B) + 1 f32 scale per row", "data": GRAPH_DATA_NOTE[args.graph_data],
                "sharding": f"id-range x{world}, one independent Vamana sub-graph per GPU, every query on every shard, one all-gather + merge of per-shard top-{k}",
                "l2_policy": "index larger than L2 (rows x 2304 B >> 126 MB)"}

        def build_index(lo_, hi_, seed_off):
            vl_ = dk.VectorList(D, device=local_rank, id_base=lo_)
            vl_.reserve(hi_ - lo_)
            for c0 in range(lo_ // data.chunk * data.chunk, hi_, data.chunk):
                xb = data.rows(c0, data.chunk)
                a, b = max(lo_, c0) - c0, min(hi_, c0 + data.chunk) - c0
                xs = xb[a:b].contiguous()
                vl_.add_f16_dev(xs.data_ptr(), b - a, stream)
                del xb, xs
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dk.random_fill_graph(vl_, R, seed=1 + seed_off)
            med_ = dk.medioid(vl_)
            bst_ = dk.build_graph(vl_, med_, dk.IndexBuildConfig(r=R, l=192, maxc=750), seed=7 + seed_off)
            return vl_, med_, bst_, time.perf_counter() - t0

        vl, med, bst, build_s = build_index(lo, hi, rank)
        q16 = data.queries(nq, n_total)
        q32 = q16.float().contiguous()
        q16_host = q16.cpu().pin_memory()
        # ground truth: exact top-k of the whole index = sharded flat search
        gt_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
        gt_sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
        grp.flat_search_dev(vl, q32.data_ptr(), nq, k, gt_ids.data_ptr(), gt_sc.data_ptr(), stream)
        grp.check()

        def recall(ids):
            a = ids.long().unsqueeze(2) == gt_ids.long().unsqueeze(1)
            return float(a.any(dim=2).float().sum().item() / (nq * k))

        out_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
        out_sc = torch.empty((nq, k), dtype=torch.int64, device=dev)
        dist_d = torch.empty(nq, dtype=torch.int64, device=dev)
        res_host = torch.empty((nq, k), dtype=torch.int32).pin_memory()

        # ---- exact-row greedy_search (lib.rs:183-211) on every shard, merged
        def greedy_resident():
            grp.graph_search_dev(vl, q16.data_ptr(), nq, L, med, k, out_ids.data_ptr(), out_sc.data_ptr(), dist_d.data_ptr(), stream)

        def greedy_e2e():
            q16.copy_(q16_host, non_blocking=True)
            greedy_resident()
            res_host.copy_(out_ids, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        # the ground-truth flat search just before leaves the board at its power-cap clocks (r03g: 1237 MHz during a 125 ms timed
        # region): warm up for ~0.5 s so that the timed steps see the clocks a search-only workload runs at
        with ClockSampler(local_rank) as cs:
            ms_g, launches_g = timed(greedy_resident, max(warmup, 100), max(steps, 40))
        gclocks = cs.summary()
        dk.greedy_search_check(vl, nq)
        rec_g = recall(out_ids)
        ms_g_e2e, _ = timed(greedy_e2e, min(warmup, 2), steps)
        l_ids = torch.empty((nq, L), dtype=torch.int32, device=dev)
        l_sc = torch.empty((nq, L), dtype=torch.int64, device=dev)
        l_len = torch.empty(nq, dtype=torch.int32, device=dev)
        ms_gk = kernel_ms(lambda: dk.greedy_search_dev(vl, q16.data_ptr(), nq, L, med, l_ids.data_ptr(), l_sc.data_ptr(), l_len.data_ptr(),
                                                       dist_d.data_ptr(), stream), steps)
        n_dist = float(dist_d.double().sum().item())
        g_bytes = n_dist * D * 2
        trg, trg_src = NCU_TRAFFIC["greedy_bytes_per_row_byte"]
        greedy = {"value": nq / (ms_g * 1e-3), "unit": "queries/s", "recall_at_10": rec_g, "L": L, "ms_per_step": ms_g,
                  "e2e": {"value": nq / (ms_g_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": nq * D * 2, "d2h_bytes_per_step": nq * k * 4, "ms_per_step": ms_g_e2e},
                  "gpu_launches_per_step": launches_g / max(steps, 40), "distances_per_query_per_shard": n_dist / nq,
                  "roofline": {"kernel": "k_greedy_search_wq<18> (one warp per query, half a warp per gathered row, exact fp16 rows)", "bound": "hbm",
                               "achieved": g_bytes / (ms_gk * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s", "frac": g_bytes / (ms_gk * 1e-3) / 1e9 / pk["hbm"],
                               "algorithmic_bytes": "distances x 2304 B (gathered rows; adjacency lists and the query are < 2 %)",
                               "kernel_ms_per_step": ms_gk, "kernel_share_of_step": ms_gk / ms_g, "traffic": trg * g_bytes,
                               "traffic_note": f"ratio from the ncu capture {trg_src} (12.5 M rows, L = 64): every gathered row leaves HBM once"}}

        # ---- C4 proper: RabitQ codes for the frontier, exact rows for expanded nodes (query_disk_index.rs:144-212), beam W
        t0 = time.perf_counter()
        rq = dk.RabitQ.train(vl, sample_rows=100_000, output_dims=512, seed=11)   # rabitq.py:11-28 on the device: mean + random orthogonal P
        train_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        rq.encode_index(vl, 0)                                    # rows are encoded where they lie in HBM
        torch.cuda.synchronize()
        enc_s = time.perf_counter() - t0
        qtm = torch.empty((nq, 513), dtype=torch.float32, device=dev)
        cm_d = torch.empty(nq, dtype=torch.int64, device=dev)
        pc_d = torch.empty(nq, dtype=torch.int64, device=dev)

        def make_beam(Lb, Wb):
            def beam_resident():
                rq.query_dev(q32.data_ptr(), nq, qtm.data_ptr(), stream)
                grp.beam_search_dev(vl, q16.data_ptr(), nq, Lb, Wb, med, k, out_ids.data_ptr(), out_sc.data_ptr(), cm_d.data_ptr(), pc_d.data_ptr(),
                                    stream, d_qtm=qtm.data_ptr(), rabitq=rq)
            return beam_resident

        variants, head_name = {}, None
        for spec in args.graph_sweep.split(","):
            Lb, Wb = (int(v) for v in spec.split(":"))
            fn = make_beam(Lb, Wb)
            with ClockSampler(local_rank) as cs:
                ms_b, launches_b = timed(fn, min(warmup, 2), steps)
            dk.greedy_search_check(vl, nq)
            exact_rows, codes = float(cm_d.double().sum().item()), float(pc_d.double().sum().item())
            variants[spec] = {"L": Lb, "W": Wb, "value": nq / (ms_b * 1e-3), "unit": "queries/s", "recall_at_10": recall(out_ids), "ms_per_step": ms_b,
                              "exact_rows_per_query_per_shard": exact_rows / nq, "code_cmps_per_query_per_shard": codes / nq,
                              "gpu_launches": launches_b, "clocks": cs.summary(),
                              "bytes": exact_rows * (D * 2 + R * 4) + codes * (64 + 4)}
            if head_name is None and variants[spec]["recall_at_10"] >= args.recall_target:
                head_name = spec
                break                                              # the sweep is ordered cheapest first
        target_met = head_name is not None
        if head_name is None:
            head_name = max(variants, key=lambda s: variants[s]["recall_at_10"])
        hv = variants[head_name]
        Lh, Wh = hv["L"], hv["W"]
        head_fn = make_beam(Lh, Wh)

        def beam_e2e():
            q16.copy_(q16_host, non_blocking=True)
            q32.copy_(q16)                                        # f32 copy of the fp16 query for the codec's query side
            head_fn()
            res_host.copy_(out_ids, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        ms_b_e2e, _ = timed(beam_e2e, min(warmup, 2), steps)
        t_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
        t_sc = torch.empty((nq, k), dtype=torch.int64, device=dev)
        t_len = torch.empty(nq, dtype=torch.int32, device=dev)
        ms_bk = kernel_ms(lambda: dk.beam_search_dev(vl, q16.data_ptr(), nq, Lh, Wh, med, k, t_ids.data_ptr(), t_sc.data_ptr(), t_len.data_ptr(),
                                                     cm_d.data_ptr(), pc_d.data_ptr(), stream, d_qtm=qtm.data_ptr(), rabitq=rq), steps)
        # the merge is checked against an independent gather of every rank's own device top-k
        merge_ok = None
        if world > 1:
            head_fn()
            torch.cuda.synchronize()
            loc = torch.where(t_ids == -1, t_ids.long(), t_ids.long() + lo)
            a_ids = [torch.empty_like(loc) for _ in range(world)]
            a_sc = [torch.empty_like(t_sc) for _ in range(world)]
            dist.all_gather(a_ids, loc)
            dist.all_gather(a_sc, t_sc)
            ci, cs_ = torch.cat(a_ids, dim=1), torch.cat(a_sc, dim=1)
            cs_ = torch.where(ci < 0, torch.full_like(cs_, -2 ** 62), cs_)
            order = torch.argsort(ci, dim=1, stable=True)                        # id asc, then stable by score desc
            ci, cs_ = torch.gather(ci, 1, order), torch.gather(cs_, 1, order)
            order = torch.argsort(cs_, dim=1, descending=True, stable=True)
            mi = torch.gather(ci, 1, order)[:, :k]
            merge_ok = all_true(bool(torch.equal(mi.int(), out_ids)))
        graph = {"metric": GRAPH_METRIC, "value": hv["value"], "unit": "queries/s", "recall_at_10": hv["recall_at_10"], "recall_target": args.recall_target,
                 "recall_target_met": target_met, "L": Lh, "W": Wh, "n_gpus": world, "ms_per_step": hv["ms_per_step"], "higher_is_better": True,
                 "scaling": "weak", "dtype": "1-bit RabitQ codes (f32 estimate) for candidates, fp16 rows / f32 FMA / i64 fixed point for expanded nodes",
                 "config": gcfg, "clocks": hv["clocks"],
                 "e2e": {"value": nq / (ms_b_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": nq * D * 2, "d2h_bytes_per_step": nq * k * 4, "ms_per_step": ms_b_e2e},
                 "gpu_launches": hv["gpu_launches"], "merge_matches_independent_gather": merge_ok,
                 "roofline": {"kernel": "k_beam_search_wq<18> (one warp per query; estimates straight from the sign codes, exact rows for expanded nodes)",
                              "bound": "hbm", "achieved": hv["bytes"] / (ms_bk * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                              "frac": hv["bytes"] / (ms_bk * 1e-3) / 1e9 / pk["hbm"],
                              "algorithmic_bytes": "expanded nodes x (2304 B row + 256 B adjacency) + scored candidates x (64 B code + 4 B scale), from the kernel's counters",
                              "kernel_ms_per_step": ms_bk, "kernel_share_of_step": ms_bk / hv["ms_per_step"],
                              "traffic": NCU_TRAFFIC["beam_l512_bytes_per_launch"][0] if (Lh, Wh, per_gpu) == (512, 4, 12_500_000) else None,
                              "traffic_note": "ncu capture " + NCU_TRAFFIC["beam_l512_bytes_per_launch"][1] + " (L = 512, W = 4, 12.5 M rows): 16 x the algorithmic "
                                              "bytes -- the visited-set tables (cleared per query, one cold sector per probe)"},
                 "sweep": {s: {kk: vv for kk, vv in v.items() if kk not in ("clocks", "bytes", "unit")} for s, v in variants.items()},
                 "greedy_exact": greedy, "greedy_clocks": gclocks,
                 "build": {"seconds": build_s, "points_per_s": n_local / build_s, "stats": bst, "rabitq_train_seconds": train_s, "rabitq_encode_seconds": enc_s}}
        rq.close()
        vl.close()
        del vl
        torch.cuda.empty_cache()

        # ---- CPU baseline of the graph path: the oracle's greedy_search and RabitQ beam search on a separate, smaller index
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            O.build()
            n_cpu = min(args.cpu_sample_graph_rows, per_gpu) // data.chunk * data.chunk or min(args.cpu_sample_graph_rows, per_gpu)
            svl, smed, _, sbuild_s = build_index(0, n_cpu, 100)
            x_host = np.empty((n_cpu, D), np.float16)
            for c0 in range(0, n_cpu, data.chunk):
                x_host[c0:c0 + data.chunk] = data.rows(c0, data.chunk)[: n_cpu - c0].cpu().numpy()
            adj, deg = svl.get_graph()
            ns = min(args.cpu_sample_graph_queries, nq)
            qs = q16_host.numpy()[:ns]
            og = O.IndexGraph(n_cpu, adj.shape[1])
            og.set(adj, deg)
            ocfg = O.make_config(r=R, l=L, maxc=750)
            O.greedy_search_batch(smed, qs[:32], x_host, og, ocfg)
            t0 = time.perf_counter()
            o_ids, o_sc, o_len, o_dist = O.greedy_search_batch(smed, qs, x_host, og, ocfg)
            dt_g = time.perf_counter() - t0
            s_ids = torch.empty((ns, L), dtype=torch.int32, device=dev)
            s_sc = torch.empty((ns, L), dtype=torch.int64, device=dev)
            s_len = torch.empty(ns, dtype=torch.int32, device=dev)
            s_dist = torch.empty(ns, dtype=torch.int64, device=dev)
            dk.set_graph_mode(dk.GRAPH_MODE_WARP)
            dk.greedy_search_dev(svl, q16.data_ptr(), ns, L, smed, s_ids.data_ptr(), s_sc.data_ptr(), s_len.data_ptr(), s_dist.data_ptr(), stream)
            dk.greedy_search_check(svl, ns)
            same_g = bool(np.array_equal(o_ids, s_ids.cpu().numpy().view(np.uint32)) and np.array_equal(o_sc, s_sc.cpu().numpy()))
            graph["greedy_exact"]["cpu_baseline"] = {"value": ns / dt_g, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                                                     "sample": f"{ns} of the {nq} queries on a separate {n_cpu}-row index of the same data (graph built on the GPU), "
                                                               f"oracle greedy_search L={L}, OpenMP over queries",
                                                     "distances_per_query": float(o_dist.mean()), "gpu_results_bit_identical": same_g}
            # RabitQ beam on the CPU: one query per thread through the oracle's orc_beam_search_rabitq
            srq = dk.RabitQ.train(svl, sample_rows=100_000, output_dims=512, seed=11)
            srq.encode_index(svl, 0)
            codes_h, norms_h, dots_h = srq.quantize(x_host)
            scale_h = (norms_h * dots_h).astype(np.float32)
            offsets = np.zeros(n_cpu + 1, np.uint64)
            np.cumsum(deg, out=offsets[1:])
            csr = adj[np.arange(adj.shape[1])[None, :] < deg[:, None]]
            sqtm = torch.empty((ns, 513), dtype=torch.float32, device=dev)
            srq.query_dev(q32.data_ptr(), ns, sqtm.data_ptr(), stream)
            torch.cuda.synchronize()
            qtm_h = sqtm.cpu().numpy()
            rq_scale = float(np.float32(1.0 / np.sqrt(np.float64(D))))
            from concurrent.futures import ThreadPoolExecutor
            nsb = min(ns, 256)

            def one(i):
                return O.beam_search(x_host, csr, offsets, codes_h, None, smed, qs[i], Lh, Wh, code_scale=scale_h, rabitq_qtm=qtm_h[i], rabitq_scale=rq_scale)
            with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
                list(ex.map(one, range(min(nsb, 16))))
                t0 = time.perf_counter()
                outs = list(ex.map(one, range(nsb)))
                dt_b = time.perf_counter() - t0
            b_ids = torch.empty((nsb, k), dtype=torch.int32, device=dev)
            b_sc = torch.empty((nsb, k), dtype=torch.int64, device=dev)
            b_len = torch.empty(nsb, dtype=torch.int32, device=dev)
            b_cm = torch.empty(nsb, dtype=torch.int64, device=dev)
            b_pc = torch.empty(nsb, dtype=torch.int64, device=dev)
            dk.beam_search_dev(svl, q16.data_ptr(), nsb, Lh, Wh, smed, k, b_ids.data_ptr(), b_sc.data_ptr(), b_len.data_ptr(), b_cm.data_ptr(), b_pc.data_ptr(),
                               stream, d_qtm=sqtm.data_ptr(), rabitq=srq)
            dk.greedy_search_check(svl, nsb)
            dk.set_graph_mode(dk.GRAPH_MODE_AUTO)
            gi, gs = b_ids.cpu().numpy().view(np.uint32), b_sc.cpu().numpy()
            same_b = True
            for i, (vi, vs, _) in enumerate(outs):
                order = np.lexsort((np.arange(len(vs)), -vs))[:k]           # stable: score desc, visit order
                same_b &= bool(np.array_equal(vi[order], gi[i, : len(order)]) and np.array_equal(vs[order], gs[i, : len(order)]))
            graph["cpu_baseline"] = {"value": nsb / dt_b, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                                     "sample": f"{nsb} of the {nq} queries on a separate {n_cpu}-row index of the same data (graph + codes built on the GPU), "
                                               f"oracle beam search over RabitQ codes L={Lh} W={Wh}, one query per host thread",
                                     "gpu_results_bit_identical": same_b}
            srq.close()
            svl.close()
            del svl
            torch.cuda.empty_cache()
        full["graph"] = graph

    # ============================================================ C1: one text query -> text tower (batch 1) -> top-10 over 1k rows
    if "c1" in wl:
        wpath = os.path.join(tempfile.gettempdir(), f"mse_bench_text27_rank{rank}.msew")
        sd = random_openclip_text_state_dict(dev)
        mse_b200.weights.save_weights(wpath, sd, mse_b200.weights.config_for(sd))
        del sd
        tenc = mse_b200.Encoder(wpath, device=local_rank, max_batch=8)
        os.remove(wpath)
        ids_np, x16 = c1_inputs()
        cix = mse_b200.FlatIndex.from_f16(x16, device=local_rank)
        ids_h = torch.from_numpy(ids_np).pin_memory()
        ids_d = ids_h.to(dev)
        emb_d = torch.empty((1, D), dtype=torch.float16, device=dev)
        q_d = torch.zeros((1, D), dtype=torch.float32, device=dev)
        top_i = torch.empty((1, 10), dtype=torch.int32, device=dev)
        top_s = torch.empty((1, 10), dtype=torch.float32, device=dev)
        emb_h = np.empty((1, D), np.float16)
        out_i, out_s = np.empty((1, 10), np.uint32), np.empty((1, 10), np.float32)

        def c1_resident():
            tenc.encode_text_dev(ids_d.data_ptr(), 1, emb_d.data_ptr(), stream)
            q_d.zero_()
            mse_b200.check(mse_b200.lib().mse_query_accumulate_f16_dev(local_rank, emb_d.data_ptr(), 1.0, q_d.data_ptr(), 1, D, stream), "mse_query_accumulate_f16_dev")
            cix.search_dev(q_d.data_ptr(), 1, 10, top_i.data_ptr(), top_s.data_ptr(), stream)

        def c1_e2e():   # what a client of the reference does: embed over the clip_server boundary, then Index::search (src/main.rs:899-934)
            mse_b200.check(mse_b200.lib().mse_encode_text_ids(tenc._h, ids_h.data_ptr(), 1, emb_h.ctypes.data), "mse_encode_text_ids")
            qf = emb_h.astype(np.float32)
            mse_b200.check(mse_b200.lib().mse_search_flat(cix._h, qf.ctypes.data, 1, 10, out_i.ctypes.data, out_s.ctypes.data), "mse_search_flat")

        n_c1 = max(steps, 20)
        with ClockSampler(local_rank) as cs:
            ms_c1, launches_c1 = timed(c1_resident, max(warmup, 3), n_c1)
        cix.check()
        ms_c1_e2e, _ = timed(c1_e2e, 3, n_c1)
        ms_tower = kernel_ms(lambda: tenc.encode_text_dev(ids_d.data_ptr(), 1, emb_d.data_ptr(), stream), n_c1)
        c1_resident()
        cix.check()
        torch.cuda.synchronize()
        same = bool(np.array_equal(top_i.cpu().numpy().view(np.uint32), out_i))
        full["c1"] = {"metric": C1_METRIC, "value": 1e3 / ms_c1, "unit": "queries/s", "ms_per_step": ms_c1, "n_gpus": 1, "higher_is_better": True,
                      "dtype": "fp16 storage, fp32 accumulate; f32 query, fp64 flat scan",
                      "config": {"workload": "text_b1_top10_1k", "tokens": "synthetic ids (rng seed 1; no tokenizer model offline)", "index_rows": 1000, "k": 10,
                                 "l2_policy": "text tower weights (0.83 GB) larger than L2"},
                      "clocks": cs.summary(), "steps": n_c1,
                      "e2e": {"value": 1e3 / ms_c1_e2e, "unit": "queries/s", "h2d_bytes_per_step": 64 * 4 + D * 4, "d2h_bytes_per_step": D * 2 + 10 * 8, "ms_per_step": ms_c1_e2e},
                      "gpu_launches": launches_c1 / n_c1, "resident_and_e2e_ids_identical": same,
                      "roofline": {"kernel": "text tower forward at batch 1 (tmega::k_text_blocks: the 27 blocks as one persistent cooperative kernel, + embed / final LN / projection kernels)", "bound": "hbm",
                                   "achieved": TEXT_WEIGHT_BYTES / (ms_tower * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                   "frac": TEXT_WEIGHT_BYTES / (ms_tower * 1e-3) / 1e9 / pk["hbm"],
                                   "algorithmic_bytes": "0.826 GB of fp16 weights touched once per forward (SURVEY 8d C1)", "kernel_ms_per_step": ms_tower,
                                   "kernel_share_of_step": ms_tower / ms_c1, "traffic": NCU_TRAFFIC["text_blocks_bytes_per_launch"][0],
                                   "traffic_note": "k_text_blocks, ncu capture " + NCU_TRAFFIC["text_blocks_bytes_per_launch"][1] + ", not measured by this run"}}
        cix.close()
        tenc.close()

    # ============================================================ end to end (C5 at reduced scale): encode -> index -> build -> serve
    if "e2e" in wl:
        try:   # the default N = 8 run does this workload at a size it was never run at: a failure here must not cost the whole line
            B, nq_t, nq_i, Ls, k = 256, 500, 500, 64, 10
            n_img = (args.e2e_images_total // world if args.e2e_images_total else args.e2e_images) // B * B
            wpath = os.path.join(tempfile.gettempdir(), f"mse_bench_e2e_rank{rank}.msew")
            sd = random_openclip_state_dict(dev)
            sd.update(random_openclip_text_state_dict(dev))
            mse_b200.weights.save_weights(wpath, sd, mse_b200.weights.config_for(sd))
            del sd
            enc = mse_b200.Encoder(wpath, device=local_rank, max_batch=B)
            os.remove(wpath)
            vl = dk.VectorList(D, device=local_rank, id_base=rank * n_img)      # this rank's id range of the index under construction
            vl.reserve(n_img)
            feat = torch.empty((B, D), dtype=torch.float16, device=dev)

            def image_source(seed):
                gg = torch.Generator(device=dev).manual_seed(seed)
                base = torch.randint(0, 256, (B, 24, 24, 3), generator=gg, device=dev, dtype=torch.uint8)

                def make_batch():
                    # blocky pattern + per-pixel noise so that images (and their embeddings) differ from each other
                    up = base[torch.randperm(B, generator=gg, device=dev)].repeat_interleave(16, 1).repeat_interleave(16, 2).to(torch.int16)
                    noise = torch.randint(-40, 41, (B, 384, 384, 3), generator=gg, device=dev, dtype=torch.int16)
                    return (up + noise).clamp_(0, 255).to(torch.uint8).contiguous()
                return make_batch
            make_batch = image_source(6 + rank)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
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
            if world > 1:
                t = torch.tensor([enc_ms, build_s], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                enc_ms, build_s = float(t[0].item()), float(t[1].item())
            # serve: the same 500 text + 500 image queries on every rank (replicated), every shard searched, one all-gather + merge
            gid = torch.Generator(device="cpu").manual_seed(8)
            ids = torch.ones((nq_t, 64), dtype=torch.int32)
            for i in range(nq_t):
                Lt = int(torch.randint(3, 17, (1,), generator=gid))
                ids[i, :Lt] = torch.randint(2, 32000, (Lt,), generator=gid, dtype=torch.int32)
            ids_dev = ids.to(dev)
            nq = nq_t + nq_i
            q_feat = torch.empty((nq, D), dtype=torch.float16, device=dev)
            make_query_batch = image_source(66)
            q_imgs = [make_query_batch() for _ in range((nq_i + B - 1) // B)]
            o_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
            o_sc = torch.empty((nq, k), dtype=torch.int64, device=dev)
            res_host = torch.empty((nq, k), dtype=torch.int32).pin_memory()

            def embed_queries():
                for b0 in range(0, nq_t, B):
                    m = min(B, nq_t - b0)
                    enc.encode_text_dev(ids_dev[b0:b0 + m].data_ptr(), m, q_feat[b0:b0 + m].data_ptr(), stream)
                for bi, b0 in enumerate(range(0, nq_i, B)):
                    m = min(B, nq_i - b0)
                    enc.encode_image_dev(q_imgs[bi].data_ptr(), m, q_feat[nq_t + b0:nq_t + b0 + m].data_ptr(), stream)

            def serve():
                embed_queries()
                grp.graph_search_dev(vl, q_feat.data_ptr(), nq, Ls, med, k, o_ids.data_ptr(), o_sc.data_ptr(), 0, stream)
                res_host.copy_(o_ids, non_blocking=True)
                torch.cuda.current_stream().synchronize()

            serve()
            ms_serve, _ = timed(serve, 1, 3)
            dk.greedy_search_check(vl, nq)
            # recall@10 of the served answers against the exact top-10 over all shards
            q32 = q_feat.float().contiguous()
            gt_i = torch.empty((nq, k), dtype=torch.int32, device=dev)
            gt_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
            grp.flat_search_dev(vl, q32.data_ptr(), nq, k, gt_i.data_ptr(), gt_s.data_ptr(), stream)
            grp.check()
            rec = float((o_ids.long().unsqueeze(2) == gt_i.long().unsqueeze(1)).any(dim=2).float().mean().item())
            wall = time.perf_counter() - t_wall0
            full["e2e_pipeline"] = {"metric": "end to end: encode + index + build + serve (BASELINE configs[4])", "n_gpus": world,
                                    "config": {"workload": "e2e_encode_build_serve", "images_total": n_img * world, "images_per_gpu": n_img, "graph": "R 64, L 192, C 750, one sub-graph per GPU",
                                               "queries": "500 text (token ids) + 500 image, replicated, L = 64, top-10 merged over the shards",
                                               "note": "configs[4] names 1M images on 8 GPUs (>= 61 s of encoder work at the tensor roofline); default at N = 8: 1M / 8 per rank"},
                                    "encode": {"images_per_s": world * n_img / (enc_ms * 1e-3), "seconds": enc_ms * 1e-3},
                                    "build": {"seconds": build_s, "points_per_s": world * n_img / build_s, "stats": bst},
                                    "serve": {"queries_per_s": nq / (ms_serve * 1e-3), "ms_per_batch_of_1000": ms_serve, "recall_at_10": rec,
                                              "includes": "text tower (500) + image tower (500) on every rank + sharded graph search + gather/merge + D2H of top-10 ids"},
                                    "wall_seconds_total": wall}
            vl.close()
            enc.close()
        except Exception as ex:
            full["e2e_pipeline_error"] = f"{type(ex).__name__}: {ex}"
            print("e2e workload failed:", full["e2e_pipeline_error"], file=sys.stderr)

    # ============================================================ CPU baselines (rank 0, N = 1) and the output line
    comm = {"backend": "nccl" if world > 1 else None, **grp.info()}
    if world > 1 and nccl_log and rank == 0 and os.path.exists(nccl_log):
        init = [ln.strip() for ln in open(nccl_log, errors="replace") if "nranks" in ln][:4]
        comm["nccl_init_lines"] = init
        for ln in init:
            print(ln, file=sys.stderr)
    grp.close()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    if world == 1 and not args.no_cpu_baseline:
        if "tower" in wl:
            ips, dt, cores = cpu_reference_tower(args.cpu_sample_images, 2, 1)
            result["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                                      "sample": f"2 steps of {args.cpu_sample_images} images, full 27-block tower, transformers SiglipVisionModel fp32 "
                                                f"(stand-in for clip_server.py device=cpu; open_clip is not installable here)"}
        if "flat" in wl:
            qps, dt, cores = cpu_reference_flat(args.rows, args.k, 1, 1, args.cpu_sample_rows, args.cpu_sample_queries)
            full["search"]["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                                              "sample": f"{args.cpu_sample_queries} queries x {args.cpu_sample_rows}-row slice, scaled by rows to {args.rows}"}
        if "c1" in wl:
            qps, dt, cores = cpu_reference_c1(2, 1)
            full["c1"]["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                                          "sample": "2 queries: transformers SiglipTextModel fp32 (all cores) + C brute force over 1k rows"}
    if not result:   # a run without the tower: promote the first workload that ran, so the line still has the contract's keys
        for name in ("search", "graph", "c1", "e2e_pipeline"):
            if name in full:
                result = dict(full.pop(name), steps=steps, warmup=warmup, vs_baseline=None, data="synthetic")
                break
    result["comm"] = comm
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"bench_full_n{world}.json"), "w") as f:
        json.dump({**result, **full}, f, indent=1)

    # compact summaries, LAST in the line so that they survive in a truncated tail
    def brief(w, extra=()):
        r = w.get("roofline", {})
        b = {"value": round(w["value"], 1), "unit": w["unit"], "e2e": round(w["e2e"]["value"], 1), "ms_per_step": round(w["ms_per_step"], 3),
             "roofline": {"bound": r.get("bound"), "frac": round(r["frac"], 4) if r.get("frac") else None,
                          "kernel_ms_per_step": round(r["kernel_ms_per_step"], 3) if r.get("kernel_ms_per_step") else None, "traffic": r.get("traffic")},
             "cpu_baseline": round(w["cpu_baseline"]["value"], 2) if w.get("cpu_baseline") else None}
        for key in extra:
            if w.get(key) is not None:
                b[key] = w[key]
        return b
    result["details"] = f"gpurun_out/bench_full_n{world}.json"
    if "c1" in full:
        result["c1"] = brief(full["c1"])
    if "search" in full:
        result["search"] = brief(full["search"], ("sharded_ids_match", "all_gathers_per_step"))
        result["search"]["index_rows"] = args.rows
    if "graph" in full:
        g = full["graph"]
        result["graph"] = brief(g, ("recall_at_10", "recall_target_met", "L", "W", "merge_matches_independent_gather"))
        result["graph"]["index_rows"] = g["config"]["index_rows"]
        ge = g["greedy_exact"]
        result["graph"]["greedy_exact"] = {"value": round(ge["value"], 1), "recall_at_10": round(ge["recall_at_10"], 4), "L": ge["L"],
                                           "e2e": round(ge["e2e"]["value"], 1), "frac": round(ge["roofline"]["frac"], 4)}
        result["graph"]["recall_at_10"] = round(g["recall_at_10"], 4)
        result["graph"]["build_s"] = round(g["build"]["seconds"], 1)
    if "e2e_pipeline" in full:
        e = full["e2e_pipeline"]
        result["e2e_pipeline"] = {"images_total": e["config"]["images_total"], "encode_images_per_s": round(e["encode"]["images_per_s"], 1),
                                  "encode_s": round(e["encode"]["seconds"], 1), "build_s": round(e["build"]["seconds"], 2),
                                  "serve_queries_per_s": round(e["serve"]["queries_per_s"], 1), "serve_recall_at_10": round(e["serve"]["recall_at_10"], 4)}
    elif "e2e_pipeline_error" in full:
        result["e2e_pipeline"] = {"error": full["e2e_pipeline_error"][:200]}
    sys.stdout.flush()
    print(json.dumps(result), flush=True)


if __name__ == "__main__":
    main()
