"""Flat search at small query counts over a 10M x 1152 fp16 index (HBM-bound regime: one 23 GB pass per batch), CUDA events."""
import json, sys
sys.path.insert(0, ".")
import torch
import mse_b200
from bench import peaks

D, rows, k = 1152, int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 10
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
ix = mse_b200.FlatIndex(D)
ix.reserve(rows)
gen = torch.Generator(device=dev)
for c0 in range(0, rows, 1 << 19):
    m = min(1 << 19, rows - c0)
    gen.manual_seed(2_000_006 + c0)
    xb = torch.randn((m, D), generator=gen, device=dev)
    xb = (xb / xb.norm(dim=1, keepdim=True)).to(torch.float16).contiguous()
    ix.add_f16_dev(xb.data_ptr(), m, stream)
    del xb
out = {"rows": rows, "hbm_peak_gbs": peaks()["hbm"], "one_pass_ms_at_peak": rows * D * 2 / (peaks()["hbm"] * 1e9) * 1e3}
for nq in (1, 2, 4, 16, 64, 128):
    q = torch.randn((nq, D), device=dev)
    q = (q / q.norm(dim=1, keepdim=True)).contiguous()
    ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
    for _ in range(2):
        ix.search_dev(q.data_ptr(), nq, k, ids.data_ptr(), sc.data_ptr(), stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        ix.search_dev(q.data_ptr(), nq, k, ids.data_ptr(), sc.data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    st = ix.stats()
    out[f"nq_{nq}"] = {"ms": ms, "queries_per_s": nq / ms * 1e3, "index_gbs": rows * D * 2 / (ms * 1e-3) / 1e9,
                       "frac_of_hbm_peak_one_pass": rows * D * 2 / (ms * 1e-3) / 1e9 / peaks()["hbm"], "exact_queries": st["exact_queries"], "tensor_queries": st["tensor_queries"]}
print(json.dumps(out))
