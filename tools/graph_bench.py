"""Graph search schedules side by side on one B200: CTA-per-query vs warp-per-query greedy search (kernel time by CUDA events,
achieved gather bandwidth), on a GPU-built Vamana graph over a clustered synthetic index.  usage: graph_bench.py [rows] [L]"""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import mse_b200
from mse_b200 import diskann as dk

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 64
D, R, nq = 1152, 64, 4096
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device=dev).manual_seed(4)
cent = torch.randn((4096, D), generator=g, device=dev); cent /= cent.norm(dim=1, keepdim=True)
def draw(m, seed):
    gg = torch.Generator(device=dev).manual_seed(seed)
    a = torch.randint(0, 4096, (m,), generator=gg, device=dev)
    x = cent[a] + 0.3 * torch.randn((m, D), generator=gg, device=dev) / D ** 0.5
    return (x / x.norm(dim=1, keepdim=True)).to(torch.float16).contiguous()
vl = dk.VectorList(D)
vl.reserve(n)
for c0 in range(0, n, 1 << 18):
    m = min(1 << 18, n - c0)
    xb = draw(m, 4_000_003 + c0); vl.add_f16_dev(xb.data_ptr(), m, stream); del xb
q16 = draw(nq, 5)
torch.cuda.synchronize()
out = {"rows": n, "L": L, "queries": nq}
for mode, name in ((1, "cta_per_query"), (2, "warp_per_query")):
    dk.set_graph_mode(mode)
    dk.random_fill_graph(vl, R, seed=1)
    med = dk.medioid(vl)
    t0 = time.time(); st = dk.build_graph(vl, med, dk.IndexBuildConfig(r=R, l=192, maxc=750), seed=7); tb = time.time() - t0
    ids = torch.empty((nq, L), dtype=torch.int32, device=dev); sc = torch.empty((nq, L), dtype=torch.int64, device=dev)
    ln = torch.empty(nq, dtype=torch.int32, device=dev); dist = torch.empty(nq, dtype=torch.int64, device=dev)
    def run():
        dk.greedy_search_dev(vl, q16.data_ptr(), nq, L, med, ids.data_ptr(), sc.data_ptr(), ln.data_ptr(), dist.data_ptr(), stream)
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    dk.greedy_search_check(vl, nq)
    nd = float(dist.double().sum().item())
    out[name] = {"build_s": tb, "build_points_per_s": n / tb, "build_distances": st["distances"], "search_ms": ms, "qps": nq / ms * 1e3,
                 "distances_per_query": nd / nq, "gather_gbs": nd * D * 2 / (ms * 1e-3) / 1e9}
dk.set_graph_mode(0)
print(json.dumps(out))
