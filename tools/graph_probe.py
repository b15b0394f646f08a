"""Graph-search kernels alone on one B200 (CUDA events, resident queries): builds an index of `rows` rows of bench.py's C4 data on
the GPU, then times the exact greedy_search ("g:L") and the RabitQ beam search ("b:L:W") variants named on the command line and
reports q/s, recall@10 against the flat ground truth and the kernels' own counters.  Also the workload for ncu captures:
  ncu --set full --import-source on -k regex:k_beam_search_wq -c 1 -o gpurun_out/x python tools/graph_probe.py 1000000 families b:64:4
usage: graph_probe.py [rows] [families|mixture|latent] [variant ...] [--reps N] [--cuda-profiler]"""
import json, sys, time
sys.path.insert(0, ".")
import torch
import mse_b200
from mse_b200 import diskann as dk
from bench import GraphData, D

args = [a for a in sys.argv[1:] if not a.startswith("--")]
reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 5
args = [a for a in args if a != str(reps) or "--reps" not in sys.argv]
n = int(args[0]) if args else 1_000_000
kind = args[1] if len(args) > 1 else "families"
variants = args[2:] or ["g:64", "b:64:4", "b:256:4"]
R, nq, k = 64, 4096, 10
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
data = GraphData(kind, dev)
vl = dk.VectorList(D)
vl.reserve(n)
for c0 in range(0, n, data.chunk):
    xb = data.rows(c0, data.chunk)[: n - c0].contiguous()
    vl.add_f16_dev(xb.data_ptr(), xb.shape[0], stream)
    del xb
q16 = data.queries(nq, n)
q32 = q16.float().contiguous()
torch.cuda.synchronize()
t0 = time.time()
dk.random_fill_graph(vl, R, seed=1)
med = dk.medioid(vl)
bst = dk.build_graph(vl, med, dk.IndexBuildConfig(r=R, l=192, maxc=750), seed=7)
out = {"rows": n, "data": kind, "build_s": time.time() - t0, "build_stats": bst}
gt = torch.empty((nq, k), dtype=torch.int32, device=dev)
gts = torch.empty((nq, k), dtype=torch.float32, device=dev)
vl.search_dev(q32.data_ptr(), nq, k, gt.data_ptr(), gts.data_ptr(), stream)
vl.check()
rq = dk.RabitQ.train(vl, sample_rows=100_000, output_dims=512, seed=11)
rq.encode_index(vl, 0)
qtm = torch.empty((nq, 513), dtype=torch.float32, device=dev)
rq.query_dev(q32.data_ptr(), nq, qtm.data_ptr(), stream)


def recall(ids):
    return float((ids[:, :k].long().unsqueeze(2) == gt.long().unsqueeze(1)).any(dim=2).float().mean().item())


def timed(fn):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if "--cuda-profiler" in sys.argv:      # ncu --profile-from-start off: capture the search kernels, not the build's
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
for v in variants:
    p = v.split(":")
    if p[0] == "g":                      # g:L[:pf=<rows prefetched>][:hs=<visited-set scale>]  (tuning aids of csrc/graph.cu)
        L = int(p[1])
        import os
        for kv in p[2:]:
            key, val = kv.split("=")
            os.environ[{"pf": "MSE_GREEDY_PF", "hs": "MSE_GREEDY_HASH_SCALE"}[key]] = val
        ids = torch.empty((nq, L), dtype=torch.int32, device=dev); sc = torch.empty((nq, L), dtype=torch.int64, device=dev)
        ln = torch.empty(nq, dtype=torch.int32, device=dev); dist = torch.empty(nq, dtype=torch.int64, device=dev)
        ms = timed(lambda: dk.greedy_search_dev(vl, q16.data_ptr(), nq, L, med, ids.data_ptr(), sc.data_ptr(), ln.data_ptr(), dist.data_ptr(), stream))
        try:
            dk.greedy_search_check(vl, nq)
        except Exception as e:   # e.g. a visited-set overflow with a tuning override
            out[v] = {"ms": ms, "error": str(e)}
            print(json.dumps({v: out[v]}), flush=True)
            continue
        nd = float(dist.double().sum().item())
        out[v] = {"ms": ms, "qps": nq / ms * 1e3, "recall_at_10": recall(ids), "distances_per_query": nd / nq, "row_gbs": nd * D * 2 / (ms * 1e-3) / 1e9}
    else:
        L, W = int(p[1]), int(p[2])                      # b:L:W[:hp=<visited_adjacent slots per unit of L>]
        import os
        for kv in p[3:]:
            key, val = kv.split("=")
            os.environ[{"hp": "MSE_BEAM_HASH_PER_L"}[key]] = val
        ti = torch.empty((nq, k), dtype=torch.int32, device=dev); ts = torch.empty((nq, k), dtype=torch.int64, device=dev)
        tl = torch.empty(nq, dtype=torch.int32, device=dev); cm = torch.empty(nq, dtype=torch.int64, device=dev); pc = torch.empty(nq, dtype=torch.int64, device=dev)
        ms = timed(lambda: dk.beam_search_dev(vl, q16.data_ptr(), nq, L, W, med, k, ti.data_ptr(), ts.data_ptr(), tl.data_ptr(), cm.data_ptr(), pc.data_ptr(),
                                              stream, d_qtm=qtm.data_ptr(), rabitq=rq))
        try:
            dk.greedy_search_check(vl, nq)
        except Exception as e:
            out[v] = {"ms": ms, "error": str(e)}
            print(json.dumps({v: out[v]}), flush=True)
            continue
        ex, co = float(cm.double().sum().item()), float(pc.double().sum().item())
        out[v] = {"ms": ms, "qps": nq / ms * 1e3, "recall_at_10": recall(ti), "exact_rows_per_query": ex / nq, "codes_per_query": co / nq,
                  "algorithmic_gbs": (ex * (D * 2 + R * 4) + co * 68) / (ms * 1e-3) / 1e9}
    print(json.dumps({v: out[v]}), flush=True)
print(json.dumps(out))
