"""Graph path measurements on one B200: GPU Vamana build, greedy / beam (PQ) search QPS and recall@10 vs flat ground truth."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import mse_b200
from mse_b200 import diskann as dk

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
R, Lb, maxc = 64, 192, 750
nq, Ls, W = 4096, 64, 4
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(4)
ncl = 4096
cent = torch.randn((ncl, 1152), generator=g, device=dev); cent /= cent.norm(dim=1, keepdim=True)
def draw(m, seed):
    gg = torch.Generator(device=dev).manual_seed(seed)
    a = torch.randint(0, ncl, (m,), generator=gg, device=dev)
    x = cent[a] + 0.3 * torch.randn((m, 1152), generator=gg, device=dev) / 1152 ** 0.5
    return (x / x.norm(dim=1, keepdim=True)).to(torch.float16)
x = draw(n, 4).cpu().numpy(); q = draw(nq, 5).cpu().numpy()
vl = dk.VectorList.from_f16s(x)
t0 = time.time(); dk.random_fill_graph(vl, R, seed=1); med = dk.medioid(vl); t_med = time.time() - t0
cfg = dk.IndexBuildConfig(r=R, l=Lb, maxc=maxc)
t0 = time.time(); st = dk.build_graph(vl, med, cfg, seed=7); t_build = time.time() - t0
adj, deg = vl.get_graph()
# ground truth by flat search
sc, lab = vl.search(q.astype(np.float32), 10)
cfg_s = dk.IndexBuildConfig(r=R, l=Ls, maxc=maxc)
for _ in range(2): res = dk.greedy_search(vl, q, med, cfg_s)
t0 = time.time(); res = dk.greedy_search(vl, q, med, cfg_s); t_g = time.time() - t0
rec_g = np.mean([len(set(res.ids[i][:10].tolist()) & set(lab[i].tolist())) / 10 for i in range(nq)])
# PQ codec: random orthogonal rotation, centroids sampled from the data
T = torch.linalg.qr(torch.randn((1152, 1152), generator=g, device=dev))[0].cpu().numpy().astype(np.float32)
pq = dk.ProductQuantizer(x[np.random.default_rng(0).choice(n, 256, replace=False)].astype(np.float32), T, 18)
t0 = time.time(); codes = pq.quantize_batch(x.astype(np.float32)); t_enc = time.time() - t0
vl.set_pq_codes(codes)
luts = pq.preprocess_query(q.astype(np.float32))
for _ in range(2): out, cmps, pqc = dk.beam_search(vl, q, luts, med, Ls, W, out_cap=2048)
t0 = time.time(); out, cmps, pqc = dk.beam_search(vl, q, luts, med, Ls, W, out_cap=2048); t_b = time.time() - t0
def top10(ids, s):
    o = np.argsort(-s, kind="stable")[:10]; return set(ids[o].tolist())
rec_b = np.mean([len(top10(*out[i]) & set(lab[i].tolist())) / 10 for i in range(nq)])
print(json.dumps({"n": n, "R": R, "L_build": Lb, "maxc": maxc, "medioid_s": t_med, "build_s": t_build, "build_points_per_s": n / t_build, "build_stats": st,
                  "deg_mean": float(deg.mean()), "greedy_L64_qps_e2e": nq / t_g, "greedy_recall10": float(rec_g), "greedy_dist_per_q": float(res.distances.mean()),
                  "pq_encode_s": t_enc, "beam_W4_L64_qps_e2e": nq / t_b, "beam_recall10": float(rec_b), "beam_cmps": float(cmps.mean()), "beam_pq_cmps": float(pqc.mean())}))
