"""Latency of the text tower at small batches (BASELINE configs[0]: one text query), CUDA events, 50 repetitions each.
At batch 1 the 27 blocks are weight-bandwidth bound: 0.826 GB of fp16 weights / HBM copy peak = the floor printed below."""
import json, os, sys, tempfile
sys.path.insert(0, ".")
import torch
import mse_b200
from bench import random_openclip_text_state_dict, peaks

dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
sd = random_openclip_text_state_dict(dev)
path = os.path.join(tempfile.gettempdir(), "mse_text27.msew")
mse_b200.weights.save_weights(path, sd, mse_b200.weights.config_for(sd))
del sd
enc = mse_b200.Encoder(path, device=0, max_batch=64)
os.remove(path)
out = {"floor_ms_weights_over_hbm": 0.826e9 / (peaks()["hbm"] * 1e9) * 1e3}
for B in ([int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 32, 64]):   # usage: text_latency.py [batch ...]
    ids = torch.randint(2, 32000, (B, 64), dtype=torch.int32, device=dev)
    feat = torch.empty((B, 1152), dtype=torch.float16, device=dev)
    for _ in range(5):
        enc.encode_text_dev(ids.data_ptr(), B, feat.data_ptr(), stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50):
        enc.encode_text_dev(ids.data_ptr(), B, feat.data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    if B <= 16 and not os.environ.get('MSE_NO_GRAPH'):
        out[f"batch_{B}"] = {"ms": ms, "texts_per_s": B / ms * 1e3, "graph": True}
        continue
    enc.profile(True); enc.encode_text_dev(ids.data_ptr(), B, feat.data_ptr(), stream); st = enc.stats(); enc.profile(False)
    out[f"batch_{B}"] = {"ms": ms, "texts_per_s": B / ms * 1e3, "gemm_ms": st["gemm_ns"] * 1e-6, "gemm_launches": st["gemm_launches"], "attn_ms": st["attn_ns"] * 1e-6,
                         "launches": st["launches"]}
print(json.dumps(out))
