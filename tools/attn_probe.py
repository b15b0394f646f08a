"""The tcgen05 attention kernel alone (C-ABI mse_debug_attention: random fp16 qkv, CUDA events): time per launch for the given
batch / debug-mode list; also the workload for ncu captures.  usage: attn_probe.py [batch] [mode ...]
modes (bit mask, profiling only): 0 shipping kernel, 1 skip ex2, 2 skip P stores, 4 skip PV MMAs, 8 skip S MMAs, 16 skip K/V TMA loads"""
import ctypes as C, sys
sys.path.insert(0, ".")
import mse_b200
l = mse_b200.lib()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
modes = [int(m) for m in sys.argv[2:]] or [0]
for mode in modes:
    ms = C.c_float()
    rc = l.mse_debug_attention(0, batch, 729, mode, 5, C.byref(ms))
    flop = 4.0 * 729 * 729 * 72 * 16 * batch
    print(f"mode {mode} rc {rc} ms {ms.value:.4f} TFLOP/s {flop / (ms.value * 1e-3) / 1e12:.1f} (x27 layers x batch 256/{batch}: {ms.value * 27 * 256 / batch:.1f} ms/step)", flush=True)
