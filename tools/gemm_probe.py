"""Linear-layer GEMM alone (C-ABI mse_debug_gemm: random fp16 operands, fp16 output, CUDA events): time and weight-stream rate for
the text tower's four layer shapes at a given row count.  usage: gemm_probe.py [rows ...]   (env MSE_GEMM_NO_PANEL=1: r01 skinny kernel)"""
import ctypes as C, sys
sys.path.insert(0, ".")
import mse_b200
l = mse_b200.lib()
shapes = {"qkv": (3456, 1152, 2), "proj": (1152, 1152, 10), "fc1": (4304, 1152, 6), "fc2": (1152, 4304, 10)}
for M in [int(a) for a in sys.argv[1:]] or [64]:
    tot = 0.0
    for name, (N, K, mode) in shapes.items():
        ms = C.c_float()
        rc = l.mse_debug_gemm(0, M, N, K, 0, mode, 50, C.byref(ms))
        tot += ms.value
        print(f"M={M:4d} {name:5s} N={N} K={K} rc={rc} us={ms.value * 1e3:7.2f} weights GB/s={N * K * 2 / (ms.value * 1e-3) / 1e9:7.1f}", flush=True)
    print(f"M={M:4d} one block's four GEMMs: {tot * 1e3:.1f} us (x27 = {tot * 27:.3f} ms)")
