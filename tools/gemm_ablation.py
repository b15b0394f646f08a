import ctypes as C, os, sys
sys.path.insert(0, '.')
import mse_b200
l = mse_b200.lib()
T = 256 * 729
shapes = {"qkv": (T, 3456, 1152, 2), "proj": (T, 1152, 1152, 10), "fc1": (T, 4304, 1152, 6), "fc2": (T, 1152, 4304, 10)}
for name, (M, N, K, mode) in shapes.items():
    for direct in (0, 1):
        if direct: os.environ["MSE_GEMM_DIRECT_STORE"] = "1"
        else: os.environ.pop("MSE_GEMM_DIRECT_STORE", None)
        ms = C.c_float()
        rc = l.mse_debug_gemm(0, M, N, K, 256 if direct else 0, mode, 10, C.byref(ms))
        tf = 2.0 * M * N * K / (ms.value * 1e-3) / 1e12 if rc == 0 else 0
        print(f"{name:5s} {'direct   ' if direct else 'tma-store'} mode={mode:2d} rc={rc} ms={ms.value:.4f} TF/s={tf:.0f}", flush=True)
