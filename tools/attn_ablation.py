import ctypes as C, sys
sys.path.insert(0, '.')
import mse_b200
l = mse_b200.lib()
for mode in [0, 1, 2, 3, 4, 8, 12, 16, 28, 31, 7]:
    ms = C.c_float()
    rc = l.mse_debug_attention(0, 64, 729, mode, 5, C.byref(ms))
    print("mode", mode, "rc", rc, "ms", round(ms.value, 4), flush=True)
