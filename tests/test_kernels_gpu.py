"""GPU parity of the building-block kernels: fast_dot (bit-exact i64) and the tcgen05 GEMM (fp32 tolerance)."""
import ctypes as C

import numpy as np
import pytest

from helpers import index_f16

pytestmark = pytest.mark.gpu


def test_fast_dot_bit_exact(mse, oracle):
    """diskann/src/vector.rs:255-306: i64 fixed-point scores must equal the AVX2 oracle bit for bit."""
    x = index_f16(3, 5000)
    rng = np.random.default_rng(0)
    x[10] *= np.float16(40.0)
    x[11, ::5] = np.float16(6e-8)
    ids = rng.integers(0, 5000, 20000).astype(np.uint32)
    out = np.empty(ids.size, np.int64)
    rc = mse.lib().mse_fast_dot_batch(0, x[10].ctypes.data, x.ctypes.data, 5000, 1152, ids.ctypes.data, ids.size, out.ctypes.data)
    assert rc == 0, mse.last_error()
    want = oracle.fast_dot_batch(x[10], x)[ids]
    assert np.array_equal(out, want)
    out2 = np.empty(5000, np.int64)
    rc = mse.lib().mse_fast_dot_batch(0, x[0].ctypes.data, x.ctypes.data, 5000, 1152, None, 5000, out2.ctypes.data)
    assert rc == 0 and np.array_equal(out2, oracle.fast_dot_batch(x[0], x))
    # d = 64 (the smallest legal size, vector.rs:197)
    y = index_f16(4, 100, d=64)
    out3 = np.empty(100, np.int64)
    assert mse.lib().mse_fast_dot_batch(0, y[1].ctypes.data, y.ctypes.data, 100, 64, None, 100, out3.ctypes.data) == 0
    assert np.array_equal(out3, oracle.fast_dot_batch(y[1], y))


def _gemm(mse, a, b, bias=None, act=0):
    M, K = a.shape
    N = b.shape[0]
    c = np.empty((M, N), np.float32)
    rc = mse.lib().mse_gemm_f16_tn(0, a.ctypes.data, b.ctypes.data, M, N, K, None if bias is None else bias.ctypes.data, act, c.ctypes.data)
    assert rc == 0, mse.last_error()
    return c


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 128, 64), (128, 256, 1152), (256, 512, 128), (200, 300, 72),
                                   (729, 1152, 1152), (1000, 4304, 1152), (500, 1152, 4304), (64, 3456, 1152), (129, 257, 200)])
def test_gemm_vs_fp32_reference(mse, M, N, K):
    """Tolerance: inputs are exact fp16, products exact in fp32, so only fp32 accumulation order differs:
    |err| <= 1e-4 * sum|a||b| per output (K <= 4304)."""
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    a = (rng.standard_normal((M, K)) / np.sqrt(K)).astype(np.float16)
    b = rng.standard_normal((N, K)).astype(np.float16)
    c = _gemm(mse, a, b)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    bound = 1e-4 * (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64).T) + 1e-6
    assert np.isfinite(c).all()
    assert (np.abs(c - ref) <= bound).all(), float(np.abs(c - ref).max())


def test_gemm_bias_gelu(mse):
    from math import erf
    rng = np.random.default_rng(1)
    a = (rng.standard_normal((130, 256)) / 16).astype(np.float16)
    b = rng.standard_normal((200, 256)).astype(np.float16)
    bias = rng.standard_normal(200).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64).T + bias
    c1 = _gemm(mse, a, b, bias, act=1)
    g = 0.5 * ref * (1 + np.vectorize(erf)(ref / np.sqrt(2)))
    assert np.abs(c1 - g).max() < 2e-5
    c2 = _gemm(mse, a, b, bias, act=2)
    t = 0.5 * ref * (1 + np.tanh(np.sqrt(2 / np.pi) * (ref + 0.044715 * ref ** 3)))
    assert np.abs(c2 - t).max() < 2e-5
