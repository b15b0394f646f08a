"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, exports every symbol the header
declares, and refuses to compute without a GPU (no fallback)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import HAS_GPU


def test_library_exports_every_header_symbol(mse):
    from mse_b200 import _lib
    mse.build()
    l = mse.lib()
    syms = _lib.header_symbols()
    assert len(syms) >= 15
    missing = [s for s in syms if not hasattr(l, s)]
    assert not missing, missing
    # the Python binding table covers the header exactly
    assert sorted(_lib._SIGS) == syms


def test_header_is_plain_c():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = "#include \"mse_b200.h\"\nint main(void){return MSE_OK;}\n"
    p = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-x", "c", "-", "-o", "/dev/null"],
                       input=src, text=True, capture_output=True)
    assert p.returncode == 0, p.stderr


def test_library_built_for_sm100a_with_tcgen05(mse):
    out = subprocess.run(["cuobjdump", "-lelf", mse.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", mse.lib_path()], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(mse):
    with pytest.raises(mse.MseError) as e:
        mse.FlatIndex(1152)
    assert "no CPU fallback" in str(e.value)
    q = np.zeros(64, np.float16)
    out = np.zeros(1, np.int64)
    rc = mse.lib().mse_fast_dot_batch(0, q.ctypes.data, q.ctypes.data, 1, 64, None, 1, out.ctypes.data)
    assert rc == -2


def test_argument_validation_without_device(mse):
    l = mse.lib()
    h = C.c_void_p()
    assert l.mse_index_create(None, 0, 7, 0, 0, C.byref(h)) == -1     # d % 8
    assert l.mse_index_create(None, 5, 64, 0, 0, C.byref(h)) == -1    # NULL rows with n > 0
    assert l.mse_search_flat(None, None, 1, 1, None, None) == -1
    assert b"NULL" in l.mse_last_error()
    assert l.mse_index_ntotal(None) == 0


def test_graph_entries_validate_arguments_without_device(mse):
    """Every graph / codec entry added behind the diskann boundary rejects NULL handles and buffers before touching a device
    (error code -1 = MSE_ERR_INVALID, message through mse_last_error) -- no crash, no silent success."""
    l = mse.lib()
    assert l.mse_search_graph_dev(None, None, 1, 8, None, 0, 0, 0, None, None, None, None, None) == -1
    assert l.mse_search_graph_check(None, 1) == -1
    assert l.mse_search_graph_set_mode(7) == -1 and b"mode" in l.mse_last_error()
    assert l.mse_search_graph_set_mode(0) == 0
    assert l.mse_search_beam_dev(None, None, None, None, 0, 0, None, 1, 8, 1, None, 0, 256, 10, None, None, None, None, None, None) == -1
    assert l.mse_search_beam_scaled(None, None, None, None, None, 1, 8, 1, None, 0, 256, None, None, None, 1, None, None) == -1
    assert l.mse_dedup_topk_dev(None, 1, 0.95, 10, None, None, None, None, None) == -1
    assert l.mse_index_set_code_scales(None, None) == -1
    assert l.mse_index_encode_rabitq(None, None, 0) == -1
    assert l.mse_index_robust_stitch(None, None, None, 0) == -1
    assert l.mse_rabitq_preprocess_query(None, None, 1, None, None) == -1
    assert l.mse_rabitq_query_dev(None, None, 1, None, None) == -1


def test_r03_entries_validate_arguments_without_device(mse):
    """The entries added in the second half of round 2 (shard k-means / assignment, resize) reject NULL handles, NULL buffers and
    out-of-range sizes before touching a device."""
    l = mse.lib()
    cnt = np.zeros(4, np.uint32)
    assert l.mse_kmeans_assign(None, None, 2, 2, 1, cnt.ctypes.data, None) == -1
    assert l.mse_kmeans_anneal(None, 2, 2, 1, 0, None, None, None) == -1
    st = np.zeros(2, np.uint64)
    assert l.mse_shard_assign(None, None, 2, 2, 0.2, st.ctypes.data, st.ctypes.data, None) == -1
    px = np.zeros((2, 2, 3), np.uint8)
    out = np.zeros((2, 2, 3), np.uint8)
    assert l.mse_resize_rgb_u8(0, px.ctypes.data, 2, 2, 2, 2, 7, out.ctypes.data) == -1      # filter out of range
    assert l.mse_resize_rgb_u8(0, px.ctypes.data, 2, 2, 2, 2, 0, None) == -1                  # NULL output
    assert l.mse_encode_images_resized(None, None, None, None, 1, None) == -1
