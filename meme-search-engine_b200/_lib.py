"""ctypes binding of libmse_b200.so (the declarations mirror include/mse_b200.h one to one)."""
from __future__ import annotations

import ctypes as C
import json
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class MseError(RuntimeError):
    pass


def lib_path() -> str:
    return os.path.join(_HERE, "libmse_b200.so")


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into libmse_b200.so (in-tree)."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j", str(os.cpu_count() or 4)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise MseError("building libmse_b200.so failed")
    return lib_path()


def header_symbols() -> list[str]:
    """Every function include/mse_b200.h declares."""
    hdr = os.path.join(os.path.dirname(_HERE), "include", "mse_b200.h")
    txt = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(mse_[a-z0-9_]+)\s*\(", txt)))


_vp, _u64, _u32, _i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
_SIGS = {
    "mse_last_error": (C.c_char_p, []),
    "mse_device_info": (_i32, [_i32, C.c_char_p, C.c_size_t]),
    "mse_launch_count": (_u64, []),
    "mse_fast_dot_batch": (_i32, [_i32, _vp, _vp, _u64, _u32, _vp, _u64, _vp]),
    "mse_index_create": (_i32, [_vp, _u64, _u32, _i32, _u32, C.POINTER(_vp)]),
    "mse_index_add": (_i32, [_vp, _vp, _u64]),
    "mse_index_add_f16": (_i32, [_vp, _vp, _u64]),
    "mse_index_add_f16_dev": (_i32, [_vp, _vp, _u64, _vp]),
    "mse_index_reserve": (_i32, [_vp, _u64]),
    "mse_index_ntotal": (_u64, [_vp]),
    "mse_index_dim": (_u32, [_vp]),
    "mse_index_device": (_i32, [_vp]),
    "mse_index_vectors_dev": (_vp, [_vp]),
    "mse_index_destroy": (None, [_vp]),
    "mse_search_flat": (_i32, [_vp, _vp, _u32, _u32, _vp, _vp]),
    "mse_search_flat_dev": (_i32, [_vp, _vp, _u32, _u32, _vp, _vp, _vp]),
    "mse_search_flat_check": (_i32, [_vp, _vp]),
    "mse_query_accumulate_f16_dev": (_i32, [_i32, _vp, C.c_float, _vp, _u32, _u32, _vp]),
    "mse_search_flat_stats": (_i32, [_vp, _vp]),
    "mse_shard_range": (_i32, [_u64, _i32, _i32, _vp, _vp]),
    "mse_shard_group_unique_id": (_i32, [_vp]),
    "mse_shard_group_create": (_i32, [_vp, _i32, _i32, _i32, C.POINTER(_vp)]),
    "mse_shard_group_info": (_i32, [_vp, _vp]),
    "mse_shard_group_gathers": (_u64, [_vp]),
    "mse_shard_group_destroy": (None, [_vp]),
    "mse_search_flat_sharded_dev": (_i32, [_vp, _vp, _vp, _u32, _u32, _vp, _vp, _vp]),
    "mse_search_sharded_check": (_i32, [_vp, _vp]),
    "mse_search_graph_sharded_dev": (_i32, [_vp, _vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp]),
    "mse_search_beam_sharded_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _vp, _u32, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp]),
    "mse_search_flat_set_mode": (_i32, [_vp, _i32]),
    "mse_search_flat_profile": (_i32, [_vp, _i32]),
    "mse_merge_topk_dev": (_i32, [_i32, _vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp]),
    "mse_index_set_graph": (_i32, [_vp, _vp, _vp, _u32]),
    "mse_index_get_graph": (_i32, [_vp, _vp, _vp, _vp]),
    "mse_index_set_pq_codes": (_i32, [_vp, _vp, _u32]),
    "mse_index_set_descriptors": (_i32, [_vp, _vp, _u32, _vp]),
    "mse_search_graph": (_i32, [_vp, _vp, _u32, _u32, _vp, _u32, _i32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32]),
    "mse_search_graph_dev": (_i32, [_vp, _vp, _u32, _u32, _vp, _u32, _i32, _u32, _vp, _vp, _vp, _vp, _vp]),
    "mse_search_graph_check": (_i32, [_vp, _u32]),
    "mse_search_graph_set_mode": (_i32, [_i32]),
    "mse_search_beam": (_i32, [_vp, _vp, _vp, _vp, _u32, _u32, _u32, _vp, _u32, _i32, _u32, _vp, _vp, _vp, _u32, _vp, _vp]),
    "mse_search_beam_scaled": (_i32, [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _vp, _u32, _u32, _vp, _vp, _vp, _u32, _vp, _vp]),
    "mse_index_set_code_scales": (_i32, [_vp, _vp]),
    "mse_dedup_topk_dev": (_i32, [_vp, _u32, C.c_float, _u32, _vp, _vp, _vp, _vp, _vp]),
    "mse_search_beam_dev": (_i32, [_vp, _vp, _vp, _vp, _u32, _u32, _vp, _u32, _u32, _u32, _vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mse_scores_i64": (_i32, [_vp, _vp, _vp]),
    "mse_robust_prune": (_i32, [_vp, _u32, _vp, _vp, _u32, _vp, _vp, _vp]),
    "mse_index_random_fill_graph": (_i32, [_vp, _u32, _u64]),
    "mse_index_medioid": (_i32, [_vp, _vp]),
    "mse_index_robust_stitch": (_i32, [_vp, _vp, _vp, _u64]),
    "mse_index_build_vamana": (_i32, [_vp, _u32, _vp, _u64, _u32, _vp]),
    "mse_pq_create": (_i32, [_vp, _vp, _u32, _u32, _u32, _i32, C.POINTER(_vp)]),
    "mse_pq_load": (_i32, [_vp, C.c_size_t, _i32, C.POINTER(_vp)]),
    "mse_pq_info": (_i32, [_vp, _vp]),
    "mse_pq_apply_transform": (_i32, [_vp, _vp, _u64, _vp]),
    "mse_pq_encode": (_i32, [_vp, _vp, _u64, _vp]),
    "mse_pq_preprocess_query": (_i32, [_vp, _vp, _u32, _vp]),
    "mse_pq_adc": (_i32, [_vp, _vp, _vp, _u64, _vp]),
    "mse_pq_destroy": (None, [_vp]),
    "mse_rabitq_create": (_i32, [_vp, _vp, _u32, _u32, _i32, C.POINTER(_vp)]),
    "mse_rabitq_load": (_i32, [_vp, C.c_size_t, _i32, C.POINTER(_vp)]),
    "mse_rabitq_train": (_i32, [_vp, _u64, _u32, _u64, C.POINTER(_vp)]),
    "mse_rabitq_info": (_i32, [_vp, _vp]),
    "mse_rabitq_export": (_i32, [_vp, _vp, _vp]),
    "mse_rabitq_encode": (_i32, [_vp, _vp, _u64, _vp, _vp, _vp]),
    "mse_rabitq_estimate": (_i32, [_vp, _vp, _vp, _vp, _vp, _u64, _vp]),
    "mse_rabitq_preprocess_query": (_i32, [_vp, _vp, _u32, _vp, _vp]),
    "mse_index_encode_rabitq": (_i32, [_vp, _vp, _i32]),
    "mse_rabitq_query_dev": (_i32, [_vp, _vp, _u32, _vp, _vp]),
    "mse_rabitq_destroy": (None, [_vp]),
    "mse_kmeans_assign": (_i32, [_vp, _vp, _u32, _u32, _i32, _vp, _vp]),
    "mse_kmeans_anneal": (_i32, [_vp, _u32, _u32, _u32, _u64, _vp, _vp, _vp]),
    "mse_shard_assign": (_i32, [_vp, _vp, _u32, _u32, C.c_double, _vp, _vp, _vp]),
    "mse_encoder_create": (_i32, [C.c_char_p, _i32, _i32, C.POINTER(_vp)]),
    "mse_encoder_config": (_i32, [_vp, _vp]),
    "mse_encode_images_u8": (_i32, [_vp, _vp, _i32, _vp]),
    "mse_encode_images_u8_dev": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "mse_encode_images_bmp": (_i32, [_vp, _vp, _vp, _i32, _vp]),
    "mse_resize_rgb_u8": (_i32, [_i32, _vp, _u32, _u32, _u32, _u32, _i32, _vp]),
    "mse_encode_images_resized": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp]),
    "mse_encode_text_ids": (_i32, [_vp, _vp, _i32, _vp]),
    "mse_encode_text_ids_dev": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "mse_encode_images_hidden": (_i32, [_vp, _vp, _i32, _i32, _vp]),
    "mse_encode_text_hidden": (_i32, [_vp, _vp, _i32, _i32, _vp]),
    "mse_encoder_profile": (_i32, [_vp, _i32]),
    "mse_encoder_stats": (_i32, [_vp, _vp]),
    "mse_encoder_destroy": (None, [_vp]),
    "mse_debug_gemm": (_i32, [_i32, _u32, _u32, _u32, _i32, _i32, _i32, _vp]),
    "mse_debug_attention": (_i32, [_i32, _i32, _i32, _i32, _i32, _vp]),
    "mse_gemm_f16_tn": (_i32, [_i32, _vp, _vp, _u32, _u32, _u32, _vp, _i32, _vp]),
}


def lib():
    """Load the CUDA library.  Fails loudly when it has not been built -- there is no other implementation."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise MseError(f"{path} is missing: run __graft_entry__.build() (nvcc, sm_100a). There is no CPU fallback.")
        l = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            f = getattr(l, name)
            f.restype, f.argtypes = res, args
        _LIB = l
    return _LIB


def last_error() -> str:
    return (lib().mse_last_error() or b"").decode("utf-8", "replace")


def check(rc: int, what: str = ""):
    if rc != 0:
        raise MseError(f"{what or 'mse call'} failed ({rc}): {last_error()}")


def launch_count() -> int:
    return int(lib().mse_launch_count())


def device_info(device: int = 0) -> dict:
    buf = C.create_string_buffer(1024)
    check(lib().mse_device_info(device, buf, 1024), "mse_device_info")
    return json.loads(buf.value.decode())
