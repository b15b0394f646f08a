"""Host-side handle on the SigLIP towers (C ABI: mse_encoder_*).  Mirrors what clip_server.py does with the
OpenCLIP model object: ``encode_image`` / ``encode_text`` return L2-normalised fp16 features (clip_server.py:98-116)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import MseError, check, lib


class Encoder:
    def __init__(self, weights_path: str, device: int = 0, max_batch: int = 128):
        self._h = C.c_void_p()
        check(lib().mse_encoder_create(weights_path.encode(), device, max_batch, C.byref(self._h)), "mse_encoder_create")
        cfg = (C.c_int32 * 16)()
        check(lib().mse_encoder_config(self._h, cfg), "mse_encoder_config")
        (self.image_size, self.patch, self.dim, self.depth_v, self.heads, self.mlp, self.vocab, self.ctx, self.act, self.has_vision,
         self.has_text, self.depth_t) = [int(v) for v in cfg[:12]]
        self.max_batch = max_batch
        self.device = device
        self.tokens_per_image = (self.image_size // self.patch) ** 2

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            try:
                lib().mse_encoder_destroy(self._h)
            except Exception:  # interpreter shutdown
                pass
            self._h = C.c_void_p()

    __del__ = close

    def encode_image(self, images_u8: np.ndarray) -> np.ndarray:
        """images_u8: [B, H, W, 3] uint8 RGB (already at image_size, as the reference's clients send them). -> [B, dim] fp16."""
        x = np.ascontiguousarray(images_u8, np.uint8)
        if x.ndim != 4 or x.shape[1:] != (self.image_size, self.image_size, 3):
            raise MseError(f"encode_image expects [B,{self.image_size},{self.image_size},3] uint8, got {x.shape}")
        out = np.empty((x.shape[0], self.dim), np.float16)
        check(lib().mse_encode_images_u8(self._h, x.ctypes.data_as(C.c_void_p), x.shape[0], out.ctypes.data_as(C.c_void_p)), "mse_encode_images_u8")
        return out

    def encode_image_bmp(self, files) -> np.ndarray:
        """files: image_size x image_size 24-bit BMP files (bytes / uint8 arrays) as the reference's clients send them; unpacked on the
        device.  Raises MseError (MSE_ERR_UNSUPPORTED) for anything else."""
        bufs = [np.frombuffer(f, np.uint8) if isinstance(f, (bytes, bytearray, memoryview)) else np.ascontiguousarray(f, np.uint8) for f in files]
        n = len(bufs)
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        lens = (C.c_size_t * n)(*[b.size for b in bufs])
        out = np.empty((n, self.dim), np.float16)
        check(lib().mse_encode_images_bmp(self._h, ptrs, lens, n, out.ctypes.data_as(C.c_void_p)), "mse_encode_images_bmp")
        return out

    def encode_image_resized(self, images) -> np.ndarray:
        """images: decoded RGB8 arrays [h, w, 3] of ANY size (what the ingest client holds before resize_for_embed_sync,
        src/common.rs:31-54): resized on the device (Hamming when both dimensions shrink, else Lanczos3) into the tower's input."""
        arrs = [np.ascontiguousarray(a, np.uint8) for a in images]
        for a in arrs:
            if a.ndim != 3 or a.shape[2] != 3:
                raise MseError(f"encode_image_resized expects [h,w,3] uint8 arrays, got {a.shape}")
        n = len(arrs)
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        ws = (C.c_uint32 * n)(*[a.shape[1] for a in arrs])
        hs = (C.c_uint32 * n)(*[a.shape[0] for a in arrs])
        out = np.empty((n, self.dim), np.float16)
        check(lib().mse_encode_images_resized(self._h, ptrs, ws, hs, n, out.ctypes.data_as(C.c_void_p)), "mse_encode_images_resized")
        return out

    def encode_text(self, ids: np.ndarray) -> np.ndarray:
        """ids: [B, ctx] token ids (int). -> [B, dim] fp16."""
        t = np.ascontiguousarray(ids, np.int32)
        if t.ndim != 2 or t.shape[1] != self.ctx:
            raise MseError(f"encode_text expects [B,{self.ctx}] token ids, got {t.shape}")
        out = np.empty((t.shape[0], self.dim), np.float16)
        check(lib().mse_encode_text_ids(self._h, t.ctypes.data_as(C.c_void_p), t.shape[0], out.ctypes.data_as(C.c_void_p)), "mse_encode_text_ids")
        return out

    def encode_image_dev(self, img_ptr: int, batch: int, out_ptr: int, stream: int = 0):
        check(lib().mse_encode_images_u8_dev(self._h, C.c_void_p(img_ptr), batch, C.c_void_p(out_ptr), C.c_void_p(stream)), "mse_encode_images_u8_dev")

    def encode_text_dev(self, ids_ptr: int, batch: int, out_ptr: int, stream: int = 0):
        check(lib().mse_encode_text_ids_dev(self._h, C.c_void_p(ids_ptr), batch, C.c_void_p(out_ptr), C.c_void_p(stream)), "mse_encode_text_ids_dev")

    def profile(self, enable: bool = True):
        check(lib().mse_encoder_profile(self._h, int(enable)), "mse_encoder_profile")

    def stats(self) -> dict:
        out = (C.c_uint64 * 8)()
        check(lib().mse_encoder_stats(self._h, out), "mse_encoder_stats")
        return {"gemm_ns": int(out[0]), "gemm_launches": int(out[1]), "attn_ns": int(out[2]), "attn_launches": int(out[3]),
                "launches": int(out[4]), "gemm_mflop": int(out[5])}

    # per-layer parity hooks
    def image_hidden(self, images_u8: np.ndarray, n_blocks: int) -> np.ndarray:
        x = np.ascontiguousarray(images_u8, np.uint8)
        out = np.empty((x.shape[0], self.tokens_per_image, self.dim), np.float16)
        check(lib().mse_encode_images_hidden(self._h, x.ctypes.data_as(C.c_void_p), x.shape[0], n_blocks, out.ctypes.data_as(C.c_void_p)), "mse_encode_images_hidden")
        return out

    def text_hidden(self, ids: np.ndarray, n_blocks: int) -> np.ndarray:
        t = np.ascontiguousarray(ids, np.int32)
        out = np.empty((t.shape[0], self.ctx, self.dim), np.float16)
        check(lib().mse_encode_text_hidden(self._h, t.ctypes.data_as(C.c_void_p), t.shape[0], n_blocks, out.ctypes.data_as(C.c_void_p)), "mse_encode_text_hidden")
        return out


def resize_for_embed(image: np.ndarray, size: tuple[int, int], device: int = 0, filter: int = 0) -> np.ndarray:
    """resize_for_embed_sync (src/common.rs:31-54) on the device: RGB8 [h, w, 3] -> [size[1], size[0], 3] (size = (width, height) as
    InferenceServerConfig.image_size); filter 0 = the reference's rule, 1 = Hamming, 2 = Lanczos3."""
    a = np.ascontiguousarray(image, np.uint8)
    if a.ndim != 3 or a.shape[2] != 3:
        raise MseError(f"resize_for_embed expects an [h,w,3] uint8 array, got {a.shape}")
    out = np.empty((size[1], size[0], 3), np.uint8)
    check(lib().mse_resize_rgb_u8(device, a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[0], size[0], size[1], filter,
                                  out.ctypes.data_as(C.c_void_p)), "mse_resize_rgb_u8")
    return out

