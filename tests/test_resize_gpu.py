"""f3: resize_for_embed_sync (src/common.rs:31-54) on the device.  The reference calls the fast_image_resize crate (absent offline); the
kernel follows Pillow's ImagingResample -- the fixed-point separable convolution that crate descends from -- and Pillow is what can be
executed here: the CUDA result must equal PIL.Image.resize bit for bit (Hamming for shrinking both ways, Lanczos3 otherwise, the
reference's rule at :43-44).  Unpinned against the crate itself (its i16 coefficients may round a code value differently)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _image(seed, h, w):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 100 * np.sin(xx / 17.0 + seed), 128 + 100 * np.cos(yy / 23.0), (xx * 3 + yy * 5) % 256], axis=2)
    return np.clip(base + rng.normal(0, 20, (h, w, 3)), 0, 255).astype(np.uint8)


def test_resize_matches_pillow(mse):
    from PIL import Image
    from mse_b200.encoder import resize_for_embed
    cases = [(500, 700), (384, 600), (1000, 384), (200, 150), (384, 384), (97, 1031), (2048, 1536), (1, 1), (385, 383)]
    for i, (h, w) in enumerate(cases):
        img = _image(i, h, w)
        got = resize_for_embed(img, (384, 384))
        flt = Image.Resampling.HAMMING if (w > 384 and h > 384) else Image.Resampling.LANCZOS     # common.rs:43-44
        want = np.asarray(Image.fromarray(img, "RGB").resize((384, 384), flt))
        assert got.shape == want.shape
        assert np.array_equal(got, want), (h, w, int(np.abs(got.astype(int) - want.astype(int)).max()))
    # explicit filters and a non-square target
    img = _image(99, 300, 420)
    for f, flt in ((1, Image.Resampling.HAMMING), (2, Image.Resampling.LANCZOS)):
        got = resize_for_embed(img, (128, 200), filter=f)
        assert np.array_equal(got, np.asarray(Image.fromarray(img, "RGB").resize((128, 200), flt)))


def test_encode_image_resized_equals_host_resize(mse):
    """decoded images of any size -> device resize -> tower  ==  Pillow resize on the host -> mse_encode_images_u8."""
    import os, tempfile
    from PIL import Image
    from oracle import towers as T
    v, t = T.build_vision(depth=1, seed=42), T.build_text(depth=1, seed=43)
    sd = T.export_openclip(v, t)
    wpath = os.path.join(tempfile.gettempdir(), "mse_resize_towers1.msew")
    mse.weights.save_weights(wpath, sd, mse.weights.config_for(sd))
    enc = mse.Encoder(wpath, max_batch=4)
    os.remove(wpath)
    imgs = [_image(1, 500, 640), _image(2, 384, 384), _image(3, 120, 90)]
    host = np.stack([np.asarray(Image.fromarray(a, "RGB").resize((384, 384), Image.Resampling.HAMMING if (a.shape[1] > 384 and a.shape[0] > 384)
                                                                  else Image.Resampling.LANCZOS)) for a in imgs])
    assert np.array_equal(enc.encode_image_resized(imgs), enc.encode_image(host))
    enc.close()
