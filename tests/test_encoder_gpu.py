"""GPU parity of the SigLIP towers through the C ABI against the fp32 oracle (oracle/towers.py).
Floating point: north_star's tolerance is cosine >= 1 - 1e-3 of the fp16 output against the fp32 reference,
checked per block (token activations) and end to end."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _cos(a, b):
    a = a.reshape(-1, a.shape[-1]).astype(np.float64)
    b = b.reshape(-1, b.shape[-1]).astype(np.float64)
    return (a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1))


@pytest.fixture(scope="module")
def setup(tmp_path_factory, mse):
    from oracle import towers as T
    v, t = T.build_vision(depth=2, seed=42), T.build_text(depth=2, seed=43)
    sd = T.export_openclip(v, t)
    path = str(tmp_path_factory.mktemp("w") / "towers2.msew")
    mse.weights.save_weights(path, sd, mse.weights.config_for(sd))
    enc = mse.Encoder(path, max_batch=5)
    return T, v, t, enc


def test_config(setup):
    _, _, _, enc = setup
    assert (enc.image_size, enc.dim, enc.depth_v, enc.depth_t, enc.ctx, enc.tokens_per_image) == (384, 1152, 2, 2, 64, 729)


def test_image_tower_per_block_and_final(setup):
    T, v, _, enc = setup
    imgs = T.synthetic_images(1, 2)
    ref, hs = T.encode_image(v, imgs, hidden_states=True)
    for n in range(3):  # after patch-embed + pos, after block 1, after block 2
        got = enc.image_hidden(imgs, n)
        c = _cos(got, hs[n])
        assert c.min() >= 1 - TOL, (n, float(c.min()))
        assert np.abs(got.astype(np.float32) - hs[n]).max() < 0.05 * np.abs(hs[n]).max()
    out = enc.encode_image(imgs)
    assert out.dtype == np.float16 and out.shape == (2, 1152)
    assert _cos(out, ref).min() >= 1 - TOL
    assert np.allclose(np.linalg.norm(out.astype(np.float32), axis=1), 1.0, atol=2e-3)
    # golden fixture generated in the build container (tests/golden/make_tower_golden.py)
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "towers_depth2.npz"))
    assert _cos(out, g["image_features"]).min() >= 1 - TOL


def test_text_tower_per_block_and_final(setup):
    T, _, t, enc = setup
    ids = T.synthetic_token_ids(1, 3)
    ref, hs = T.encode_text(t, ids, hidden_states=True)
    for n in range(3):
        got = enc.text_hidden(ids, n)
        c = _cos(got, hs[n])
        assert c.min() >= 1 - TOL, (n, float(c.min()))
    out = enc.encode_text(ids)
    assert _cos(out, ref).min() >= 1 - TOL
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "towers_depth2.npz"))
    assert _cos(out, g["text_features"]).min() >= 1 - TOL


def test_batch_sizes_and_determinism(setup):
    T, v, _, enc = setup
    imgs = T.synthetic_images(7, 5)
    full = enc.encode_image(imgs)
    one = np.concatenate([enc.encode_image(imgs[i:i + 1]) for i in range(5)])
    assert _cos(full, one).min() >= 1 - 1e-5  # a row's embedding does not depend on its batch neighbours
    assert np.array_equal(full, enc.encode_image(imgs))
    ref = T.encode_image(v, imgs)
    assert _cos(full, ref).min() >= 1 - TOL


def test_batch_limit_error(setup, mse):
    T, _, _, enc = setup
    with pytest.raises(mse.MseError) as e:
        enc.encode_image(T.synthetic_images(2, 6))
    assert "max batch size is 5" in str(e.value)  # clip_server.py:136,139
    with pytest.raises(mse.MseError):
        enc.encode_text(np.ones((6, 64), np.int32))


def test_full_depth_image_tower(tmp_path_factory, mse):
    """All 27 blocks at SO400M size, batch 2."""
    from oracle import towers as T
    v = T.build_vision(depth=27, seed=42)
    sd = T.export_openclip(vision=v)
    path = str(tmp_path_factory.mktemp("w27") / "vision27.msew")
    mse.weights.save_weights(path, sd, mse.weights.config_for(sd))
    enc = mse.Encoder(path, max_batch=2)
    imgs = T.synthetic_images(3, 2)
    ref = T.encode_image(v, imgs)
    out = enc.encode_image(imgs)
    assert _cos(out, ref).min() >= 1 - TOL, float(_cos(out, ref).min())


def test_full_depth_text_tower(tmp_path_factory, mse):
    """All 27 blocks of the text tower at SO400M size, batch 1 and 3 (64 and 192 rows: the skinny GEMM path with K = 1152 and
    K = 4304, and the tcgen05 path) against the fp32 oracle."""
    from oracle import towers as T
    t = T.build_text(depth=27, seed=43)
    sd = T.export_openclip(text=t)
    path = str(tmp_path_factory.mktemp("t27") / "text27.msew")
    mse.weights.save_weights(path, sd, mse.weights.config_for(sd))
    enc = mse.Encoder(path, max_batch=8)
    ids = T.synthetic_token_ids(9, 8)
    ref = T.encode_text(t, ids)
    out3 = enc.encode_text(ids[:3])
    out1 = enc.encode_text(ids[:1])
    assert _cos(out3, ref[:3]).min() >= 1 - TOL, float(_cos(out3, ref[:3]).min())
    assert _cos(out1, ref[:1]).min() >= 1 - TOL, float(_cos(out1, ref[:1]).min())
    # the first call at a batch size runs eagerly and records a CUDA graph; later calls replay it -- same bits, other inputs too
    assert np.array_equal(out1, enc.encode_text(ids[:1])) and np.array_equal(out1, enc.encode_text(ids[:1]))
    assert np.array_equal(out3, enc.encode_text(ids[:3]))
    other = enc.encode_text(ids[3:4])
    assert _cos(other, ref[3:4]).min() >= 1 - TOL and not np.array_equal(other, out1)
    # batch 8 = 512 rows: the 128 x 128 single-CTA tiles (between the skinny path and the 256 x 256 pair tiles)
    out8 = enc.encode_text(ids)
    assert _cos(out8, ref).min() >= 1 - TOL, float(_cos(out8, ref).min())
    assert np.array_equal(out8, enc.encode_text(ids))


def test_bmp_files_unpacked_on_device(setup, mse):
    """mse_encode_images_bmp: the 24-bit BMP files the reference's clients send (src/common.rs:42-53) give the same features as the
    decoded pixels through mse_encode_images_u8 -- bit for bit; bottom-up and top-down row order; other formats are refused."""
    import io
    from PIL import Image
    T, _, _, enc = setup
    imgs = T.synthetic_images(31, 3)
    want = enc.encode_image(imgs)

    def bmp(a):
        buf = io.BytesIO()
        Image.fromarray(a).save(buf, format="BMP")
        return buf.getvalue()
    files = [bmp(a) for a in imgs]
    assert np.array_equal(enc.encode_image_bmp(files), want)
    # top-down variant: negative height, rows in display order
    f = bytearray(files[0])
    stride, off = 384 * 3, int.from_bytes(f[10:14], "little")
    rows = [bytes(f[off + r * stride: off + (r + 1) * stride]) for r in range(384)]
    f[22:26] = (-384).to_bytes(4, "little", signed=True)
    f[off:] = b"".join(reversed(rows))
    assert np.array_equal(enc.encode_image_bmp([bytes(f)]), want[:1])
    small = io.BytesIO()
    Image.fromarray(imgs[0][:100, :100]).save(small, format="BMP")
    png = io.BytesIO()
    Image.fromarray(imgs[0]).save(png, format="PNG")
    for bad in (small.getvalue(), png.getvalue(), files[0][:1000]):
        with pytest.raises(mse.MseError):
            enc.encode_image_bmp([bad])
