"""CPU checks of the resize oracle (oracle/resize.py, restating reference src/vision.rs:164-198 + fast_image_resize's
published algorithm): identity at the model resolution, crop box arithmetic, agreement with Pillow (an independent
implementation of the same windows / kernels with 22-bit coefficients) to within 1 LSB."""
import numpy as np
import pytest

from oracle import resize as RZ


def structured(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 120 * np.sin(xx / 37.0 + seed), 127 + 120 * np.cos(yy / 23.0), (xx * 3 + yy * 5) % 256], -1)
    img = img + rng.normal(0, 25, size=img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def test_identity_at_resolution():
    x = np.random.default_rng(0).integers(0, 256, (96, 96, 3), dtype=np.uint8)
    for interp in ("bicubic", "bilinear", "nearest"):
        for mode in ("shortest", "squash"):
            assert np.array_equal(RZ.resize_rgb8(x, 96, interp, mode), x)


def test_crop_box():
    assert RZ.crop_box(640, 480, 224, "squash") == (0.0, 0.0, 640.0, 480.0)
    left, top, cw, ch = RZ.crop_box(640, 480, 224, "shortest")
    assert abs(cw - 480) < 1e-9 and abs(ch - 480) < 1e-9 and abs(left - 80) < 1e-9 and abs(top) < 1e-9
    left, top, cw, ch = RZ.crop_box(1944, 2592, 384, "anything-else")
    assert abs(cw - 1944) < 1e-9 and abs(top - 324) < 1e-9


@pytest.mark.parametrize("shape,size,interp,mode", [
    ((480, 640), 224, "bicubic", "shortest"), ((1000, 750), 384, "bicubic", "squash"),
    ((300, 500), 256, "bilinear", "shortest"), ((97, 211), 128, "bicubic", "shortest"),
    ((60, 40), 96, "bilinear", "squash")])
def test_against_pillow(shape, size, interp, mode):
    from PIL import Image

    a = structured(shape[0], shape[1], seed=size)
    mine = RZ.resize_rgb8(a, size, interp, mode)
    left, top, cw, ch = RZ.crop_box(shape[1], shape[0], size, mode)
    box = (max(left, 0.0), max(top, 0.0), min(left + cw, shape[1]), min(top + ch, shape[0]))
    pil = np.asarray(Image.fromarray(a).resize((size, size), Image.BICUBIC if interp == "bicubic" else Image.BILINEAR, box=box))
    d = np.abs(mine.astype(int) - pil.astype(int))
    assert d.max() <= 1, f"max diff {d.max()}"
    assert (d > 0).mean() < 0.05


def test_weights_sum_and_precision():
    xs, xn, xw = RZ.precompute_coefficients(2592, 324.0, 2268.0, 384, RZ.catmull_rom, 2.0)
    assert np.allclose(xw.sum(1), 1.0) and xw.shape[1] == int(np.ceil(2.0 * (1944 / 384))) * 2 + 1
    w16, p = RZ.normalise_i16(xw)
    assert 8 <= p <= 15 and np.abs(w16.sum(1) - (1 << p)).max() <= xw.shape[1]
    assert int(RZ._round_half_away(2.5)) == 3 and int(RZ._round_half_away(-2.5)) == -3


@pytest.mark.parametrize("size", [224, 256, 384])
@pytest.mark.parametrize("mode", ["shortest", "squash"])
def test_reference_photos_against_pillow(size, mode):
    """The photos the reference ships (assets/img, 1944x2592 ... 5312x2988 JPEGs; tests/integration_test.rs and the README
    examples embed them): the oracle's bicubic resize is within 1 LSB of Pillow's on every one of them, for the centre
    crop of `resize_mode: shortest` and for `squash`, at the three resolutions the reference's models use."""
    from PIL import Image

    from conftest import reference_photos

    photos = reference_photos()
    if not photos:
        pytest.skip("reference photos not available (no /root/reference/assets/img and no tests/_ref_assets)")
    for name, a in photos:
        h, w = a.shape[:2]
        mine = RZ.resize_rgb8(a, size, "bicubic", mode)
        left, top, cw, ch = RZ.crop_box(w, h, size, mode)
        box = (max(left, 0.0), max(top, 0.0), min(left + cw, w), min(top + ch, h))  # f64 rounding can leave -1e-13
        pil = np.asarray(Image.fromarray(a).resize((size, size), Image.BICUBIC, box=box))
        d = np.abs(mine.astype(int) - pil.astype(int))
        assert d.max() <= 1, f"{name} {w}x{h} -> {size} {mode}: max diff {d.max()}"
        assert (d > 0).mean() < 0.05, f"{name}: {(d > 0).mean():.3f} of the bytes differ"
