"""GPU resize (clipb200_resize_rgb8 / clipb200_vision_embed_rgb8_var) against the CPU oracle: bit-exact pixels for
every interpolation / resize_mode the reference distinguishes (src/vision.rs:176-192), photo-sized inputs, up- and
down-scaling, and end-to-end embeddings of mixed-size batches."""
import json
import os

import numpy as np
import pytest

from conftest import cosine_rows
from test_resize_cpu import structured

pytestmark = pytest.mark.gpu


def _with_preproc(make_model, config, interpolation, resize_mode, tmp_path_factory):
    """A copy of the model directory whose preprocess_cfg uses the requested interpolation / resize mode."""
    import shutil

    src = make_model(config)
    dst = str(tmp_path_factory.mktemp(f"{config}_{interpolation}_{resize_mode}"))
    for f in os.listdir(src):
        if f.endswith(".onnx") or f.endswith(".data"):
            os.symlink(os.path.join(src, f), os.path.join(dst, f))
        else:
            shutil.copy(os.path.join(src, f), os.path.join(dst, f))
    cfg = json.load(open(os.path.join(dst, "open_clip_config.json")))
    cfg["preprocess_cfg"]["interpolation"] = interpolation
    cfg["preprocess_cfg"]["resize_mode"] = resize_mode
    json.dump(cfg, open(os.path.join(dst, "open_clip_config.json"), "w"))
    return dst


@pytest.mark.parametrize("interpolation,resize_mode", [("bicubic", "shortest"), ("bicubic", "squash"),
                                                       ("bilinear", "shortest"), ("nearest", "squash")])
def test_resize_bit_exact(make_model, tmp_path_factory, interpolation, resize_mode):
    import clip_embedder_rs_b200 as cb
    from oracle import resize as RZ

    mdir = _with_preproc(make_model, "tiny_clip", interpolation, resize_mode, tmp_path_factory)
    emb = cb.VisionEmbedder.from_local_dir(mdir).build()
    size = emb.config.model_cfg.vision_cfg.image_size  # 64
    for i, (h, w) in enumerate([(1944, 2592), (480, 640), (375, 500), (64, 64), (40, 97), (64, 200), (333, 64)]):
        img = structured(h, w, seed=i)
        want = RZ.resize_rgb8(img, size, interpolation, resize_mode)
        got = emb.resize(img)
        assert got.shape == (size, size, 3) and got.dtype == np.uint8
        assert np.array_equal(got, want), f"{h}x{w}: {np.abs(got.astype(int) - want.astype(int)).max()} LSB off"


def test_mixed_size_batch_embeddings(make_model):
    import clip_embedder_rs_b200 as cb
    from oracle import reference_forward as R

    mdir = make_model("tiny_siglip")  # bicubic + squash, 64 px
    clip = cb.Clip.from_local_dir(mdir).build()
    o = R.OracleClip(mdir)
    imgs = [structured(h, w, seed=10 + i) for i, (h, w) in enumerate([(480, 640), (64, 64), (900, 300), (50, 70), (64, 64)])]
    got = clip.vision.embed_images(imgs)
    want = o.embed_images(imgs)
    cos = cosine_rows(got, want)
    print(f"\n[resize] mixed-size batch cos min {cos.min():.6f} max_abs {np.abs(got - want).max():.2e}")
    assert cos.min() >= 0.999
    pv = clip.vision.preprocess_batch(imgs)
    pc = clip.vision.config.preprocess_cfg
    assert np.array_equal(pv, R.preprocess_batch(imgs, 64, pc.mean, pc.std, pc.interpolation, pc.resize_mode))
    labels = ["a photo of a cat", "a photo of a dog", "a photo of a beignet"]
    assert [l for l, _ in clip.classify(imgs[0], labels)] == [l for l, _ in o.classify(imgs[0], labels)]


@pytest.mark.parametrize("size,patch", [(224, 32), (256, 32), (384, 32)])
def test_reference_photos_bit_exact(model_root, size, patch):
    """GPU resize == oracle, byte for byte, on the reference's own photos (assets/img: seven JPEGs between 2592x1456 and
    5312x2988) for bicubic / shortest and bicubic / squash at 224, 256 and 384 px, singly (clipb200_resize_rgb8) and as
    one mixed-size batch through the pipelined embed path (clipb200_vision_embed_rgb8_var)."""
    import clip_embedder_rs_b200 as cb
    import export_synthetic as ex
    from conftest import reference_photos
    from oracle import reference_forward as R
    from oracle import resize as RZ

    photos = reference_photos()
    if not photos:
        pytest.skip("reference photos not available (no /root/reference/assets/img and no tests/_ref_assets)")
    spec = ex._clip(f"photo-{size}", size, patch, 128, 1, 2, 256, 64, 128, 1, 2, 256)
    for mode in ("shortest", "squash"):
        spec.resize_mode = mode
        mdir = ex.write_model_dir(spec, os.path.join(model_root, f"photo_{size}_{mode}"), seed=0, towers=("vision",))
        emb = cb.VisionEmbedder.from_local_dir(mdir).micro_batch(4).build()
        want_px = []
        for name, a in photos:
            want = RZ.resize_rgb8(a, size, "bicubic", mode)
            got = emb.resize(a)
            assert np.array_equal(got, want), f"{name} -> {size} {mode}: {np.abs(got.astype(int) - want.astype(int)).max()} LSB off"
            want_px.append(want)
        # the batch path (groups of photos staged, uploaded and resized while the previous micro-batch runs) must see
        # exactly the pixels of the single-image path: embeddings of the photos == embeddings of the resized photos
        imgs = [a for _, a in photos]
        got_e = emb.embed_images(imgs)
        want_e = emb.embed_images(np.stack(want_px))
        assert np.array_equal(got_e, want_e)
        pc = emb.config.preprocess_cfg
        assert np.array_equal(emb.preprocess_batch(imgs[:2]), R.preprocess_batch(imgs[:2], size, pc.mean, pc.std, "bicubic", mode))
