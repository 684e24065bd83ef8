"""In-process multi-GPU pool (`clipb200_pool_*`, `.devices([...])` on the builders): the scaled-up `duplicate()` of the
reference (src/vision.rs:87-91, src/text.rs:104-108).  The pool must return exactly what one engine returns — same
kernels, independent rows — whatever the split.  On a one-GPU box the replicas share the device (`devices=[0, 0, 0]`),
which exercises the same host threads, row ranges and output placement; with two or more GPUs the same checks run
across devices."""
import numpy as np
import pytest

from conftest import random_images, random_texts
from test_resize_cpu import structured

pytestmark = pytest.mark.gpu


def _device_lists():
    from clip_embedder_rs_b200 import _native

    n = _native.lib.clipb200_device_count()
    lists = [[0], [0, 0, 0]]
    if n >= 2:
        lists.append(list(range(min(n, 8))))
    return lists


def test_pool_matches_single_engine_vision(make_model):
    import clip_embedder_rs_b200 as cb

    mdir = make_model("tiny_siglip")
    single = cb.VisionEmbedder.from_local_dir(mdir).micro_batch(4).build()
    imgs = random_images(23, 64, seed=11)
    want = single.embed_images(imgs)
    photos = [structured(h, w, seed=i) for i, (h, w) in enumerate([(480, 640), (64, 64), (300, 200), (90, 70), (64, 64), (700, 333), (128, 128)])]
    want_var = single.embed_images(photos)
    for devs in _device_lists():
        pool = cb.VisionEmbedder.from_local_dir(mdir).micro_batch(4).devices(devs).build()
        assert pool.session.is_pool and pool.session.devices == devs
        assert pool.session.embed_dim == single.session.embed_dim and pool.session.image_size == 64
        assert pool.input_name == single.input_name
        assert np.array_equal(pool.embed_images(imgs), want), devs
        assert np.array_equal(pool.embed_images(imgs[:2]), want[:2]), "fewer rows than replicas"
        assert np.array_equal(pool.embed_image(imgs[5]), want[5])
        assert np.array_equal(pool.embed_images(photos), want_var), "arbitrary-size path"
        with pytest.raises(cb.ClipError):
            pool.embed_images([])
        pool.session.close()


def test_pool_matches_single_engine_text_and_classify(make_model):
    import clip_embedder_rs_b200 as cb

    mdir = make_model("tiny_clip")
    single = cb.Clip.from_local_dir(mdir).build()
    texts = random_texts(17, seed=3)
    want = single.text.embed_texts(texts)
    img = random_images(1, 64, seed=5)[0]
    for devs in _device_lists():
        pool = cb.Clip.from_local_dir(mdir).devices(devs).build()
        assert np.array_equal(pool.text.embed_texts(texts), want), devs
        assert pool.classify(img, texts[:3]) == single.classify(img, texts[:3])
        # an out-of-vocabulary id in the LAST shard must surface as the call's error, naming the replica
        ids, _ = pool.text.tokenize(texts)
        ids[-1, 1] = 10 ** 6
        with pytest.raises(cb.ClipError) as ei:
            pool.text.embed_ids(ids)
        assert "replica" in str(ei.value) and "vocab" in str(ei.value)


def test_pool_rejects_bad_device(make_model):
    import clip_embedder_rs_b200 as cb

    with pytest.raises(cb.ClipError) as ei:
        cb.VisionEmbedder.from_local_dir(make_model("tiny_clip")).devices([0, 99]).build()
    assert "replica 1" in str(ei.value)
