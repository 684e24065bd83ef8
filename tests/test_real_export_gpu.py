"""GPU parity on real `torch.onnx.export` graphs (the files the reference's `ort::Session` loads): the engine binds
the weights from graph structure (csrc/onnx_graph.cc) and must agree with the oracle ONNX interpreter executing the
same file in fp32 — per-embedding cosine >= 0.999 (BASELINE.json north_star), identical classify label order."""
import os
import warnings

import numpy as np
import pytest
import torch

from conftest import cosine_rows, random_images, random_texts

pytestmark = pytest.mark.gpu

COS_BAR = 0.999
CASES = [("tiny_clip", False), ("tiny_clip_p14", True), ("tiny_siglip", True), ("small_siglip", True)]


@pytest.mark.parametrize("config,anonymize", CASES)
def test_real_export_embeddings_match_interpreter(make_real_model, config, anonymize):
    import clip_embedder_rs_b200 as cb
    from oracle import onnx_interp as oi
    from oracle import reference_forward as R

    mdir = make_real_model(config, anonymize=anonymize)
    clip = cb.Clip.from_local_dir(mdir).build()
    size = clip.vision.config.model_cfg.vision_cfg.image_size
    pc = clip.vision.config.preprocess_cfg
    imgs = random_images(6, size, seed=31)
    texts = random_texts(9, seed=32) + ["a photo of a cat", ""]
    pv = R.preprocess_batch(list(imgs), size, pc.mean, pc.std)
    want_v = oi.OnnxSession(os.path.join(mdir, "visual.onnx")).run({"pixel_values": pv})
    ids, _ = R.tokenize(mdir, texts)
    got_ids, _ = clip.text.tokenize(texts)
    assert np.array_equal(got_ids, ids)
    want_t = oi.OnnxSession(os.path.join(mdir, "text.onnx")).run({"input_ids": ids})
    got_v, got_t = clip.vision.embed_images(imgs), clip.text.embed_texts(texts)
    cv, ct = cosine_rows(got_v, want_v), cosine_rows(got_t, want_t)
    print(f"\n[{config}{' anonymised' if anonymize else ''}] vision cos >= {cv.min():.6f} (max abs "
          f"{np.abs(got_v - want_v).max():.2e}), text cos >= {ct.min():.6f} (max abs {np.abs(got_t - want_t).max():.2e})")
    assert cv.min() >= COS_BAR and ct.min() >= COS_BAR
    # the same tail as Clip::classify (src/clip.rs:94-132) on the interpreter's embeddings gives the same ranking
    labels = texts[:5]
    got = clip.classify(imgs[0], labels)
    mc = clip.get_model_config()
    probs = R.probabilities(want_t[:5], want_v[0], {"logit_scale": mc.logit_scale, "logit_bias": mc.logit_bias,
                                                   "activation_function": mc.activation_function})
    want = R.sort_desc(list(zip(labels, probs.tolist())))
    assert [l for l, _ in got] == [l for l, _ in want]


def test_real_and_initializer_only_files_give_identical_embeddings(make_real_model, make_model):
    """Same seeded weights reach the engine through two routes (graph recogniser vs binding by parameter name): the
    embeddings must be bit-identical, i.e. the recogniser changes nothing but where tensors are found."""
    import clip_embedder_rs_b200 as cb

    for config in ("tiny_clip", "tiny_siglip"):
        a = cb.Clip.from_local_dir(make_real_model(config, anonymize=True)).build()
        b = cb.Clip.from_local_dir(make_model(config)).build()
        size = a.vision.config.model_cfg.vision_cfg.image_size
        imgs = random_images(5, size, seed=41)
        texts = random_texts(5, seed=42)
        assert np.array_equal(a.vision.embed_images(imgs), b.vision.embed_images(imgs)), config
        assert np.array_equal(a.text.embed_texts(texts), b.text.embed_texts(texts)), config


def test_third_party_export_hf_clip_vision_on_gpu(tmp_path):
    """HF transformers' CLIP vision tower exported with torch.onnx.export: split q/k/v projections, pooling before the
    final LayerNorm, foreign parameter names.  Engine vs the exporting module itself."""
    from clip_embedder_rs_b200.onnx import OnnxSession
    from test_onnx_graph_cpu import HFVision, _hf_clip
    import torch_export as te
    import ctypes as C
    from clip_embedder_rs_b200 import _native

    m = _hf_clip()
    path = str(tmp_path / "visual.onnx")
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        te.export_tower(HFVision(m), torch.randn(2, 3, 64, 64), path, "pixel_values", "image_embeddings")
    x = np.random.default_rng(5).standard_normal((9, 3, 64, 64)).astype(np.float32)
    with torch.no_grad():
        want = HFVision(m)(torch.from_numpy(x)).numpy()
    s = OnnxSession(path)
    assert s.find_input(["pixel_values", "input"]) == "pixel_values" and s.embed_dim == 96 and s.image_size == 64
    out = np.empty((9, 96), dtype=np.float32)
    s.check(_native.lib.clipb200_vision_embed_f32(s.handle, x.ctypes.data_as(C.c_void_p), 9, out.ctypes.data_as(C.c_void_p)))
    cos = cosine_rows(out, want)
    print(f"\n[hf clip vision] cos >= {cos.min():.6f} max abs {np.abs(out - want).max():.2e}")
    assert cos.min() >= COS_BAR


@pytest.mark.parametrize("config", ["tiny_clip", "tiny_siglip"])
def test_engine_matches_opencv_dnn_on_the_same_file(tmp_path, config):
    """The engine against an independent ONNX runtime (OpenCV DNN) executing the very same exported file (fixed batch,
    inline weights): the closest stand-in for the reference's onnxruntime that this image offers."""
    cv2 = pytest.importorskip("cv2")
    import ctypes as C

    from clip_embedder_rs_b200 import _native
    from clip_embedder_rs_b200.onnx import OnnxSession
    from test_onnx_graph_cpu import _export_static

    path, in_name, feed, _ = _export_static(tmp_path, config, "vision")
    net = cv2.dnn.readNetFromONNX(path)
    net.setInput(feed, in_name)
    want = net.forward()
    s = OnnxSession(path)
    out = np.empty_like(want)
    s.check(_native.lib.clipb200_vision_embed_f32(s.handle, feed.ctypes.data_as(C.c_void_p), feed.shape[0],
                                                  out.ctypes.data_as(C.c_void_p)))
    cos = cosine_rows(out, want)
    print(f"\n[{config} vs OpenCV DNN] cos >= {cos.min():.6f} max abs {np.abs(out - want).max():.2e}")
    assert cos.min() >= COS_BAR


def test_declined_graph_is_refused_not_guessed(make_real_model, tmp_path):
    """A file that carries an executable graph is bound from that graph or not at all.  Here the graph is a real CLIP
    vision export whose GELU was swapped for a Relu: every parameter keeps its open_clip name, so a name-bound loader
    would accept it and silently compute QuickGELU where onnxruntime computes what the graph says.  The engine must
    refuse it with CLIPB200_ERR_UNSUPPORTED and the recogniser's reason; a graph that declares an attention_mask input is
    refused the same way."""
    import shutil

    import clip_embedder_rs_b200 as cb
    import onnx_proto as op
    from clip_embedder_rs_b200 import _native
    from clip_embedder_rs_b200.onnx import OnnxSession, inspect_onnx

    src = make_real_model("tiny_clip", anonymize=False)
    dst = tmp_path / "model"
    shutil.copytree(src, dst)

    def rewrite(path, node_fn=None, extra_input=None):
        buf = memoryview(open(path, "rb").read())
        out = bytearray()
        for f, w, val in op._fields(buf):
            if f != 7:
                out += op.f_bytes(f, bytes(val)) if w == 2 else op.f_varint(f, val)
                continue
            g = bytearray()
            for gf, gw, gval in op._fields(val):
                if gf == 1 and node_fn is not None:
                    g += op.f_bytes(1, node_fn(gval))
                elif gw == 2:
                    g += op.f_bytes(gf, bytes(gval))
                else:
                    g += op.f_varint(gf, gval)
            if extra_input is not None:
                g += op.f_bytes(11, extra_input)
            out += op.f_bytes(7, bytes(g))
        open(path, "wb").write(out)

    def sigmoid_to_relu(node):  # QuickGELU is x * sigmoid(1.702 x): turn its Sigmoid into a Relu
        nb = bytearray()
        for nf, nw, nval in op._fields(node):
            if nf == 4 and bytes(nval) == b"Sigmoid":
                nb += op.f_str(4, "Relu")
            elif nw == 2:
                nb += op.f_bytes(nf, bytes(nval))
            else:
                nb += op.f_varint(nf, nval)
        return bytes(nb)

    vis = str(dst / "visual.onnx")
    rewrite(vis, node_fn=sigmoid_to_relu)
    j = inspect_onnx(vis)
    assert j["graph"]["attempted"] and not j["graph"]["recognized"]
    with pytest.raises(cb.ClipError) as ei:
        OnnxSession(vis)
    assert ei.value.code == _native.ERR_UNSUPPORTED and "graph recogniser" in str(ei.value)

    txt = str(dst / "text.onnx")
    rewrite(txt, extra_input=op.value_info("attention_mask", op.INT64, ("batch_size", 77)))
    with pytest.raises(cb.ClipError) as ei:
        OnnxSession(txt)
    assert ei.value.code == _native.ERR_UNSUPPORTED and "attention_mask" in str(ei.value)


@pytest.mark.parametrize("config", ["tiny_mobileclip", "tiny_mobileclip5"])
def test_fastvit_real_export(make_real_model, make_model, config):
    """A re-parameterised FastViT trunk exported as a real graph (Conv / BatchNormalization / ... nodes, the attention
    Linears renamed by the exporter, no `clipb200.*` metadata, image size only in the input's declared shape): the engine
    must agree with the ONNX interpreter executing the same file, and bit-for-bit with itself loaded from the
    initializer-only file of the same weights."""
    import clip_embedder_rs_b200 as cb
    from oracle import onnx_interp as oi
    from oracle import reference_forward as R

    mdir = make_real_model(config, towers=("vision",))
    vis = cb.VisionEmbedder.from_local_dir(mdir).build()
    size = vis.config.model_cfg.vision_cfg.image_size
    assert vis.session.image_size == size
    pc = vis.config.preprocess_cfg
    imgs = random_images(5, size, seed=41)
    pv = R.preprocess_batch(list(imgs), size, pc.mean, pc.std)
    want = oi.OnnxSession(os.path.join(mdir, "visual.onnx")).run({"pixel_values": pv})
    got = vis.embed_images(imgs)
    c = cosine_rows(got, want)
    print(f"\n[{config} real graph] vision cos >= {c.min():.6f} (max abs {np.abs(got - want).max():.2e})")
    assert c.min() >= COS_BAR
    same = cb.VisionEmbedder.from_local_dir(make_model(config)).build().embed_images(imgs)
    assert np.array_equal(got, same)
