"""Real `torch.onnx.export` graphs (what the reference's `ort::Session` loads, pull_onnx.py:169-181) on the CPU:

* the oracle ONNX interpreter (`oracle/onnx_interp.py`) reproduces the exporting `nn.Module` and agrees with the
  functional oracle (`oracle/reference_forward.py`) on the same seeded weights — the *file* is what is compared;
* the engine's graph recogniser (`csrc/onnx_graph.cc`, reached through the parse-only C ABI entry points) binds every
  parameter from graph structure alone: hyper-parameters (heads, eps, activation, pooling, causal mask) and tensors
  (canonical name, `[out, in]` layout) must equal what the exporter was given, also when every initializer has been
  renamed to `val_<n>`, and for a third-party model (HF transformers CLIP: split q/k/v projections, pooling before
  the final LayerNorm)."""
import os
import warnings

import numpy as np
import pytest
import torch

import export_synthetic as ex
import torch_export as te
from clip_embedder_rs_b200 import error
from clip_embedder_rs_b200.onnx import inspect_onnx, read_onnx_tensor
from oracle import onnx_interp as oi
from oracle import reference_forward as R

CASES = [("tiny_clip", False), ("tiny_siglip", True), ("tiny_clip_p14", True)]


def _weights(spec, seed=0):
    w = {}
    ex.gen_vision(spec, seed, lambda n, a: w.__setitem__(n, np.array(a)))
    ex.gen_text(spec, seed, lambda n, a: w.__setitem__(n, np.array(a)))
    return w


@pytest.mark.parametrize("config,anonymize", CASES)
def test_interpreter_matches_module_and_functional_oracle(make_real_model, make_model, config, anonymize):
    spec = ex.CONFIGS[config]
    real = make_real_model(config, anonymize=anonymize)
    synth = make_model(config)
    model = te.build_model(spec, 0)
    s = spec.vision.image_size
    x = np.random.default_rng(0).standard_normal((3, 3, s, s)).astype(np.float32)
    ids, _ = R.tokenize(synth, ["a photo of a cat", "a dog", "quite a long sentence about nothing in particular", ""])
    with torch.no_grad():
        want_v = te.VisualWrapper(model)(torch.from_numpy(x)).numpy()
        want_t = te.TextWrapper(model)(torch.from_numpy(ids)).numpy()
    sv, st = oi.OnnxSession(os.path.join(real, "visual.onnx")), oi.OnnxSession(os.path.join(real, "text.onnx"))
    assert sv.input_names == ["pixel_values"] and st.input_names == ["input_ids"]  # src/vision.rs:73-75, src/text.rs:87
    got_v, got_t = sv.run({"pixel_values": x}), st.run({"input_ids": ids})
    assert np.abs(got_v - want_v).max() < 2e-6 and np.abs(got_t - want_t).max() < 2e-6
    # batch is a dynamic axis (pull_onnx.py:172-177): a batch the exporter never saw
    assert sv.run({"pixel_values": x[:1]}).shape == (1, spec.embed_dim)
    assert np.abs(sv.run({"pixel_values": x[:1]}) - want_v[:1]).max() < 2e-6
    # the independent restatement of the architecture agrees with the executed file
    fv = R.vision_forward(R.Tower(os.path.join(synth, "visual.onnx")), x)
    ft = R.text_forward(R.Tower(os.path.join(synth, "text.onnx")), ids)
    assert np.abs(fv - got_v).max() < 1e-5 and np.abs(ft - got_t).max() < 1e-5
    # fp64 execution of the same file bounds the interpreter's own rounding
    assert np.abs(sv.run({"pixel_values": x}, dtype=torch.float64) - got_v).max() < 1e-5


@pytest.mark.parametrize("config,anonymize", CASES)
def test_recogniser_recovers_hyperparameters(make_real_model, config, anonymize):
    spec = ex.CONFIGS[config]
    real = make_real_model(config, anonymize=anonymize)
    jv, jt = inspect_onnx(os.path.join(real, "visual.onnx")), inspect_onnx(os.path.join(real, "text.onnx"))
    for j in (jv, jt):
        assert j["graph"]["attempted"] and j["graph"]["recognized"], j["graph"]["error"]
        assert j["num_nodes"] > 100 and j["metadata"]["clipb200.binding"] == "graph"
    v, t = spec.vision, spec.text
    mv, mt = jv["metadata"], jt["metadata"]
    assert mv["clipb200.tower"] == "vision" and mt["clipb200.tower"] == "text"
    assert int(mv["clipb200.heads"]) == v.heads and int(mt["clipb200.heads"]) == t.heads
    assert int(mv["clipb200.layers"]) == v.layers and int(mt["clipb200.layers"]) == t.layers
    assert int(mv["clipb200.mlp_dim"]) == v.mlp_dim and int(mt["clipb200.mlp_dim"]) == t.mlp_dim
    assert int(mv["clipb200.act"]) == ex.ACT_IDS[v.act] and int(mt["clipb200.act"]) == ex.ACT_IDS[t.act]
    assert abs(float(mv["clipb200.eps"]) - v.eps) < 1e-9 and abs(float(mt["clipb200.eps"]) - t.eps) < 1e-9
    assert mv["clipb200.pool"] == v.pool and mt["clipb200.pool"] == t.pool
    assert mv["clipb200.family"] == v.family
    assert int(mt["clipb200.causal"]) == int(t.causal)
    assert int(mv["clipb200.embed_dim"]) == spec.embed_dim == int(mt["clipb200.embed_dim"])
    assert int(mv["clipb200.patch"]) == v.patch and int(mt["clipb200.context_length"]) == t.context_length
    if anonymize:  # nothing but structure was available
        assert all(b["source"].startswith(("val_", "<", "concat(")) for b in jv["graph"]["bindings"])
    # Linear weights that the exporter pre-transposed into MatMul operands are flagged as such
    assert any(b["transposed"] for b in jv["graph"]["bindings"]) and any(b["transposed"] for b in jt["graph"]["bindings"])


@pytest.mark.parametrize("config,anonymize", CASES)
def test_recogniser_binds_every_tensor(make_real_model, config, anonymize):
    """Every parameter the exporter was given comes back under its open_clip / timm name, bit-identical, in
    `[out, in]` layout, although the file stores Linear weights transposed under `onnx::MatMul_<n>` / `val_<n>`."""
    spec = ex.CONFIGS[config]
    real = make_real_model(config, anonymize=anonymize)
    w = _weights(spec)
    text_prefix = "model." if spec.text.family == "clip" else "model.text."
    checked = 0
    for name, arr in w.items():
        if ".attn_pool.latent" in name or ".attn_pool.q." in name:
            continue  # folded into clipb200.map_query (below)
        is_text = name.startswith(text_prefix) and ".visual." not in name
        path = os.path.join(real, "text.onnx" if is_text else "visual.onnx")
        canonical = "model." + name[len(text_prefix):] if is_text else name  # text towers are reported CLIP-style
        got = read_onnx_tensor(path, canonical)
        assert got.size == arr.size, (name, got.shape, arr.shape)
        assert np.array_equal(got.reshape(arr.shape), arr), name
        checked += 1
    assert checked >= 40
    if spec.vision.family == "timm":
        p = "model.visual.trunk.attn_pool"
        hd = spec.vision.width // spec.vision.heads
        q = (w[f"{p}.latent"].reshape(-1).astype(np.float64) @ w[f"{p}.q.weight"].astype(np.float64).T
             + w[f"{p}.q.bias"]) * hd ** -0.5
        got = read_onnx_tensor(os.path.join(real, "visual.onnx"), "clipb200.map_query")
        assert np.abs(got - q).max() < 1e-6 * max(1.0, np.abs(q).max())


def _hf_clip(seed=0):
    from transformers import CLIPConfig, CLIPModel

    cfg = CLIPConfig(
        text_config=dict(hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                         vocab_size=1000, max_position_embeddings=77),
        vision_config=dict(hidden_size=192, intermediate_size=640, num_hidden_layers=3, num_attention_heads=3,
                           image_size=64, patch_size=16),
        projection_dim=96)
    cfg._attn_implementation = "eager"
    torch.manual_seed(seed)
    m = CLIPModel(cfg).eval()
    with torch.no_grad():  # default init has identical LayerNorms (de-duplicated by the exporter): make them distinct
        for n, p in m.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
            elif "norm" in n or n.endswith("bias"):
                p.copy_(0.02 * torch.randn_like(p))
    return m


class HFVision(torch.nn.Module):
    def __init__(self, m):
        super().__init__()
        self.m = m

    def forward(self, x):
        f = self.m.visual_projection(self.m.vision_model(pixel_values=x).pooler_output)
        return torch.nn.functional.normalize(f, dim=-1)


@pytest.fixture(scope="session")
def hf_clip_vision(model_root):
    m = _hf_clip()
    path = os.path.join(model_root, "hf_clip_visual.onnx")
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        te.export_tower(HFVision(m), torch.randn(2, 3, 64, 64), path, "pixel_values", "image_embeddings")
    return m, path


def test_third_party_export_hf_clip_vision(hf_clip_vision):
    """A graph this repo did not shape: HF transformers' CLIP vision tower (separate q/k/v Linear layers, `q * scale`,
    pooling before `post_layernorm`, parameters named `vision_model.encoder.layers...`)."""
    m, path = hf_clip_vision
    x = np.random.default_rng(3).standard_normal((2, 3, 64, 64)).astype(np.float32)
    with torch.no_grad():
        want = HFVision(m)(torch.from_numpy(x)).numpy()
    assert np.abs(oi.OnnxSession(path).run({"pixel_values": x}) - want).max() < 2e-6
    j = inspect_onnx(path)
    assert j["graph"]["recognized"], j["graph"]["error"]
    md = j["metadata"]
    assert (md["clipb200.family"], md["clipb200.pool"], int(md["clipb200.heads"]), int(md["clipb200.layers"]),
            int(md["clipb200.act"]), int(md["clipb200.embed_dim"])) == ("clip", "cls", 3, 3, 1, 96)
    sd = {k: v.numpy() for k, v in m.state_dict().items()}
    for i in range(3):
        hp, cp = f"vision_model.encoder.layers.{i}", f"model.visual.transformer.resblocks.{i}"
        qkv_w = np.concatenate([sd[f"{hp}.self_attn.{n}_proj.weight"] for n in "qkv"], 0)
        qkv_b = np.concatenate([sd[f"{hp}.self_attn.{n}_proj.bias"] for n in "qkv"], 0)
        assert np.array_equal(read_onnx_tensor(path, f"{cp}.attn.in_proj_weight"), qkv_w)
        assert np.array_equal(read_onnx_tensor(path, f"{cp}.attn.in_proj_bias"), qkv_b)
        assert np.array_equal(read_onnx_tensor(path, f"{cp}.attn.out_proj.weight"), sd[f"{hp}.self_attn.out_proj.weight"])
        assert np.array_equal(read_onnx_tensor(path, f"{cp}.mlp.c_fc.weight"), sd[f"{hp}.mlp.fc1.weight"])
        assert np.array_equal(read_onnx_tensor(path, f"{cp}.mlp.c_proj.bias"), sd[f"{hp}.mlp.fc2.bias"])
        assert np.array_equal(read_onnx_tensor(path, f"{cp}.ln_2.weight"), sd[f"{hp}.layer_norm2.weight"])
    assert np.array_equal(read_onnx_tensor(path, "model.visual.proj"), sd["visual_projection.weight"].T)
    assert np.array_equal(read_onnx_tensor(path, "model.visual.class_embedding"), sd["vision_model.embeddings.class_embedding"])
    assert np.array_equal(read_onnx_tensor(path, "model.visual.positional_embedding"),
                          sd["vision_model.embeddings.position_embedding.weight"])
    assert np.array_equal(read_onnx_tensor(path, "model.visual.ln_post.bias"), sd["vision_model.post_layernorm.bias"])


def test_unrecognised_graph_is_reported_not_guessed(tmp_path):
    """A graph that is not one of the supported tower layouts must be declined with a reason; the engine refuses such a
    file (tests/test_real_export_gpu.py::test_declined_graph_is_refused_not_guessed) instead of binding it by name."""

    class Odd(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(3, 8, 4, 4)
            self.fc = torch.nn.Linear(8, 4)

        def forward(self, x):
            return torch.nn.functional.normalize(self.fc(self.conv(x).flatten(2).mean(-1)), dim=-1)

    path = str(tmp_path / "visual.onnx")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        te.export_tower(Odd().eval(), torch.randn(2, 3, 16, 16), path, "pixel_values", "image_embeddings")
    j = inspect_onnx(path)
    assert j["graph"]["attempted"] and not j["graph"]["recognized"]
    assert "graph recogniser" in j["graph"]["error"]
    with pytest.raises(error.Ort):
        read_onnx_tensor(path, "model.visual.conv1.weight")
    assert read_onnx_tensor(path, "conv.weight").shape == (8, 3, 4, 4)  # exported names stay readable


def test_initializer_only_files_skip_the_recogniser(make_model):
    j = inspect_onnx(os.path.join(make_model("tiny_clip"), "visual.onnx"))
    assert not j["graph"]["attempted"] and j["num_nodes"] == 0


# ---------------------------------------------------------------------------------------------------------------
# An independent ONNX runtime.  onnxruntime (what the reference links) is not in this image, but OpenCV's DNN module is:
# a third-party ONNX importer + CPU executor that shares no code with torch, with this repo's interpreter or with the
# engine.  It cannot shape-infer a dynamic batch axis, so the towers are exported with the batch baked in (no
# dynamic_axes) and weights inline; the CLIP text tower's argmax pooling (Range / Flatten / Gather arithmetic) is beyond
# its Gather, so text is checked on the SigLIP tower (last-token pooling).
def _export_static(tmp_path, config, tower):
    spec = ex.CONFIGS[config]
    model = te.build_model(spec, 0)
    path = str(tmp_path / f"{config}_{tower}_static.onnx")
    if tower == "vision":
        s = spec.vision.image_size
        feed = np.random.default_rng(7).standard_normal((2, 3, s, s)).astype(np.float32)
        wrapper, in_name, out_name = te.VisualWrapper(model), "pixel_values", "image_embeddings"
    else:
        feed = np.random.default_rng(8).integers(1, spec.text.vocab_size - 2, (2, spec.text.context_length)).astype(np.int64)
        wrapper, in_name, out_name = te.TextWrapper(model), "input_ids", "text_embeddings"
    te.export_tower(wrapper, torch.from_numpy(feed), path, in_name, out_name, dynamic_batch=False, external_data=False)
    with torch.no_grad():
        want = wrapper(torch.from_numpy(feed)).numpy()
    return path, in_name, feed, want


@pytest.mark.parametrize("config,tower", [("tiny_clip", "vision"), ("tiny_siglip", "vision"), ("tiny_siglip", "text")])
def test_opencv_dnn_runs_the_same_file(tmp_path, make_model, config, tower):
    cv2 = pytest.importorskip("cv2")
    path, in_name, feed, want = _export_static(tmp_path, config, tower)
    net = cv2.dnn.readNetFromONNX(path)
    net.setInput(feed, in_name)
    got_cv = net.forward()
    got_interp = oi.OnnxSession(path).run({in_name: feed})
    synth = make_model(config)
    got_oracle = (R.vision_forward(R.Tower(os.path.join(synth, "visual.onnx")), feed) if tower == "vision"
                  else R.text_forward(R.Tower(os.path.join(synth, "text.onnx")), feed))
    assert got_cv.shape == want.shape
    # four implementations of the same file / weights: torch module, OpenCV DNN, the oracle interpreter, the functional oracle
    assert np.abs(got_cv - want).max() < 2e-6
    assert np.abs(got_cv - got_interp).max() < 2e-6
    assert np.abs(got_cv - got_oracle).max() < 1e-5
    # and the engine's recogniser binds the fixed-batch graph too (class token / pool query expanded along batch 2)
    j = inspect_onnx(path)
    assert j["graph"]["recognized"], j["graph"]["error"]
    spec = ex.CONFIGS[config]
    assert int(j["metadata"]["clipb200.heads"]) == (spec.vision.heads if tower == "vision" else spec.text.heads)


# ---------------------------------------------------------------------------------------------------------------------
# Re-parameterised FastViT (MobileCLIP2) as a REAL exported graph: Conv / BatchNormalization / Sigmoid / Relu / Erf nodes.
# Until round 2 the FastViT restatement of the oracle had no independent implementation beside it; now the same seeded
# weights go through (1) the timm-shaped nn.Module of tools/torch_export.py, (2) the file torch.onnx.export writes,
# executed by the oracle's ONNX interpreter and (3) by OpenCV's DNN module, and (4) the functional oracle.
FASTVIT_CASES = ["tiny_mobileclip", "tiny_mobileclip5"]   # 4 stages / attention in the last; 5 stages / attention in the last two


@pytest.mark.parametrize("config", FASTVIT_CASES)
def test_fastvit_real_graph_matches_module_interpreter_and_functional_oracle(make_real_model, make_model, config):
    spec = ex.CONFIGS[config]
    real = make_real_model(config, towers=("vision",))
    synth = make_model(config)
    model = te.build_model(spec, 0, towers=("vision",))
    s = spec.vision.image_size
    x = np.random.default_rng(0).random((3, 3, s, s)).astype(np.float32)
    with torch.no_grad():
        want = te.VisualWrapper(model)(torch.from_numpy(x)).numpy()
    sess = oi.OnnxSession(os.path.join(real, "visual.onnx"))
    assert sess.input_names == ["pixel_values"]
    got = sess.run({"pixel_values": x})
    assert np.abs(got - want).max() < 2e-6
    assert np.abs(sess.run({"pixel_values": x[:1]}) - want[:1]).max() < 2e-6     # dynamic batch axis
    fv = R.vision_forward(R.Tower(os.path.join(synth, "visual.onnx")), x)
    assert np.abs(fv - got).max() < 1e-5, "the functional FastViT restatement and the executed graph disagree"
    # embeddings of different images must differ (a collapsed random network would make every comparison vacuous)
    assert (got[0] @ got[1]) < 0.999


@pytest.mark.parametrize("config", FASTVIT_CASES)
def test_fastvit_real_graph_binds_by_name_plus_graph(make_real_model, config):
    """`torch.onnx.export` keeps every FastViT parameter under its module name except the attention blocks' two Linear
    weights (pre-transposed and renamed `onnx::MatMul_<n>`); `bind_fastvit_graph` (csrc/onnx_graph.cc) finds those from
    the named BatchNorm scale / proj bias next to them.  Every tensor the engine loads must come back exactly."""
    spec = ex.CONFIGS[config]
    path = os.path.join(make_real_model(config, towers=("vision",)), "visual.onnx")
    j = inspect_onnx(path)
    assert j["graph"]["attempted"] and not j["graph"]["recognized"] and j["graph"]["fastvit_by_name"], j["graph"]
    n_attn = sum(spec.vision.depths[len(spec.vision.dims) - spec.vision.attn_stages:])
    assert int(j["metadata"]["clipb200.fastvit_graph_linears"]) == 2 * n_attn
    w = {}
    ex.gen_vision(spec, 0, lambda n, a: w.__setitem__(n, np.array(a)))
    renamed = [k for k in w if k.endswith("token_mixer.qkv.weight") or k.endswith("token_mixer.proj.weight")]
    assert len(renamed) == 2 * n_attn
    for k, v in w.items():
        got = read_onnx_tensor(path, k)
        assert got.shape == v.shape and np.array_equal(got, v.astype(np.float32)), k


def test_opencv_dnn_runs_the_fastvit_graph(tmp_path, make_model):
    cv2 = pytest.importorskip("cv2")
    config = "tiny_mobileclip"
    path, in_name, feed, want = _export_static(tmp_path, config, "vision")
    net = cv2.dnn.readNetFromONNX(path)
    net.setInput(feed, in_name)
    got_cv = net.forward()
    got_interp = oi.OnnxSession(path).run({in_name: feed})
    got_oracle = R.vision_forward(R.Tower(os.path.join(make_model(config), "visual.onnx")), feed)
    assert got_cv.shape == want.shape
    assert np.abs(got_cv - want).max() < 5e-6
    assert np.abs(got_cv - got_interp).max() < 5e-6
    assert np.abs(got_cv - got_oracle).max() < 1e-5
    assert inspect_onnx(path)["graph"]["fastvit_by_name"]
