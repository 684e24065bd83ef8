"""CPU-only checks of the oracle and the host-side pieces (no GPU needed):
  * golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py) pin the exporter, the oracle numerics and
    the tokenizer;
  * an INDEPENDENT implementation of the same architectures (HF transformers' CLIPModel / SiglipModel, present in
    the image) is loaded with the same weights and must agree with the oracle: this is the strongest pin available,
    because the reference's own ort CPU path cannot run here and its tests hold no vectors;
  * reference-documented behaviours: Fixed(ctx) right padding with pad_id, truncation keeping specials, lowercase,
    empty batch error, softmax / sigmoid / stable descending sort of clip.rs.
"""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import cosine_rows, random_images, random_texts

from oracle import reference_forward as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONFIGS = ["tiny_clip", "tiny_clip_p14", "tiny_siglip"]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


@pytest.mark.parametrize("config", CONFIGS)
def test_golden(make_model, config):
    g = np.load(os.path.join(GOLDEN, f"{config}.npz"))
    mdir = make_model(config)
    assert _sha(os.path.join(mdir, "visual.onnx.data")) == str(g["visual_data_sha256"]), "exporter drifted"
    assert _sha(os.path.join(mdir, "text.onnx.data")) == str(g["text_data_sha256"]), "exporter drifted"
    o = R.OracleClip(mdir, threads=2)
    size = int(o.config["model_cfg"]["vision_cfg"]["image_size"])
    imgs = random_images(3, size, seed=int(g["image_seed"]))
    texts = [str(t) for t in g["texts"]]
    ids, mask = R.tokenize(mdir, texts)
    assert np.array_equal(ids, g["ids"]) and np.array_equal(mask, g["mask"])
    pc = o.config["preprocess_cfg"]
    pv = R.preprocess_batch(list(imgs), size, pc["mean"], pc["std"])
    assert np.array_equal(pv[0, :, 0, :8], g["pixel_first"])
    assert np.allclose([pv.astype(np.float64).sum(), np.abs(pv).astype(np.float64).sum()], g["pixel_checksum"], rtol=1e-12)
    assert np.allclose(o.embed_images(list(imgs)), g["image_embeddings"], atol=2e-5)
    assert np.allclose(o.embed_texts(texts), g["text_embeddings"], atol=2e-5)
    got = o.classify(imgs[0], texts[:3])
    assert [l for l, _ in got] == [str(l) for l in g["classify_order"]]
    assert np.allclose([p for _, p in got], g["classify_probs"], atol=1e-4)


def test_normalize_pixels_expression():
    """vision.rs:253-254, all 256 byte values, OpenAI and SigLIP statistics."""
    img = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, axis=2)
    for mean, std in (([0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]),
                      ([0.5, 0.5, 0.5], [0.5, 0.5, 0.5]), ([0.0, 0.0, 0.0], [1.0, 1.0, 1.0])):
        out = R.normalize_pixels(img, mean, std)
        for c in range(3):
            want = np.asarray([(np.float32(v) / np.float32(255.0) - np.float32(mean[c])) / np.float32(std[c])
                               for v in range(256)], dtype=np.float32)
            assert np.array_equal(out[c].reshape(-1), want)
    with pytest.raises(ValueError, match="Empty batch"):
        R.preprocess_batch([], 16, [0, 0, 0], [1, 1, 1])


def test_tokenizer_padding_truncation(make_model):
    for config, pad_id in (("tiny_clip", 0), ("tiny_siglip", 0)):
        mdir = make_model(config)
        ctx = json.load(open(os.path.join(mdir, "open_clip_config.json")))["model_cfg"]["text_cfg"]["context_length"]
        ids, mask = R.tokenize(mdir, ["a photo of a cat", "", "word " * 200])
        assert ids.shape == (3, ctx) and ids.dtype == np.int64 and mask.shape == (3, ctx)
        n0 = int(mask[0].sum())
        assert np.all(ids[0, n0:] == pad_id) and np.all(mask[0, n0:] == 0), "Fixed(ctx) right padding with pad_id"
        assert mask[2].sum() == ctx, "truncated to ctx"
        if config == "tiny_clip":
            assert ids[0, 0] == 49406 and ids[0, n0 - 1] == 49407 and ids[2, 0] == 49406 and ids[2, -1] == 49407, \
                "special tokens survive truncation"
            assert ids[0].argmax() == n0 - 1, "EOT is the arg-max id (EOT-argmax pooling)"
        else:
            up, _ = R.tokenize(mdir, ["A Photo Of A CAT"])
            assert np.array_equal(up[0], ids[0]), "tokenizer_needs_lowercase (text.rs:115-117)"


def test_tail_functions():
    logits = np.asarray([1.0, 2.0, 3.0, -50.0], dtype=np.float32)
    p = R.softmax(logits)
    assert abs(float(p.sum()) - 1.0) < 1e-6 and np.all(np.diff(p[:3]) > 0)
    assert np.allclose(p, torch.softmax(torch.tensor(logits), 0).numpy(), atol=1e-7)
    assert abs(float(R.sigmoid(0.0)) - 0.5) < 1e-7 and float(R.sigmoid(-200.0)) == 0.0
    items = [("a", 0.2), ("b", float("nan")), ("c", 0.2), ("d", 0.9)]
    assert [k for k, _ in R.sort_desc([("a", 0.2), ("c", 0.2), ("d", 0.9)])] == ["d", "a", "c"], "stable for ties"
    assert len(R.sort_desc(items)) == 4  # NaN compares Equal (clip.rs:129) and must not raise
    mc = {"logit_scale": 100.0, "logit_bias": None, "activation_function": None}
    e = np.eye(4, dtype=np.float32)[:3]
    assert np.allclose(R.probabilities(e, e[1], mc), R.softmax(np.asarray([0, 100, 0], dtype=np.float32)))
    mc = {"logit_scale": 10.0, "logit_bias": -5.0, "activation_function": "sigmoid"}
    assert np.allclose(R.probabilities(e, e[1], mc), R.sigmoid(np.asarray([-5, 5, -5], dtype=np.float32)))


# ----------------------------------------------------------------------------------------------------------
# independent implementation: HF transformers
# ----------------------------------------------------------------------------------------------------------
def _copy(dst: torch.nn.Parameter, src: torch.Tensor):
    assert tuple(dst.shape) == tuple(src.shape), (dst.shape, src.shape)
    with torch.no_grad():
        dst.copy_(src)


def _copy_clip_layers(layers, w, prefix, D):
    for i, layer in enumerate(layers):
        p = f"{prefix}.resblocks.{i}"
        wi, bi = w[f"{p}.attn.in_proj_weight"], w[f"{p}.attn.in_proj_bias"]
        for j, proj in enumerate((layer.self_attn.q_proj, layer.self_attn.k_proj, layer.self_attn.v_proj)):
            _copy(proj.weight, wi[j * D:(j + 1) * D]); _copy(proj.bias, bi[j * D:(j + 1) * D])
        _copy(layer.self_attn.out_proj.weight, w[f"{p}.attn.out_proj.weight"])
        _copy(layer.self_attn.out_proj.bias, w[f"{p}.attn.out_proj.bias"])
        _copy(layer.layer_norm1.weight, w[f"{p}.ln_1.weight"]); _copy(layer.layer_norm1.bias, w[f"{p}.ln_1.bias"])
        _copy(layer.layer_norm2.weight, w[f"{p}.ln_2.weight"]); _copy(layer.layer_norm2.bias, w[f"{p}.ln_2.bias"])
        _copy(layer.mlp.fc1.weight, w[f"{p}.mlp.c_fc.weight"]); _copy(layer.mlp.fc1.bias, w[f"{p}.mlp.c_fc.bias"])
        _copy(layer.mlp.fc2.weight, w[f"{p}.mlp.c_proj.weight"]); _copy(layer.mlp.fc2.bias, w[f"{p}.mlp.c_proj.bias"])


@pytest.mark.parametrize("config", ["tiny_clip", "tiny_clip_p14"])
def test_oracle_vs_transformers_clip(make_model, config):
    from transformers import CLIPConfig, CLIPModel

    import export_synthetic as ex

    spec = ex.CONFIGS[config]
    mdir = make_model(config)
    o = R.OracleClip(mdir, threads=2)
    v, t = spec.vision, spec.text
    cfg = CLIPConfig(
        vision_config=dict(hidden_size=v.width, intermediate_size=v.mlp_dim, num_hidden_layers=v.layers,
                           num_attention_heads=v.heads, image_size=v.image_size, patch_size=v.patch,
                           hidden_act="quick_gelu", layer_norm_eps=v.eps, projection_dim=spec.embed_dim),
        text_config=dict(hidden_size=t.width, intermediate_size=t.mlp_dim, num_hidden_layers=t.layers,
                         num_attention_heads=t.heads, max_position_embeddings=t.context_length,
                         vocab_size=t.vocab_size, hidden_act="quick_gelu", layer_norm_eps=t.eps,
                         projection_dim=spec.embed_dim, eos_token_id=2, bos_token_id=0, pad_token_id=1),
        projection_dim=spec.embed_dim)
    m = CLIPModel(cfg).eval()
    w = o.vision.w
    vm = m.vision_model
    _copy(vm.embeddings.class_embedding, w["model.visual.class_embedding"])
    _copy(vm.embeddings.patch_embedding.weight, w["model.visual.conv1.weight"])
    _copy(vm.embeddings.position_embedding.weight, w["model.visual.positional_embedding"])
    _copy(vm.pre_layrnorm.weight, w["model.visual.ln_pre.weight"]); _copy(vm.pre_layrnorm.bias, w["model.visual.ln_pre.bias"])
    _copy_clip_layers(vm.encoder.layers, w, "model.visual.transformer", v.width)
    _copy(vm.post_layernorm.weight, w["model.visual.ln_post.weight"]); _copy(vm.post_layernorm.bias, w["model.visual.ln_post.bias"])
    _copy(m.visual_projection.weight, w["model.visual.proj"].t())
    wt = o.text.w
    tm = m.text_model
    _copy(tm.embeddings.token_embedding.weight, wt["model.token_embedding.weight"])
    _copy(tm.embeddings.position_embedding.weight, wt["model.positional_embedding"])
    _copy_clip_layers(tm.encoder.layers, wt, "model.transformer", t.width)
    _copy(tm.final_layer_norm.weight, wt["model.ln_final.weight"]); _copy(tm.final_layer_norm.bias, wt["model.ln_final.bias"])
    _copy(m.text_projection.weight, wt["model.text_projection"].t())

    imgs = random_images(3, v.image_size, seed=5)
    pc = o.config["preprocess_cfg"]
    pv = R.preprocess_batch(list(imgs), v.image_size, pc["mean"], pc["std"])
    with torch.no_grad():
        hf_v = m.get_image_features(pixel_values=torch.from_numpy(pv))
        hf_v = getattr(hf_v, "pooler_output", hf_v)
        hf_v = torch.nn.functional.normalize(hf_v, dim=-1).numpy()
    mine_v = R.vision_forward(o.vision, pv)
    assert cosine_rows(mine_v, hf_v).min() > 0.99999 and np.abs(mine_v - hf_v).max() < 5e-5
    ids, _ = R.tokenize(mdir, random_texts(4, seed=6))
    with torch.no_grad():
        hf_t = m.get_text_features(input_ids=torch.from_numpy(ids))
        hf_t = getattr(hf_t, "pooler_output", hf_t)
        hf_t = torch.nn.functional.normalize(hf_t, dim=-1).numpy()
    mine_t = R.text_forward(o.text, ids)
    assert cosine_rows(mine_t, hf_t).min() > 0.99999 and np.abs(mine_t - hf_t).max() < 5e-5


def test_oracle_vs_transformers_siglip(make_model):
    from transformers import SiglipConfig, SiglipModel

    import export_synthetic as ex

    config = "tiny_siglip"
    spec = ex.CONFIGS[config]
    mdir = make_model(config)
    o = R.OracleClip(mdir, threads=2)
    v, t = spec.vision, spec.text
    cfg = SiglipConfig(
        vision_config=dict(hidden_size=v.width, intermediate_size=v.mlp_dim, num_hidden_layers=v.layers,
                           num_attention_heads=v.heads, image_size=v.image_size, patch_size=v.patch,
                           hidden_act="gelu_pytorch_tanh", layer_norm_eps=v.eps),
        text_config=dict(hidden_size=t.width, intermediate_size=t.mlp_dim, num_hidden_layers=t.layers,
                         num_attention_heads=t.heads, max_position_embeddings=t.context_length,
                         vocab_size=t.vocab_size, hidden_act="gelu_pytorch_tanh", layer_norm_eps=t.eps,
                         projection_size=spec.embed_dim))
    m = SiglipModel(cfg).eval()
    w = o.vision.w
    D = v.width
    vm = m.vision_model
    pre = "model.visual.trunk"
    _copy(vm.embeddings.patch_embedding.weight, w[f"{pre}.patch_embed.proj.weight"])
    _copy(vm.embeddings.patch_embedding.bias, w[f"{pre}.patch_embed.proj.bias"])
    _copy(vm.embeddings.position_embedding.weight, w[f"{pre}.pos_embed"][0])
    for i, layer in enumerate(vm.encoder.layers):
        p = f"{pre}.blocks.{i}"
        wi, bi = w[f"{p}.attn.qkv.weight"], w[f"{p}.attn.qkv.bias"]
        for j, proj in enumerate((layer.self_attn.q_proj, layer.self_attn.k_proj, layer.self_attn.v_proj)):
            _copy(proj.weight, wi[j * D:(j + 1) * D]); _copy(proj.bias, bi[j * D:(j + 1) * D])
        _copy(layer.self_attn.out_proj.weight, w[f"{p}.attn.proj.weight"]); _copy(layer.self_attn.out_proj.bias, w[f"{p}.attn.proj.bias"])
        _copy(layer.layer_norm1.weight, w[f"{p}.norm1.weight"]); _copy(layer.layer_norm1.bias, w[f"{p}.norm1.bias"])
        _copy(layer.layer_norm2.weight, w[f"{p}.norm2.weight"]); _copy(layer.layer_norm2.bias, w[f"{p}.norm2.bias"])
        _copy(layer.mlp.fc1.weight, w[f"{p}.mlp.fc1.weight"]); _copy(layer.mlp.fc1.bias, w[f"{p}.mlp.fc1.bias"])
        _copy(layer.mlp.fc2.weight, w[f"{p}.mlp.fc2.weight"]); _copy(layer.mlp.fc2.bias, w[f"{p}.mlp.fc2.bias"])
    _copy(vm.post_layernorm.weight, w[f"{pre}.norm.weight"]); _copy(vm.post_layernorm.bias, w[f"{pre}.norm.bias"])
    ap = f"{pre}.attn_pool"
    head = vm.head
    _copy(head.probe, w[f"{ap}.latent"])
    _copy(head.attention.in_proj_weight, torch.cat([w[f"{ap}.q.weight"], w[f"{ap}.kv.weight"]], 0))
    _copy(head.attention.in_proj_bias, torch.cat([w[f"{ap}.q.bias"], w[f"{ap}.kv.bias"]], 0))
    _copy(head.attention.out_proj.weight, w[f"{ap}.proj.weight"]); _copy(head.attention.out_proj.bias, w[f"{ap}.proj.bias"])
    _copy(head.layernorm.weight, w[f"{ap}.norm.weight"]); _copy(head.layernorm.bias, w[f"{ap}.norm.bias"])
    _copy(head.mlp.fc1.weight, w[f"{ap}.mlp.fc1.weight"]); _copy(head.mlp.fc1.bias, w[f"{ap}.mlp.fc1.bias"])
    _copy(head.mlp.fc2.weight, w[f"{ap}.mlp.fc2.weight"]); _copy(head.mlp.fc2.bias, w[f"{ap}.mlp.fc2.bias"])
    wt = o.text.w
    tm = m.text_model
    _copy(tm.embeddings.token_embedding.weight, wt["model.text.token_embedding.weight"])
    _copy(tm.embeddings.position_embedding.weight, wt["model.text.positional_embedding"])
    _copy_clip_layers(tm.encoder.layers, wt, "model.text.transformer", t.width)
    _copy(tm.final_layer_norm.weight, wt["model.text.ln_final.weight"]); _copy(tm.final_layer_norm.bias, wt["model.text.ln_final.bias"])
    _copy(tm.head.weight, wt["model.text.text_projection.weight"]); _copy(tm.head.bias, wt["model.text.text_projection.bias"])

    imgs = random_images(3, v.image_size, seed=5)
    pc = o.config["preprocess_cfg"]
    pv = R.preprocess_batch(list(imgs), v.image_size, pc["mean"], pc["std"])
    with torch.no_grad():
        hf_v = m.get_image_features(pixel_values=torch.from_numpy(pv))
        hf_v = getattr(hf_v, "pooler_output", hf_v)
        hf_v = torch.nn.functional.normalize(hf_v, dim=-1).numpy()
    mine_v = R.vision_forward(o.vision, pv)
    assert cosine_rows(mine_v, hf_v).min() > 0.99999 and np.abs(mine_v - hf_v).max() < 5e-5
    ids, _ = R.tokenize(mdir, random_texts(4, seed=6))
    with torch.no_grad():
        hf_t = m.get_text_features(input_ids=torch.from_numpy(ids))
        hf_t = getattr(hf_t, "pooler_output", hf_t)
        hf_t = torch.nn.functional.normalize(hf_t, dim=-1).numpy()
    mine_t = R.text_forward(o.text, ids)
    assert cosine_rows(mine_t, hf_t).min() > 0.99999 and np.abs(mine_t - hf_t).max() < 5e-5
