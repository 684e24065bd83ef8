"""Synthetic stand-in for the reference's offline exporter (`/root/reference/pull_onnx.py`).

There is no network here, so instead of `open_clip.create_model_and_transforms("hf-hub:<id>")`
(`pull_onnx.py:102`) this script draws random-init weights for the SAME architectures,
names them the way open_clip / timm name their parameters under the exporter's `model.` wrapper prefix
(`pull_onnx.py:53-68`), and writes a model directory with the nine files the reference insists on
(`/root/reference/src/model_manager.rs:8-18`):

    visual.onnx  visual.onnx.data  text.onnx  text.onnx.data  open_clip_config.json  model_config.json
    tokenizer.json  tokenizer_config.json  special_tokens_map.json

I/O contract per `pull_onnx.py:279-302`: `pixel_values` f32 [batch,3,S,S] -> `image_embeddings` f32 [batch,D];
`input_ids` i64 [batch,ctx] -> `text_embeddings` f32 [batch,D]; opset 18; dynamic batch axis.
`model_config.json` follows `pull_onnx.py:128-150`; `open_clip_config.json` has the fields
`/root/reference/src/config.rs:24-57` parses (extra keys are ignored by serde).

The .onnx files carry the initializers (open_clip names, torch layouts, fp32, external data) plus
`metadata_props` entries `clipb200.*` with the hyper-parameters that are not derivable from tensor shapes
(heads, activation, eps, pooling, causal mask).

Usage:  python tools/export_synthetic.py --config vit_b32 --output /tmp/models
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from dataclasses import asdict, dataclass, field
from typing import Dict, List, Optional

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import onnx_proto as op  # noqa: E402

OPENAI_MEAN = [0.48145466, 0.4578275, 0.40821073]
OPENAI_STD = [0.26862954, 0.26130258, 0.27577711]


@dataclass
class VisionSpec:
    family: str  # "clip" (open_clip VisionTransformer) | "timm" (timm trunk, SigLIP style)
    image_size: int
    patch: int
    width: int
    layers: int
    heads: int
    mlp_dim: int
    act: str  # quick_gelu | gelu_tanh | gelu
    eps: float
    pool: str  # cls | map | avg
    # FastViT (family "fastvit", MobileCLIP2): per-stage widths / depths, SE on the downsample of those stages
    dims: tuple = ()
    depths: tuple = ()
    se_down: tuple = ()
    mlp_ratio: int = 3
    layer_scale: float = 0.1  # synthetic layer-scale gamma (conditioning of the random network, see fastvit_synth.py)
    attn_stages: int = 1  # trailing stages whose token mixer is attention (+ RepCPE): 1 for MCi2, 2 for MCi3 / MCi4


@dataclass
class TextSpec:
    family: str  # "clip" (model.* names) | "custom" (model.text.* names)
    context_length: int
    vocab_size: int
    width: int
    layers: int
    heads: int
    mlp_dim: int
    act: str
    eps: float
    causal: bool
    pool: str  # argmax | last
    proj_bias: bool


@dataclass
class ModelSpec:
    name: str
    embed_dim: int
    vision: VisionSpec
    text: TextSpec
    mean: List[float]
    std: List[float]
    interpolation: str
    resize_mode: str
    logit_scale: float
    logit_bias: float
    activation_function: str  # softmax | sigmoid
    tokenizer_needs_lowercase: bool
    pad_id: int
    timm_model_name: Optional[str] = None
    extra: Dict = field(default_factory=dict)


def _siglip2(name, width, layers, heads, mlp, embed, timm_name, image=384, patch=16, tlayers=27, twidth=1152,
             tmlp=4304, theads=16, vocab=256000, ctx=64):
    return ModelSpec(
        name=name, embed_dim=embed,
        vision=VisionSpec("timm", image, patch, width, layers, heads, mlp, "gelu_tanh", 1e-6, "map"),
        text=TextSpec("custom", ctx, vocab, twidth, tlayers, theads, tmlp, "gelu_tanh", 1e-6, False, "last", True),
        mean=[0.5, 0.5, 0.5], std=[0.5, 0.5, 0.5], interpolation="bicubic", resize_mode="squash",
        logit_scale=112.0, logit_bias=-16.5, activation_function="sigmoid", tokenizer_needs_lowercase=True,
        pad_id=0, timm_model_name=timm_name)


def _clip(name, image, patch, width, layers, heads, mlp, embed, twidth, tlayers, theads, tmlp, ctx=77,
          vocab=49408):
    return ModelSpec(
        name=name, embed_dim=embed,
        vision=VisionSpec("clip", image, patch, width, layers, heads, mlp, "quick_gelu", 1e-5, "cls"),
        text=TextSpec("clip", ctx, vocab, twidth, tlayers, theads, tmlp, "quick_gelu", 1e-5, True, "argmax", False),
        mean=OPENAI_MEAN, std=OPENAI_STD, interpolation="bicubic", resize_mode="shortest",
        logit_scale=100.0, logit_bias=0.0, activation_function="softmax", tokenizer_needs_lowercase=False,
        pad_id=0)


CONFIGS: Dict[str, ModelSpec] = {
    # BASELINE.json configs[0]: timm/vit_base_patch32_clip_224.openai
    "vit_b32": _clip("vit_base_patch32_clip_224.openai", 224, 32, 768, 12, 12, 3072, 512, 512, 12, 8, 2048),
    # BASELINE.json configs[2]: ViT-SO400M-16-SigLIP2-384
    "so400m_siglip2_384": _siglip2("ViT-SO400M-16-SigLIP2-384", 1152, 27, 16, 4304, 1152,
                                   "vit_so400m_patch16_siglip_384"),
    # BASELINE.json configs[3]: DFN5B-CLIP-ViT-H-14-378
    "dfn5b_h14_378": _clip("DFN5B-CLIP-ViT-H-14-378", 378, 14, 1280, 32, 16, 5120, 1024, 1024, 24, 16, 4096),
    # BASELINE.json configs[4]: ViT-gopt-16-SigLIP2-384
    "gopt_siglip2_384": _siglip2("ViT-gopt-16-SigLIP2-384", 1536, 40, 16, 6144, 1536,
                                 "vit_giantopt_patch16_siglip_384"),
    # Small shapes for fast CPU/GPU parity tests; they keep the awkward properties of the big ones
    # (head_dim 72, MLP not a multiple of 64, CLS token, T not a multiple of 64, K=588-style patch size).
    "tiny_clip": _clip("tiny-clip", 64, 16, 128, 2, 2, 512, 64, 128, 2, 2, 512, ctx=77, vocab=49408),
    "tiny_clip_p14": _clip("tiny-clip-p14", 70, 14, 160, 2, 2, 640, 96, 128, 2, 2, 512, ctx=77, vocab=49408),
    "tiny_siglip": _siglip2("tiny-siglip", 144, 2, 2, 536, 144, "tiny_siglip", image=64, patch=16,
                            tlayers=2, twidth=144, tmlp=536, theads=2, vocab=49412, ctx=64),
    "small_siglip": _siglip2("small-siglip", 576, 4, 8, 2152, 576, "small_siglip", image=384, patch=16,
                             tlayers=3, twidth=576, tmlp=2152, theads=8, vocab=49412, ctx=64),
}


def _mobileclip(name, image, dims, depths, embed, twidth=512, tlayers=12, theads=8, tmlp=2048, mlp_ratio=3,
                attn_stages=1, timm_name="fastvit_mci2", layer_scale=0.1):
    """MobileCLIP2 shapes: re-parameterised FastViT trunk + non-causal text tower.  S2 = MCi2
    (4 stages, attention in the last); S3 / S4 = MCi3 / MCi4 (5 stages, attention in the last two, MLP ratio 4)."""
    n = len(dims)
    return ModelSpec(
        name=name, embed_dim=embed,
        vision=VisionSpec("fastvit", image, 4, dims[-1], sum(depths), dims[-1] // 32, dims[-1] * mlp_ratio, "gelu", 1e-5, "avg",
                          dims=tuple(dims), depths=tuple(depths), se_down=tuple(i >= 2 for i in range(n)),
                          mlp_ratio=mlp_ratio, attn_stages=attn_stages, layer_scale=layer_scale),
        text=TextSpec("custom", 77, 49408, twidth, tlayers, theads, tmlp, "gelu", 1e-5, False, "argmax", False),
        mean=[0.0, 0.0, 0.0], std=[1.0, 1.0, 1.0], interpolation="bilinear", resize_mode="shortest",
        logit_scale=100.0, logit_bias=0.0, activation_function="softmax", tokenizer_needs_lowercase=False, pad_id=0,
        timm_model_name=timm_name)


# BASELINE.json configs[1]: MobileCLIP2-S2
CONFIGS["mobileclip2_s2"] = _mobileclip("MobileCLIP2-S2", 256, (80, 160, 320, 640), (4, 12, 24, 4), 512)
CONFIGS["tiny_mobileclip"] = _mobileclip("tiny-mobileclip", 64, (32, 64, 128, 256), (1, 2, 2, 2), 64, twidth=128,
                                         tlayers=2, theads=2, tmlp=512)
# MobileCLIP2-S3 / S4 (benches/model_bench.rs:10-12): 5-stage FastViT-MCi3 / MCi4 [upstream recall: widths, depths and
# text-tower sizes as in timm / open_clip; only the layout matters for the engine, which derives it from the file]
CONFIGS["mobileclip2_s3"] = _mobileclip("MobileCLIP2-S3", 256, (96, 192, 384, 768, 1536), (2, 12, 24, 4, 2), 768,
                                        twidth=768, tlayers=12, theads=12, tmlp=3072, mlp_ratio=4, attn_stages=2,
                                        timm_name="fastvit_mci3", layer_scale=0.05)
CONFIGS["mobileclip2_s4"] = _mobileclip("MobileCLIP2-S4", 256, (128, 256, 512, 1024, 2048), (2, 12, 24, 4, 4), 768,
                                        twidth=768, tlayers=12, theads=12, tmlp=3072, mlp_ratio=4, attn_stages=2,
                                        timm_name="fastvit_mci4", layer_scale=0.05)
CONFIGS["tiny_mobileclip5"] = _mobileclip("tiny-mobileclip-5stage", 128, (16, 32, 64, 128, 256), (1, 2, 2, 2, 1), 64,
                                          twidth=128, tlayers=2, theads=2, tmlp=512, mlp_ratio=4, attn_stages=2,
                                          timm_name="fastvit_mci3", layer_scale=0.05)
CONFIGS["tiny_siglip"].text.vocab_size = 49412
CONFIGS["small_siglip"].text.vocab_size = 49412


# ----------------------------------------------------------------------------- weights
class WeightGen:
    """Seeded fp32 generator.  Scales follow SURVEY.md Appendix A ("Init scales for random fixtures"); every LN
    gamma/beta and bias gets small distinct noise so that gamma/beta/bias bugs are visible in parity tests."""

    def __init__(self, seed: int):
        self.rng = np.random.default_rng(seed)

    def normal(self, shape, std):
        a = self.rng.standard_normal(shape, dtype=np.float32)
        a *= np.float32(std)
        return a

    def ln_w(self, n):
        return (1.0 + 0.1 * self.rng.standard_normal(n, dtype=np.float32)).astype(np.float32)

    def small(self, n, std=0.02):
        return self.normal((n,), std)


def _clip_resblocks(g: WeightGen, prefix: str, width: int, layers: int, mlp: int, emit) -> None:
    attn_std = width ** -0.5
    proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
    fc_std = (2 * width) ** -0.5
    for i in range(layers):
        p = f"{prefix}.resblocks.{i}"
        emit(f"{p}.ln_1.weight", g.ln_w(width)); emit(f"{p}.ln_1.bias", g.small(width))
        emit(f"{p}.attn.in_proj_weight", g.normal((3 * width, width), attn_std))
        emit(f"{p}.attn.in_proj_bias", g.small(3 * width))
        emit(f"{p}.attn.out_proj.weight", g.normal((width, width), proj_std))
        emit(f"{p}.attn.out_proj.bias", g.small(width))
        emit(f"{p}.ln_2.weight", g.ln_w(width)); emit(f"{p}.ln_2.bias", g.small(width))
        emit(f"{p}.mlp.c_fc.weight", g.normal((mlp, width), fc_std)); emit(f"{p}.mlp.c_fc.bias", g.small(mlp))
        emit(f"{p}.mlp.c_proj.weight", g.normal((width, mlp), proj_std)); emit(f"{p}.mlp.c_proj.bias", g.small(width))


def gen_vision(spec: ModelSpec, seed: int, emit) -> None:
    v = spec.vision
    g = WeightGen(seed)
    D, P = v.width, v.patch
    T = (v.image_size // P) ** 2
    if v.family == "clip":
        pre = "model.visual"
        emit(f"{pre}.conv1.weight", g.normal((D, 3, P, P), (3 * P * P) ** -0.5))
        emit(f"{pre}.class_embedding", g.normal((D,), D ** -0.5))
        emit(f"{pre}.positional_embedding", g.normal((T + 1, D), D ** -0.5))
        emit(f"{pre}.ln_pre.weight", g.ln_w(D)); emit(f"{pre}.ln_pre.bias", g.small(D))
        _clip_resblocks(g, f"{pre}.transformer", D, v.layers, v.mlp_dim, emit)
        emit(f"{pre}.ln_post.weight", g.ln_w(D)); emit(f"{pre}.ln_post.bias", g.small(D))
        emit(f"{pre}.proj", g.normal((D, spec.embed_dim), D ** -0.5))
    elif v.family == "timm":
        pre = "model.visual.trunk"
        w_std = 0.02 if D >= 512 else D ** -0.5  # keep tiny test models from collapsing to their biases
        emit(f"{pre}.patch_embed.proj.weight", g.normal((D, 3, P, P), (3 * P * P) ** -0.5))
        emit(f"{pre}.patch_embed.proj.bias", g.small(D))
        emit(f"{pre}.pos_embed", g.normal((1, T, D), 0.02 if D >= 512 else 0.2))
        for i in range(v.layers):
            p = f"{pre}.blocks.{i}"
            emit(f"{p}.norm1.weight", g.ln_w(D)); emit(f"{p}.norm1.bias", g.small(D))
            emit(f"{p}.attn.qkv.weight", g.normal((3 * D, D), w_std)); emit(f"{p}.attn.qkv.bias", g.small(3 * D))
            emit(f"{p}.attn.proj.weight", g.normal((D, D), w_std)); emit(f"{p}.attn.proj.bias", g.small(D))
            emit(f"{p}.norm2.weight", g.ln_w(D)); emit(f"{p}.norm2.bias", g.small(D))
            emit(f"{p}.mlp.fc1.weight", g.normal((v.mlp_dim, D), w_std)); emit(f"{p}.mlp.fc1.bias", g.small(v.mlp_dim))
            emit(f"{p}.mlp.fc2.weight", g.normal((D, v.mlp_dim), w_std)); emit(f"{p}.mlp.fc2.bias", g.small(D))
        emit(f"{pre}.norm.weight", g.ln_w(D)); emit(f"{pre}.norm.bias", g.small(D))
        ap = f"{pre}.attn_pool"
        emit(f"{ap}.latent", g.normal((1, 1, D), D ** -0.5))
        emit(f"{ap}.q.weight", g.normal((D, D), w_std)); emit(f"{ap}.q.bias", g.small(D))
        emit(f"{ap}.kv.weight", g.normal((2 * D, D), w_std)); emit(f"{ap}.kv.bias", g.small(2 * D))
        emit(f"{ap}.proj.weight", g.normal((D, D), w_std)); emit(f"{ap}.proj.bias", g.small(D))
        emit(f"{ap}.norm.weight", g.ln_w(D)); emit(f"{ap}.norm.bias", g.small(D))
        emit(f"{ap}.mlp.fc1.weight", g.normal((v.mlp_dim, D), w_std)); emit(f"{ap}.mlp.fc1.bias", g.small(v.mlp_dim))
        emit(f"{ap}.mlp.fc2.weight", g.normal((D, v.mlp_dim), w_std)); emit(f"{ap}.mlp.fc2.bias", g.small(D))
    elif v.family == "fastvit":
        import fastvit_synth  # calibrated folded-BatchNorm generator (needs torch for the calibration forward)

        fastvit_synth.gen_fastvit(spec, g, emit)
    else:
        raise ValueError(v.family)


def gen_fastvit(spec: ModelSpec, g: "WeightGen", emit) -> None:
    """Re-parameterised (pull_onnx.py:110-116) FastViT trunk with eval-mode BatchNorm folded into the preceding conv,
    timm parameter names: every MobileOne / RepMixer / large-kernel block is a single `reparam_conv`."""
    v = spec.vision
    pre = "model.visual.trunk"

    def conv(name, cout, cin_per_group, k, bias_std=0.02, gain=1.0):
        fan_in = cin_per_group * k * k
        emit(f"{name}.weight", g.normal((cout, cin_per_group, k, k), gain * fan_in ** -0.5))
        emit(f"{name}.bias", g.small(cout, bias_std))

    def se(name, ch):
        rd = max(ch // 16, 8)
        conv(f"{name}.fc1", rd, ch, 1)
        conv(f"{name}.fc2", ch, rd, 1)

    def mlp(name, c):
        conv(f"{name}.conv.conv", c, 1, 7)  # depthwise 7x7 with its BatchNorm folded in
        conv(f"{name}.fc1", v.mlp_ratio * c, c, 1, gain=1.4)
        conv(f"{name}.fc2", c, v.mlp_ratio * c, 1)

    d0 = v.dims[0]
    conv(f"{pre}.stem.0.reparam_conv", d0, 3, 3, gain=1.4)
    conv(f"{pre}.stem.1.reparam_conv", d0, 1, 3, gain=1.4)
    conv(f"{pre}.stem.2.reparam_conv", d0, d0, 1, gain=1.4)
    prev = d0
    for i, (c, depth) in enumerate(zip(v.dims, v.depths)):
        st = f"{pre}.stages.{i}"
        if i > 0:
            conv(f"{st}.downsample.proj.0.reparam_conv", c, 1, 7, gain=1.4)  # groups = prev, multiplier c / prev
            if v.se_down[i]:
                se(f"{st}.downsample.proj.0.se", c)
            conv(f"{st}.downsample.proj.1.reparam_conv", c, c, 1, gain=1.4)
        last = i == len(v.dims) - 1
        if last:
            conv(f"{st}.pos_emb.reparam_conv", c, 1, 7)
        for j in range(depth):
            b = f"{st}.blocks.{j}"
            if not last:
                conv(f"{b}.token_mixer.reparam_conv", c, 1, 3)
                mlp(f"{b}.mlp", c)
                emit(f"{b}.layer_scale.gamma", (0.3 + 0.05 * g.rng.standard_normal((c, 1, 1))).astype(np.float32))
            else:
                emit(f"{b}.norm.weight", g.ln_w(c)); emit(f"{b}.norm.bias", g.small(c))
                emit(f"{b}.norm.running_mean", g.small(c, 0.1))
                emit(f"{b}.norm.running_var", (1.0 + 0.2 * g.rng.random(c)).astype(np.float32))
                emit(f"{b}.token_mixer.qkv.weight", g.normal((3 * c, c), c ** -0.5))
                emit(f"{b}.token_mixer.proj.weight", g.normal((c, c), c ** -0.5))
                emit(f"{b}.token_mixer.proj.bias", g.small(c))
                emit(f"{b}.layer_scale_1.gamma", (0.3 + 0.05 * g.rng.standard_normal((c, 1, 1))).astype(np.float32))
                mlp(f"{b}.mlp", c)
                emit(f"{b}.layer_scale_2.gamma", (0.3 + 0.05 * g.rng.standard_normal((c, 1, 1))).astype(np.float32))
        prev = c
    cf = 2 * prev
    conv(f"{pre}.final_conv.reparam_conv", cf, 1, 3, gain=1.4)
    se(f"{pre}.final_conv.se", cf)
    emit(f"{pre}.head.fc.weight", g.normal((spec.embed_dim, cf), cf ** -0.5))
    emit(f"{pre}.head.fc.bias", g.small(spec.embed_dim))


def gen_text(spec: ModelSpec, seed: int, emit) -> None:
    t = spec.text
    g = WeightGen(seed + 1000003)
    D = t.width
    pre = "model" if t.family == "clip" else "model.text"
    emit(f"{pre}.token_embedding.weight", g.normal((t.vocab_size, D), 0.02))
    emit(f"{pre}.positional_embedding", g.normal((t.context_length, D), 0.01))
    _clip_resblocks(g, f"{pre}.transformer", D, t.layers, t.mlp_dim, emit)
    emit(f"{pre}.ln_final.weight", g.ln_w(D)); emit(f"{pre}.ln_final.bias", g.small(D))
    if t.proj_bias:
        emit(f"{pre}.text_projection.weight", g.normal((spec.embed_dim, D), D ** -0.5))
        emit(f"{pre}.text_projection.bias", g.small(spec.embed_dim))
    else:
        emit(f"{pre}.text_projection", g.normal((D, spec.embed_dim), D ** -0.5))


# ----------------------------------------------------------------------------- tokenizer
def _bytes_to_unicode():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + \
        list(range(ord("\xae"), ord("\xff") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


def build_tokenizer(kind: str, vocab_size: int):
    """A CLIP-style byte-level BPE with a synthetic, deterministic merge table (letters -> bigrams -> trigrams ->
    4-grams) sized so that the two special tokens are the two highest ids (EOT = max id, which EOT-argmax pooling
    relies on).  kind="clip": <|startoftext|>/<|endoftext|> wrapped around the text (open_clip SimpleTokenizer
    behaviour); kind="siglip": <pad>=0,<eos>=1,<bos>=2,<unk>=3 first, EOS appended."""
    from tokenizers import Regex, Tokenizer, decoders, models, normalizers, pre_tokenizers, processors

    b2u = _bytes_to_unicode()
    chars = [b2u[b] for b in range(256)]
    n_special_front = 4 if kind == "siglip" else 0
    n_special_back = 2 if kind == "clip" else 0
    budget = min(vocab_size, 49408 + n_special_front) - n_special_front - n_special_back
    vocab: Dict[str, int] = {}
    if kind == "siglip":
        for i, s in enumerate(["<pad>", "<eos>", "<bos>", "<unk>"]):
            vocab[s] = i
    for c in chars:
        vocab[c] = len(vocab)
    for c in chars:
        vocab[c + "</w>"] = len(vocab)
    merges = []
    letters = [chr(c) for c in range(ord("a"), ord("z") + 1)]

    def add(a, b):
        if len(vocab) - n_special_front >= budget:
            return False
        tok = a + b
        if tok in vocab:
            return True
        merges.append((a, b))
        vocab[tok] = len(vocab)
        return True

    ok = True
    for x in letters:
        for y in letters:
            ok = ok and add(x, y + "</w>") and add(x, y)
    for x in letters:
        for y in letters:
            for z in letters:
                if not ok:
                    break
                ok = add(x + y, z + "</w>") and add(x + y, z)
    for x in letters:
        for y in letters:
            for z in letters:
                for w in letters:
                    if not ok:
                        break
                    ok = add(x + y, z + w + "</w>")
    if kind == "clip":
        vocab["<|startoftext|>"] = len(vocab)
        vocab["<|endoftext|>"] = len(vocab)
        unk = "<|endoftext|>"
    else:
        unk = "<unk>"
    model = models.BPE(vocab=vocab, merges=merges, unk_token=unk, continuing_subword_prefix="",
                       end_of_word_suffix="</w>", fuse_unk=False)
    tok = Tokenizer(model)
    if kind == "clip":
        tok.normalizer = normalizers.Sequence([normalizers.NFC(), normalizers.Replace(Regex(r"\s+"), " "),
                                               normalizers.Lowercase()])
    else:
        tok.normalizer = normalizers.Sequence([normalizers.NFC(), normalizers.Replace(Regex(r"\s+"), " ")])
    pattern = r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+"
    tok.pre_tokenizer = pre_tokenizers.Sequence([
        pre_tokenizers.Split(Regex(pattern), behavior="removed", invert=True),
        pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    tok.decoder = decoders.ByteLevel()
    if kind == "clip":
        sot, eot = vocab["<|startoftext|>"], vocab["<|endoftext|>"]
        tok.post_processor = processors.RobertaProcessing(sep=("<|endoftext|>", eot), cls=("<|startoftext|>", sot),
                                                          trim_offsets=False, add_prefix_space=False)
        tok.add_special_tokens(["<|startoftext|>", "<|endoftext|>"])
        specials = {"bos_token": "<|startoftext|>", "eos_token": "<|endoftext|>", "unk_token": "<|endoftext|>",
                    "pad_token": "<|endoftext|>"}
    else:
        tok.post_processor = processors.TemplateProcessing(single="$A <eos>", special_tokens=[("<eos>", 1)])
        tok.add_special_tokens(["<pad>", "<eos>", "<bos>", "<unk>"])
        specials = {"bos_token": "<bos>", "eos_token": "<eos>", "unk_token": "<unk>", "pad_token": "<pad>"}
    return tok, specials


# ----------------------------------------------------------------------------- model dir
ACT_IDS = {"none": 0, "quick_gelu": 1, "gelu_tanh": 2, "gelu": 3}


def _write_tower(path: str, graph_name: str, in_name: str, in_type: int, in_shape, out_name: str, out_dim: int,
                 meta: Dict[str, str], gen) -> None:
    w = op.ModelWriter(path, graph_name)
    w.add_input(in_name, in_type, in_shape)
    w.add_output(out_name, op.FLOAT, ["batch_size", out_dim])
    for k, v in meta.items():
        w.add_metadata(k, str(v))
    gen(lambda name, arr: w.add_initializer(name, arr))
    w.close()


def vision_meta(spec: ModelSpec) -> Dict[str, str]:
    v = spec.vision
    return {"clipb200.tower": "vision", "clipb200.family": v.family, "clipb200.image_size": v.image_size,
            "clipb200.patch": v.patch, "clipb200.width": v.width, "clipb200.layers": v.layers,
            "clipb200.heads": v.heads, "clipb200.mlp_dim": v.mlp_dim, "clipb200.act": ACT_IDS[v.act],
            "clipb200.eps": repr(v.eps), "clipb200.pool": v.pool, "clipb200.embed_dim": spec.embed_dim,
            **({"clipb200.dims": ",".join(map(str, v.dims)), "clipb200.depths": ",".join(map(str, v.depths)),
                "clipb200.se_down": ",".join(str(int(x)) for x in v.se_down), "clipb200.mlp_ratio": v.mlp_ratio,
                "clipb200.attn_stages": v.attn_stages}
               if v.family == "fastvit" else {})}


def text_meta(spec: ModelSpec) -> Dict[str, str]:
    t = spec.text
    return {"clipb200.tower": "text", "clipb200.family": t.family, "clipb200.context_length": t.context_length,
            "clipb200.vocab_size": t.vocab_size, "clipb200.width": t.width, "clipb200.layers": t.layers,
            "clipb200.heads": t.heads, "clipb200.mlp_dim": t.mlp_dim, "clipb200.act": ACT_IDS[t.act],
            "clipb200.eps": repr(t.eps), "clipb200.pool": t.pool, "clipb200.causal": int(t.causal),
            "clipb200.embed_dim": spec.embed_dim}


def write_model_dir(spec: ModelSpec, out_dir: str, seed: int = 0, towers=("vision", "text")) -> str:
    os.makedirs(out_dir, exist_ok=True)
    v, t = spec.vision, spec.text
    if "vision" in towers:
        _write_tower(os.path.join(out_dir, "visual.onnx"), "visual", "pixel_values", op.FLOAT,
                     ["batch_size", 3, v.image_size, v.image_size], "image_embeddings", spec.embed_dim,
                     vision_meta(spec), lambda emit: gen_vision(spec, seed, emit))
    else:  # the reference refuses a directory without all nine files (model_manager.rs:57-65)
        for fn in ("visual.onnx", "visual.onnx.data"):
            open(os.path.join(out_dir, fn), "ab").close()
    if "text" in towers:
        _write_tower(os.path.join(out_dir, "text.onnx"), "text", "input_ids", op.INT64,
                     ["batch_size", t.context_length], "text_embeddings", spec.embed_dim,
                     text_meta(spec), lambda emit: gen_text(spec, seed, emit))
    else:
        for fn in ("text.onnx", "text.onnx.data"):
            open(os.path.join(out_dir, fn), "ab").close()

    vision_cfg = {"image_size": v.image_size, "patch_size": v.patch, "width": v.width, "layers": v.layers,
                  "heads": v.heads, "mlp_dim": v.mlp_dim}
    if spec.timm_model_name:
        vision_cfg.update({"timm_model_name": spec.timm_model_name,
                           "timm_pool": "avg" if v.family == "fastvit" else "map", "timm_proj": "none"})
    open_clip_config = {
        "model_cfg": {
            "embed_dim": spec.embed_dim,
            "vision_cfg": vision_cfg,
            "text_cfg": {"context_length": t.context_length, "vocab_size": t.vocab_size, "width": t.width,
                         "heads": t.heads, "layers": t.layers, "mlp_dim": t.mlp_dim,
                         "no_causal_mask": not t.causal, "pool_type": t.pool, "proj_bias": t.proj_bias},
            "quick_gelu": v.act == "quick_gelu",
        },
        "preprocess_cfg": {"mean": spec.mean, "std": spec.std, "interpolation": spec.interpolation,
                           "resize_mode": spec.resize_mode},
    }
    if spec.activation_function == "sigmoid":
        open_clip_config["model_cfg"]["init_logit_bias"] = -10
    with open(os.path.join(out_dir, "open_clip_config.json"), "w") as f:
        json.dump(open_clip_config, f, indent=2)
    model_config = {"logit_scale": spec.logit_scale, "logit_bias": spec.logit_bias,
                    "activation_function": spec.activation_function,
                    "tokenizer_needs_lowercase": spec.tokenizer_needs_lowercase, "pad_id": spec.pad_id,
                    "vocab_size": t.vocab_size}
    with open(os.path.join(out_dir, "model_config.json"), "w") as f:
        json.dump(model_config, f, indent=2)
    kind = "clip" if t.family == "clip" else "siglip"
    tok, specials = build_tokenizer(kind, t.vocab_size)
    tok.save(os.path.join(out_dir, "tokenizer.json"))
    with open(os.path.join(out_dir, "tokenizer_config.json"), "w") as f:
        json.dump({"model_max_length": t.context_length, "tokenizer_class": "PreTrainedTokenizerFast", **specials},
                  f, indent=2)
    with open(os.path.join(out_dir, "special_tokens_map.json"), "w") as f:
        json.dump(specials, f, indent=2)
    return out_dir


def main() -> None:
    ap = argparse.ArgumentParser(description="Write a synthetic open_clip_inference model directory.")
    ap.add_argument("--config", required=True, choices=sorted(CONFIGS))
    ap.add_argument("--output", required=True, help="base output directory; the model goes to <output>/<config>")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--towers", default="vision,text")
    a = ap.parse_args()
    d = write_model_dir(CONFIGS[a.config], os.path.join(a.output, a.config), a.seed, tuple(a.towers.split(",")))
    print(d)


if __name__ == "__main__":
    main()
