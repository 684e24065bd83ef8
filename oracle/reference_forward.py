"""CPU ORACLE (test infrastructure, NOT the product).  PARITY UNPINNED.

A plain torch-fp32 / numpy restatement of what the reference computes on its hot path.  It exists only so that
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs can check and time
something on the CPU; nothing under `clip_embedder_rs_b200/` may import it.

Why "parity unpinned": the reference delegates all model math to ONNX Runtime (crate `ort` 2.0.0-rc.12,
`/root/reference/Cargo.toml:13`), which is a downloaded binary that is absent from this image together with
`onnx`, `onnxruntime`, `open_clip` and `timm`; the reference's only test
(`/root/reference/tests/integration_test.rs:9-36`) needs network weights and holds no numeric vectors.  So this
restatement cannot be checked against the reference's own outputs.  What it follows, line by line:

  * preprocessing ........ `/root/reference/src/vision.rs:235-259` (normalize_pixels) and `:164-198` (resize; for
                           inputs already at the model resolution the convolution resize is the identity)
  * tokenisation ......... `/root/reference/src/text.rs:70-85,111-139`, executed by the SAME Rust crate
                           (`tokenizers` 0.22.2, `Cargo.lock:2807-2808`) through its Python binding
  * graph semantics ...... `/root/reference/pull_onnx.py:53-68` (`encode_image/encode_text(normalize=True)`, eval
                           mode) for the open_clip 3.2.0 / timm architectures listed in SURVEY.md Appendix A
  * similarity tail ...... `/root/reference/src/clip.rs:81-185`

The forward passes read their weights from the SAME `visual.onnx` / `text.onnx` (+ `.onnx.data`) files that the
engine loads, through the independent Python reader in `tools/onnx_proto.py`.

What it IS checked against (tests/test_oracle_cpu.py, tests/test_onnx_graph_cpu.py): HF transformers' independent CLIP /
SiglipModel implementations loaded with the same weights; real `torch.onnx.export` graphs of the same models executed by
`oracle/onnx_interp.py`; and the same exported files executed by a third-party ONNX runtime that is in the image, OpenCV
DNN (vision towers incl. the SigLIP MAP head, SigLIP text; 1e-5).  None of these is onnxruntime, hence the label above.
"""
from __future__ import annotations

import json
import math
import os
import sys
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(_ROOT, "tools"))
import onnx_proto  # noqa: E402


# ----------------------------------------------------------------------------------------------------------
# preprocessing  (src/vision.rs)
# ----------------------------------------------------------------------------------------------------------
def normalize_pixels(pixels_hwc_u8: np.ndarray, mean: Sequence[float], std: Sequence[float]) -> np.ndarray:
    """vision.rs:235-259:  out[c][i] = ((px[i*3+c] as f32) / 255.0 - mean[c]) / std[c]   (all in f32)."""
    px = np.asarray(pixels_hwc_u8)
    assert px.dtype == np.uint8 and px.ndim == 3 and px.shape[2] == 3
    out = np.empty((3, px.shape[0], px.shape[1]), dtype=np.float32)
    for c in range(3):
        val = px[:, :, c].astype(np.float32) / np.float32(255.0)
        out[c] = (val - np.float32(mean[c])) / np.float32(std[c])
    return out


def preprocess_batch(images_hwc_u8: Sequence[np.ndarray], image_size: int, mean, std,
                     interpolation: str = "bicubic", resize_mode: str = "shortest") -> np.ndarray:
    """vision.rs:120-135: resize (oracle/resize.py restates vision.rs:164-198; the identity for images that are already
    image_size x image_size) then normalise.  Empty batch -> error, as vision.rs:121-123."""
    if len(images_hwc_u8) == 0:
        raise ValueError("Empty batch")
    out = np.zeros((len(images_hwc_u8), 3, image_size, image_size), dtype=np.float32)
    for i, im in enumerate(images_hwc_u8):
        if im.shape[0] != image_size or im.shape[1] != image_size:
            from oracle import resize as _resize

            im = _resize.resize_rgb8(im, image_size, interpolation, resize_mode)
        out[i] = normalize_pixels(im, mean, std)
    return out


# ----------------------------------------------------------------------------------------------------------
# tokenisation  (src/text.rs)
# ----------------------------------------------------------------------------------------------------------
def load_tokenizer(model_dir: str):
    """text.rs:66-85: Tokenizer::from_file, pad id from model_config.pad_id else vocab["<pad>"],
    PaddingStrategy::Fixed(ctx) with that pad id, truncation to ctx."""
    from tokenizers import Tokenizer

    with open(os.path.join(model_dir, "model_config.json")) as f:
        mc = json.load(f)
    with open(os.path.join(model_dir, "open_clip_config.json")) as f:
        oc = json.load(f)
    tok = Tokenizer.from_file(os.path.join(model_dir, "tokenizer.json"))
    pad_id = mc.get("pad_id")
    if pad_id is None:
        pad_id = tok.get_vocab(True).get("<pad>")
    if pad_id is None:
        raise ValueError("No pad token found in tokenizer")
    ctx = int(oc["model_cfg"]["text_cfg"]["context_length"])
    tok.enable_padding(length=ctx, pad_id=int(pad_id))  # PaddingParams defaults: right, pad_token "[PAD]", type id 0
    tok.enable_truncation(max_length=ctx)
    return tok, bool(mc.get("tokenizer_needs_lowercase", False)), ctx


def tokenize(model_dir: str, texts: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    """text.rs:111-139 -> (ids i64 [B,ctx], mask i64 [B,ctx])."""
    tok, lower, ctx = load_tokenizer(model_dir)
    if lower:
        texts = [t.lower() for t in texts]  # Rust str::to_lowercase (Unicode) ~ Python str.lower
    enc = tok.encode_batch(list(texts), add_special_tokens=True)
    ids = np.asarray([e.ids for e in enc], dtype=np.int64).reshape(len(enc), ctx)
    mask = np.asarray([e.attention_mask for e in enc], dtype=np.int64).reshape(len(enc), ctx)
    return ids, mask


# ----------------------------------------------------------------------------------------------------------
# towers  (what session.run executes; pull_onnx.py:53-68)
# ----------------------------------------------------------------------------------------------------------
def _act(x: torch.Tensor, act: int) -> torch.Tensor:
    if act == 1:
        return x * torch.sigmoid(1.702 * x)  # open_clip QuickGELU
    if act == 2:
        return F.gelu(x, approximate="tanh")
    if act == 3:
        return F.gelu(x)
    return x


class Tower:
    def __init__(self, onnx_path: str, dtype=torch.float32):
        m = onnx_proto.read_model(onnx_path)
        self.meta = {k.replace("clipb200.", ""): v for k, v in m["metadata"].items()}
        self.inputs = m["inputs"]
        self.outputs = m["outputs"]
        self.dtype = dtype
        self.w: Dict[str, torch.Tensor] = {k: torch.from_numpy(np.array(v)).to(dtype)
                                           for k, v in m["initializers"].items()}

    def i(self, k: str) -> int:
        return int(self.meta[k])


def _mha(x, w_in, b_in, w_out, b_out, heads: int, mask: Optional[torch.Tensor]):
    B, T, D = x.shape
    hd = D // heads
    qkv = F.linear(x, w_in, b_in).reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    s = (q * (hd ** -0.5)) @ k.transpose(-1, -2)
    if mask is not None:
        s = s + mask
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, T, D)
    return F.linear(o, w_out, b_out)


def _clip_blocks(t: Tower, x, prefix: str, layers: int, heads: int, act: int, eps: float, mask):
    w = t.w
    D = x.shape[-1]
    for i in range(layers):
        p = f"{prefix}.resblocks.{i}"
        h = F.layer_norm(x, (D,), w[f"{p}.ln_1.weight"], w[f"{p}.ln_1.bias"], eps)
        x = x + _mha(h, w[f"{p}.attn.in_proj_weight"], w[f"{p}.attn.in_proj_bias"], w[f"{p}.attn.out_proj.weight"],
                     w[f"{p}.attn.out_proj.bias"], heads, mask)
        h = F.layer_norm(x, (D,), w[f"{p}.ln_2.weight"], w[f"{p}.ln_2.bias"], eps)
        h = _act(F.linear(h, w[f"{p}.mlp.c_fc.weight"], w[f"{p}.mlp.c_fc.bias"]), act)
        x = x + F.linear(h, w[f"{p}.mlp.c_proj.weight"], w[f"{p}.mlp.c_proj.bias"])
    return x


@torch.no_grad()
def vision_forward(t: Tower, pixel_values: np.ndarray, normalize: bool = True) -> np.ndarray:
    """pixel_values f32 [B,3,S,S] -> image_embeddings f32 [B,D], L2-normalised (pull_onnx.py:58-59)."""
    w = t.w
    x = torch.from_numpy(np.ascontiguousarray(pixel_values)).to(t.dtype)
    fam, P, D, L, H = t.meta["family"], t.i("patch"), t.i("width"), t.i("layers"), t.i("heads")
    act, eps = t.i("act"), float(t.meta["eps"])
    B = x.shape[0]
    if fam == "clip":
        pre = "model.visual"
        x = F.conv2d(x, w[f"{pre}.conv1.weight"], None, stride=P)
        x = x.reshape(B, D, -1).permute(0, 2, 1)
        cls = w[f"{pre}.class_embedding"].reshape(1, 1, D).expand(B, 1, D)
        x = torch.cat([cls, x], dim=1) + w[f"{pre}.positional_embedding"]
        x = F.layer_norm(x, (D,), w[f"{pre}.ln_pre.weight"], w[f"{pre}.ln_pre.bias"], eps)
        x = _clip_blocks(t, x, f"{pre}.transformer", L, H, act, eps, None)
        pooled = F.layer_norm(x[:, 0], (D,), w[f"{pre}.ln_post.weight"], w[f"{pre}.ln_post.bias"], eps)
        out = pooled @ w[f"{pre}.proj"]
    elif fam == "timm":
        pre = "model.visual.trunk"
        x = F.conv2d(x, w[f"{pre}.patch_embed.proj.weight"], w[f"{pre}.patch_embed.proj.bias"], stride=P)
        x = x.flatten(2).transpose(1, 2) + w[f"{pre}.pos_embed"]
        hd = D // H
        for i in range(L):
            p = f"{pre}.blocks.{i}"
            h = F.layer_norm(x, (D,), w[f"{p}.norm1.weight"], w[f"{p}.norm1.bias"], eps)
            x = x + _mha(h, w[f"{p}.attn.qkv.weight"], w[f"{p}.attn.qkv.bias"], w[f"{p}.attn.proj.weight"],
                         w[f"{p}.attn.proj.bias"], H, None)
            h = F.layer_norm(x, (D,), w[f"{p}.norm2.weight"], w[f"{p}.norm2.bias"], eps)
            h = _act(F.linear(h, w[f"{p}.mlp.fc1.weight"], w[f"{p}.mlp.fc1.bias"]), act)
            x = x + F.linear(h, w[f"{p}.mlp.fc2.weight"], w[f"{p}.mlp.fc2.bias"])
        x = F.layer_norm(x, (D,), w[f"{pre}.norm.weight"], w[f"{pre}.norm.bias"], eps)
        # timm AttentionPoolLatent (latent_len 1, pool "token")
        ap = f"{pre}.attn_pool"
        N = x.shape[1]
        q = F.linear(w[f"{ap}.latent"].expand(B, -1, -1), w[f"{ap}.q.weight"], w[f"{ap}.q.bias"])
        q = q.reshape(B, 1, H, hd).transpose(1, 2)
        kv = F.linear(x, w[f"{ap}.kv.weight"], w[f"{ap}.kv.bias"]).reshape(B, N, 2, H, hd).permute(2, 0, 3, 1, 4)
        k, v = kv[0], kv[1]
        s = (q * (hd ** -0.5)) @ k.transpose(-1, -2)
        o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, 1, D)
        o = F.linear(o, w[f"{ap}.proj.weight"], w[f"{ap}.proj.bias"])
        h = F.layer_norm(o, (D,), w[f"{ap}.norm.weight"], w[f"{ap}.norm.bias"], eps)
        h = _act(F.linear(h, w[f"{ap}.mlp.fc1.weight"], w[f"{ap}.mlp.fc1.bias"]), act)
        o = o + F.linear(h, w[f"{ap}.mlp.fc2.weight"], w[f"{ap}.mlp.fc2.bias"])
        out = o[:, 0]
    elif fam == "fastvit":
        out = _fastvit_forward(t, x)
    else:
        raise ValueError(fam)
    if normalize:
        out = F.normalize(out, dim=-1)
    return out.to(torch.float32).numpy()


def _fastvit_forward(t: Tower, x: torch.Tensor) -> torch.Tensor:
    """timm FastViT (fastvit_mci2 for MobileCLIP2-S2) after `reparameterize_model` (pull_onnx.py:110-116) with eval
    BatchNorm folded: stem (3x3 s2, dw3x3 s2, 1x1) -> 4 or 5 stages (RepMixer blocks; last stage(s): conditional positional
    encoding + attention blocks) with 7x7 s2 depthwise + 1x1 downsampling -> dw3x3 expansion + SE -> GAP -> Linear.
    SURVEY.md Appendix A ("C2 MobileCLIP2-S2"); architecture is upstream recall, see DESIGN.md."""
    w = t.w
    pre = "model.visual.trunk"
    dims = [int(v) for v in t.meta["dims"].split(",")]
    depths = [int(v) for v in t.meta["depths"].split(",")]
    se_down = [bool(int(v)) for v in t.meta["se_down"].split(",")]

    def conv(name, x, stride=1, groups=1):
        wt = w[f"{name}.weight"]
        return F.conv2d(x, wt, w[f"{name}.bias"], stride=stride, padding=wt.shape[-1] // 2, groups=groups)

    def se(name, x):
        s = x.mean((2, 3), keepdim=True)
        s = F.relu(conv(f"{name}.fc1", s))
        return x * torch.sigmoid(conv(f"{name}.fc2", s))

    def mlp(name, x):
        c = x.shape[1]
        h = conv(f"{name}.conv.conv", x, groups=c)
        h = F.gelu(conv(f"{name}.fc1", h))
        return conv(f"{name}.fc2", h)

    d0 = dims[0]
    x = F.gelu(conv(f"{pre}.stem.0.reparam_conv", x, stride=2))
    x = F.gelu(conv(f"{pre}.stem.1.reparam_conv", x, stride=2, groups=d0))
    x = F.gelu(conv(f"{pre}.stem.2.reparam_conv", x))
    prev = d0
    for i, (c, depth) in enumerate(zip(dims, depths)):
        st = f"{pre}.stages.{i}"
        if i > 0:
            x = conv(f"{st}.downsample.proj.0.reparam_conv", x, stride=2, groups=prev)
            if se_down[i]:
                x = se(f"{st}.downsample.proj.0.se", x)
            x = F.gelu(x)
            x = F.gelu(conv(f"{st}.downsample.proj.1.reparam_conv", x))
        last = i >= len(dims) - int(t.meta.get("attn_stages", 1))  # attention stages (fastvit_mci3 / mci4: the last two)
        if last:
            x = conv(f"{st}.pos_emb.reparam_conv", x, groups=c)  # RepCPE, identity branch folded in
        for j in range(depth):
            b = f"{st}.blocks.{j}"
            if not last:
                x = conv(f"{b}.token_mixer.reparam_conv", x, groups=c)  # RepMixer, re-parameterised
                x = x + w[f"{b}.layer_scale.gamma"] * mlp(f"{b}.mlp", x)
            else:
                B, C, H, W = x.shape
                h = F.batch_norm(x, w[f"{b}.norm.running_mean"], w[f"{b}.norm.running_var"], w[f"{b}.norm.weight"],
                                 w[f"{b}.norm.bias"], False, 0.0, 1e-5)
                tok = h.flatten(2).transpose(1, 2)  # [B, N, C]
                heads = C // 32
                a = _mha(tok, w[f"{b}.token_mixer.qkv.weight"], None, w[f"{b}.token_mixer.proj.weight"],
                         w[f"{b}.token_mixer.proj.bias"], heads, None)
                a = a.transpose(1, 2).reshape(B, C, H, W)
                x = x + w[f"{b}.layer_scale_1.gamma"] * a
                x = x + w[f"{b}.layer_scale_2.gamma"] * mlp(f"{b}.mlp", x)
        prev = c
    x = conv(f"{pre}.final_conv.reparam_conv", x, groups=prev)
    x = F.gelu(se(f"{pre}.final_conv.se", x))
    x = x.mean((2, 3))
    return F.linear(x, w[f"{pre}.head.fc.weight"], w[f"{pre}.head.fc.bias"])


@torch.no_grad()
def text_forward(t: Tower, input_ids: np.ndarray, normalize: bool = True) -> np.ndarray:
    """input_ids i64 [B,ctx] -> text_embeddings f32 [B,D], L2-normalised (pull_onnx.py:67-68)."""
    w = t.w
    ids = torch.from_numpy(np.ascontiguousarray(input_ids)).long()
    fam, D, L, H = t.meta["family"], t.i("width"), t.i("layers"), t.i("heads")
    act, eps, causal, pool = t.i("act"), float(t.meta["eps"]), bool(t.i("causal")), t.meta["pool"]
    pre = "model" if fam == "clip" else "model.text"
    B, T = ids.shape
    x = w[f"{pre}.token_embedding.weight"][ids] + w[f"{pre}.positional_embedding"][:T]
    mask = None
    if causal:
        mask = torch.full((T, T), float("-inf"), dtype=t.dtype).triu_(1)
    x = _clip_blocks(t, x, f"{pre}.transformer", L, H, act, eps, mask)
    x = F.layer_norm(x, (D,), w[f"{pre}.ln_final.weight"], w[f"{pre}.ln_final.bias"], eps)
    if pool == "argmax":
        pooled = x[torch.arange(B), ids.argmax(dim=-1)]
    elif pool == "last":
        pooled = x[:, -1]
    else:
        raise ValueError(pool)
    if f"{pre}.text_projection.weight" in w:
        out = F.linear(pooled, w[f"{pre}.text_projection.weight"], w[f"{pre}.text_projection.bias"])
    else:
        out = pooled @ w[f"{pre}.text_projection"]
    if normalize:
        out = F.normalize(out, dim=-1)
    return out.to(torch.float32).numpy()


# ----------------------------------------------------------------------------------------------------------
# similarity tail  (src/clip.rs)
# ----------------------------------------------------------------------------------------------------------
def _mul_add_f32(s: np.ndarray, scale: float, bias: float) -> np.ndarray:
    """f32::mul_add (fused): computed exactly in f64 (24x24-bit product is exact) and rounded once."""
    return (s.astype(np.float64) * np.float64(np.float32(scale)) + np.float64(np.float32(bias))).astype(np.float32)


def softmax(logits: np.ndarray) -> np.ndarray:
    """clip.rs:174-179."""
    logits = np.asarray(logits, dtype=np.float32)
    m = np.float32(-np.inf)
    for v in logits:
        m = max(m, v)
    exps = np.exp((logits - m).astype(np.float32)).astype(np.float32)
    total = np.float32(0.0)
    for e in exps:
        total = np.float32(total + e)
    return (exps / total).astype(np.float32)


def sigmoid(logit):
    """clip.rs:183-185."""
    x = np.asarray(logit, dtype=np.float32)
    return (np.float32(1.0) / (np.float32(1.0) + np.exp(-x).astype(np.float32))).astype(np.float32)


def probabilities(embs: np.ndarray, query: np.ndarray, model_config: dict) -> np.ndarray:
    """clip.rs:102-121 / :144-163: dot -> mul_add(scale, bias) -> sigmoid | softmax."""
    sims = (embs.astype(np.float32) @ query.astype(np.float32)).astype(np.float32)
    scale = model_config.get("logit_scale")
    bias = model_config.get("logit_bias")
    logits = _mul_add_f32(sims, 1.0 if scale is None else scale, 0.0 if bias is None else bias)
    if (model_config.get("activation_function") or "softmax") == "sigmoid":
        return sigmoid(logits)
    return softmax(logits)


def sort_desc(items: List[Tuple], key_index: int = 1) -> List[Tuple]:
    """clip.rs:129 / :167: stable sort, descending by probability, NaN compares Equal."""
    import functools

    def cmp(a, b):
        x, y = a[key_index], b[key_index]
        if math.isnan(x) or math.isnan(y):
            return 0
        return -1 if y < x else (1 if y > x else 0)

    return sorted(items, key=functools.cmp_to_key(cmp))


class OracleClip:
    """The whole reference pipeline on the CPU: Clip::{compare, classify, rank_images} (clip.rs:81-170)."""

    def __init__(self, model_dir: str, towers=("vision", "text"), threads: Optional[int] = None):
        if threads:
            torch.set_num_threads(threads)
        self.model_dir = model_dir
        with open(os.path.join(model_dir, "open_clip_config.json")) as f:
            self.config = json.load(f)
        with open(os.path.join(model_dir, "model_config.json")) as f:
            self.model_config = json.load(f)
        self.vision = Tower(os.path.join(model_dir, "visual.onnx")) if "vision" in towers else None
        self.text = Tower(os.path.join(model_dir, "text.onnx")) if "text" in towers else None

    def embed_images(self, images: Sequence[np.ndarray]) -> np.ndarray:
        pc = self.config["preprocess_cfg"]
        size = int(self.config["model_cfg"]["vision_cfg"]["image_size"])
        return vision_forward(self.vision, preprocess_batch(images, size, pc["mean"], pc["std"],
                                                            pc.get("interpolation", "bicubic"),
                                                            pc.get("resize_mode", "shortest")))

    def embed_texts(self, texts: Sequence[str]) -> np.ndarray:
        ids, _ = tokenize(self.model_dir, texts)
        return text_forward(self.text, ids)

    def classify(self, image: np.ndarray, labels: Sequence[str]) -> List[Tuple[str, float]]:
        v = self.embed_images([image])[0]
        t = self.embed_texts(labels)
        probs = probabilities(t, v, self.model_config)
        return sort_desc([(l, float(p)) for l, p in zip(labels, probs)])

    def rank_images(self, images: Sequence[np.ndarray], text: str) -> List[Tuple[int, float]]:
        v = self.embed_images(images)
        t = self.embed_texts([text])[0]
        probs = probabilities(v, t, self.model_config)
        return sort_desc([(i, float(p)) for i, p in enumerate(probs)])

    def compare(self, image: np.ndarray, text: str) -> float:
        v = self.embed_images([image])[0]
        t = self.embed_texts([text])[0]
        sim = np.float32(np.dot(v, t))
        return float(_mul_add_f32(np.asarray([sim]), self.model_config.get("logit_scale") or 1.0,
                                  self.model_config.get("logit_bias") or 0.0)[0])
