"""Real ONNX graphs for the synthetic model directories: `torch.onnx.export`, the call `pull_onnx.py:169-181` makes.

`tools/export_synthetic.py` writes initializer-only files (enough for a loader that binds by parameter name).  This
tool produces what the reference's `ort::Session` actually consumes: executable graphs, exported from `nn.Module`s
that are laid out like the classes `open_clip.create_model_and_transforms` builds:

  * `ClipVisionTower` / `ClipTextTower`  — open_clip `VisionTransformer` / `CLIP.encode_text` / `TextTransformer`:
    `nn.MultiheadAttention` residual blocks, class token, `ln_pre` / `ln_post`, `x @ proj`, EOT-argmax or last pooling;
  * `TimmVisionTower`  — timm `VisionTransformer` trunk with fused-qkv `Attention` (scaled_dot_product_attention) and
    the `AttentionPoolLatent` MAP head;
  * `FastVitVisionTower`  — timm `FastVit` (MobileCLIP2's fastvit_mci2 / mci3 / mci4) as `reparameterize_model` leaves
    it: one `reparam_conv` per block branch set, ConvMlp, RepCPE + BatchNorm + attention blocks in the last stage(s).

They are wrapped exactly like `pull_onnx.py:53-68` (`self.model = ...`, `encode_image(x, normalize=True)`), exported
with `opset_version=18, do_constant_folding=True, dynamic_axes={name: {0: "batch_size"}}` and dummy batch 2
(`pull_onnx.py:41-42,169-181,279-302`).  open_clip / timm / onnxscript are not installed in this image, so the modules
are restated here and the TorchScript-based exporter (`dynamo=False`) is used; its output has what a name-bound loader
cannot handle: `Linear` weights renamed `onnx::MatMul_<n>` and stored transposed, `Identity`-deduplicated tensors,
head counts only inside Reshape/shape arithmetic.  The graph bytes are kept verbatim; initializers are moved to the
sibling `*.onnx.data` file the reference's directory check requires (`src/model_manager.rs:16-17`).  `anonymize=True`
additionally renames every initializer to `val_<n>` (what an optimiser pass may leave behind), so nothing but the
graph structure identifies a tensor.

The weights are the same seeded tensors `export_synthetic.py` generates, so the functional oracle
(`oracle/reference_forward.py`), the ONNX interpreter (`oracle/onnx_interp.py`) and the engine can all be compared on
identical parameters.
"""
from __future__ import annotations

import argparse
import math
import os
import shutil
import sys
import tempfile
import warnings
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import export_synthetic as ex  # noqa: E402
import onnx_proto as op  # noqa: E402


# ----------------------------------------------------------------------------- open_clip-shaped modules
class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


def _act_layer(act: str) -> nn.Module:
    if act == "quick_gelu":
        return QuickGELU()
    if act == "gelu_tanh":
        return nn.GELU(approximate="tanh")
    return nn.GELU()


class ResidualAttentionBlock(nn.Module):
    def __init__(self, width: int, heads: int, mlp: int, act: str, eps: float):
        super().__init__()
        self.ln_1 = nn.LayerNorm(width, eps=eps)
        self.attn = nn.MultiheadAttention(width, heads, batch_first=True)
        self.ln_2 = nn.LayerNorm(width, eps=eps)
        self.mlp = nn.Sequential()
        self.mlp.add_module("c_fc", nn.Linear(width, mlp))
        self.mlp.add_module("gelu", _act_layer(act))
        self.mlp.add_module("c_proj", nn.Linear(mlp, width))

    def forward(self, x, attn_mask: Optional[torch.Tensor] = None):
        h = self.ln_1(x)
        x = x + self.attn(h, h, h, need_weights=False, attn_mask=attn_mask)[0]
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width, layers, heads, mlp, act, eps):
        super().__init__()
        self.resblocks = nn.ModuleList([ResidualAttentionBlock(width, heads, mlp, act, eps) for _ in range(layers)])

    def forward(self, x, attn_mask=None):
        for r in self.resblocks:
            x = r(x, attn_mask)
        return x


class ClipVisionTower(nn.Module):
    def __init__(self, v: ex.VisionSpec, embed_dim: int):
        super().__init__()
        D, P = v.width, v.patch
        T = (v.image_size // P) ** 2
        self.conv1 = nn.Conv2d(3, D, P, P, bias=False)
        self.class_embedding = nn.Parameter(torch.zeros(D))
        self.positional_embedding = nn.Parameter(torch.zeros(T + 1, D))
        self.ln_pre = nn.LayerNorm(D, eps=v.eps)
        self.transformer = Transformer(D, v.layers, v.heads, v.mlp_dim, v.act, v.eps)
        self.ln_post = nn.LayerNorm(D, eps=v.eps)
        self.proj = nn.Parameter(torch.zeros(D, embed_dim))

    def forward(self, x):
        x = self.conv1(x)
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
        cls = self.class_embedding.view(1, 1, -1).expand(x.shape[0], -1, -1)
        x = torch.cat([cls.to(x.dtype), x], dim=1)
        x = x + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = self.transformer(x)
        x = self.ln_post(x)
        pooled = x[:, 0]
        return pooled @ self.proj


class TextTower(nn.Module):
    """open_clip CLIP text side (parameters live directly on the CLIP model) or `TextTransformer` (CustomTextCLIP:
    parameters under `.text`)."""

    def __init__(self, t: ex.TextSpec, embed_dim: int):
        super().__init__()
        _init_text(self, t, embed_dim)

    def forward(self, text):
        return _encode_text(self, text)


def _init_text(m: nn.Module, t: ex.TextSpec, embed_dim: int) -> None:
    D = t.width
    m.token_embedding = nn.Embedding(t.vocab_size, D)
    m.positional_embedding = nn.Parameter(torch.zeros(t.context_length, D))
    m.transformer = Transformer(D, t.layers, t.heads, t.mlp_dim, t.act, t.eps)
    m.ln_final = nn.LayerNorm(D, eps=t.eps)
    if t.proj_bias:
        m.text_projection = nn.Linear(D, embed_dim)
    else:
        m.text_projection = nn.Parameter(torch.zeros(D, embed_dim))
    m.text_pool_type = t.pool
    if t.causal:
        mask = torch.empty(t.context_length, t.context_length).fill_(float("-inf")).triu_(1)
        m.register_buffer("attn_mask", mask, persistent=False)
    else:
        m.attn_mask = None


def _encode_text(m: nn.Module, text):
    x = m.token_embedding(text)
    x = x + m.positional_embedding
    x = m.transformer(x, attn_mask=m.attn_mask)
    x = m.ln_final(x)
    if m.text_pool_type == "argmax":
        x = x[torch.arange(x.shape[0]), text.argmax(dim=-1)]
    else:
        x = x[:, -1]
    if isinstance(m.text_projection, nn.Linear):
        return m.text_projection(x)
    return x @ m.text_projection


# ----------------------------------------------------------------------------- timm-shaped modules
class TimmMlp(nn.Module):
    def __init__(self, width, hidden, act):
        super().__init__()
        self.fc1 = nn.Linear(width, hidden)
        self.act = _act_layer(act)
        self.fc2 = nn.Linear(hidden, width)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class TimmAttention(nn.Module):
    def __init__(self, width, heads):
        super().__init__()
        self.num_heads = heads
        self.head_dim = width // heads
        self.qkv = nn.Linear(width, 3 * width)
        self.proj = nn.Linear(width, width)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        x = F.scaled_dot_product_attention(q, k, v)
        x = x.transpose(1, 2).reshape(B, N, C)
        return self.proj(x)


class TimmBlock(nn.Module):
    def __init__(self, width, heads, mlp, act, eps):
        super().__init__()
        self.norm1 = nn.LayerNorm(width, eps=eps)
        self.attn = TimmAttention(width, heads)
        self.norm2 = nn.LayerNorm(width, eps=eps)
        self.mlp = TimmMlp(width, mlp, act)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class PatchEmbed(nn.Module):
    def __init__(self, width, patch):
        super().__init__()
        self.proj = nn.Conv2d(3, width, patch, patch)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class AttentionPoolLatent(nn.Module):
    def __init__(self, width, heads, mlp, act, eps):
        super().__init__()
        self.num_heads = heads
        self.head_dim = width // heads
        self.latent = nn.Parameter(torch.zeros(1, 1, width))
        self.q = nn.Linear(width, width)
        self.kv = nn.Linear(width, 2 * width)
        self.proj = nn.Linear(width, width)
        self.norm = nn.LayerNorm(width, eps=eps)
        self.mlp = TimmMlp(width, mlp, act)

    def forward(self, x):
        B, N, C = x.shape
        q_latent = self.latent.expand(B, -1, -1)
        q = self.q(q_latent).reshape(B, 1, self.num_heads, self.head_dim).transpose(1, 2)
        kv = self.kv(x).reshape(B, N, 2, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        k, v = kv.unbind(0)
        x = F.scaled_dot_product_attention(q, k, v)
        x = x.transpose(1, 2).reshape(B, 1, C)
        x = self.proj(x)
        x = x + self.mlp(self.norm(x))
        return x[:, 0]


class TimmTrunk(nn.Module):
    def __init__(self, v: ex.VisionSpec):
        super().__init__()
        D = v.width
        T = (v.image_size // v.patch) ** 2
        self.patch_embed = PatchEmbed(D, v.patch)
        self.pos_embed = nn.Parameter(torch.zeros(1, T, D))
        self.blocks = nn.Sequential(*[TimmBlock(D, v.heads, v.mlp_dim, v.act, v.eps) for _ in range(v.layers)])
        self.norm = nn.LayerNorm(D, eps=v.eps)
        self.attn_pool = AttentionPoolLatent(D, v.heads, v.mlp_dim, v.act, v.eps)

    def forward(self, x):
        x = self.patch_embed(x)
        x = x + self.pos_embed
        x = self.blocks(x)
        x = self.norm(x)
        return self.attn_pool(x)


class TimmVisionTower(nn.Module):
    """open_clip `TimmModel`: `.trunk` + `.head` (Identity for `timm_proj: none`)."""

    def __init__(self, v: ex.VisionSpec):
        super().__init__()
        self.trunk = TimmTrunk(v)

    def forward(self, x):
        return self.trunk(x)


# ----------------------------------------------------------------------------- timm FastViT (MobileCLIP2), re-parameterised
class _SqueezeExcite(nn.Module):  # timm SqueezeExcite with 1x1 convs: fc1 -> ReLU -> fc2 -> sigmoid gate
    def __init__(self, ch: int, rd: int):
        super().__init__()
        self.fc1 = nn.Conv2d(ch, rd, 1)
        self.fc2 = nn.Conv2d(rd, ch, 1)

    def forward(self, x):
        s = x.mean((2, 3), keepdim=True)
        return x * torch.sigmoid(self.fc2(F.relu(self.fc1(s))))


class _RepConv(nn.Module):
    """timm `MobileOneBlock` / `ReparamLargeKernelConv` / `RepCPE` / `RepMixer` after `reparameterize_model`
    (pull_onnx.py:110-116): every branch and BatchNorm folded into one `reparam_conv`, then the optional SE and GELU."""

    def __init__(self, cin, cout, k, stride=1, groups=1, se_rd: int = 0, act: bool = True):
        super().__init__()
        self.reparam_conv = nn.Conv2d(cin, cout, k, stride, k // 2, groups=groups)
        if se_rd:
            self.se = _SqueezeExcite(cout, se_rd)
        self.has_se, self.has_act = bool(se_rd), act

    def forward(self, x):
        x = self.reparam_conv(x)
        if self.has_se:
            x = self.se(x)
        return F.gelu(x) if self.has_act else x


class _ConvNorm(nn.Module):  # timm ConvNormAct of ConvMlp with the BatchNorm folded: `.conv`
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 7, 1, 3, groups=c)

    def forward(self, x):
        return self.conv(x)


class _ConvMlp(nn.Module):
    def __init__(self, c, hidden):
        super().__init__()
        self.conv = _ConvNorm(c)
        self.fc1 = nn.Conv2d(c, hidden, 1)
        self.fc2 = nn.Conv2d(hidden, c, 1)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(self.conv(x))))


class _LayerScale2d(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(c, 1, 1))

    def forward(self, x):
        return x * self.gamma


class _RepMixerBlock(nn.Module):
    def __init__(self, c, hidden):
        super().__init__()
        self.token_mixer = _RepConv(c, c, 3, groups=c, act=False)
        self.mlp = _ConvMlp(c, hidden)
        self.layer_scale = _LayerScale2d(c)

    def forward(self, x):
        x = self.token_mixer(x)
        return x + self.layer_scale(self.mlp(x))


class _FastVitAttention(nn.Module):  # timm fastvit.Attention: tokens = flattened pixels, head dim 32, no qkv bias
    def __init__(self, c):
        super().__init__()
        self.heads = c // 32
        self.qkv = nn.Linear(c, 3 * c, bias=False)
        self.proj = nn.Linear(c, c)

    def forward(self, x):
        B, C, H, W = x.shape
        t = x.flatten(2).transpose(-2, -1)
        qkv = self.qkv(t).reshape(B, H * W, 3, self.heads, 32).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, H * W, C)
        return self.proj(o).transpose(-2, -1).reshape(B, C, H, W)


class _AttentionBlock(nn.Module):
    def __init__(self, c, hidden):
        super().__init__()
        self.norm = nn.BatchNorm2d(c)
        self.token_mixer = _FastVitAttention(c)
        self.layer_scale_1 = _LayerScale2d(c)
        self.mlp = _ConvMlp(c, hidden)
        self.layer_scale_2 = _LayerScale2d(c)

    def forward(self, x):
        x = x + self.layer_scale_1(self.token_mixer(self.norm(x)))
        return x + self.layer_scale_2(self.mlp(x))


class _PatchEmbed(nn.Module):  # stage downsampling: 7x7 s2 depthwise (+ SE) -> GELU -> 1x1 -> GELU
    def __init__(self, cin, cout, se_rd):
        super().__init__()
        self.proj = nn.Sequential(_RepConv(cin, cout, 7, 2, groups=cin, se_rd=se_rd), _RepConv(cout, cout, 1))

    def forward(self, x):
        return self.proj(x)


class _FastVitStage(nn.Module):
    def __init__(self, cin, c, depth, hidden, down_se_rd, attention: bool):
        super().__init__()
        if cin != c:
            self.downsample = _PatchEmbed(cin, c, down_se_rd)
        if attention:
            self.pos_emb = _RepConv(c, c, 7, groups=c, act=False)   # RepCPE, identity branch folded in
        self.has_down, self.has_cpe = cin != c, attention
        self.blocks = nn.Sequential(*[(_AttentionBlock if attention else _RepMixerBlock)(c, hidden) for _ in range(depth)])

    def forward(self, x):
        if self.has_down:
            x = self.downsample(x)
        if self.has_cpe:
            x = self.pos_emb(x)
        return self.blocks(x)


class _FastVitHead(nn.Module):  # timm ClassifierHead: global average pool + fc
    def __init__(self, cin, embed):
        super().__init__()
        self.fc = nn.Linear(cin, embed)

    def forward(self, x):
        return self.fc(x.mean((2, 3)))


class FastVitTrunk(nn.Module):
    """timm `FastVit` (fastvit_mci2 / mci3 / mci4) in its re-parameterised inference form; parameter names as timm's."""

    def __init__(self, v: ex.VisionSpec, embed_dim: int, shapes: Dict[str, tuple]):
        super().__init__()
        pre = "model.visual.trunk"
        d0 = v.dims[0]
        self.stem = nn.Sequential(_RepConv(3, d0, 3, 2), _RepConv(d0, d0, 3, 2, groups=d0), _RepConv(d0, d0, 1))
        stages, prev = [], d0
        n = len(v.dims)
        for i, (c, depth) in enumerate(zip(v.dims, v.depths)):
            se_key = f"{pre}.stages.{i}.downsample.proj.0.se.fc1.weight"
            se_rd = shapes[se_key][0] if se_key in shapes else 0
            stages.append(_FastVitStage(prev, c, depth, v.mlp_ratio * c, se_rd, attention=i >= n - v.attn_stages))
            prev = c
        self.stages = nn.Sequential(*stages)
        self.final_conv = _RepConv(prev, 2 * prev, 3, groups=prev, se_rd=shapes[f"{pre}.final_conv.se.fc1.weight"][0])
        self.head = _FastVitHead(2 * prev, embed_dim)

    def forward(self, x):
        return self.head(self.final_conv(self.stages(self.stem(x))))


class FastVitVisionTower(nn.Module):
    def __init__(self, v: ex.VisionSpec, embed_dim: int, shapes):
        super().__init__()
        self.trunk = FastVitTrunk(v, embed_dim, shapes)

    def forward(self, x):
        return self.trunk(x)


# ----------------------------------------------------------------------------- CLIP containers + export wrappers
class ClipModel(nn.Module):
    """`open_clip.CLIP`: `.visual`, text parameters on the model itself."""

    def __init__(self, spec: ex.ModelSpec, towers):
        super().__init__()
        if "vision" in towers:
            self.visual = ClipVisionTower(spec.vision, spec.embed_dim)
        if "text" in towers:
            _init_text(self, spec.text, spec.embed_dim)

    def encode_image(self, x, normalize: bool = False):
        f = self.visual(x)
        return F.normalize(f, dim=-1) if normalize else f

    def encode_text(self, text, normalize: bool = False):
        f = _encode_text(self, text)
        return F.normalize(f, dim=-1) if normalize else f


class CustomTextClipModel(nn.Module):
    """`open_clip.CustomTextCLIP`: `.visual` (TimmModel) and `.text` (TextTransformer)."""

    def __init__(self, spec: ex.ModelSpec, towers, shapes=None):
        super().__init__()
        if "vision" in towers:
            self.visual = (FastVitVisionTower(spec.vision, spec.embed_dim, shapes) if spec.vision.family == "fastvit"
                           else TimmVisionTower(spec.vision))
        if "text" in towers:
            self.text = TextTower(spec.text, spec.embed_dim)

    def encode_image(self, x, normalize: bool = False):
        f = self.visual(x)
        return F.normalize(f, dim=-1) if normalize else f

    def encode_text(self, text, normalize: bool = False):
        f = self.text(text)
        return F.normalize(f, dim=-1) if normalize else f


class VisualWrapper(nn.Module):  # pull_onnx.py:53-59
    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, x):
        return self.model.encode_image(x, normalize=True)


class TextWrapper(nn.Module):  # pull_onnx.py:62-68
    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, x):
        return self.model.encode_text(x, normalize=True)


def build_model(spec: ex.ModelSpec, seed: int = 0, towers=("vision", "text")) -> nn.Module:
    clip_style = spec.text.family == "clip"
    weights: Dict[str, np.ndarray] = {}
    if "vision" in towers:
        ex.gen_vision(spec, seed, lambda n, a: weights.__setitem__(n, a))
    if "text" in towers:
        ex.gen_text(spec, seed, lambda n, a: weights.__setitem__(n, a))
    shapes = {k: tuple(np.shape(v)) for k, v in weights.items()}
    model = ClipModel(spec, towers) if clip_style else CustomTextClipModel(spec, towers, shapes)
    sd = {k[len("model."):]: torch.from_numpy(np.array(v)) for k, v in weights.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    missing = [k for k in missing if not k.endswith("num_batches_tracked")]   # BatchNorm bookkeeping, not a weight
    assert not missing, missing
    return model.eval()


# ----------------------------------------------------------------------------- export + repack
def _patch_exporter() -> None:
    # the TorchScript exporter only needs the `onnx` package to splice onnxscript functions, which we never use
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils as U

    U._add_onnxscript_fn = lambda model_bytes, custom_opsets: model_bytes


def _rename_in_node(node_bytes: memoryview, rename: Dict[str, str]) -> bytes:
    out = bytearray()
    for f, w, val in op._fields(node_bytes):
        if f in (1, 2) and w == 2:
            s = bytes(val).decode()
            out += op.f_str(f, rename.get(s, s))
        elif w == 2:
            out += op.f_bytes(f, bytes(val))
        elif w == 0:
            out += op.f_varint(f, val)
        else:
            out += op._key(f, w) + val
    return bytes(out)


def repack(src_path: str, dst_path: str, anonymize: bool = False, metadata: Optional[Dict[str, str]] = None) -> None:
    """Copies the exported ModelProto, keeping every graph node byte-for-byte, and moves initializers >= 1 KiB into
    `<dst>.data` (external data, 64-byte aligned).  `anonymize` renames all initializers to `val_<n>`."""
    src = op.read_model(src_path)
    rename: Dict[str, str] = {}
    if anonymize:
        for i, name in enumerate(src["initializers"]):
            rename[name] = f"val_{i}"
    with open(src_path, "rb") as f:
        buf = memoryview(f.read())
    w = op.ModelWriter(dst_path, "main_graph", producer="pytorch")
    graph_rest = bytearray()
    top_rest = bytearray()
    for f_, w_, val in op._fields(buf):
        if f_ == 7:
            for gf, gw, gval in op._fields(val):
                if gf == 5:
                    continue  # initializers are re-emitted below
                if gf == 1:
                    w.add_node(_rename_in_node(gval, rename) if rename else bytes(gval))
                elif gf in (11, 12):
                    (w._inputs if gf == 11 else w._outputs).append(bytes(gval))
                elif gw == 2 and gf != 2:
                    graph_rest += op.f_bytes(gf, bytes(gval))
        elif f_ == 8:
            for of, _, oval in op._fields(val):
                if of == 2:
                    w.opset = int(oval)
    # graph inputs listed for initializers (older IR) are not emitted by torch >= 1.x, nothing to filter
    for name, arr in src["initializers"].items():
        w.add_initializer(rename.get(name, name), np.ascontiguousarray(arr))
    for k, v in (metadata or {}).items():
        w.add_metadata(k, v)
    w.close()


def export_tower(module: nn.Module, dummy: torch.Tensor, path: str, in_name: str, out_name: str,
                 anonymize: bool = False, dynamic_batch: bool = True, external_data: bool = True) -> None:
    """`dynamic_batch=False` bakes the dummy batch into the graph (what `torch.onnx.export` does without
    `dynamic_axes`); `external_data=False` keeps the exporter's file as it is, weights inline — the form OpenCV's DNN
    importer (an independent ONNX runtime, used as a cross-check in the tests) can read."""
    _patch_exporter()
    module.eval()   # a fresh wrapper is in training mode; the exporter restores that mode afterwards, recursively, which
                    # would leave the wrapped model's BatchNorm layers (FastViT) using batch statistics in eager runs
    kw = dict(dynamic_axes={in_name: {0: "batch_size"}, out_name: {0: "batch_size"}}) if dynamic_batch else {}
    with tempfile.TemporaryDirectory() as tmp, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        raw = os.path.join(tmp, "raw.onnx")
        torch.onnx.export(module, dummy, raw, input_names=[in_name], output_names=[out_name],
                          opset_version=18, do_constant_folding=True, dynamo=False, **kw)
        if external_data or anonymize:
            repack(raw, path, anonymize=anonymize)
        else:
            shutil.copyfile(raw, path)


def export_model_dir(spec: ex.ModelSpec, out_dir: str, seed: int = 0, towers=("vision", "text"),
                     anonymize: bool = False) -> str:
    """Same nine-file directory as `export_synthetic.write_model_dir`, but `visual.onnx` / `text.onnx` hold the graphs
    `torch.onnx.export` produced (no `clipb200.*` metadata: everything must come from the graph)."""
    ex.write_model_dir(spec, out_dir, seed=seed, towers=())  # configs + tokenizer; tower files are replaced below
    model = build_model(spec, seed, towers)
    if "vision" in towers:
        s = spec.vision.image_size
        export_tower(VisualWrapper(model), torch.randn(2, 3, s, s), os.path.join(out_dir, "visual.onnx"),
                     "pixel_values", "image_embeddings", anonymize)
    if "text" in towers:
        ids = torch.randint(0, spec.text.vocab_size, (2, spec.text.context_length))
        export_tower(TextWrapper(model), ids, os.path.join(out_dir, "text.onnx"), "input_ids", "text_embeddings",
                     anonymize)
    return out_dir


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--config", required=True, choices=sorted(ex.CONFIGS))
    ap.add_argument("--out", required=True)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--towers", default="vision,text")
    ap.add_argument("--anonymize", action="store_true")
    a = ap.parse_args()
    export_model_dir(ex.CONFIGS[a.config], a.out, a.seed, tuple(a.towers.split(",")), a.anonymize)
    print(a.out)


if __name__ == "__main__":
    main()
