"""Random-init weights for a re-parameterised FastViT trunk (MobileCLIP2, `timm` names) with *calibrated* folded
BatchNorm.

A real export (`/root/reference/pull_onnx.py:110-116`: `reparameterize_model`, eval mode) has every BatchNorm folded
into the preceding convolution, which keeps each conv output roughly zero-mean / unit-variance per channel.  Plain
random weights without that property collapse: after ~50 GELU layers and a global average pool every image maps to
(almost) the same embedding and a parity test could not tell a correct engine from a broken one.  So this generator
runs a small calibration batch through the trunk while it draws the weights and, for every conv that carries a
BatchNorm in the trained model, rescales the weight and sets the bias from the batch statistics (plus noise, so the
normalisation is imperfect like real running statistics).  The calibration forward is the exporter's own (this is
the exporter's job, like `pull_onnx.py` instantiating the model); the CPU oracle has an independent restatement.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _calibration_images(n: int, size: int, seed: int) -> torch.Tensor:
    rng = np.random.default_rng(seed)
    block = max(size // 8, 4)
    g = size // block
    base = rng.integers(0, 256, size=(n, g, g, 3)).astype(np.float32)
    img = np.repeat(np.repeat(base, block, axis=1), block, axis=2)
    img = np.clip(img + rng.normal(0, 20, size=img.shape), 0, 255).astype(np.float32) / 255.0
    return torch.from_numpy(img).permute(0, 3, 1, 2).contiguous()


@torch.no_grad()
def gen_fastvit(spec, g, emit) -> None:
    v = spec.vision
    pre = "model.visual.trunk"
    rng = g.rng
    torch.set_num_threads(max(torch.get_num_threads(), 4))
    x = _calibration_images(6, v.image_size, seed=12345)
    x = (x - torch.tensor(spec.mean).view(1, 3, 1, 1)) / torch.tensor(spec.std).view(1, 3, 1, 1)

    def draw(cout, cin_g, k, gain=1.0):
        fan_in = cin_g * k * k
        return torch.from_numpy(g.normal((cout, cin_g, k, k), gain * fan_in ** -0.5))

    def conv(name, inp, cout, cin_g, k, stride=1, groups=1, bn=True, gain=1.0):
        """Draws a conv, optionally folds a calibrated BatchNorm into it, emits it, returns its output."""
        w = draw(cout, cin_g, k, gain)
        b = torch.from_numpy(g.small(cout, 0.02))
        y = F.conv2d(inp, w, None, stride=stride, padding=k // 2, groups=groups)
        if bn:
            mean = y.mean((0, 2, 3))
            std = y.std((0, 2, 3)) + 1e-3
            gamma = torch.from_numpy((1.0 + 0.1 * rng.standard_normal(cout)).astype(np.float32))
            beta = torch.from_numpy((0.1 * rng.standard_normal(cout)).astype(np.float32))
            scale = gamma / std
            w = w * scale.view(-1, 1, 1, 1)
            b = beta - mean * scale
            y = y * scale.view(1, -1, 1, 1)
        y = y + b.view(1, -1, 1, 1)
        emit(f"{name}.weight", w.numpy().astype(np.float32))
        emit(f"{name}.bias", b.numpy().astype(np.float32))
        return y

    def se(name, inp, ch):
        rd = max(ch // 16, 8)
        s = inp.mean((2, 3), keepdim=True)
        s = F.relu(conv(f"{name}.fc1", s, rd, ch, 1, bn=False, gain=1.4))
        s = torch.sigmoid(conv(f"{name}.fc2", s, ch, rd, 1, bn=False, gain=2.0))
        return inp * s

    def mlp(name, inp, c):
        h = conv(f"{name}.conv.conv", inp, c, 1, 7, groups=c, bn=True)
        h = F.gelu(conv(f"{name}.fc1", h, v.mlp_ratio * c, c, 1, bn=False, gain=1.4))
        return conv(f"{name}.fc2", h, c, v.mlp_ratio * c, 1, bn=False)

    def gamma(name, c):
        # layer-scale ~0.1: with 0.3 the random network amplifies perturbations so much that even the fp32 oracle
        # with bf16-rounded GEMM operands only reaches cosine 0.989 against itself; at 0.1 it is well conditioned
        # (0.99996) while different images still map to clearly different embeddings (cosine ~0.9).
        ls = getattr(v, "layer_scale", 0.1)  # the deeper 5-stage trunks (ratio-4 MLPs, two attention stages) need 0.05
        gm = (ls + 0.2 * ls * rng.standard_normal((c, 1, 1))).astype(np.float32)
        emit(name, gm)
        return torch.from_numpy(gm)

    d0 = v.dims[0]
    x = F.gelu(conv(f"{pre}.stem.0.reparam_conv", x, d0, 3, 3, stride=2))
    x = F.gelu(conv(f"{pre}.stem.1.reparam_conv", x, d0, 1, 3, stride=2, groups=d0))
    x = F.gelu(conv(f"{pre}.stem.2.reparam_conv", x, d0, d0, 1))
    prev = d0
    for i, (c, depth) in enumerate(zip(v.dims, v.depths)):
        st = f"{pre}.stages.{i}"
        if i > 0:
            x = conv(f"{st}.downsample.proj.0.reparam_conv", x, c, 1, 7, stride=2, groups=prev)
            if v.se_down[i]:
                x = se(f"{st}.downsample.proj.0.se", x, c)
            x = F.gelu(x)
            x = F.gelu(conv(f"{st}.downsample.proj.1.reparam_conv", x, c, c, 1))
        last = i >= len(v.dims) - getattr(v, "attn_stages", 1)  # attention stage(s): MCi2 the last one, MCi3 / MCi4 the last two
        if last:
            # RepCPE: x + dwconv7x7(x), folded into one conv (identity added to the centre tap)
            w = draw(c, 1, 7, 0.5)
            w[:, 0, 3, 3] += 1.0
            b = torch.from_numpy(g.small(c, 0.02))
            emit(f"{st}.pos_emb.reparam_conv.weight", w.numpy().astype(np.float32))
            emit(f"{st}.pos_emb.reparam_conv.bias", b.numpy().astype(np.float32))
            x = F.conv2d(x, w, b, padding=3, groups=c)
        for j in range(depth):
            blk = f"{st}.blocks.{j}"
            if not last:
                x = conv(f"{blk}.token_mixer.reparam_conv", x, c, 1, 3, groups=c, bn=True)
                m = mlp(f"{blk}.mlp", x, c)
                x = x + gamma(f"{blk}.layer_scale.gamma", c) * m
            else:
                B, C, H, W = x.shape
                mean = x.mean((0, 2, 3)) + 0.05 * torch.randn(C, generator=torch.Generator().manual_seed(j))
                var = x.var((0, 2, 3)) * torch.from_numpy((1.0 + 0.1 * rng.random(C)).astype(np.float32))
                nw = torch.from_numpy(g.ln_w(C))
                nb = torch.from_numpy(g.small(C))
                emit(f"{blk}.norm.weight", nw.numpy()); emit(f"{blk}.norm.bias", nb.numpy())
                emit(f"{blk}.norm.running_mean", mean.numpy().astype(np.float32))
                emit(f"{blk}.norm.running_var", var.numpy().astype(np.float32))
                h = F.batch_norm(x, mean, var, nw, nb, False, 0.0, 1e-5)
                wqkv = torch.from_numpy(g.normal((3 * C, C), C ** -0.5))
                wproj = torch.from_numpy(g.normal((C, C), C ** -0.5))
                bproj = torch.from_numpy(g.small(C))
                emit(f"{blk}.token_mixer.qkv.weight", wqkv.numpy())
                emit(f"{blk}.token_mixer.proj.weight", wproj.numpy())
                emit(f"{blk}.token_mixer.proj.bias", bproj.numpy())
                tok = h.flatten(2).transpose(1, 2)
                heads = C // 32
                qkv = F.linear(tok, wqkv).reshape(B, H * W, 3, heads, 32).permute(2, 0, 3, 1, 4)
                att = torch.softmax((qkv[0] * 32 ** -0.5) @ qkv[1].transpose(-1, -2), -1) @ qkv[2]
                a = F.linear(att.transpose(1, 2).reshape(B, H * W, C), wproj, bproj)
                x = x + gamma(f"{blk}.layer_scale_1.gamma", C) * a.transpose(1, 2).reshape(B, C, H, W)
                m = mlp(f"{blk}.mlp", x, C)
                x = x + gamma(f"{blk}.layer_scale_2.gamma", C) * m
        prev = c
    cf = 2 * prev
    x = conv(f"{pre}.final_conv.reparam_conv", x, cf, 1, 3, groups=prev)
    x = F.gelu(se(f"{pre}.final_conv.se", x, cf))
    emit(f"{pre}.head.fc.weight", g.normal((spec.embed_dim, cf), cf ** -0.5))
    emit(f"{pre}.head.fc.bias", g.small(spec.embed_dim))
