"""CPU restatements of arithmetic that lives inside the CUDA kernels, with the constants read from the kernel sources,
so that a changed coefficient cannot silently degrade the numerics the GPU parity tests only see through a cosine."""
import os
import re

import numpy as np

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "clip_embedder_rs_b200", "csrc")


def _exp2_poly_constants():
    src = open(os.path.join(CSRC, "attn_sm100.cuh")).read()
    body = src[src.index("void exp2_poly_pair("):]
    body = body[:body.index("\n}\n")]
    clamp = float(re.search(r"fmaxf\(x0, (-?[0-9.]+)f\)", body).group(1))
    magic = float(re.search(r"add2\(x, pack2\(([0-9.]+)f,", body).group(1))
    c3, c2 = re.search(r"fma2\(pack2\(([0-9.]+)f, [0-9.]+f\), fr, pack2\(([0-9.]+)f,", body).groups()
    c1 = re.search(r"p = fma2\(p, fr, pack2\(([0-9.]+)f, [0-9.]+f\)\);\n  p = fma2", body).group(1)
    c0 = re.findall(r"p = fma2\(p, fr, pack2\(([0-9.]+)f,", body)[-1]
    return clamp, magic, [np.float32(c) for c in (c0, c1, c2, c3)]


def _exp2_poly(x, clamp, magic, c):
    """attn_sm100.cuh exp2_poly_pair in numpy float32: Cody-Waite split by the 1.5 * 2^23 magic add, degree-3 polynomial
    for 2^f, integer add of round(x) into the exponent field."""
    x = np.maximum(x.astype(np.float32), np.float32(clamp))
    xf = (x + np.float32(magic)).astype(np.float32)
    xi = (xf - np.float32(magic)).astype(np.float32)
    fr = (x - xi).astype(np.float32)
    p = (c[3] * fr + c[2]).astype(np.float32)
    p = (p * fr + c[1]).astype(np.float32)
    p = (p * fr + c[0]).astype(np.float32)
    bits = p.view(np.int32) + (xf.view(np.int32) << 23)     # int32 wrap-around is the kernel's behaviour too
    return bits.astype(np.int32).view(np.float32)


def test_attention_exp2_polynomial_accuracy_and_range():
    clamp, magic, c = _exp2_poly_constants()
    assert magic == 12582912.0 and clamp == -125.0
    rng = np.random.default_rng(0)
    # the kernel's arguments are s * scale - m <= 8 (lazy rescaling keeps P <= 2^8) and arbitrarily negative
    x = np.concatenate([rng.uniform(-125.0, 9.0, 400000), np.linspace(-126.0, 9.0, 100001), np.arange(-125, 10, 0.5)])
    got = _exp2_poly(x.astype(np.float32), clamp, magic, c).astype(np.float64)
    want = np.exp2(np.maximum(x.astype(np.float32), np.float32(clamp)).astype(np.float64))
    rel = np.abs(got / want - 1.0)
    assert rel.max() < 1.0e-4, rel.max()                    # bf16 rounding of P is 3.9e-3
    # masked scores (-inf) and anything below the clamp come out as 2^-125: positive, finite, invisible next to 2^0
    tiny = _exp2_poly(np.array([-np.inf, -1.0e30, -126.0, -125.0], np.float32), clamp, magic, c)
    assert np.all(np.isfinite(tiny)) and np.all(tiny > 0) and np.all(tiny < 1e-37)
    # exact at integers up to the polynomial's constant-term error
    ints = np.arange(-120, 9, dtype=np.float32)
    assert np.abs(_exp2_poly(ints, clamp, magic, c).astype(np.float64) / np.exp2(ints.astype(np.float64)) - 1).max() < 1e-4


def test_fast_erf_gelu_matches_exact_gelu():
    """gemm_sm100.cuh gelu_erf_fast (Abramowitz-Stegun 7.1.26 form used by the GEMM / conv epilogues for ACT_GELU_ERF)
    restated in float32 with the constants read from the source, against x * Phi(x) from scipy's erf in float64."""
    from scipy.special import erf

    src = open(os.path.join(CSRC, "gemm_sm100.cuh")).read()
    body = src[src.index("float gelu_erf_fast(float x) {"):]
    body = body[:body.index("\n}\n")]
    p = np.float32(re.search(r"fmaf\(([0-9.]+)f, fabsf\(x\), 1\.0f\)", body).group(1))
    k = np.float32(re.search(r"x \* x \* (-[0-9.]+)f", body).group(1))
    a5, a4 = (np.float32(v) for v in re.search(r"q = fmaf\(t, ([0-9.]+)f, (-[0-9.]+)f\);", body).groups())
    rest = [np.float32(v) for v in re.findall(r"q = fmaf\(t, q, (-?[0-9.]+)f\);", body)]
    assert len(rest) == 3
    a3, a2, a1 = rest
    assert abs(float(k) + 0.5 / np.log(2.0)) < 1e-7          # exp(-x^2/2) as exp2(x^2 * -0.5 / ln 2)

    x = np.concatenate([np.linspace(-12, 12, 480001), np.random.default_rng(0).normal(0, 2, 200000)]).astype(np.float32)
    t = (np.float32(1) / (p * np.abs(x) + np.float32(1))).astype(np.float32)
    e = np.exp2((x * x * k).astype(np.float32)).astype(np.float32)
    q = (t * a5 + a4).astype(np.float32)
    for a in (a3, a2, a1):
        q = (t * q + a).astype(np.float32)
    h = (q * t * e).astype(np.float32)
    got = (x * np.where(x < 0, h, np.float32(1) - h)).astype(np.float64)
    xd = x.astype(np.float64)
    want = 0.5 * xd * (1.0 + erf(xd / np.sqrt(2.0)))
    # 4.3e-7 is the formula's own bound; float32 evaluation and the approx MUFU ops add a few ulps
    assert np.abs(got - want).max() < 2.0e-6, np.abs(got - want).max()


def test_fused_mlp_relu_form_gelu_matches_exact_gelu():
    """fused_mlp_sm100.cuh gelu_erf_pair_bf16: gelu(x) = relu(x) - |x| * Phi(-|x|) in packed f32x2 arithmetic (no sign
    select), same Abramowitz-Stegun constants as gelu_erf_fast.  Restated in float32 with the constants read from the
    source; the result is rounded to bf16 by the kernel, so 2e-6 absolute is three orders below what H can resolve."""
    from scipy.special import erf

    src = open(os.path.join(CSRC, "fused_mlp_sm100.cuh")).read()
    body = src[src.index("uint32_t gelu_erf_pair_bf16("):]
    body = body[:body.index("\n}\n")]
    body = body[body.index("#endif"):]   # skip the timing-only experiment branch
    p = np.float32(re.search(r"f2_fma\(nax, f2_splat\((-[0-9.]+)f\), f2_splat\(1\.0f\)\)", body).group(1))
    k = np.float32(re.search(r"f2_mul\(f2_mul\(nax, nax\), f2_splat\((-[0-9.]+)f\)\)", body).group(1))
    a5, a4 = (np.float32(v) for v in re.search(r"q = f2_fma\(t, f2_splat\(([0-9.]+)f\), f2_splat\((-[0-9.]+)f\)\);", body).groups())
    rest = [np.float32(v) for v in re.findall(r"q = f2_fma\(t, q, f2_splat\((-?[0-9.]+)f\)\);", body)]
    assert len(rest) == 3 and p < 0
    assert "f2_fma(nax, h, f2_pack(fmaxf(xa, 0.f), fmaxf(xb, 0.f)))" in body      # relu(x) + (-|x|) * h
    a3, a2, a1 = rest

    x = np.concatenate([np.linspace(-12, 12, 480001), np.random.default_rng(1).normal(0, 2, 200000)]).astype(np.float32)
    nax = -np.abs(x)
    t = (np.float32(1) / (nax * p + np.float32(1))).astype(np.float32)
    e = np.exp2(((nax * nax).astype(np.float32) * k).astype(np.float32)).astype(np.float32)
    q = (t * a5 + a4).astype(np.float32)
    for a in (a3, a2, a1):
        q = (t * q + a).astype(np.float32)
    h = ((q * t).astype(np.float32) * e).astype(np.float32)
    got = (nax * h + np.maximum(x, np.float32(0))).astype(np.float64)
    xd = x.astype(np.float64)
    want = 0.5 * xd * (1.0 + erf(xd / np.sqrt(2.0)))
    assert np.abs(got - want).max() < 2.0e-6, np.abs(got - want).max()
