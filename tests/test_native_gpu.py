"""Runs the standalone native checks (built by `make gemm_test attn_test`) as part of the GPU suite:
tests/native/gemm_test.cu — tcgen05 GEMM (1-CTA and CTA-pair modes, all epilogues, M/N/K tails) against a SIMT fp32
reference; tests/native/attn_test.cu — tcgen05 attention against the mma.sync kernel and an fp32 SIMT reference."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
NATIVE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native")


@pytest.mark.parametrize("binary,arg,marker", [("gemm_test.bin", "1", "GEMM TEST PASSED"), ("attn_test.bin", None, "ATTN TEST PASSED")])
def test_native_binary(binary, arg, marker):
    path = os.path.join(NATIVE, binary)
    if not os.path.exists(path):
        pytest.fail(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    cmd = [path] + ([arg] if arg else [])
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    print(out.stdout[-3000:])
    assert out.returncode == 0 and marker in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_mma_latency_microbenchmark_runs():
    """tests/native/mma_latency_test.cu (per-instruction tcgen05.mma cost, commit and mbarrier hand-off latencies quoted in
    DESIGN.md section 3.2) runs to completion: a protocol error in it would trap or hang, not print the end marker."""
    path = os.path.join(NATIVE, "mma_latency_test.bin")
    if not os.path.exists(path):
        pytest.fail(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    out = subprocess.run([path], capture_output=True, text=True, timeout=120)
    print(out.stdout[-3000:])
    assert out.returncode == 0 and "MMA LATENCY TEST DONE" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("switch", ["CLIPB200_ATTN_SINGLE_S", "CLIPB200_ATTN_DOUBLE_S"])
def test_attention_protocol_switches(switch):
    """The tcgen05 attention kernel has two MMA <-> softmax hand-off protocols (one S tile, or two S tiles with P written
    in place) and picks one per head dim by measurement; the environment switches force the other one, and every
    correctness case (head dims 64 / 72 / 80 / 96, kv tails, causal, multi-tile, persistent loop) must pass with it."""
    path = os.path.join(NATIVE, "attn_test.bin")
    if not os.path.exists(path):
        pytest.fail(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    env = dict(os.environ, **{switch: "1"})
    for case in range(11):   # cases 0..10 are the correctness cases, 11+ the timed ones
        out = subprocess.run([path, str(case)], capture_output=True, text=True, timeout=120, env=env)
        assert out.returncode == 0 and "ATTN TEST PASSED" in out.stdout, (case, out.stdout[-2000:] + out.stderr[-2000:])


def test_short_sequence_attention_kernel():
    """`attn_short_sm100.cuh` (T <= 80, head dim 64: the CLIP text towers) against the mma.sync kernel and the fp32
    reference: ViT-B/32 vision-like T = 50, causal T = 77, one token, 1 / 2 / 3 active softmax warps, T = 80, and more
    items than resident CTAs.  `ATTN_NO_VT=1` selects the natural V layout, where the engine's `attn_auto` dispatch picks
    the short kernel; `CLIPB200_ATTN_SHORT=0` must give the long kernel the same cases."""
    path = os.path.join(NATIVE, "attn_test.bin")
    if not os.path.exists(path):
        pytest.fail(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    for extra in ({}, {"CLIPB200_ATTN_SHORT": "0"}):
        env = dict(os.environ, ATTN_NO_VT="1", **extra)
        for case in (6, 8, 19, 20, 21, 22, 23, 24):
            out = subprocess.run([path, str(case)], capture_output=True, text=True, timeout=120, env=env)
            assert out.returncode == 0 and "ATTN TEST PASSED" in out.stdout, (extra, case, out.stdout[-2000:] + out.stderr[-2000:])


def test_loader_accepts_f16_bf16_and_identity_aliases(make_model, tmp_path):
    """The ONNX loader's claims beyond the synthetic exporter's defaults: fp16 / bf16 initializers, inline
    `float_data`, and exporter-style de-duplication (an `Identity` node aliasing one initializer under a second name)."""
    import numpy as np

    import clip_embedder_rs_b200 as cb
    import onnx_proto as op
    from conftest import cosine_rows, random_images

    src = make_model("tiny_clip")
    ref = cb.VisionEmbedder.from_local_dir(src).build()
    imgs = random_images(3, 64, seed=9)
    want = ref.embed_images(imgs)

    m = op.read_model(os.path.join(src, "visual.onnx"))
    dst = tmp_path / "variant"
    dst.mkdir()
    for f in os.listdir(src):
        if not f.startswith("visual.onnx"):
            os.symlink(os.path.join(src, f), dst / f)
    w = op.ModelWriter(str(dst / "visual.onnx"), "visual")
    w.add_input("pixel_values", op.FLOAT, ["batch_size", 3, 64, 64])
    w.add_output("image_embeddings", op.FLOAT, ["batch_size", 64])
    for k, v in m["metadata"].items():
        w.add_metadata(k, v)
    alias_src = "model.visual.transformer.resblocks.0.ln_1.bias"
    alias_dst = "model.visual.transformer.resblocks.1.ln_1.bias"
    for i, (name, arr) in enumerate(m["initializers"].items()):
        arr = np.array(arr)
        if name == alias_dst:
            continue  # provided through the Identity node below (made equal to resblocks.0's bias)
        if name.endswith("mlp.c_fc.weight"):  # fp16 storage
            w._inits.append(op.tensor_proto(name, arr.astype(np.float16)))
        elif name.endswith("attn.out_proj.weight"):  # bf16 storage (upper 16 bits of fp32, round-to-nearest-even)
            u = arr.astype(np.float32).view(np.uint32)
            bf = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
            w._inits.append(op.f_bytes(1, b"".join(op._varint(int(d)) for d in arr.shape)) + op.f_varint(2, op.BFLOAT16) +
                            op.f_str(8, name) + op.f_bytes(9, bf.tobytes()))
        else:
            w.add_initializer(name, arr)
    w.add_node(op.node("Identity", [alias_src], [alias_dst]))
    w.close()
    got = cb.VisionEmbedder.from_local_dir(dst).build().embed_images(imgs)
    # same model up to fp16/bf16 storage rounding of two weight families and one aliased LayerNorm bias
    assert cosine_rows(got, want).min() > 0.995
    assert np.abs(got - want).max() > 0  # the variant really differs (alias + rounding), i.e. it was loaded


def test_cpp_host_mirror_matches_python_mirror(make_model, tmp_path):
    """include/clipb200.hpp (C++ mirror of the crate API: Clip / VisionEmbedder / TextEmbedder, configs, errors) driven
    by tests/native/host_api_test.cpp on the same model directory and inputs as the Python mirror: same embeddings,
    same classify / rank_images order, same preprocess values, `Empty batch` error, working `duplicate()`."""
    import json

    import numpy as np

    import clip_embedder_rs_b200 as cb
    from conftest import random_images, random_texts

    path = os.path.join(NATIVE, "host_api_test.bin")
    if not os.path.exists(path):
        pytest.fail(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    for config in ("tiny_clip", "tiny_siglip"):
        mdir = make_model(config)
        clip = cb.Clip.from_local_dir(mdir).build()
        n, w, h = 4, 100, 80  # not the model resolution: the GPU resize path is part of the call
        imgs = np.random.default_rng(61).integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
        texts = random_texts(5, seed=62)
        ids, mask = clip.text.tokenize(texts)
        lens = mask.sum(axis=1).astype(np.int64)
        imgs.tofile(tmp_path / "images.u8")
        ids.astype(np.int64).tofile(tmp_path / "ids.i64")
        lens.tofile(tmp_path / "lens.i64")
        out = subprocess.run([path, "run", mdir, str(tmp_path / "images.u8"), str(n), str(w), str(h), str(tmp_path / "ids.i64"),
                              str(tmp_path / "lens.i64"), str(len(texts))], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        j = json.loads(out.stdout)
        D = clip.vision.session.embed_dim
        assert j["embed_dim"] == D and j["input_name"] == "pixel_values" and j["id_name"] == "input_ids"
        assert j["empty_batch_error"] and j["duplicate_matches"]
        want_v = clip.vision.embed_images(list(imgs))
        want_t = clip.text.embed_texts(texts)
        got_v = np.asarray(j["image_embeddings"], dtype=np.float32).reshape(n, D)
        got_t = np.asarray(j["text_embeddings"], dtype=np.float32).reshape(len(texts), D)
        assert np.allclose(got_v, want_v, atol=1e-6) and np.allclose(got_t, want_t, atol=1e-6), config
        assert np.allclose(np.asarray(j["preprocess0"], dtype=np.float32), clip.vision.preprocess(imgs[0]).reshape(-1)[:64])
        want_cls = clip.classify(imgs[0], texts)
        assert [int(l[1:]) for l, _ in j["classify"]] == [texts.index(l) for l, _ in want_cls]
        assert np.allclose([p for _, p in j["classify"]], [p for _, p in want_cls], atol=1e-5)
        want_rank = clip.rank_images(list(imgs), texts[0])
        assert [i for i, _ in j["rank_images"]] == [i for i, _ in want_rank]
        assert abs(j["compare"] - clip.compare(imgs[0], texts[1])) < 1e-4
