"""CPU-only checks of the host-side mirror and of the C-ABI library surface (no compute calls: there is no GPU in
the build container and the engine has no CPU path)."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "clipb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(clipb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from clip_embedder_rs_b200 import _native

    declared = _header_symbols()
    assert len(declared) >= 30
    exported = subprocess.run(["nm", "-D", "--defined-only", os.fspath(_native.LIB_PATH)], capture_output=True,
                              text=True, check=True).stdout
    exported = set(re.findall(r" T (clipb200_[a-z0-9_]+)", exported))
    assert set(declared) <= exported, sorted(set(declared) - exported)
    assert set(declared) == set(_native.SIGNATURES), sorted(set(declared) ^ set(_native.SIGNATURES))
    assert b"sm_100a" in _native.lib.clipb200_version()


def test_struct_layouts_match_header():
    from clip_embedder_rs_b200 import _native

    assert C.sizeof(_native.Opts) == 32 and C.sizeof(_native.Preproc) == 32
    assert C.sizeof(_native.Profile) == 8 * 8 + 8 * 8 + 8 + 8


def test_onnx_inspect_and_loader_errors(make_model, tmp_path):
    from clip_embedder_rs_b200 import _native

    import onnx_proto

    mdir = make_model("tiny_siglip")
    buf = C.create_string_buffer(1 << 16)
    for fn, inp, out in (("visual.onnx", "pixel_values", "image_embeddings"), ("text.onnx", "input_ids", "text_embeddings")):
        path = os.path.join(mdir, fn)
        assert _native.lib.clipb200_onnx_inspect(path.encode(), buf, len(buf)) == _native.OK, _native.last_error()
        info = json.loads(buf.value.decode())
        py = onnx_proto.read_model(path)
        assert info["inputs"] == [inp] and info["outputs"] == [out] and info["opset"] == 18
        assert info["num_initializers"] == len(py["initializers"])
        assert info["initializer_bytes"] == sum(a.nbytes for a in py["initializers"].values())
        assert info["metadata"] == py["metadata"]
    assert _native.lib.clipb200_onnx_inspect(os.path.join(mdir, "missing.onnx").encode(), buf, len(buf)) == _native.ERR_IO
    assert _native.lib.clipb200_onnx_inspect(os.path.join(mdir, "tokenizer.json").encode(), buf, len(buf)) == _native.ERR_PARSE
    # external data file missing -> IO error naming the file; truncated external data -> PARSE error
    broken = tmp_path / "broken"
    broken.mkdir()
    (broken / "visual.onnx").write_bytes(open(os.path.join(mdir, "visual.onnx"), "rb").read())
    assert _native.lib.clipb200_onnx_inspect(str(broken / "visual.onnx").encode(), buf, len(buf)) == _native.ERR_IO
    assert "visual.onnx.data" in _native.last_error()
    (broken / "visual.onnx.data").write_bytes(b"\0" * 100)
    assert _native.lib.clipb200_onnx_inspect(str(broken / "visual.onnx").encode(), buf, len(buf)) == _native.ERR_PARSE
    assert "outside" in _native.last_error()


def test_external_data_must_stay_inside_the_model_directory(tmp_path):
    """An untrusted .onnx must not be able to map files outside its own directory through an external-data `location`
    (absolute path, or `..` components): onnxruntime refuses those, so does the loader — before opening anything."""
    import onnx_proto as op
    from clip_embedder_rs_b200 import _native, error
    from clip_embedder_rs_b200.onnx import inspect_onnx

    secret = tmp_path / "secret.bin"
    secret.write_bytes(np.arange(64, dtype=np.float32).tobytes())
    mdir = tmp_path / "model"
    mdir.mkdir()
    (mdir / "inside.bin").write_bytes(np.arange(64, dtype=np.float32).tobytes())
    (mdir / "sub").mkdir()
    (mdir / "sub" / "deep.bin").write_bytes(np.arange(64, dtype=np.float32).tobytes())

    def model_with(location: str) -> str:
        g = op.f_str(2, "g") + op.f_bytes(5, op.tensor_proto("w", dims=(8, 8), data_type=op.FLOAT, external=(location, 0, 256)))
        g += op.f_bytes(11, op.value_info("pixel_values", op.FLOAT, ("b", 3, 8, 8)))
        m = op.f_varint(1, 8) + op.f_bytes(7, g) + op.f_bytes(8, op.f_str(1, "") + op.f_varint(2, 18))
        path = mdir / "visual.onnx"
        path.write_bytes(m)
        return str(path)

    for ok in ("inside.bin", "./inside.bin", "sub/deep.bin", "sub/../inside.bin"):
        assert inspect_onnx(model_with(ok))["num_initializers"] == 1, ok
    for bad in ("../secret.bin", str(secret), "sub/../../secret.bin", "..", "/etc/passwd", "..\\secret.bin"):
        with pytest.raises(error.Ort) as ei:
            inspect_onnx(model_with(bad))
        assert ei.value.code == _native.ERR_PARSE and "escapes the model directory" in str(ei.value), (bad, str(ei.value))


def test_onnx_writer_reader_roundtrip(tmp_path):
    import onnx_proto as op

    path = str(tmp_path / "m.onnx")
    w = op.ModelWriter(path, "g", external_threshold=64)
    big = np.arange(1000, dtype=np.float32).reshape(10, 100)
    small = np.asarray([1.5, -2.0], dtype=np.float32)
    ints = np.asarray([[1, -2, 3]], dtype=np.int64)
    w.add_input("x", op.FLOAT, ["batch_size", 3])
    w.add_output("y", op.FLOAT, ["batch_size", 4])
    w.add_initializer("big", big); w.add_initializer("small", small); w.add_initializer("ints", ints)
    w.add_node(op.node("Identity", ["small"], ["small_alias"]))
    w.add_metadata("k", "v")
    w.close()
    m = op.read_model(path)
    assert m["inputs"] == ["x"] and m["outputs"] == ["y"] and m["metadata"] == {"k": "v"}
    assert np.array_equal(m["initializers"]["big"], big) and np.array_equal(m["initializers"]["small"], small)
    assert np.array_equal(m["initializers"]["ints"], ints) and m["nodes"] == [("Identity", ["small"], ["small_alias"])]
    assert os.path.getsize(path + ".data") >= big.nbytes


def test_config_and_model_manager(make_model, tmp_path):
    import clip_embedder_rs_b200 as cb
    from clip_embedder_rs_b200 import error, model_manager

    mdir = make_model("tiny_clip")
    oc = cb.OpenClipConfig.from_file(os.path.join(mdir, "open_clip_config.json"))
    assert oc.model_cfg.embed_dim == 64 and oc.model_cfg.vision_cfg.image_size == 64
    assert oc.model_cfg.text_cfg.context_length == 77 and oc.preprocess_cfg.interpolation == "bicubic"
    assert oc.preprocess_cfg.resize_mode == "shortest" and len(oc.preprocess_cfg.mean) == 3
    mc = cb.ModelConfig.from_file(os.path.join(mdir, "model_config.json"))
    assert mc.logit_scale == 100.0 and mc.activation_function == "softmax" and mc.pad_id == 0
    (tmp_path / "mc.json").write_text("{}")
    d = cb.ModelConfig.from_file(tmp_path / "mc.json")  # serde defaults (config.rs:7-14)
    assert d.tokenizer_needs_lowercase is False and d.logit_scale is None and d.pad_id is None
    (tmp_path / "oc.json").write_text(json.dumps({"model_cfg": {"embed_dim": 8, "vision_cfg": {"image_size": 32},
                                                                "text_cfg": {"context_length": 7}},
                                                  "preprocess_cfg": {"mean": [0, 0, 0], "std": [1, 1, 1]}}))
    o2 = cb.OpenClipConfig.from_file(tmp_path / "oc.json")  # defaults config.rs:59-64
    assert o2.preprocess_cfg.interpolation == "bicubic" and o2.preprocess_cfg.resize_mode == "shortest"
    (tmp_path / "bad.json").write_text("{not json")
    with pytest.raises(error.Json):
        cb.ModelConfig.from_file(tmp_path / "bad.json")
    with pytest.raises(error.Io):
        cb.ModelConfig.from_file(tmp_path / "absent.json")
    assert len(model_manager.MODEL_FILES) == 9
    model_manager.verify_model_dir(mdir)
    with pytest.raises(error.ModelFolderNotFound):
        model_manager.verify_model_dir(tmp_path / "nope")
    part = tmp_path / "part"
    part.mkdir()
    (part / "model_config.json").write_text("{}")
    with pytest.raises(error.MissingModelFile) as ei:
        model_manager.verify_model_dir(part)
    assert ei.value.file == "open_clip_config.json"
    assert str(model_manager.get_default_base_folder()).endswith(".cache/open_clip_rs")
    with pytest.raises(error.HfHub):
        cb.Clip.from_hf("RuteNL/MobileCLIP2-S2-OpenCLIP-ONNX").build()


def test_no_gpu_means_loud_failure(make_model):
    """There is no CPU fallback: without a CUDA device the engine refuses to load (ClipError::Ort, code CUDA)."""
    import clip_embedder_rs_b200 as cb
    from clip_embedder_rs_b200 import _native, error

    if _native.lib.clipb200_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(error.Ort) as ei:
        cb.VisionEmbedder.from_local_dir(make_model("tiny_clip")).build()
    assert ei.value.code == _native.ERR_CUDA and "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing in the package or the C sources may reference it."""
    pkg = os.path.join(ROOT, "clip_embedder_rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|oracle[/.]reference_forward)", src), f


def test_shard_ranges_and_counter_corpus():
    from clip_embedder_rs_b200 import sharding

    for n in (0, 1, 7, 8, 100000, 1024):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    a = sharding.counter_images(10, 14, 16, seed=3)
    b = sharding.counter_images(12, 13, 16, seed=3)
    assert a.shape == (4, 16, 16, 3) and a.dtype == np.uint8 and np.array_equal(a[2], b[0])
    assert not np.array_equal(a[0], a[1]) and 100 < a.mean() < 155


_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch.distributed as dist
from clip_embedder_rs_b200 import sharding
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
N, D = 37, 8
def fake_embed(start, stop):   # stands in for VisionEmbedder.embed_images on this rank's GPU
    idx = np.arange(start, stop, dtype=np.float32)[:, None]
    return idx * 10 + np.arange(D, dtype=np.float32)[None, :]
start, rows = sharding.embed_sharded(fake_embed, N, rank, world)
full = sharding.gather_rows(rows, N, D, rank, world, dst=0)
if rank == 0:
    want = np.arange(N, dtype=np.float32)[:, None] * 10 + np.arange(D, dtype=np.float32)[None, :]
    assert full.shape == (N, D) and np.array_equal(full, want), "gathered shards out of order"
    print("GATHER_OK")
else:
    assert full is None
dist.barrier(); dist.destroy_process_group()
"""


def test_two_rank_sharding_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = 29500 + (os.getpid() % 500)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0]


def test_cpp_host_mirror_errors_and_configs(tmp_path):
    """include/clipb200.hpp without a GPU: `verify_model_dir` errors with the crate's messages (src/error.rs:29-37),
    `ModelConfig` / `OpenClipConfig::from_file` (serde defaults, optional fields, JSON / IO errors), softmax / sigmoid."""
    import subprocess

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "host_api_test.bin")
    if not os.path.exists(path):
        pytest.fail(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    out = subprocess.run([path, "errors", str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "HOST API ERRORS PASSED" in out.stdout, out.stdout + out.stderr


def test_committed_bench_lines_follow_the_contract():
    """The JSON lines committed under profiles/ (what `bench.py` printed on a B200) carry every key the driver reads:
    the base contract, `e2e`, `gpu_launches`, `clocks`, `roofline` (with `traffic`) and `cpu_baseline`; the reference arm
    carries `impl` and a zero-copy `e2e`."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    own = json.loads(open(os.path.join(root, "profiles", "r01_bench_final.json")).read())
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in own, k
    assert own["unit"] == "images/s" and own["higher_is_better"] is True and own["scaling"] == "weak"
    assert own["vs_baseline"] is None and own["dtype"] == "bf16" and own["data"] == "synthetic"
    assert "workload" in own["config"] and "model" not in own["config"]
    assert own["warmup"] >= 3 and own["gpu_launches"] > 0
    assert set(own["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert own["e2e"]["h2d_bytes_per_step"] == 1024 * 384 * 384 * 3 and own["e2e"]["d2h_bytes_per_step"] == 1024 * 1152 * 4
    assert abs(own["e2e"]["value"] - own["value"]) / own["value"] < 0.05
    r = own["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is not None and 0.5 < r["frac"] < 1.0
    assert set(own["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(own["clocks"]["reasons"])
    c = own["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    ref = json.loads(open(os.path.join(root, "profiles", "r01_bench_reference_arm.json")).read())
    assert ref["impl"] == "reference" and ref["metric"] == own["metric"] and ref["unit"] == own["unit"]
    assert ref["config"]["workload"] == own["config"]["workload"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0 and ref["e2e"]["value"] == ref["value"]
