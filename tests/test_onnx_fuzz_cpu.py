"""Robustness of the file-facing part of the C ABI: `clipb200_engine_create` / `clipb200_onnx_inspect` parse
user-supplied ONNX files (csrc/onnx_loader.cc) and interpret their graphs (csrc/onnx_graph.cc).  Byte-mutated copies of
real `torch.onnx.export` graphs (bit flips, truncation, splices, swapped blocks, varint tweaks) must be declined or
accepted — never crash, overflow or read out of bounds.  Checked twice: through the shipped library, and through an
AddressSanitizer + UBSan build of the same two source files."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "clip_embedder_rs_b200", "csrc")


def _mutations(data: bytes, n: int, seed: int):
    rng = np.random.default_rng(seed)
    for it in range(n):
        b = bytearray(data)
        mode = it % 5
        if mode == 0:
            for _ in range(int(rng.integers(1, 8))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        elif mode == 1:
            b = b[: int(rng.integers(1, len(b)))]
        elif mode == 2:
            i, k = int(rng.integers(0, len(b) - 64)), int(rng.integers(1, 64))
            b[i:i + k] = bytes(rng.integers(0, 256, size=k, dtype=np.uint8))
        elif mode == 3:
            i, j = int(rng.integers(0, len(b) - 8)), int(rng.integers(0, len(b) - 8))
            b[i:i + 8], b[j:j + 8] = b[j:j + 8], b[i:i + 8]
        else:
            for _ in range(int(rng.integers(1, 4))):
                i = int(rng.integers(0, len(b)))
                b[i] = [0, 1, 0xFF, (b[i] + 1) & 0xFF, (b[i] - 1) & 0xFF, 0x7F, 0x80][int(rng.integers(0, 7))]
        yield bytes(b)


def _write_corpus(src_dir: str, fname: str, out_dir, n: int, seed: int):
    os.makedirs(out_dir, exist_ok=True)
    os.symlink(os.path.join(src_dir, fname + ".data"), os.path.join(out_dir, fname + ".data"))
    data = open(os.path.join(src_dir, fname), "rb").read()
    paths = []
    for i, blob in enumerate(_mutations(data, n, seed)):
        # every variant keeps the name of the external-data file it references, so it lives in its own directory
        d = os.path.join(out_dir, f"m{i}")
        os.makedirs(d)
        os.symlink(os.path.join(src_dir, fname + ".data"), os.path.join(d, fname + ".data"))
        p = os.path.join(d, fname)
        with open(p, "wb") as f:
            f.write(blob)
        paths.append(p)
    return paths


@pytest.mark.parametrize("config,fname,seed", [("tiny_clip", "visual.onnx", 1), ("tiny_siglip", "text.onnx", 2),
                                               ("tiny_mobileclip", "visual.onnx", 7)])
def test_mutated_graphs_never_crash_the_library(make_real_model, tmp_path, config, fname, seed):
    if config == "tiny_mobileclip":   # FastViT conv graph: the name + graph-edge binder (bind_fastvit_graph)
        mdir = make_real_model(config, towers=("vision",))
    else:
        mdir = make_real_model(config, anonymize=(config == "tiny_siglip"))
    paths = _write_corpus(mdir, fname, str(tmp_path / "corpus"), 160, seed)
    script = (
        "import sys, ctypes as C\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from clip_embedder_rs_b200 import _native\n"
        "buf = C.create_string_buffer(1 << 22)\n"
        "codes = {}\n"
        "for p in open(sys.argv[1]).read().split('\\n'):\n"
        "    if p:\n"
        "        rc = _native.lib.clipb200_onnx_inspect(p.encode(), buf, len(buf))\n"
        "        codes[rc] = codes.get(rc, 0) + 1\n"
        "print('INSPECT DONE', sorted(codes.items()))\n")
    listing = tmp_path / "files.txt"
    listing.write_text("\n".join(paths))
    out = subprocess.run([sys.executable, "-c", script, str(listing)], capture_output=True, text=True, errors="replace",
                         timeout=600)
    assert out.returncode == 0 and "INSPECT DONE" in out.stdout, (out.returncode, out.stdout[-500:], out.stderr[-1500:])


def test_mutated_graphs_under_address_and_ub_sanitizers(make_real_model, make_model, tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "onnx_fuzz_harness")
    build = subprocess.run([gxx, "-std=c++17", "-g", "-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                            "-fno-omit-frame-pointer", "-I", CSRC, os.path.join(ROOT, "tests", "native", "onnx_fuzz_harness.cc"),
                            os.path.join(CSRC, "onnx_loader.cc"), os.path.join(CSRC, "onnx_graph.cc"), "-o", exe],
                           capture_output=True, text=True, timeout=600)
    if build.returncode != 0 and "sanitize" in build.stderr:
        pytest.skip("sanitizer runtime not available: " + build.stderr[-300:])
    assert build.returncode == 0, build.stderr[-2000:]
    paths = []
    for config, fname, seed in (("tiny_clip", "visual.onnx", 3), ("tiny_clip", "text.onnx", 4), ("tiny_siglip", "visual.onnx", 5)):
        mdir = make_real_model(config, anonymize=(config == "tiny_siglip"))
        paths += _write_corpus(mdir, fname, str(tmp_path / f"corpus_{config}_{fname}"), 120, seed)
        paths.append(os.path.join(mdir, fname))  # and the unmodified file: must be recognised
    # a FastViT conv graph: declined by the recogniser, its attention Linears bound through graph edges
    fv = make_real_model("tiny_mobileclip", towers=("vision",))
    paths += _write_corpus(fv, "visual.onnx", str(tmp_path / "corpus_fastvit"), 120, 8)
    paths.append(os.path.join(fv, "visual.onnx"))
    # initializer-only files (tools/export_synthetic.py): the loader alone (typed data fields, external-data records)
    paths += _write_corpus(make_model("tiny_clip"), "visual.onnx", str(tmp_path / "corpus_synthetic"), 120, 6)
    for i in range(0, len(paths), 64):
        out = subprocess.run([exe] + paths[i:i + 64], capture_output=True, text=True, errors="replace", timeout=600)
        assert out.returncode == 0 and "FUZZ HARNESS DONE" in out.stdout, (paths[i:i + 64][:2], out.stderr[-3000:])
    out = subprocess.run([exe] + [p for p in paths if "/corpus_" not in p], capture_output=True, text=True, timeout=600)
    assert "loaded=4 recognised=3 fastvit=1" in out.stdout, out.stdout
