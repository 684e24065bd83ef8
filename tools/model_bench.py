"""The reference's own benchmark (`/root/reference/benches/model_bench.rs:28-47`: per model `vision/preprocess`,
`vision/embed` = preprocess + inference of ONE photo, `text/embed` = tokenize + inference of ONE string) through the
host mirror, for the models listed there (`:7-14`), with random-init weights of the same architectures.  The photo is a
synthetic 1920x1280 RGB image, so `vision/embed` includes the GPU resize.  README.md:106-113 of the reference holds the
author's CPU numbers for the same three measurements (ms, "vision embedding includes 10-20 ms preprocessing").

    python tools/model_bench.py [--models so400m_siglip2_384,mobileclip2_s2,...]
"""
import argparse
import json
import os
import statistics
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import clip_embedder_rs_b200 as cb  # noqa: E402
import export_synthetic as ex  # noqa: E402

# config -> (name in benches/model_bench.rs, the reference README's CPU ms for vision / text embedding)
MODELS = {
    "so400m_siglip2_384": ("timm/ViT-SO400M-16-SigLIP2-384", 988, 136),
    "dfn5b_h14_378": ("apple/DFN5B-CLIP-ViT-H-14-378", 1860, 131),
    "mobileclip2_s2": ("timm/MobileCLIP2-S2-OpenCLIP", 75, 19),
    "mobileclip2_s3": ("timm/MobileCLIP2-S3-OpenCLIP", 116, 35),
    "mobileclip2_s4": ("timm/MobileCLIP2-S4-OpenCLIP", 192, 38),
    "gopt_siglip2_384": ("timm/ViT-gopt-16-SigLIP2-384", 2354, 128),
}


def med_ms(fn, n=20, warm=4):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", default=",".join(MODELS))
    ap.add_argument("--root", default=os.path.join(tempfile.gettempdir(), "clipb200_model_bench"))
    args = ap.parse_args()
    photo = np.random.default_rng(11).integers(0, 256, size=(1280, 1920, 3), dtype=np.uint8)
    text = "A photo of rocks"
    for cfg in args.models.split(","):
        name, ref_v, ref_t = MODELS[cfg]
        mdir = ex.write_model_dir(ex.CONFIGS[cfg], os.path.join(args.root, cfg), seed=0)
        clip = cb.Clip.from_local_dir(mdir).micro_batch(8).build()
        row = {"model": name, "config": cfg,
               "vision/preprocess_ms": round(med_ms(lambda: clip.vision.preprocess(photo)), 3),
               "vision/embed_ms": round(med_ms(lambda: clip.vision.embed_image(photo)), 3),
               "text/embed_ms": round(med_ms(lambda: clip.text.embed_text(text)), 3),
               "reference_readme_cpu_ms": {"vision": ref_v, "text": ref_t}}
        print(json.dumps(row), flush=True)
        del clip


if __name__ == "__main__":
    main()
