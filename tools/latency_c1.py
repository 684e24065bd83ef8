"""BASELINE config 1 latency: `Clip::classify`, 1 image + 3 labels, ViT-B/32 (random-init weights), through the host
mirror.  Prints the engine's median latency and the CPU oracle's, plus batch-1 embed latencies for the big towers."""
import os
import statistics
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import clip_embedder_rs_b200 as cb  # noqa: E402
import export_synthetic as ex  # noqa: E402
from conftest import random_texts  # noqa: E402


def med(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts)


def main():
    base = os.path.join(tempfile.gettempdir(), "clipb200_models")
    mdir = ex.write_model_dir(ex.CONFIGS["vit_b32"], os.path.join(base, "vit_b32_lat"), seed=0)
    clip = cb.Clip.from_local_dir(mdir).micro_batch(8).build()
    img = np.random.default_rng(1).integers(0, 256, size=(224, 224, 3), dtype=np.uint8)
    labels = random_texts(3, seed=2)
    print(f"C1 classify (1 image + 3 labels) engine: {med(lambda: clip.classify(img, labels)):.3f} ms median")
    print(f"   vision embed_image: {med(lambda: clip.vision.embed_image(img)):.3f} ms, "
          f"text embed_texts(3): {med(lambda: clip.text.embed_texts(labels)):.3f} ms, "
          f"tokenize(3): {med(lambda: clip.text.tokenize(labels)):.3f} ms")
    if "--cpu" in sys.argv:
        from oracle import reference_forward as R

        o = R.OracleClip(mdir)
        print(f"C1 classify CPU oracle ({os.cpu_count()} threads): {med(lambda: o.classify(img, labels), n=5, warm=1):.1f} ms")
    for cfg, size in (("so400m_siglip2_384", 384),):
        d = ex.write_model_dir(ex.CONFIGS[cfg], os.path.join(base, cfg + "_lat"), seed=0, towers=("vision",))
        v = cb.VisionEmbedder.from_local_dir(d).micro_batch(8).build()
        im = np.random.default_rng(2).integers(0, 256, size=(size, size, 3), dtype=np.uint8)
        print(f"{cfg} batch-1 vision embed: {med(lambda: v.embed_image(im), n=20):.3f} ms (reference README: 988 ms on the author's CPU)")


if __name__ == "__main__":
    main()
