"""Generates tests/golden/*.npz: seeded inputs and the CPU oracle's outputs for the small configs.

The reference holds no golden vectors for this path (its only test needs network weights), and the
reference itself cannot run here, so these fixtures pin (a) the synthetic exporter (weights are a pure function of
config + seed; a checksum is stored), (b) the oracle's numerics at the commit that generated them, and (c) the
host tokenizer.  Regenerate with:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import export_synthetic as ex  # noqa: E402
from conftest import random_images, random_texts  # noqa: E402
from oracle import reference_forward as R  # noqa: E402

CONFIGS = ["tiny_clip", "tiny_clip_p14", "tiny_siglip"]


def file_sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    with tempfile.TemporaryDirectory() as tmp:
        for config in CONFIGS:
            mdir = ex.write_model_dir(ex.CONFIGS[config], os.path.join(tmp, config), seed=0)
            o = R.OracleClip(mdir, threads=1)
            size = int(o.config["model_cfg"]["vision_cfg"]["image_size"])
            imgs = random_images(3, size, seed=101)
            texts = random_texts(4, seed=102) + ["a photo of a cat"]
            ids, mask = R.tokenize(mdir, texts)
            pc = o.config["preprocess_cfg"]
            pv = R.preprocess_batch(list(imgs), size, pc["mean"], pc["std"])
            np.savez_compressed(
                os.path.join(out_dir, f"{config}.npz"),
                image_seed=101, text_seed=102, texts=np.asarray(texts),
                ids=ids, mask=mask,
                pixel_checksum=np.asarray([float(pv.astype(np.float64).sum()), float(np.abs(pv).astype(np.float64).sum())]),
                pixel_first=pv[0, :, 0, :8].copy(),
                image_embeddings=o.embed_images(list(imgs)), text_embeddings=o.embed_texts(texts),
                classify_probs=np.asarray([p for _, p in o.classify(imgs[0], texts[:3])], dtype=np.float32),
                classify_order=np.asarray([l for l, _ in o.classify(imgs[0], texts[:3])]),
                visual_data_sha256=file_sha(os.path.join(mdir, "visual.onnx.data")),
                text_data_sha256=file_sha(os.path.join(mdir, "text.onnx.data")))
            print("wrote", config)


if __name__ == "__main__":
    main()
