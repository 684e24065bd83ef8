#!/usr/bin/env python
"""Expected values for `rust/open_clip_inference_b200/tests/b200_parity.rs`, computed by the CPU oracle.

    python tools/rust_expectations.py <model_dir> > expect.json

The Rust test builds the same deterministic image (pixel (x, y, c) = (7x + 13y + 101c) % 256, 640 x 480) and the same
three labels, so no file has to travel with it.  Test infrastructure only (imports oracle/)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_image(w: int, h: int) -> np.ndarray:
    y, x, c = np.meshgrid(np.arange(h), np.arange(w), np.arange(3), indexing="ij")
    return ((7 * x + 13 * y + 101 * c) % 256).astype(np.uint8)


def main() -> None:
    from oracle import reference_forward as R

    mdir = sys.argv[1]
    o = R.OracleClip(mdir)
    img = test_image(640, 480)
    labels = ["a photo of a cat", "a photo of a dog", "a photo of a beignet"]
    out = {"classify": [[l, float(p)] for l, p in o.classify(img, labels)],
           "image_embedding": [float(v) for v in o.embed_images([img])[0]]}
    json.dump(out, sys.stdout)


if __name__ == "__main__":
    main()
