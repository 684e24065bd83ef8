import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with `-m gpu` on the GPU box")


@pytest.fixture(scope="session")
def model_root(tmp_path_factory):
    return tmp_path_factory.mktemp("models")


@pytest.fixture(scope="session")
def make_model(model_root):
    """Writes (once per session) a synthetic model directory in the reference's layout and returns its path."""
    import export_synthetic as ex

    cache = {}

    def _make(config: str, seed: int = 0) -> str:
        key = (config, seed)
        if key not in cache:
            cache[key] = ex.write_model_dir(ex.CONFIGS[config], os.path.join(model_root, f"{config}_s{seed}"), seed)
        return cache[key]

    return _make


@pytest.fixture(scope="session")
def make_real_model(model_root):
    """Model directory whose visual.onnx / text.onnx are executable graphs written by `torch.onnx.export`
    (tools/torch_export.py), optionally with every initializer renamed to `val_<n>`."""
    import export_synthetic as ex
    import torch_export as te

    cache = {}

    def _make(config: str, seed: int = 0, anonymize: bool = False, towers=("vision", "text")) -> str:
        key = (config, seed, anonymize, tuple(towers))
        if key not in cache:
            out = os.path.join(model_root, f"real_{config}_s{seed}_{'anon' if anonymize else 'named'}_{'-'.join(towers)}")
            cache[key] = te.export_model_dir(ex.CONFIGS[config], out, seed, tuple(towers), anonymize)
        return cache[key]

    return _make


def random_images(n: int, size: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 256, size=(n, size, size, 3), dtype=np.uint8)


def random_texts(n: int, seed: int):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        words = []
        for _ in range(int(rng.integers(3, 9))):
            words.append("".join(chr(ord("a") + int(c)) for c in rng.integers(0, 26, size=int(rng.integers(2, 10)))))
        out.append(" ".join(words))
    return out


def cosine_rows(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return (a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1))


REF_PHOTO_DIRS = ("/root/reference/assets/img", os.path.join(ROOT, "tests", "_ref_assets"))


def reference_photos():
    """The reference's own example photos (`/root/reference/assets/img/*.jpg`, what its README and integration test
    embed).  They are not part of this repository: `__graft_entry__.build()` copies them into the git-ignored
    `tests/_ref_assets/` when the reference checkout is present, which is how they reach the GPU box.  Returns
    [(name, uint8 [H,W,3])], decoded with Pillow like `image::open(..).to_rgb8()` (vision.rs:171)."""
    from PIL import Image

    for d in REF_PHOTO_DIRS:
        if os.path.isdir(d):
            names = sorted(f for f in os.listdir(d) if f.lower().endswith((".jpg", ".jpeg", ".png")))
            if names:
                return [(n, np.asarray(Image.open(os.path.join(d, n)).convert("RGB"))) for n in names]
    return []
