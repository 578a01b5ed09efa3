import json
import lzma
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


with open(os.path.join(GOLDEN, "manifest.json")) as _f:
    MANIFEST = json.load(_f)


def golden_names():
    return sorted(MANIFEST)


def load_golden_frame(name: str) -> np.ndarray:
    """The frame the compiled reference rendered (uint32 XRGB8888, (H, W))."""
    info = MANIFEST[name]
    w, h = info["width"], info["height"]
    with open(os.path.join(GOLDEN, name + ".frame.xz"), "rb") as f:
        planar = np.frombuffer(lzma.decompress(f.read()), dtype=np.uint8).reshape(3, h, w).astype(np.uint32)
    return (planar[0] << 16) | (planar[1] << 8) | planar[2]


def load_golden_scene(name: str):
    from gp1_raytracer_2223_b200.scene_file import load_rtsc
    return load_rtsc(os.path.join(GOLDEN, name + ".rtsc"))


def compare_frames(got: np.ndarray, want: np.ndarray):
    """(fraction of pixels identical in every channel, max per-channel |diff|, differing pixel count)."""
    assert got.shape == want.shape, (got.shape, want.shape)
    diff_px = got != want
    n_diff = int(diff_px.sum())
    if n_diff == 0:
        return 1.0, 0, 0
    g = got.view(np.uint8).reshape(got.shape + (4,)).astype(np.int16)
    w = want.view(np.uint8).reshape(want.shape + (4,)).astype(np.int16)
    max_err = int(np.abs(g - w).max())
    return 1.0 - n_diff / got.size, max_err, n_diff


# SURVEY.md / BASELINE.json parity bar
MIN_IDENTICAL = 0.999
MAX_LSB = 1
