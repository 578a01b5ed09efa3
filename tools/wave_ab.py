#!/usr/bin/env python
"""Kernel variants side by side (tiled 1, persistent 3, wavefront 4, auto 0) on fixtures: kernel ms + diff vs golden."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import MANIFEST, load_golden_frame, load_golden_scene
from gp1_raytracer_2223_b200 import Renderer
names = sys.argv[1:] or ["optional_320", "optional_640", "optional_4k", "bunny_640", "bunny_4k", "w4ref_640"]
for name in names:
    info = MANIFEST[name]
    for variant, label in ((1, "tiled"), (3, "persistent"), (4, "wavefront"), (0, "auto")):
        r = Renderer(info["width"], info["height"])
        r.SetScene(load_golden_scene(name))
        r.ctx.set_kernel_variant(variant)
        for _ in range(5):
            r.render_device()
        ms = [r.render_device()["kernel_ms"] for _ in range(30)]
        diff = int((r.download() != load_golden_frame(name)).sum())
        print(f"{name:16s} {label:10s} kernel ms mean {np.mean(ms):.4f} min {np.min(ms):.4f} launches {r.ctx.timing()['kernel_launches']} diff_px {diff}")
        r.close()
