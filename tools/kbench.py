#!/usr/bin/env python
"""Quick kernel A/B: renders fixtures N times through rt_render_device, prints mean/min kernel ms
and checks the frame against the reference-rendered golden.  Usage: python tools/kbench.py [names...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import MANIFEST, load_golden_frame, load_golden_scene  # noqa: E402
from gp1_raytracer_2223_b200 import Renderer  # noqa: E402

names = [a for a in sys.argv[1:] if not a.startswith("-")] or ["bunny_4k", "bunny_640", "w3_640", "w4ref_640"]
reps = 20
paths = {"slab/scalar": (1, 1), "slab/persist": (1, 3), "bvh/scalar": (2, 1), "bvh/persist": (2, 3)}
for name in names:
  for pname, (path, variant) in paths.items():
    info = MANIFEST[name]
    scene = load_golden_scene(name)
    if path == 2 and not scene.meshes:
        continue
    r = Renderer(info["width"], info["height"])
    for _ in range((info["mode"] - 3) % 4):
        r.CycleLightingMode()
    if not info["shadows"]:
        r.ToggleShadows()
    r.SetScene(scene)
    r.ctx.set_mesh_path(path)
    r.ctx.set_kernel_variant(variant)
    for _ in range(3):
        r.render_device()
    ms = [r.render_device()["kernel_ms"] for _ in range(reps)]
    got = r.download()
    diff = int((got != load_golden_frame(name)).sum())
    print(f"{name:24s} {pname:12s} kernel ms mean {np.mean(ms):.4f} min {np.min(ms):.4f}  diff_px {diff}")
    r.close()
