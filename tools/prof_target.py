#!/usr/bin/env python
"""Smallest program that launches the flagship kernel (bunny 4K, Combined + shadows) a few times: the ncu target."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import MANIFEST, load_golden_scene  # noqa: E402
from gp1_raytracer_2223_b200 import Renderer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "bunny_4k"
w, h = MANIFEST[name]["width"], MANIFEST[name]["height"]
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
r = Renderer(w, h)
r.SetScene(load_golden_scene(name))
if variant:
    r.ctx.set_kernel_variant(variant)
for _ in range(4):
    print(r.render_device()["kernel_ms"])
r.close()
