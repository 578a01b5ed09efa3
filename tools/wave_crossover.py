#!/usr/bin/env python
"""Where does RT_KERNEL_WAVEFRONT stop paying?  Scene_W4_OptionalScene at growing frame sizes, persistent / tiled / wavefront
forced, frames compared with each other (no golden at these sizes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import load_golden_scene
from gp1_raytracer_2223_b200 import Renderer
scene_name = sys.argv[1] if len(sys.argv) > 1 else "optional_320"
for w, h in ((320, 240), (640, 480), (960, 720), (1280, 960), (1920, 1080), (2560, 1440)):
    frames = {}
    line = f"{scene_name} {w}x{h}:"
    for variant, label in ((1, "tiled"), (3, "persistent"), (4, "wavefront")):
        r = Renderer(w, h)
        r.SetScene(load_golden_scene(scene_name))
        r.ctx.set_kernel_variant(variant)
        for _ in range(3):
            r.render_device()
        ms = [r.render_device()["kernel_ms"] for _ in range(15)]
        frames[label] = r.download()
        line += f"  {label} {np.mean(ms):.4f} ms"
        r.close()
    same = all(np.array_equal(frames["tiled"], f) for f in frames.values())
    print(line, " identical" if same else "  FRAMES DIFFER")
