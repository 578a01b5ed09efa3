#!/usr/bin/env python
"""Per-job clocks of the wavefront walk kernels (RT_B200_WAVE_TIMING + RT_B200_WAVE_JOB_CLOCKS): is a frame bound by the
sum of its jobs or by the longest one?  usage: wave_jobs.py [fixture ...]"""
import os, sys
os.environ["RT_B200_WAVE_TIMING"] = "1"
os.environ["RT_B200_WAVE_JOB_CLOCKS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import MANIFEST, load_golden_frame, load_golden_scene
from gp1_raytracer_2223_b200 import Renderer
for name in sys.argv[1:] or ["optional_320", "optional_640"]:
    info = MANIFEST[name]
    r = Renderer(info["width"], info["height"])
    r.SetScene(load_golden_scene(name))
    r.ctx.set_kernel_variant(4)
    print(f"== {name}", file=sys.stderr, flush=True)
    for _ in range(3):
        r.render_device()
    diff = int((r.download() != load_golden_frame(name)).sum())
    print(f"== {name} diff_px {diff}", file=sys.stderr, flush=True)
    r.close()
