#!/usr/bin/env python
"""A/B of library builds: python tools/ab.py build/ab/a.so build/ab/b.so ... [-- fixture ...]
Each library is loaded in its own process (RT_B200_LIB), renders the fixtures with the default kernel choice
(rt_render_device) and prints mean / min kernel ms plus the differing pixels against the reference frame."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
from conftest import MANIFEST, load_golden_frame, load_golden_scene
from gp1_raytracer_2223_b200 import Renderer
for name in sys.argv[1:]:
    info = MANIFEST[name]
    r = Renderer(info["width"], info["height"])
    r.SetScene(load_golden_scene(name))
    for _ in range(5):
        r.render_device()
    ms = [r.render_device()["kernel_ms"] for _ in range(40)]
    diff = int((r.download() != load_golden_frame(name)).sum())
    print(f"  {name:16s} kernel ms mean {np.mean(ms):.4f} min {np.min(ms):.4f} diff_px {diff}")
    r.close()
''' % (ROOT, ROOT)

args = sys.argv[1:]
fixtures = ["bunny_4k"]
if "--" in args:
    fixtures = args[args.index("--") + 1:]
    args = args[:args.index("--")]
for lib in args or [os.path.join(ROOT, "gp1_raytracer_2223_b200", "librt_b200.so")]:
    print(lib)
    env = dict(os.environ, RT_B200_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, "-c", CHILD] + fixtures, env=env, capture_output=True, text=True)
    print(out.stdout + out.stderr[-800:])
