#!/usr/bin/env python
"""What does the progressive present's per-tile completion signal cost the kernel?  The whole 4K bunny frame rendered
into the context's own frame buffer with and without band counters (no copies either way), CUDA-event timed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import load_golden_scene
from gp1_raytracer_2223_b200 import Renderer
r = Renderer(3840, 2160)
r.SetScene(load_golden_scene("bunny_4k"))
stream = torch.cuda.Stream()
def timed(fn, n=30):
    for _ in range(5):
        fn()
    stream.synchronize()
    ms = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); stream.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.mean(ms)), float(np.min(ms))
for variant, label in ((0, "auto"), (1, "tiled"), (3, "persistent")):
    r.ctx.set_kernel_variant(variant)
    print(label, "plain  %.4f (min %.4f)" % timed(lambda: r.render_strips_to_frame(0, 1, 0, stream.cuda_stream)),
          " with band counters (16 bands) %.4f (min %.4f)" % timed(lambda: r.render_strips_to_frame_banded(0, 1, 0, 16, stream.cuda_stream)))
r.close()
