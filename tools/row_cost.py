#!/usr/bin/env python
"""Kernel time of 32-row bands of the 4K bunny frame, top to bottom: where the expensive tiles are."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from conftest import load_golden_scene
from gp1_raytracer_2223_b200 import Renderer

r = Renderer(3840, 2160)
r.SetScene(load_golden_scene("bunny_4k"))
r.ctx.set_kernel_variant(1)
ROWS = 32
band = torch.empty((ROWS, 3840), dtype=torch.int32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
out = []
for y in range(0, 2160, ROWS):
    n = min(ROWS, 2160 - y)
    for _ in range(2):
        r.render_rows_device(y, n, band.data_ptr(), stream)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
    for a, b in ev:
        a.record(); r.render_rows_device(y, n, band.data_ptr(), stream); b.record()
    torch.cuda.synchronize()
    out.append(float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3)
print("us per 32-row band, top to bottom:")
print(" ".join(f"{v:.0f}" for v in out))
print("sum", sum(out))
r.close()
