#!/usr/bin/env python
"""One GPU rendering only rank r's share of the 4K bunny frame (strips r, r + world, ...): how the per-GPU kernel
time shrinks with the share, without any multi-GPU effect.  Prints ms per launch for each kernel variant."""
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
frame = torch.empty((2160, 3840), dtype=torch.int32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for variant, label in ((1, "tiled"), (3, "persistent")):
    r.ctx.set_kernel_variant(variant)
    for world in (1, 2, 4, 8):
        times = []
        for rank in range(world) if world > 1 else (0,):
            for _ in range(3):
                r.render_strips_device(rank, world, frame.data_ptr(), stream)
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
            for a, b in ev:
                a.record(); r.render_strips_device(rank, world, frame.data_ptr(), stream); b.record()
            torch.cuda.synchronize()
            times.append(float(np.median([a.elapsed_time(b) for a, b in ev])))
        print(f"{label:>10s} world {world}: share of the slowest rank {max(times):.4f} ms, fastest {min(times):.4f} ms, ideal {0.805 / world:.4f} ms")
r.close()
