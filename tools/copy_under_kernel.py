#!/usr/bin/env python
"""Is a device-to-host copy slower while the pixel kernel runs?  33 MB in 16 chunks on a copy stream, alone and while the
4K bunny frame renders on another stream (CUDA events on the copy stream)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import load_golden_scene
from gp1_raytracer_2223_b200 import Renderer
r = Renderer(3840, 2160)
r.SetScene(load_golden_scene("bunny_4k"))
frame = torch.zeros((2160, 3840), dtype=torch.int32, device="cuda")
other = torch.zeros((2160, 3840), dtype=torch.int32, device="cuda")
host = torch.zeros((2160, 3840), dtype=torch.int32).pin_memory()
ks, cs = torch.cuda.Stream(), torch.cuda.Stream()
def copies(chunks):
    rows = 2160 // chunks
    with torch.cuda.stream(cs):
        for c in range(chunks):
            host[c * rows:(c + 1) * rows].copy_(other[c * rows:(c + 1) * rows], non_blocking=True)
for chunks in (1, 16):
    for with_kernel in (False, True):
        ms = []
        for rep in range(12):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if with_kernel:
                r.render_strips_device(0, 1, frame.data_ptr(), ks.cuda_stream)
            a.record(cs); copies(chunks); b.record(cs)
            torch.cuda.synchronize()
            if rep >= 2:
                ms.append(a.elapsed_time(b))
        print(f"{chunks:2d} chunk(s), kernel running: {with_kernel}:  copy {np.mean(ms):.4f} ms -> {33.1776 / np.mean(ms):.1f} GB/s")
r.close()
