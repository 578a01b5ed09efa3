#!/usr/bin/env python
"""One rank's share of the 4K bunny frame through rt_render_strips_to_host on one GPU (no PCIe contention from other
ranks): kernel ms / kernel + copies ms by kernel variant.  python tools/share_e2e.py [world ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import load_golden_frame, load_golden_scene
from gp1_raytracer_2223_b200 import Renderer
worlds = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]
want = load_golden_frame("bunny_4k")
for world in worlds:
    for variant, label in ((0, "auto"), (1, "tiled"), (3, "persistent")):
        r = Renderer(3840, 2160)
        r.SetScene(load_golden_scene("bunny_4k"))
        r.ctx.set_kernel_variant(variant)
        host = torch.zeros((2160, 3840), dtype=torch.int32).pin_memory()
        for _ in range(5):
            r.render_strips_to_host(0, world, host.data_ptr(), 3840 * 4)
        k = t = 0.0
        t0 = time.perf_counter()
        n = 40
        for _ in range(n):
            tm = r.render_strips_to_host(0, world, host.data_ptr(), 3840 * 4)
            k += tm["kernel_ms"]; t += tm["total_ms"]
        wall = (time.perf_counter() - t0) / n * 1e3
        rows = np.zeros(2160, dtype=bool)
        for s0 in range(0, 2160, world * 8):
            rows[s0:s0 + 8] = True
        ok = np.array_equal(host.numpy().view(np.uint32)[rows], want[rows])
        dev = [r.render_strips_device(0, world, 0) for _ in range(0)]
        print(f"share 1/{world} {label:10s} kernel {k / n:.4f} ms  kernel + copies {t / n:.4f} ms  host call {wall:.4f} ms  {'ok' if ok else 'WRONG'}")
        r.close()
