#!/usr/bin/env python
"""A/B of library builds on the end-to-end call: rt_upload_mesh + rt_render into pinned host memory (4K bunny; RT_E2E_FIXTURE=name for another golden scene with a mesh).
python tools/e2e_ab.py build/ab/a.so ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, torch
from conftest import load_golden_frame, load_golden_scene
from gp1_raytracer_2223_b200 import Renderer
from conftest import MANIFEST
name = os.environ.get("RT_E2E_FIXTURE", "bunny_4k")
W, H = MANIFEST[name]["width"], MANIFEST[name]["height"]
scene = load_golden_scene(name)
r = Renderer(W, H); r.SetScene(scene)
host = torch.empty((H, W), dtype=torch.int32).pin_memory()
desc = r.ctx.mesh_descriptor(scene.meshes[0])
for _ in range(10):
    r.ctx.upload_mesh_descriptor(0, desc); r.render_host_ptr(host.data_ptr(), W * 4)
k, t, n = 0.0, 0.0, 60
t0 = time.perf_counter()
for _ in range(n):
    r.ctx.upload_mesh_descriptor(0, desc); tm = r.render_host_ptr(host.data_ptr(), W * 4); k += tm["kernel_ms"]; t += tm["total_ms"]
wall = (time.perf_counter() - t0) / n * 1e3
diff = int((host.numpy().view(np.uint32) != load_golden_frame(name)).sum())
dev = [r.render_device()["kernel_ms"] for _ in range(30)]
print(f"  e2e {wall:.4f} ms  (device: kernel {k / n:.4f}, kernel + copies {t / n:.4f})  device-only kernel {np.mean(dev):.4f}  diff_px {diff}")
''' % (ROOT, ROOT)
for lib in sys.argv[1:]:
    print(lib)
    out = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, RT_B200_LIB=os.path.abspath(lib)), capture_output=True, text=True)
    print(out.stdout + out.stderr[-600:])
