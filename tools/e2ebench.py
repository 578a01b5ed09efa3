#!/usr/bin/env python
"""End-to-end rt_render (pinned host buffer) timing for a few pipeline band counts."""
import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 1:
    for bands in (1, 4, 8, 16, 32, 64):
        env = dict(os.environ, RT_B200_PIPELINE_BANDS=str(bands))
        subprocess.run([sys.executable, __file__, str(bands)], env=env, check=True)
    env = dict(os.environ, RT_B200_PIPELINE_BANDS="4", RT_B200_NO_STREAM_WAIT="1")
    subprocess.run([sys.executable, __file__, "4 (multi-launch fallback)"], env=env, check=True)
    sys.exit(0)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import load_golden_scene, load_golden_frame
from gp1_raytracer_2223_b200 import Renderer
scene = load_golden_scene("bunny_4k")
r = Renderer(3840, 2160); r.SetScene(scene)
host = torch.empty((2160, 3840), dtype=torch.int32).pin_memory()
for _ in range(5): r.render_host_ptr(host.data_ptr(), 3840 * 4)
t0 = time.perf_counter(); n = 50
for _ in range(n):
    r.ctx.upload_mesh(0, scene.meshes[0]); tm = r.render_host_ptr(host.data_ptr(), 3840 * 4)
dt = (time.perf_counter() - t0) / n * 1e3
ok = np.array_equal(host.numpy().view(np.uint32), load_golden_frame("bunny_4k"))
print(f"bands {sys.argv[1]:>2s}: e2e {dt:.3f} ms/frame  timing {tm}  exact={ok}")
