import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden_scene
from gp1_raytracer_2223_b200 import Renderer
r = Renderer(3840, 2160); r.SetScene(load_golden_scene("bunny_4k")); r.ctx.set_mesh_path(1)
for _ in range(4): print(r.render_device()["kernel_ms"])
