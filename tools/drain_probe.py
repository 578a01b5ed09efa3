import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import load_golden_scene
from gp1_raytracer_2223_b200 import Renderer
r = Renderer(3840, 2160)
r.SetScene(load_golden_scene("bunny_4k"))
r.ctx.set_kernel_variant(3)
frame = torch.empty((2160, 3840), dtype=torch.int32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for world in (1, 8):
    for rep in range(6):
        r.render_strips_device(0, world, frame.data_ptr(), stream)
    torch.cuda.synchronize()
r.close()
