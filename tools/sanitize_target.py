#!/usr/bin/env python
"""Small workload for compute-sanitizer: every kernel variant / mesh body / present path on small frames."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import MANIFEST, load_golden_frame, load_golden_scene
from gp1_raytracer_2223_b200 import Renderer
from gp1_raytracer_2223_b200.scene_file import load_rtms
bad = 0
for name in ("bunny_333x77", "w4ref_101x203", "w3_320_brdf", "optional_320"):
    info = MANIFEST[name]
    r = Renderer(info["width"], info["height"])
    for _ in range((info["mode"] - 3) % 4): r.CycleLightingMode()
    if not info["shadows"]: r.ToggleShadows()
    sc = load_golden_scene(name); r.SetScene(sc)
    want = load_golden_frame(name)
    for path in (1, 2):
        if path == 2 and not sc.meshes: continue
        for variant in (1, 2, 3, 4):                          # 4 = wavefront: BVH body only (falls back to AUTO otherwise)
            if variant == 4 and path != 2: continue
            r.ctx.set_mesh_path(path); r.ctx.set_kernel_variant(variant)
            got = r.Render()
            d = int((got != want).sum()); bad += d > want.size // 1000
            print(name, path, variant, "diff", d)
    if sc.meshes: r.count_frame(mesh_path=2)
    r.count_frame(mesh_path=1)
    r.close()
r = Renderer(320, 240); sc = load_golden_scene("bunny_320_yaw10"); r.SetScene(sc)
src = load_rtms(os.path.join(ROOT, "tests", "golden", "bunny_320_yaw10.rtms"))[0]
r.ctx.upload_mesh_source(0, src.positions, src.indices, src.normals, sc.meshes[0].cull_mode, sc.meshes[0].material_index)
r.ctx.transform_mesh(0, src.transform)
print("device transform diff", int((r.Render() != load_golden_frame("bunny_320_yaw10")).sum()))
r.close()
sys.exit(1 if bad else 0)
