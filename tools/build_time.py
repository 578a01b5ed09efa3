#!/usr/bin/env python
"""Time of update_transforms_bvh_kernel alone: N rt_transform_mesh calls queued, executed back to back by one flush
(rt_read_mesh_build), wall clock / N.  Knobs through the environment: RT_B200_BUILD_GLOBAL=1 (work arrays in global
memory), RT_B200_BUILD_LOCAL=<n> (subtree size a warp finishes alone)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import GOLDEN, load_golden_scene
from gp1_raytracer_2223_b200 import Renderer
from gp1_raytracer_2223_b200.scene_file import load_rtmp

N = 200
for scene_name, steps_name in (("bunny_320_yaw10", "bunny_320_steps3"), ("optional_320", "optional_320_steps2")):
    scene = load_golden_scene(scene_name)
    steps = load_rtmp(os.path.join(GOLDEN, steps_name + ".rtmp"))[0]
    r = Renderer(320, 240)
    r.SetScene(scene)
    r.ctx.upload_mesh_source(0, steps.positions, steps.indices, steps.normals, scene.meshes[0].cull_mode, scene.meshes[0].material_index)
    r.ctx.set_mesh_device_bvh(0, True)
    T = steps.indices.shape[0]
    best = 1e9
    for rep in range(4):
        for k in range(N):
            r.ctx.transform_mesh(0, steps.transforms[k % len(steps.transforms)])
        t0 = time.perf_counter()
        r.ctx.read_mesh_build(0, T)
        best = min(best, (time.perf_counter() - t0) / N * 1e6)
    print(f"{scene_name}: {T} triangles, {best:.1f} us per UpdateTransforms + BuildBVH "
          f"(global={os.environ.get('RT_B200_BUILD_GLOBAL', '0')}, local={os.environ.get('RT_B200_BUILD_LOCAL', 'default')})")
    r.close()
