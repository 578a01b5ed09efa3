#!/usr/bin/env python
"""TEST TOOLING (lives under tests/ because its host arm runs the CPU oracle as a stand-in for the reference's host code).

Animated bunny at 3840x2160: one TriangleMesh::UpdateTransforms per frame, three ways.

  host     UpdateTransforms + BuildBVH on the host (the oracle's C restatement standing in for the reference's own
           host code), rt_upload_mesh of the result, BVH body
  xform    rt_transform_mesh only (transform_mesh_kernel), slab + linear body
  device   rt_set_mesh_device_bvh: transform + BuildBVH on the device (update_transforms_bvh_kernel), BVH body

Every arm renders into the same pinned host buffer through rt_render; frames are compared between arms every step.
Prints wall-clock ms per frame (host work included) and the device-timed pixel kernel."""
import copy
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from conftest import GOLDEN, load_golden_scene
from gp1_raytracer_2223_b200 import Renderer
from gp1_raytracer_2223_b200.scene_file import load_rtmp
from oracle import rt_oracle

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
SCENE, STEPS = ("optional_320", "optional_320_steps2") if "--optional" in sys.argv else ("bunny_4k", "bunny_320_steps3")
N = 40


def pose(base, k):
    """base (S * R * T of the reference's scene) with an extra yaw in front: any matrix does, all arms share it."""
    a = np.float32(0.05 * (k + 1))
    c, s = np.float32(np.cos(a)), np.float32(np.sin(a))
    rot = np.array([[c, 0, -s, 0], [0, 1, 0, 0], [s, 0, c, 0], [0, 0, 0, 1]], dtype=np.float32)
    return (rot @ base).astype(np.float32)


def main():
    scene = load_golden_scene(SCENE)
    scene.width, scene.height = W, H
    all_steps = load_rtmp(os.path.join(GOLDEN, STEPS + ".rtmp"))
    frames = {}
    for arm in ("host", "xform", "device"):
        sc = copy.deepcopy(scene)
        r = Renderer(W, H)
        r.SetScene(sc)
        state = []
        for i, steps in enumerate(all_steps):
            idx = np.ascontiguousarray(steps.indices, dtype=np.int32).copy()
            nrm = np.ascontiguousarray(steps.normals, dtype=np.float32).copy()
            state.append((idx, nrm))
            if arm != "host":
                r.ctx.upload_mesh_source(i, steps.positions, steps.indices, steps.normals, sc.meshes[i].cull_mode, sc.meshes[i].material_index)
                if arm == "device":
                    r.ctx.set_mesh_device_bvh(i, True)
        host = torch.empty((H, W), dtype=torch.int32).pin_memory()
        out = []
        kernel_ms = []
        t_total = 0.0
        for k in range(N + 3):
            t0 = time.perf_counter()
            for i, steps in enumerate(all_steps):
                m = pose(steps.transforms[0], k)
                if arm == "host":
                    idx, nrm = state[i]
                    pos, tnrm, nodes = rt_oracle.update_transforms_bvh(steps.positions, idx, nrm, m)
                    mesh = sc.meshes[i]
                    mesh.positions, mesh.normals, mesh.indices, mesh.bvh_nodes = pos, tnrm, idx.reshape(-1, 3), nodes
                    r.ctx.upload_mesh(i, mesh)
                else:
                    r.ctx.transform_mesh(i, m)
            tm = r.render_host_ptr(host.data_ptr(), W * 4)
            dt = time.perf_counter() - t0
            if k >= 3:
                t_total += dt
                kernel_ms.append(tm["kernel_ms"])
            if k % 8 == 0:
                out.append(host.numpy().view(np.uint32).copy())
        frames[arm] = out
        print(f"{arm:>6s}: {t_total / N * 1e3:.3f} ms/frame wall (host work + H2D + kernels + D2H), pixel kernel {np.median(kernel_ms):.3f} ms, launches/frame {tm['kernel_launches']}")
        r.close()
    for arm in ("xform", "device"):
        same = all(np.array_equal(a, b) for a, b in zip(frames["host"], frames[arm]))
        print(f"{arm} frames identical to host arm: {same}")
        assert same


if __name__ == "__main__":
    main()
