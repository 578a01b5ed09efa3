#!/usr/bin/env python
"""Differential fuzzing: seeded random scenes (tests/random_scenes.py) rendered by the CUDA path (both mesh bodies,
every kernel variant) and by the CPU oracle; reports every differing pixel.  Usage: python tests/tools/fuzz_parity.py [first_seed] [count] [max triangles per mesh]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from random_scenes import random_scene
from gp1_raytracer_2223_b200 import Renderer
from oracle import rt_oracle

first, count = (int(sys.argv[1]) if len(sys.argv) > 1 else 1000), (int(sys.argv[2]) if len(sys.argv) > 2 else 100)
max_triangles = int(sys.argv[3]) if len(sys.argv) > 3 else 120          # per mesh; thousands make trees deep enough for subtrees with several parts
bad = total = 0
for seed in range(first, first + count):
    rng = np.random.default_rng(seed)
    pow_materials = bool(rng.integers(0, 2))
    scene = random_scene(seed, n_spheres=int(rng.integers(0, 7)), n_planes=int(rng.integers(0, 8)), n_meshes=int(rng.integers(0, 4)),
                         n_triangles=int(rng.integers(1, max_triangles)), n_lights=int(rng.integers(0, 6)), pow_materials=pow_materials)
    # random camera jitter, occasionally axis aligned (zero direction components)
    if rng.integers(0, 4) == 0:
        scene.camera.right[:] = (1, 0, 0); scene.camera.up[:] = (0, 1, 0); scene.camera.forward[:] = (0, 0, 1)
    W, H = int(rng.integers(33, 200)), int(rng.integers(9, 90))
    mode, shadows = int(rng.integers(0, 4)), bool(rng.integers(0, 2))
    r = Renderer(W, H)
    for _ in range((mode - 3) % 4): r.CycleLightingMode()
    if not shadows: r.ToggleShadows()
    r.SetScene(scene)
    for gpu_path, oracle_path in ((1, rt_oracle.MESH_SLAB_LINEAR), (2, rt_oracle.MESH_BVH)):
        if gpu_path == 2 and not scene.meshes: continue
        want = rt_oracle.render(scene, W, H, mode, shadows, mesh_path=oracle_path)
        for variant in (1, 2, 3, 4):                      # 4 = wavefront, BVH body only
            if variant == 4 and gpu_path != 2: continue
            r.ctx.set_mesh_path(gpu_path); r.ctx.set_kernel_variant(variant)
            got = r.Render()
            if variant == 4:
                again = r.Render()         # the second wavefront frame chooses its units by the first one's job counts
                if not np.array_equal(got, again):
                    bad += 1
                    print(f"seed {seed}: two wavefront frames of the same scene differ in {int((got != again).sum())} px")
            total += 1
            nd = int((got != want).sum())
            if nd:
                g = got.view(np.uint8).reshape(H, W, 4).astype(int); w = want.view(np.uint8).reshape(H, W, 4).astype(int)
                mx = int(np.abs(g - w).max())
                uses_pow = pow_materials and mode in (2, 3)
                if not uses_pow or mx > 1 or nd > W * H // 1000:
                    bad += 1
                print(f"seed {seed} path {gpu_path} variant {variant} mode {mode} shadows {shadows} {W}x{H}: {nd} px differ, max {mx} LSB, pow={uses_pow}")
    r.close()
print(f"{total} frames compared, {bad} outside the parity bar")
sys.exit(1 if bad else 0)
