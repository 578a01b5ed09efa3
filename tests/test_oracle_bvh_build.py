"""The oracle's restatement of TriangleMesh::UpdateTransforms + BuildBVH (reference source/DataTypes.h:210-236,
294-483) against the reference itself: tests/golden/<case>.rtmp holds the meshes BEFORE a sequence of
UpdateTransforms calls made by the compiled reference (oracle/ref_driver.cpp --yaw-steps) and the transform of each
call; <case>.rtsc holds what the reference had AFTER the last one (transformed vertices, reordered indices and
normals, BVHNode array).  Replaying the calls through the restatement must reproduce all of it bit for bit."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden_scene
from gp1_raytracer_2223_b200.scene_file import load_rtmp
from oracle import rt_oracle

STEP_CASES = ["bunny_320_steps3", "w4ref_320_steps2", "optional_320_steps2"]


def same_nodes(got, want):
    """Bitwise equality of two BVHNode arrays, except leftNode of leaves: the reference never writes it for a leaf
    (source/DataTypes.h:371-379), so there it holds whatever an earlier build left in pBVHNodes."""
    got, want = got.copy(), want.copy()
    got["left_node"][got["idx_count"] > 0] = 0
    want["left_node"][want["idx_count"] > 0] = 0
    return got.tobytes() == want.tobytes()


def replay(steps):
    """Runs every step of one mesh through the oracle; returns the state after the last one."""
    idx = np.ascontiguousarray(steps.indices, dtype=np.int32).copy()
    nrm = np.ascontiguousarray(steps.normals, dtype=np.float32).copy()
    out = None
    for transform in steps.transforms:
        out = rt_oracle.update_transforms_bvh(steps.positions, idx, nrm, transform)
    return idx, nrm, out


@pytest.mark.parametrize("name", STEP_CASES)
def test_bvh_build_restatement_reproduces_the_reference(name):
    scene = load_golden_scene(name)
    meshes = load_rtmp(os.path.join(GOLDEN, name + ".rtmp"))
    assert len(meshes) == len(scene.meshes)
    for steps, want in zip(meshes, scene.meshes):
        idx, _, (pos, tnrm, nodes) = replay(steps)
        assert np.array_equal(idx, want.indices), "triangle order after the builds differs"
        assert np.array_equal(pos.view(np.uint32), want.positions.view(np.uint32))
        assert np.array_equal(tnrm.view(np.uint32), want.normals.view(np.uint32))
        assert len(nodes) == len(want.bvh_nodes)
        assert same_nodes(nodes, want.bvh_nodes), "BVH nodes differ from the reference's"


def test_every_build_depends_on_the_previous_order():
    """Guards the fixture itself: replaying only the LAST step from the initial order must NOT give the reference's
    final triangle order (the in-place partition makes the build history-dependent)."""
    name = "bunny_320_steps3"
    scene = load_golden_scene(name)
    steps = load_rtmp(os.path.join(GOLDEN, name + ".rtmp"))[0]
    idx = np.ascontiguousarray(steps.indices, dtype=np.int32).copy()
    nrm = np.ascontiguousarray(steps.normals, dtype=np.float32).copy()
    rt_oracle.update_transforms_bvh(steps.positions, idx, nrm, steps.transforms[-1])
    assert not np.array_equal(idx, scene.meshes[0].indices)
