"""SURVEY.md 8(f) N1, second half: TriangleMesh::UpdateTransforms WITH BuildBVH on the device
(rt_set_mesh_device_bvh; reference source/DataTypes.h:210-236, 294-483).

tests/golden/<case>.rtmp holds the meshes before a sequence of UpdateTransforms calls made by the compiled
reference and the transform of each call; <case>.rtsc what the reference had after the last one (reordered
indices, BVHNode array) and <case>.frame.xz the frame it then rendered.  Replaying the calls on the device must
leave the same triangle order and the same tree - boxes bit for bit - and render the same frame."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, MANIFEST, MAX_LSB, MIN_IDENTICAL, compare_frames, load_golden_frame, load_golden_scene
from gp1_raytracer_2223_b200.scene_file import load_rtmp

pytestmark = pytest.mark.gpu

STEP_CASES = ["bunny_320_steps3", "w4ref_320_steps2", "optional_320_steps2"]


def renderer_with_steps(name):
    from gp1_raytracer_2223_b200 import Renderer
    info = MANIFEST[name]
    scene = load_golden_scene(name)
    r = Renderer(info["width"], info["height"])
    r.SetScene(scene)
    meshes = load_rtmp(os.path.join(GOLDEN, name + ".rtmp"))
    for i, (steps, mesh) in enumerate(zip(meshes, scene.meshes)):
        r.ctx.upload_mesh_source(i, steps.positions, steps.indices, steps.normals, mesh.cull_mode, mesh.material_index)
        r.ctx.set_mesh_device_bvh(i, True)
    return r, scene, meshes


def assert_same_tree(dev, ref):
    """Walks both trees from the root in IntersectionTest_BVH's order (left child, then left + 1; reference
    source/Utils.h:246-288).  Node numbers differ by design; everything the walk reads must not."""
    visited = 0
    stack = [(0, 0, -1)]
    while stack:
        d, r, escape = stack.pop()
        visited += 1
        dn, rn = dev[d], ref[r]
        assert dn["min_aabb"].tobytes() == rn["min_aabb"].tobytes() and dn["max_aabb"].tobytes() == rn["max_aabb"].tobytes(), (d, r)
        assert int(dn["escape"]) == escape, (d, int(dn["escape"]), escape)
        if rn["idx_count"] > 0:
            assert int(dn["triangle_count"]) * 3 == int(rn["idx_count"]) and int(dn["first"]) * 3 == int(rn["first_idx"]), (d, r)
        else:
            assert int(dn["triangle_count"]) == 0, (d, r)
            left, rleft = int(dn["first"]), int(rn["left_node"])
            stack.append((left + 1, rleft + 1, escape))
            stack.append((left, rleft, left + 1))
    return visited


@pytest.mark.parametrize("name", STEP_CASES)
def test_device_build_reproduces_the_reference(name):
    r, scene, meshes = renderer_with_steps(name)
    n_steps = meshes[0].transforms.shape[0]
    for s in range(n_steps):                       # all calls queued before the first frame: executed in call order
        for i, steps in enumerate(meshes):
            r.ctx.transform_mesh(i, steps.transforms[s])
    for i, want in enumerate(scene.meshes):
        idx, _, nodes = r.ctx.read_mesh_build(i, want.triangle_count)
        assert np.array_equal(idx.reshape(-1, 3), want.indices), "triangle order after the builds differs"
        assert len(nodes) == len(want.bvh_nodes)
        assert assert_same_tree(nodes, want.bvh_nodes) == len(nodes)
    r.ctx.set_mesh_path(2)                         # the BVH body, or an error if a mesh came out without nodes
    for variant in (1, 2, 3):
        r.ctx.set_kernel_variant(variant)
        got = r.Render()
        identical, max_err, n_diff = compare_frames(got, load_golden_frame(name))
        if name.startswith("bunny"):
            assert n_diff == 0, (variant, n_diff)
        assert identical >= MIN_IDENTICAL and max_err <= MAX_LSB, (variant, n_diff, max_err)
    r.close()


def test_frames_between_the_builds_do_not_disturb_them():
    """One frame per step, like the reference's main loop (Scene::Update, then Render): same final state."""
    from oracle import rt_oracle
    name = "bunny_320_steps3"
    r, scene, meshes = renderer_with_steps(name)
    steps = meshes[0]
    idx = np.ascontiguousarray(steps.indices, dtype=np.int32).copy()
    nrm = np.ascontiguousarray(steps.normals, dtype=np.float32).copy()
    for transform in steps.transforms:
        r.ctx.transform_mesh(0, transform)
        got = r.Render()
        pos, tnrm, nodes = rt_oracle.update_transforms_bvh(steps.positions, idx, nrm, transform)
        scene.meshes[0].positions, scene.meshes[0].normals, scene.meshes[0].indices, scene.meshes[0].bvh_nodes = pos, tnrm, idx.reshape(-1, 3).copy(), nodes
        assert np.array_equal(got, rt_oracle.render(scene, 320, 240))
        got_idx, got_nrm, _ = r.ctx.read_mesh_build(0, idx.size // 3)
        assert np.array_equal(got_idx, idx.reshape(-1))
        assert np.array_equal(got_nrm.view(np.uint32), nrm.reshape(-1, 3).view(np.uint32))
    assert np.array_equal(got, load_golden_frame(name))
    r.close()


def test_mesh_block_rewrite_keeps_the_built_meshes():
    """Uploading another mesh rewrites the whole mesh block from the host mirror; the device-built meshes have to be
    put back without running a build again (a second build would advance their triangle order)."""
    from gp1_raytracer_2223_b200 import Renderer
    name = "w4ref_320_steps2"
    scene = load_golden_scene(name)
    meshes = load_rtmp(os.path.join(GOLDEN, name + ".rtmp"))
    r = Renderer(MANIFEST[name]["width"], MANIFEST[name]["height"])
    r.SetScene(scene)                              # all meshes in their final state, from the host
    for i in (0, 1):                               # ... then two of the three handed to the device builder
        r.ctx.upload_mesh_source(i, meshes[i].positions, meshes[i].indices, meshes[i].normals, scene.meshes[i].cull_mode, scene.meshes[i].material_index)
        r.ctx.set_mesh_device_bvh(i, True)
        for transform in meshes[i].transforms:
            r.ctx.transform_mesh(i, transform)
    want = load_golden_frame(name)
    identical, max_err, _ = compare_frames(r.Render(), want)
    assert identical >= MIN_IDENTICAL and max_err <= MAX_LSB
    first = r.Render()
    r.ctx.upload_mesh(2, scene.meshes[2])          # mesh block rewritten; meshes 0 and 1 must survive
    assert np.array_equal(r.Render(), first)
    before = r.ctx.read_mesh_build(0, scene.meshes[0].triangle_count)[0]
    assert np.array_equal(before.reshape(-1, 3), scene.meshes[0].indices)
    r.close()


def test_read_back_order_feeds_the_slab_body():
    """The order read back from the device is a valid TriangleMesh state: uploaded again as a plain source
    (transform only, slab + linear body) it renders the same picture."""
    name = "bunny_320_steps3"
    r, scene, meshes = renderer_with_steps(name)
    for transform in meshes[0].transforms:
        r.ctx.transform_mesh(0, transform)
    first = r.Render()
    idx, nrm, _ = r.ctx.read_mesh_build(0, scene.meshes[0].triangle_count)
    r.ctx.upload_mesh_source(0, meshes[0].positions, idx.reshape(-1, 3), nrm, scene.meshes[0].cull_mode, scene.meshes[0].material_index)
    r.ctx.transform_mesh(0, meshes[0].transforms[-1])
    assert np.array_equal(r.Render(), first)
    assert np.array_equal(first, load_golden_frame(name))
    r.close()


def test_device_bvh_must_be_chosen_before_the_first_transform():
    from gp1_raytracer_2223_b200 import RtError
    name = "bunny_320_steps3"
    r, scene, meshes = renderer_with_steps(name)
    r.ctx.transform_mesh(0, meshes[0].transforms[0])
    with pytest.raises(RtError, match="before the first rt_transform_mesh"):
        r.ctx.set_mesh_device_bvh(0, True)
    r.close()


def made_up_mesh(rng, n_triangles, kind):
    """Triangle soups that stress the builder: sizes around the team / warp / leaf limits, flat and repeated geometry."""
    n_vertices = max(3, (n_triangles * 2) // 3 + 3)
    pos = rng.uniform(-1.0, 1.0, size=(n_vertices, 3)).astype(np.float32)
    if kind == "flat":
        pos[:, 2] = np.float32(0.25)                 # one axis without extent: skipped by FindBestSplitPlane
    if kind == "grid":
        pos = np.round(pos * 4).astype(np.float32) / 4   # many equal centroids and ties between planes
    idx = rng.integers(0, n_vertices, size=(n_triangles, 3)).astype(np.int32)
    if kind == "repeated":
        idx[:] = idx[0]                                 # every centroid the same: no axis splits, one wide leaf
    e1, e2 = pos[idx[:, 1]] - pos[idx[:, 0]], pos[idx[:, 2]] - pos[idx[:, 0]]
    nrm = np.cross(e1, e2).astype(np.float32)
    length = np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm = np.where(length > 0, nrm / np.maximum(length, 1e-30), np.float32([0, 1, 0])).astype(np.float32)
    return pos, idx, nrm


MADE_UP = [(1, "soup"), (2, "soup"), (3, "soup"), (9, "soup"), (33, "soup"), (64, "grid"), (65, "soup"), (257, "flat"), (300, "repeated"),
           (1025, "soup"), (1500, "grid"), (2600, "soup"), (5000, "soup")]


@pytest.mark.parametrize("n_triangles,kind", MADE_UP)
def test_device_build_matches_the_oracle_on_made_up_meshes(n_triangles, kind):
    """Three successive builds (each starts from the order the last one left) of a made-up mesh: triangle order, normals
    order and tree against the oracle's restatement after every one, and the frame after the last."""
    from gp1_raytracer_2223_b200 import Renderer
    from oracle import rt_oracle
    rng = np.random.default_rng(1000 + n_triangles)
    scene = load_golden_scene("bunny_320_yaw10")
    pos, idx, nrm = made_up_mesh(rng, n_triangles, kind)
    w, h = 96, 72
    scene.width, scene.height = w, h
    r = Renderer(w, h)
    r.SetScene(scene)
    mesh = scene.meshes[0]
    r.ctx.upload_mesh_source(0, pos, idx, nrm, mesh.cull_mode, mesh.material_index)
    r.ctx.set_mesh_device_bvh(0, True)
    idx, nrm = idx.copy(), nrm.copy()
    for step in range(3):
        a = np.float32(0.7 * step + 0.2)
        c, s_ = np.float32(np.cos(a)), np.float32(np.sin(a))
        m = np.array([[2 * c, 0, -2 * s_, 0], [0, 2, 0, 0], [2 * s_, 0, 2 * c, 0], [0.1 * step, 1.5, 0, 1]], dtype=np.float32)
        r.ctx.transform_mesh(0, m)
        tpos, tnrm, nodes = rt_oracle.update_transforms_bvh(pos, idx, nrm, m)
        got_idx, got_nrm, got_nodes = r.ctx.read_mesh_build(0, n_triangles)
        assert np.array_equal(got_idx.reshape(-1, 3), idx), f"triangle order differs after build {step}"
        assert np.array_equal(got_nrm.view(np.uint32), nrm.view(np.uint32))
        assert len(got_nodes) == len(nodes)
        assert assert_same_tree(got_nodes, nodes) == len(nodes)
    mesh.positions, mesh.indices, mesh.normals, mesh.bvh_nodes = tpos, idx, tnrm, nodes
    r.ctx.set_mesh_path(2)
    assert np.array_equal(r.Render(), rt_oracle.render(scene, w, h))
    r.close()
