"""the OBJ loader.  gp1_raytracer_2223_b200/obj_file.py restates Utils::ParseOBJ (reference
source/Utils.h:377-451); the fixtures tests/golden/*.rtmp hold what the compiled reference had in
TriangleMesh::positions / indices / normals after loading the same files (and one BuildBVH, which only permutes the
triangles), so positions must match bit for bit in order, triangles and their normals bit for bit as a set.
The .obj files are the reference's input data; oracle/Makefile copies them next to the reference binary
(oracle/_ref/Resources, not in git): without them the file tests skip."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, MANIFEST, MAX_LSB, MIN_IDENTICAL, ROOT, compare_frames, load_golden_frame, load_golden_scene
from gp1_raytracer_2223_b200.obj_file import face_normals, parse_obj
from gp1_raytracer_2223_b200.scene_file import load_rtmp, load_rtms

RESOURCES = os.path.join(ROOT, "oracle", "_ref", "Resources")
CASES = [("lowpoly_bunny2.obj", "bunny_320_steps3"), ("Assignment3D1.obj", "optional_320_steps2")]


def resource(name):
    path = os.path.join(RESOURCES, name)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/Resources is not populated (make -C oracle ref needs /root/reference)")
    return path


def triangle_table(indices, normals):
    return sorted((tuple(int(v) for v in tri), nrm.view(np.uint32).tobytes()) for tri, nrm in zip(indices, normals))


@pytest.mark.parametrize("obj,fixture", CASES)
def test_parse_obj_reproduces_the_reference_mesh(obj, fixture):
    mesh = parse_obj(resource(obj))
    want = load_rtmp(os.path.join(GOLDEN, fixture + ".rtmp"))[0]
    assert mesh.positions.shape == want.positions.shape
    assert np.array_equal(mesh.positions.view(np.uint32), want.positions.view(np.uint32))
    assert mesh.indices.shape == want.indices.shape
    assert triangle_table(mesh.indices, mesh.normals) == triangle_table(want.indices, want.normals)


def test_face_normals_on_the_reference_triangles_in_their_order():
    """Same arithmetic, independent of the file: normals of the fixture's own (permuted) triangles."""
    for fixture in ("bunny_320_steps3", "optional_320_steps2", "w4ref_320_steps2"):
        for want in load_rtmp(os.path.join(GOLDEN, fixture + ".rtmp")):
            got = face_normals(want.positions, want.indices)
            if fixture.startswith("w4ref"):
                continue        # those triangles carry hand-written normals (source/Scene.cpp), not ParseOBJ's
            assert np.array_equal(got.view(np.uint32), want.normals.view(np.uint32)), fixture


def test_loader_semantics_on_a_made_up_file(tmp_path):
    text = "\n".join([
        "# comment v 9 9 9",
        "o thing",
        "v 0 0 0",
        "v 1.5 0 0   trailing words are dropped",
        "v 0 2e0 0",
        "v 0 0 -3",
        "vn 0 0 1",
        "vt 0.5 0.5",
        "f 1/7/7 2/8/8 3/9/9",
        "f 1.9 3 4 2",              # corners are read as floats and truncated; a fourth corner is ignored
        "f 1 2",                    # not three corners: skipped
        "",
        "s off",
    ]) + "\n"
    path = tmp_path / "made_up.obj"
    path.write_text(text)
    mesh = parse_obj(str(path))
    assert mesh.positions.tolist() == [[0, 0, 0], [1.5, 0, 0], [0, 2, 0], [0, 0, -3]]
    assert mesh.indices.tolist() == [[0, 1, 2], [0, 2, 3]]
    assert mesh.normals[0].tolist() == [0, 0, 1]
    assert mesh.normals[1].tolist() == [-1, 0, 0]
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nf 1 2 3\n")
    with pytest.raises(ValueError, match="outside"):
        parse_obj(str(bad))
    degenerate = tmp_path / "degenerate.obj"
    degenerate.write_text("v 0 0 0\nv 1 1 1\nv 2 2 2\nf 1 2 3\n")
    assert np.isnan(parse_obj(str(degenerate)).normals).all()          # 0 / 0, like the reference


@pytest.mark.gpu
@pytest.mark.parametrize("obj,scale,name,posed", [("lowpoly_bunny2.obj", 2.0, "bunny_640", False), ("lowpoly_bunny2.obj", 2.0, "bunny_320_yaw05", True),
                                                  ("Assignment3D1.obj", 0.03, "optional_320", False), ("Assignment3D1.obj", 0.03, "optional_320_steps2", True)])
def test_obj_file_to_frame_on_the_device(obj, scale, name, posed):
    """File -> parse_obj -> rt_upload_mesh_source -> device UpdateTransforms + BuildBVH -> frame, with nothing of the
    mesh taken from the reference: the first build is the one of Initialize() (scale only, source/Scene.cpp:413-417,
    450-454), posed fixtures add the build of their pose.  Must be the frame the reference rendered."""
    from gp1_raytracer_2223_b200 import Renderer
    mesh = parse_obj(resource(obj))
    info = MANIFEST[name]
    scene = load_golden_scene(name)
    r = Renderer(info["width"], info["height"])
    r.SetScene(scene)
    r.ctx.upload_mesh_source(0, mesh.positions, mesh.indices, mesh.normals, scene.meshes[0].cull_mode, scene.meshes[0].material_index)
    r.ctx.set_mesh_device_bvh(0, True)
    s = np.float32(scale)
    r.ctx.transform_mesh(0, np.diag([s, s, s, np.float32(1)]).astype(np.float32))          # Initialize(): Scale + UpdateTransforms
    if posed and os.path.exists(os.path.join(GOLDEN, name + ".rtmp")):
        for transform in load_rtmp(os.path.join(GOLDEN, name + ".rtmp"))[0].transforms:       # the fixture's poses, one UpdateTransforms each
            r.ctx.transform_mesh(0, transform)
    elif posed:
        r.ctx.transform_mesh(0, load_rtms(os.path.join(GOLDEN, name + ".rtms"))[0].transform)   # the fixture's pose: one more UpdateTransforms
    got = r.Render()
    identical, max_err, n_diff = compare_frames(got, load_golden_frame(name))
    if name.startswith("bunny"):
        assert n_diff == 0
    assert identical >= MIN_IDENTICAL and max_err <= MAX_LSB, (n_diff, max_err)
    idx = r.ctx.read_mesh_build(0, mesh.indices.shape[0])[0]
    assert np.array_equal(idx.reshape(-1, 3), scene.meshes[0].indices), "triangle order differs from the reference's after the same builds"
    r.close()
