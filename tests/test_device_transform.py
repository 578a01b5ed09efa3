"""TriangleMesh::UpdateTransforms on the device.  The untransformed mesh is uploaded once,
each pose is 16 floats; frames must equal what the reference rendered after its own host-side UpdateTransforms."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, MANIFEST, MAX_LSB, MIN_IDENTICAL, compare_frames, load_golden_frame, load_golden_scene
from gp1_raytracer_2223_b200.scene_file import load_rtms

pytestmark = pytest.mark.gpu

CASES = ["bunny_320_yaw05", "bunny_320_yaw10", "bunny_320_time2", "w4ref_320_time13", "optional_320"]


def renderer_with_sources(name, source_name=None):
    from gp1_raytracer_2223_b200 import Renderer
    info = MANIFEST[name]
    scene = load_golden_scene(name)
    r = Renderer(info["width"], info["height"])
    r.SetScene(scene)
    sources = load_rtms(os.path.join(GOLDEN, (source_name or name) + ".rtms"))
    for i, (src, mesh) in enumerate(zip(sources, scene.meshes)):
        r.ctx.upload_mesh_source(i, src.positions, src.indices, src.normals, mesh.cull_mode, mesh.material_index)
    return r, scene, sources


@pytest.mark.parametrize("name", CASES)
def test_device_transform_renders_the_reference_frame(name):
    r, scene, sources = renderer_with_sources(name)
    from gp1_raytracer_2223_b200 import RtError
    with pytest.raises(RtError, match="no transform yet"):
        r.Render()
    for i, src in enumerate(sources):
        r.ctx.transform_mesh(i, src.transform)
    for variant in (1, 2, 3):
        r.ctx.set_kernel_variant(variant)
        got = r.Render()
        identical, max_err, n_diff = compare_frames(got, load_golden_frame(name))
        if name.startswith("bunny"):
            assert n_diff == 0, (variant, n_diff)
        assert identical >= MIN_IDENTICAL and max_err <= MAX_LSB, (variant, n_diff, max_err)
    r.close()


def test_new_pose_is_sixteen_floats():
    """Source uploaded once (in the triangle order of the yaw 0.5 run), then posed like the yaw 1.0 and the
    timer-driven runs: only the matrix changes."""
    r, scene, sources = renderer_with_sources("bunny_320_yaw05")
    for other in ("bunny_320_yaw10", "bunny_320_time2", "bunny_320_yaw05"):
        r.ctx.transform_mesh(0, load_rtms(os.path.join(GOLDEN, other + ".rtms"))[0].transform)
        assert np.array_equal(r.Render(), load_golden_frame(other)), other
    r.close()


def test_device_transform_matches_cpu_restatement_on_a_made_up_pose():
    from oracle import rt_oracle
    r, scene, sources = renderer_with_sources("bunny_320_yaw10")
    c, s_ = np.float32(np.cos(0.37)), np.float32(np.sin(0.37))
    m = np.array([[1.5 * c, 0, -1.5 * s_, 0], [0, 2.25, 0, 0], [1.75 * s_, 0, 1.75 * c, 0], [0.4, 0.1, -0.3, 1]], dtype=np.float32)
    r.ctx.transform_mesh(0, m)
    pos, nrm = rt_oracle.transform_mesh(sources[0].positions, sources[0].normals, m)
    scene.meshes[0].positions, scene.meshes[0].normals, scene.meshes[0].bvh_nodes = pos, nrm, None
    want = rt_oracle.render(scene, 320, 240)
    assert np.array_equal(r.Render(), want)
    r.close()
