"""Parity of the CUDA path (through the C ABI) with the reference.

Bar (BASELINE.json north_star): per frame, >= 99.9 % of pixels identical in every RGBA8
channel and no channel off by more than 1 LSB, against the frame the compiled reference
rendered for the same scene and camera (tests/golden/).  Scenes whose shading never calls
powf (Solid / Lambert materials: W1, W2, both triangle scenes in ObservedArea / Radiance
modes, the whole bunny scene) must match bit for bit.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import (MANIFEST, MAX_LSB, MIN_IDENTICAL, compare_frames, golden_names, load_golden_frame,
                      load_golden_scene)

pytestmark = pytest.mark.gpu

EXACT = {n for n in MANIFEST if n.startswith(("bunny", "w1", "w2")) or n.endswith(("observed", "radiance"))}


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    from gp1_raytracer_2223_b200 import build
    build.ensure()
    return torch


def make_renderer(name, **kw):
    from gp1_raytracer_2223_b200 import Renderer
    info = MANIFEST[name]
    r = Renderer(info["width"], info["height"], **kw)
    for _ in range((info["mode"] - 3) % 4):
        r.CycleLightingMode()
    if not info["shadows"]:
        r.ToggleShadows()
    r.SetScene(load_golden_scene(name))
    return r


@pytest.mark.parametrize("mesh_path", [1, 2], ids=["slab_linear", "bvh"])
@pytest.mark.parametrize("name", golden_names())
def test_frame_matches_reference(gpu, name, mesh_path):
    """Both bodies of HitTest_TriangleMesh (reference source/Utils.h:296-325): the slab + linear loop and
    the BVH walk over the reference's own nodes."""
    if mesh_path == 2 and not load_golden_scene(name).meshes:
        pytest.skip("scene has no triangle mesh")
    r = make_renderer(name)
    r.ctx.set_mesh_path(mesh_path)
    got = r.Render()
    want = load_golden_frame(name)
    identical, max_err, n_diff = compare_frames(got, want)
    if name in EXACT:
        assert n_diff == 0, f"{name}: {n_diff} pixels differ (max {max_err} LSB) on a powf-free scene"
    assert identical >= MIN_IDENTICAL and max_err <= MAX_LSB, f"{name}: {n_diff} px differ, max {max_err} LSB"
    r.close()


@pytest.mark.parametrize("mesh_path", [1, 2], ids=["slab_linear", "bvh"])
@pytest.mark.parametrize("name", golden_names())
def test_packed_kernel_matches_reference(gpu, name, mesh_path):
    """RT_KERNEL_PACKED (two pixels per thread on FFMA2) must be the same function as the scalar kernel."""
    if mesh_path == 2 and not load_golden_scene(name).meshes:
        pytest.skip("scene has no triangle mesh")
    r = make_renderer(name)
    r.ctx.set_mesh_path(mesh_path)
    r.ctx.set_kernel_variant(2)
    got = r.Render().copy()
    r.ctx.set_kernel_variant(1)
    assert np.array_equal(got, r.Render()), "packed and scalar kernels disagree"
    r.ctx.set_kernel_variant(3)
    assert np.array_equal(got, r.Render()), "persistent and tiled kernels disagree"
    r.ctx.set_kernel_variant(4)          # rays as the unit of work (five launches); runs AUTO's choice where it cannot (slab body, no meshes)
    assert np.array_equal(got, r.Render()), "wavefront and tiled kernels disagree"
    identical, max_err, n_diff = compare_frames(got, load_golden_frame(name))
    if name in EXACT:
        assert n_diff == 0
    assert identical >= MIN_IDENTICAL and max_err <= MAX_LSB
    r.close()


def test_frame_matches_cpu_oracle_on_fresh_pose(gpu):
    """A pose that is in no fixture: oracle and CUDA path on the same inputs."""
    from oracle import rt_oracle
    scene = load_golden_scene("bunny_320_cam")
    scene.camera.origin[:] = (-1.5, 2.25, -6.5)
    r = make_renderer("bunny_320_cam")
    r.SetScene(scene)
    want = rt_oracle.render(scene, 320, 240)
    assert np.array_equal(r.Render(), want)
    r.close()


def test_counters_match_oracle(gpu):
    from oracle import rt_oracle
    for name in ("bunny_320_yaw10", "w4ref_320_cam", "w3_320_fov90"):
        info = MANIFEST[name]
        scene = load_golden_scene(name)
        r = make_renderer(name)
        for gpu_path, oracle_path in ((1, rt_oracle.MESH_SLAB_LINEAR), (2, rt_oracle.MESH_BVH)):
            if gpu_path == 2 and not scene.meshes:
                continue
            _, want = rt_oracle.render(scene, info["width"], info["height"], info["mode"], bool(info["shadows"]),
                                       mesh_path=oracle_path, counters=True)
            got = r.count_frame(mesh_path=gpu_path)
            assert np.array_equal(got[:38], want[:38]), (name, gpu_path, np.nonzero(got[:38] != want[:38]))
        r.close()


def test_row_bands_and_strips_are_bit_identical(gpu):
    """Fake multi-GPU on one device: N contiguous bands, and N interleaved strip bands + unstripe."""
    torch = gpu
    from gp1_raytracer_2223_b200 import bands
    name = "bunny_333x77"
    r = make_renderer(name)
    full = r.Render()
    W, H = 333, 77
    for parts in (2, 3, 8):
        out = torch.zeros((H, W), dtype=torch.int32, device="cuda")
        edges = np.linspace(0, H, parts + 1).astype(int)
        for a, b in zip(edges[:-1], edges[1:]):
            r.render_rows_device(int(a), int(b - a), out[a:].data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy().view(np.uint32), full)

        spr = bands.strips_per_rank(H, parts)
        packed = torch.zeros((parts, spr * bands.STRIP_ROWS, W), dtype=torch.int32, device="cuda")
        for rank in range(parts):
            r.render_strips_device(rank, parts, packed[rank].data_ptr())
        frame = torch.zeros((H, W), dtype=torch.int32, device="cuda")
        r.unstripe_device(packed.data_ptr(), frame.data_ptr(), parts, spr)
        torch.cuda.synchronize()
        assert np.array_equal(frame.cpu().numpy().view(np.uint32), full)
        assert np.array_equal(bands.unstripe_numpy(packed.cpu().numpy().view(np.uint32), W, H, parts), full)
    r.close()


def test_strips_on_torch_stream_vector_path(gpu):
    torch = gpu
    from gp1_raytracer_2223_b200 import bands
    r = make_renderer("bunny_640")
    full = load_golden_frame("bunny_640")
    W, H, world = 640, 480, 4
    spr = bands.strips_per_rank(H, world)
    packed = torch.zeros((world, spr * bands.STRIP_ROWS, W), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for rank in range(world):
        r.render_strips_device(rank, world, packed[rank].data_ptr(), stream)
    frame = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    r.unstripe_device(packed.data_ptr(), frame.data_ptr(), world, spr, stream)
    torch.cuda.synchronize()
    assert np.array_equal(frame.cpu().numpy().view(np.uint32), full)
    r.close()


@pytest.mark.parametrize("name,world,origin", [("bunny_4k", 8, None), ("bunny_4k", 4, None), ("bunny_640", 4, None), ("bunny_640", 3, (2.5, 3.0, -9.0)),
                                               ("bunny_640", 2, (0.0, 6.5, -7.0)), ("bunny_333x77", 2, None), ("w4ref_101x203", 3, None), ("optional_320", 4, None)])
def test_mesh_rectangle_first_order_of_the_persistent_kernel(gpu, name, world, origin):
    """Device-only launches of the persistent kernel hand out the tiles under the meshes' screen rectangle first
    (tiles_to_render_first): every queue position must still mean exactly one tile.  Shares of a rank / world split,
    cameras that push the rectangle partly out of the frame, odd frame sizes; against the same frame rendered in
    the plain order by the tiled kernel (and the reference frame where the camera is the fixture's)."""
    torch = gpu
    from gp1_raytracer_2223_b200 import bands
    r = make_renderer(name)
    scene = load_golden_scene(name)
    if origin is not None:
        scene.camera.origin[:] = origin
        r.SetScene(scene)
    info = MANIFEST[name]
    W, H = info["width"], info["height"]
    spr = bands.strips_per_rank(H, world)
    frames = {}
    for variant in (1, 3):
        r.ctx.set_kernel_variant(variant)
        packed = torch.full((world, spr * bands.STRIP_ROWS, W), -1, dtype=torch.int32, device="cuda")
        for rank in range(world):
            r.render_strips_device(rank, world, packed[rank].data_ptr())
        torch.cuda.synchronize()
        frames[variant] = bands.unstripe_numpy(packed.cpu().numpy().view(np.uint32), W, H, world)
    assert np.array_equal(frames[1], frames[3])
    # the same share again and again: from the second launch on the cells are walked in the order of their measured
    # cost (prepare_cell_order), and the order keeps changing with the measurements
    for rank in range(world):
        for repeat in range(4):
            packed[rank].fill_(-1)
            r.render_strips_device(rank, world, packed[rank].data_ptr())
            torch.cuda.synchronize()
            got = bands.unstripe_numpy(packed.cpu().numpy().view(np.uint32), W, H, world)
            mine = np.zeros(H, dtype=bool)
            for s0 in range(rank * 8, H, world * 8):
                mine[s0:s0 + 8] = True
            assert np.array_equal(got[mine], frames[1][mine]), (rank, repeat)
    if origin is None:
        identical, max_err, n_diff = compare_frames(frames[3], load_golden_frame(name))
        if name.startswith("bunny"):
            assert n_diff == 0
        assert identical >= MIN_IDENTICAL and max_err <= MAX_LSB
    r.close()


def test_pixel_format_and_pitch(gpu):
    """Other SDL surface formats: BGR shifts + alpha mask, and a surface pitch wider than 4*W."""
    from oracle import rt_oracle
    name = "w4ref_101x203"
    scene = load_golden_scene(name)
    r = make_renderer(name, shifts=(0, 8, 16), alpha_mask=0xFF000000)
    want = rt_oracle.render(scene, 101, 203, shifts=(0, 8, 16), alpha_mask=0xFF000000)
    wide = np.zeros((203, 128), dtype=np.uint32)
    view = wide[:, :101]
    r.Render(out=view)
    identical, max_err, _ = compare_frames(np.ascontiguousarray(view), want)
    assert identical >= MIN_IDENTICAL and max_err <= MAX_LSB
    assert not wide[:, 101:].any()
    r.close()


def test_mesh_reupload_follows_update_transforms(gpu):
    """Scene::Update re-transforms the mesh every frame: re-upload and compare with the posed fixture."""
    r = make_renderer("bunny_320_yaw05")
    first = r.Render().copy()
    r.ctx.upload_mesh(0, load_golden_scene("bunny_320_yaw25").meshes[0])
    assert np.array_equal(r.Render(), load_golden_frame("bunny_320_yaw25"))
    r.ctx.upload_mesh(0, load_golden_scene("bunny_320_yaw05").meshes[0])
    assert np.array_equal(r.Render(), first)
    r.close()


def test_lighting_mode_cycle_and_shadow_toggle(gpu):
    """F3 / F2 semantics of the reference Renderer (Renderer.cpp:189-193, Renderer.h:34-36)."""
    r = make_renderer("bunny_640")
    assert np.array_equal(r.Render(), load_golden_frame("bunny_640"))
    for name in ("bunny_640_observed", "bunny_640_radiance", "bunny_640_brdf", "bunny_640"):
        r.CycleLightingMode()
        assert np.array_equal(r.Render(), load_golden_frame(name)), name
    r.ToggleShadows()
    assert np.array_equal(r.Render(), load_golden_frame("bunny_640_noshadow"))
    r.close()


def test_empty_scene_and_errors(gpu):
    from gp1_raytracer_2223_b200 import Renderer, RtError
    from gp1_raytracer_2223_b200.scene_file import FlatScene, Camera
    z3 = np.zeros((3, 0), dtype=np.float32)
    empty = FlatScene(z3, np.zeros(0, np.float32), np.zeros(0, np.uint8), z3, z3, np.zeros(0, np.uint8), z3, z3, z3,
                      np.zeros(0, np.float32), np.zeros(0, np.int32), load_golden_scene("w1_640").materials[:1], [],
                      Camera(np.zeros(3, np.float32), 0.4142, np.array([1, 0, 0], np.float32),
                             np.array([0, 1, 0], np.float32), np.array([0, 0, 1], np.float32)))
    r = Renderer(33, 9)
    r.SetScene(empty)
    assert not r.Render().any()                       # nothing to hit: black, like the reference
    lib, h = r.ctx.lib, r.ctx.handle
    assert lib.rt_set_mesh_count(h, 10_000) == 5      # RT_ERR_CAPACITY
    assert lib.rt_set_mesh_count(h, 1) == 0
    with pytest.raises(RtError, match="never uploaded"):
        r.Render()
    assert lib.rt_set_mesh_count(h, 0) == 0
    assert lib.rt_render(h, None, None, None, 0) != 0
    assert lib.rt_download_frame(h, None, 0) != 0
    r.close()


def test_default_path_is_the_shipped_bvh_and_bad_trees_are_rejected(gpu):
    import ctypes as C
    from gp1_raytracer_2223_b200._abi import SceneViews
    scene = load_golden_scene("bunny_320_yaw10")
    r = make_renderer("bunny_320_yaw10")
    assert np.array_equal(r.Render(), load_golden_frame("bunny_320_yaw10"))     # auto -> BVH (nodes were uploaded)
    # a mesh without nodes under a forced BVH path is an error, not a silent fallback
    v = SceneViews.__new__(SceneViews)
    v._keep = []
    desc = SceneViews.mesh_desc(v, scene.meshes[0], with_bvh=False)
    assert r.ctx.lib.rt_upload_mesh(r.ctx.handle, 0, C.byref(desc)) == 0
    r.ctx.set_mesh_path(2)
    cam_ok = False
    try:
        r.Render()
    except Exception as e:
        cam_ok = "without BVH nodes" in str(e)
    assert cam_ok
    r.ctx.set_mesh_path(0)
    assert np.array_equal(r.Render(), load_golden_frame("bunny_320_yaw10"))     # auto -> slab + linear now
    # corrupt trees: child index out of range, and a cycle
    import copy
    for corrupt in ("range", "cycle"):
        bad = copy.deepcopy(scene.meshes[0])
        inner = int(np.nonzero(bad.bvh_nodes["idx_count"] == 0)[0][-1])
        bad.bvh_nodes["left_node"][inner] = 10_000_000 if corrupt == "range" else 0
        d = SceneViews.mesh_desc(v, bad)
        assert r.ctx.lib.rt_upload_mesh(r.ctx.handle, 0, C.byref(d)) != 0
    r.close()


def test_full_size_properties(gpu):
    """BASELINE.json full size (3840x2160): determinism, band independence, reference hash."""
    from oracle import rt_oracle
    torch = gpu
    r = make_renderer("bunny_4k")
    a = r.Render()
    b = r.Render()
    assert np.array_equal(a, b)
    assert f"{rt_oracle.fnv1a64(a):016x}" == MANIFEST["bunny_4k"]["fnv1a64"]
    band = torch.zeros((64, 3840), dtype=torch.int32, device="cuda")
    r.render_rows_device(1403, 64, band.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(band.cpu().numpy().view(np.uint32), a[1403:1467])
    r.close()


def test_in_process_multi_device_matches_single(gpu):
    """rt_create with several devices: strips are rendered on every GPU and stored straight into device 0's
    frame (peer stores).  Must be bit-identical to the one-GPU frame."""
    torch = gpu
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    want = load_golden_frame("bunny_640")
    for ids in ([0, 1], list(range(n)), [1, 0]):
        r = make_renderer("bunny_640", device_ids=ids)
        assert r.ctx.device_count == len(ids)
        for path in (1, 2):
            r.ctx.set_mesh_path(path)
            assert np.array_equal(r.Render(), want), (ids, path)
        r.close()
    r = make_renderer("bunny_4k", device_ids=list(range(n)))
    assert np.array_equal(r.Render(), load_golden_frame("bunny_4k"))
    print("multi-device timing", r.ctx.timing())
    r.close()


@pytest.mark.parametrize("name,world,extra_pitch", [("bunny_640", 3, 0), ("bunny_333x77", 2, 0), ("bunny_333x77", 5, 64), ("w4ref_101x203", 4, 0),
                                                    ("bunny_4k", 8, 0)])
def test_direct_present_of_rank_strips(gpu, name, world, extra_pitch):
    """rt_render_strips_to_host: every rank copies exactly the strips it rendered into the shared host surface.  The
    ranks run one after the other on this GPU here; together they must produce the reference frame, and a rank must
    not touch the rows of the others (surface pre-filled with a marker; pitch wider than the rows included)."""
    r = make_renderer(name)
    info = MANIFEST[name]
    w, h = info["width"], info["height"]
    pitch = 4 * w + extra_pitch
    surface = np.full((h, pitch // 4), 0xDEADBEEF, dtype=np.uint32)
    want = load_golden_frame(name)
    for rank in range(world):
        before = surface.copy()
        r.render_strips_to_host(rank, world, surface.ctypes.data, pitch)
        mine = np.zeros(h, dtype=bool)
        for s0 in range(rank * 8, h, world * 8):
            mine[s0:s0 + 8] = True
        assert np.array_equal(surface[~mine], before[~mine]), f"rank {rank} wrote outside its strips"
        assert np.array_equal(surface[mine, :w], want[mine]), f"rank {rank}: its strips differ from the reference frame"
        assert np.all(surface[:, w:] == 0xDEADBEEF)
    assert np.array_equal(surface[:, :w], want)
    r.close()


def test_in_process_multi_device_presents(gpu):
    """Several devices in one context, host destination: direct present (default) and the gather through device 0
    (RT_B200_PRESENT=gather is read once per process, so the gather flow is exercised through rt_render_device +
    rt_download_frame).  Odd frame sizes included."""
    torch = gpu
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    for name in ("bunny_333x77", "w4ref_101x203", "bunny_4k"):
        want = load_golden_frame(name)
        r = make_renderer(name, device_ids=list(range(n)))
        assert np.array_equal(r.Render(), want), name
        r.render_device()
        assert np.array_equal(r.download(), want), name
        print(name, "direct present timing", r.ctx.timing())
        r.close()


def test_wavefront_tables_follow_the_scene(gpu):
    """One context, RT_KERNEL_WAVEFRONT, scenes swapped under it: three meshes, then one small mesh, then one deep mesh, and
    back - the split tables are rebuilt for whatever meshes the launch finds (ensure_splits), block by block."""
    r = make_renderer("w4ref_640")
    r.ctx.set_mesh_path(2)
    r.ctx.set_kernel_variant(4)
    for name in ("w4ref_640", "bunny_640", "optional_640", "w4ref_640", "optional_640", "bunny_640"):
        info = MANIFEST[name]
        assert (info["width"], info["height"]) == (640, 480)
        r.SetScene(load_golden_scene(name))
        for frame in range(2):
            identical, max_err, n_diff = compare_frames(r.Render(), load_golden_frame(name))
            assert (n_diff == 0) if name in EXACT else (identical >= MIN_IDENTICAL and max_err <= MAX_LSB), (name, frame, n_diff, max_err)
    r.close()


def test_in_process_multi_device_wavefront(gpu):
    """Several devices in one context with RT_KERNEL_WAVEFRONT: every device renders its strips with the five launches
    and its own copy of the split tables (built once per mesh, copied per device when its first launch wants them);
    a re-uploaded mesh rebuilds them.  Frames against the reference's."""
    torch = gpu
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    for name in ("optional_320", "w4ref_101x203", "bunny_333x77"):
        want = load_golden_frame(name)
        r = make_renderer(name, device_ids=list(range(n)))
        r.ctx.set_mesh_path(2)
        r.ctx.set_kernel_variant(4)
        scene = load_golden_scene(name)
        for frame in range(3):
            if frame == 2:
                r.ctx.upload_mesh_descriptor(0, r.ctx.mesh_descriptor(scene.meshes[0]))      # the same mesh again: new tables
            identical, max_err, n_diff = compare_frames(r.Render(), want)
            assert (n_diff == 0) if name in EXACT else (identical >= MIN_IDENTICAL and max_err <= MAX_LSB), (name, frame, n_diff, max_err)
            r.render_device()
            identical, max_err, n_diff = compare_frames(r.download(), want)
            assert identical >= MIN_IDENTICAL and max_err <= MAX_LSB, (name, frame, n_diff, max_err)
        r.close()


def test_axis_parallel_rays_take_the_literal_slab_test(gpu):
    """Rays with a zero direction component have an infinite 1/dir; on a box face through the ray origin the
    slab products are 0 * inf = NaN and std::min/max semantics decide (reference source/Utils.h:197-215).
    Odd width + axis-aligned camera puts d.x == 0 on the centre column; fov 0 makes every ray (0, 0, 1)."""
    from oracle import rt_oracle
    scene = load_golden_scene("w4ref_101x203")
    for origin, fov in (((-0.75, 4.5, -9.0), scene.camera.fov), ((0.75, 3.0, -9.0), scene.camera.fov),
                        ((-0.75, 5.0, -9.0), 0.0), ((0.0, 4.5, -9.0), 0.0)):
        scene.camera.origin[:] = origin
        scene.camera.fov = fov
        r = make_renderer("w4ref_101x203")
        r.SetScene(scene)
        for gpu_path, oracle_path in ((1, rt_oracle.MESH_SLAB_LINEAR), (2, rt_oracle.MESH_BVH)):
            r.ctx.set_mesh_path(gpu_path)
            want = rt_oracle.render(scene, 101, 203, mesh_path=oracle_path)
            got = r.Render()
            identical, max_err, n_diff = compare_frames(got, want)
            assert n_diff == 0 or (identical >= MIN_IDENTICAL and max_err <= MAX_LSB), (origin, fov, gpu_path, n_diff, max_err)
        r.close()


def test_bounce_buffer_path_when_the_surface_cannot_be_pinned(gpu, tmp_path):
    """RT_B200_FORCE_STAGING makes prepare_host behave as if cudaHostRegister had failed: frames go through the pinned
    bounce buffer and are copied into the caller's surface by the CPU - whole frames (rt_render) and a rank's strips only
    (rt_render_strips_to_host).  The switch is read once per process, hence the child process."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    code = '''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
from conftest import load_golden_frame
from test_gpu_parity import make_renderer
r = make_renderer("bunny_333x77")
want = load_golden_frame("bunny_333x77")
assert np.array_equal(r.Render(), want)
assert np.array_equal(r.Render(), want)
surface = np.full((77, 333), 0xDEADBEEF, dtype=np.uint32)
for rank in range(3):
    r.render_strips_to_host(rank, 3, surface.ctypes.data, 333 * 4)
    rows = np.zeros(77, dtype=bool)
    for s0 in range(rank * 8, 77, 24):
        rows[s0:s0 + 8] = True
    assert np.array_equal(surface[rows], want[rows])
    if rank < 2:
        assert (surface[~rows] == 0xDEADBEEF).any()
assert np.array_equal(surface, want)
r.close()
print("ok")
''' % (ROOT, os.path.join(ROOT, "tests"))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, RT_B200_FORCE_STAGING="1"), timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout + res.stderr


def test_registered_surface_lifetime_and_unpinned_surfaces(gpu):
    """rt_register_surface / rt_unregister_surface (rt_b200.h): a caller-pinned surface is written directly, a pageable
    one goes through the context's bounce buffer - and the library never keeps a registration of memory the caller may
    free (round-1 advisor finding: a fresh 33 MB numpy array per frame is unmapped and mapped again at the same address)."""
    want = load_golden_frame("bunny_4k")
    r = make_renderer("bunny_4k")
    for _ in range(3):
        out = np.empty((2160, 3840), dtype=np.uint32)          # above glibc's mmap threshold: a new mapping every time
        out[...] = 0xDEADBEEF
        assert np.array_equal(r.Render(out), want)
        del out
    surface = np.full((2160, 3840), 0xDEADBEEF, dtype=np.uint32)
    r.register_surface(surface.ctypes.data, surface.nbytes)
    with pytest.raises(Exception):
        r.register_surface(surface.ctypes.data, surface.nbytes)   # twice
    assert np.array_equal(r.Render(surface), want)
    r.unregister_surface(surface.ctypes.data)
    with pytest.raises(Exception):
        r.unregister_surface(surface.ctypes.data)                 # not registered any more
    surface[...] = 0
    assert np.array_equal(r.Render(surface), want)                # pageable again: bounce buffer
    r.close()


def test_clear_frame_poisons_device_frame(gpu):
    r = make_renderer("bunny_333x77")
    r.render_device()
    r.clear_frame(0xDEADBEEF)
    assert (r.download() == 0xDEADBEEF).all()
    r.clear_frame(0x01010101)
    assert (r.download() == 0x01010101).all()
    r.render_device()
    assert np.array_equal(r.download(), load_golden_frame("bunny_333x77"))
    r.close()


def test_uploads_after_a_launch_on_a_foreign_stream_wait_for_it(gpu):
    """A frame launched on a caller-supplied stream, then - without synchronising - a new mesh and the next frame: the
    upload must queue up behind the first kernel (it still reads the old geometry) and the second kernel behind the
    upload (round-1 advisor finding).  Both frames must be what their own geometry renders to."""
    torch = gpu
    a, b = load_golden_scene("bunny_320_yaw05"), load_golden_scene("bunny_320_yaw25")
    want_a, want_b = load_golden_frame("bunny_320_yaw05"), load_golden_frame("bunny_320_yaw25")
    r = make_renderer("bunny_320_yaw05")
    stream = torch.cuda.Stream()
    out_a = torch.zeros((240, 320), dtype=torch.int32, device="cuda")
    out_b = torch.zeros((240, 320), dtype=torch.int32, device="cuda")
    for _ in range(20):
        r.ctx.upload_mesh(0, a.meshes[0])
        r.render_strips_device(0, 1, out_a.data_ptr(), stream.cuda_stream)
        r.ctx.upload_mesh(0, b.meshes[0])                          # no synchronisation in between
        r.render_strips_device(0, 1, out_b.data_ptr(), stream.cuda_stream)
        stream.synchronize()
        assert np.array_equal(out_a.cpu().numpy().view(np.uint32), want_a)
        assert np.array_equal(out_b.cpu().numpy().view(np.uint32), want_b)
    r.close()


def test_wavefront_kernel_in_every_mode_and_on_strips(gpu):
    """RT_KERNEL_WAVEFRONT (rt_kernel_wave.cuh): closest hit merged over subtrees with atomicMin on (t, primitive), any hit
    with atomicOr - bit-identical to the reference frames in all four lighting modes, with shadows off, on an odd-sized
    frame, on a rank's strips and through the progressive present."""
    for name in ("bunny_640_observed", "bunny_640_radiance", "bunny_640_brdf", "bunny_640_noshadow", "bunny_333x77", "w4ref_101x203",
                 "optional_320", "optional_320_steps2", "w4ref_320_cam"):
        r = make_renderer(name)
        r.ctx.set_mesh_path(2)
        r.ctx.set_kernel_variant(4)
        identical, max_err, n_diff = compare_frames(r.Render(), load_golden_frame(name))
        assert (n_diff == 0) if name in EXACT else (identical >= MIN_IDENTICAL and max_err <= MAX_LSB), (name, n_diff, max_err)
        r.render_device()
        identical, max_err, n_diff = compare_frames(r.download(), load_golden_frame(name))
        assert identical >= MIN_IDENTICAL and max_err <= MAX_LSB, (name, n_diff, max_err)
        r.close()
    # strips of three "ranks" into one host surface
    r = make_renderer("optional_320")
    r.ctx.set_kernel_variant(4)
    want = load_golden_frame("optional_320")
    surface = np.zeros((240, 320), dtype=np.uint32)
    for rank in range(3):
        r.render_strips_to_host(rank, 3, surface.ctypes.data, 320 * 4)
    identical, max_err, n_diff = compare_frames(surface, want)
    assert identical >= MIN_IDENTICAL and max_err <= MAX_LSB, (n_diff, max_err)
    r.close()


@pytest.mark.parametrize("knobs", [{"RT_B200_WAVE_PARTS_BELOW": "0"}, {"RT_B200_WAVE_PARTS_BELOW": "1000000"},
                                   {"RT_B200_WAVE_PARTS_BELOW": "1000000", "RT_B200_WAVE_PARTS": "2", "RT_B200_WAVE_SUBTREES": "7"},
                                   {"RT_B200_WAVE_NO_CHAIN": "1"}])
def test_wavefront_units_whole_subtrees_and_parts(gpu, knobs):
    """The walk kernels of RT_KERNEL_WAVEFRONT take whole subtrees (from global memory) or parts of subtrees (copied into
    shared memory, after the boxes between the subtree's root and the part) as their units; the host picks per launch.
    Forced either way, with an odd cut of the tree, and without programmatic dependent launch: the reference's frames
    all the same.  The knobs are read once per process, hence the child process."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    code = '''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
from conftest import MAX_LSB, MIN_IDENTICAL, compare_frames, load_golden_frame
from test_gpu_parity import EXACT, make_renderer
for name in ("optional_320", "optional_320_steps2", "w4ref_101x203", "bunny_333x77", "bunny_640_noshadow", "w4ref_320_cam"):
    r = make_renderer(name)
    r.ctx.set_mesh_path(2)
    r.ctx.set_kernel_variant(4)
    for frame in range(3):          # the second frame on, the host knows the job counts of the one before
        identical, max_err, n_diff = compare_frames(r.Render(), load_golden_frame(name))
        assert (n_diff == 0) if name in EXACT else (identical >= MIN_IDENTICAL and max_err <= MAX_LSB), (name, frame, n_diff, max_err)
    assert r.ctx.timing()["kernel_launches"] in (3, 5), r.ctx.timing()          # K1 K2 [K3 K4] K5
    r.close()
print("ok")
''' % (ROOT, os.path.join(ROOT, "tests"))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, **knobs), timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout + res.stderr
