"""The drop-in itself: the reference's own Scene / Camera / Material / Timer sources and its own
Renderer.h, linked once with the reference's Renderer.cpp (oracle/_ref/ref_render) and once with
gp1_raytracer_2223_b200/host/Renderer.cpp -> librt_b200.so (oracle/_ref/ref_render_b200).
Same command line, same scene code, frames must meet the parity bar (bit-exact without powf)."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import MAX_LSB, MIN_IDENTICAL, ROOT, compare_frames

pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "oracle", "_ref", "ref_render")
DROPIN = os.path.join(ROOT, "oracle", "_ref", "ref_render_b200")

CASES = [
    (["--scene", "W4_Bunny", "--width", "640", "--height", "480"], True),
    (["--scene", "W4_Bunny", "--width", "322", "--height", "181", "--time", "3.7", "--mode", "0"], True),
    (["--scene", "W4_Bunny", "--width", "320", "--height", "240", "--mesh-yaw", "1.9", "--cam-origin", "1", "2.5", "-8", "--cam-rot", "0.1", "-0.15"], True),
    (["--scene", "W4_Reference", "--width", "640", "--height", "480", "--time", "0.9"], False),
    (["--scene", "W3", "--width", "640", "--height", "480"], False),
    (["--scene", "W3_Test", "--width", "400", "--height", "300", "--mode", "2"], False),
    (["--scene", "W2", "--width", "640", "--height", "480", "--shadows", "0"], True),
    (["--scene", "W1", "--width", "64", "--height", "48", "--shadows", "0"], True),
    (["--scene", "W4_Optional", "--width", "320", "--height", "240", "--time", "1.1"], False),
]


def run(binary, args, out, env=None):
    res = subprocess.run([binary] + args + ["--out", out], check=True, capture_output=True, text=True,
                         env=dict(os.environ, **(env or {})))
    assert "rt_b200:" not in res.stderr, res.stderr
    return json.loads(res.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("args,exact", CASES, ids=[" ".join(c[0][1:2] + c[0][6:]) for c in CASES])
def test_dropin_renders_what_the_reference_renders(tmp_path, args, exact):
    if not (os.path.exists(REF) and os.path.exists(DROPIN)):
        pytest.skip("oracle/_ref binaries are not built (they need /root/reference at build time)")
    a, b = str(tmp_path / "ref.bin"), str(tmp_path / "b200.bin")
    info_ref = run(REF, args, a)
    info_gpu = run(DROPIN, args, b)
    w, h = info_ref["width"], info_ref["height"]
    want = np.fromfile(a, dtype=np.uint32).reshape(h, w)
    got = np.fromfile(b, dtype=np.uint32).reshape(h, w)
    identical, max_err, n_diff = compare_frames(got, want)
    if exact:
        assert info_ref["fnv1a64"] == info_gpu["fnv1a64"] and n_diff == 0
    assert identical >= MIN_IDENTICAL and max_err <= MAX_LSB, (n_diff, max_err)
    assert "B200" in info_gpu["path"]


@pytest.mark.parametrize("args", [c[0] for c in CASES if c[0][1].startswith("W4")], ids=[" ".join(c[0][1:2] + c[0][6:]) for c in CASES if c[0][1].startswith("W4")])
def test_dropin_with_device_side_update_transforms(tmp_path, args):
    """RT_B200_DEVICE_TRANSFORM=1: the drop-in sends the untransformed mesh once and finalTransform per frame
    (SURVEY.md 8(f) N1); frames must still be the reference's."""
    if not (os.path.exists(REF) and os.path.exists(DROPIN)):
        pytest.skip("oracle/_ref binaries are not built (they need /root/reference at build time)")
    a, b = str(tmp_path / "ref.bin"), str(tmp_path / "b200.bin")
    info_ref = run(REF, args, a)
    run(DROPIN, args + ["--frames", "3"], b, env={"RT_B200_DEVICE_TRANSFORM": "1"})
    w, h = info_ref["width"], info_ref["height"]
    want = np.fromfile(a, dtype=np.uint32).reshape(h, w)
    got = np.fromfile(b, dtype=np.uint32).reshape(h, w)
    identical, max_err, n_diff = compare_frames(got, want)
    if "W4_Bunny" in args:
        assert n_diff == 0
    assert identical >= MIN_IDENTICAL and max_err <= MAX_LSB, (n_diff, max_err)


STEP_RUNS = [
    ["--scene", "W4_Bunny", "--width", "320", "--height", "240", "--yaw-steps", "0.4,1.3,2.9"],
    ["--scene", "W4_Reference", "--width", "320", "--height", "240", "--yaw-steps", "0.7,2.2"],
    ["--scene", "W4_Optional", "--width", "320", "--height", "240", "--yaw-steps", "0.5,1.0"],
]


@pytest.mark.parametrize("args", STEP_RUNS, ids=[a[1] for a in STEP_RUNS])
def test_dropin_with_device_side_bvh_builds(tmp_path, args):
    """RT_B200_DEVICE_TRANSFORM=2: UpdateTransforms WITH BuildBVH on the device.  The reference poses the meshes and
    runs UpdateTransforms on the host after every pose (--yaw-steps); the drop-in's host only poses them and renders
    (--steps-on-device), every change of pose being one build on the device.  Same frame at the end - also when more
    frames follow without a new pose (no further build may run)."""
    if not (os.path.exists(REF) and os.path.exists(DROPIN)):
        pytest.skip("oracle/_ref binaries are not built (they need /root/reference at build time)")
    a, b = str(tmp_path / "ref.bin"), str(tmp_path / "b200.bin")
    info_ref = run(REF, args, a)
    info_gpu = run(DROPIN, args + ["--steps-on-device", "--frames", "3"], b, env={"RT_B200_DEVICE_TRANSFORM": "2"})
    w, h = info_ref["width"], info_ref["height"]
    want = np.fromfile(a, dtype=np.uint32).reshape(h, w)
    got = np.fromfile(b, dtype=np.uint32).reshape(h, w)
    identical, max_err, n_diff = compare_frames(got, want)
    if "W4_Bunny" in args:
        assert info_ref["fnv1a64"] == info_gpu["fnv1a64"] and n_diff == 0
    assert identical >= MIN_IDENTICAL and max_err <= MAX_LSB, (n_diff, max_err)
