#!/usr/bin/env python
"""Regenerates tests/golden/ from the UNMODIFIED reference compiled by oracle/Makefile.

Run in the build container (needs /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

For every case below it runs oracle/_ref/ref_render (the reference's own Scene::Initialize /
Scene::Update / Renderer::Render) and stores
  <case>.rtsc      the flattened scene RenderPixel saw (input fixture)
  <case>.frame.xz  the frame the reference rendered: LZMA of planar R, G, B bytes (output fixture)
and manifest.json with the command line, size, FNV-1a-64 of the raw XRGB8888 frame.
The reference ships no golden vectors; these are the pinning vectors.
"""
import json
import lzma
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_render")

S, L = (640, 480), (3840, 2160)
SMALL = (320, 240)

# name -> (scene, (w, h), mode, shadows, extra args)
CASES = {
    # BASELINE.json configs[0..3] at pose P0
    "w1_640": ("W1", S, 3, 0, []),
    "w2_640": ("W2", S, 3, 1, []),
    "w3_640": ("W3", S, 3, 1, []),
    "w3test_640": ("W3_Test", S, 3, 1, []),
    "w4ref_640": ("W4_Reference", S, 3, 1, []),
    "bunny_640": ("W4_Bunny", S, 3, 1, []),
    # the other lighting modes / shadows off on the bunny scene
    "bunny_640_observed": ("W4_Bunny", S, 0, 1, []),
    "bunny_640_radiance": ("W4_Bunny", S, 1, 1, []),
    "bunny_640_brdf": ("W4_Bunny", S, 2, 1, []),
    "bunny_640_noshadow": ("W4_Bunny", S, 3, 0, []),
    # Cook-Torrance / Phong in every mode
    "w3_320_observed": ("W3", SMALL, 0, 1, []),
    "w3_320_radiance": ("W3", SMALL, 1, 1, []),
    "w3_320_brdf": ("W3", SMALL, 2, 1, []),
    "w3_320_noshadow": ("W3", SMALL, 3, 0, []),
    "w3test_320_brdf": ("W3_Test", SMALL, 2, 1, []),
    # animated / moved poses through the reference's own Update, RotateY and camera code
    "bunny_320_yaw05": ("W4_Bunny", SMALL, 3, 1, ["--mesh-yaw", "0.5"]),
    "bunny_320_yaw10": ("W4_Bunny", SMALL, 3, 1, ["--mesh-yaw", "1.0"]),
    "bunny_320_yaw25": ("W4_Bunny", SMALL, 3, 1, ["--mesh-yaw", "2.5"]),
    "bunny_320_yaw40": ("W4_Bunny", SMALL, 3, 1, ["--mesh-yaw", "4.0"]),
    "bunny_320_time2": ("W4_Bunny", SMALL, 3, 1, ["--time", "2.0"]),
    "bunny_320_cam": ("W4_Bunny", SMALL, 3, 1, ["--cam-origin", "2.5", "4.0", "-7.0", "--cam-rot", "0.25", "-0.35"]),
    "w4ref_320_time13": ("W4_Reference", SMALL, 3, 1, ["--time", "1.3"]),
    "w4ref_320_cam": ("W4_Reference", SMALL, 3, 1, ["--time", "0.7", "--cam-origin", "-3.0", "5.0", "-6.0", "--cam-rot", "0.3", "0.4"]),
    "w3_320_fov90": ("W3", SMALL, 3, 1, ["--fov", "90", "--cam-origin", "1.0", "2.0", "-4.0", "--cam-rot", "-0.1", "-0.2"]),
    # odd sizes: ragged tiles, width not a multiple of 4 (scalar store path)
    "bunny_333x77": ("W4_Bunny", (333, 77), 3, 1, []),
    "w4ref_101x203": ("W4_Reference", (101, 203), 3, 1, []),
    # next row N3: 3082-triangle mesh, Cook-Torrance
    "optional_320": ("W4_Optional", SMALL, 3, 1, []),
    # sequences of UpdateTransforms + BuildBVH calls (every build starts from the triangle order the previous one
    # left): pin the BuildBVH restatement and the device-side build
    "bunny_320_steps3": ("W4_Bunny", SMALL, 3, 1, ["--yaw-steps", "0.5,1.0,1.7"]),
    "w4ref_320_steps2": ("W4_Reference", SMALL, 3, 1, ["--yaw-steps", "0.7,2.1"]),
    "optional_320_steps2": ("W4_Optional", SMALL, 3, 1, ["--yaw-steps", "0.3,0.9"]),
    # BASELINE.json configs[4]: the headline frame
    "bunny_4k": ("W4_Bunny", L, 3, 1, []),
    # the frames SURVEY.md 0.4 / 8(c) found most sensitive to a single contracted operation (shadow-terminator pixels,
    # Renderer.cpp:137-141) at the headline size, plus the spheres + planes + SolidColor check of config 0
    "w2_4k": ("W2", L, 3, 1, []),
    "w3_4k": ("W3", L, 3, 1, []),
    "w4ref_4k": ("W4_Reference", L, 3, 1, []),
    # N3 at the BASELINE sizes
    "optional_640": ("W4_Optional", S, 3, 1, []),
    "optional_4k": ("W4_Optional", L, 3, 1, []),
}


# cases that also get <case>.rtms: the untransformed meshes + finalTransform (input of the device-side
# TriangleMesh::UpdateTransforms)
# cases that also get <case>.rtmp: the meshes before the --yaw-steps sequence + the transform of every step
MESH_STEPS = {"bunny_320_steps3", "w4ref_320_steps2", "optional_320_steps2"}

MESH_SOURCES = {"bunny_320_yaw05", "bunny_320_yaw10", "bunny_320_time2", "w4ref_320_time13", "optional_320"}


def pack_frame(frame_u32: np.ndarray) -> bytes:
    b = frame_u32.view(np.uint8).reshape(-1, 4)          # little-endian XRGB8888: B, G, R, X
    assert not b[:, 3].any(), "X byte must be 0 for XRGB8888"
    planar = np.ascontiguousarray(b[:, [2, 1, 0]].T)     # R plane, G plane, B plane
    return lzma.compress(planar.tobytes(), preset=9 | lzma.PRESET_EXTREME)


def main():
    if not os.path.exists(REF):
        sys.exit("oracle/_ref/ref_render missing: run `make -C oracle ref` first")
    manifest = {}
    only = set(sys.argv[1:])            # `make_golden.py case ...` regenerates just those cases
    if only:
        with open(os.path.join(HERE, "manifest.json")) as f:
            manifest.update(json.load(f))
    for name, (scene, (w, h), mode, shadows, extra) in CASES.items():
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            raw = os.path.join(tmp, "frame.bin")
            rtsc = os.path.join(HERE, name + ".rtsc")
            args = ["--scene", scene, "--width", str(w), "--height", str(h), "--mode", str(mode),
                    "--shadows", str(shadows)] + extra
            extra_out = ["--dump-mesh-source", os.path.join(HERE, name + ".rtms")] if name in MESH_SOURCES else []
            if name in MESH_STEPS:
                extra_out += ["--dump-mesh-steps", os.path.join(HERE, name + ".rtmp")]
            out = subprocess.run([REF] + args + ["--out", raw, "--dump-scene", rtsc] + extra_out, check=True,
                                 capture_output=True, text=True).stdout
            info = json.loads(out)
            frame = np.fromfile(raw, dtype=np.uint32)
            assert frame.size == w * h
            with open(os.path.join(HERE, name + ".frame.xz"), "wb") as f:
                f.write(pack_frame(frame))
        manifest[name] = {"scene": scene, "width": w, "height": h, "mode": mode, "shadows": shadows,
                          "ref_render_args": args, "fnv1a64": info["fnv1a64"]}
        print(name, info["fnv1a64"], os.path.getsize(os.path.join(HERE, name + ".frame.xz")))
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    main()
