#!/usr/bin/env python
"""NVLink evidence for the fused render + gather: ONE process, one context over devices 0 and 1,
gather flow (RT_B200_PRESENT=gather is not needed: rt_render_device always stores every device's strips into device 0's
frame).  Device 1's pixel kernel writes its half of the 4K frame (16.6 MB) straight into device 0's memory with 128-bit
peer stores.  Prints the NVLink data counters of both GPUs (nvidia-smi nvlink -gt d) around N frames; under ncu
(--devices 1) the same launches give nvltx / aperture_peer counters per launch.
Usage: python tools/r02_nvlink_probe.py [frames]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from conftest import load_golden_frame, load_golden_scene  # noqa: E402
from gp1_raytracer_2223_b200 import Renderer  # noqa: E402


def nvlink_kib(gpu):
    """Sum over links of (rx KiB, tx KiB) of one GPU."""
    out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu)], capture_output=True, text=True).stdout
    rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
    tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
    return rx, tx, out


frames = int(sys.argv[1]) if len(sys.argv) > 1 else 200
r = Renderer(3840, 2160, device_ids=[0, 1])
r.SetScene(load_golden_scene("bunny_4k"))
for _ in range(3):
    r.render_device()
before = [nvlink_kib(g) for g in (0, 1)]
ms = [r.render_device()["kernel_ms"] for _ in range(frames)]
after = [nvlink_kib(g) for g in (0, 1)]
diff = int((r.download() != load_golden_frame("bunny_4k")).sum())
print(f"frames {frames}  kernel ms (max over the two devices) mean {np.mean(ms):.4f}  differing pixels vs reference frame {diff}")
half = 3840 * 2160 * 4 / 2
for g in (0, 1):
    drx, dtx = after[g][0] - before[g][0], after[g][1] - before[g][1]
    print(f"GPU {g}: NVLink data rx {drx} KiB ({drx * 1024 / frames / 1e6:.2f} MB / frame), tx {dtx} KiB ({dtx * 1024 / frames / 1e6:.2f} MB / frame); "
          f"half a 4K frame = {half / 1e6:.2f} MB")
if not any(after[g][0] or after[g][1] for g in (0, 1)):
    print("nvidia-smi reports no NVLink data counters on this box; raw output follows\n" + after[1][2][:600])
r.close()
