#!/usr/bin/env python
"""What can the device-to-host leg of the present do on this box?    One process, N GPUs, CUDA runtime
through ctypes: every GPU copies its share of a 3840x2160 XRGB frame into ONE pinned host surface, all GPUs at once.
  contiguous   one cudaMemcpyAsync of the share (as if strips were packed on both sides)
  strips       one cudaMemcpy2DAsync: "rows" of one 8-row strip (122 880 B), destination stride = N strips (what
               rt_render_strips_to_host issues per band today)
  strips/16    the same in 16 bands (16 calls per GPU)
Prints ms for the slowest GPU and the aggregate GB/s.  Usage: python tools/r02_d2h_probe.py [max_gpus]"""
import ctypes as C
import glob
import os
import sys
import time

import torch  # noqa: F401  (loads libcudart)

rt = None
for pat in ("libcudart.so*",):
    for p in glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", pat)) + glob.glob("/usr/local/cuda/lib64/" + pat):
        try:
            rt = C.CDLL(p)
            break
        except OSError:
            pass
    if rt:
        break
assert rt, "libcudart not found"
W, H, STRIP = 3840, 2160, 8
FRAME = W * H * 4
ROWB = W * 4


def chk(e):
    assert e == 0, f"CUDA error {e}"


n_max = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
host = C.c_void_p()
chk(rt.cudaHostAlloc(C.byref(host), C.c_size_t(FRAME), C.c_uint(1)))      # portable
C.memset(host, 0, FRAME)
for n in [g for g in (1, 2, 4, 8) if g <= n_max]:
    devs = []
    for g in range(n):
        chk(rt.cudaSetDevice(g))
        buf, stream, e0, e1 = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        chk(rt.cudaMalloc(C.byref(buf), C.c_size_t(FRAME)))
        chk(rt.cudaStreamCreateWithFlags(C.byref(stream), C.c_uint(1)))
        chk(rt.cudaEventCreate(C.byref(e0))); chk(rt.cudaEventCreate(C.byref(e1)))
        devs.append((buf, stream, e0, e1))
    total_strips = H // STRIP
    share = FRAME // n

    def run(mode):
        best = 1e9
        for rep in range(6):
            t0 = time.perf_counter()
            for g, (buf, stream, e0, e1) in enumerate(devs):
                chk(rt.cudaSetDevice(g))
                chk(rt.cudaEventRecord(e0, stream))
                if mode == "contiguous":
                    chk(rt.cudaMemcpyAsync(C.c_void_p(host.value + g * share), buf, C.c_size_t(share), C.c_int(2), stream))
                else:
                    bands = 1 if mode == "strips" else 16
                    mine = total_strips // n
                    per = (mine + bands - 1) // bands
                    for b in range(bands):
                        s0, s1 = b * per, min(mine, (b + 1) * per)
                        if s1 <= s0:
                            break
                        first = g + s0 * n                      # frame strip index
                        chk(rt.cudaMemcpy2DAsync(C.c_void_p(host.value + first * STRIP * ROWB), C.c_size_t(n * STRIP * ROWB),
                                                 C.c_void_p(buf.value + first * STRIP * ROWB), C.c_size_t(n * STRIP * ROWB),
                                                 C.c_size_t(STRIP * ROWB), C.c_size_t(s1 - s0), C.c_int(2), stream))
                chk(rt.cudaEventRecord(e1, stream))
            worst = 0.0
            for g, (buf, stream, e0, e1) in enumerate(devs):
                chk(rt.cudaSetDevice(g))
                chk(rt.cudaStreamSynchronize(stream))
                ms = C.c_float()
                chk(rt.cudaEventElapsedTime(C.byref(ms), e0, e1))
                worst = max(worst, ms.value)
            wall = (time.perf_counter() - t0) * 1e3
            if rep > 0:
                best = min(best, max(worst, 0.0))
        return best, wall

    for mode in ("contiguous", "strips", "strips/16"):
        ms, wall = run(mode)
        print(f"{n} GPU(s) {mode:11s}: slowest GPU {ms:.3f} ms for {share / 1e6:.2f} MB -> {share / ms / 1e6:.1f} GB/s per GPU, {FRAME / ms / 1e6:.1f} GB/s aggregate (host wall of the last repetition {wall:.3f} ms)")
    for g, (buf, stream, e0, e1) in enumerate(devs):
        chk(rt.cudaSetDevice(g)); chk(rt.cudaFree(buf)); chk(rt.cudaStreamDestroy(stream))
