"""Builds the CUDA library in-tree: gp1_raytracer_2223_b200/librt_b200.so (sm_100a only).

nvcc cross-compiles without a GPU.  Flags that matter for parity:
``--fmad=false`` (belt and braces: the kernels already spell every FP32 operation with
non-contractable ``__f*_rn`` intrinsics), IEEE division / square root (nvcc defaults; never
``--use_fast_math``).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librt_b200.so")
SOURCES = ["rt_api.cu"]
HEADERS = ["rt_kernel.cuh", "rt_kernel_x2.cuh", "rt_device.cuh", "rt_bvh_build.cuh", os.path.join("..", "..", "include", "rt_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [__file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get("RT_B200_NVCC_EXTRA", "").split() + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    # nvcc's default host compiler must be the system g++ (a CXX override in the environment
    # points at a toolchain without all runtime pieces)
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log)
    if verbose:
        print(log)
    return LIB


def ensure() -> str:
    """Build only when the library is missing.  Used by bench.py and the tests: on the GPU box the snapshot's
    file times say nothing about freshness, and N ranks must not start N concurrent nvcc runs over one file.
    (`__graft_entry__.build()` / `python -m gp1_raytracer_2223_b200.build` are what rebuild after an edit.)"""
    if os.path.exists(LIB):
        return LIB
    return build(force=True)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
