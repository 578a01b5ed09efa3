"""Builds the CUDA library in-tree: gp1_raytracer_2223_b200/librt_b200.so (sm_100a only).

nvcc cross-compiles without a GPU.  Flags that matter for parity:
``--fmad=false`` (belt and braces: the kernels already spell every FP32 operation with
non-contractable ``__f*_rn`` intrinsics), IEEE division / square root (nvcc defaults; never
``--use_fast_math``).
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librt_b200.so")
# One translation unit per family of kernel instantiations: ptxas is the long pole (dozens of template instantiations
# of a 6 000-instruction kernel), so the units are compiled in parallel and linked once.  No relocatable device code:
# every kernel is complete inside its unit; the units only exchange host function pointers (rt_pick.h).
SOURCES = ["rt_api.cu", "rt_pick_tiled.cu", "rt_pick_persistent.cu", "rt_pick_x2.cu", "rt_pick_wave.cu"]
HEADERS = ["rt_kernel.cuh", "rt_kernel_x2.cuh", "rt_device.cuh", "rt_bvh_build.cuh", "rt_aux_kernels.cuh", "rt_kernel_wave.cuh", "rt_wave_params.h", "rt_pick.h",
           os.path.join("..", "..", "include", "rt_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]
OBJ_DIR = os.path.join(os.path.dirname(HERE), "build", "obj")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


HASH_FILE = LIB + ".hash"
LOCK_FILE = LIB + ".lock"


def _extra_flags():
    return os.environ.get("RT_B200_NVCC_EXTRA", "").split()


def source_hash() -> str:
    """sha256 over every source the library is compiled from plus the nvcc flags: what a given .so must have been
    built from.  Stored next to the library (librt_b200.so.hash) by build(); ensure() rebuilds when it differs, so a
    stale library can neither be tested nor benchmarked after an edit (file times mean nothing on the GPU box)."""
    h = hashlib.sha256()
    for rel in sorted(SOURCES + HEADERS):
        path = os.path.normpath(os.path.join(CSRC, rel))
        h.update(os.path.basename(path).encode() + b"\0")
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + _extra_flags()).encode())
    return h.hexdigest()


def built_hash():
    try:
        with open(HASH_FILE) as f:
            return f.read().strip()
    except OSError:
        return None


def needs_build() -> bool:
    return not os.path.exists(LIB) or built_hash() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    want = source_hash()
    tmp = LIB + f".tmp{os.getpid()}"
    os.makedirs(OBJ_DIR, exist_ok=True)
    # nvcc's default host compiler must be the system g++ (a CXX override in the environment
    # points at a toolchain without all runtime pieces)
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    only = set(os.environ.get("RT_B200_BUILD_ONLY", "").split())       # dev loop: recompile just these units, reuse the other objects

    def compile_unit(src):
        obj = os.path.join(OBJ_DIR, f"{os.path.splitext(src)[0]}.{os.getpid()}.o")
        cmd = [_nvcc()] + NVCC_FLAGS + _extra_flags() + ccbin + ["-c", "-o", obj, os.path.join(CSRC, src)]
        keep = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        if only and src not in only and os.path.exists(keep):
            return src, keep, 0, "(reused)\n", cmd
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode == 0:
            os.replace(obj, keep)
        return src, keep, proc.returncode, proc.stdout + proc.stderr, cmd

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_unit, SOURCES))
    log = ""
    for src, obj, rc, out, cmd in results:
        log += " ".join(cmd) + "\n" + out
    failed = [src for src, _, rc, _, _ in results if rc != 0]
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + ccbin + ["-o", tmp] + [obj for _, obj, _, _, _ in results]
    if not failed:
        proc = subprocess.run(link, capture_output=True, text=True)
        log += " ".join(link) + "\n" + proc.stdout + proc.stderr
        if proc.returncode != 0:
            failed = ["link"]
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(log)
    if failed:
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise RuntimeError(f"nvcc failed ({failed}):\n" + log[-6000:])
    os.replace(tmp, LIB)                      # readers never see a half-written library
    with open(HASH_FILE, "w") as f:
        f.write(want + "\n")
    if verbose:
        print(log)
    return LIB


def ensure() -> str:
    """The library, guaranteed to be built from the sources in the tree: rebuilds when it is missing or when the hash
    stored beside it differs from source_hash().  Used by bench.py and the tests.  N ranks starting together serialise
    on a lock file, and whoever gets the lock second finds the work done."""
    if not needs_build():
        return LIB
    with open(LOCK_FILE, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if needs_build():
                build(force=True)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


def provenance() -> dict:
    """What bench.py prints about the library it measured."""
    from . import _lib
    override = os.environ.get("RT_B200_LIB")
    return {"library": os.path.relpath(_lib.LIB_PATH, os.path.dirname(HERE)), "source_sha256_16": source_hash()[:16],
            "built_from_sha256_16": (built_hash() or "unknown")[:16] if not override else "unchecked (RT_B200_LIB override)",
            "fresh": (not override) and built_hash() == source_hash()}


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
