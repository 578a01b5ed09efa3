"""TEST INFRASTRUCTURE: ctypes front end of oracle/port/librt_oracle.so (the CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from gp1_raytracer_2223_b200._abi import (SceneViews, camera_struct, frame_struct, rt_camera, rt_counters,
                                          rt_frame_desc, rt_lights_soa, rt_material_desc, rt_mesh_desc,
                                          rt_planes_soa, rt_spheres_soa)
from gp1_raytracer_2223_b200.scene_file import FlatScene

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "port", "librt_oracle.so")

MESH_SLAB_LINEAR = 0
MESH_BVH = 1


class rto_mesh(C.Structure):
    _fields_ = [("desc", rt_mesh_desc)]


class rto_scene(C.Structure):
    _fields_ = [("spheres", rt_spheres_soa), ("planes", rt_planes_soa), ("lights", rt_lights_soa),
                ("materials", C.POINTER(rt_material_desc)), ("material_count", C.c_int32),
                ("meshes", C.POINTER(rto_mesh)), ("mesh_count", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc only; does not need /root/reference)."""
    src = os.path.join(HERE, "port", "rt_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "port"], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.rto_render_rows.restype = C.c_int
        _lib.rto_render_rows.argtypes = [C.POINTER(rto_scene), C.POINTER(rt_camera), C.POINTER(rt_frame_desc),
                                         C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                         C.POINTER(rt_counters)]
        _lib.rto_transform_mesh.restype = None
        _lib.rto_transform_mesh.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.rto_update_transforms_bvh.restype = C.c_int32
        _lib.rto_update_transforms_bvh.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        _lib.rto_fnv1a64.restype = C.c_uint64
        _lib.rto_fnv1a64.argtypes = [C.c_void_p, C.c_uint64]
    return _lib


def fnv1a64(buf: np.ndarray) -> int:
    a = np.ascontiguousarray(buf)
    return int(lib().rto_fnv1a64(a.ctypes.data, a.nbytes))


class OracleScene:
    def __init__(self, scene: FlatScene):
        self.views = SceneViews(scene)
        n = len(scene.meshes)
        self._meshes = (rto_mesh * max(n, 1))()
        self.has_bvh = n > 0 and all(m.bvh_nodes is not None and len(m.bvh_nodes) for m in scene.meshes)
        for i in range(n):
            self._meshes[i].desc = self.views.meshes[i]
        self.c = rto_scene(self.views.spheres, self.views.planes, self.views.lights,
                           C.cast(self.views.materials, C.POINTER(rt_material_desc)), self.views.material_count,
                           C.cast(self._meshes, C.POINTER(rto_mesh)), n)


def render(scene: FlatScene, width: int, height: int, lighting_mode: int = 3, shadows: bool = True,
           mesh_path: int = MESH_SLAB_LINEAR, row_begin: int = 0, row_count: int | None = None, threads: int = 0,
           counters: bool = False, camera=None, aspect_ratio: float | None = None, shifts=(16, 8, 0),
           alpha_mask: int = 0):
    """Render rows with the CPU restatement.  Returns uint32 (rows, width) [and counters]."""
    if row_count is None:
        row_count = height - row_begin
    osc = OracleScene(scene)
    cam = camera_struct(camera if camera is not None else scene.camera)
    frame = frame_struct(width, height, lighting_mode, shadows, aspect_ratio, shifts, alpha_mask)
    out = np.empty((row_count, width), dtype=np.uint32)
    cnt = rt_counters() if counters else None
    rc = lib().rto_render_rows(C.byref(osc.c), C.byref(cam), C.byref(frame), mesh_path, row_begin, row_count,
                               out.ctypes.data, threads, C.byref(cnt) if counters else None)
    if rc != 0:
        raise RuntimeError(f"rto_render_rows failed ({rc})")
    if counters:
        return out, np.array(list(cnt.slot), dtype=np.uint64)
    return out


def transform_mesh(positions: np.ndarray, normals: np.ndarray, transform: np.ndarray):
    """CPU restatement of TransformPoint / TransformVector().Normalized() over a mesh.  Returns (positions, normals)."""
    pos = np.ascontiguousarray(positions, dtype=np.float32)
    nrm = np.ascontiguousarray(normals, dtype=np.float32)
    m = np.ascontiguousarray(transform, dtype=np.float32).reshape(16)
    out_p, out_n = np.empty_like(pos), np.empty_like(nrm)
    lib().rto_transform_mesh(pos.ctypes.data, pos.shape[0], nrm.ctypes.data, nrm.shape[0], m.ctypes.data, out_p.ctypes.data, out_n.ctypes.data)
    return out_p, out_n


def update_transforms_bvh(positions: np.ndarray, indices: np.ndarray, normals: np.ndarray, transform: np.ndarray):
    """CPU restatement of TriangleMesh::UpdateTransforms with BuildBVH (reference source/DataTypes.h:210-236, 294-483).
    `indices` (T, 3) int32 and `normals` (T, 3) float32 are REORDERED IN PLACE like the reference's members (they
    must be C-contiguous arrays of exactly those dtypes).  Returns (transformed positions, transformed normals in
    the new order, BVH nodes as a BVH_NODE_DTYPE array)."""
    from gp1_raytracer_2223_b200.scene_file import BVH_NODE_DTYPE
    pos = np.ascontiguousarray(positions, dtype=np.float32)
    assert indices.dtype == np.int32 and indices.flags.c_contiguous and normals.dtype == np.float32 and normals.flags.c_contiguous
    m = np.ascontiguousarray(transform, dtype=np.float32).reshape(16)
    n_t = indices.shape[0]
    out_p, out_n = np.empty_like(pos), np.empty_like(normals)
    nodes = np.zeros(max(2 * n_t, 1), dtype=BVH_NODE_DTYPE)
    used = lib().rto_update_transforms_bvh(pos.ctypes.data, pos.shape[0], indices.ctypes.data, normals.ctypes.data, n_t,
                                           m.ctypes.data, out_p.ctypes.data, out_n.ctypes.data, nodes.ctypes.data, len(nodes))
    if used < 0:
        raise RuntimeError("rto_update_transforms_bvh: node array too small")
    return out_p, out_n, nodes[:used].copy()
