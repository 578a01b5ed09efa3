"""Flattened scene files ("RTSC0001") and the SoA view the C ABI takes.

An RTSC file is what ``oracle/ref_driver.cpp --dump-scene`` writes after running the
reference's own ``Scene::Initialize`` / ``Scene::Update``: exactly the data
``Renderer::RenderPixel`` reads (reference source/Renderer.cpp:100-182,
source/Scene.cpp:29-96).  Little-endian, int32 / float32 throughout:

    char[8]  "RTSC0001"
    i32      width, height, lighting_mode, shadows_enabled
    f32      aspect_ratio
    f32[13]  camera: origin xyz, fov, right xyz, up xyz, forward xyz
    i32      n_spheres, n_planes, n_lights, n_materials, n_meshes
    spheres    n x { f32 origin xyz, f32 radius, i32 material }
    planes     n x { f32 origin xyz, f32 normal xyz, i32 material }
    lights     n x { f32 origin xyz, f32 direction xyz, f32 color rgb, f32 intensity, i32 type }
    materials  n x { i32 tag, f32 color rgb, f32 p0, p1, p2, reserved }
    meshes     n x { i32 n_vertices, n_triangles, cull_mode, material, n_bvh_nodes,
                     f32[3*n_vertices] transformedPositions, i32[3*n_triangles] indices,
                     f32[3*n_triangles] transformedNormals,
                     n_bvh_nodes x { f32 min xyz, f32 max xyz, u32 firstIdx, idxCount, leftNode } }

This module is host-side plumbing (numpy only): it never renders anything.
"""
from __future__ import annotations

import dataclasses
import struct
from typing import List, Optional

import numpy as np

MAGIC = b"RTSC0001"

LIGHTING_OBSERVED_AREA, LIGHTING_RADIANCE, LIGHTING_BRDF, LIGHTING_COMBINED = 0, 1, 2, 3
CULL_FRONT_FACE, CULL_BACK_FACE, CULL_NONE = 0, 1, 2
MATERIAL_SOLID_COLOR, MATERIAL_LAMBERT, MATERIAL_LAMBERT_PHONG, MATERIAL_COOK_TORRENCE = 0, 1, 2, 3

MATERIAL_DTYPE = np.dtype([("tag", "<i4"), ("color", "<f4", (3,)), ("p0", "<f4"), ("p1", "<f4"),
                           ("p2", "<f4"), ("reserved", "<f4")])
BVH_NODE_DTYPE = np.dtype([("min_aabb", "<f4", (3,)), ("max_aabb", "<f4", (3,)), ("first_idx", "<u4"),
                           ("idx_count", "<u4"), ("left_node", "<u4")])


@dataclasses.dataclass
class Camera:
    """The 13 floats RenderPixel reads from dae::Camera (reference source/Camera.h:24-40)."""
    origin: np.ndarray
    fov: float
    right: np.ndarray
    up: np.ndarray
    forward: np.ndarray


@dataclasses.dataclass
class Mesh:
    """One TriangleMesh after UpdateTransforms (reference source/DataTypes.h:210-236)."""
    positions: np.ndarray            # (V, 3) float32, world space
    indices: np.ndarray              # (T, 3) int32
    normals: np.ndarray              # (T, 3) float32, world space face normals
    cull_mode: int
    material_index: int
    bvh_nodes: Optional[np.ndarray] = None   # BVH_NODE_DTYPE, only used by the oracle's BVH path

    @property
    def triangle_count(self) -> int:
        return int(self.indices.shape[0])


@dataclasses.dataclass
class FlatScene:
    """Scene data in the SoA layout of include/rt_b200.h."""
    sphere_origin: np.ndarray        # (3, S) float32: rows are the x, y, z arrays
    sphere_radius: np.ndarray        # (S,)
    sphere_material: np.ndarray      # (S,) uint8
    plane_origin: np.ndarray         # (3, P)
    plane_normal: np.ndarray         # (3, P)
    plane_material: np.ndarray       # (P,) uint8
    light_origin: np.ndarray         # (3, L)
    light_direction: np.ndarray      # (3, L)
    light_color: np.ndarray          # (3, L)
    light_intensity: np.ndarray      # (L,)
    light_type: np.ndarray           # (L,) int32
    materials: np.ndarray            # (K,) MATERIAL_DTYPE
    meshes: List[Mesh]
    camera: Camera
    # the frame state the dump was taken with (tests may override)
    width: int = 640
    height: int = 480
    aspect_ratio: float = 640 / 480
    lighting_mode: int = LIGHTING_COMBINED
    shadows_enabled: int = 1

    def triangle_count(self) -> int:
        return sum(m.triangle_count for m in self.meshes)


class _Reader:
    def __init__(self, data: bytes):
        self.data = data
        self.pos = 0

    def take(self, dtype, count):
        dt = np.dtype(dtype)
        n = dt.itemsize * count
        if self.pos + n > len(self.data):
            raise ValueError("truncated RTSC file")
        out = np.frombuffer(self.data, dtype=dt, count=count, offset=self.pos).copy()
        self.pos += n
        return out

    def i32(self, count=1):
        return self.take("<i4", count)

    def f32(self, count=1):
        return self.take("<f4", count)


def load_rtsc(path) -> FlatScene:
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != MAGIC:
        raise ValueError(f"{path}: not an RTSC0001 file")
    r = _Reader(data)
    r.pos = 8
    width, height, mode, shadows = (int(v) for v in r.i32(4))
    aspect = float(r.f32(1)[0])
    cam = r.f32(13)
    camera = Camera(origin=cam[0:3].copy(), fov=float(cam[3]), right=cam[4:7].copy(), up=cam[7:10].copy(),
                    forward=cam[10:13].copy())
    n_s, n_p, n_l, n_k, n_m = (int(v) for v in r.i32(5))

    sph = r.take(np.dtype([("o", "<f4", (3,)), ("r", "<f4"), ("m", "<i4")]), n_s)
    pla = r.take(np.dtype([("o", "<f4", (3,)), ("n", "<f4", (3,)), ("m", "<i4")]), n_p)
    lig = r.take(np.dtype([("o", "<f4", (3,)), ("d", "<f4", (3,)), ("c", "<f4", (3,)), ("i", "<f4"), ("t", "<i4")]), n_l)
    mats = r.take(MATERIAL_DTYPE, n_k)
    meshes = []
    for _ in range(n_m):
        n_v, n_t, cull, mat, n_nodes = (int(v) for v in r.i32(5))
        pos = r.f32(3 * n_v).reshape(n_v, 3)
        idx = r.i32(3 * n_t).reshape(n_t, 3)
        nrm = r.f32(3 * n_t).reshape(n_t, 3)
        nodes = r.take(BVH_NODE_DTYPE, n_nodes) if n_nodes else None
        meshes.append(Mesh(pos, idx, nrm, cull, mat, nodes))
    if r.pos != len(data):
        raise ValueError(f"{path}: {len(data) - r.pos} trailing bytes")

    def soa(a):  # (N,3) AoS -> (3,N) contiguous SoA
        return np.ascontiguousarray(a.reshape(-1, 3).T, dtype=np.float32)

    return FlatScene(
        sphere_origin=soa(sph["o"]), sphere_radius=np.ascontiguousarray(sph["r"], dtype=np.float32),
        sphere_material=sph["m"].astype(np.uint8),
        plane_origin=soa(pla["o"]), plane_normal=soa(pla["n"]), plane_material=pla["m"].astype(np.uint8),
        light_origin=soa(lig["o"]), light_direction=soa(lig["d"]), light_color=soa(lig["c"]),
        light_intensity=np.ascontiguousarray(lig["i"], dtype=np.float32), light_type=lig["t"].astype(np.int32),
        materials=mats, meshes=meshes, camera=camera, width=width, height=height, aspect_ratio=aspect,
        lighting_mode=mode, shadows_enabled=shadows)


@dataclasses.dataclass
class MeshSource:
    """What TriangleMesh::UpdateTransforms consumes (reference source/DataTypes.h:210-230)."""
    positions: np.ndarray     # (V, 3) float32, untransformed
    indices: np.ndarray       # (T, 3) int32
    normals: np.ndarray       # (T, 3) float32, untransformed face normals
    transform: np.ndarray     # (4, 4) float32: Matrix::data rows of finalTransform = S * R * T


def load_rtms(path) -> List[MeshSource]:
    """"RTMS0001": i32 n_meshes, then per mesh i32 n_vertices, n_triangles, f32 positions, i32 indices,
    f32 normals, f32[16] finalTransform (written by oracle/ref_driver.cpp --dump-mesh-source)."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != b"RTMS0001":
        raise ValueError(f"{path}: not an RTMS0001 file")
    r = _Reader(data)
    r.pos = 8
    out = []
    for _ in range(int(r.i32(1)[0])):
        n_v, n_t = (int(v) for v in r.i32(2))
        pos = r.f32(3 * n_v).reshape(n_v, 3)
        idx = r.i32(3 * n_t).reshape(n_t, 3)
        nrm = r.f32(3 * n_t).reshape(n_t, 3)
        out.append(MeshSource(pos, idx, nrm, r.f32(16).reshape(4, 4)))
    if r.pos != len(data):
        raise ValueError(f"{path}: trailing bytes")
    return out


@dataclasses.dataclass
class MeshSteps:
    """Input of a sequence of TriangleMesh::UpdateTransforms calls with BuildBVH (reference
    source/DataTypes.h:210-236, 294-389): the mesh as it is BEFORE the first call (BuildBVH reorders indices and
    normals in place, so every build starts from the order the previous one left) and one finalTransform per call."""
    positions: np.ndarray     # (V, 3) float32, untransformed
    indices: np.ndarray       # (T, 3) int32, order before the first call
    normals: np.ndarray       # (T, 3) float32, untransformed face normals, same order
    transforms: np.ndarray    # (n_steps, 4, 4) float32


def load_rtmp(path) -> List[MeshSteps]:
    """"RTMP0001": i32 n_meshes, n_steps, then per mesh i32 n_vertices, n_triangles, f32 positions, i32 indices,
    f32 normals, n_steps x f32[16] finalTransform (written by oracle/ref_driver.cpp --yaw-steps --dump-mesh-steps)."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != b"RTMP0001":
        raise ValueError(f"{path}: not an RTMP0001 file")
    r = _Reader(data)
    r.pos = 8
    n_meshes, n_steps = (int(v) for v in r.i32(2))
    out = []
    for _ in range(n_meshes):
        n_v, n_t = (int(v) for v in r.i32(2))
        pos = r.f32(3 * n_v).reshape(n_v, 3)
        idx = r.i32(3 * n_t).reshape(n_t, 3)
        nrm = r.f32(3 * n_t).reshape(n_t, 3)
        out.append(MeshSteps(pos, idx, nrm, r.f32(16 * n_steps).reshape(n_steps, 4, 4)))
    if r.pos != len(data):
        raise ValueError(f"{path}: trailing bytes")
    return out
