"""ctypes mirror of include/rt_b200.h (struct layouts only; no library is loaded here)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .scene_file import BVH_NODE_DTYPE, Camera, FlatScene, Mesh

RT_COUNTER_SLOTS = 40
RT_B200_ABI_VERSION = 3
MESH_PATH_AUTO, MESH_PATH_SLAB_LINEAR, MESH_PATH_BVH = 0, 1, 2

c_float_p = C.POINTER(C.c_float)
c_u8_p = C.POINTER(C.c_uint8)
c_i32_p = C.POINTER(C.c_int32)


class rt_material_desc(C.Structure):
    _fields_ = [("tag", C.c_int32), ("color", C.c_float * 3), ("p0", C.c_float), ("p1", C.c_float),
                ("p2", C.c_float), ("reserved", C.c_float)]


class rt_spheres_soa(C.Structure):
    _fields_ = [("origin_x", c_float_p), ("origin_y", c_float_p), ("origin_z", c_float_p), ("radius", c_float_p),
                ("material_index", c_u8_p), ("count", C.c_int32)]


class rt_planes_soa(C.Structure):
    _fields_ = [("origin_x", c_float_p), ("origin_y", c_float_p), ("origin_z", c_float_p),
                ("normal_x", c_float_p), ("normal_y", c_float_p), ("normal_z", c_float_p),
                ("material_index", c_u8_p), ("count", C.c_int32)]


class rt_lights_soa(C.Structure):
    _fields_ = [("origin_x", c_float_p), ("origin_y", c_float_p), ("origin_z", c_float_p),
                ("direction_x", c_float_p), ("direction_y", c_float_p), ("direction_z", c_float_p),
                ("color_r", c_float_p), ("color_g", c_float_p), ("color_b", c_float_p),
                ("intensity", c_float_p), ("type", c_i32_p), ("count", C.c_int32)]


class rt_bvh_node(C.Structure):
    _fields_ = [("min_aabb", C.c_float * 3), ("max_aabb", C.c_float * 3), ("first_idx", C.c_uint32),
                ("idx_count", C.c_uint32), ("left_node", C.c_uint32)]


class rt_mesh_desc(C.Structure):
    _fields_ = [("positions", c_float_p), ("vertex_count", C.c_int32), ("indices", c_i32_p),
                ("normals", c_float_p), ("triangle_count", C.c_int32), ("cull_mode", C.c_int32),
                ("material_index", C.c_uint8), ("aabb_min", c_float_p), ("aabb_max", c_float_p),
                ("bvh_nodes", C.POINTER(rt_bvh_node)), ("bvh_node_count", C.c_int32)]


class rt_mesh_source(C.Structure):
    _fields_ = [("positions", c_float_p), ("vertex_count", C.c_int32), ("indices", c_i32_p), ("normals", c_float_p),
                ("triangle_count", C.c_int32), ("cull_mode", C.c_int32), ("material_index", C.c_uint8)]


class rt_built_node(C.Structure):
    _fields_ = [("min_aabb", C.c_float * 3), ("max_aabb", C.c_float * 3), ("first", C.c_int32), ("triangle_count", C.c_int32),
                ("escape", C.c_int32)]


class rt_camera(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("fov", C.c_float), ("right", C.c_float * 3), ("up", C.c_float * 3),
                ("forward", C.c_float * 3)]


class rt_frame_desc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("aspect_ratio", C.c_float),
                ("lighting_mode", C.c_int32), ("shadows_enabled", C.c_int32),
                ("r_shift", C.c_uint8), ("g_shift", C.c_uint8), ("b_shift", C.c_uint8), ("reserved", C.c_uint8),
                ("alpha_mask", C.c_uint32)]


class rt_timing(C.Structure):
    _fields_ = [("kernel_ms", C.c_float), ("gather_ms", C.c_float), ("d2h_ms", C.c_float), ("total_ms", C.c_float),
                ("kernel_launches", C.c_int32), ("reserved", C.c_int32)]


class rt_counters(C.Structure):
    _fields_ = [("slot", C.c_uint64 * RT_COUNTER_SLOTS)]


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_float_p)


class SceneViews:
    """Holds the numpy arrays alive and exposes them as the C structs of rt_b200.h."""

    def __init__(self, scene: FlatScene):
        self.scene = scene
        s = scene
        self._keep = []

        def row(a, k):
            r = np.ascontiguousarray(a[k], dtype=np.float32)
            self._keep.append(r)
            return _fp(r)

        def vec(a, dtype, ptr):
            r = np.ascontiguousarray(a, dtype=dtype)
            self._keep.append(r)
            return r.ctypes.data_as(ptr)

        self.spheres = rt_spheres_soa(row(s.sphere_origin, 0), row(s.sphere_origin, 1), row(s.sphere_origin, 2),
                                      vec(s.sphere_radius, np.float32, c_float_p),
                                      vec(s.sphere_material, np.uint8, c_u8_p), int(s.sphere_radius.shape[0]))
        self.planes = rt_planes_soa(row(s.plane_origin, 0), row(s.plane_origin, 1), row(s.plane_origin, 2),
                                    row(s.plane_normal, 0), row(s.plane_normal, 1), row(s.plane_normal, 2),
                                    vec(s.plane_material, np.uint8, c_u8_p), int(s.plane_material.shape[0]))
        self.lights = rt_lights_soa(row(s.light_origin, 0), row(s.light_origin, 1), row(s.light_origin, 2),
                                    row(s.light_direction, 0), row(s.light_direction, 1), row(s.light_direction, 2),
                                    row(s.light_color, 0), row(s.light_color, 1), row(s.light_color, 2),
                                    vec(s.light_intensity, np.float32, c_float_p),
                                    vec(s.light_type, np.int32, c_i32_p), int(s.light_type.shape[0]))
        k = len(s.materials)
        self.materials = (rt_material_desc * max(k, 1))()
        for i in range(k):
            m = s.materials[i]
            self.materials[i].tag = int(m["tag"])
            for c in range(3):
                self.materials[i].color[c] = float(m["color"][c])
            self.materials[i].p0 = float(m["p0"])
            self.materials[i].p1 = float(m["p1"])
            self.materials[i].p2 = float(m["p2"])
        self.material_count = k
        self.meshes = [self.mesh_desc(m) for m in s.meshes]

    def mesh_desc(self, m: Mesh, with_bvh: bool = True) -> rt_mesh_desc:
        pos = np.ascontiguousarray(m.positions, dtype=np.float32)
        idx = np.ascontiguousarray(m.indices, dtype=np.int32)
        nrm = np.ascontiguousarray(m.normals, dtype=np.float32)
        self._keep += [pos, idx, nrm]
        nodes, n_nodes = None, 0
        if with_bvh and m.bvh_nodes is not None and len(m.bvh_nodes):
            arr = np.ascontiguousarray(m.bvh_nodes, dtype=BVH_NODE_DTYPE)
            self._keep.append(arr)
            nodes, n_nodes = C.cast(arr.ctypes.data, C.POINTER(rt_bvh_node)), len(arr)
        return rt_mesh_desc(_fp(pos.reshape(-1)), int(pos.shape[0]), idx.reshape(-1).ctypes.data_as(c_i32_p),
                            _fp(nrm.reshape(-1)), int(idx.shape[0]), int(m.cull_mode), int(m.material_index),
                            None, None, nodes, n_nodes)


def camera_struct(cam: Camera) -> rt_camera:
    c = rt_camera()
    for k in range(3):
        c.origin[k] = float(cam.origin[k])
        c.right[k] = float(cam.right[k])
        c.up[k] = float(cam.up[k])
        c.forward[k] = float(cam.forward[k])
    c.fov = float(cam.fov)
    return c


def frame_struct(width: int, height: int, lighting_mode: int = 3, shadows: bool = True,
                 aspect_ratio: float | None = None, shifts=(16, 8, 0), alpha_mask: int = 0) -> rt_frame_desc:
    """aspect_ratio defaults to width / float32(height) as in reference source/Renderer.cpp:31."""
    if aspect_ratio is None:
        aspect_ratio = float(np.float32(width) / np.float32(height))
    return rt_frame_desc(width, height, aspect_ratio, int(lighting_mode), int(bool(shadows)),
                         shifts[0], shifts[1], shifts[2], 0, alpha_mask)
