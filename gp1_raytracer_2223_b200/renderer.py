"""Host-side mirror of the reference's ``dae::Renderer`` over the C ABI.

The reference drives the path with ``Renderer(pWindow)``, ``Render(pScene)``,
``CycleLightingMode()`` and ``ToggleShadows()`` (reference source/Renderer.h:20-36,
source/Renderer.cpp:24-98, 189-193).  This class keeps those names and semantics
(start state Combined + shadows on, F3 order ObservedArea -> Radiance -> BRDF -> Combined)
and forwards the frame to ``librt_b200.so``.  It is plumbing for tests and the benchmark;
the C++ drop-in for the reference's own ``main.cpp`` is ``host/Renderer.cpp``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._abi import SceneViews, camera_struct, frame_struct, rt_counters, rt_timing
from .scene_file import FlatScene, LIGHTING_COMBINED


class RtError(RuntimeError):
    pass


class Context:
    """One ``rt_context``: device state for one Renderer."""

    def __init__(self, device_ids: Optional[Sequence[int]] = None):
        self.lib = _lib.load()
        self.handle = C.c_void_p()
        if device_ids:
            arr = (C.c_int32 * len(device_ids))(*device_ids)
            rc = self.lib.rt_create(arr, len(device_ids), C.byref(self.handle))
        else:
            rc = self.lib.rt_create(None, 0, C.byref(self.handle))
        if rc != 0:
            msg = self.lib.rt_last_error(None)
            raise RtError(f"rt_create failed ({rc}): {msg.decode() if msg else ''}")

    def check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.rt_last_error(self.handle)
            raise RtError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if self.handle:
            self.lib.rt_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_count(self) -> int:
        return int(self.lib.rt_device_count(self.handle))

    def upload_scene(self, scene: FlatScene) -> SceneViews:
        v = SceneViews(scene)
        self.check(self.lib.rt_upload_spheres(self.handle, C.byref(v.spheres)), "rt_upload_spheres")
        self.check(self.lib.rt_upload_planes(self.handle, C.byref(v.planes)), "rt_upload_planes")
        self.check(self.lib.rt_upload_lights(self.handle, C.byref(v.lights)), "rt_upload_lights")
        self.check(self.lib.rt_upload_materials(self.handle, v.materials, v.material_count), "rt_upload_materials")
        self.check(self.lib.rt_set_mesh_count(self.handle, len(v.meshes)), "rt_set_mesh_count")
        for i, m in enumerate(v.meshes):
            self.check(self.lib.rt_upload_mesh(self.handle, i, C.byref(m)), "rt_upload_mesh")
        return v

    def set_mesh_path(self, mesh_path: int) -> None:
        """The reference's compile-time `#define BVH` as a run-time choice (0 auto, 1 slab + linear, 2 BVH)."""
        self.check(self.lib.rt_set_mesh_path(self.handle, int(mesh_path)), "rt_set_mesh_path")

    def upload_mesh_source(self, mesh_id: int, positions, indices, normals, cull_mode: int, material_index: int) -> None:
        """Untransformed mesh, once (device-side TriangleMesh::UpdateTransforms, reference source/DataTypes.h:210-230)."""
        from ._abi import rt_mesh_source, c_float_p, c_i32_p
        pos = np.ascontiguousarray(positions, dtype=np.float32)
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        nrm = np.ascontiguousarray(normals, dtype=np.float32)
        src = rt_mesh_source(pos.ctypes.data_as(c_float_p), int(pos.shape[0]), idx.ctypes.data_as(c_i32_p), nrm.ctypes.data_as(c_float_p),
                             int(idx.shape[0]), int(cull_mode), int(material_index))
        self.check(self.lib.rt_upload_mesh_source(self.handle, mesh_id, C.byref(src)), "rt_upload_mesh_source")

    def transform_mesh(self, mesh_id: int, transform) -> None:
        """finalTransform = scale * rotation * translation as the 16 floats of Matrix::data (per frame: 64 bytes)."""
        m = np.ascontiguousarray(transform, dtype=np.float32).reshape(16)
        self.check(self.lib.rt_transform_mesh(self.handle, mesh_id, m.ctypes.data_as(C.POINTER(C.c_float))), "rt_transform_mesh")

    def set_mesh_device_bvh(self, mesh_id: int, enable: bool = True) -> None:
        """Every later transform_mesh is a full TriangleMesh::UpdateTransforms of the reference, BuildBVH included
        (reference source/DataTypes.h:210-236, 294-483), run on the device; the mesh is rendered by the BVH body."""
        self.check(self.lib.rt_set_mesh_device_bvh(self.handle, mesh_id, 1 if enable else 0), "rt_set_mesh_device_bvh")

    def read_mesh_build(self, mesh_id: int, triangle_count: int):
        """(indices, normals, nodes) the device-side builds left behind: TriangleMesh::indices / normals in their new
        order and the tree as a structured array (min_aabb, max_aabb, first, triangle_count, escape)."""
        from ._abi import rt_built_node, c_float_p, c_i32_p
        idx = np.zeros(3 * triangle_count, dtype=np.int32)
        nrm = np.zeros((triangle_count, 3), dtype=np.float32)
        capacity = max(2 * triangle_count - 1, 1)
        nodes = (rt_built_node * capacity)()
        count = C.c_int32(0)
        self.check(self.lib.rt_read_mesh_build(self.handle, mesh_id, idx.ctypes.data_as(c_i32_p), nrm.ctypes.data_as(c_float_p), nodes, capacity,
                                               C.byref(count)), "rt_read_mesh_build")
        dt = np.dtype([("min_aabb", np.float32, 3), ("max_aabb", np.float32, 3), ("first", np.int32), ("triangle_count", np.int32),
                       ("escape", np.int32)])
        return idx, nrm, np.frombuffer(nodes, dtype=dt, count=count.value).copy()

    def set_kernel_variant(self, variant: int) -> None:
        """0 auto (persistent warps on deep frames, else tiled), 1 tiled one-pixel kernel, 2 packed two-pixel FFMA2
        kernel, 3 persistent warps."""
        self.check(self.lib.rt_set_kernel_variant(self.handle, int(variant)), "rt_set_kernel_variant")

    def upload_mesh(self, mesh_id: int, mesh) -> None:
        """Re-upload one mesh after TriangleMesh::UpdateTransforms (reference source/DataTypes.h:210-236)."""
        self.upload_mesh_descriptor(mesh_id, self.mesh_descriptor(mesh))

    def mesh_descriptor(self, mesh):
        """The rt_mesh_desc of a mesh, built once: a caller that re-uploads the same host arrays every frame (what the
        C++ drop-in does with the TriangleMesh's vectors) pays for the ctypes marshalling only here."""
        v = SceneViews.__new__(SceneViews)
        v._keep = []
        desc = SceneViews.mesh_desc(v, mesh)
        desc._keepalive = v
        return desc

    def upload_mesh_descriptor(self, mesh_id: int, desc) -> None:
        self.check(self.lib.rt_upload_mesh(self.handle, mesh_id, C.byref(desc)), "rt_upload_mesh")

    def measure_fp32_peak(self, use_fma: bool) -> dict:
        tf, ms = C.c_double(), C.c_float()
        self.check(self.lib.rt_measure_fp32_peak(self.handle, int(use_fma), C.byref(tf), C.byref(ms)),
                   "rt_measure_fp32_peak")
        return {"tflops": tf.value, "ms": ms.value}

    def timing(self) -> dict:
        t = rt_timing()
        self.check(self.lib.rt_get_timing(self.handle, C.byref(t)), "rt_get_timing")
        return {"kernel_ms": t.kernel_ms, "gather_ms": t.gather_ms, "d2h_ms": t.d2h_ms, "total_ms": t.total_ms,
                "kernel_launches": t.kernel_launches}


class Renderer:
    """``dae::Renderer`` with the pixel loop on the GPU."""

    def __init__(self, width: int, height: int, device_ids: Optional[Sequence[int]] = None,
                 shifts=(16, 8, 0), alpha_mask: int = 0):
        self.width = int(width)                       # m_Width
        self.height = int(height)                     # m_Height
        self.aspect_ratio = float(np.float32(self.width) / np.float32(self.height))   # Renderer.cpp:31
        self.lighting_mode = LIGHTING_COMBINED        # Renderer.h:49
        self.shadows_enabled = True                   # Renderer.h:50
        self.shifts = shifts
        self.alpha_mask = alpha_mask
        self.ctx = Context(device_ids)
        self.scene: Optional[FlatScene] = None
        self._views = None

    # -- reference API ------------------------------------------------------------------
    def CycleLightingMode(self):                      # Renderer.cpp:189-193
        self.lighting_mode = (self.lighting_mode + 1) % 4

    def ToggleShadows(self):                          # Renderer.h:34-36
        self.shadows_enabled = not self.shadows_enabled

    def SetScene(self, scene: FlatScene):
        """Upload what Render reads from the Scene (Renderer.cpp:36-38): call after Initialize/Update."""
        self.scene = scene
        self._views = self.ctx.upload_scene(scene)

    def Render(self, out: Optional[np.ndarray] = None, camera=None) -> np.ndarray:
        """Renderer::Render (Renderer.cpp:34-98): blocking, returns the uint32 surface (H, W)."""
        if self.scene is None:
            raise RtError("SetScene must be called before Render")
        if out is None:
            out = np.empty((self.height, self.width), dtype=np.uint32)
        if out.dtype != np.uint32 or out.shape != (self.height, self.width) or out.strides[1] != 4:
            raise RtError("out must be a uint32 (height, width) array with contiguous rows")
        cam = camera_struct(camera if camera is not None else self.scene.camera)
        frame = self._frame()
        self.ctx.check(self.ctx.lib.rt_render(self.ctx.handle, C.byref(cam), C.byref(frame), out.ctypes.data,
                                              out.strides[0]), "rt_render")
        return out

    # -- measurement helpers --------------------------------------------------------------
    def render_device(self, camera=None) -> dict:
        cam = camera_struct(camera if camera is not None else self.scene.camera)
        frame = self._frame()
        self.ctx.check(self.ctx.lib.rt_render_device(self.ctx.handle, C.byref(cam), C.byref(frame)), "rt_render_device")
        return self.ctx.timing()

    def render_host_ptr(self, host_ptr: int, pitch_bytes: int, camera=None) -> dict:
        """rt_render into caller memory given as an address (e.g. a pinned torch tensor)."""
        cam = camera_struct(camera if camera is not None else self.scene.camera)
        frame = self._frame()
        self.ctx.check(self.ctx.lib.rt_render(self.ctx.handle, C.byref(cam), C.byref(frame), host_ptr, pitch_bytes),
                       "rt_render")
        return self.ctx.timing()

    def download(self) -> np.ndarray:
        out = np.empty((self.height, self.width), dtype=np.uint32)
        self.ctx.check(self.ctx.lib.rt_download_frame(self.ctx.handle, out.ctypes.data, out.strides[0]),
                       "rt_download_frame")
        return out

    def render_rows_device(self, row_begin: int, row_count: int, device_ptr: int, stream: int = 0, camera=None):
        cam = camera_struct(camera if camera is not None else self.scene.camera)
        frame = self._frame()
        self.ctx.check(self.ctx.lib.rt_render_rows_device(self.ctx.handle, C.byref(cam), C.byref(frame), row_begin,
                                                          row_count, device_ptr, stream), "rt_render_rows_device")

    def render_strips_device(self, strip_first: int, strip_step: int, device_ptr: int, stream: int = 0, camera=None):
        cam = camera_struct(camera if camera is not None else self.scene.camera)
        frame = self._frame()
        self.ctx.check(self.ctx.lib.rt_render_strips_device(self.ctx.handle, C.byref(cam), C.byref(frame), strip_first,
                                                            strip_step, device_ptr, stream), "rt_render_strips_device")

    def unstripe_device(self, src_ptr: int, dst_ptr: int, world: int, strips_per_rank: int, stream: int = 0):
        self.ctx.check(self.ctx.lib.rt_unstripe_device(self.ctx.handle, src_ptr, dst_ptr, self.width, self.height, world,
                                                       strips_per_rank, stream), "rt_unstripe_device")

    # -- one process per GPU, gather fused into the kernel (rt_b200.h) -----------------------------------
    def frame_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self.ctx.check(self.ctx.lib.rt_frame_export(self.ctx.handle, self.width, self.height, buf), "rt_frame_export")
        return buf.raw

    def frame_import(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        self.ctx.check(self.ctx.lib.rt_frame_import(self.ctx.handle, C.c_char_p(handle), C.byref(ptr)), "rt_frame_import")
        return int(ptr.value)

    def frame_release(self, ptr: int) -> None:
        self.ctx.check(self.ctx.lib.rt_frame_release(self.ctx.handle, ptr), "rt_frame_release")

    def render_strips_to_frame(self, strip_first: int, strip_step: int, frame_ptr: int = 0, stream: int = 0, camera=None):
        cam = camera_struct(camera if camera is not None else self.scene.camera)
        frame = self._frame()
        self.ctx.check(self.ctx.lib.rt_render_strips_to_frame(self.ctx.handle, C.byref(cam), C.byref(frame), strip_first,
                                                              strip_step, frame_ptr or None, stream), "rt_render_strips_to_frame")

    def render_strips_to_host(self, strip_first: int, strip_step: int, host_ptr: int, pitch_bytes: int, camera=None) -> dict:
        """This rank's strips rendered and copied into the (shared) host surface over this GPU's own PCIe link."""
        cam = camera_struct(camera if camera is not None else self.scene.camera)
        frame = self._frame()
        self.ctx.check(self.ctx.lib.rt_render_strips_to_host(self.ctx.handle, C.byref(cam), C.byref(frame), strip_first, strip_step,
                                                             host_ptr, pitch_bytes), "rt_render_strips_to_host")
        return self.ctx.timing()

    def frame_signal(self, frame_ptr: int, stream: int = 0) -> None:
        self.ctx.check(self.ctx.lib.rt_frame_signal(self.ctx.handle, frame_ptr, self.width, self.height, stream), "rt_frame_signal")

    def frame_wait(self, expected: int, stream: int = 0) -> None:
        self.ctx.check(self.ctx.lib.rt_frame_wait(self.ctx.handle, expected & 0xFFFFFFFF, stream), "rt_frame_wait")

    def render_strips_to_frame_banded(self, strip_first: int, strip_step: int, frame_ptr: int, bands: int, stream: int = 0, camera=None):
        cam = camera_struct(camera if camera is not None else self.scene.camera)
        frame = self._frame()
        self.ctx.check(self.ctx.lib.rt_render_strips_to_frame_banded(self.ctx.handle, C.byref(cam), C.byref(frame), strip_first, strip_step,
                                                                     frame_ptr or None, bands, stream), "rt_render_strips_to_frame_banded")

    def frame_present(self, host_ptr: int, pitch_bytes: int, bands: int, frame_number: int) -> None:
        self.ctx.check(self.ctx.lib.rt_frame_present(self.ctx.handle, host_ptr, pitch_bytes, bands, frame_number & 0xFFFFFFFF), "rt_frame_present")

    def clear_frame(self, pixel: int = 0xDEADBEEF) -> None:
        """Poison device 0's frame buffer (a later frame check cannot pass on a stale frame)."""
        self.ctx.check(self.ctx.lib.rt_clear_frame(self.ctx.handle, pixel & 0xFFFFFFFF), "rt_clear_frame")

    def register_surface(self, host_ptr: int, nbytes: int) -> None:
        """Pin a caller-owned surface for direct device-to-host copies; it must outlive the registration (rt_b200.h)."""
        self.ctx.check(self.ctx.lib.rt_register_surface(self.ctx.handle, host_ptr, nbytes), "rt_register_surface")

    def unregister_surface(self, host_ptr: int) -> None:
        self.ctx.check(self.ctx.lib.rt_unregister_surface(self.ctx.handle, host_ptr), "rt_unregister_surface")

    def download_to(self, host_ptr: int, pitch_bytes: int) -> None:
        self.ctx.check(self.ctx.lib.rt_download_frame(self.ctx.handle, host_ptr, pitch_bytes), "rt_download_frame")

    def count_frame(self, camera=None, mesh_path: int = 1) -> np.ndarray:
        """Test histogram of one frame; mesh_path 1 = slab + linear (the algorithmic counts), 2 = BVH."""
        cam = camera_struct(camera if camera is not None else self.scene.camera)
        frame = self._frame()
        cnt = rt_counters()
        self.ctx.check(self.ctx.lib.rt_count_frame(self.ctx.handle, C.byref(cam), C.byref(frame), int(mesh_path),
                                                   C.byref(cnt)), "rt_count_frame")
        return np.array(list(cnt.slot), dtype=np.uint64)

    def _frame(self):
        key = (self.width, self.height, self.lighting_mode, self.shadows_enabled, self.aspect_ratio, self.shifts, self.alpha_mask)
        if getattr(self, "_frame_key", None) != key:
            self._frame_key = key
            self._frame_struct = frame_struct(self.width, self.height, self.lighting_mode, self.shadows_enabled, self.aspect_ratio,
                                              self.shifts, self.alpha_mask)
        return self._frame_struct

    def close(self):
        self.ctx.close()
