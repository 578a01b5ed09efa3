"""Loads gp1_raytracer_2223_b200/librt_b200.so and declares the C ABI (include/rt_b200.h).

No fallback: a missing library or a missing symbol raises.
"""
from __future__ import annotations

import ctypes as C
import os

from ._abi import (rt_built_node, rt_mesh_source, rt_camera, rt_counters, rt_frame_desc, rt_lights_soa, rt_material_desc, rt_mesh_desc,
                   rt_planes_soa, rt_spheres_soa, rt_timing)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RT_B200_LIB") or os.path.join(HERE, "librt_b200.so")   # override: A/B builds

# every symbol include/rt_b200.h declares: name -> (restype, argtypes)
_ctx = C.c_void_p
SYMBOLS = {
    "rt_abi_version": (C.c_int, []),
    "rt_create": (C.c_int, [C.POINTER(C.c_int32), C.c_int32, C.POINTER(_ctx)]),
    "rt_destroy": (C.c_int, [_ctx]),
    "rt_last_error": (C.c_char_p, [_ctx]),
    "rt_device_count": (C.c_int, [_ctx]),
    "rt_upload_spheres": (C.c_int, [_ctx, C.POINTER(rt_spheres_soa)]),
    "rt_upload_planes": (C.c_int, [_ctx, C.POINTER(rt_planes_soa)]),
    "rt_upload_lights": (C.c_int, [_ctx, C.POINTER(rt_lights_soa)]),
    "rt_upload_materials": (C.c_int, [_ctx, C.POINTER(rt_material_desc), C.c_int32]),
    "rt_set_mesh_count": (C.c_int, [_ctx, C.c_int32]),
    "rt_set_mesh_path": (C.c_int, [_ctx, C.c_int32]),
    "rt_upload_mesh_source": (C.c_int, [_ctx, C.c_int32, C.POINTER(rt_mesh_source)]),
    "rt_transform_mesh": (C.c_int, [_ctx, C.c_int32, C.POINTER(C.c_float)]),
    "rt_set_mesh_device_bvh": (C.c_int, [_ctx, C.c_int32, C.c_int32]),
    "rt_read_mesh_build": (C.c_int, [_ctx, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(rt_built_node), C.c_int32,
                                     C.POINTER(C.c_int32)]),
    "rt_set_kernel_variant": (C.c_int, [_ctx, C.c_int32]),
    "rt_upload_mesh": (C.c_int, [_ctx, C.c_int32, C.POINTER(rt_mesh_desc)]),
    "rt_render": (C.c_int, [_ctx, C.POINTER(rt_camera), C.POINTER(rt_frame_desc), C.c_void_p, C.c_int32]),
    "rt_render_device": (C.c_int, [_ctx, C.POINTER(rt_camera), C.POINTER(rt_frame_desc)]),
    "rt_download_frame": (C.c_int, [_ctx, C.c_void_p, C.c_int32]),
    "rt_clear_frame": (C.c_int, [_ctx, C.c_uint32]),
    "rt_register_surface": (C.c_int, [_ctx, C.c_void_p, C.c_size_t]),
    "rt_unregister_surface": (C.c_int, [_ctx, C.c_void_p]),
    "rt_render_rows_device": (C.c_int, [_ctx, C.POINTER(rt_camera), C.POINTER(rt_frame_desc), C.c_int32, C.c_int32,
                                        C.c_void_p, C.c_void_p]),
    "rt_render_strips_device": (C.c_int, [_ctx, C.POINTER(rt_camera), C.POINTER(rt_frame_desc), C.c_int32, C.c_int32,
                                          C.c_void_p, C.c_void_p]),
    "rt_unstripe_device": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_void_p]),
    "rt_frame_export": (C.c_int, [_ctx, C.c_int32, C.c_int32, C.c_void_p]),
    "rt_frame_import": (C.c_int, [_ctx, C.c_void_p, C.POINTER(C.c_void_p)]),
    "rt_frame_release": (C.c_int, [_ctx, C.c_void_p]),
    "rt_render_strips_to_frame": (C.c_int, [_ctx, C.POINTER(rt_camera), C.POINTER(rt_frame_desc), C.c_int32, C.c_int32,
                                            C.c_void_p, C.c_void_p]),
    "rt_render_strips_to_host": (C.c_int, [_ctx, C.POINTER(rt_camera), C.POINTER(rt_frame_desc), C.c_int32, C.c_int32,
                                           C.c_void_p, C.c_int32]),
    "rt_frame_signal": (C.c_int, [_ctx, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "rt_frame_wait": (C.c_int, [_ctx, C.c_uint32, C.c_void_p]),
    "rt_render_strips_to_frame_banded": (C.c_int, [_ctx, C.POINTER(rt_camera), C.POINTER(rt_frame_desc), C.c_int32, C.c_int32,
                                                   C.c_void_p, C.c_int32, C.c_void_p]),
    "rt_frame_present": (C.c_int, [_ctx, C.c_void_p, C.c_int32, C.c_int32, C.c_uint32]),
    "rt_host_arrive_and_wait": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_double]),
    "rt_get_timing": (C.c_int, [_ctx, C.POINTER(rt_timing)]),
    "rt_measure_fp32_peak": (C.c_int, [_ctx, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_float)]),
    "rt_count_frame": (C.c_int, [_ctx, C.POINTER(rt_camera), C.POINTER(rt_frame_desc), C.c_int32,
                                 C.POINTER(rt_counters)]),
}

_lib = None


def load():
    """dlopen the CUDA library (this does not need a GPU; rt_create does)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m gp1_raytracer_2223_b200.build` "
                               "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)      # AttributeError if the export is missing
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib
