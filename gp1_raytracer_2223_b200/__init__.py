"""B200-native per-pixel path of GP1_Raytracer_2223 (Renderer::Render -> RenderPixel).

The product is ``librt_b200.so`` (CUDA, sm_100a) behind the C ABI in ``include/rt_b200.h``;
this package is the thin host-side plumbing around it.  Nothing here computes a pixel on
the CPU.
"""
from .scene_file import FlatScene, Camera, Mesh, load_rtsc  # noqa: F401
from .renderer import Renderer, Context, RtError  # noqa: F401
from .obj_file import ObjMesh, parse_obj  # noqa: F401
