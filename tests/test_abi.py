"""The C-ABI library loads on a CPU box and exports every symbol include/rt_b200.h declares."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from gp1_raytracer_2223_b200 import _abi, _lib, build


@pytest.fixture(scope="module")
def lib():
    build.ensure()
    return _lib.load()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:const\s+char\s*\*|int)\s+(rt_[a-z0-9_]+)\s*\(", text, flags=re.M)))


def test_header_and_binding_agree(lib):
    names = declared_functions()
    assert len(names) >= 20
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in rt_b200.h but not exported"


def test_abi_version(lib):
    assert lib.rt_abi_version() == _abi.RT_B200_ABI_VERSION


def test_struct_sizes_match_header():
    # layouts the header fixes (x86-64 SysV)
    assert C.sizeof(_abi.rt_material_desc) == 32
    assert C.sizeof(_abi.rt_camera) == 52
    assert C.sizeof(_abi.rt_frame_desc) == 28
    assert C.sizeof(_abi.rt_timing) == 24
    assert C.sizeof(_abi.rt_counters) == 8 * _abi.RT_COUNTER_SLOTS
    assert C.sizeof(_abi.rt_spheres_soa) == 48
    assert C.sizeof(_abi.rt_mesh_desc) == 80
    assert C.sizeof(_abi.rt_bvh_node) == 36


def test_no_cpu_fallback(lib):
    """Without a CUDA device rt_create must fail loudly (RT_ERR_NO_DEVICE), not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.rt_create(None, 0, C.byref(h))
    assert rc == 3 and not h.value
    assert b"no CPU path" in lib.rt_last_error(None)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gp1_raytracer_2223_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "rt_oracle" not in src and "oracle/" not in src.replace("oracle/ref_driver", ""), f


def test_library_is_built_from_the_sources_in_the_tree(lib):
    """build.ensure() stores a hash of csrc/* + flags beside the library and rebuilds on a mismatch."""
    assert not build.needs_build()
    assert build.built_hash() == build.source_hash()
    info = build.provenance()
    assert info["fresh"] and info["source_sha256_16"] == build.source_hash()[:16]
