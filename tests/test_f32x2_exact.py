"""The packed kernel's arithmetic claim, checked on the GPU itself: FFMA2 with run-time constants is
bit-identical to scalar FMUL / FADD / FSUB for arbitrary operands (tools/micro/f32x2_exact.cu)."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_ffma2_formulation_is_bit_exact(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this box")
    exe = str(tmp_path / "f32x2_exact")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "--fmad=false", "-ccbin", "/usr/bin/g++", "-o", exe,
                    os.path.join(ROOT, "tools", "micro", "f32x2_exact.cu")], check=True, capture_output=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "mismatches with FFMA2 + run-time constants: 0" in res.stdout.replace("runtime", "run-time"), res.stdout


def test_shared_reciprocal_normalise_is_bit_exact(tmp_path):
    """rt::normalize_and_invert (rt_device.cuh) shares one refined reciprocal between the three divisions of
    Vector3::Normalize and replaces nvcc's per-division range check by one test on the operands: it must equal
    __fsqrt_rn / __fdiv_rn / __frcp_rn bit for bit, inside its fast range, at the range's edges and outside
    (tools/micro/div_exact.cu: 2.5e9 random vectors incl. zeros, denormals, infinities, NaN)."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this box")
    exe = str(tmp_path / "div_exact")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "--fmad=false", "-ccbin", "/usr/bin/g++", "-o", exe,
                    os.path.join(ROOT, "tools", "micro", "div_exact.cu")], check=True, capture_output=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "mismatches against __fsqrt_rn / __fdiv_rn / __frcp_rn: 0 " in res.stdout, res.stdout
