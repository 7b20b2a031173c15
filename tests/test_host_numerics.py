"""Host-side accuracy of the epilogue's atan / tan helper (csrc/inflx_device.cuh `inflx_atan_tan`):
the device header compiled for the host (tests/native/cuda_host_shim.h) against libquadmath on
4 * 10^5 arguments over 60 decades - delta within 1 ulp of atan(y), T within 3 ulp of tan(delta) for
the delta actually returned (libdevice: 1.1 / 1.5 ulp).  The GPU twin with the real hardware seeds
is tests/test_gpu_numerics.py::test_atan_tan_helper_accuracy.  Also pins WHY the polynomial is
evaluated as Estrin pairs: an even / odd split has the same depth but loses more than a bit."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(tmp_path, chains: int, n: int) -> dict:
    exe = tmp_path / f"atan_check_{chains}"
    subprocess.run(
        ["g++", "-O1", "-std=c++17", "-ffp-contract=off", f"-DINFLX_ATAN_CHAINS={chains}",
         f"-I{ROOT}/tests/native",
         f'-DINFLX_DEVICE_HEADER="{ROOT}/inflatox_b200/csrc/inflx_device.cuh"',
         f"{ROOT}/tests/native/atan_check.cpp", "-o", str(exe), "-lquadmath"],
        check=True,
    )  # fmt: skip
    out = subprocess.run([str(exe), str(n)], capture_output=True, text=True, check=True).stdout
    return {k: float(v) for k, v in re.findall(r"(\w+)=([\d.]+)", out)}


@pytest.mark.parametrize("chains", [1, 2])
def test_atan_within_one_ulp_and_tan_within_three(tmp_path, chains):
    r = _run(tmp_path, chains, 400000)
    assert r["worst_atan_ulp"] <= 1.0, r
    assert r["worst_tan_ulp"] <= 3.0, r
