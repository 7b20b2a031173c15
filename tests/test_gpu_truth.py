"""Conditioning check against a __float128 evaluation of the reference-generated C.

Where the CUDA result and the CPU oracle differ by more than 1e-10, neither is "right": the
expression itself loses the digits.  This test bounds the GPU's error against the quad-precision
truth by the CPU oracle's own error envelope: the GPU must be within 1e-10 of the truth on at
least as large a fraction of the points as the oracle (minus 0.5 %), and its median error must
not exceed twice the oracle's."""
import numpy as np
import pytest

import cases
import oracle
from inflatox_b200 import libinflx_rs as rs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model,n", [("angular", 64), ("egno", 48), ("d5", 32)])
def test_gpu_error_within_cpu_error_envelope(model, n):
    lib = rs.open_inflx_dylib(cases.artifact(model).shared_object_path, False)
    lib.set_devices([0])
    p, ext = cases.params(model), cases.EXTENT[model]
    gpu = np.zeros((n, n, 6))
    rs.complete_analysis(lib, p, gpu, np.array(ext).reshape(2, 2), False, 0)
    cpu = oracle.Oracle(model).complete_analysis(p, n, n, ext)
    truth = oracle.Oracle(model, quad=True).complete_analysis(p, n, n, ext)
    for k in range(6):
        eg, fg, _, _ = cases.rel_err(gpu[..., k], truth[..., k])
        ec, fc, _, _ = cases.rel_err(cpu[..., k], truth[..., k])
        both = fg & fc
        if not both.any():
            continue
        fg10, fc10 = (eg[both] <= 1e-10).mean(), (ec[both] <= 1e-10).mean()
        assert fg10 >= fc10 - 0.005, (model, k, fg10, fc10)
        assert np.median(eg[both]) <= 2 * np.median(ec[both]) + 1e-15, (model, k)


@pytest.mark.parametrize("model", ["angular", "egno", "d5"])
def test_residue_against_the_oracle_is_glibc_misrounding(model):
    """Attribution of what is left of the parity residue on the three ill-conditioned models.

    The CUDA path evaluates the hoisted libm calls (EGNO: pow(x, -3 alpha) once per row; d5: log per
    row, sin / cos per column) correctly rounded (csrc/inflx_crmath.cuh).  glibc's pow is not
    correctly rounded in ~1e-3 of its calls (tests/test_crmath.py), which on EGNO hits a handful
    of whole rows (angular: rows and columns, through the hoisted pow(x, n)) and is amplified past
    1e-10 by the model's cancellation.  Against the oracle variant whose libm IS correctly rounded
    - same generated C, same flags, same restated loop - every finite point agrees within 1e-10,
    NaN masks included (measured: 100 % on all planes of all three models, max 1.1e-11;
    profiles/parity_cr_r1.json)."""
    n = 512
    lib = rs.open_inflx_dylib(cases.artifact(model).shared_object_path, False)
    lib.set_devices([0])
    p, ext = cases.params(model), cases.EXTENT[model]
    gpu = np.zeros((n, n, 6))
    rs.complete_analysis(lib, p, gpu, np.array(ext).reshape(2, 2), False, 0)
    ref = oracle.Oracle(model).complete_analysis(p, n, n, ext)
    ref_cr = oracle.Oracle(model, libm="cr").complete_analysis(p, n, n, ext)
    for k in range(6):
        e, fin, nan_mm, inf_mm = cases.rel_err(gpu[..., k], ref[..., k])
        ec, finc, nan_mmc, inf_mmc = cases.rel_err(gpu[..., k], ref_cr[..., k])
        assert nan_mm == 0 and inf_mm == 0 and nan_mmc == 0 and inf_mmc == 0, (model, k)
        if not finc.any():
            continue
        frac, frac_cr = (e[fin] <= 1e-10).mean(), (ec[finc] <= 1e-10).mean()
        assert frac_cr >= 0.9999, (model, k, frac_cr)
        assert frac_cr >= frac, (model, k, frac, frac_cr)
        assert frac >= 0.985, (model, k, frac)  # the reference's own libm: a few rows per 512
