"""Device helpers of csrc/inflx_device.cuh against the compiler's IEEE operators (GPU)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SRC = r"""
extern "C" __global__ void t_div(const double* a, const double* b, double* q, double* qe,
                                 unsigned char* bad, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  inflx_chk f;
  q[i] = inflx_div_s(a[i], b[i], f);
  bad[i] = f.any();
  qe[i] = a[i] / b[i];
}
extern "C" __global__ void t_sqrt(const double* a, double* q, double* qe, unsigned char* bad, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  inflx_chk f;
  q[i] = inflx_sqrt_s(a[i], f);
  bad[i] = f.any();
  qe[i] = sqrt(a[i]);
}
extern "C" __global__ void t_pow(const double* a, double* p3, double* p4, double* p7, double* pm2,
                                 double* ph, double* pmh, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = a[i];
  p3[i] = inflx_powi<3>(x);
  p4[i] = inflx_powi<4>(x);
  p7[i] = inflx_powi<7>(x);
  pm2[i] = inflx_powi_neg<2>(x, inflx_exact());
  ph[i] = inflx_powh<1>(x, inflx_exact());
  pmh[i] = inflx_powh_neg<0>(x, inflx_exact());
}
extern "C" __global__ void t_atan_tan(const double* y, const double* yinv, double* d, double* t,
                                      double* d_ref, double* t_ref, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  inflx_atan_tan(y[i], yinv[i], d[i], t[i], inflx_exact());
  d_ref[i] = atan(y[i]);       // libdevice, for comparison
  t_ref[i] = tan(d_ref[i]);
}
"""


@pytest.fixture(scope="module")
def mod():
    from gpu_kernels import Module

    return Module(SRC)


def _random_doubles(rng, n):
    """mixture: wide-exponent bit patterns, moderate values, specials"""
    bits = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    wide = bits.view(np.float64).copy()
    moderate = rng.standard_normal(n) * 10.0 ** rng.uniform(-12, 12, n)
    out = np.where(rng.random(n) < 0.5, wide, moderate)
    specials = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 5e-324, 2.2e-308,
                         1.7e308, 1e-300, 1e300, 3.0, 1 / 3])
    out[: specials.size] = specials
    return np.ascontiguousarray(out)


def _same(a, b):
    return (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))


def test_speculative_division_is_ieee_whenever_it_claims_so(mod):
    rng = np.random.default_rng(1)
    n = 1 << 22
    a, b = _random_doubles(rng, n), _random_doubles(rng, n)[::-1].copy()
    q, qe, bad = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.uint8)
    mod.launch("t_div", n, [a, b], [q, qe, bad])
    ok = bad == 0
    assert ok.mean() > 0.3  # the fast path must actually be taken for ordinary operands
    assert _same(q[ok], qe[ok]).all()
    # ordinary operands never need the slow path
    m = np.abs(np.log10(np.abs(a) + 1e-300)) < 100
    m &= np.abs(np.log10(np.abs(b) + 1e-300)) < 100
    m &= np.isfinite(a) & np.isfinite(b) & (a != 0) & (b != 0)
    assert (bad[m] == 0).all()


def test_speculative_sqrt_is_ieee_whenever_it_claims_so(mod):
    rng = np.random.default_rng(2)
    n = 1 << 22
    a = _random_doubles(rng, n)
    a[: n // 2] = np.abs(a[: n // 2])
    a[:8] = [0.0, np.inf, np.nan, -0.0, 4.0, -1.0, -1e-320, -np.inf]
    q, qe, bad = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.uint8)
    mod.launch("t_sqrt", n, [a], [q, qe, bad])
    ok = bad == 0
    assert ok.mean() > 0.3
    assert _same(q[ok], qe[ok]).all()
    m = (a > 1e-100) & (a < 1e100)
    assert (bad[m] == 0).all()
    # the test is on |high word| read as a float: conservative beyond 2^1017 (nvcc: 2^1024)
    # zero, inf, -0 and NaN take the slow path (a NaN argument means a NaN upstream, whose quotients
    # have flagged the point anyway); a negative argument of ordinary magnitude is NaN on the fast
    # path too and must NOT be flagged (omega / eta are NaN by design on 10-60 % of a grid)
    assert bad[0] == 1 and bad[1] == 1 and bad[2] == 1 and bad[3] == 1
    assert bad[5] == 0 and np.isnan(q[[2, 5, 7]]).all()
    neg = (a < -1e-290) & (a > -1e300)
    assert neg.sum() > 1000 and (bad[neg] == 0).all() and np.isnan(q[neg]).all()


def test_double_double_powers_are_correctly_rounded(mod):
    """x^n / x^(n+1/2) helpers against exact rational arithmetic (python fractions)."""
    from fractions import Fraction
    import math

    rng = np.random.default_rng(3)
    n = 4096
    a = np.ascontiguousarray(np.abs(rng.standard_normal(n)) * 10.0 ** rng.uniform(-3, 3, n) + 1e-6)
    outs = [np.zeros(n) for _ in range(6)]
    mod.launch("t_pow", n, [a], outs)
    p3, p4, p7, pm2, ph, pmh = outs

    def nearest(fr: Fraction) -> float:
        return float(fr)  # Fraction -> float is correctly rounded

    for k in range(0, n, 8):
        x = Fraction(float(a[k]))
        assert p3[k] == nearest(x**3)
        assert p4[k] == nearest(x**4)
        assert p7[k] == nearest(x**7)
        assert pm2[k] == nearest(1 / x**2)
    # irregular arguments of the half-integer powers (exact policy = what the slow path runs)
    special = np.array([0.0, np.inf, -1.0, np.nan, 1e-320, 1e300, 4.0])
    outs2 = [np.zeros(special.size) for _ in range(6)]
    mod.launch("t_pow", special.size, [special], outs2)
    ph2, pmh2 = outs2[4], outs2[5]
    assert ph2[0] == 0.0 and ph2[1] == np.inf and np.isnan(ph2[2]) and np.isnan(ph2[3])
    assert pmh2[0] == np.inf and pmh2[1] == 0.0 and np.isnan(pmh2[2]) and np.isnan(pmh2[3])
    assert ph2[6] == 8.0 and pmh2[6] == 0.5
    assert abs(pmh2[4] / (1e-320 ** -0.5) - 1) < 1e-15 and abs(pmh2[5] / 1e-150 - 1) < 1e-15
    # half-integer powers: compare with a 200-bit evaluation
    import mpmath

    with mpmath.workprec(200):
        for k in range(0, n, 8):
            x = mpmath.mpf(float(a[k]))
            assert ph[k] == float(x * mpmath.sqrt(x))
            assert pmh[k] == float(1 / mpmath.sqrt(x))


def test_atan_tan_helper_accuracy(mod):
    """delta = atan(y) within 1 ulp; T within 3 ulp of tan(delta) for the delta actually returned
    (200-bit mpmath references), over 600 decades of y and with the 1/y input off by up to 1 ulp
    the way the epilogue's own quotient is."""
    import mpmath

    rng = np.random.default_rng(4)
    n = 6000
    y = 10.0 ** rng.uniform(-30, 30, n)
    y[:2000] = 10.0 ** rng.uniform(-2, 2, 2000)
    y[2000:2200] = 10.0 ** rng.uniform(14, 18, 200)  # delta rounds to within a few ulp of pi/2
    y[-6:] = [0.0, 1.0, np.inf, 1.0000000000000002, 0.9999999999999999, 1e300]
    yinv = 1.0 / y
    yinv = np.where(rng.random(n) < 0.5, np.nextafter(yinv, np.inf), yinv)
    yinv = np.where(rng.random(n) < 0.3, np.nextafter(yinv, 0), yinv)
    yinv[~np.isfinite(1.0 / y)] = 0.0
    yinv[y == 0] = np.inf
    outs = [np.zeros(n) for _ in range(4)]
    mod.launch("t_atan_tan", n, [np.ascontiguousarray(y), np.ascontiguousarray(yinv)], outs)
    d, t, d_ref, t_ref = outs
    worst_d = worst_t = worst_d_lib = worst_t_lib = 0.0
    with mpmath.workprec(250):
        for k in range(n):
            yy = mpmath.mpf(float(y[k])) if np.isfinite(y[k]) else mpmath.inf
            true_d = mpmath.atan(yy)
            ulp_d = np.spacing(float(true_d))
            worst_d = max(worst_d, abs(float((mpmath.mpf(float(d[k])) - true_d) / ulp_d)))
            worst_d_lib = max(worst_d_lib, abs(float((mpmath.mpf(float(d_ref[k])) - true_d) / ulp_d)))
            for dd, tt, which in ((d[k], t[k], 0), (d_ref[k], t_ref[k], 1)):
                true_t = mpmath.tan(mpmath.mpf(float(dd)))
                e = abs(float((mpmath.mpf(float(tt)) - true_t) / np.spacing(abs(float(true_t)) or 5e-324)))
                if which == 0:
                    worst_t = max(worst_t, e)
                else:
                    worst_t_lib = max(worst_t_lib, e)
    print(f"atan: ours {worst_d:.2f} ulp, libdevice {worst_d_lib:.2f} ulp; "
          f"tan(delta): ours {worst_t:.2f} ulp, libdevice {worst_t_lib:.2f} ulp")
    assert worst_d <= 1.0, worst_d
    assert worst_t <= 3.0, worst_t
    assert d[-6] == 0.0 and t[-6] == 0.0  # y = 0
    assert d[-4] == 1.5707963267948966 and abs(t[-4] / 1.633123935319537e16 - 1) < 1e-15  # y = inf


# ---------------------------------------------------------------------------------------------
# correctly rounded libm subset (csrc/inflx_crmath.cuh): the device build must return the bits of
# the host build of the same file, which tests/test_crmath.py proves correctly rounded
# ---------------------------------------------------------------------------------------------
CR_SRC = r"""
extern "C" __global__ void t_cr(const double* x, const double* y, double* p, double* l, double* e,
                                double* s, double* c, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  p[i] = inflx_cr_pow(x[i], y[i]);
  l[i] = inflx_cr_log(x[i]);
  e[i] = inflx_cr_exp(y[i]);
  s[i] = inflx_cr_sin(y[i]);
  c[i] = inflx_cr_cos(y[i]);
}
"""
CR_HOST = r"""
#include INFLX_CRMATH_HEADER
void host_cr(const double* x, const double* y, double* p, double* l, double* e, double* s,
             double* c, long n) {
  for (long i = 0; i < n; i++) {
    p[i] = inflx_cr_pow(x[i], y[i]);
    l[i] = inflx_cr_log(x[i]);
    e[i] = inflx_cr_exp(y[i]);
    s[i] = inflx_cr_sin(y[i]);
    c[i] = inflx_cr_cos(y[i]);
  }
}
"""


@pytest.mark.parametrize("fmad", [False, True])
def test_correctly_rounded_libm_matches_the_host_build(tmp_path, fmad):
    import ctypes
    import os
    import subprocess

    from gpu_kernels import ROOT, Module

    header = os.path.join(ROOT, "inflatox_b200", "csrc", "inflx_crmath.cuh")
    src, so = tmp_path / "host_cr.c", tmp_path / "host_cr.so"
    src.write_text(CR_HOST)
    subprocess.run(
        ["gcc", "-O2", "-march=native", "-ffp-contract=off", "-shared", "-fPIC",
         f'-DINFLX_CRMATH_HEADER="{header}"', str(src), "-o", str(so), "-lm"], check=True,
    )  # fmt: skip
    host = ctypes.CDLL(str(so))
    rng = np.random.default_rng(5)
    n = 1 << 18
    x = np.ascontiguousarray(np.exp(rng.uniform(-30, 30, n)))
    y = np.ascontiguousarray(rng.uniform(-20, 20, n))
    x[:4], y[:4] = [0.47, 0.5, 1.0, 25098.05], [-1.5, -3.0, 7.0, 0.3]  # EGNO / d5 magnitudes
    x[4:10] = [0.0, -1.0, np.inf, np.nan, 5e-324, 1e308]   # irregular: libm / libdevice path
    outs_d = [np.zeros(n) for _ in range(5)]
    outs_h = [np.zeros(n) for _ in range(5)]
    with open(header) as fh:
        mod = Module(fh.read() + CR_SRC, fmad=fmad)
    mod.launch("t_cr", n, [x, y], outs_d)
    dp = ctypes.POINTER(ctypes.c_double)
    host.host_cr(*[a.ctypes.data_as(dp) for a in [x, y] + outs_h], ctypes.c_long(n))
    regular = np.ones(n, dtype=bool)
    regular[4:10] = False
    for name, d, h in zip(("pow", "log", "exp", "sin", "cos"), outs_d, outs_h):
        same = _same(d, h)
        assert same[regular].all(), (name, int((~same[regular]).sum()))
    # irregular arguments: same class of result (NaN / inf / zero) as libm
    for d, h in zip(outs_d[:2], outs_h[:2]):
        assert (np.isnan(d[4:10]) == np.isnan(h[4:10])).all()
        assert (np.isinf(d[4:10]) == np.isinf(h[4:10])).all()


# ---------------------------------------------------------------------------------------------
# the reference host's libm restated (csrc/inflx_glibcmath.cuh): the device build must return the
# bits of the HOST'S libm itself (tests/test_glibcmath.py proves the host build does)
# ---------------------------------------------------------------------------------------------
GL_SRC = r"""
extern "C" __global__ void t_gl(const double* x, const double* y, double* p, double* l, double* e,
                                double* s, double* c, double* t, double* m, double* a, double* q,
                                int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  a[i] = inflx_gl_atan(x[i]);
  q[i] = inflx_gl_tan(inflx_gl_atan(fabs(y[i])));  // the epilogue's composition: argument in [0, pi/2]
  p[i] = inflx_gl_pow(x[i], y[i]);
  l[i] = inflx_gl_log(x[i]);
  e[i] = inflx_gl_exp(y[i]);
  s[i] = inflx_gl_sin(y[i]);
  c[i] = inflx_gl_cos(y[i]);
  t[i] = inflx_gl_tanh(y[i]);
  m[i] = inflx_gl_expm1(y[i]);
}
"""
GL_HOST = r"""
#include <math.h>
void host_libm(const double* x, const double* y, double* p, double* l, double* e, double* s,
               double* c, double* t, double* m, double* a, double* q, long n) {
  for (long i = 0; i < n; i++) {
    a[i] = atan(x[i]);
    q[i] = tan(atan(fabs(y[i])));
    p[i] = pow(x[i], y[i]);
    l[i] = log(x[i]);
    e[i] = exp(y[i]);
    s[i] = sin(y[i]);
    c[i] = cos(y[i]);
    t[i] = tanh(y[i]);
    m[i] = expm1(y[i]);
  }
}
"""


@pytest.mark.parametrize("fmad", [False, True])
def test_glibc_libm_on_the_device_has_the_bits_of_the_host_libm(tmp_path, fmad):
    import ctypes
    import os
    import subprocess

    from gpu_kernels import ROOT, Module
    from test_glibcmath import _has_fma, _host_glibc

    if not (_host_glibc() == "2.39" and _has_fma()):
        pytest.skip("bit identity is pinned to glibc 2.39's FMA ifunc variants")
    csrc = os.path.join(ROOT, "inflatox_b200", "csrc")
    src, so = tmp_path / "host_libm.c", tmp_path / "host_libm.so"
    src.write_text(GL_HOST)
    subprocess.run(["gcc", "-O2", "-fno-builtin", "-shared", "-fPIC", str(src), "-o", str(so), "-lm"],
                   check=True)  # fmt: skip
    host = ctypes.CDLL(str(so))
    rng = np.random.default_rng(6)
    n = 1 << 21
    q = n // 4
    x = np.empty(n)
    y = np.empty(n)
    x[:q], y[:q] = np.exp(rng.uniform(-30, 30, q)), rng.uniform(-20, 20, q)
    x[q:2 * q], y[q:2 * q] = rng.uniform(0.4, 0.6, q), -3.0 * rng.uniform(0.05, 2.0, q)  # EGNO rows
    x[2 * q:3 * q] = rng.uniform(0, 36, q)                                               # d5 rows
    y[2 * q:3 * q] = np.where(rng.random(q) < 0.5, rng.integers(1, 9, q) * 0.5, 4 * np.pi * rng.random(q))
    x[3 * q:], y[3 * q:] = _random_doubles(rng, n - 3 * q), _random_doubles(rng, n - 3 * q)[::-1]
    x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
    outs_d = [np.zeros(n) for _ in range(9)]
    outs_h = [np.zeros(n) for _ in range(9)]
    with open(os.path.join(csrc, "inflx_glibc_tables.cuh")) as fh:
        tables = fh.read()
    with open(os.path.join(csrc, "inflx_glibcmath.cuh")) as fh:
        header = fh.read().replace('#include "inflx_glibc_tables.cuh"', tables)
    mod = Module(header + GL_SRC, fmad=fmad)
    mod.launch("t_gl", n, [x, y], outs_d)
    dp = ctypes.POINTER(ctypes.c_double)
    host.host_libm(*[a.ctypes.data_as(dp) for a in [x, y] + outs_h], ctypes.c_long(n))
    names = ("pow", "log", "exp", "sin", "cos", "tanh", "expm1", "atan", "tan(atan)")
    for name, d, h in zip(names, outs_d, outs_h):
        same = _same(d, h)
        bad = np.flatnonzero(~same)
        assert same.all(), (name, bad.size, x[bad[:3]], y[bad[:3]], d[bad[:3]], h[bad[:3]])
