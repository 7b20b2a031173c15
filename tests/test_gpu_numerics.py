"""Device helpers of csrc/inflx_device.cuh against the compiler's IEEE operators (GPU)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SRC = r"""
extern "C" __global__ void t_div(const double* a, const double* b, double* q, double* qe,
                                 unsigned char* bad, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool f = false;
  q[i] = inflx_div_s(a[i], b[i], f);
  bad[i] = f;
  qe[i] = a[i] / b[i];
}
extern "C" __global__ void t_sqrt(const double* a, double* q, double* qe, unsigned char* bad, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool f = false;
  q[i] = inflx_sqrt_s(a[i], f);
  bad[i] = f;
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
    a = np.abs(_random_doubles(rng, n))
    a[:5] = [0.0, np.inf, np.nan, -1.0, 4.0]
    q, qe, bad = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.uint8)
    mod.launch("t_sqrt", n, [a], [q, qe, bad])
    ok = bad == 0
    assert ok.mean() > 0.3
    assert _same(q[ok], qe[ok]).all()
    m = (a > 1e-100) & (a < 1e100)
    assert (bad[m] == 0).all()
    assert bad[0] == 1 and bad[1] == 1 and bad[2] == 1 and bad[3] == 1


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
    # half-integer powers: compare with a 200-bit evaluation
    import mpmath

    with mpmath.workprec(200):
        for k in range(0, n, 8):
            x = mpmath.mpf(float(a[k]))
            assert ph[k] == float(x * mpmath.sqrt(x))
            assert pmh[k] == float(1 / mpmath.sqrt(x))
