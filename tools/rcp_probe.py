#!/usr/bin/env python3
"""GPU diagnostic for the `INFLX_EXPERIMENT_RCP4` question: how accurate is the
MUFU.RCP64H seed, and how often does the division built on a 4-step reciprocal differ from the
IEEE quotient?  Usage (GPU box): python tools/rcp_probe.py [log2_n=26]

Prints, for n random operand pairs: max / mean relative error of the seed, and the number of
quotients that differ from `a / b` for the default (5-step) and the experimental (4-step)
reciprocal, among the pairs whose fast-path validity test passes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("INFLATOX_CACHE_DIR", os.path.join(ROOT, "tests", ".cubin_cache"))

SRC = r"""
__device__ __forceinline__ double rcp4(double b) {
  const double y0 = inflx_mufu_rcp64h(b);
  const double e = fma(y0, -b, 1.0);
  const double y1 = fma(y0, e, y0);
  const double e2 = fma(y1, -b, 1.0);
  return fma(y1, e2, y1);
}
extern "C" __global__ void probe(const double* a, const double* b, double* seed_err,
                                 unsigned char* d5, unsigned char* d4, unsigned char* flagged, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = a[i], y = b[i];
  const double y0 = inflx_mufu_rcp64h(y);
  seed_err[i] = fabs(fma(y0, -y, 1.0));          // |1 - b*y0| = relative error of the seed
  bool f5 = false, f4 = false;
  const double q5 = inflx_div_y(x, y, inflx_rcp_s(y), f5);
  const double q4 = inflx_div_y(x, y, rcp4(y), f4);
  const double q = x / y;
  flagged[i] = f5 || f4;
  d5[i] = !(f5) && (__double_as_longlong(q5) != __double_as_longlong(q));
  d4[i] = !(f4) && (__double_as_longlong(q4) != __double_as_longlong(q));
}
"""


def main():
    from gpu_kernels import Module

    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
    chunk = 1 << 22
    rng = np.random.default_rng(0)
    mod = Module(SRC)
    tot = dict(n=0, flagged=0, d5=0, d4=0, seed_max=0.0, seed_sum=0.0)
    for _ in range((1 << log2n) // chunk):
        # mantissas uniform, exponents moderate: the fast path's own domain
        a = np.ldexp(1.0 + rng.random(chunk), rng.integers(-60, 60, chunk)) * rng.choice([-1.0, 1.0], chunk)
        b = np.ldexp(1.0 + rng.random(chunk), rng.integers(-60, 60, chunk)) * rng.choice([-1.0, 1.0], chunk)
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        seed = np.zeros(chunk)
        d5, d4, fl = (np.zeros(chunk, dtype=np.uint8) for _ in range(3))
        mod.launch("probe", chunk, [a, b], [seed, d5, d4, fl])
        tot["n"] += chunk
        tot["flagged"] += int(fl.sum())
        tot["d5"] += int(d5.sum())
        tot["d4"] += int(d4.sum())
        tot["seed_max"] = max(tot["seed_max"], float(seed.max()))
        tot["seed_sum"] += float(seed.sum())
    print(f"pairs {tot['n']}, flagged {tot['flagged']}")
    print(f"MUFU.RCP64H seed: max rel. error {tot['seed_max']:.3e} (2^{np.log2(tot['seed_max']):.2f}), "
          f"mean {tot['seed_sum'] / tot['n']:.3e}")
    print(f"quotients != a/b: 5-step reciprocal {tot['d5']}, 4-step reciprocal {tot['d4']}")


if __name__ == "__main__":
    main()
