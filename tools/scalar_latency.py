#!/usr/bin/env python3
"""Latency of the scalar entry points (reference src/lib.rs:309-339, 384-419: calc_V / calc_H) on a
GPU box: microseconds per call over N calls, through the facade and through the C ABI directly.

    python tools/scalar_latency.py [--n 10000] [--model doc]
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("INFLATOX_CACHE_DIR", os.path.join(ROOT, "tests", ".cubin_cache"))
os.environ.setdefault("INFLATOX_QUIET", "1")

import cases  # noqa: E402
from inflatox_b200 import _native  # noqa: E402
from inflatox_b200.consistency_conditions import GeneralisedAL  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10000)
    ap.add_argument("--models", nargs="*", default=["doc", "egno", "d5"])
    a = ap.parse_args()
    out = {}
    for model in a.models:
        al = GeneralisedAL(cases.artifact(model))
        p, ext = cases.params(model), cases.EXTENT[model]
        x = np.array([0.3 * ext[0] + 0.7 * ext[1], 0.6 * ext[2] + 0.4 * ext[3]])
        for _ in range(200):
            al.calc_V(x, p)
            al.calc_H(x, p)
        t0 = time.perf_counter()
        for _ in range(a.n):
            al.calc_V(x, p)
        t1 = time.perf_counter()
        for _ in range(a.n):
            al.calc_H(x, p)
        t2 = time.perf_counter()
        # the C ABI alone (no numpy conversions)
        lib = _native.lib()
        h = al.dylib._h if hasattr(al.dylib, "_h") else None
        abi = None
        if h is not None:
            dp = ctypes.POINTER(ctypes.c_double)
            v = ctypes.c_double()
            xp, pp = x.ctypes.data_as(dp), p.ctypes.data_as(dp)
            t3 = time.perf_counter()
            for _ in range(a.n):
                lib.inflx_potential(h, xp, 2, pp, p.size, ctypes.byref(v))
            abi = (time.perf_counter() - t3) / a.n * 1e6
        out[model] = {"calc_V_us": (t1 - t0) / a.n * 1e6, "calc_H_us": (t2 - t1) / a.n * 1e6,
                      "inflx_potential_c_abi_us": abi, "calls": a.n}
        print(model, json.dumps(out[model]), flush=True)


if __name__ == "__main__":
    main()
