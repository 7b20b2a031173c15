#!/usr/bin/env python3
"""Parity + timing report of the CUDA path against the CPU oracle on a GPU box.

    python tools/gpu_parity_report.py [--n 512] [--quad] [--fmad] [--out gpurun_out/parity.json]

For each model: complete_analysis and consistency_only on an n x n grid over the reference tests'
extent, compared with oracle/ (the restated reference path): fraction of finite points within
1e-10 relative, NaN/inf mask mismatches, and optionally both implementations' errors against the
__float128 truth build.
"""
import argparse
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
import oracle  # noqa: E402
from inflatox_b200 import libinflx_rs as rs  # noqa: E402

NAMES6 = ["consistency", "eps_V", "eps_H", "eta", "delta", "omega"]


def stats(gpu, ref, truth=None):
    out = {}
    err, fin, nanmm, infmm = cases.rel_err(gpu, ref)
    e = err[fin]
    out["points"] = int(gpu.size)
    out["finite"] = int(fin.sum())
    out["nan_mask_mismatch"] = nanmm
    out["inf_mask_mismatch"] = infmm
    out["frac_within_1e-10"] = float((e <= 1e-10).mean()) if e.size else 1.0
    out["frac_bit_identical"] = float((gpu[fin] == ref[fin]).mean()) if e.size else 1.0
    out["median_rel"] = float(np.median(e)) if e.size else 0.0
    out["p99_rel"] = float(np.quantile(e, 0.99)) if e.size else 0.0
    out["max_rel"] = float(e.max()) if e.size else 0.0
    if truth is not None:
        eg, fg, _, _ = cases.rel_err(gpu, truth)
        ec, fc, _, _ = cases.rel_err(ref, truth)
        both = fg & fc
        out["vs_truth"] = {
            "gpu_median": float(np.median(eg[both])) if both.any() else 0.0,
            "cpu_median": float(np.median(ec[both])) if both.any() else 0.0,
            "gpu_p99": float(np.quantile(eg[both], 0.99)) if both.any() else 0.0,
            "cpu_p99": float(np.quantile(ec[both], 0.99)) if both.any() else 0.0,
            "gpu_frac_1e-10": float((eg[both] <= 1e-10).mean()) if both.any() else 1.0,
            "cpu_frac_1e-10": float((ec[both] <= 1e-10).mean()) if both.any() else 1.0,
            "gpu_worse_than_cpu_x10": float((eg[both] > 10 * ec[both] + 1e-13).mean()) if both.any() else 0.0,
        }
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--quad", action="store_true")
    ap.add_argument("--quad-n", type=int, default=96)
    ap.add_argument("--fmad", action="store_true")
    ap.add_argument("--cr", action="store_true",
                    help="also compare with the oracle variant whose libm is correctly rounded")
    ap.add_argument("--libm", default=None, help="libm flavour of the artefacts (cudagen.LIBM_FLAVOURS)")
    ap.add_argument("--no-timing", action="store_true")
    ap.add_argument("--models", nargs="*", default=list(cases.MODELS))
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    report = {"n": a.n, "fmad": a.fmad, "libm": a.libm or "glibc (default)", "models": {}}
    for m in a.models:
        t0 = time.time()
        art = cases.artifact(m, a.fmad, a.libm)
        lib = rs.open_inflx_dylib(art.shared_object_path, False)
        t_compile = time.time() - t0
        p, ext = cases.params(m), cases.EXTENT[m]
        ss = np.array([[ext[0], ext[1]], [ext[2], ext[3]]])
        n = a.n
        orc = oracle.Oracle(m)
        entry = {"compile_s": round(t_compile, 2), "flops_per_point": art.metadata["flops_per_point"]}
        # complete analysis
        out = np.zeros((n, n, 6))
        rs.complete_analysis(lib, p, out, ss, False, 0)
        ref = orc.complete_analysis(p, n, n, ext)
        truth = None
        entry["complete_analysis"] = {NAMES6[k]: stats(out[..., k], ref[..., k]) for k in range(6)}
        out1 = np.zeros((n, n))
        rs.consistency_only(lib, p, out1, ss, False, 0)
        entry["consistency_only"] = stats(out1, orc.consistency_only(p, n, n, ext))
        if a.cr:
            ref_cr = oracle.Oracle(m, libm="cr").complete_analysis(p, n, n, ext)
            entry["complete_analysis_vs_cr_libm_oracle"] = {
                NAMES6[k]: stats(out[..., k], ref_cr[..., k]) for k in range(6)
            }
        if a.quad:
            qn = a.quad_n
            oq = oracle.Oracle(m, quad=True)
            truth = oq.complete_analysis(p, qn, qn, ext)
            g = np.zeros((qn, qn, 6))
            rs.complete_analysis(lib, p, g, ss, False, 0)
            c = orc.complete_analysis(p, qn, qn, ext)
            entry["complete_analysis_vs_truth"] = {
                NAMES6[k]: stats(g[..., k], c[..., k], truth[..., k]) for k in range(6)
            }
        # timing: device-resident kernel-only via report, host end-to-end
        big = 4096
        if not a.no_timing:
            outb = rs.pinned_empty((big, big, 6))
            for _ in range(2):
                rep = rs.grid_eval(lib, "complete_analysis", p, outb, big, big, ss.reshape(4))
            entry["timing_4096"] = {
                "kernel_ms_incl_copies": rep["kernel_ms"],
                "total_ms": rep["total_ms"],
                "e2e_points_per_s": big * big / (rep["total_ms"] / 1e3),
                "launches": rep["launches"],
            }
        report["models"][m] = entry
        print(m, json.dumps(entry, indent=None)[:1500], flush=True)
    if a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        with open(a.out, "w") as fh:
            json.dump(report, fh, indent=1)


if __name__ == "__main__":
    main()
