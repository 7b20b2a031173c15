#!/usr/bin/env python3
"""Interleaved A/B timing of compile-time variants of one grid kernel on a GPU box: every round
times each variant once (device-resident output, CUDA events around the grid kernel), so clock /
thermal drift hits all variants alike; reports the median over the rounds and the ratio to the
first variant.

    python tools/ab.py egno complete_analysis 16384 '[{"name": "base"}, {"name": "rcp5", "extra": ["-DINFLX_RCP_NVCC"]}]' [rounds]

Variant keys: name, rpt (16), block (128), minb, fmad, extra (NVRTC flags), libm, cols, run_rpt."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("INFLATOX_CACHE_DIR", os.path.join(ROOT, "tests", ".cubin_cache"))
os.environ.setdefault("INFLATOX_QUIET", "1")
import numpy as np
import torch

import cases
import inflatox_b200 as ix
from inflatox_b200 import libinflx_rs as rs


def main():
    model, op, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
    variants = json.loads(sys.argv[4])
    rounds = int(sys.argv[5]) if len(sys.argv) > 5 else 7
    m = cases.load_model(model)
    cse = cases.golden_cse(model)
    per = 6 if op == "complete_analysis" else 1
    d = torch.empty(n * n * per, dtype=torch.float64, device="cuda:0")
    p, ext = cases.params(model), cases.EXTENT[model]
    libs = []
    for v in variants:
        flags = ["--gpu-architecture=sm_100a", "--std=c++17",
                 f"--fmad={'true' if v.get('fmad') else 'false'}", "--prec-div=true",
                 "--prec-sqrt=true", "-lineinfo", f"-DINFLX_RPT={v.get('rpt', 16)}",
                 f"-DINFLX_BLOCK={v.get('block', 128)}"]  # fmt: skip
        if "minb" in v:
            flags.append(f"-DINFLX_MIN_BLOCKS={v['minb']}")
        flags += v.get("extra", [])
        comp = ix.Compiler(m, silent=True, cse=cse, compiler_flags=flags)
        if "libm" in v:
            comp.libm = v["libm"]
        if "cols" in v:
            comp.cols_prepass = v["cols"]
        if "store" in v:
            comp.store_mode = v["store"]
        art = comp.compile()
        lib = rs.open_inflx_dylib(art.shared_object_path, False)
        lib.set_devices([0])
        libs.append((v, art, lib))
    times = {i: [] for i in range(len(libs))}
    outs = {}
    for r in range(rounds + 1):
        for i, (v, art, lib) in enumerate(libs):
            if "run_rpt" in v:
                os.environ["INFLATOX_RPT"] = str(v["run_rpt"])
            else:
                os.environ.pop("INFLATOX_RPT", None)
            rep = rs.grid_eval(lib, op, p, None, n, n, ext, device=0, out_device_ptr=d.data_ptr())
            if r:
                times[i].append(rep["grid_ms"])
            elif v.get("check", True):
                # checksum of the output bits: variants that claim to change no bit must agree
                outs[i] = int(d.view(torch.int64).sum().item())
    base = float(np.median(times[0]))
    for i, (v, art, lib) in enumerate(libs):
        t = float(np.median(times[i]))
        print(f"{model} {op} {n}^2 {v.get('name', json.dumps(v))}: median {t:.3f} ms  min {min(times[i]):.3f}  "
              f"x{t / base:.4f} vs first  bits {'same' if outs.get(i) == outs.get(0) else 'DIFFER'}",
              flush=True)


if __name__ == "__main__":
    main()
