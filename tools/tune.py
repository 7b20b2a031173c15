#!/usr/bin/env python3
"""Kernel-only timing of compile-time variants (rows per thread, CTA width, register cap, fmad)
of the grid kernel on a GPU box.  Usage: python tools/tune.py egno complete_analysis 8192 ['[{"rpt": 16, "block": 128, "minb": 5, "extra": ["-DINFLX_RCP_NVCC"], "libm": "glibc-all"}]']"""
import itertools
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
    variants = json.loads(sys.argv[4]) if len(sys.argv) > 4 else None
    m = cases.load_model(model)
    cse = cases.golden_cse(model)
    per = 6 if op == "complete_analysis" else 1
    d = torch.empty(n * n * per, dtype=torch.float64, device="cuda:0")
    p, ext = cases.params(model), cases.EXTENT[model]
    if variants is None:
        variants = [dict(rpt=r, block=b, minb=mb, fmad=0) for r, b, mb in
                    itertools.product((2, 4, 8, 16), (64, 128, 256), (1, 2, 3, 4))]
    rows = []
    for v in variants:
        flags = ["--gpu-architecture=sm_100a", "--std=c++17",
                 f"--fmad={'true' if v.get('fmad') else 'false'}", "--prec-div=true",
                 "--prec-sqrt=true", "-lineinfo", f"-DINFLX_RPT={v['rpt']}",
                 f"-DINFLX_BLOCK={v["block"]}"] + ([f"-DINFLX_MIN_BLOCKS={v["minb"]}"] if "minb" in v else []) + v.get("extra", [])
        try:
            comp = ix.Compiler(m, silent=True, cse=cse, compiler_flags=flags)
            if "libm" in v:
                comp.libm = v["libm"]
            if "cols" in v:
                comp.cols_prepass = v["cols"]
            if "store" in v:
                comp.store_mode = v["store"]
            art = comp.compile()
        except Exception as e:
            print(v, "compile failed", str(e)[:200])
            continue
        if "run_rpt" in v:  # rows per CTA the engine uses at launch time (<= the compiled maximum)
            os.environ["INFLATOX_RPT"] = str(v["run_rpt"])
        else:
            os.environ.pop("INFLATOX_RPT", None)
        lib = rs.open_inflx_dylib(art.shared_object_path, False)
        lib.set_devices([0])
        ts = []
        for _ in range(5):
            rep = rs.grid_eval(lib, op, p, None, n, n, ext, device=0, out_device_ptr=d.data_ptr())
            ts.append(rep["grid_ms"])
        best = min(ts[1:])
        rows.append((best, v))
        print(f"{model} {op} {n}^2 {v}: {best:.3f} ms  {n * n / best / 1e6:.1f} Mpt/ms", flush=True)
    rows.sort(key=lambda t: t[0])
    print("BEST", rows[:3])


if __name__ == "__main__":
    main()
