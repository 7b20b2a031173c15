#!/usr/bin/env python3
"""GPU diagnostic: BASELINE C2 (angular, consistency_only, 4096^2) under forced launch geometries -
(rows per CTA, rows per CTA of the launch's tail), None = the engine's policy.  Prints the median
grid-kernel and whole-step device times in ms.  Run from the repo root: python tools/c2_probe.py

Round 2, one B200: policy 0.228 ms; uniform tiles 0.234; (16,0) 0.244; (16,4) 0.232; (8,2) 0.229;
(4,0) 0.242; (2,0) 0.279 - the engine's pick is the best of them."""
import os, sys, statistics
sys.path.insert(0, "."); sys.path.insert(0, "tests")
os.environ.setdefault("INFLATOX_CACHE_DIR", "tests/.cubin_cache"); os.environ.setdefault("INFLATOX_QUIET", "1")
import torch, cases
from inflatox_b200 import libinflx_rs as rs
art = cases.artifact("angular"); lib = rs.open_inflx_dylib(art.shared_object_path, False); lib.set_devices([0])
p, ext = cases.params("angular"), cases.EXTENT["angular"]
n = 4096
d = torch.empty(n * n, dtype=torch.float64, device="cuda:0")
fl = torch.empty(1 << 28, dtype=torch.uint8, device="cuda:0")
combos = [(None, None), (None, 0), (16, 0), (16, 4), (8, 0), (8, 2), (4, 0), (4, 2), (4, 1), (2, 0)]
t = {c: [] for c in combos}
for r in range(9):
    for c in combos:
        for k, v in (("INFLATOX_RPT", c[0]), ("INFLATOX_RPT_TAIL", c[1])):
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = str(v)
        fl.fill_(1); torch.cuda.synchronize()
        rep = rs.grid_eval(lib, "consistency_only", p, None, n, n, ext, device=0, out_device_ptr=d.data_ptr())
        if r > 1: t[c].append((rep["grid_ms"], rep["kernel_ms"]))
for c, v in t.items():
    print(c, round(statistics.median(x[0] for x in v), 4), round(statistics.median(x[1] for x in v), 4))
