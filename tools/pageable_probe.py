"""Diagnostic: end-to-end throughput into a pageable numpy array (staging ring + parallel memcpy)
versus a pinned one (direct DMA)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("INFLATOX_CACHE_DIR", os.path.join(ROOT, "tests", ".cubin_cache"))
os.environ.setdefault("INFLATOX_QUIET", "1")
import numpy as np
import cases
from inflatox_b200 import libinflx_rs as rs

model, n = "egno", 8192
p, ext = cases.params(model), cases.EXTENT[model]
ss = np.array(ext).reshape(2, 2)
lib = rs.open_inflx_dylib(cases.artifact(model).shared_object_path, False)
lib.set_devices([0])
for tag, alloc in (("pinned", lambda: rs.pinned_empty((n, n, 6))), ("pageable, pre-faulted", lambda: np.ones((n, n, 6))),
                   ("pageable, fresh np.zeros each call", None)):
    out = alloc() if alloc else None
    for i in range(4):
        if alloc is None:
            out = np.zeros((n, n, 6))
        t0 = time.perf_counter()
        rs.complete_analysis(lib, p, out, ss, False, 0)
        dt = time.perf_counter() - t0
        if i:
            print(f"{tag}: {dt * 1e3:.1f} ms, {n * n / dt:.3e} points/s, {n * n * 48 / dt / 1e9:.1f} GB/s", flush=True)
