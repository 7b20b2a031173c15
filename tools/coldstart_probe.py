"""Diagnostic: cold-start cost of one facade call (pinned pool allocation vs staged pageable output)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("INFLATOX_CACHE_DIR", os.path.join(ROOT, "tests", ".cubin_cache"))
os.environ.setdefault("INFLATOX_QUIET", "1")
import numpy as np
t0 = time.perf_counter()
import cases
from inflatox_b200.consistency_conditions import GeneralisedAL
art = cases.artifact("egno")
t1 = time.perf_counter()
al = GeneralisedAL(art)
t2 = time.perf_counter()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
p, ext = cases.params("egno"), cases.EXTENT["egno"]
for k in range(3):
    t = time.perf_counter()
    out = al.complete_analysis(p, *ext, n, n)
    dt = time.perf_counter() - t
    print(f"PINNED={os.environ.get('INFLATOX_PINNED', '1')} call {k}: {dt * 1e3:.0f} ms ({n * n / dt:.3e} points/s)", flush=True)
    del out
print(f"import+compile {t1 - t0:.2f} s, open+basis check {t2 - t1:.2f} s")
