"""Diagnostic: cold-start cost of the facade's first calls under each output-pool mode
(INFLATOX_PIN_MODE = deferred | eager | sync | off).  Usage: coldstart_probe.py [n] [pause_s]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("INFLATOX_CACHE_DIR", os.path.join(ROOT, "tests", ".cubin_cache"))
os.environ.setdefault("INFLATOX_QUIET", "1")
import numpy as np
t0 = time.perf_counter()
import cases
from inflatox_b200.consistency_conditions import GeneralisedAL, InflationCondition
art = cases.artifact("egno")
t1 = time.perf_counter()
al = GeneralisedAL.__new__(GeneralisedAL)
InflationCondition.__init__(al, art, validate_basis=False)
t2 = time.perf_counter()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
pause = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
p, ext = cases.params("egno"), cases.EXTENT["egno"]
mode = os.environ.get("INFLATOX_PIN_MODE", "deferred")
tb = time.perf_counter()
for k in range(4):
    t = time.perf_counter()
    out = al.complete_analysis(p, *ext, n, n)
    dt = time.perf_counter() - t
    print(f"mode={mode} pause={pause} call {k}: {dt * 1e3:.0f} ms ({n * n / dt:.3e} points/s)", flush=True)
    del out
    time.sleep(pause)
print(f"mode={mode}: 4 calls + pauses {time.perf_counter() - tb:.2f} s; import+compile {t1 - t0:.2f} s, open {t2 - t1:.2f} s", flush=True)
