#!/usr/bin/env python3
"""GPU diagnostic: where a SHARD of C3 loses time against 1/N of the whole grid.

    python tools/shard_probe.py [model=egno] [rounds=7]

(1) rows per CTA: the 16384-column C3 grid, row shards of 16384 / 4096 / 2048 / 1024 rows (what one
    of 1 / 4 / 8 / 16 GPUs evaluates), with the engine's launch policy and with INFLATOX_RPT /
    INFLATOX_RPT_TAIL (tile height of the launch's last ~1.5 waves; 0 = uniform tiles) forced;
    median device time of the grid kernel and of the whole step, relative to rows/16384 of the
    full grid's time.  Variants are interleaved round-robin.
(2) host-issue gaps: the same step enqueued behind a 2 ms sleep kernel on a caller stream (every
    launch is queued before the GPU gets to the first one) against the step on an idle stream,
    both timed with events on that stream.  The difference is the time the GPU spends waiting for
    the host to issue the next launch of the step."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("INFLATOX_CACHE_DIR", os.path.join(ROOT, "tests", ".cubin_cache"))
os.environ.setdefault("INFLATOX_QUIET", "1")
import torch

import cases
from inflatox_b200 import libinflx_rs as rs


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "egno"
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    art = cases.artifact(model)
    lib = rs.open_inflx_dylib(art.shared_object_path, False)
    lib.set_devices([0])
    p, ext = cases.params(model), cases.EXTENT[model]
    n = 16384
    d = torch.empty(n * n * 6, dtype=torch.float64, device="cuda:0")

    def step(rows, rpt, tail=None, stream=None):
        for key, v in (("INFLATOX_RPT", rpt), ("INFLATOX_RPT_TAIL", tail)):
            if v is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = str(v)
        return rs.grid_eval(lib, "complete_analysis", p, None, n, n, ext, rows=(4096, 4096 + rows)
                            if rows < n else (0, n), device=0, out_device_ptr=d.data_ptr(), stream=stream)

    # (rows per CTA, rows per CTA of the launch's tail); None = the engine's policy, tail 0 = uniform
    combos = [(None, None), (None, 0), (16, 0), (16, 4), (16, 2), (8, 0), (8, 2), (4, 0), (4, 2)]
    variants = [(rows,) + c for rows in (16384, 4096, 2048, 1024) for c in combos]
    for v in variants:
        step(*v)
    t = {v: ([], []) for v in variants}
    for _ in range(rounds):
        for v in variants:
            rep = step(*v)
            t[v][0].append(rep["grid_ms"])
            t[v][1].append(rep["kernel_ms"])
    full = statistics.median(t[(16384, 16, 0)][0])
    print(f"{model} complete_analysis, 16384 columns; full grid kernel, uniform 16-row tiles: {full:.4f} ms")
    print("rows  rpt tail     grid_ms  vs ideal   step_ms  step-grid us")
    for (rows, rpt, tail), (g, k) in t.items():
        gm, km = statistics.median(g), statistics.median(k)
        print(f"{rows:5d} {str(rpt or 'auto'):>5} {str('auto' if tail is None else tail):>4} {gm:10.4f} "
              f"{gm / (full * rows / n):8.4f} {km:10.4f} {1e3 * (km - gm):8.1f}")

    # (2) host-issue gaps
    os.environ.pop("INFLATOX_RPT", None)
    os.environ.pop("INFLATOX_RPT_TAIL", None)
    s = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = {"idle": [], "queued": []}
    with torch.cuda.stream(s):
        for _ in range(rounds + 2):
            for mode in ("idle", "queued"):
                torch.cuda.synchronize()
                if mode == "queued":
                    torch.cuda._sleep(4_000_000)  # ~2 ms at 1.965 GHz
                e0.record(s)
                step(2048, None, None, stream=s.cuda_stream)
                e1.record(s)
                torch.cuda.synchronize()
                res[mode].append(e0.elapsed_time(e1))
    for mode, v in res.items():
        print(f"2048-row step on a caller stream, {mode}: median {statistics.median(v[2:]):.4f} ms")


if __name__ == "__main__":
    main()
