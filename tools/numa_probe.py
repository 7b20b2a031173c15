#!/usr/bin/env python3
"""Concurrent device->host bandwidth of all visible GPUs into page-locked memory placed (a) on the
NUMA node each GPU is attached to (inflx_host_alloc_on, the default) and (b) wherever the
allocating thread happens to run (INFLATOX_NUMA=0, round 1's behaviour).

    python tools/numa_probe.py [GiB per GPU = 2]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from inflatox_b200 import _native  # noqa: E402
from inflatox_b200 import libinflx_rs as rs  # noqa: E402


def run(devs, gib, tag):
    bufs = []
    for d in devs:
        rs.set_output_placement([d])
        host = rs.pinned_empty((gib << 30,), np.uint8)
        with torch.cuda.device(d):
            src = torch.empty(gib << 30, dtype=torch.uint8, device=f"cuda:{d}")
            bufs.append((d, src, torch.from_numpy(host), torch.cuda.Stream(device=d), host))
    for d, s, h, st, _ in bufs:
        with torch.cuda.stream(st):
            h.copy_(s, non_blocking=True)
    for d in devs:
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    for d, s, h, st, _ in bufs:
        with torch.cuda.stream(st):
            for _ in range(3):
                h.copy_(s, non_blocking=True)
    for d in devs:
        torch.cuda.synchronize(d)
    dt = time.perf_counter() - t0
    print(f"{tag} {devs}: {3 * gib * len(devs) * 1.0737 / dt:.1f} GB/s total, "
          f"{3 * gib * 1.0737 / dt:.1f} per GPU", flush=True)


def main():
    gib = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n = torch.cuda.device_count()
    print("numa node per device:", [_native.lib().inflx_device_numa_node(d) for d in range(n)])
    os.system("lscpu | grep -i -E 'numa|socket' | head -8; nvidia-smi topo -m | head -14")
    sets = [[0]] + ([[0, 1]] if n >= 2 else []) + ([[0, 1, 2, 3]] if n >= 4 else []) + \
        ([[0, 4], list(range(8))] if n >= 8 else [])
    for numa in ("1", "0"):
        os.environ["INFLATOX_NUMA"] = numa
        for devs in sets:
            run(devs, gib, f"INFLATOX_NUMA={numa}")


if __name__ == "__main__":
    main()
