"""Diagnostic: concurrent device->host bandwidth of the visible GPUs into pinned memory."""
import os, sys, time, threading
import torch
n = torch.cuda.device_count()
os.system("nvidia-smi topo -m | head -20; lscpu | grep -i -E 'numa|socket|model name' | head; ")
for d in range(n):
    bdf = torch.cuda.get_device_properties(d).pci_bus_id if hasattr(torch.cuda.get_device_properties(d), 'pci_bus_id') else None
    print(d, bdf)
GB = 4
def run(devs, tag):
    bufs = []
    for d in devs:
        with torch.cuda.device(d):
            src = torch.empty(GB << 30, dtype=torch.uint8, device=f"cuda:{d}")
            dst = torch.empty(GB << 30, dtype=torch.uint8).pin_memory()
            bufs.append((d, src, dst, torch.cuda.Stream(device=d)))
    for d, s, h, st in bufs:
        with torch.cuda.stream(st):
            h.copy_(s, non_blocking=True)
    for d in devs: torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    for d, s, h, st in bufs:
        with torch.cuda.stream(st):
            for _ in range(3): h.copy_(s, non_blocking=True)
    for d in devs: torch.cuda.synchronize(d)
    dt = time.perf_counter() - t0
    print(tag, devs, f"{3 * GB * len(devs) * 1.0737 / dt:.1f} GB/s total, {3 * GB * 1.0737 / dt:.1f} per GPU", flush=True)
run([0], "single")
if n >= 2:
    run([1], "single")
    run([0, 1], "pair")
if n >= 4:
    run([0, 2], "pair"); run([0, 1, 2, 3], "quad")
if n >= 8:
    run([0, 4], "pair"); run(list(range(8)), "all8")
