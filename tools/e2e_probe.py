"""Diagnostic: end-to-end (host output) throughput of grid_eval per device / concurrently."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("INFLATOX_CACHE_DIR", os.path.join(ROOT, "tests", ".cubin_cache"))
os.environ.setdefault("INFLATOX_QUIET", "1")
import numpy as np
import cases
from inflatox_b200 import libinflx_rs as rs, _native

model, n = "egno", 16384
p, ext = cases.params(model), cases.EXTENT[model]
lib = rs.open_inflx_dylib(cases.artifact(model).shared_object_path, False)
ndev = _native.lib().inflx_device_count()
rank = int(os.environ.get("RANK", "-1"))
world = int(os.environ.get("WORLD_SIZE", "1"))
def run(devs, rows, tag):
    lib.set_devices(devs)
    out = rs.pinned_empty((rows[1] - rows[0], n, 6))
    for i in range(4):
        rep = rs.grid_eval(lib, "complete_analysis", p, out.reshape(-1), n, n, ext, rows=rows)
        if i:
            print(f"[rank {rank}] {tag} devs={devs} rows={rows}: {rep['total_ms']:.1f} ms, "
                  f"{rep['d2h_bytes'] / rep['total_ms'] / 1e6:.1f} GB/s, kernel_ms(incl waits) {rep['kernel_ms']:.1f}", flush=True)
if rank < 0:
    run([0], (0, n // 2), "half grid on one device")
    if ndev > 1:
        run([1], (0, n // 2), "half grid on one device")
        run([0, 1], (0, n), "full grid, in-process 2 devices")
else:
    import torch, torch.distributed as dist
    torch.cuda.set_device(rank)
    if os.environ.get("USE_NCCL", "1") == "1":
        dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo")
    dist.barrier()
    run([rank], (n * rank // world, n * (rank + 1) // world), f"torchrun world {world}")
    dist.barrier()
    dist.destroy_process_group()
