"""Diagnostic: how fast can a multi-GB output block be page-locked?

Compares cudaHostAlloc with mmap(+MADV_HUGEPAGE) -> parallel pre-fault -> cudaHostRegister, and
checks that the D2H DMA rate into either kind of block is the same.  Feeds the design of
inflx_host_alloc (end-to-end section)."""
import ctypes, sys, time, threading
import torch

GB = 1 << 30
nbytes = int(float(sys.argv[1]) * GB) if len(sys.argv) > 1 else 12 * GB
nbytes = (nbytes + (1 << 21) - 1) & ~((1 << 21) - 1)
libc = ctypes.CDLL("libc.so.6", use_errno=True)
libc.mmap.restype = ctypes.c_void_p
libc.mmap.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long]
libc.munmap.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
libc.madvise.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
MADV_HUGEPAGE, MADV_POPULATE_WRITE = 14, 23
PROT_RW, MAP_PRIVATE_ANON = 3, 0x22
torch.cuda.init()
cudart_raw = ctypes.CDLL(torch.__path__[0] + "/../nvidia/cuda_runtime/lib/libcudart.so.12")
cudart_raw.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
cudart_raw.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
cudart_raw.cudaHostUnregister.argtypes = [ctypes.c_void_p]
cudart_raw.cudaFreeHost.argtypes = [ctypes.c_void_p]
cudart_raw.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
dev = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")


def d2h_rate(ptr):
    """GB/s of 1 GiB device->host copies cycling over the block."""
    n = min(nbytes // GB, 8)
    s = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record()
        for i in range(n):
            cudart_raw.cudaMemcpyAsync(ptr + i * GB, dev.data_ptr(), GB, 2, ctypes.c_void_p(s.cuda_stream))
        e1.record()
    s.synchronize()
    return n * GB / (e0.elapsed_time(e1) * 1e-3) / 1e9


def populate(ptr, n, threads):
    def work(k):
        lo = (n * k // threads) & ~((1 << 21) - 1)
        hi = (n * (k + 1) // threads) & ~((1 << 21) - 1) if k + 1 < threads else n
        if libc.madvise(ptr + lo, hi - lo, MADV_POPULATE_WRITE) != 0:
            ctypes.memset(ptr + lo, 0, hi - lo)
    ts = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    [t.start() for t in ts]; [t.join() for t in ts]


# A: cudaHostAlloc
t = time.perf_counter()
p = ctypes.c_void_p()
r = cudart_raw.cudaHostAlloc(ctypes.byref(p), nbytes, 1)
dt = time.perf_counter() - t
print(f"A cudaHostAlloc {nbytes / GB:.1f} GiB: rc={r} {dt:.2f} s ({nbytes / dt / 1e9:.1f} GB/s)  d2h {d2h_rate(p.value):.1f} GB/s", flush=True)
t = time.perf_counter(); cudart_raw.cudaFreeHost(p); print(f"  free {time.perf_counter() - t:.2f} s", flush=True)

for huge in (False, True):
    for threads in (1, 8, 16):
        t0 = time.perf_counter()
        ptr = libc.mmap(None, nbytes, PROT_RW, MAP_PRIVATE_ANON, -1, 0)
        if huge:
            libc.madvise(ptr, nbytes, MADV_HUGEPAGE)
        populate(ptr, nbytes, threads)
        t1 = time.perf_counter()
        r = int(cudart_raw.cudaHostRegister(ptr, nbytes, 1))
        t2 = time.perf_counter()
        rate = d2h_rate(ptr) if r == 0 else float("nan")
        t3 = time.perf_counter()
        cudart_raw.cudaHostUnregister(ptr)
        t4 = time.perf_counter()
        libc.munmap(ptr, nbytes)
        print(f"B mmap huge={huge} threads={threads}: populate {t1 - t0:.2f} s, register rc={r} {t2 - t1:.2f} s, "
              f"d2h {rate:.1f} GB/s, unregister {t4 - t3:.2f} s", flush=True)

# C: register in 1 GiB pieces (can be interleaved with use / aborted at exit); copies that span pieces
ptr = libc.mmap(None, nbytes, PROT_RW, MAP_PRIVATE_ANON, -1, 0)
libc.madvise(ptr, nbytes, MADV_HUGEPAGE)
populate(ptr, nbytes, 16)
t0 = time.perf_counter()
rcs = [int(cudart_raw.cudaHostRegister(ptr + o, min(GB, nbytes - o), 1)) for o in range(0, nbytes, GB)]
t1 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
rc = cudart_raw.cudaMemcpyAsync(ptr + GB // 2, dev.data_ptr(), GB, 2, None)  # spans two registrations
e1.record(); torch.cuda.synchronize()
print(f"C piecewise register: {t1 - t0:.2f} s rcs={set(rcs)}; spanning copy rc={rc} {GB / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s")
