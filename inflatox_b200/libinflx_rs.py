"""Drop-in for the reference's Rust extension module `inflatox.libinflx_rs`.

Exports exactly the names the reference registers in its pyo3 module (reference
src/lib.rs:68-92) with the same argument order, shapes and error classes, implemented as a thin
ctypes binding of the C ABI in include/inflx_b200.h (engine: csrc/inflx_engine.cpp, CUDA driver
API).  Caller-allocated numpy outputs are filled in place, as the reference does through
`PyReadwriteArray*`.  `solve_eom_rk4` / `solve_eom_rkf` (the sequential background ODE solver,
reference src/background_solver.rs) are outside the grid-evaluation path and raise
NotImplementedError.
"""
from __future__ import annotations

import ctypes
import os
import sys
import threading
import time

import numpy as np

from . import _native

__all__ = [
    "InflatoxPyDyLib", "open_inflx_dylib", "log_info", "log_warn", "flag_quantum_dif_py",
    "consistency_only", "consistency_rapidturn_only", "epsilon_v_only", "complete_analysis",
    "complete_analysis_on_trajectory", "consistency_only_on_trajectory",
    "consistency_rapidturn_only_on_trajectory", "epsilon_v_only_on_trajectory", "solve_eom_rk4",
    "solve_eom_rkf", "PanicException", "pinned_empty", "host_output", "sweep", "grid_eval",
]  # fmt: skip

_DP = ctypes.POINTER(ctypes.c_double)


class PanicException(BaseException):
    """Stands in for pyo3_runtime.PanicException: the reference `panic!`s on non-contiguous
    arrays (reference src/anguelova.rs:189-191, 210-212)."""


def log_info(msg: str) -> None:  # reference src/lib.rs:94-97
    print(f"\033[1;35m[Inflatox Info]\033[0m\n{msg}", file=sys.stderr)


def log_warn(msg: str) -> None:  # reference src/lib.rs:99-102
    print(f"\033[1;33m[Inflatox Warning]\033[0m\n{msg}", file=sys.stderr)


def _f64_in(a, what: str) -> np.ndarray:
    """Read-only fp64 input.  Lists are accepted (more lenient than pyo3); an ndarray must be
    C-contiguous like the reference demands."""
    if isinstance(a, np.ndarray):
        if a.dtype != np.float64:
            raise TypeError(f"{what} must be a float64 array, got {a.dtype}")
        if not a.flags.c_contiguous:
            raise PanicException(f"{what.upper()} SHOULD BE C-CONTIGUOUS")
        return a
    return np.ascontiguousarray(a, dtype=np.float64)


def _out(a, dtype, what: str) -> np.ndarray:
    if not isinstance(a, np.ndarray) or a.dtype != dtype:
        raise TypeError(f"{what} must be a numpy array of dtype {np.dtype(dtype)}")
    if not a.flags.c_contiguous:
        raise PanicException("OUTPUT ARRAY SHOULD BE C-CONTIGUOUS")
    if not a.flags.writeable:
        raise TypeError(f"{what} is read-only")
    return a


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(_DP)


class _EngineGate:
    """Counts engine calls in flight.  Page-locking a multi-GB block holds the CUDA driver's lock
    for seconds, so the background pin jobs of the output pool (below) wait here until the
    engine is idle instead of stalling the very call they were started from."""

    def __init__(self):
        self.cond = threading.Condition()
        self.busy = 0
        self.completed = 0
        self.closing = False

    def __enter__(self):
        with self.cond:
            self.busy += 1

    def __exit__(self, *exc):
        with self.cond:
            self.busy -= 1
            self.completed += 1
            self.cond.notify_all()

    def wait_idle_after(self, completed: int, timeout: float) -> None:
        """Block until a call newer than `completed` has finished and none is running (or until
        `timeout` seconds have passed: the caller may never use the array it asked for)."""
        with self.cond:
            self.cond.wait_for(
                lambda: self.closing or (self.busy == 0 and self.completed > completed), timeout
            )

    def close(self) -> None:
        with self.cond:
            self.closing = True
            self.cond.notify_all()


_gate = _EngineGate()
try:  # wake waiting pin jobs when the interpreter starts to shut down (before threads are joined)
    threading._register_atexit(_gate.close)
except Exception:  # private hook missing: the jobs' own timeout bounds the wait
    pass


def _ss(start_stop) -> np.ndarray:
    ss = _f64_in(start_stop, "start_stop array")
    if ss.ndim != 2:
        raise TypeError("start_stop must be a 2D float64 array")
    return ss


class InflatoxPyDyLib:
    """Handle of an opened model artefact (reference src/lib.rs:104-106, methods :205-463)."""

    def __init__(self, handle: ctypes.c_void_p, path: str):
        self._h = handle
        self.path = path

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _native.lib().inflx_close(h)
            except Exception:
                pass

    # -- metadata -----------------------------------------------------------------------------
    @property
    def n_fields(self) -> int:
        return int(_native.lib().inflx_n_fields(self._h))

    @property
    def n_parameters(self) -> int:
        return int(_native.lib().inflx_n_parameters(self._h))

    @property
    def name(self) -> str:
        return _native.lib().inflx_model_name(self._h).decode()

    def set_devices(self, ordinals) -> None:
        arr = (ctypes.c_int * len(ordinals))(*ordinals)
        _native.raise_for_status(_native.lib().inflx_set_devices(self._h, arr, len(ordinals)))
        set_output_placement(self.devices())  # pinned outputs: one NUMA-local slice per device

    def devices(self) -> list[int]:
        buf = (ctypes.c_int * 64)()
        n = _native.lib().inflx_get_devices(self._h, buf, 64)
        return [buf[i] for i in range(min(n, 64))]

    # -- reference methods --------------------------------------------------------------------
    def validate_basis_on_domain(self, num_points, p, start_stop, accuracy) -> None:
        num_points = np.ascontiguousarray(num_points, dtype=np.uint32)
        p, ss = _f64_in(p, "parameter array"), _ss(start_stop)
        rc = _native.lib().inflx_validate_basis_on_domain(
            self._h, num_points.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), num_points.size,
            _dptr(p), p.size, _dptr(ss), ss.shape[0], ss.shape[1], float(accuracy),
        )  # fmt: skip
        _native.raise_for_status(rc)

    def potential(self, x, p) -> float:
        x, p = _f64_in(x, "field-space array"), _f64_in(p, "parameter array")
        if x.ndim != 1 or p.ndim != 1:
            x_len = x.size if x.ndim == 1 else 0  # forces the reference's Shape error
            p_len = p.size if p.ndim == 1 else (1 << 62)
        else:
            x_len, p_len = x.size, p.size
        val = ctypes.c_double()
        rc = _native.lib().inflx_potential(
            self._h, _dptr(x), x_len, _dptr(p), p_len, ctypes.byref(val)
        )
        _native.raise_for_status(rc)
        return val.value

    def potential_array(self, x, p, start_stop) -> None:
        x = _out(x, np.float64, "x")
        p, ss = _f64_in(p, "parameter array"), _ss(start_stop)
        if x.ndim != self.n_fields:  # reference src/lib.rs:355-363
            raise Exception(
                f"Expected array with shape [], received array with shape {list(x.shape)}. "
                "Context: expected an array with with the same number of axes as there are "
                "field-space coordinates"
            )
        with _gate:
            rc = _native.lib().inflx_potential_array(
                self._h, _dptr(x), x.shape[0], x.shape[1], _dptr(p), p.size, _dptr(ss),
                ss.shape[0], ss.shape[1],
            )  # fmt: skip
        _native.raise_for_status(rc)

    def hesse(self, x, p) -> np.ndarray:
        x, p = _f64_in(x, "field-space array"), _f64_in(p, "parameter array")
        out = np.zeros((2, 2), dtype=np.float64)
        rc = _native.lib().inflx_hesse(self._h, _dptr(x), x.size, _dptr(p), p.size, _dptr(out))
        _native.raise_for_status(rc)
        return out

    def hesse_array(self, nx, p, start_stop) -> np.ndarray:
        nx = np.atleast_1d(np.asarray(nx)).astype(np.int64)
        p, ss = _f64_in(p, "parameter array"), _ss(start_stop)
        if nx.size != self.n_fields:  # reference src/lib.rs:437-444
            raise Exception(
                f"Expected array with shape [{self.n_fields}], received array with shape "
                f"[{nx.size}]. Context: expected a 1D array with as many elements as there are "
                "field-space coordinates"
            )
        out = np.zeros((2, 2, int(nx[0]), int(nx[1])), dtype=np.float64)
        with _gate:
            rc = _native.lib().inflx_hesse_array(
                self._h, _dptr(out), int(nx[0]), int(nx[1]), _dptr(p), p.size, _dptr(ss),
                ss.shape[0], ss.shape[1],
            )  # fmt: skip
        _native.raise_for_status(rc)
        return out


def open_inflx_dylib(lib_path: str, check_basis: bool) -> InflatoxPyDyLib:
    """reference src/lib.rs:108-115"""
    h = ctypes.c_void_p()
    rc = _native.lib().inflx_open(str(lib_path).encode(), int(bool(check_basis)), ctypes.byref(h))
    _native.raise_for_status(rc)
    return InflatoxPyDyLib(h, str(lib_path))


# ----------------------------------------------------------------------------------------------
# grid pyfunctions (reference src/anguelova.rs:176-550, 569-626)
# ----------------------------------------------------------------------------------------------
def _grid2(fn_name: str, lib, p, out, start_stop, progress, threads) -> None:
    p, ss = _f64_in(p, "parameter array"), _ss(start_stop)
    out = _out(out, np.float64, "out")
    if out.ndim != 2:
        raise TypeError("out must be a 2D float64 array")
    with _gate:
        rc = getattr(_native.lib(), fn_name)(
            lib._h, _dptr(p), p.size, _dptr(out), out.shape[0], out.shape[1], _dptr(ss),
            ss.shape[0], ss.shape[1], int(bool(progress)), int(threads),
        )  # fmt: skip
    _native.raise_for_status(rc)


def consistency_only(lib, p, out, start_stop, progress, threads) -> None:
    _grid2("inflx_consistency_only", lib, p, out, start_stop, progress, threads)


def consistency_rapidturn_only(lib, p, out, start_stop, progress, threads) -> None:
    _grid2("inflx_consistency_rapidturn_only", lib, p, out, start_stop, progress, threads)


def epsilon_v_only(lib, p, out, start_stop, progress, threads) -> None:
    _grid2("inflx_epsilon_v_only", lib, p, out, start_stop, progress, threads)


def complete_analysis(lib, p, out, start_stop, progress, threads) -> None:
    p, ss = _f64_in(p, "parameter array"), _ss(start_stop)
    out = _out(out, np.float64, "out")
    if out.ndim != 3:
        raise TypeError("out must be a 3D float64 array")
    with _gate:
        rc = _native.lib().inflx_complete_analysis(
            lib._h, _dptr(p), p.size, _dptr(out), out.shape[0], out.shape[1], out.shape[2],
            _dptr(ss), ss.shape[0], ss.shape[1], int(bool(progress)), int(threads),
        )  # fmt: skip
    _native.raise_for_status(rc)


def flag_quantum_dif_py(lib, p, x, start_stop, progress, accuracy) -> None:
    p, ss = _f64_in(p, "parameter array"), _ss(start_stop)
    x = _out(x, np.bool_, "x")
    if x.ndim != 2:
        raise TypeError("x must be a 2D bool array")
    with _gate:
        rc = _native.lib().inflx_flag_quantum_dif(
            lib._h, _dptr(p), p.size, x.ctypes.data_as(ctypes.c_void_p), x.shape[0], x.shape[1],
            _dptr(ss), ss.shape[0], ss.shape[1], int(bool(progress)), float(accuracy),
        )  # fmt: skip
    _native.raise_for_status(rc)


# ----------------------------------------------------------------------------------------------
# on-trajectory pyfunctions (reference src/anguelova.rs:633-977)
# ----------------------------------------------------------------------------------------------
def _traj_x(x) -> np.ndarray:
    x = _f64_in(x, "field-space array")
    if x.ndim != 2:
        raise TypeError("x must be a 2D float64 array")
    return x


def complete_analysis_on_trajectory(lib, p, x, out, progress, threads) -> None:
    p, x = _f64_in(p, "parameter array"), _traj_x(x)
    out = _out(out, np.float64, "out")
    if out.ndim != 2:
        raise TypeError("out must be a 2D float64 array")
    rc = _native.lib().inflx_complete_analysis_on_trajectory(
        lib._h, _dptr(p), p.size, _dptr(x), x.shape[0], x.shape[1], _dptr(out), out.shape[0],
        out.shape[1], int(bool(progress)), int(threads),
    )  # fmt: skip
    _native.raise_for_status(rc)


def _traj1(fn_name: str, lib, p, x, out, progress, threads) -> None:
    p, x = _f64_in(p, "parameter array"), _traj_x(x)
    out = _out(out, np.float64, "out")
    if out.ndim != 1:
        raise TypeError("out must be a 1D float64 array")
    rc = getattr(_native.lib(), fn_name)(
        lib._h, _dptr(p), p.size, _dptr(x), x.shape[0], x.shape[1], _dptr(out), out.shape[0],
        int(bool(progress)), int(threads),
    )  # fmt: skip
    _native.raise_for_status(rc)


def consistency_only_on_trajectory(lib, p, x, out, progress, threads) -> None:
    _traj1("inflx_consistency_only_on_trajectory", lib, p, x, out, progress, threads)


def consistency_rapidturn_only_on_trajectory(lib, p, x, out, progress, threads) -> None:
    _traj1("inflx_consistency_rapidturn_only_on_trajectory", lib, p, x, out, progress, threads)


def epsilon_v_only_on_trajectory(lib, p, x, out, progress, threads) -> None:
    _traj1("inflx_epsilon_v_only_on_trajectory", lib, p, x, out, progress, threads)


def solve_eom_rk4(*args, **kwargs):
    raise NotImplementedError(
        "the background ODE solver (reference src/background_solver.rs) is a sequential CPU "
        "integrator outside the grid-evaluation path this package accelerates"
    )


solve_eom_rkf = solve_eom_rk4


# ----------------------------------------------------------------------------------------------
# extensions: pinned outputs, fused parameter sweep, row shards, device-resident output
# ----------------------------------------------------------------------------------------------
class _PinnedBlock:
    """Owner of one cuMemHostAlloc block; freed when the last numpy view dies."""

    pinned = True

    total = 0  # bytes currently page-locked through this pool

    def __init__(self, nbytes: int, placement: tuple = ()):
        ptr = ctypes.c_void_p()
        devs = (ctypes.c_int * max(1, len(placement)))(*placement)
        _native.raise_for_status(
            _native.lib().inflx_host_alloc_on(nbytes, devs, len(placement), ctypes.byref(ptr))
        )
        self.ptr, self.nbytes, self.key = ptr, nbytes, (nbytes, tuple(placement))
        _PinnedBlock.total += nbytes

    def address(self) -> int:
        return self.ptr.value

    def __del__(self):
        p, self.ptr = getattr(self, "ptr", None), None
        if p:
            try:
                _native.lib().inflx_host_free(p)
                _PinnedBlock.total -= self.nbytes
            except Exception:
                pass


class _PageableBlock:
    """Plain (pageable) host block; reused so that its pages are faulted in only once.  Ordinary
    4 KiB pages on purpose: the copy-out threads fault them in in parallel at memcpy speed,
    whereas first-touching a MADV_HUGEPAGE mapping from those threads cost 2-6 s per 12 GiB on the
    round-1 boxes (tools/coldstart_probe.py)."""

    pinned = False

    def __init__(self, nbytes: int):
        self.buf = np.empty(nbytes + 64, dtype=np.uint8)
        self.nbytes, self.key = nbytes, nbytes

    def address(self) -> int:
        return (self.buf.ctypes.data + 63) & ~63


# re-entrant: a lease's __del__ (below) takes it too, and the cyclic GC may run that finaliser on
# the very thread that is inside one of the locked regions
_pool_lock = threading.RLock()
_pin_pool: dict[tuple, list[_PinnedBlock]] = {}  # (bytes, placement) -> free blocks
_page_pool: dict[int, list[_PageableBlock]] = {}
_pin_jobs: dict[tuple, threading.Thread] = {}
_POOL_DEPTH = 4
# devices whose row shards the next pinned outputs will receive (NUMA placement of the block's
# slices, inflx_host_alloc_on); () = the engine's default devices.  Follows the last
# open_inflx_dylib / InflatoxPyDyLib.set_devices call.
_placement: tuple = ()


def set_output_placement(devices) -> None:
    global _placement
    _placement = tuple(int(d) for d in devices)


class _Lease:
    """Returns a block to its pool (instead of freeing / unpinning it) when the array dies."""

    def __init__(self, block):
        self.block = block

    def __del__(self):
        try:
            pools = _pin_pool if self.block.pinned else _page_pool
            with _pool_lock:
                pool = pools.setdefault(self.block.key, [])
                if len(pool) < _POOL_DEPTH:
                    pool.append(self.block)
        except Exception:  # interpreter shutdown: module globals may already be gone
            pass


def _pin_budget() -> int:
    """Upper bound on page-locked pool memory: $INFLATOX_PINNED_MAX_GB, else a third of the RAM."""
    env = os.environ.get("INFLATOX_PINNED_MAX_GB")
    if env:
        return int(float(env) * (1 << 30))
    try:
        return os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") // 3
    except (ValueError, OSError):
        return 16 << 30


def _round_block(shape, dtype) -> tuple[tuple, int, int]:
    shape = tuple(int(s) for s in (shape if hasattr(shape, "__len__") else (shape,)))
    count = int(np.prod(shape))
    nbytes = max(1, count * np.dtype(dtype).itemsize)
    return shape, count, (nbytes + (1 << 21) - 1) & ~((1 << 21) - 1)


def _as_array(block, shape, count, dtype) -> np.ndarray:
    lease = _Lease(block)
    buf = (ctypes.c_char * block.nbytes).from_address(block.address())
    arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
    buf._inflx_lease = lease  # keep the lease alive as long as any view of the buffer exists
    return arr


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array in page-locked host memory, allocated NOW (pinning runs at ~2.5 GB/s, so this
    blocks for seconds on a multi-GB array) and pooled.  Outputs allocated here are written by DMA
    straight from the GPU(s)."""
    shape, count, nbytes = _round_block(shape, dtype)
    key = (nbytes, _placement)
    with _pool_lock:
        pool = _pin_pool.get(key)
        block = pool.pop() if pool else None
    if block is None:
        block = _PinnedBlock(nbytes, _placement)
    return _as_array(block, shape, count, dtype)


def _pool_debug(msg: str) -> None:
    if os.environ.get("INFLATOX_DEBUG_POOL"):
        print(f"[inflatox pool {time.perf_counter():.3f}] {msg}", file=sys.stderr, flush=True)


def host_output(shape, dtype=np.float64) -> np.ndarray:
    """Output array for the facade.  A pinned block from the pool when one is free (direct DMA,
    ~55 GB/s); otherwise a pooled pageable block for THIS call (staged copy-out, 30-47 GB/s) while
    a background thread page-locks a block of that size for the next ones, so that a cold call
    does not wait the seconds it takes to pin its own output first.  Page-locking holds the
    driver's lock, so the job waits until the engine call this array is for has returned
    ($INFLATOX_PIN_MODE: `deferred` (default) | `eager`: pin concurrently | `sync`: pin in the
    call, the pre-pool behaviour | `off`: pageable only)."""
    mode = os.environ.get("INFLATOX_PIN_MODE", "deferred")
    if mode == "sync":
        return pinned_empty(shape, dtype)
    shape, count, nbytes = _round_block(shape, dtype)
    key = (nbytes, _placement)
    with _pool_lock:
        pool = _pin_pool.get(key)
        block = pool.pop() if pool else None
        if block is None:
            job = _pin_jobs.get(key)
            if (
                mode != "off"
                and (job is None or not job.is_alive())
                and _PinnedBlock.total + nbytes <= _pin_budget()
            ):
                seen = _gate.completed

                def pin(nb=nbytes, key=key):
                    if mode != "eager":
                        _gate.wait_idle_after(seen, timeout=10.0)
                    if _gate.closing:
                        return
                    _pool_debug(f"pinning {nb >> 20} MiB")
                    try:
                        blk = _PinnedBlock(nb, key[1])
                    except Exception:
                        return  # no GPU / out of lockable memory: stay on the staged path
                    _pool_debug(f"pinned {nb >> 20} MiB")
                    with _pool_lock:
                        _pin_pool.setdefault(key, []).append(blk)

                # not a daemon: interpreter shutdown waits for a page-lock in flight instead of
                # tearing the driver down under it
                job = threading.Thread(target=pin, daemon=False)
                _pin_jobs[key] = job
                job.start()
            pages = _page_pool.get(nbytes)
            block = pages.pop() if pages else None
    if block is None:
        block = _PageableBlock(nbytes)
    _pool_debug(f"output {nbytes >> 20} MiB: {'pinned' if block.pinned else 'pageable'}")
    return _as_array(block, shape, count, dtype)


def grid_eval(lib, op: str, p, out, n0: int, n1: int, start_stop, rows=None, accuracy: float = 0.0,
              device: int = -1, out_device_ptr: int | None = None, stream: int | None = None):  # fmt: skip
    """Extended grid interface (inflx_grid_eval): row shard `rows=(begin,end)` with global
    coordinates, parameter sweep (`p` of shape (S,P)), optional device-resident output.  Returns
    the engine's report as a dict."""
    p = _f64_in(p, "parameter array")
    p2 = p.reshape(1, -1) if p.ndim == 1 else p
    if p2.shape[1] != lib.n_parameters:
        raise Exception(
            f"Expected array with shape [2], received array with shape [{p2.shape[1]}]. "
            f'Context: model "{lib.name}" has {lib.n_parameters} paramters'
        )
    ss = np.ascontiguousarray(start_stop, dtype=np.float64).reshape(4)
    r0, r1 = rows if rows is not None else (0, n0)
    rq = _native.GridRequest()
    rq.op = _native.OPS[op]
    rq.params = _dptr(p2)
    rq.n_vectors = p2.shape[0]
    rq.n0, rq.n1 = n0, n1
    rq.start_stop = (ctypes.c_double * 4)(*ss)
    rq.row_begin, rq.row_end = r0, r1
    rq.aux = accuracy
    if out_device_ptr is not None:
        rq.out = ctypes.c_void_p(out_device_ptr)
        rq.out_is_device = 1
        rq.stream = ctypes.c_void_p(stream or 0)
    else:
        per = {"complete_analysis": 6, "hesse": 4}.get(op, 1)
        need = p2.shape[0] * (r1 - r0) * n1 * per
        want = np.bool_ if op == "flag_quantum_dif" else np.float64
        if (not isinstance(out, np.ndarray) or not out.flags.c_contiguous or out.size != need
                or out.dtype != want or not out.flags.writeable):  # fmt: skip
            raise TypeError(
                f"out must be a writeable C-contiguous {np.dtype(want).name} array of {need} elements"
            )
        rq.out = out.ctypes.data_as(ctypes.c_void_p)
        rq.out_is_device = 0
    rq.device = device
    rep = _native.GridReport()
    with _gate:
        rc = _native.lib().inflx_grid_eval(lib._h, ctypes.byref(rq), ctypes.byref(rep))
    _native.raise_for_status(rc)
    return {
        "kernel_ms": rep.kernel_ms, "grid_ms": rep.grid_ms, "total_ms": rep.total_ms, "launches": int(rep.launches),
        "d2h_bytes": int(rep.d2h_bytes), "h2d_bytes": int(rep.h2d_bytes),
        "n_devices": rep.n_devices,
    }  # fmt: skip


def sweep(lib, op: str, params, out, start_stop, accuracy: float = 0.0):
    """Fused parameter sweep: `params` (S,P), `out` (S,N0,N1[,6]); the sweep axis is part of the
    launch grid (BASELINE config C5)."""
    n0, n1 = out.shape[1], out.shape[2]
    return grid_eval(lib, op, params, out, n0, n1, start_stop, accuracy=accuracy)
