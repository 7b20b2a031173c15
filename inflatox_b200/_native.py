"""Loader (and in-tree builder) of the C-ABI engine `libinflx_b200.so` (include/inflx_b200.h).

The shared object is built IN TREE next to this file (git-ignored, shipped with the working tree)
from csrc/inflx_engine.cpp with the host C++ compiler; it binds the CUDA driver and NVRTC at run
time (csrc/inflx_cuda_dl.h), so it loads on a machine without a GPU while every compute entry
point fails loudly with INFLX_ERR_CUDA there.  There is no Python/CPU fall-back for any of them.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libinflx_b200.so")
HEADER_PATH = os.path.join(_ROOT, "include", "inflx_b200.h")

_lock = threading.Lock()
_lib: ctypes.CDLL | None = None

# inflx_status values (include/inflx_b200.h)
OK, ERR_IO, ERR_MISSING_SYMBOL, ERR_VERSION, ERR_THREADS, ERR_SHAPE = 0, 1, 2, 3, 4, 5
ERR_FIELD_DIM, ERR_BASIS_NORM, ERR_BASIS_OTH, ERR_CUDA, ERR_NVRTC = 6, 7, 8, 9, 10

OPS = {
    "complete_analysis": 0, "consistency_only": 1, "consistency_rapidturn_only": 2,
    "epsilon_v_only": 3, "flag_quantum_dif": 4, "potential": 5, "hesse": 6, "basis": 7,
}  # fmt: skip


def _sources() -> list[str]:
    return [
        os.path.join(_CSRC, "inflx_engine.cpp"),
        os.path.join(_CSRC, "inflx_cuda_dl.h"),
        HEADER_PATH,
    ]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False) -> str:
    """Compile the engine in tree.  Needs g++ and the CUDA toolkit headers (cuda.h, nvrtc.h)."""
    if not force and not needs_build():
        return LIB_PATH
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    tmp = LIB_PATH + f".{os.getpid()}.tmp"
    cmd = [
        os.environ.get("CXX", "g++"), "-O2", "-fPIC", "-shared", "-std=c++17", "-Wall",
        f"-I{cuda_home}/include", "-o", tmp, os.path.join(_CSRC, "inflx_engine.cpp"),
        "-ldl", "-lpthread",
    ]  # fmt: skip
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise ImportError(f"building libinflx_b200.so failed:\n{' '.join(cmd)}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


class GridRequest(ctypes.Structure):
    _fields_ = [
        ("op", ctypes.c_int),
        ("params", ctypes.POINTER(ctypes.c_double)),
        ("n_vectors", ctypes.c_uint64),
        ("n0", ctypes.c_uint64),
        ("n1", ctypes.c_uint64),
        ("start_stop", ctypes.c_double * 4),
        ("row_begin", ctypes.c_uint64),
        ("row_end", ctypes.c_uint64),
        ("aux", ctypes.c_double),
        ("out", ctypes.c_void_p),
        ("out_is_device", ctypes.c_int),
        ("device", ctypes.c_int),
        ("stream", ctypes.c_void_p),
    ]


class GridReport(ctypes.Structure):
    _fields_ = [
        ("kernel_ms", ctypes.c_double),
        ("grid_ms", ctypes.c_double),
        ("total_ms", ctypes.c_double),
        ("launches", ctypes.c_uint64),
        ("d2h_bytes", ctypes.c_uint64),
        ("h2d_bytes", ctypes.c_uint64),
        ("n_devices", ctypes.c_int),
    ]


def _declare(lib: ctypes.CDLL) -> None:
    c = ctypes
    dp, sz, vp, ci = c.POINTER(c.c_double), c.c_size_t, c.c_void_p, c.c_int
    lib.inflx_last_error.restype = c.c_char_p
    lib.inflx_build_info.restype = c.c_char_p
    lib.inflx_model_name.restype = c.c_char_p
    lib.inflx_model_name.argtypes = [vp]
    lib.inflx_n_fields.restype = c.c_uint32
    lib.inflx_n_fields.argtypes = [vp]
    lib.inflx_n_parameters.restype = c.c_uint32
    lib.inflx_n_parameters.argtypes = [vp]
    lib.inflx_abi_version.restype = None
    lib.inflx_abi_version.argtypes = [vp, c.POINTER(c.c_uint16)]
    lib.inflx_kernel_launches.restype = c.c_uint64
    lib.inflx_device_count.restype = ci
    lib.inflx_free.restype = None
    lib.inflx_free.argtypes = [vp]
    lib.inflx_close.restype = None
    lib.inflx_close.argtypes = [vp]
    lib.inflx_open.argtypes = [c.c_char_p, ci, c.POINTER(vp)]
    lib.inflx_nvrtc_compile.argtypes = [
        c.c_char_p, c.c_char_p, c.POINTER(c.c_char_p), ci, c.POINTER(vp), c.POINTER(sz),
        c.POINTER(vp),
    ]  # fmt: skip
    lib.inflx_nvrtc_version.argtypes = [c.POINTER(ci), c.POINTER(ci)]
    lib.inflx_shard_of.argtypes = [c.c_uint64, c.c_uint64, c.c_uint64, c.c_uint64, c.POINTER(c.c_uint64)]
    lib.inflx_set_devices.argtypes = [vp, c.POINTER(ci), ci]
    lib.inflx_get_devices.argtypes = [vp, c.POINTER(ci), ci]
    lib.inflx_complete_analysis.argtypes = [vp, dp, sz, dp, sz, sz, sz, dp, sz, sz, ci, sz]
    for n in ("inflx_consistency_only", "inflx_consistency_rapidturn_only", "inflx_epsilon_v_only"):
        getattr(lib, n).argtypes = [vp, dp, sz, dp, sz, sz, dp, sz, sz, ci, sz]
    lib.inflx_flag_quantum_dif.argtypes = [vp, dp, sz, vp, sz, sz, dp, sz, sz, ci, c.c_double]
    lib.inflx_complete_analysis_on_trajectory.argtypes = [vp, dp, sz, dp, sz, sz, dp, sz, sz, ci, sz]
    for n in (
        "inflx_consistency_only_on_trajectory",
        "inflx_consistency_rapidturn_only_on_trajectory",
        "inflx_epsilon_v_only_on_trajectory",
    ):
        getattr(lib, n).argtypes = [vp, dp, sz, dp, sz, sz, dp, sz, ci, sz]
    lib.inflx_potential.argtypes = [vp, dp, sz, dp, sz, dp]
    lib.inflx_hesse.argtypes = [vp, dp, sz, dp, sz, dp]
    lib.inflx_potential_array.argtypes = [vp, dp, sz, sz, dp, sz, dp, sz, sz]
    lib.inflx_hesse_array.argtypes = [vp, dp, sz, sz, dp, sz, dp, sz, sz]
    lib.inflx_validate_basis_on_domain.argtypes = [
        vp, c.POINTER(c.c_uint32), sz, dp, sz, dp, sz, sz, c.c_double,
    ]  # fmt: skip
    lib.inflx_grid_eval.argtypes = [vp, c.POINTER(GridRequest), c.POINTER(GridReport)]
    lib.inflx_points_eval.argtypes = [vp, ci, dp, dp, c.c_uint64, c.c_double, dp]
    lib.inflx_measure_fp64_peak.argtypes = [ci, ci, dp, dp]
    lib.inflx_host_alloc.argtypes = [sz, c.POINTER(vp)]
    lib.inflx_host_alloc_on.argtypes = [sz, c.POINTER(ci), ci, c.POINTER(vp)]
    lib.inflx_device_numa_node.argtypes = [ci]
    lib.inflx_device_numa_node.restype = ci
    lib.inflx_host_free.argtypes = [vp]


def lib() -> ctypes.CDLL:
    """The loaded engine; builds it first when the in-tree .so is missing or stale."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if needs_build():
                    build()
                handle = ctypes.CDLL(LIB_PATH)
                _declare(handle)
                _lib = handle
    return _lib


def last_error() -> str:
    return lib().inflx_last_error().decode("utf-8", "replace")


def raise_for_status(code: int) -> None:
    """inflx_status -> the Python exception class the reference raises for the corresponding
    LibInflxRsErr (reference src/err.rs:63-74)."""
    if code == OK:
        return
    msg = last_error()
    if code == ERR_IO:
        raise IOError(msg)
    if code in (ERR_MISSING_SYMBOL, ERR_VERSION, ERR_THREADS, ERR_CUDA):
        raise SystemError(msg)
    raise Exception(msg)
