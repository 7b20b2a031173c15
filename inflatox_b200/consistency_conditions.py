"""Numerical façade: `InflationCondition` and `GeneralisedAL`.

Same classes, method names, argument order, defaults and return shapes as the reference's
python/inflatox/consistency_conditions.py:31-715; every method allocates the numpy output, builds
`start_stop` and hands the work to `libinflx_rs` (here: the CUDA engine) exactly like the
reference does with its Rust extension.  The only differences a caller can observe:

  * outputs larger than a few MiB come from a host pool: page-locked blocks (the GPUs DMA straight
    into the array the caller receives) once a background thread has pinned one of that size,
    pooled pageable blocks until then (`INFLATOX_PINNED=0` restores plain np.zeros);
  * `threads` / `progress` are accepted and ignored (there is no CPU thread pool);
  * `GeneralisedAL.sweep_complete_analysis` is an addition (fused parameter sweep, BASELINE C5).
"""
from __future__ import annotations

import os

import numpy as np

from .compiler import CompilationArtifact
from .libinflx_rs import *  # noqa: F401,F403  (the reference does the same, :25)
from . import libinflx_rs as _rs

__all__ = ["InflationCondition", "GeneralisedAL"]

_PIN_THRESHOLD = 4 << 20


def _new_output(shape, dtype=float) -> np.ndarray:
    """np.zeros for the caller-visible result; pinned (pooled) when large enough to matter."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    if nbytes >= _PIN_THRESHOLD and os.environ.get("INFLATOX_PINNED", "1") != "0":
        return _rs.host_output(shape, dtype)
    return np.zeros(shape, dtype=dtype)


def _start_stop(x0_start, x0_stop, x1_start, x1_stop) -> np.ndarray:
    return np.array([[float(x0_start), float(x0_stop)], [float(x1_start), float(x1_stop)]])


class InflationCondition:
    """Base class of all inflation conditions: evaluates the potential and the projected Hesse
    matrix of a compiled model (reference consistency_conditions.py:31-196)."""

    def __init__(self, compiled_artifact: CompilationArtifact, validate_basis: bool = True):
        self.artifact = compiled_artifact
        self.dylib = open_inflx_dylib(compiled_artifact.shared_object_path, validate_basis)

    def calc_V(self, x: np.ndarray, args: np.ndarray) -> float:
        """Scalar potential at field-space point `x` with model parameters `args`."""
        return self.dylib.potential(x, args)

    def calc_V_array(self, args, start, stop, N=None) -> np.ndarray:
        """Potential on a regular grid: axis k runs over [start[k], stop[k]) with N[k] samples
        (8000 per axis when `N` is None, as in the reference)."""
        n_fields = self.artifact.n_fields
        start_stop = np.array([[float(a), float(b)] for (a, b) in zip(start, stop)])
        N = tuple(N) if N is not None else tuple(8000 for _ in range(n_fields))
        x = _new_output(N)
        self.dylib.potential_array(x, np.asarray(args, dtype=float), start_stop)
        return x

    def calc_H(self, x: np.ndarray, args: np.ndarray) -> np.ndarray:
        """Projected Hesse matrix [[v00, v01], [v10, v11]] at `x`."""
        return self.dylib.hesse(x, args)

    def calc_H_array(
        self,
        args,
        x0_start,
        x0_stop,
        x1_start=None,
        x1_stop=None,
        N=None,
    ) -> np.ndarray:
        """Projected Hesse matrix on a regular grid; result has shape (2, 2, *N).

        Signature of the reference (consistency_conditions.py:119-127): `(args, x0_start, x0_stop,
        x1_start, x1_stop, N=None)`.  Upstream the call cannot succeed - the wrapper passes
        `np.array(n_fields)` instead of `N` and the Rust side asserts `p.len() == n_fields`
        (consistency_conditions.py:156, src/hesse_bindings.rs:163) - so only those two bugs are
        fixed here.  As an extra, the list style of `calc_V_array` is accepted too:
        `calc_H_array(args, start, stop, N)` with `start` / `stop` sequences."""
        n_fields = self.artifact.n_fields
        if np.ndim(x0_start) > 0:  # (args, start, stop[, N]) - `x1_start` then carries N
            if x1_stop is not None:
                raise TypeError("calc_H_array(args, start, stop, N): too many positional arguments")
            start, stop = x0_start, x0_stop
            N = x1_start if N is None else N
            start_stop = np.array([[float(a), float(b)] for (a, b) in zip(start, stop)])
        else:
            if x1_start is None or x1_stop is None:
                raise TypeError("calc_H_array() missing required arguments: 'x1_start', 'x1_stop'")
            start_stop = np.array(
                [[float(x0_start), float(x0_stop)], [float(x1_start), float(x1_stop)]]
            )
        N = N if N is not None else [8000 for _ in range(n_fields)]
        return self.dylib.hesse_array(
            np.array(N, dtype=np.int64), np.asarray(args, dtype=float), start_stop
        )

    def validate_basis_on_domain(self, args, start, stop, N=100, accuracy: float = 1e-3) -> None:
        """Checks orthonormality of the basis {v, w} along the axes of the given domain."""
        n_fields = self.artifact.n_fields
        start_stop = np.array([[float(a), float(b)] for (a, b) in zip(start, stop)])
        N = np.array(N if isinstance(N, (list, tuple, np.ndarray)) else [N] * n_fields, dtype=np.uint32)
        self.dylib.validate_basis_on_domain(N, np.asarray(args, dtype=float), start_stop, accuracy)


class GeneralisedAL(InflationCondition):
    """Generalised Anguelova-Lazaroiu consistency condition and the quantities derived from it
    (reference consistency_conditions.py:199-715)."""

    def __init__(self, compiled_artifact: CompilationArtifact):
        super().__init__(compiled_artifact)

    # -- grids --------------------------------------------------------------------------------
    def complete_analysis(
        self,
        args: np.ndarray,
        x0_start: float,
        x0_stop: float,
        x1_start: float,
        x1_stop: float,
        N_x0: int = 1_000,
        N_x1: int = 1_000,
        progress: bool = True,
        threads: None | int = None,
    ):
        """Six (N_x0, N_x1) arrays: consistency |lhs-rhs|/(|lhs|+|rhs|), ε_V, ε_H, η_H, δ, ω."""
        out = _new_output((N_x0, N_x1, 6))
        start_stop = _start_stop(x0_start, x0_stop, x1_start, x1_stop)
        threads = threads if threads is not None else 0
        complete_analysis(self.dylib, args, out, start_stop, progress, threads)
        return tuple(out[:, :, k] for k in range(6))

    def consistency(
        self,
        args: np.ndarray,
        x0_start: float,
        x0_stop: float,
        x1_start: float,
        x1_stop: float,
        N_x0: int = 1_000,
        N_x1: int = 1_000,
        progress: bool = True,
        threads: None | int = None,
    ) -> np.ndarray:
        """||lhs|-|rhs||/(|lhs|+|rhs|) of the AL consistency condition on the grid."""
        out = _new_output((N_x0, N_x1))
        start_stop = _start_stop(x0_start, x0_stop, x1_start, x1_stop)
        threads = threads if threads is not None else 0
        consistency_only(self.dylib, args, out, start_stop, progress, threads)
        return out

    def epsilon_v(
        self,
        args: np.ndarray,
        x0_start: float,
        x0_stop: float,
        x1_start: float,
        x1_stop: float,
        N_x0: int = 1_000,
        N_x1: int = 1_000,
        progress: bool = True,
        threads: None | int = None,
    ) -> np.ndarray:
        """First potential slow-roll parameter ε_V on the grid."""
        out = _new_output((N_x0, N_x1))
        start_stop = _start_stop(x0_start, x0_stop, x1_start, x1_stop)
        threads = threads if threads is not None else 0
        epsilon_v_only(self.dylib, args, out, start_stop, progress, threads)
        return out

    def consistency_rapidturn(
        self,
        args: np.ndarray,
        x0_start: float,
        x0_stop: float,
        x1_start: float,
        x1_stop: float,
        N_x0: int = 1_000,
        N_x1: int = 1_000,
        progress: bool = True,
        threads: None | int = None,
    ) -> np.ndarray:
        """Consistency condition in the rapid-turn approximation on the grid."""
        out = _new_output((N_x0, N_x1))
        start_stop = _start_stop(x0_start, x0_stop, x1_start, x1_stop)
        threads = threads if threads is not None else 0
        consistency_rapidturn_only(self.dylib, args, out, start_stop, progress, threads)
        return out

    def flag_quantum_dif(
        self,
        args: np.ndarray,
        x0_start: float,
        x0_stop: float,
        x1_start: float,
        x1_stop: float,
        N_x0: int = 10_000,
        N_x1: int = 10_000,
        progress=True,
        accuracy=1e-3,
    ) -> np.ndarray:
        """Boolean grid: True where every component of the basis vector v is <= `accuracy`."""
        x = np.zeros((N_x0, N_x1), dtype=bool)
        start_stop = _start_stop(x0_start, x0_stop, x1_start, x1_stop)
        flag_quantum_dif_py(self.dylib, args, x, start_stop, progress, accuracy)
        return x

    # -- on a trajectory ------------------------------------------------------------------------
    def complete_analysis_ot(
        self, args: np.ndarray, x: np.ndarray, progress: bool = True, threads: None | int = None
    ):
        """`complete_analysis` at the explicit points x[n, 2]; six (n, 1) arrays."""
        out = np.zeros((x.shape[0], 6), dtype=float)
        threads = threads if threads is not None else 1
        complete_analysis_on_trajectory(self.dylib, args, x, out, progress, threads)
        return np.split(out, 6, 1)

    def consistency_ot(
        self, args: np.ndarray, x: np.ndarray, progress: bool = True, threads: None | int = None
    ) -> np.ndarray:
        out = np.zeros((x.shape[0]), dtype=float)
        threads = threads if threads is not None else 1
        consistency_only_on_trajectory(self.dylib, args, x, out, progress, threads)
        return out

    def consistency_rapidturn_ot(
        self, args: np.ndarray, x: np.ndarray, progress: bool = True, threads: None | int = None
    ) -> np.ndarray:
        out = np.zeros((x.shape[0]), dtype=float)
        threads = threads if threads is not None else 1
        consistency_rapidturn_only_on_trajectory(self.dylib, args, x, out, progress, threads)
        return out

    def epsilon_v_ot(
        self, args: np.ndarray, x: np.ndarray, progress: bool = True, threads: None | int = None
    ) -> np.ndarray:
        out = np.zeros((x.shape[0]), dtype=float)
        threads = threads if threads is not None else 1
        epsilon_v_only_on_trajectory(self.dylib, args, x, out, progress, threads)
        return out

    # -- addition: fused parameter sweep (BASELINE config C5) -----------------------------------
    def sweep_complete_analysis(
        self, args_batch: np.ndarray, x0_start, x0_stop, x1_start, x1_stop, N_x0=1_000, N_x1=1_000
    ) -> np.ndarray:
        """`complete_analysis` for S parameter vectors `args_batch[S, P]` in one fused launch
        grid; returns an (S, N_x0, N_x1, 6) array."""
        args_batch = np.ascontiguousarray(args_batch, dtype=float)
        out = _new_output((args_batch.shape[0], N_x0, N_x1, 6))
        _rs.sweep(
            self.dylib, "complete_analysis", args_batch, out,
            _start_stop(x0_start, x0_stop, x1_start, x1_stop),
        )  # fmt: skip
        return out
