"""Shared test inputs: the five models of BASELINE.json with the parameter vectors and extents of
the reference's own tests (reference tests/test_doc.py:53, README.md:59-66,
tests/test_angular.py:63-68, tests/test_egno.py:80-90, tests/test_d5.py:144-158)."""
import functools
import math
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

MODELS = ("doc", "hyper", "angular", "egno", "d5")

PARAMS = {
    "doc": [1.0],
    "hyper": [1.0, 1.0, 1.0],
    "angular": [1 / 600, 6e-5, 2e-5],
    "egno": [1e-3, 0.5, 1000.0, 1.0],
    "d5": [-1.17e-8, 1e-3, 5.0, 1.0, 50 * 501.961, 501.961, 5e-4, 1e-3, 0.01, 1000.0],
}
EXTENT = {  # x0_start, x0_stop, x1_start, x1_stop
    "doc": (0.0, 2.5, 0.0, math.pi),
    "hyper": (-1.0, 1.0, -1.0, 1.0),
    "angular": (-1.05, 1.05, -1.05, 1.05),
    "egno": (0.46, 0.50, 0.0, math.pi),
    "d5": (0.0, 36.0, 0.0, 4 * math.pi),
}


def params(model):
    return np.array(PARAMS[model], dtype=np.float64)


@functools.lru_cache(maxsize=None)
def golden_cse(model: str) -> bool:
    """Whether the reference's test compiles this model with cse=True (tests/golden/c/*.json)."""
    import json

    with open(os.path.join(GOLDEN, "c", f"{model}.json")) as fh:
        return bool(json.load(fh)["cse"])


def load_model(model: str):
    """The symbolic model of one of the reference's tests, built by the reference's own
    `InflationModelBuilder` and pickled by tests/golden/make_golden.py (a dict of sympy
    expressions).  TEST FIXTURE LOADER: unpickling executes code, so this lives under tests/ and
    only ever reads the files committed under tests/golden/models/."""
    import gzip
    import pickle

    import inflatox_b200 as ix

    with gzip.open(os.path.join(GOLDEN, "models", f"{model}.pkl.gz"), "rb") as fh:
        d = pickle.load(fh)
    return ix.InflationModel(**{k: d[k] for k in ix.InflationModel.FIELDS})


@functools.lru_cache(maxsize=None)
def artifact(model: str, fmad: bool = False, libm: str | None = None):
    """Compile the pickled reference-built model with inflatox_b200.Compiler (cached per process;
    cubins additionally cached on disk by content hash).  `libm`: flavour of the hoisted libm
    calls (cudagen.LIBM_FLAVOURS), default = the Compiler's."""
    import inflatox_b200 as ix

    m = load_model(model)
    flags = None
    if fmad:
        flags = [f.replace("--fmad=false", "--fmad=true") for f in ix.Compiler.default_nvrtc_flags]
    c = ix.Compiler(m, silent=True, cse=golden_cse(model), cleanup=True, compiler_flags=flags)
    if libm is not None:
        c.libm = libm
    return c.compile()


def load_checked(open_fn, attempts: int = 6):
    """Run `open_fn()` - something that opens an artefact WITH the load-time basis validation.
    The reference validates the basis at 100 random points for RANDOM parameters in [-10, 10]
    (src/lib.rs:142-203, unseeded), and for the angular model that check fails in ~8 % of the loads
    on the reference's own CPU path (negative alpha: tests/test_oracle.py pins the rate with the
    oracle).  The product reproduces that behaviour; tests that only need a loaded model retry."""
    last = None
    for _ in range(attempts):
        try:
            return open_fn()
        except Exception as e:  # BasisNorm / BasisOth map to plain Exception (src/err.rs:63-74)
            if "Expected basis vector" not in str(e):
                raise
            last = e
    raise last


def trajectory(model: str) -> np.ndarray:
    """(n, 2) trajectory fixtures the reference's tests evaluate on (copied data files)."""
    d = os.path.join(GOLDEN, "trajectories")
    if model == "angular":
        return np.ascontiguousarray(
            np.stack([np.load(f"{d}/angular_phix.npy"), np.load(f"{d}/angular_phiy.npy")], axis=1)
        )
    if model == "egno":
        return np.ascontiguousarray(
            np.stack([np.load(f"{d}/egno_r.npy"), np.load(f"{d}/egno_theta.npy")], axis=1)
        )
    if model == "d5":
        return np.ascontiguousarray(np.loadtxt(f"{d}/d5_trajectory.dat")[:, :2])
    raise KeyError(model)


def rel_err(a, b):
    """elementwise |a-b| / max(|b|, tiny) on the jointly finite entries; (err array, mask stats)"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    fin = np.isfinite(a) & np.isfinite(b)
    err = np.zeros_like(a)
    denom = np.maximum(np.abs(b[fin]), np.finfo(np.float64).tiny)
    err[fin] = np.abs(a[fin] - b[fin]) / denom
    nan_mismatch = int((np.isnan(a) != np.isnan(b)).sum())
    inf_mismatch = int(((np.isinf(a) != np.isinf(b)) | (np.isinf(a) & (np.sign(a) != np.sign(b)))).sum())
    return err, fin, nan_mismatch, inf_mismatch
