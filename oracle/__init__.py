"""TEST INFRASTRUCTURE - NOT PRODUCT CODE.

CPU oracle for the inflatox grid-evaluation hot path: builds and drives oracle/inflx_oracle.c (the
restatement of the reference's Rust ops/grid loops) over the reference's own generated C model
artefacts (tests/golden/c/*.c.gz).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; nothing under inflatox_b200/ does.

Build products go to oracle/_build/<key>/ (git-ignored); <key> hashes the sources, the flags and
the host CPU flags because the reference's flag set contains -march=native (reference
python/inflatox/compiler.py:299-310) and the build may happen on a different host than the one
that wrote the snapshot.  `oracle/_ref/` is unused: the reference's Rust crate cannot be compiled
in this image (no rustc/cargo/maturin), see DESIGN.md.
"""
from __future__ import annotations

import ctypes
import gzip
import hashlib
import json
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN_C = os.path.join(ROOT, "tests", "golden", "c")

# the reference's flag set, verbatim (compiler.py:299-310) ...
REFERENCE_FLAGS = [
    "-O3", "-Wall", "-Werror", "-fpic", "-lm", "-march=native", "-shared", "-std=c17",
    "-fno-math-errno", "-fno-signed-zeros",
]  # fmt: skip
# ... plus the frozen contraction mode of the oracle (gcc's ISO-mode default anyway)
ORACLE_EXTRA = ["-ffp-contract=off"]
MODELS = ("doc", "hyper", "angular", "egno", "d5")

_DBL = ctypes.POINTER(ctypes.c_double)


def _cpu_key() -> str:
    try:
        with open("/proc/cpuinfo") as fh:
            for ln in fh:
                if ln.startswith("flags"):
                    return hashlib.sha1(ln.encode()).hexdigest()[:8]
    except OSError:
        pass
    return "generic"


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"oracle build failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")


def golden_c_text(model: str) -> str:
    with gzip.open(os.path.join(GOLDEN_C, f"{model}.c.gz"), "rt") as fh:
        return fh.read()


def golden_meta(model: str) -> dict:
    with open(os.path.join(GOLDEN_C, f"{model}.json")) as fh:
        return json.load(fh)


def quad_transliteration(c_text: str) -> str:
    """The same translation unit with every fp64 object and libm call replaced by its
    __float128 / libquadmath counterpart and every floating literal made a Q literal, so the
    expression is evaluated to ~34 digits.  Decimal literals (incl. the preamble's short M_PI)
    are taken at face value: the truth of *the text the reference generated*."""
    body = c_text
    body = re.sub(r"#include <math.h>", "#include <quadmath.h>", body)
    body = re.sub(r"\bdouble\b", "__float128", body)
    fns = (
        "pow sqrt cbrt exp exp2 expm1 log log2 log10 log1p sin cos tan asin acos atan atan2 sinh "
        "cosh tanh asinh acosh atanh fabs hypot erf erfc tgamma lgamma floor ceil fmin fmax fmod"
    ).split()
    body = re.sub(r"\b(" + "|".join(fns) + r")\(", lambda m: m.group(1) + "q(", body)
    # floating literals -> Q suffix (leave integers, array indices and identifiers alone)
    body = re.sub(
        r"(?<![A-Za-z_0-9\.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![A-Za-z_0-9\.])",
        lambda m: m.group(1) + "Q",
        body,
    )
    return body


class _Build:
    def __init__(self):
        srcs = [open(os.path.join(HERE, "inflx_oracle.c"), "rb").read()]
        for m in MODELS:
            p = os.path.join(GOLDEN_C, f"{m}.c.gz")
            if os.path.exists(p):
                srcs.append(open(p, "rb").read())
        h = hashlib.sha1(b"".join(srcs) + " ".join(REFERENCE_FLAGS + ORACLE_EXTRA).encode())
        self.dir = os.path.join(HERE, "_build", f"{_cpu_key()}-{h.hexdigest()[:10]}")
        os.makedirs(self.dir, exist_ok=True)

    def driver(self, quad: bool = False) -> str:
        out = os.path.join(self.dir, "liboracle_quad.so" if quad else "liboracle.so")
        if not os.path.exists(out):
            tmp = out + f".{os.getpid()}.tmp"
            cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=gnu17", "-fopenmp", "-ffp-contract=off",
                   "-fno-fast-math", "-Wall", "-o", tmp, os.path.join(HERE, "inflx_oracle.c")]  # fmt: skip
            if quad:
                cmd += ["-DINFLX_QUAD", "-lquadmath"]
            cmd += ["-lm", "-ldl"]
            _run(cmd)
            os.replace(tmp, out)
        return out

    def model(self, name: str, quad: bool = False, contract: str = "off", libm: str = "glibc") -> str:
        """Compile the reference-generated C of `name` with the reference's flags.

        `libm="cr"` is an ATTRIBUTION variant, not the parity oracle: the same text with libm's
        log / exp / pow / sin / cos replaced by the correctly rounded versions of
        inflatox_b200/csrc/inflx_crmath.cuh (host build).  Comparing the CUDA path with both
        shows which part of a residue is glibc's own misrounding (~1e-3 of the pow calls)."""
        tag = "quad" if quad else f"fpc_{contract}"
        if libm == "cr":
            with open(os.path.join(ROOT, "inflatox_b200", "csrc", "inflx_crmath.cuh"), "rb") as fh:
                tag += "_cr2" + hashlib.sha1(fh.read()).hexdigest()[:8]
        out = os.path.join(self.dir, f"{name}.{tag}.so")
        if not os.path.exists(out):
            text = golden_c_text(name)
            if quad:
                text = quad_transliteration(text)
            elif libm == "cr":
                hdr = os.path.join(ROOT, "inflatox_b200", "csrc", "inflx_crmath.cuh")
                # a negative base with an integer exponent (fields range over negative values and the
                # text is full of pow(x, 3)): correctly rounded through |x|, sign restored
                wrap = (
                    "static double inflx_oracle_cr_pow(double x, double y) {\n"
                    "  if (x < 0 && y == rint(y) && fabs(y) < 9e15) {\n"
                    "    const double r = inflx_cr_pow(-x, y);\n"
                    "    return fmod(fabs(y), 2.0) == 1.0 ? -r : r;\n"
                    "  }\n"
                    "  return inflx_cr_pow(x, y);\n"
                    "}\n#define pow inflx_oracle_cr_pow\n"
                )
                defs = "".join(f"#define {f} inflx_cr_{f}\n" for f in ("log", "exp", "sin", "cos"))
                text = text.replace(
                    "#include <math.h>", f'#include <math.h>\n#include "{hdr}"\n{wrap}{defs}', 1
                )
            src = os.path.join(self.dir, f"{name}.{tag}.c")
            with open(src, "w") as fh:
                fh.write(text)
            tmp = out + f".{os.getpid()}.tmp"
            if quad:
                cmd = ["gcc", "-O1", "-fpic", "-shared", "-std=gnu17", "-o", tmp, src,
                       "-lquadmath", "-lm"]  # fmt: skip
            else:
                flags = REFERENCE_FLAGS
                if libm == "cr":  # the header's unused static functions would trip -Werror
                    flags = [f for f in flags if f not in ("-Wall", "-Werror")]
                cmd = ["gcc", "-o", tmp, src, *flags, f"-ffp-contract={contract}"]
            _run(cmd)
            os.replace(tmp, out)
        return out


_build: _Build | None = None


def build_all(models=MODELS, quad: bool = False) -> str:
    """Compile the driver and the model artefacts (called by __graft_entry__.build())."""
    global _build
    _build = _build or _Build()
    _build.driver(False)
    for m in models:
        _build.model(m)
    if quad:
        _build.driver(True)
        for m in models:
            _build.model(m, quad=True)
    return _build.dir


class Oracle:
    """ctypes front-end of one model artefact opened by the restated loader."""

    def __init__(self, model: str, quad: bool = False, contract: str = "off", libm: str = "glibc"):
        global _build
        _build = _build or _Build()
        self.quad = quad
        self.sfx = "_quad" if quad else ""
        self.lib = ctypes.CDLL(_build.driver(quad))
        self.path = _build.model(model, quad=quad, contract=contract, libm=libm)
        self.h = ctypes.c_void_p()
        fn = self._fn("oracle_open")
        fn.restype = ctypes.c_int
        rc = fn(self.path.encode(), ctypes.byref(self.h))
        if rc != 0:
            raise RuntimeError(f"oracle_open({self.path}) failed with code {rc}")
        f = self._fn("oracle_n_fields")
        f.restype = ctypes.c_uint32
        self.n_fields = f(self.h)
        f = self._fn("oracle_n_params")
        f.restype = ctypes.c_uint32
        self.n_params = f(self.h)
        self.meta = golden_meta(model)

    def _fn(self, name):
        return getattr(self.lib, name + self.sfx)

    @staticmethod
    def _d(a):
        return a.ctypes.data_as(_DBL)

    @staticmethod
    def _ss(extent):
        ss = np.ascontiguousarray(np.asarray(extent, dtype=np.float64).reshape(4))
        return ss

    def _grid(self, fname, p, n0, n1, extent, per_point, rows=None, threads=0, dtype=np.float64,
              extra=()):  # fmt: skip
        p = np.ascontiguousarray(p, dtype=np.float64)
        assert p.shape == (self.n_params,)
        r0, r1 = rows if rows is not None else (0, n0)
        shape = (r1 - r0, n1) + ((per_point,) if per_point > 1 else ())
        out = np.zeros(shape, dtype=dtype)
        ss = self._ss(extent)
        fn = self._fn(fname)
        fn.restype = None
        fn(self.h, self._d(p), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(n0),
           ctypes.c_uint64(n1), self._d(ss), ctypes.c_uint64(r0), ctypes.c_uint64(r1), *extra,
           ctypes.c_int(threads))  # fmt: skip
        return out

    # extent = (x0_start, x0_stop, x1_start, x1_stop)
    def complete_analysis(self, p, n0, n1, extent, rows=None, threads=0):
        return self._grid("oracle_complete_analysis", p, n0, n1, extent, 6, rows, threads)

    def consistency_only(self, p, n0, n1, extent, rows=None, threads=0):
        return self._grid("oracle_consistency_only", p, n0, n1, extent, 1, rows, threads)

    def consistency_rapidturn_only(self, p, n0, n1, extent, rows=None, threads=0):
        return self._grid("oracle_consistency_rapidturn_only", p, n0, n1, extent, 1, rows, threads)

    def epsilon_v_only(self, p, n0, n1, extent, rows=None, threads=0):
        return self._grid("oracle_epsilon_v_only", p, n0, n1, extent, 1, rows, threads)

    def model_functions(self, p, n0, n1, extent, rows=None, threads=0):
        return self._grid("oracle_model_functions", p, n0, n1, extent, 5, rows, threads)

    def flag_quantum_dif(self, p, n0, n1, extent, accuracy, rows=None, threads=0):
        return self._grid("oracle_flag_quantum_dif", p, n0, n1, extent, 1, rows, threads,
                          dtype=np.uint8, extra=(ctypes.c_double(accuracy),)).astype(bool)  # fmt: skip

    def _traj(self, fname, p, xs, per_point, threads=0):
        p = np.ascontiguousarray(p, dtype=np.float64)
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        n = xs.shape[0]
        out = np.zeros((n, per_point) if per_point > 1 else (n,), dtype=np.float64)
        fn = self._fn(fname)
        fn.restype = None
        fn(self.h, self._d(p), self._d(xs), self._d(out), ctypes.c_uint64(n), ctypes.c_int(threads))
        return out

    def complete_analysis_on_trajectory(self, p, xs, threads=0):
        return self._traj("oracle_complete_analysis_on_trajectory", p, xs, 6, threads)

    def consistency_only_on_trajectory(self, p, xs, threads=0):
        return self._traj("oracle_consistency_only_on_trajectory", p, xs, 1, threads)

    def consistency_rapidturn_only_on_trajectory(self, p, xs, threads=0):
        return self._traj("oracle_consistency_rapidturn_only_on_trajectory", p, xs, 1, threads)

    def epsilon_v_only_on_trajectory(self, p, xs, threads=0):
        return self._traj("oracle_epsilon_v_only_on_trajectory", p, xs, 1, threads)

    def potential(self, x, p) -> float:
        x = np.ascontiguousarray(x, dtype=np.float64)
        p = np.ascontiguousarray(p, dtype=np.float64)
        fn = self._fn("oracle_potential")
        fn.restype = ctypes.c_double
        return fn(self.h, self._d(x), self._d(p))

    def grad_norm_squared(self, x, p) -> float:
        x = np.ascontiguousarray(x, dtype=np.float64)
        p = np.ascontiguousarray(p, dtype=np.float64)
        fn = self._fn("oracle_grad_norm_squared")
        fn.restype = ctypes.c_double
        return fn(self.h, self._d(x), self._d(p))

    def hesse(self, x, p):
        x = np.ascontiguousarray(x, dtype=np.float64)
        p = np.ascontiguousarray(p, dtype=np.float64)
        out = np.zeros((2, 2))
        fn = self._fn("oracle_hesse")
        fn.restype = None
        fn(self.h, self._d(x), self._d(p), self._d(out))
        return out

    def basis(self, which, x, p):
        x = np.ascontiguousarray(x, dtype=np.float64)
        p = np.ascontiguousarray(p, dtype=np.float64)
        out = np.zeros(2)
        fn = self._fn("oracle_basis")
        fn.restype = None
        fn(self.h, ctypes.c_int(which), self._d(x), self._d(p), self._d(out))
        return out

    def inner_prod(self, x, p, v1, v2) -> float:
        a = [np.ascontiguousarray(t, dtype=np.float64) for t in (x, p, v1, v2)]
        fn = self._fn("oracle_inner_prod")
        fn.restype = ctypes.c_double
        return fn(self.h, *[self._d(t) for t in a])

    def potential_array(self, p, n0, n1, extent):
        p = np.ascontiguousarray(p, dtype=np.float64)
        out = np.zeros((n0, n1))
        ss = self._ss(extent)
        fn = self._fn("oracle_potential_array")
        fn.restype = None
        fn(self.h, self._d(p), self._d(out), ctypes.c_uint64(n0), ctypes.c_uint64(n1), self._d(ss))
        return out

    def hesse_array(self, p, n0, n1, extent):
        p = np.ascontiguousarray(p, dtype=np.float64)
        out = np.zeros((2, 2, n0, n1))
        ss = self._ss(extent)
        fn = self._fn("oracle_hesse_array")
        fn.restype = None
        fn(self.h, self._d(p), self._d(out), ctypes.c_uint64(n0), ctypes.c_uint64(n1), self._d(ss))
        return out

    def close(self):
        if self.h:
            fn = self._fn("oracle_close")
            fn.restype = None
            fn(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
