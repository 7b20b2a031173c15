"""Assembles an importable `inflatox` package = the reference's UNMODIFIED symbolic front-end on top
of this back-end, so that `import inflatox; inflatox.Compiler(model).compile()` and
`from inflatox.consistency_conditions import GeneralisedAL` - the calls of BASELINE.json's north_star
and of the reference's own tests - run on the GPU path without a single edit in user code.

    python -m inflatox_b200.overlay /path/to/reference/checkout /path/to/site-dir

Layout written under the destination (put `<dest>` first and `<dest>/_stubs` last on PYTHONPATH;
git-ignored here as baseline/_ref, built by `__graft_entry__.build()` when the reference checkout is
present):

    inflatox/__init__.py, symbolic.py, version.py, background.py   copied verbatim from the reference
                                                                   (python/inflatox/; symbolic model
                                                                   builder = SURVEY.md §2 #12, out of
                                                                   scope and used as is)
    inflatox/compiler.py, consistency_conditions.py, libinflx_rs.py
                                                                   three-line shims re-exporting
                                                                   inflatox_b200's modules under the
                                                                   reference's module names
    inflatox-<version>.dist-info/METADATA                          so that version.py's
                                                                   importlib.metadata lookup resolves
    _stubs/interruptingcow.py, _stubs/IPython/display.py           stand-ins for the two third-party
                                                                   packages the reference imports that
                                                                   this image lacks (a SIGALRM `timeout`
                                                                   context manager; `display`, which
                                                                   symbolic.py:281 calls when a builder
                                                                   is not silent).  Put `_stubs` LAST on
                                                                   PYTHONPATH: a real installation wins
    reference_tests/                                               the reference's tests/ directory,
                                                                   verbatim (acceptance tests)
    MANIFEST.json                                                  sha256 of every copied file

Nothing of the reference is stored in this repository: the copies are made at build time from the
checkout and live in a git-ignored directory.
"""
from __future__ import annotations

import hashlib
import json
import os
import re
import shutil
import sys

COPIED_MODULES = ("__init__.py", "symbolic.py", "version.py", "background.py")

SHIMS = {
    "compiler.py": (
        '"""inflatox.compiler -> inflatox_b200.compiler (CUDA back-end; same names and arguments as\n'
        'reference python/inflatox/compiler.py)."""\n'
        "from inflatox_b200.compiler import *  # noqa: F401,F403\n"
        "from inflatox_b200.compiler import (  # noqa: F401\n"
        "    CInflatoxPrinter, CompilationArtifact, Compiler, GSLInflatoxPrinter,\n"
        ")\n"
    ),
    "consistency_conditions.py": (
        '"""inflatox.consistency_conditions -> inflatox_b200.consistency_conditions."""\n'
        "from inflatox_b200.consistency_conditions import *  # noqa: F401,F403\n"
        "from inflatox_b200.consistency_conditions import GeneralisedAL, InflationCondition  # noqa: F401\n"
    ),
    "libinflx_rs.py": (
        '"""inflatox.libinflx_rs -> inflatox_b200.libinflx_rs (ctypes binding of the C ABI in\n'
        'include/inflx_b200.h; replaces the pyo3 module of reference src/lib.rs:68-92)."""\n'
        "from inflatox_b200.libinflx_rs import *  # noqa: F401,F403\n"
        "from inflatox_b200.libinflx_rs import (  # noqa: F401\n"
        "    InflatoxPyDyLib, log_info, log_warn, open_inflx_dylib, solve_eom_rk4, solve_eom_rkf,\n"
        ")\n"
    ),
}

INTERRUPTINGCOW = '''"""Minimal stand-in for the `interruptingcow` package (absent from this image): the one name the
reference uses, `timeout(seconds, exception)` (reference python/inflatox/symbolic.py:22, 231-265),
as a SIGALRM context manager.  Outside the main thread no alarm can be armed and the block simply
runs without a time limit."""
import contextlib
import signal
import threading


@contextlib.contextmanager
def timeout(seconds, exception=RuntimeError):
    if threading.current_thread() is not threading.main_thread() or seconds is None or seconds <= 0:
        yield
        return

    def _raise(signum, frame):
        raise exception

    previous = signal.signal(signal.SIGALRM, _raise)
    signal.setitimer(signal.ITIMER_REAL, float(seconds))
    try:
        yield
    finally:
        signal.setitimer(signal.ITIMER_REAL, 0.0)
        signal.signal(signal.SIGALRM, previous)
'''


IPYTHON_DISPLAY = '''"""Minimal stand-in for IPython.display (IPython is absent from this image): the reference's
non-silent model builder shows its intermediate expressions with `display`
(reference python/inflatox/symbolic.py:274-284); here they are pretty-printed to stdout."""


def display(*objs, **kwargs):
    try:
        import sympy

        for o in objs:
            print(sympy.pretty(o))
    except Exception:
        for o in objs:
            print(o)
'''


def _sha(path: str) -> str:
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def reference_version(reference_root: str) -> str:
    with open(os.path.join(reference_root, "pyproject.toml")) as fh:
        m = re.search(r'^version\s*=\s*"([^"]+)"', fh.read(), re.M)
    if not m:
        raise RuntimeError("no version in the reference's pyproject.toml")
    return m.group(1)


def assemble(reference_root: str, dest: str) -> dict:
    """Build the overlay; returns the manifest (also written to <dest>/MANIFEST.json)."""
    src_pkg = os.path.join(reference_root, "python", "inflatox")
    if not os.path.isdir(src_pkg):
        raise FileNotFoundError(f"{src_pkg}: not a checkout of smups/inflatox")
    version = reference_version(reference_root)
    pkg = os.path.join(dest, "inflatox")
    tests = os.path.join(dest, "reference_tests")
    for d in (pkg, tests):
        shutil.rmtree(d, ignore_errors=True)
    os.makedirs(pkg)
    manifest = {"reference_version": version, "copied": {}, "shims": sorted(SHIMS)}
    for name in COPIED_MODULES:
        shutil.copyfile(os.path.join(src_pkg, name), os.path.join(pkg, name))
        manifest["copied"][f"inflatox/{name}"] = _sha(os.path.join(pkg, name))
    for name, text in SHIMS.items():
        with open(os.path.join(pkg, name), "w") as fh:
            fh.write(text)
    stubs = os.path.join(dest, "_stubs")
    shutil.rmtree(stubs, ignore_errors=True)
    os.makedirs(os.path.join(stubs, "IPython"))
    with open(os.path.join(stubs, "interruptingcow.py"), "w") as fh:
        fh.write(INTERRUPTINGCOW)
    with open(os.path.join(stubs, "IPython", "__init__.py"), "w") as fh:
        fh.write('"""Stand-in package, see display.py."""\n')
    with open(os.path.join(stubs, "IPython", "display.py"), "w") as fh:
        fh.write(IPYTHON_DISPLAY)
    if os.path.exists(os.path.join(dest, "interruptingcow.py")):
        os.remove(os.path.join(dest, "interruptingcow.py"))
    info = os.path.join(dest, f"inflatox-{version}.dist-info")
    shutil.rmtree(info, ignore_errors=True)
    os.makedirs(info)
    with open(os.path.join(info, "METADATA"), "w") as fh:
        fh.write(f"Metadata-Version: 2.1\nName: inflatox\nVersion: {version}\n"
                 "Summary: reference symbolic front-end + inflatox_b200 CUDA back-end (overlay)\n")
    with open(os.path.join(info, "INSTALLER"), "w") as fh:
        fh.write("inflatox_b200.overlay\n")
    shutil.copytree(os.path.join(reference_root, "tests"), tests)
    for root, _dirs, files in os.walk(tests):
        for f in sorted(files):
            path = os.path.join(root, f)
            manifest["copied"]["reference_tests/" + os.path.relpath(path, tests)] = _sha(path)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    return manifest


if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    m = assemble(sys.argv[1], sys.argv[2])
    print(f"inflatox {m['reference_version']} overlay: {len(m['copied'])} files copied, "
          f"{len(m['shims'])} shims -> {sys.argv[2]}")
