"""The reference's OWN test files, unmodified, as the acceptance gate:
they are copied verbatim from the reference checkout into the git-ignored overlay
(`inflatox_b200.overlay.assemble`, run by `__graft_entry__.build()`), next to an `inflatox` package
made of the reference's untouched symbolic front-end and three shims that re-export this back-end
under the reference's module names.  Each file runs in its own pytest subprocess with the overlay on
PYTHONPATH (joblib's worker processes of symbolic.py:376 re-import the package, so the path must
come from the environment, not from sys.path edits).

test_compiler.py / test_symbolic.py need no GPU (printer strings, symbolic identities);
test_doc.py / test_angular.py / test_egno.py / test_d5.py build each model with the reference's
InflationModelBuilder, compile it with `inflatox.Compiler` (here: CUDA back-end) and run the
GeneralisedAL analyses on the GPU - test_doc.py holds the reference's only known answers
(calc_V == 1.9166666666666667, calc_H, nanmax(consistency) <= 1)."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

import cases

OVERLAY = os.path.join(cases.ROOT, "baseline", "_ref")
REFERENCE = "/root/reference"


def _overlay() -> str:
    if not os.path.isdir(os.path.join(OVERLAY, "inflatox")):
        if os.path.isdir(REFERENCE):
            from inflatox_b200 import overlay

            overlay.assemble(REFERENCE, OVERLAY)
        else:
            pytest.skip("no reference overlay in this tree (run __graft_entry__.build() where the "
                        "reference checkout exists)")
    return OVERLAY


def _run_reference_test(name: str, timeout: int) -> None:
    site = _overlay()
    env = dict(os.environ)
    # overlay first, the stand-ins for absent third-party packages last (a real one wins)
    env["PYTHONPATH"] = os.pathsep.join(
        [site, cases.ROOT] + [p for p in [env.get("PYTHONPATH")] if p] + [os.path.join(site, "_stubs")]
    )
    env.setdefault("INFLATOX_CACHE_DIR", os.path.join(cases.ROOT, "tests", ".cubin_cache"))
    for attempt in range(4):
        r = subprocess.run(
            [sys.executable, "-m", "pytest", os.path.join(site, "reference_tests", name), "-q", "-x",
             "-p", "no:cacheprovider", "--rootdir", os.path.join(site, "reference_tests")],
            capture_output=True, text=True, timeout=timeout, env=env, cwd=site,
        )  # fmt: skip
        # the reference's load-time basis check draws RANDOM parameters (src/lib.rs:142-203) and
        # fails in ~8 % of the loads of the angular model on the reference's own CPU path as well
        # (tests/test_oracle.py): that upstream flake, and only that, is retried
        if r.returncode == 0 or "Expected basis vector" not in r.stdout + r.stderr:
            break
    assert r.returncode == 0, f"{name}\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-500:]


def test_overlay_holds_the_reference_files_verbatim():
    site = _overlay()
    with open(os.path.join(site, "MANIFEST.json")) as fh:
        manifest = json.load(fh)
    assert manifest["shims"] == ["compiler.py", "consistency_conditions.py", "libinflx_rs.py"]
    for rel, sha in manifest["copied"].items():
        with open(os.path.join(site, rel), "rb") as fh:
            assert hashlib.sha256(fh.read()).hexdigest() == sha, rel
        if os.path.isdir(REFERENCE):  # where the checkout exists: byte-identical to it
            src = rel.replace("inflatox/", "python/inflatox/", 1) if rel.startswith("inflatox/") \
                else rel.replace("reference_tests/", "tests/", 1)
            with open(os.path.join(REFERENCE, src), "rb") as fh:
                assert hashlib.sha256(fh.read()).hexdigest() == sha, rel
    for f in ("test_doc.py", "test_angular.py", "test_egno.py", "test_d5.py"):
        assert f"reference_tests/{f}" in manifest["copied"]


@pytest.mark.parametrize("name", ["test_compiler.py", "test_symbolic.py"])
def test_reference_unit_tests_unmodified(name):
    _run_reference_test(name, 600)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["test_doc.py", "test_angular.py", "test_egno.py", "test_d5.py"])
def test_reference_integration_tests_unmodified_on_the_gpu(name):
    _run_reference_test(name, 1500)
