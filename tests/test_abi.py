"""The C-ABI library loads without a GPU and exports every symbol include/inflx_b200.h declares;
argument validation and error mapping mirror the reference (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest

import cases
from inflatox_b200 import _native
from inflatox_b200 import libinflx_rs as rs


def declared_functions() -> list[str]:
    with open(_native.HEADER_PATH) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(inflx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.lib()
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/inflx_b200.h but not exported"


def test_library_is_in_tree_and_not_linked_against_libcuda():
    assert os.path.dirname(_native.LIB_PATH).endswith("inflatox_b200")
    with open(_native.LIB_PATH, "rb") as fh:
        blob = fh.read()
    # bound at run time with dlopen (csrc/inflx_cuda_dl.h), so the ABI loads on a CPU-only host
    assert b"libcuda.so.1" in blob


def test_open_errors_mirror_the_reference(tmp_path):
    with pytest.raises(IOError, match="Could not load Inflatox Compilation Artefact"):
        rs.open_inflx_dylib(str(tmp_path / "missing.bin"), False)
    junk = tmp_path / "junk.bin"
    junk.write_bytes(b"\x7fELF" + b"\0" * 400)
    with pytest.raises(IOError):
        rs.open_inflx_dylib(str(junk), False)
    # ABI check: major.minor must match (reference src/inflatox_version.rs:48-53)
    art = cases.artifact("doc")
    blob = bytearray(open(art.shared_object_path, "rb").read())
    blob[12:14] = (4).to_bytes(2, "little")
    old = tmp_path / "old.bin"
    old.write_bytes(bytes(blob))
    with pytest.raises(SystemError, match="compiled for Inflatox ABI v4.0.0"):
        rs.open_inflx_dylib(str(old), False)
    blob[12:14] = (5).to_bytes(2, "little")
    blob[16:18] = (9).to_bytes(2, "little")  # patch level is ignored
    ok = tmp_path / "patch.bin"
    ok.write_bytes(bytes(blob))
    assert rs.open_inflx_dylib(str(ok), False).n_fields == 2
    # container layout: an artefact of an older kernel ABI must be refused, not launched
    blob[8:12] = (1).to_bytes(4, "little")
    stale = tmp_path / "stale.bin"
    stale.write_bytes(bytes(blob))
    with pytest.raises(SystemError, match="container layout v1"):
        rs.open_inflx_dylib(str(stale), False)


def test_handle_metadata():
    lib = rs.open_inflx_dylib(cases.artifact("d5").shared_object_path, False)
    assert (lib.n_fields, lib.n_parameters, lib.name) == (2, 10, "d5")
    lib.set_devices([3, 1])
    assert lib.devices() == [3, 1]


def test_argument_validation_precedes_any_device_work():
    lib = rs.open_inflx_dylib(cases.artifact("egno").shared_object_path, False)
    ss = np.array([[0.0, 1.0], [0.0, 1.0]])
    good_p = cases.params("egno")
    with pytest.raises(Exception, match='model "egno" has 4 paramters'):
        rs.complete_analysis(lib, np.zeros(3), np.zeros((4, 4, 6)), ss, False, 0)
    with pytest.raises(Exception, match="Last axis must have lenght 6"):
        rs.complete_analysis(lib, good_p, np.zeros((4, 4, 5)), ss, False, 0)
    with pytest.raises(Exception, match="start_stop array should have 2 rows"):
        rs.consistency_only(lib, good_p, np.zeros((4, 4)), np.zeros((3, 2)), False, 0)
    with pytest.raises(rs.PanicException, match="C-CONTIGUOUS"):
        rs.consistency_only(lib, good_p, np.zeros((4, 8))[:, ::2], ss, False, 0)
    with pytest.raises(Exception, match="First axis of output array"):
        rs.complete_analysis_on_trajectory(lib, good_p, np.zeros((5, 2)), np.zeros((4, 6)), False, 1)
    with pytest.raises(Exception, match="as many elements as there are field-space coordinates"):
        lib.potential(np.zeros(3), good_p)
    with pytest.raises(TypeError):
        rs.consistency_only(lib, good_p, np.zeros((4, 4), dtype=np.float32), ss, False, 0)


@pytest.mark.skipif(cases is None, reason="")
def test_compute_fails_loudly_without_a_gpu():
    if _native.lib().inflx_device_count() > 0:
        pytest.skip("a GPU is present")
    lib = rs.open_inflx_dylib(cases.artifact("doc").shared_object_path, False)
    with pytest.raises(SystemError, match="CUDA"):
        rs.complete_analysis(lib, np.array([1.0]), np.zeros((4, 4, 6)),
                             np.array([[0.0, 1.0], [0.0, 1.0]]), False, 0)  # fmt: skip
    with pytest.raises(SystemError, match="CUDA"):
        rs.open_inflx_dylib(cases.artifact("doc").shared_object_path, True)  # basis check needs the GPU


def test_product_code_never_touches_the_oracle():
    root = os.path.join(cases.ROOT, "inflatox_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), f
                assert "liboracle" not in text and "oracle/" not in text.replace("# oracle/", ""), f


def test_cpp_host_example_builds_against_the_header_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/complete_analysis.cpp: a C++ embedder over include/inflx_b200.h only.  It must
    compile warning-free, read the artefact's metadata without a GPU, and - on a box without a
    CUDA driver - stop with the engine's own message instead of producing numbers."""
    import subprocess

    from inflatox_b200 import _native

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    _native.lib()  # builds the engine in tree if it is missing
    libdir = os.path.dirname(_native.LIB_PATH)
    exe = tmp_path / "complete_analysis"
    subprocess.run(
        ["g++", "-std=c++17", "-Wall", "-Werror", f"-I{root}/include",
         f"{root}/examples/complete_analysis.cpp", "-o", str(exe), f"-L{libdir}", "-linflx_b200",
         f"-Wl,-rpath,{libdir}"], check=True,
    )  # fmt: skip
    art = cases.artifact("doc")
    r = subprocess.run(
        [str(exe), art.shared_object_path, "32", "32", "0", "2.5", "0", "3.14", "1.0"],
        capture_output=True, text=True,
    )  # fmt: skip
    assert "2 fields, 1 parameters, artefact ABI 5.0.0" in r.stdout
    if _native.lib().inflx_device_count() < 1:
        assert r.returncode == 1 and "failed (status 9)" in r.stderr
    else:
        assert r.returncode == 0 and "consistency" in r.stdout
