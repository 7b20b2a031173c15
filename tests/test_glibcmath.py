"""csrc/inflx_glibcmath.cuh - the reference host's libm (glibc 2.39, x86_64 FMA ifunc variants)
restated for the device - compiled for the HOST and compared with the host's own libm: every result
must be bit-identical (random arguments over every exponent, model-like ranges, irregular operands).
This is the CPU half of the claim "hoisted libm calls return the reference's bits"; the device build
of the same file differs only in how + - * fma are spelled (_rn intrinsics), which
tests/test_gpu_numerics.py checks on the GPU against this host build.

INFLX_GLIBC_CHECK_N raises the sample (default 10^7 arguments per function, a few seconds; the run
with 10^9 is recorded in profiles/glibcmath_r2.txt)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "inflatox_b200", "csrc")
HEADER = os.path.join(CSRC, "inflx_glibcmath.cuh")


def _host_glibc() -> str:
    f = ctypes.CDLL("libc.so.6").gnu_get_libc_version
    f.restype = ctypes.c_char_p
    return f().decode()


def _has_fma() -> bool:
    with open("/proc/cpuinfo") as fh:
        flags = next((ln for ln in fh if ln.startswith("flags")), "")
    return " fma " in flags + " " and " avx2 " in flags + " "


pinned_host = pytest.mark.skipif(
    not (_host_glibc() == "2.39" and _has_fma()),
    reason="bit identity is pinned to glibc 2.39's FMA ifunc variants (the oracle host's libm)",
)


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = tmp_path_factory.mktemp("glibcmath") / "glibcmath_check"
    subprocess.run(
        ["gcc", "-O2", "-march=native", "-ffp-contract=off", "-fno-builtin", "-fopenmp", "-x", "c",
         f"-I{CSRC}", f'-DINFLX_GLIBCMATH_HEADER="{HEADER}"',
         os.path.join(ROOT, "tests", "native", "glibcmath_check.c"), "-o", str(exe), "-lm"],
        check=True,
    )  # fmt: skip
    return str(exe)


@pinned_host
def test_every_result_has_the_bits_of_the_host_libm(checker):
    n = int(os.environ.get("INFLX_GLIBC_CHECK_N", "10000000"))
    out = subprocess.run([checker, str(n)], capture_output=True, text=True, check=True).stdout
    c = {k: int(v) for k, v in re.findall(r"(\w+)=(\d+)", out.splitlines()[0])}
    assert c["n"] == n
    for k in ("bad_pow", "bad_exp", "bad_log", "bad_sin", "bad_cos", "bad_tanh", "bad_expm1", "bad_atan", "bad_tan",
              "bad_special"):
        assert c[k] == 0, out


@pinned_host
def test_tables_are_the_host_libm_s():
    """The committed tables are the ones tools/gen_glibc_tables.py reads out of this host's libm."""
    gen = subprocess.run(
        ["python", os.path.join(ROOT, "tools", "gen_glibc_tables.py")],
        capture_output=True, text=True, check=True,
    ).stdout  # fmt: skip
    with open(os.path.join(CSRC, "inflx_glibc_tables.cuh")) as fh:
        assert gen == fh.read()
