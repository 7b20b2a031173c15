#!/usr/bin/env python3
"""Static cost of the row loop of a generated grid kernel, from its SASS (no GPU needed).

Measured on B200 (profiles/ncu_c3_r1.md, round 2 A/B runs): the grid kernels are issue bound with an
FP64 warp instruction occupying its SM sub-partition's issue port for 2 cycles and every other
instruction for 1, i.e.

    cycles per warp and grid row ~= 2 * N_fp64 + N_other      (hot path, slow path excluded)
    kernel time ~= cycles * (points / 32) / (n_SM * 4 sub-partitions * f_SM)

which reproduces ncu's FP64-pipe utilisation (EGNO: 2*436 / 1120 = 77.9 % vs 77.7 % measured), its
issue-slot utilisation (684 / 1120 = 61.1 % vs 61 %) and the kernel time (8.08 ms vs 8.17 ms).  So
a change can be costed offline: one FP64 instruction saved = 2 cycles, any other = 1.

    python tools/sass_cost.py egno [complete_analysis] [--minb 5] [--extra=-DINFLX_...] [--cols never]
"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("INFLATOX_CACHE_DIR", os.path.join(ROOT, "tests", ".cubin_cache"))
os.environ.setdefault("INFLATOX_QUIET", "1")

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
GROUP_OF = {"complete_analysis": "cmp", "consistency_only": "con", "consistency_rapidturn_only": "con",
            "epsilon_v_only": "eps", "potential": "pot", "hesse": "hes"}  # fmt: skip


def group_cubin(path: str, group: str) -> bytes:
    from inflatox_b200.compiler import _GROUP, _HEADER

    blob = open(path, "rb").read()
    n_groups = _HEADER.unpack(blob[: _HEADER.size])[8]
    for i in range(n_groups):
        g = _GROUP.unpack(blob[_HEADER.size + i * _GROUP.size: _HEADER.size + (i + 1) * _GROUP.size])
        if g[0].decode().strip("\0") == group:
            return blob[g[4]: g[4] + g[5]]
    raise KeyError(group)


def sass(cubin: bytes, kernel: str):
    with tempfile.NamedTemporaryFile(suffix=".cubin") as fh:
        fh.write(cubin)
        fh.flush()
        out = subprocess.run(["cuobjdump", "-sass", "-fun", kernel, fh.name], capture_output=True,
                             text=True, check=True).stdout  # fmt: skip
        res = subprocess.run(["cuobjdump", "-res-usage", fh.name], capture_output=True, text=True).stdout
    ins = []
    for ln in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    usage = ""
    lines = res.splitlines()
    for j, ln in enumerate(lines):
        if f"Function {kernel}:" in ln:
            usage = lines[j + 1].strip()
    return ins, usage


def hot_path(ins):
    """(loop start, loop end, skipped range): the row loop is the longest backward branch; inside it
    the one long forward branch is the `if (!bad)` skip over the exact recomputation."""
    loop = None
    for a, t in ins:
        m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            tgt = int(m.group(1), 16)
            if loop is None or a - tgt > loop[1] - loop[0]:
                loop = (tgt, a)
    if loop is None:
        raise RuntimeError("no loop found")
    skip = (0, 0)
    for a, t in ins:
        m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if m and loop[0] <= a <= loop[1]:
            tgt = int(m.group(1), 16)
            if a < tgt <= loop[1] and tgt - a > skip[1] - skip[0] and tgt - a > 0x400:
                skip = (a + 0x10, tgt)
    return loop, skip


def cost(ins):
    loop, skip = hot_path(ins)
    c = collections.Counter()
    for a, t in ins:
        if not (loop[0] <= a <= loop[1]) or (skip[0] <= a < skip[1]):
            continue
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        op = t.split()[0]
        base = op.split(".")[0]
        if base == "IMAD" and ".MOV" in op:
            base = "IMAD.MOV"
        c[base] += 1
    n = sum(c.values())
    fp64 = sum(v for k, v in c.items() if k in FP64)
    return {"instructions": n, "fp64": fp64, "other": n - fp64, "cycles": 2 * fp64 + n - fp64,
            "fp64_busy": 2 * fp64 / (2 * fp64 + n - fp64), "mix": c}  # fmt: skip


def main():
    import cases
    import inflatox_b200 as ix

    ap = argparse.ArgumentParser()
    ap.add_argument("model")
    ap.add_argument("op", nargs="?", default="complete_analysis")
    ap.add_argument("--minb", type=int)
    ap.add_argument("--rpt", type=int, default=16)
    ap.add_argument("--extra", action="append", default=[])
    ap.add_argument("--libm")
    ap.add_argument("--cols")
    ap.add_argument("--points", type=float, default=16384.0**2)
    ap.add_argument("--mhz", type=float, default=1965.0)
    a = ap.parse_args()
    flags = [f for f in ix.Compiler.default_nvrtc_flags if not f.startswith("-DINFLX_RPT=")]
    flags.append(f"-DINFLX_RPT={a.rpt}")
    if a.minb:
        flags.append(f"-DINFLX_MIN_BLOCKS={a.minb}")
    flags += a.extra
    comp = ix.Compiler(cases.load_model(a.model), silent=True, cse=cases.golden_cse(a.model),
                       compiler_flags=flags)  # fmt: skip
    if a.libm:
        comp.libm = a.libm
    if a.cols:
        comp.cols_prepass = a.cols
    art = comp.compile()
    ins, usage = sass(group_cubin(art.shared_object_path, GROUP_OF[a.op]), f"inflx_grid_{a.op}")
    r = cost(ins)
    ms = r["cycles"] * (a.points / 32) / (148 * 4 * a.mhz * 1e6) * 1e3
    top = ", ".join(f"{k} {v}" for k, v in r["mix"].most_common(16))
    print(f"{a.model} {a.op}: {usage.split(' SHARED')[0]}")
    print(f"  hot path per row: {r['instructions']} instructions, {r['fp64']} FP64 + {r['other']} other "
          f"-> {r['cycles']} issue cycles per warp, FP64 pipe {100 * r['fp64_busy']:.1f} % busy; "
          f"model time for {a.points:.3g} points: {ms:.2f} ms")
    print(f"  {top}")


if __name__ == "__main__":
    main()
