#!/usr/bin/env python3
"""Per-model ncu table (profiles/ncu_models_r1.md) from one `ncu --set full` capture per config.

    python tools/ncu_models.py C1=gpurun_out/prof_C1_r1k.ncu-rep C2=… --out profiles/ncu_models_r1.md
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_summary import raw  # noqa: E402

ROWS = [
    ("kernel time", "gpu__time_duration.sum"),
    ("CTAs", "launch__grid_size"),
    ("registers/thread", "launch__registers_per_thread"),
    ("CTAs/SM allowed by registers", "launch__occupancy_limit_registers"),
    ("CTAs/SM allowed by shared memory", "launch__occupancy_limit_shared_mem"),
    ("achieved occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("FP64 pipe busy %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("XU/SFU pipe busy % (MUFU seeds, conversions)", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("DRAM written", "dram__bytes_write.sum"),
    ("DRAM read", "dram__bytes_read.sum"),
]  # fmt: skip
TITLES = {
    "C1": "C1: hyperinflation, complete_analysis, 1000^2",
    "C2": "C2: angular, consistency_only, 4096^2",
    "C3": "C3: EGNO, complete_analysis, 16384^2",
    "C4": "C4: D5-brane, complete_analysis, 16384^2",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--out", required=True)
    ap.add_argument("--notes", default="")
    a = ap.parse_args()
    reps = [r.split("=", 1) for r in a.reports]
    data = {lab: raw(path) for lab, path in reps}
    lines = [
        "# ncu per model - dominant kernel `inflx_grid_<op>`, 1 x B200",
        "",
        "One `ncu --set full --clock-control none --import-source on -k regex:inflx_grid -c 1 python "
        "bench.py --config Cx --steps 2 --warmup 1 --no-cpu --no-e2e` per config, each after the "
        "same command had exited 0 without ncu (reports: "
        + ", ".join(p for _, p in reps)
        + "; binary, not committed).",
        "",
        "| metric | " + " | ".join(TITLES.get(lab, lab) for lab, _ in reps) + " |",
        "|---|" + "---|" * len(reps),
    ]
    for title, key in ROWS:
        cells = []
        for lab, _ in reps:
            unit, val = data[lab].get(key, ("", ""))
            try:
                val = f"{float(val):.4g}"
            except ValueError:
                pass
            cells.append(f"{val} {unit}".strip())
        lines.append(f"| {title} (`{key}`) | " + " | ".join(cells) + " |")
    if a.notes:
        lines += ["", a.notes]
    with open(a.out, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print(open(a.out).read())


if __name__ == "__main__":
    main()
