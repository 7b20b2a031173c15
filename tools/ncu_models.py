#!/usr/bin/env python3
"""Per-model ncu table (profiles/ncu_models_r<N>.md) from one `ncu --set full` capture per config.

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
    ("bytes written L1 -> L2 (output + spill stores)", "l1tex__m_l1tex2xbar_write_bytes.sum"),
    ("local-memory (spill) store sectors (32 B)", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"),
    ("local-memory (spill) load sectors", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"),
    ("... of which missed L1 (reloaded from L2)", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld_lookup_miss.sum"),
    ("warp instructions executed", "smsp__inst_executed.sum"),
    ("stall: wait / issue", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall: math pipe throttle / issue", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall: not selected / issue", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    ("stall: long scoreboard / issue", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
]  # fmt: skip
TITLES = {
    "C1": "C1: hyperinflation, complete_analysis, 1000^2",
    "C2": "C2: angular, consistency_only, 4096^2",
    "C3": "C3: EGNO, complete_analysis, 16384^2",
    "C4": "C4: D5-brane, complete_analysis, 16384^2",
    "C6": "C6: angular, complete_analysis, 16384^2",
    "C7": "C7: hyperinflation, complete_analysis, 16384^2",
    "C8": "C8: doc model, complete_analysis, 16384^2",
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
        "bench.py --config Cx --steps 2 --warmup 1 --no-cpu --no-e2e --no-all` per config, each after the "
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
