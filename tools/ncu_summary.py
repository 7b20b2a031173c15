#!/usr/bin/env python3
"""Summarise .ncu-rep captures (read here, no GPU needed) into profiles/*.md + ncu_traffic.json.

    python tools/ncu_summary.py r1a=gpurun_out/prof_a.ncu-rep r1b=gpurun_out/prof_b.ncu-rep \
        --points 268435456 --out profiles/ncu_c3_r1.md --traffic-key C3
"""
import argparse
import collections
import csv
import io
import json
import os
import re
import subprocess

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]  # fmt: skip


def page(rep, name):
    out = subprocess.run(
        ["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True
    ).stdout
    return list(csv.reader(io.StringIO(out)))


def raw(rep):
    r = page(rep, "raw")
    return {k: (r[1][i], r[2][i]) for i, k in enumerate(r[0])}


def mix(rep, points):
    r = page(rep, "source")
    h = r[1]
    ia, isrc = h.index("Instructions Executed"), h.index("Source")
    ops = collections.Counter()
    for row in r[2:]:
        if len(row) <= ia:
            continue
        try:
            n = int(float(row[ia]))
        except ValueError:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_\.]+)", row[isrc])
        if m:
            op = m.group(2)
            key = op.split(".")[0]
            if key in ("IMAD", "MUFU", "LDG", "LDL", "STL"):
                key = ".".join(op.split(".")[:2])
            ops[key] += n
    pw = points / 32
    return sum(ops.values()) / pw, [(k, v / pw) for k, v in ops.most_common(16)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--points", type=int, default=16384 * 16384)
    ap.add_argument(
        "--title",
        default="inflx_grid_complete_analysis, BASELINE C3 (EGNO, 16384 x 16384), 1 x B200",
    )
    ap.add_argument("--notes", default="")
    ap.add_argument("--out", required=True)
    ap.add_argument("--traffic-key", default=None)
    a = ap.parse_args()
    reps = [r.split("=", 1) for r in a.reports]
    data = {lab: raw(path) for lab, path in reps}
    lines = [
        f"# ncu summary - {a.title}",
        "",
        "Command: `ncu --set full --clock-control none --import-source on -k regex:inflx_grid -c 1 "
        "python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e`,",
        "run only after the same command had exited 0 without ncu. Reports (binary, not committed): "
        + ", ".join(p for _, p in reps)
        + ".",
        "",
    ]
    lines += [
        "| metric | unit | " + " | ".join(lab for lab, _ in reps) + " |",
        "|---|---|" + "---|" * len(reps),
    ]
    for k in KEYS:
        unit = next((data[lab][k][0] for lab, _ in reps if k in data[lab]), "")
        vals = [data[lab].get(k, ("", ""))[1] for lab, _ in reps]
        vals = [f"{float(v):.4g}" if re.match(r"^-?[\d.]+(e[+-]?\d+)?$", v) else v for v in vals]
        lines.append(f"| `{k}` | {unit} | " + " | ".join(vals) + " |")
    lines.append("")
    for lab, path in reps:
        total, top = mix(path, a.points)
        lines.append(
            f"Executed warp-instructions per 32 points, `{lab}`: total {total:.0f}: "
            + ", ".join(f"{k} {v:.1f}" for k, v in top)
            + "."
        )
    if a.notes:
        lines += ["", a.notes]
    with open(a.out, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if a.traffic_key:
        lab = reps[-1][0]
        rd, wr = data[lab]["dram__bytes_read.sum"], data[lab]["dram__bytes_write.sum"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tr = float(rd[1]) * scale[rd[0]] + float(wr[1]) * scale[wr[0]]
        path = os.path.join(os.path.dirname(a.out), "ncu_traffic.json")
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur[a.traffic_key] = tr
        json.dump(cur, open(path, "w"))
    print(open(a.out).read())


if __name__ == "__main__":
    main()
