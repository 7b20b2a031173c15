"""bench.py's driver contract, checked on the arm that needs no GPU (`--impl reference`)."""
import json
import os
import subprocess
import sys

import cases


def run(*args):
    r = subprocess.run(
        [sys.executable, os.path.join(cases.ROOT, "bench.py"), *args],
        capture_output=True, text=True, timeout=600,
    )  # fmt: skip
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_prints_exactly_one_json_line():
    out = run("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-seconds", "0.3",
              "--config", "C1")  # fmt: skip
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "grid_points_per_s"
    assert d["unit"] == "points/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    assert d["config"]["workload"].startswith("C1: hyper model, complete_analysis, 1000x1000")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == os.cpu_count() and "rows [" in cb["sample"]
    assert cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}  # fmt: skip


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run(
        [sys.executable, os.path.join(cases.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
         "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env,
    )  # fmt: skip
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_default_workload_is_the_metrics_config():
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench", os.path.join(cases.ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    model, op, n0, n1, S, ext, p = bench.workload("C3")
    assert (model, op, n0, n1, S) == ("egno", "complete_analysis", 16384, 16384, 1)
    assert list(p[0]) == cases.PARAMS["egno"]
    model, op, n0, n1, S, ext, p = bench.workload("C5")
    assert S == 1024 and p.shape == (1024, 3) and (n0, n1) == (1024, 1024)
    # args order of the hyperinflation model is (m, phi0, L): BASELINE C5's ranges per column
    assert p[:, 0].min() >= 1e-3 and p[:, 0].max() <= 10 and abs(p[:, 1]).max() <= 1
    assert p[:, 2].min() >= 0.05 and p[:, 2].max() <= 2.0


def test_every_baseline_config_and_every_repo_model_has_a_bench_config():
    """BASELINE.json's five configs + complete_analysis on a 16384^2 grid for each remaining repo test
    model (north_star: ">= 50 % of the FP64 roofline ... for each repo test model"): the default
    `python bench.py` run reports all of them under `configs` next to the C3 headline."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench", os.path.join(cases.ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert sorted(bench.CONFIGS) == ["C1", "C2", "C3", "C4", "C5", "C6", "C7", "C8"]
    full = {bench.CONFIGS[c][0] for c in bench.CONFIGS
            if bench.CONFIGS[c][1:4] == ("complete_analysis", 16384, 16384)}
    assert full == {"egno", "d5", "angular", "hyper", "doc"} == set(cases.MODELS)
    assert bench.CPU_CONTRACT == "fast"  # the faster CPU build is the baseline
