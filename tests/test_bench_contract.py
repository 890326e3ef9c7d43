"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the
contract's keys; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess

import pytest
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=env, cwd=ROOT)


def test_reference_arm_line():
    r = _run({}, "--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "lpost+grad evals/s" and d["unit"] == "evals/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["config"]["n"] == 100_000_000 and d["config"]["p"] == 64 and "HMC" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "rows" in cb["sample"] and cb["value"] == d["value"]
    # measured at two sample sizes, a line through them gives the figure at the full n
    assert cb["sizes"] == [1_000_000, 4_000_000] and len(cb["ms_at_size"]) == 2 and cb["ms_at_size"][1] > cb["ms_at_size"][0]
    fit = cb["fit"]
    t_full = fit["intercept_ms"] + fit["ms_per_1e6_rows"] * 100.0
    assert d["ms_per_step_at_full_n"] == pytest.approx(t_full, rel=1e-9)
    assert d["value"] == pytest.approx(20 * 1e3 / t_full, rel=1e-9)            # L = 20 evaluations per HMC iteration
    # ms_per_step is the measured time of a step of the bounded sample (so steps x ms_per_step is this run's work)
    assert d["ms_per_step"] == pytest.approx(cb["ms_at_size"][0], rel=1e-9)
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--gpus", "2", "--steps", "1")
    assert r.returncode == 0 and r.stdout.strip() == ""
