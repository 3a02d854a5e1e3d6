"""bench.py's own arm on a small batch: one JSON line on stdout with every key of the measurement contract."""
import json
import os
import subprocess
import sys

import pytest

import util

pytestmark = pytest.mark.gpu


def test_bench_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(util.ROOT, "bench.py"), "--envs", "512", "--steps", "12", "--warmup", "3",
                        "--burn-in", "20", "--no-cpu-baseline"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in d, k
    assert d["metric"] == "env-steps/sec" and d["steps"] == 12 and d["n_gpus"] == 1 and d["dtype"] == "f64"
    # NoMove + scripted gaze: the K steps of a replica are ONE d2d_rollout launch; the per-step-launch figure rides along
    assert d["gpu_launches"] == 1 and d["value"] > 0 and d["vs_baseline"] is None and d["scaling"] == "weak"
    ssl = d["single_step_launches"]
    assert ssl["gpu_launches"] == 12 and ssl["value"] > 0 and abs(ssl["value"] - 512 / (ssl["ms_per_step"] * 1e-3)) <= 1e-6 * ssl["value"]
    assert abs(d["value"] - 512 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    assert d["method"]["replicas"] >= 2 and "workload" in d["config"] and "l2" in d["config"]   # a small batch needs replicas to stay L2-cold
    assert d["method"]["graph_replays"] >= 7 and d["method"]["timed_region_ms_this_rank"] >= 50.0
    assert len(d["per_rank"]) == 1 and abs(d["per_rank"][0]["ms_per_step_median"] - d["ms_per_step"]) < 1e-12
    assert "workloads" not in d                                              # only the default (config 2, full size) run carries them
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 512 * 8 and e["d2h_bytes_per_step"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
