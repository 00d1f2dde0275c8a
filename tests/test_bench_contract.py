"""The driver-facing contract of `bench.py --impl reference` (the CPU arm: runs without a GPU): one JSON line on stdout
with the keys the driver reads, on a small bounded sample so the test takes seconds.  Under torchrun only rank 0 prints."""

import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample-rows", "50000"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["metric"].startswith("queries/sec exact top-100") and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_clock_sampler_counts_only_samples_taken_under_load(tmp_path, monkeypatch):
    """bench.ClockSampler with a stand-in nvidia-smi: the process is spawned early (prestart), samples count from
    start() on, throttle reasons are collected, and a missing nvidia-smi degrades to a labelled empty result."""
    import importlib.util
    import stat
    import time
    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/sh\ni=0\nwhile [ $i -lt 200 ]; do\n"
                    "  if [ $i -lt 5 ]; then echo '345, 1965, 90.0, Not Active, Not Active, Not Active, Not Active';\n"
                    "  else echo '1700, 1965, 990.0, Not Active, Not Active, Not Active, Active'; fi\n"
                    "  i=$((i+1)); sleep 0.05\ndone\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", f"{tmp_path}:{os.environ['PATH']}")
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler(0)
    s.prestart()
    time.sleep(0.6)                 # the five idle samples (345 MHz) arrive before the load starts
    s.start()
    time.sleep(0.5)
    out = s.stop()
    assert out["samples"] >= 5 and out["sm_mhz"] == 1700.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]
    monkeypatch.setenv("PATH", str(tmp_path / "nowhere"))
    s2 = bench.ClockSampler(0)
    s2.start()
    assert s2.stop()["sm_mhz"] is None
