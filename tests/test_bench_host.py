"""Host-side pieces of bench.py that need no GPU: the clock sampler, the workload table, the reference arm's JSON line."""
import json
import os
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


FAKE_SMI = """#!/bin/bash
# stands in for `nvidia-smi --query-gpu=... -lms 20`: slow to start, then one CSV line every 20 ms
sleep %s
while true; do echo "0, %s, 1965, 512.3, 0x0000000000000004, Not Active, Not Active, Not Active, %s"; sleep 0.02; done
"""


def _fake_smi(tmp_path, monkeypatch, startup="0.3", clock="1965", power_cap="Active"):
    p = tmp_path / "nvidia-smi"
    p.write_text(FAKE_SMI % (startup, clock, power_cap))
    p.chmod(0o755)
    monkeypatch.setenv("PATH", f"{tmp_path}:{os.environ['PATH']}")


def test_clock_sampler_waits_for_first_sample_and_windows(tmp_path, monkeypatch):
    _fake_smi(tmp_path, monkeypatch)
    s = bench.ClockSampler(0)
    s.start()
    s.begin()                      # blocks until the slow sampler delivers its first line
    assert s.rows, "begin() must not return before a sample exists"
    time.sleep(0.15)               # the "timed region"
    out = s.stop()
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0
    assert out["samples"] >= 3
    assert out["reasons"] == ["sw_power_cap"]


def test_clock_sampler_short_region_uses_nearest_samples(tmp_path, monkeypatch):
    _fake_smi(tmp_path, monkeypatch, startup="0.05", power_cap="Not Active")
    s = bench.ClockSampler(0)
    s.start()
    s.begin()
    out = s.stop()                 # region of ~0 s
    assert out["samples"] >= 1 and out["sm_mhz"] == 1965.0 and out["reasons"] == []


def test_clock_sampler_without_nvidia_smi(tmp_path, monkeypatch):
    monkeypatch.setenv("PATH", str(tmp_path))  # nothing there
    s = bench.ClockSampler(0)
    s.start()
    s.begin(timeout_s=0.2)
    out = s.stop()
    assert out["sm_mhz"] is None and out["samples"] == 0


def test_workload_table_matches_baseline_json():
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "c1" in bench.WORKLOADS and "c0" in bench.WORKLOADS
    r, s, q, variant, m, k, B, desc = bench.WORKLOADS["c1"]
    assert (r, s, q, variant, m, k) == (128_000_000, 1_024_000_000, 0.01, 0, 1 << 30, 1)
    # the metric is the one BASELINE.json names
    assert "tuples" in base["metric"].lower() and "tuples" in bench.METRIC.lower()


def test_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` runs the compiled, unmodified reference on the host cores (no GPU involved)."""
    import oracle
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref.so")):
        pytest.skip("oracle/_ref not built")
    env = dict(os.environ, HWBRJ_BENCH_WATCHDOG_S="240")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"]
