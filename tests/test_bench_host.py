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


class _StubLib:
    """Stands in for libhwbrj_cuda.so so that bench.main() can assemble its JSON line without a GPU: every join
    'takes' fixed phase times and returns C1's published scalars."""

    def __init__(self):
        import ctypes as C
        from hwbloomradixjoin_b200 import _native as N
        self.C, self.N = C, N
        self.libc = C.CDLL(None)
        self.libc.malloc.restype = C.c_void_p
        self.libc.malloc.argtypes = [C.c_size_t]
        self.overlap = None

    def _fill(self, st):
        st.matches, st.filtered, st.checksum_pair = 10_240_000, 124_152_740, 12345
        st.ms_total, st.ms_memset, st.ms_build, st.ms_part_r = 8.6, 0.03, 0.4, 1.1
        st.ms_probe, st.ms_part_s, st.ms_join = 5.0, 1.2, 0.8
        st.ms_h2d, st.h2d_bytes, st.d2h_bytes = 166.0, 9_216_000_000, 136
        st.kernel_launches, st.radix_bits, st.range_passes, st.n_gpus = 13, 14, 2, 1

    def hwbrj_join_device(self, r, s, args, st_ref):
        self._fill(st_ref._obj)
        return 0

    def hwbrj_last_stats(self, st_ref):
        self._fill(st_ref._obj)
        return 0

    def hwbrj_host_alloc(self, nbytes):
        return 0x1000

    def hwbrj_host_free(self, p):
        pass

    def hwbrj_rel_download(self, h, p):
        return 0

    def hwbrj_set_overlap_h2d(self, on):
        self.overlap = on

    def _result(self):
        C, N = self.C, self.N
        p = C.cast(self.libc.malloc(C.sizeof(N.ResultT)), C.POINTER(N.ResultT))  # freed by bench.py like a real result_t
        p.contents.totalresults = 10_240_000
        return p

    def BPRO(self, r, s, nthreads, args):
        return self._result()

    def PRO(self, r, s, nthreads):
        return self._result()


class _StubRelation:
    _h = 1

    @classmethod
    def generate(cls, kind, n, r, q, seed):
        assert kind in (0, 1, 2)
        return cls()

    def free(self):
        pass


@pytest.mark.parametrize("workload", ["c1", "c3", "c5_zipf"])
def test_bench_line_assembles_with_a_stub_library(monkeypatch, capsys, workload):
    """bench.main() end to end on a stub of the C ABI: the ONE JSON line carries every key of the contract"""
    import hwbloomradixjoin_b200 as H
    from hwbloomradixjoin_b200 import _native as N, api, build
    stub = _StubLib()
    monkeypatch.setattr(N, "load", lambda: stub)
    monkeypatch.setattr(build, "build_library", lambda *a, **k: None)
    monkeypatch.setattr(H, "device_count", lambda: 1)
    monkeypatch.setattr(H, "set_quiet", lambda q: None)
    monkeypatch.setattr(H, "DeviceRelation", _StubRelation)
    monkeypatch.setattr(H, "join_device", api.join_device)  # the real wrapper, over the stub library
    monkeypatch.setenv("PATH", "/nonexistent")               # no nvidia-smi: the sampler must cope
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", workload, "--steps", "3", "--warmup", "1",
                                      "--no-cpu-baseline"])
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    assert bench.main() == 0
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["gpu_launches"] == 39 and d["vs_baseline"] is None
    assert abs(d["ms_per_step"] - 8.63) < 1e-3 and abs(d["ms_per_step_without_zero_fill"] - 8.6) < 1e-3
    assert d["value"] == pytest.approx((bench.WORKLOADS[workload][0] + bench.WORKLOADS[workload][1]) / 8.63e-3 / 1e6, rel=1e-3)
    rf = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in rf, key
    assert rf["bound"] == "hbm" and rf["frac"] == pytest.approx(rf["achieved"] / rf["peak"])
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in d["e2e"], key
    assert d["e2e"]["h2d_bytes_per_step"] == 9_216_000_000 and d["e2e"]["d2h_bytes_per_step"] == 136
    assert stub.overlap == 0  # the knob is switched back after the host-buffer leg


def test_secondary_bound_of_the_probe():
    """the wavefront bound of DESIGN.md section 4.4: one L1TEX wavefront per probed key plus one per 16 streamed tuples and
    range pass, at 148 SMs x 1.965 GHz -- 3.96 ms for C1 (two range passes), stated only where a key probes exactly once"""
    import bench
    r, s, q, variant, m, k, B, _ = bench.WORKLOADS["c1"]
    sec = bench.secondary_bound(s, k, variant, 2, 5.2)
    assert sec["wavefronts_per_step"] == s + 2 * (s // 16)
    assert abs(sec["lower_bound_ms"] - 3.961) < 0.01
    assert abs(sec["frac"] - 3.961 / 5.2) < 0.005
    one_pass = bench.secondary_bound(256_000_000, 1, 0, 1, 1.03)
    assert 0.85 < one_pass["frac"] < 0.92                       # C0: 1.03 ms for 256 M keys
    assert bench.secondary_bound(s, 4, 1, 2, 9.3)["achieved"] is None   # BLOCKED k = 4: probes per key depend on the data
    assert bench.secondary_bound(s, 3, 0, 1, 9.3)["achieved"] is None   # BASIC k > 1
