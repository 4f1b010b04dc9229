#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 Bloom-filter radix join (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c0|c3|...]

A "step" is one complete join (Bloom build over R, S pre-filter, radix partitioning, per-partition build+probe)
over one batch of synthetic input with the reference generator's key multiset. `value` is timed with CUDA
events on the library's own stream with the inputs resident in HBM (the reference's timed region: relations
in memory, filter allocated and zeroed beforehand); `e2e` goes through the reference-facing C-ABI call
BPRO()/PRO() with pinned HOST buffers, host->device copies and the result read-back inside the timed region.
`--impl reference` times the UNMODIFIED reference (oracle/_ref/libref.so, compiled from /root/reference/src in
the authoring container) on the box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (r, s, q, bloom variant or None, m, k, B, description)
    "c1": (128_000_000, 1_024_000_000, 0.01, 0, 1 << 30, 1, 512,
           "README canonical: -r 128000000 -s 1024000000 -q 0.01 -b basic -m 1073741824 -k 1"),
    "c0": (16_000_000, 256_000_000, 0.01, 0, 1 << 27, 1, 512,
           "PR1 ref: -r 16000000 -s 256000000 -q 0.01 -m 134217728 -k 1"),
    "c1_blocked": (128_000_000, 1_024_000_000, 0.01, 1, 1 << 30, 4, 256,
                   "canonical inputs, -b blocked -B 256 -k 4"),
    "c1_blocked_k1": (128_000_000, 1_024_000_000, 0.01, 1, 1 << 30, 1, 512, "canonical inputs, -b blocked -B 512 -k 1"),
    "c3": (128_000_000, 128_000_000, 1.0, None, 0, 0, 0, "Workload B plain PRO: -r 128000000 -s 128000000"),
    "c1_quarter": (32_000_000, 256_000_000, 0.01, 0, 1 << 28, 1, 512, "1/4 of the canonical workload (per-GPU sizes of C1 on 8 GPUs when run on 2)"),
    "small": (1_000_000, 8_000_000, 0.01, 0, 1 << 23, 1, 512, "1M x 8M smoke-sized"),
    # q < 0 means: S holds Zipf-distributed foreign keys with exponent -q (mchashjoins -z, create_relation_zipf)
    "c5_zipf": (128_000_000, 1_024_000_000, -1.0, 0, 1 << 30, 1, 512,
                "Zipf-skewed S: -r 128000000 -s 1024000000 -z 1.0 -b basic -m 1073741824 -k 1"),
}


def s_generator(q: float):
    """(kind, parameter) of hwbrj_rel_generate for the probe relation of a workload"""
    return (2, -q) if q < 0 else (1, q)
METRIC = "M input tuples/s ((|R|+|S|)/time)"
UNIT = "Mtuples/s"


def ncu_traffic(kernel_key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/r2_traffic.json, else r1_traffic.json; null when there is no capture for this
    workload)."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        return json.load(open(p)).get(kernel_key)
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0):
        self.rows = []  # (monotonic time, fields)
        self.proc = None
        self.gpu = gpu_index
        self.t0 = None

    def start(self):
        """Spawn the sampler (do this before the warm-up: nvidia-smi needs a few hundred ms before its first line)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [x.strip() for x in line.split(",")]))

    def begin(self, timeout_s: float = 10.0):
        """Call right before the timed region: waits until the sampler delivers, then opens the sampling window."""
        if self.proc is None:
            self.start()
        t_end = time.monotonic() + timeout_s
        while self.proc is not None and not self.rows and time.monotonic() < t_end and self.proc.poll() is None:
            time.sleep(0.01)
        self.t0 = time.monotonic()

    def stop(self) -> dict:
        t1 = time.monotonic()
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.06)  # let the sample that was being taken when the region ended arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0 = self.t0 if self.t0 is not None else 0.0
        window = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.06]
        note = None
        if not window:  # region shorter than one sampling period: take the samples closest to it
            window = [r for (_, r) in self.rows[-3:]]
            note = "timed region shorter than the sampling period: nearest samples used"
        sm, mx, reasons = [], [], set()
        for r in window:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [c for c in sm if c > 0.5 * (max(mx) if mx else 1)] or sm
        out = {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


def secondary_bound(s, k, variant, range_passes, k2_ms):
    """What actually bounds K2 on this part: the L1TEX unit of an SM accepts one wavefront per clock,
    a divergent probe is one wavefront per key, streaming S costs one wavefront per 128-byte line (16 tuples) and range
    pass. Only stated for BASIC k <= 1, where a key probes exactly once."""
    rate = 148 * 1.965  # G wavefronts/s
    if variant != 0 or k > 1:
        return {"bound": "l1tex wavefronts", "unit": "G wavefronts/s", "peak": rate, "achieved": None,
                "note": "not stated for k > 1 / BLOCKED: the number of probes per key depends on the data"}
    wavefronts = s * max(k, 1) + range_passes * (s // 16)
    return {"bound": "l1tex wavefronts (1 per SM per clock = 148 x 1.965 GHz): one per probed key + one per 16 streamed "
                     "tuples and range pass", "unit": "G wavefronts/s", "achieved": wavefronts / (k2_ms * 1e-3) / 1e9,
            "peak": rate, "frac": wavefronts / (k2_ms * 1e-3) / 1e9 / rate, "wavefronts_per_step": wavefronts,
            "lower_bound_ms": wavefronts / rate / 1e6, "probes_per_s_G": s * max(k, 1) / (k2_ms * 1e-3) / 1e9}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def wl_config(name, wl):
    """the `config` object both arms print (identical: the driver compares them)"""
    r, s, q, variant, m, k, B, desc = wl
    return {"workload": desc, "name": name, "r": r, "s": s, "q": q if q >= 0 else None, "zipf": -q if q < 0 else None,
            "bloom": None if variant is None else {"variant": "basic" if variant == 0 else "blocked", "m": m, "k": k, "B": B},
            "l2": f"inputs are {((r + s) * 8) >> 20} MiB per step, larger than the 126 MB L2 (and the host LLC); no flush needed"}


def scaled(wl, scale: int):
    r, s, q, variant, m, k, B, desc = wl
    if scale == 1:
        return wl
    return (r // scale, s // scale, q, variant, max(m // scale, 8) if variant is not None else 0, k, B, desc)


def sample_text(wl, scale: int, use_ref: bool) -> str:
    r, s, q, variant, m, k, B, _ = wl
    head = "the full workload" if scale == 1 else f"1/{scale} of the workload (same |S|/|R|, q, m/|R|, k)"
    return (f"{head}: r={r} s={s} " + (f"q={q} " if q >= 0 else f"zipf={-q} ")
            + (f"{'basic' if variant == 0 else 'blocked'} m={m} k={k} B={B}" if variant is not None else "no filter")
            + ("; unmodified reference BPRO/PRO, its own TOTAL-TIME-USECS" if use_ref else "; oracle port, wall clock"))


def cpu_join_once(oracle, R, S, wl, nthreads: int, use_ref: bool):
    """one join of the reference's CPU implementation; returns (seconds, result dict)"""
    r, s, q, variant, m, k, B, _ = wl
    if use_ref:
        res = oracle.ref_join(R, S, "PRO", nthreads, variant is not None, variant or 0, m or 8, k, B or 512)
        return res["total_usecs"] * 1e-6, res
    t0 = time.perf_counter()
    res = oracle.join(R, S, variant is not None, variant or 0, m or 8, k, B or 512)
    return time.perf_counter() - t0, res


def run_reference_arm(args, wl_name, wl):
    """The reference's own CPU implementation on the box's host cores, same config as our arm. Nothing of this repo's
    product is imported here: the inputs come from the reference's own generator (generator.c through oracle/_ref), the
    join is the unmodified BPRO/PRO."""
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    import oracle
    use_ref = oracle.ref_available()
    nthreads = os.cpu_count() or 1
    swl = scaled(wl, args.ref_scale)
    r, s, q, variant, m, k, B, _ = swl
    if use_ref:  # main.c:410-466 with the CLI's default seeds
        R = oracle.ref_generate(0, r, r, 1.0, 0.0, 12345, nthreads)
        S = oracle.ref_generate(2 if q < 0 else 1, s, r, q if q >= 0 else 1.0, -q if q < 0 else 0.0, 54321, nthreads)
    else:
        R, S = oracle.gen_R(r), (oracle.gen_zipf(s, r, -q) if q < 0 else oracle.gen_S(s, r, q))
    times, res = [], None
    for _ in range(args.steps + args.warmup):
        t, res = cpu_join_once(oracle, R, S, swl, nthreads, use_ref)
        times.append(t)
    times = times[args.warmup:]
    secs = sum(times)
    value = (r + s) * len(times) / secs / 1e6
    sample = sample_text(swl, args.ref_scale, use_ref)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / len(times) * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": wl_config(wl_name, wl), "sample": sample,
            "results": {"matches": int(res["matches"]), "filtered": int(res["filtered"])},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads if use_ref else 1,
                             "kind": "reference" if use_ref else "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-scale", type=int, default=1, help="the CPU arm runs on 1/ref-scale of the workload (1 = the same config)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer leg (default: min(steps,5))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference_arm(args, args.workload, wl)

    watchdog = threading.Timer(float(os.environ.get("HWBRJ_BENCH_WATCHDOG_S", "900")), lambda: os._exit(3))
    watchdog.daemon = True  # a stuck run must not hold the GPU box
    watchdog.start()
    rank, world, local = dist_env()
    if world > 1 or args.gpus > 1:
        from hwbloomradixjoin_b200 import bench_dist
        return bench_dist.main(args, wl, METRIC, UNIT, measured_peak, ClockSampler)

    import hwbloomradixjoin_b200 as H
    from hwbloomradixjoin_b200 import build
    build.build_library()
    if H.device_count() < 1:
        print(json.dumps({"error": "no CUDA device: this benchmark has no CPU fallback"}))
        return 1
    H.set_quiet(True)
    r, s, q, variant, m, k, B, desc = wl
    bloom = H.BloomFilterArgs(variant, m, k, B) if variant is not None else None

    # ---- inputs (reference generator's multiset, generated on the device; synthetic) ----
    dR = H.DeviceRelation.generate(0, r, r, 1.0, 1)
    dS = H.DeviceRelation.generate(s_generator(q)[0], s, r, s_generator(q)[1], 2)

    # ---- device-resident leg: `value` ----
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        res = H.join_device(dR, dS, bloom)
    sampler.begin()
    per_step, stats = [], []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        res = H.join_device(dR, dS, bloom)
        # CUDA-event time of the whole join on the library stream, filter/histogram zero-fill included
        per_step.append(res.stats["ms_total"] + res.stats["ms_memset"])
        stats.append(res.stats)
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    ms_total = sum(per_step)
    ms_per_step = ms_total / args.steps
    value = (r + s) / (ms_per_step * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (S-side probe + compaction) ----
    peak, peak_src = measured_peak()
    F = res.filtered if bloom is not None else s
    phases = {p: statistics.mean(st[p] for st in stats) for p in
              ("ms_build", "ms_part_r", "ms_probe", "ms_part_s", "ms_join", "ms_memset")}
    if bloom is not None:
        dom_name = "k_probe_compact (K2: Bloom probe of S + survivor compaction)"
        dom_bytes = 8 * s + 8 * F + m // 8          # S read once, survivors written once, filter read once
        dom_ms = phases["ms_probe"]
    else:
        dom_name = "k_build_hist + k_scatter (K3/K4: histogram and radix scatter passes of S)"
        dom_bytes = 8 * s + 16 * s * (2 if stats[-1]["radix_bits"] > 7 else 1)  # histogram read + read/write per scatter pass
        dom_ms = phases["ms_part_s"]  # histogram + offsets + both scatter passes of S (hwbrj_stats_t.phase_split == 1)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    b_alg = (24 * r + 8 * s + 16 * F + 2 * (m // 8)) if bloom is not None else (24 * r + 24 * s)
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_nominal_8000_GBps": achieved / 8000.0, "traffic": ncu_traffic(f"{args.workload}:{'k_probe_compact' if bloom is not None else 'k_scatter'}"),
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch_set": dom_bytes, "ms_per_launch_set": dom_ms,
                "algorithmic_bytes_per_launch": dom_bytes // max(stats[-1]["range_passes"] if bloom is not None else 1, 1),
                "note": "achieved = algorithmic bytes of the kernel's launches in one step / their summed CUDA-event time; traffic = ncu DRAM bytes of ONE launch",
                "launches_per_step": stats[-1]["range_passes"] if bloom is not None else 1,
                # what actually bounds K2 on this part: one divergent L1TEX access per SM per clock
                "secondary": None if bloom is None else secondary_bound(s, k, variant, stats[-1]["range_passes"], dom_ms),
                "whole_join": {"algorithmic_bytes": b_alg, "achieved": b_alg / (ms_per_step * 1e-3) / 1e9,
                               "frac": b_alg / (ms_per_step * 1e-3) / 1e9 / peak}}

    # ---- host-buffer leg through the C ABI: `e2e` ----
    from hwbloomradixjoin_b200 import _native as N
    L = N.load()
    e2e_steps = args.e2e_steps or min(args.steps, 5)
    hR = L.hwbrj_host_alloc(r * 8)
    hS = L.hwbrj_host_alloc(s * 8)
    L.hwbrj_rel_download(dR._h, hR)
    L.hwbrj_rel_download(dS._h, hS)
    relR = N.RelationT(hR, r)
    relS = N.RelationT(hS, s)
    cargs = bloom.to_c() if bloom is not None else None
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]

    def host_call():
        if bloom is not None:
            p = L.BPRO(C.byref(relR), C.byref(relS), 1, C.byref(cargs))
        else:
            p = L.PRO(C.byref(relR), C.byref(relS), 1)
        tot = p.contents.totalresults
        libc.free(C.cast(p, C.c_void_p))
        return tot
    L.hwbrj_set_overlap_h2d(1)  # S is uploaded in chunks and each chunk is probed as it lands (a public knob of the ABI)
    host_call()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        tot = host_call()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    L.hwbrj_set_overlap_h2d(0)
    assert tot == res.totalresults
    st = N.StatsT()
    L.hwbrj_last_stats(C.byref(st))
    e2e = {"value": (r + s) / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(st.h2d_bytes),
           "d2h_bytes_per_step": int(st.d2h_bytes), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
           "ms_h2d": st.ms_h2d, "api": "BPRO(relation_t*,relation_t*,int,bloom_filter_args_t*)" if bloom else "PRO(...)"}
    # ---- CPU baseline beside it: the unmodified reference on the host cores, on the SAME arrays the host-buffer leg
    # just joined (rank 0, N=1 only; one join of the full workload unless --ref-scale says otherwise) ----
    cpu = None
    if not args.no_cpu_baseline:
        try:
            import oracle
            use_ref = oracle.ref_available()
            nthreads = os.cpu_count() or 1
            sc = args.ref_scale
            swl = scaled(wl, sc)
            hRa = np.ctypeslib.as_array(C.cast(hR, C.POINTER(C.c_int64)), shape=(r,)).view(oracle.TUPLE)
            hSa = np.ctypeslib.as_array(C.cast(hS, C.POINTER(C.c_int64)), shape=(s,)).view(oracle.TUPLE)
            if sc == 1:
                Rc, Sc = hRa, hSa
            else:  # a smaller sample needs its own key multiset (the restated reference generator)
                Rc = oracle.gen_R(swl[0])
                Sc = oracle.gen_zipf(swl[1], swl[0], -q) if q < 0 else oracle.gen_S(swl[1], swl[0], q)
            secs, cres = cpu_join_once(oracle, Rc, Sc, swl, nthreads, use_ref)
            cpu = {"value": (swl[0] + swl[1]) / secs / 1e6, "unit": UNIT, "cores": nthreads if use_ref else 1,
                   "kind": "reference" if use_ref else "port", "sample": sample_text(swl, sc, use_ref), "seconds": secs,
                   "same_arrays_as_gpu": sc == 1}
            if sc == 1:  # same arrays, so the scalars must agree
                cpu["matches_equal"] = int(cres["matches"]) == int(res.totalresults)
                cpu["filtered_equal"] = bloom is None or int(cres["filtered"]) == int(res.filtered)
        except Exception as exc:  # the baseline is a reported side number: never lose the bench line over it
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": f"failed: {exc}"}
    L.hwbrj_host_free(hR)
    L.hwbrj_host_free(hS)
    dR.free()
    dS.free()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": wl_config(args.workload, wl),
            "details": {"radix_bits": stats[-1]["radix_bits"], "range_passes": stats[-1]["range_passes"],
                        "timed_region": "CUDA events on the library stream around every launch of the join, filter/histogram zero-fill included"},
            "results": {"matches": res.totalresults, "filtered": res.filtered, "checksum_pair": res.checksum_pair},
            "phases_ms": phases, "wall_ms_per_step": wall / args.steps * 1e3,
            "ms_per_step_without_zero_fill": ms_per_step - phases["ms_memset"],  # the reference's own timed region (:1583)
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(sum(st_["kernel_launches"] for st_ in stats)), "clocks": clocks}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
