#!/usr/bin/env python
"""Sweep driver in the style of the reference's measurements/run.py (run.py:70-156, 272-373): runs
build/mchashjoins_gpu over a grid of (r, s, q, filter, m, k, B), parses its stdout with the same regular expressions
and writes one CSV row per run (the reference pickles a DataFrame with the same column names).

  python tools/run_sweep.py --grid basic_vs_blocked --out gpurun_out/sweep.csv
"""
import argparse
import csv
import itertools
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse_result(res: str) -> dict:  # same expressions as measurements/run.py:100-129
    s_size = int(re.search(r"relation S with size = [\d.]+ MiB, #tuples = (\d+) : OK", res).group(1))
    filtered = re.search(r"S-tuples after filter: (\d+)\n", res)
    filtered = int(filtered.group(1)) if filtered else None
    runtime_cycles, build_cycles, part_cycles = re.search(
        r"RUNTIME TOTAL, BUILD, PART \(cycles\):\W+(\d+)\W+(\d+)\W+(\d+)", res).groups()
    time_usecs, out_tuples, nsec = re.search(
        r"TOTAL-TIME-USECS, TOTAL-TUPLES, NSEC-PER-TUPLE:\W+([\d.]+)\W+(\d+)\W+([\d.]+)", res).groups()
    part_us, probe_us, join_us = re.search(
        r"PARTITION-TIME-USECS, PROBE-TIME-USECS, JOIN-TIME-USECS:\W+([\d.]+)\W+([\d.]+)\W+([\d.]+)", res).groups()
    return {"filtered": filtered, "filtered-pct": filtered / s_size * 100 if filtered is not None else None,
            "runtime-cycles": int(runtime_cycles), "build-cycles": int(build_cycles), "part-cycles": int(part_cycles),
            "time-usecs": float(time_usecs), "out-tuples": int(out_tuples), "nsec-per-tuple": float(nsec),
            "partition-usecs": float(part_us), "probe-usecs": float(probe_us), "join-usecs": float(join_us)}


GRIDS = {
    # (r, s) x q x filter x k, filter size m = 8 bits per R tuple rounded to a power of two (config.py of the reference)
    "basic_vs_blocked": dict(sizes=[(250_000, 2_000_000), (16_000_000, 128_000_000), (128_000_000, 1_024_000_000)],
                             q=[0.01], filters=["no", "basic", "blocked"], k=[1, 2, 3, 4, 5, 6, 7], B=[512]),
    "test_parameters": dict(sizes=[(128_000_000, 1_024_000_000)], q=[0.001, 0.01, 0.1], filters=["basic", "blocked"],
                            k=[1, 2, 3, 4, 5, 6], B=[256, 512, 1024]),
    "smoke": dict(sizes=[(250_000, 2_000_000)], q=[0.01], filters=["no", "basic", "blocked"], k=[1, 3], B=[512]),
    # measurements/run.py:205-269 never_single_pass(): 1 vs 2 partitioning passes, with and without a filter
    # (NUM_PASSES is a compile-time constant there; here the HWBRJ_NUM_PASSES knob of the same binary)
    "never_single_pass": dict(sizes=[(32_000_000, 1_024_000_000)], q=[0.01, 0.1], filters=["no", "basic", "blocked"],
                              k=[1, 2, 3], B=[512], m=1 << 28, passes=[1, 2]),
    # measurements/run.py:272-330 best_bloom_filter_type(): filter sizes around the L2 capacity, k = 1..4
    "best_bloom_filter_type": dict(sizes=[(128_000_000, 1_024_000_000)], q=[0.01], filters=["basic", "blocked"],
                                   k=[1, 2, 3, 4], B=[256, 512], m_list=[1 << 28, 1 << 30, 1 << 31]),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="smoke", choices=sorted(GRIDS))
    ap.add_argument("--out", default="gpurun_out/sweep.csv")
    ap.add_argument("--algo", default="PRO")
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--gpus", type=int, default=1)
    args = ap.parse_args()
    from hwbloomradixjoin_b200 import build
    build.build_library()
    exe = build.build_driver()
    g = GRIDS[args.grid]
    rows = []
    for (r, s), q, filt, passes in itertools.product(g["sizes"], g["q"], g["filters"], g.get("passes", [0])):
        m_auto = 1
        while m_auto < 8 * r:
            m_auto <<= 1
        m_list = g.get("m_list", [g.get("m", m_auto)]) if filt != "no" else [0]
        for m in m_list:
            for k, B in (itertools.product(g["k"], g["B"] if filt == "blocked" else g["B"][:1]) if filt != "no" else [(0, 0)]):
                for _ in range(args.repeat):
                    cmd = [exe, "-a", args.algo, "-n", "1", "-r", str(r), "-s", str(s), "-q", str(q), "-b", filt]
                    if filt != "no":
                        cmd += ["-m", str(m), "-k", str(k), "-B", str(B)]
                    if args.gpus > 1:
                        cmd += ["--gpus", str(args.gpus)]
                    env = dict(os.environ)
                    if passes:
                        env["HWBRJ_NUM_PASSES"] = str(passes)
                    p = subprocess.run(cmd, capture_output=True, text=True, env=env)
                    if p.returncode != 0:
                        print("failed:", " ".join(cmd), p.stdout[-300:], file=sys.stderr)
                        continue
                    bits = re.search(r"radix bits = (\d+)", p.stdout)
                    row = {"r-size": r, "s-size": s, "s-sel": q, "bloom-filter": filt, "bloom-hashes": k, "bloom-size": m,
                           "bloom-block-size": B, "num-passes": passes or "auto", "gpus": args.gpus,
                           "radix-bits": int(bits.group(1)) if bits else None, **parse_result(p.stdout)}
                    rows.append(row)
                    print(row, flush=True)
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
        w.writeheader()
        w.writerows(rows)
    print(f"{len(rows)} rows -> {args.out}")


if __name__ == "__main__":
    main()
