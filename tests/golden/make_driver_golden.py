#!/usr/bin/env python
"""Golden outputs of the UNMODIFIED reference binary (oracle/_ref/mchashjoins, built by oracle/Makefile from
/root/reference/src) for the generator variants whose arrays depend on glibc rand() and the -x/-y seeds:
--non-unique and --full-range (main.c:421-452, generator.c:531-651) and -z (genzipf.c). The GPU host driver
(host/mchashjoins_gpu.c) restates those serial generators and must print the same `Results` and `S-tuples after filter`
for the same command line (tests/test_gpu_driver.py).

    python tests/golden/make_driver_golden.py      # needs oracle/_ref (i.e. /root/reference): run in the authoring container
"""
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
EXE = os.path.join(ROOT, "oracle", "_ref", "mchashjoins")

CASES = [
    "-a PRO -n 2 -r 200000 -s 1000000 -q 0.1 -b basic -m 2097152 -k 1 --non-unique",
    "-a PRO -n 2 -r 200000 -s 1000000 -q 0.1 -b basic -m 2097152 -k 1 --full-range",
    "-a PRO -n 4 -r 200000 -s 1000000 -q 0.5 --non-unique",
    "-a PRO -n 4 -r 300000 -s 1500000 -q 0.25 -b blocked -m 4194304 -k 3 -B 512 --full-range -x 7 -y 9",
    "-a RJ -n 1 -r 100000 -s 700000 -q 0.9 -b basic -m 1048576 -k 2 --non-unique -x 99 -y 3",
    "-a PRO -n 2 -r 100000 -s 500000 -z 1.0 -b basic -m 1048576 -k 1",
    "-a PRO -n 2 -r 150000 -s 800000 -z 0.5 -y 17",
]


def main():
    out = []
    for c in CASES:
        p = subprocess.run([EXE] + c.split(), capture_output=True, text=True, cwd="/tmp", timeout=600)
        assert p.returncode == 0, p.stdout + p.stderr
        res = int(re.search(r"Results = (\d+)", p.stdout).group(1))
        f = re.search(r"S-tuples after filter: (-?\d+)", p.stdout)
        out.append({"args": c, "results": res, "filtered": int(f.group(1)) if f else None})
        print(out[-1])
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "driver_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
