"""Profiling target: one warm-up join and one measured join of a BASELINE workload (default C1), nothing else."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hwbloomradixjoin_b200 as H
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS

name = sys.argv[1] if len(sys.argv) > 1 else "c1"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
r, s, q, variant, m, k, B, desc = WORKLOADS[name]
H.set_quiet(True)
dR = H.DeviceRelation.generate(0, r, r, 1.0, 1)
dS = H.DeviceRelation.generate(2 if q < 0 else 1, s, r, -q if q < 0 else q, 2)  # q < 0: Zipf exponent -q
bloom = H.BloomFilterArgs(variant, m, k, B) if variant is not None else None
for i in range(reps):
    res = H.join_device(dR, dS, bloom)
st = res.stats
print(f"{name}: matches={res.totalresults} filtered={res.filtered} total={st['ms_total']:.3f} ms "
      f"[build {st['ms_build']:.3f} partR {st['ms_part_r']:.3f} probe {st['ms_probe']:.3f} partS {st['ms_part_s']:.3f} "
      f"join {st['ms_join']:.3f}] launches={st['kernel_launches']} bits={st['radix_bits']} ranges={st['range_passes']}")
