"""Development helper: time a BASELINE workload for several numbers of filter range passes (results must not change)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hwbloomradixjoin_b200 as H
from bench import WORKLOADS

name = sys.argv[1] if len(sys.argv) > 1 else "c1"
passes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2, 3, 4]
r, s, q, variant, m, k, B, desc = WORKLOADS[name]
H.set_quiet(True)
dR = H.DeviceRelation.generate(0, r, r, 1.0, 1)
dS = H.DeviceRelation.generate(2 if q < 0 else 1, s, r, -q if q < 0 else q, 2)  # q < 0: Zipf exponent -q
bloom = H.BloomFilterArgs(variant, m, k, B)
ref = None
for nr in passes:
    H.set_range_passes(nr)
    for i in range(3):
        res = H.join_device(dR, dS, bloom)
    st = res.stats
    key = (res.totalresults, res.filtered, res.checksum_pair)
    ref = ref or key
    print(f"{name} passes={st['range_passes']}: total={st['ms_total']:.3f} ms probe {st['ms_probe']:.3f} build {st['ms_build']:.3f} "
          f"same_result={key == ref}", flush=True)
