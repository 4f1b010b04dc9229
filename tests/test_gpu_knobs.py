"""Parity of the pipeline variants that are selected by environment knobs (staged probe for k >= 2, forced filter range
passes, forced hash / radix partitioning, 1 or 2 scatter passes), of the device Zipf generator and of random filter
points. Each case runs in its own process because the library reads its knobs once, at the first call. All of these ran
on hardware in round 2 (they were opt-in in round 1)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, sys
sys.path.insert(0, %(root)r)
import numpy as np
import oracle
import hwbloomradixjoin_b200 as H
H.set_quiet(True)
R = oracle.gen_R(300001); S = oracle.gen_S(2500003, 300001, 0.02)
bad = []
for variant, m, k, B in %(cases)r:
    args = H.BloomFilterArgs(variant, m, k, B)
    r = H.BPRO(R, S, 4, args); o = oracle.join(R, S, True, variant, m, k, B)
    got = (r.totalresults, r.filtered, r.checksum_pair, r.checksum_key)
    want = (o["matches"], o["filtered"], o["checksum_pair"], o["checksum_key"])
    if got != want:
        bad.append([variant, m, k, B, got, want])
    bm = oracle.bloom_build(R, variant, m, k, B)
    n, surv = H.bloom_probe(bm, S, args, want_survivors=True)
    no, so = oracle.bloom_filter(bm, S, variant, m, k, B, want_survivors=True)
    if n != no or not (np.sort(surv, order=["key", "payload"]) == np.sort(so, order=["key", "payload"])).all():
        bad.append([variant, m, k, B, "survivor multiset"])
print(json.dumps(bad))
"""

CASES = [(1, 1 << 22, 4, 256), (1, 1 << 22, 2, 64), (1, 1 << 22, 8, 512), (1, 1 << 22, 3, 1 << 22), (0, 1 << 22, 2, 512),
         (0, 1 << 22, 5, 512), (1, 1 << 30, 4, 256), (1, 1 << 28, 12, 512)]


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    p = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "cases": CASES}], capture_output=True, text=True,
                       timeout=900, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


def test_staged_probe_matches_oracle():
    """k_probe_staged (the default for k >= 2): probes 2..k on compacted candidates; BLOCKED and BASIC, ranged and not"""
    assert _run({"HWBRJ_PROBE_STAGED": "1"}) == []


def test_unstaged_probe_matches_oracle():
    """HWBRJ_PROBE_STAGED=0: all k probes of a key inside k_probe_compact"""
    assert _run({"HWBRJ_PROBE_STAGED": "0"}) == []


def test_staged_probe_with_forced_range_passes():
    assert _run({"HWBRJ_PROBE_STAGED": "1", "HWBRJ_RANGE_PASSES": "4"}) == []


@pytest.mark.parametrize("env", [
    {"HWBRJ_HASH_PARTITION": "2"},                                  # slice build in shared memory wherever the slices fit
    {"HWBRJ_HASH_PARTITION": "0"},                                  # never: global atomics + radix partitions
    {"HWBRJ_HASH_PARTITION": "2", "HWBRJ_RADIX_BITS": "9"},
    {"HWBRJ_RADIX_BITS": "6", "HWBRJ_NUM_PASSES": "1"},             # the reference's NUM_RADIX_BITS / NUM_PASSES knobs
    {"HWBRJ_RADIX_BITS": "6", "HWBRJ_NUM_PASSES": "2"},
    {"HWBRJ_RADIX_BITS": "12", "HWBRJ_NUM_PASSES": "1"},            # one pass caps the fan-out at 2^7: multi-round tables
    {"HWBRJ_HASH_PARTITION": "2", "HWBRJ_RANGE_PASSES": "2", "HWBRJ_PROBE_CTAS": "2"},
    {"HWBRJ_PROBE_ADAPTIVE": "0"},                                  # no sample of S: ld.global.cg probes and the dense-survivor shape whatever S looks like
], ids=lambda e: ",".join(f"{k[6:]}={v}" for k, v in e.items()))
def test_pipeline_knobs_match_oracle(env):
    assert _run(env) == []


ZIPF_SCRIPT = r"""
import json, math, sys
sys.path.insert(0, %(root)r)
import numpy as np
import hwbloomradixjoin_b200 as H
H.set_quiet(True)
n, r, theta = 3000001, 250000, 1.0
bad = []
dS = H.DeviceRelation.generate(2, n, r, theta, 7)
S = dS.download()
if not ((S["key"] >= 1).all() and (S["key"] <= r).all()): bad.append("keys outside the alphabet")
if not (S["payload"] == np.arange(n, dtype=np.int64).astype(np.int32)).all(): bad.append("payload != position")
cnt = np.sort(np.bincount(S["key"], minlength=r + 1))[::-1].astype(np.float64)
Hn = sum(1.0 / (j ** theta) for j in range(1, r + 1))
for rank in range(1, 6):  # the five most frequent keys follow n / (rank^theta * H)
    exp = n / (rank ** theta * Hn)
    if abs(cnt[rank - 1] - exp) > 6 * math.sqrt(exp): bad.append(["rank", rank, cnt[rank - 1], exp])
# a shard generated on its own equals the slice of the whole relation
from hwbloomradixjoin_b200 import _native as N
import ctypes as C
h = N.load().hwbrj_rel_generate_shard(2, n, r, theta, 7, 1000003, 500000)
part = H.DeviceRelation(h).download()
if not (part == S[1000003:1500003]).all(): bad.append("shard differs from the slice")
# every Zipf key has a partner in R = 1..r: the join matches and the filter passes every S tuple
dR = H.DeviceRelation.generate(0, r, r, 1.0, 1)
res = H.join_device(dR, dS, H.BloomFilterArgs(0, 1 << 22, 1, 512))
if (res.totalresults, res.filtered) != (n, n): bad.append(["join", res.totalresults, res.filtered])
if res.checksum_key != int(S["key"].astype(np.uint64).sum()) %% (1 << 64): bad.append("key checksum")
print(json.dumps(bad))
"""


@pytest.mark.parametrize("adaptive", ["1", "0"])
def test_device_zipf_generator(adaptive):
    """kind 2 of hwbrj_rel_generate: alphabet, cumulated-density table and binary search of genzipf.c on the device; the
    join of the skewed relation runs K2 with L1-allocating probe loads (k_probe_sample sets the bits: repeated keys, dense survivors -> the 8-keys-per-lane shape) and, with
    HWBRJ_PROBE_ADAPTIVE=0, with the loads of the uniform case -- same scalars"""
    p = subprocess.run([sys.executable, "-c", ZIPF_SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, HWBRJ_PROBE_ADAPTIVE=adaptive))
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert json.loads(p.stdout.strip().splitlines()[-1]) == []


RANDOM_SCRIPT = r"""
import json, sys
sys.path.insert(0, %(root)r)
import numpy as np
import oracle
import hwbloomradixjoin_b200 as H
H.set_quiet(True)
rng = np.random.default_rng(77)
R = np.zeros(40_000, dtype=H.TUPLE); R["key"] = rng.integers(-2**31, 2**31, R.shape[0], dtype=np.int64).astype(np.int32)
R["payload"] = np.arange(R.shape[0])
S = np.zeros(300_000, dtype=H.TUPLE); S["key"] = rng.integers(-2**31, 2**31, S.shape[0], dtype=np.int64).astype(np.int32)
S["key"][:20000] = R["key"][rng.integers(0, R.shape[0], 20000)]; S["payload"] = np.arange(S.shape[0])
bad = []
for i in range(64):
    variant = int(rng.integers(0, 2)); log2m = int(rng.integers(10, 23)); k = int(rng.integers(0, 13))
    B = 1 << int(rng.integers(3, log2m + 1)); m = 1 << log2m
    args = H.BloomFilterArgs(variant, m, k, B)
    if not (H.bloom_build(R, args) == oracle.bloom_build(R, variant, m, k, B)).all(): bad.append(["bitmap", variant, m, k, B])
    r = H.BPRO(R, S, 2, args); o = oracle.join(R, S, True, variant, m, k, B)
    if (r.totalresults, r.filtered, r.checksum_pair) != (o["matches"], o["filtered"], o["checksum_pair"]):
        bad.append(["join", variant, m, k, B])
print(json.dumps(bad))
"""


def test_random_filter_points_match_oracle():
    """64 seeded random (variant, m = 2^10..2^22, k = 0..12, B = 8..m) points on full-range int32 keys"""
    p = subprocess.run([sys.executable, "-c", RANDOM_SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert json.loads(p.stdout.strip().splitlines()[-1]) == []
