"""Opt-in parity checks of kernels that are still behind an environment knob (not part of the default GPU suite):
    HWBRJ_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -q
Each case runs in its own process because the library reads its knobs once, at the first call."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("HWBRJ_TEST_EXPERIMENTAL") != "1",
                                 reason="experimental kernels: set HWBRJ_TEST_EXPERIMENTAL=1")]
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
    """k_probe_staged (HWBRJ_PROBE_STAGED=1): probes 2..k on compacted candidates; BLOCKED and BASIC, ranged and not"""
    assert _run({"HWBRJ_PROBE_STAGED": "1"}) == []


def test_staged_probe_with_forced_range_passes():
    assert _run({"HWBRJ_PROBE_STAGED": "1", "HWBRJ_RANGE_PASSES": "4"}) == []


@pytest.mark.parametrize("lib", sorted(f for f in os.listdir(os.path.join(ROOT, "build", "variants"))
                                       if f.endswith(".so")) if os.path.isdir(os.path.join(ROOT, "build", "variants")) else [])
def test_tuning_variant_matches_oracle(lib):
    """every build/variants/lib_*.so that is not a timing-only ablation must still be bit-exact"""
    if "no_" in lib or "only" in lib:
        pytest.skip("ablation build: results are wrong by construction")
    assert _run({"HWBRJ_LIB": os.path.join(ROOT, "build", "variants", lib)}) == []
