"""The reference's own Bloom FPR measurement (`./unittests 2 817263 1024000000 128000000 1073741824 12`,
measurements/data/bloom_filter_fpr_orig.txt) replayed on the GPU at full size: same glibc-rand() samples, same
rand() filter seed, blocked (B = 512) and basic, k = 1..12 -- every empirical FPR must print identically (3 decimals)."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TABLE = json.load(open(os.path.join(HERE, "golden", "fpr_table.json")))


def test_fpr_table_matches_reference_output(Hgpu, oracle_mod):
    assert TABLE["command"] == "./unittests 2 817263 1024000000 128000000 1073741824 12"
    seed, n_samples, n_insertions, m = 817263, 1_024_000_000, 128_000_000, 1 << 30
    R, S, filter_seed = oracle_mod.fpr_samples(seed, n_samples, n_insertions)
    dR, dS = Hgpu.DeviceRelation.upload(R), Hgpu.DeviceRelation.upload(S)
    del R, S
    try:
        assert len(TABLE["rows"]) == 24
        for row in TABLE["rows"]:
            args = Hgpu.BloomFilterArgs(0 if row["variant"] == "basic" else 1, m, row["k"], TABLE["B"])
            pos = Hgpu.fpr_count(dR, dS, args, filter_seed)
            # unit_tests.c:222-225: selectivity 0 -> neg = n_samples, tp = 0, fpr = pos / neg, printed "%.3f%%"
            assert "%.3f" % (pos / n_samples * 100) == row["fpr_emp"], (row, pos)
    finally:
        dR.free()
        dS.free()
