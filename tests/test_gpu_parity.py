"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle on the same seeded
inputs -- bit-exact for every scalar (match count, filtered count, four checksums), byte-exact for the filter
bitmap, multiset-exact for survivors and partitions -- plus the reference's golden values at all published
sizes (device generator, multiset-equivalent inputs) and size-independent properties at the full C1 size."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_results.json")))["configs"]
HASH_KAT = json.load(open(os.path.join(HERE, "golden", "hash_kat.json")))
BITMAP_KAT = json.load(open(os.path.join(HERE, "golden", "bitmap_kat.json")))

_cache = {}


def inputs(o, r, s, q, seed=0):
    key = (r, s, q, seed)
    if key not in _cache:
        _cache[key] = (o.gen_R(r, nthreads=4, seed=12345 + seed), o.gen_S(s, r, q, nthreads=4, seed=54321 + seed))
    return _cache[key]


def scalars(res):
    return (res.totalresults, res.filtered, res.checksum_pair, res.checksum_rpay, res.checksum_spay, res.checksum_key)


def oscalars(o):
    return (o["matches"], o["filtered"], o["checksum_pair"], o["checksum_rpay"], o["checksum_spay"], o["checksum_key"])


def sort_tuples(a):
    return np.sort(a, order=["key", "payload"])


# ---- K0 ------------------------------------------------------------------------------------------------------
def test_hashes_bit_exact(Hgpu, oracle_mod):
    rng = np.random.default_rng(3)
    keys = np.concatenate([rng.integers(-2**31, 2**31, 300_000, dtype=np.int64).astype(np.int32),
                           np.array([0, 1, -1, 2**31 - 1, -2**31, 0x80, 0x8080, -128], dtype=np.int32)])
    for seed in (42, 0, 0xDEADBEEF):
        for which in range(10):
            assert (Hgpu.hash_many(which, seed, keys) == oracle_mod.hash_many(which, seed, keys)).all(), (which, seed)
    for row in HASH_KAT["rows"]:
        k = np.array([row["key"]], dtype=np.int32)
        assert [int(Hgpu.hash_many(w, row["seed"], k)[0]) for w in range(10)] == row["hashes"]
    assert Hgpu.hash_many(2, 42, np.zeros(0, np.int32)).shape == (0,)
    with pytest.raises(ValueError):
        Hgpu.hash_many(10, 42, keys[:4])


# ---- K1 / K2 ---------------------------------------------------------------------------------------------------
FILTER_CASES = [(0, 1 << 21, 1, 512), (0, 1 << 21, 2, 512), (0, 1 << 21, 8, 512), (1, 1 << 21, 1, 512),
                (1, 1 << 21, 3, 64), (1, 1 << 21, 4, 256), (1, 1 << 21, 8, 1024), (1, 1 << 21, 2, 8),
                (1, 1 << 21, 5, 1 << 21), (0, 1 << 10, 3, 512), (0, 1 << 21, 0, 512), (0, 8, 1, 512)]


@pytest.mark.parametrize("case", FILTER_CASES, ids=str)
def test_filter_bitmap_and_survivors(Hgpu, oracle_mod, case):
    variant, m, k, B = case
    R, S = inputs(oracle_mod, 250_000, 2_000_000, 0.01)
    args = Hgpu.BloomFilterArgs(variant, m, k, B)
    gpu_bm = Hgpu.bloom_build(R, args)
    ora_bm = oracle_mod.bloom_build(R, variant, m, k, B)
    assert gpu_bm.tobytes() == ora_bm.tobytes()
    n, surv = Hgpu.bloom_probe(ora_bm, S, args, want_survivors=True)
    no, so = oracle_mod.bloom_filter(ora_bm, S, variant, m, k, B, want_survivors=True)
    assert n == no
    assert (sort_tuples(surv) == sort_tuples(so)).all()


def test_filter_range_passes_do_not_change_the_bitmap(Hgpu, oracle_mod):
    R, S = inputs(oracle_mod, 250_000, 2_000_000, 0.01)
    for variant, k, B in [(0, 1, 512), (1, 3, 256)]:
        args = Hgpu.BloomFilterArgs(variant, 1 << 21, k, B)
        ref_bm = oracle_mod.bloom_build(R, variant, 1 << 21, k, B)
        for passes in (1, 2, 8):
            Hgpu.set_range_passes(passes)
            try:
                assert Hgpu.bloom_build(R, args).tobytes() == ref_bm.tobytes()
                assert Hgpu.bloom_probe(ref_bm, S, args) == oracle_mod.bloom_filter(ref_bm, S, variant, 1 << 21, k, B)
            finally:
                Hgpu.set_range_passes(0)


def test_bitmap_kat_from_reference(Hgpu):
    R = np.zeros(64, dtype=Hgpu.TUPLE)
    R["key"] = np.arange(1, 65)
    S = np.zeros(1000, dtype=Hgpu.TUPLE)
    S["key"] = np.arange(65, 1065)
    for row in BITMAP_KAT:
        args = Hgpu.BloomFilterArgs(row["variant"], row["m"], row["k"], row["B"])
        bm = Hgpu.bloom_build(R, args)
        assert bm.tobytes().hex() == row["bitmap_hex"], row
        assert Hgpu.bloom_probe(bm, S, args) == row["pass"]


@pytest.mark.parametrize("variant,lgm,k,B", [(0, 31, 1, 512), (0, 32, 1, 512), (0, 32, 2, 512), (1, 32, 3, 512)])
def test_largest_filters(Hgpu, oracle_mod, variant, lgm, k, B):
    """A.5: the reference's uint32 size arithmetic allows m up to 2^32 (512 MiB, 8 filter range passes here)"""
    R, S = inputs(oracle_mod, 250_000, 2_000_000, 0.01)
    m = 1 << lgm
    res = Hgpu.BPRO(R, S, 1, Hgpu.BloomFilterArgs(variant, m, k, B))
    assert scalars(res) == oscalars(oracle_mod.join(R, S, True, variant, m, k, B))
    if k == 1 and variant == 0:
        assert res.stats["range_passes"] == (m // 8) // (64 << 20)


def test_filter_full_range_keys(Hgpu, oracle_mod):
    rng = np.random.default_rng(11)
    R = np.zeros(100_001, dtype=Hgpu.TUPLE)
    R["key"] = rng.integers(-2**31, 2**31, R.shape[0], dtype=np.int64).astype(np.int32)
    S = np.zeros(400_003, dtype=Hgpu.TUPLE)
    S["key"] = rng.integers(-2**31, 2**31, S.shape[0], dtype=np.int64).astype(np.int32)
    S["payload"] = np.arange(S.shape[0])
    for variant, m, k, B in [(0, 1 << 20, 2, 512), (1, 1 << 20, 3, 128)]:
        args = Hgpu.BloomFilterArgs(variant, m, k, B)
        bm = oracle_mod.bloom_build(R, variant, m, k, B)
        assert Hgpu.bloom_build(R, args).tobytes() == bm.tobytes()
        assert Hgpu.bloom_probe(bm, S, args) == oracle_mod.bloom_filter(bm, S, variant, m, k, B)


# ---- K3 / K4 ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bits", [0, 1, 5, 7, 8, 10, 13, 14])
def test_radix_partition(Hgpu, oracle_mod, bits):
    _, S = inputs(oracle_mod, 250_000, 2_000_000, 0.01)
    S = S[:1_234_567]
    out, off = Hgpu.radix_partition(S, bits)
    mask = (1 << bits) - 1
    cnt = np.bincount(S["key"].astype(np.uint32) & mask, minlength=1 << bits)
    assert off[0] == 0 and off[-1] == S.shape[0]
    assert (np.diff(off.astype(np.int64)) == cnt).all()
    pid = (out["key"].astype(np.uint32) & mask).astype(np.int64)
    assert (np.diff(pid) >= 0).all()  # grouped in increasing partition order == HASH_BIT_MODULO clusters
    assert (sort_tuples(out) == sort_tuples(S)).all()


def test_radix_partition_skewed_and_tiny(Hgpu):
    a = np.zeros(300_000, dtype=Hgpu.TUPLE)
    a["key"] = 12345  # everything in one partition
    a["payload"] = np.arange(a.shape[0])
    out, off = Hgpu.radix_partition(a, 12)
    assert (sort_tuples(out) == sort_tuples(a)).all() and off[(12345 & 4095) + 1] - off[12345 & 4095] == a.shape[0]
    for n in (0, 1, 2, 3, 31, 33):
        b = np.zeros(n, dtype=Hgpu.TUPLE)
        b["key"] = np.arange(n) * 7
        out, off = Hgpu.radix_partition(b, 9)
        assert off[-1] == n and (sort_tuples(out) == sort_tuples(b)).all()


# ---- full join vs oracle on identical inputs ----------------------------------------------------------------------
JOIN_CASES = [(0, 1 << 21, 1, 512), (0, 1 << 21, 2, 512), (0, 1 << 21, 7, 512), (1, 1 << 21, 1, 512),
              (1, 1 << 21, 3, 512), (1, 1 << 21, 4, 256), (1, 1 << 21, 6, 64), (1, 1 << 21, 3, 1024)]


@pytest.mark.parametrize("case", JOIN_CASES, ids=str)
def test_bpro_vs_oracle(Hgpu, oracle_mod, case):
    variant, m, k, B = case
    R, S = inputs(oracle_mod, 250_000, 2_000_000, 0.01)
    Rc, Sc = R.copy(), S.copy()
    res = Hgpu.BPRO(R, S, 8, Hgpu.BloomFilterArgs(variant, m, k, B))
    assert scalars(res) == oscalars(oracle_mod.join(R, S, True, variant, m, k, B))
    assert res.nthreads == 8
    assert (R == Rc).all() and (S == Sc).all()  # unlike the reference the inputs are left untouched


@pytest.mark.parametrize("m,k", [(1 << 21, 1), (1 << 24, 1), (1 << 16, 1), (1 << 21, 0)])
def test_filter_built_inside_the_join_is_byte_identical(Hgpu, oracle_mod, m, k):
    """BASIC k<=1 joins partition on the filter-slice index and build the filter in shared memory (K1'); the
    resulting bitmap must still be the reference's, byte for byte, whatever the radix fan-out."""
    R, S = inputs(oracle_mod, 250_000, 2_000_000, 0.01)
    exp = oracle_mod.bloom_build(R, 0, m, k, 512)
    for bits in (0, 4, 9, 13):
        Hgpu.set_radix_bits(bits)
        Hgpu.set_hash_partition(2)  # force the shared-memory build also for these small filters
        try:
            res = Hgpu.BPRO(R, S, 1, Hgpu.BloomFilterArgs(0, m, k, 512))
            assert Hgpu.last_filter(m).tobytes() == exp.tobytes(), (m, k, bits)
            assert scalars(res) == oscalars(oracle_mod.join(R, S, True, 0, m, k, 512))
            Hgpu.set_hash_partition(0)  # and the atomic build of the radix-partitioned path
            res = Hgpu.BPRO(R, S, 1, Hgpu.BloomFilterArgs(0, m, k, 512))
            assert Hgpu.last_filter(m).tobytes() == exp.tobytes(), (m, k, bits)
        finally:
            Hgpu.set_radix_bits(0)
            Hgpu.set_hash_partition(1)


@pytest.mark.parametrize("q", [0.001, 0.1, 0.5, 1.0])
def test_selectivity_sweep(Hgpu, oracle_mod, q):
    R, S = inputs(oracle_mod, 250_000, 2_000_000, q)
    res = Hgpu.BPRO(R, S, 2, Hgpu.BloomFilterArgs(0, 1 << 21, 1, 512))
    assert scalars(res) == oscalars(oracle_mod.join(R, S, True, 0, 1 << 21, 1, 512))


@pytest.mark.parametrize("name", ["PRO", "RJ", "PRH", "PRHO"])
def test_plain_joins(Hgpu, oracle_mod, name):
    R, S = inputs(oracle_mod, 250_000, 2_000_000, 0.5)
    res = Hgpu.run(name, R, S, 4)
    exp = oracle_mod.join(R, S, False)
    assert scalars(res) == (exp["matches"], -1) + oscalars(exp)[2:]
    assert res.nthreads == (1 if name == "RJ" else 4)


@pytest.mark.parametrize("name", ["RJ", "PRH", "PRHO"])
def test_bloom_aliases(Hgpu, oracle_mod, name):
    R, S = inputs(oracle_mod, 250_000, 2_000_000, 0.01)
    res = Hgpu.run(name, R, S, 4, Hgpu.BloomFilterArgs(1, 1 << 21, 3, 512))
    assert scalars(res) == oscalars(oracle_mod.join(R, S, True, 1, 1 << 21, 3, 512))


@pytest.mark.parametrize("bits", [1, 6, 9, 12, 14])
def test_result_independent_of_radix_bits(Hgpu, oracle_mod, bits):
    """NUM_RADIX_BITS / NUM_PASSES do not change the result"""
    R, S = inputs(oracle_mod, 250_000, 2_000_000, 0.01)
    exp = oscalars(oracle_mod.join(R, S, True, 0, 1 << 21, 2, 512))
    Hgpu.set_radix_bits(bits)
    try:
        res = Hgpu.BPRO(R, S, 1, Hgpu.BloomFilterArgs(0, 1 << 21, 2, 512))
        assert res.stats["radix_bits"] == bits
        assert scalars(res) == exp
    finally:
        Hgpu.set_radix_bits(0)


def test_duplicate_build_keys_and_negative_keys(Hgpu, oracle_mod):
    """--non-unique style R (bucket chains longer than one; every duplicate counts, :303-315) and signed keys"""
    rng = np.random.default_rng(5)
    R = np.zeros(300_000, dtype=Hgpu.TUPLE)
    R["key"] = rng.integers(-5000, 60_000, R.shape[0]).astype(np.int32)
    R["payload"] = np.arange(R.shape[0])
    S = np.zeros(900_001, dtype=Hgpu.TUPLE)
    S["key"] = rng.integers(-8000, 120_000, S.shape[0]).astype(np.int32)
    S["payload"] = rng.integers(0, 2**31, S.shape[0]).astype(np.int32)
    res = Hgpu.BPRO(R, S, 1, Hgpu.BloomFilterArgs(1, 1 << 20, 2, 512))
    assert scalars(res) == oscalars(oracle_mod.join(R, S, True, 1, 1 << 20, 2, 512))
    res = Hgpu.PRO(R, S, 1)
    exp = oracle_mod.join(R, S, False)
    assert (res.totalresults, res.checksum_pair) == (exp["matches"], exp["checksum_pair"])


def test_heavily_skewed_partitions(Hgpu, oracle_mod):
    """one hot key (long chains + one huge S partition split over several work items) and an R partition larger
    than one shared-memory table (multi-round build)"""
    R = np.zeros(200_000, dtype=Hgpu.TUPLE)
    R["key"] = np.arange(R.shape[0]) * 4096 + 7  # all keys share the low 12 bits -> a single partition
    R["key"][:50] = 7
    R["payload"] = np.arange(R.shape[0])
    S = np.zeros(1_000_000, dtype=Hgpu.TUPLE)
    S["key"] = 7
    S["key"][::3] = R["key"][np.arange(0, 1_000_000, 3) % R.shape[0]]
    S["payload"] = np.arange(S.shape[0])
    for args in (None, Hgpu.BloomFilterArgs(0, 1 << 22, 2, 512)):
        res = Hgpu.run("PRO", R, S, 1, args)
        exp = oracle_mod.join(R, S, args is not None, 0, 1 << 22, 2, 512)
        assert (res.totalresults, res.checksum_pair, res.checksum_key) == (exp["matches"], exp["checksum_pair"], exp["checksum_key"])


def test_one_key_dominates_the_probe_relation(Hgpu, oracle_mod):
    """more than 4 M probe tuples in ONE partition: its > 128 work items are written by the whole CTA of k_worklist and
    processed by many CTAs of k_join (what a Zipf-skewed S does at full size)"""
    R = np.zeros(60_000, dtype=Hgpu.TUPLE)
    R["key"] = np.arange(1, R.shape[0] + 1)
    R["key"][:3] = 77  # the hot key has three build tuples: three pairs per hot probe tuple
    R["payload"] = np.arange(R.shape[0]) + 11
    S = np.zeros(5_300_000, dtype=Hgpu.TUPLE)
    S["key"] = 77
    S["key"][::9] = (np.arange(0, S.shape[0], 9) % 120_000) + 1
    S["payload"] = np.arange(S.shape[0])
    for args, oargs in [(None, (False,)), (Hgpu.BloomFilterArgs(0, 1 << 21, 1, 512), (True, 0, 1 << 21, 1, 512))]:
        res = Hgpu.run("PRO", R, S, 1, args)
        exp = oracle_mod.join(R, S, *oargs)
        assert (res.totalresults, res.checksum_pair, res.checksum_key) == (exp["matches"], exp["checksum_pair"], exp["checksum_key"])
        if args is not None:
            assert res.filtered == exp["filtered"]


def test_zipf_probe_relation(Hgpu, oracle_mod):
    """BASELINE config 5 (-z 1.0) at a CPU-checkable size: every S tuple matches and passes the filter"""
    r, s = 200_000, 1_500_000
    R = oracle_mod.gen_R(r)
    S = oracle_mod.gen_zipf(s, r, 1.0)
    res = Hgpu.BPRO(R, S, 1, Hgpu.BloomFilterArgs(0, 1 << 21, 1, 512))
    assert scalars(res) == oscalars(oracle_mod.join(R, S, True, 0, 1 << 21, 1, 512))
    assert res.totalresults == s and res.filtered == s


@pytest.mark.parametrize("nr,ns", [(0, 0), (0, 10), (10, 0), (1, 1), (3, 5), (1000, 7), (8193, 100_001)])
def test_empty_and_ragged_inputs(Hgpu, oracle_mod, nr, ns):
    R = np.zeros(nr, dtype=Hgpu.TUPLE)
    R["key"] = np.arange(1, nr + 1)
    R["payload"] = np.arange(nr)
    S = np.zeros(ns, dtype=Hgpu.TUPLE)
    S["key"] = (np.arange(ns) * 3) % (2 * max(nr, 1)) + 1
    S["payload"] = np.arange(ns)
    res = Hgpu.BPRO(R, S, 1, Hgpu.BloomFilterArgs(1, 1 << 16, 3, 64))
    assert scalars(res) == oscalars(oracle_mod.join(R, S, True, 1, 1 << 16, 3, 64))
    res = Hgpu.PRO(R, S, 1)
    assert res.totalresults == oracle_mod.join(R, S, False)["matches"]


def test_result_materialisation(Hgpu, oracle_mod):
    """the output pairs {R.payload, S.payload} (JOIN_RESULT_MATERIALIZE, :307-312) as a multiset"""
    rng = np.random.default_rng(9)
    R = np.zeros(150_000, dtype=Hgpu.TUPLE)
    R["key"] = rng.integers(1, 60_000, R.shape[0]).astype(np.int32)  # duplicate build keys: several pairs per S tuple
    R["payload"] = np.arange(R.shape[0])
    S = np.zeros(400_000, dtype=Hgpu.TUPLE)
    S["key"] = rng.integers(1, 120_000, S.shape[0]).astype(np.int32)
    S["payload"] = np.arange(S.shape[0]) + 7_000_000
    for args, oargs in [(None, (False,)), (Hgpu.BloomFilterArgs(0, 1 << 20, 1, 512), (True, 0, 1 << 20, 1, 512)),
                        (Hgpu.BloomFilterArgs(1, 1 << 20, 3, 256), (True, 1, 1 << 20, 3, 256))]:
        res = Hgpu.run("PRO", R, S, 1, args)
        exp = oracle_mod.join_pairs(R, S, *oargs)
        assert res.totalresults == exp.shape[0]
        got = Hgpu.materialize_last(16)  # too small on purpose: the count comes back and the call retries
        assert got.shape[0] == exp.shape[0]
        assert (sort_tuples(got) == sort_tuples(exp)).all()
    Hgpu.set_hash_partition(2)
    try:
        res = Hgpu.run("PRO", R, S, 1, Hgpu.BloomFilterArgs(0, 1 << 20, 1, 512))
        got = Hgpu.materialize_last(res.totalresults)
        assert (sort_tuples(got) == sort_tuples(oracle_mod.join_pairs(R, S, True, 0, 1 << 20, 1, 512))).all()
    finally:
        Hgpu.set_hash_partition(1)


def test_chunked_upload_overlapped_with_probe(Hgpu, oracle_mod):
    """host-buffer call with the S upload split into chunks that are probed as they land"""
    from hwbloomradixjoin_b200 import _native as N
    R = oracle_mod.gen_R(600_000, nthreads=4)
    S = oracle_mod.gen_S(9_000_001, 600_000, 0.02, nthreads=4)
    for variant, m, k, B in [(0, 1 << 23, 1, 512), (1, 1 << 23, 3, 256), (0, 1 << 23, 2, 512)]:
        exp = oscalars(oracle_mod.join(R, S, True, variant, m, k, B))
        N.load().hwbrj_set_overlap_h2d(1)
        try:
            got = Hgpu.BPRO(R, S, 1, Hgpu.BloomFilterArgs(variant, m, k, B))
        finally:
            N.load().hwbrj_set_overlap_h2d(0)
        assert scalars(got) == exp


def test_device_resident_join_matches_host_buffer_join(Hgpu, oracle_mod):
    R, S = inputs(oracle_mod, 250_000, 2_000_000, 0.01)
    dR, dS = Hgpu.DeviceRelation.upload(R), Hgpu.DeviceRelation.upload(S)
    assert len(dR) == R.shape[0] and (dS.download() == S).all()
    a = Hgpu.join_device(dR, dS, Hgpu.BloomFilterArgs(0, 1 << 21, 1, 512))
    b = Hgpu.BPRO(R, S, 1, Hgpu.BloomFilterArgs(0, 1 << 21, 1, 512))
    assert scalars(a) == scalars(b)
    assert a.stats["kernel_launches"] > 0 and a.stats["ms_total"] > 0
    dR.free()
    dS.free()


def test_device_generator_multiset(Hgpu, oracle_mod):
    r, s, q = 100_003, 700_001, 0.01
    dR = Hgpu.DeviceRelation.generate(0, r, r, 1.0, 3)
    dS = Hgpu.DeviceRelation.generate(1, s, r, q, 4)
    R, S = dR.download(), dS.download()
    assert (np.sort(R["key"]) == np.sort(oracle_mod.gen_R(r)["key"])).all()
    assert (np.sort(S["key"]) == np.sort(oracle_mod.gen_S(s, r, q)["key"])).all()
    assert (R["payload"] == np.arange(r)).all() and (S["payload"] == np.arange(s)).all()
    assert (R["key"][:1000] != np.arange(1, 1001)).any()  # positions are shuffled
    res = Hgpu.join_device(dR, dS, Hgpu.BloomFilterArgs(0, 1 << 20, 1, 512))
    assert scalars(res) == oscalars(oracle_mod.join(R, S, True, 0, 1 << 20, 1, 512))


# ---- golden values of the reference's published data -----------------------------------------------------------------
def _gold_groups():
    groups = {}
    for c in GOLD:
        groups.setdefault((c["r"], c["s"], c["q"]), []).append(c)
    return sorted(groups.items())


@pytest.mark.parametrize("grp", _gold_groups(), ids=lambda g: f"r{g[0][0]}-s{g[0][1]}-q{g[0][2]}")
def test_golden_published_results(Hgpu, grp):
    """all 632 distinct configurations of measurements/data/pkl: `filtered` and `out-tuples` are functions of the
    key multiset only, so device-generated inputs with the reference generator's multiset must reproduce them."""
    (r, s, q), configs = grp
    dR = Hgpu.DeviceRelation.generate(0, r, r, 1.0, 1)
    dS = Hgpu.DeviceRelation.generate(1, s, r, q, 2)
    try:
        for c in configs:
            if c["bloom"] == "no":
                res = Hgpu.join_device(dR, dS, None)
            else:
                res = Hgpu.join_device(dR, dS, Hgpu.BloomFilterArgs(0 if c["bloom"] == "basic" else 1, c["m"], c["k"], c["B"]))
                assert res.filtered == c["filtered"], c
            assert res.totalresults == c["matches"], c
    finally:
        dR.free()
        dS.free()


def test_canonical_c1_and_c0(Hgpu):
    """BASELINE.json configs 0 and 1 at full size: golden scalars, closed-form key checksum, idempotence."""
    for r, s, m, filt, mt in [(16_000_000, 256_000_000, 1 << 27, 31_038_115, 2_560_000),
                              (128_000_000, 1_024_000_000, 1 << 30, 124_152_740, 10_240_000)]:
        dR = Hgpu.DeviceRelation.generate(0, r, r, 1.0, 1)
        dS = Hgpu.DeviceRelation.generate(1, s, r, 0.01, 2)
        a = Hgpu.join_device(dR, dS, Hgpu.BloomFilterArgs(0, m, 1, 512))
        assert (a.totalresults, a.filtered) == (mt, filt)
        assert a.checksum_key == mt * (mt + 1) // 2          # sum of keys 1..nb
        assert a.checksum_rpay < 2**64 and a.checksum_spay > 0
        b = Hgpu.join_device(dR, dS, Hgpu.BloomFilterArgs(0, m, 1, 512))   # idempotent: inputs are read-only
        assert scalars(a) == scalars(b)
        p = Hgpu.join_device(dR, dS, None)                   # the filter never changes the join result
        assert scalars(p)[0] == mt and scalars(p)[2:] == scalars(a)[2:]
        dR.free()
        dS.free()
