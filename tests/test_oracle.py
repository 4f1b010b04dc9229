"""The oracle (oracle/oracle.c) pinned against the reference's own artefacts: hash and bitmap known-answer vectors
generated from the compiled reference, the golden filtered/matches values mined from the reference's published
measurements, and -- when oracle/_ref was built -- the unmodified reference itself on identical arrays."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_results.json")))["configs"]
HASH_KAT = json.load(open(os.path.join(HERE, "golden", "hash_kat.json")))
BITMAP_KAT = json.load(open(os.path.join(HERE, "golden", "bitmap_kat.json")))

# SURVEY.md Appendix C, columns crc, crapwow, spooky, fnv, murmurOAAT, jenkinsOAAT, coffin, kr_v2, djb2, x17
APPENDIX_C = """42,0 9e0654ec 2da34228 73c2d6d8 72d84ddf 8ea94b8f e62a1834 55555555 024fdb2a 7c5d0f85 0032fa18
42,1 4343fe54 f253226c 82d1da52 22ac0ece 5dd8ea29 268a9bdb 55455555 02504f89 7c5d9be6 00330d68
42,2 2161776d 0b31e533 bb1148ea d27fcfbd bc603177 58a12855 55755555 0250c3e8 7c5e2847 003320bf
42,1000 ac4f752a 483d2c1b cd726bb4 aa53eec2 822a8e9e 025e1c42 abdb2aaa 0244fd85 7c4ff330 003130c4
42,128000000 3c624a9f 0547e84d 04db1297 1a30702f 9cee2a0f a1c2dda8 aabbd24a 025047d0 7c5d8b6d 00331831
42,128000001 e127e027 428b407d 9e75a484 84d3997e b307cc6e 3b2ec218 aaabd24a 0250bc2f 7c5e17ce 00332b00
42,2147483647 ab68dbac f194aa53 55619166 c8e060fb d9523664 32dc1d16 aaaaa54a 024f636a 7c5c7f41 0032e674
42,-1 299ee0d4 19d803b1 4e1ecbbc a920527b 9adfecfe 8bd74f5f 55555555 024f62ea 7c5c7ec1 0032e5f4
42,-2147483648 1cf06f94 9a1462f4 d3d6825e 1726f85f a3c9face c1754eff aaaaa54a 024fdaaa 7c5d0f05 0032f998
0,0 00000000 c6ae4c2b adce103f 4b95f515 00000000 00000000 55555555 00000000 7c5d0f85 fffd8c7d
0,1 dd45aab8 acc387b7 5a8b5863 fb69b604 f6ecd433 009dbee6 55455555 0000745f 7c5d9be6 fffd794c
0,-1 b798b438 c45889f8 502fc5f0 81ddf9b1 61654d64 f7c21986 55555555 ffff87c0 7c5c7ec1 fffda0e1"""
APPENDIX_C_COLS = [0, 2, 6, 1, 4, 5, 3, 7, 8, 9]  # column -> index in hash.h order


def test_hash_appendix_c(oracle_mod):
    for line in APPENDIX_C.splitlines():
        parts = line.split()
        seed, key = map(int, parts[0].split(","))
        for col, which in enumerate(APPENDIX_C_COLS):
            assert oracle_mod.hash_one(which, seed, key) == int(parts[1 + col], 16), (seed, key, which)


def test_hash_kat_from_reference(oracle_mod):
    for row in HASH_KAT["rows"]:
        for which in range(10):
            assert oracle_mod.hash_one(which, row["seed"], row["key"]) == row["hashes"][which], (row, which)


def test_bitmap_kat(oracle_mod):
    R = np.zeros(64, dtype=oracle_mod.TUPLE)
    R["key"] = np.arange(1, 65)
    S = np.zeros(1000, dtype=oracle_mod.TUPLE)
    S["key"] = np.arange(65, 1065)
    for row in BITMAP_KAT:
        bm = oracle_mod.bloom_build(R, row["variant"], row["m"], row["k"], row["B"])
        assert bm.tobytes().hex() == row["bitmap_hex"], row
        assert oracle_mod.bloom_filter(bm, S, row["variant"], row["m"], row["k"], row["B"]) == row["pass"]
    # the four vectors printed in SURVEY.md Appendix C (first bytes) and their pass counts
    by = {(r["variant"], r["k"], r["B"]): r for r in BITMAP_KAT}
    assert by[(0, 1, 512)]["bitmap_hex"].startswith("0800000001000000030000000400000000002000000020100100048800080000")
    assert by[(0, 1, 512)]["pass"] == 55 and by[(0, 3, 512)]["pass"] == 5
    assert by[(1, 3, 64)]["pass"] == 3 and by[(1, 1, 512)]["pass"] == 59


def _small_gold():
    rows = [c for c in GOLD if c["r"] == 250000]
    # every (s, q, variant) combination, a spread of k and m: keeps the CPU suite to about a minute
    keep = []
    for i, c in enumerate(rows):
        if c["bloom"] == "no" or c["s"] == 2000000 or i % 3 == 0:
            keep.append(c)
    return keep


@pytest.mark.parametrize("c", _small_gold(), ids=lambda c: f"r{c['r']}-s{c['s']}-q{c['q']}-{c['bloom']}-m{c['m']}-k{c['k']}")
def test_golden_small(oracle_mod, c):
    R = _cached_R(oracle_mod, c["r"])
    S = _cached_S(oracle_mod, c["s"], c["r"], c["q"])
    if c["bloom"] == "no":
        res = oracle_mod.join(R, S, False)
    else:
        res = oracle_mod.join(R, S, True, 0 if c["bloom"] == "basic" else 1, c["m"], c["k"], c["B"])
        assert res["filtered"] == c["filtered"]
    assert res["matches"] == c["matches"]


_cache = {}


def _cached_R(o, r):
    if ("R", r) not in _cache:
        _cache[("R", r)] = o.gen_R(r, nthreads=4)
    return _cache[("R", r)]


def _cached_S(o, s, r, q):
    if ("S", s, r, q) not in _cache:
        _cache[("S", s, r, q)] = o.gen_S(s, r, q, nthreads=4)
    return _cache[("S", s, r, q)]


def test_survey_probed_goldens(oracle_mod):
    """[probed] rows of SURVEY.md Appendix B that are not in the published pickles."""
    R = oracle_mod.gen_R(1_000_000)
    S = oracle_mod.gen_S(8_000_000, 1_000_000, 0.01)
    res = oracle_mod.join(R, S, True, 0, 1 << 23, 1, 512)
    assert (res["filtered"], res["matches"]) == (969_892, 80_000)
    R = _cached_R(oracle_mod, 250000)
    S = _cached_S(oracle_mod, 2000000, 250000, 0.01)
    for B, k, exp in [(64, 1, 244110), (64, 8, 67192), (128, 4, 62551), (256, 6, 57563), (1024, 3, 73754), (1024, 8, 60816)]:
        assert oracle_mod.join(R, S, True, 1, 1 << 21, k, B)["filtered"] == exp
    for q, f, mt in [(0.001, 226037, 2000), (0.1, 401833, 200000), (0.5, 1112400, 1000000), (1.0, 2000000, 2000000)]:
        S = _cached_S(oracle_mod, 2000000, 250000, q)
        res = oracle_mod.join(R, S, True, 0, 1 << 21, 1, 512)
        assert (res["filtered"], res["matches"]) == (f, mt)


def test_generator_closed_form(oracle_mod):
    """the key multiset is closed-form and independent of the generator's thread count."""
    r, s, q = 100_000, 700_001, 0.01
    for nthr in (1, 3, 8):
        R = oracle_mod.gen_R(r, nthreads=nthr, seed=nthr)
        assert (np.sort(R["key"]) == np.arange(1, r + 1)).all()
        assert (R["payload"] == np.arange(r)).all()
        S = oracle_mod.gen_S(s, r, q, nthreads=nthr, seed=nthr + 10)
        na = int(s * (1 - q))
        nb = s - na
        exp = np.sort(np.concatenate([np.arange(nb) % r + 1, r + 1 + np.arange(na)]))
        assert (np.sort(S["key"]) == exp).all()
        assert (S["payload"] == np.arange(s)).all()


def _need_ref(o, mat=False):
    if not o.ref_available(mat):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")


@pytest.mark.parametrize("cfg", [("PRO", True, 0, 1 << 21, 1, 512), ("PRO", True, 0, 1 << 21, 4, 512),
                                 ("PRO", True, 1, 1 << 21, 3, 512), ("PRO", True, 1, 1 << 21, 5, 64),
                                 ("RJ", True, 0, 1 << 20, 2, 512), ("PRH", True, 1, 1 << 21, 2, 1024),
                                 ("PRHO", True, 0, 1 << 22, 1, 512), ("PRO", False, 0, 0, 0, 0), ("RJ", False, 0, 0, 0, 0)])
def test_oracle_vs_compiled_reference(oracle_mod, cfg):
    """identical arrays through the unmodified reference (materialising build) and through the restatement"""
    _need_ref(oracle_mod, mat=True)
    algo, bloom, variant, m, k, B = cfg
    R = _cached_R(oracle_mod, 250000)
    S = _cached_S(oracle_mod, 2000000, 250000, 0.01)
    ref = oracle_mod.ref_join(R, S, algo, 4, bloom, variant, m or 1 << 20, k, B or 512, mat=True)
    orc = oracle_mod.join(R, S, bloom, variant, m or 1 << 20, k, B or 512)
    assert ref["matches"] == orc["matches"]
    if bloom and algo != "RJ":  # BRJ does not print the filtered count
        assert ref["filtered"] == orc["filtered"]
    if algo in ("PRO", "RJ"):  # only bucket_chaining_join materialises pairs (:307-312); PRH/PRHO just count
        for f in ("checksum_pair", "checksum_rpay", "checksum_spay"):
            assert ref[f] == orc[f], f


def test_oracle_filter_vs_compiled_reference(oracle_mod):
    _need_ref(oracle_mod)
    rng = np.random.default_rng(7)
    R = np.zeros(50_000, dtype=oracle_mod.TUPLE)
    R["key"] = rng.integers(-2**31, 2**31, R.shape[0], dtype=np.int64).astype(np.int32)  # full range, negative keys
    S = np.zeros(200_000, dtype=oracle_mod.TUPLE)
    S["key"] = rng.integers(-2**31, 2**31, S.shape[0], dtype=np.int64).astype(np.int32)
    S["key"][:1000] = R["key"][:1000]
    for variant, m, k, B in [(0, 1 << 19, 1, 512), (0, 1 << 19, 7, 512), (1, 1 << 19, 3, 512), (1, 1 << 19, 8, 64),
                             (1, 1 << 18, 2, 8), (1, 1 << 19, 6, 1 << 19)]:
        a = oracle_mod.bloom_build(R, variant, m, k, B)
        b = oracle_mod.ref_bloom_build(R, variant, m, k, B)
        assert (a == b).all()
        assert oracle_mod.bloom_filter(a, S, variant, m, k, B) == oracle_mod.ref_bloom_count(b, S, variant, m, k, B)


def _random_filter_configs(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        variant = int(rng.integers(0, 2))
        log2m = int(rng.integers(10, 23))
        k = int(rng.integers(0, 13))
        log2B = int(rng.integers(3, log2m + 1))  # B = 8 ... m
        out.append((variant, 1 << log2m, k, 1 << log2B))
    return out


def test_random_filter_configs_vs_compiled_reference(oracle_mod):
    """48 seeded random (variant, m, k, B) points: bitmap byte-equality and the probe count against the unmodified
    reference, on keys that cover the whole int32 range"""
    _need_ref(oracle_mod)
    rng = np.random.default_rng(2024)
    R = np.zeros(20_000, dtype=oracle_mod.TUPLE)
    R["key"] = rng.integers(-2**31, 2**31, R.shape[0], dtype=np.int64).astype(np.int32)
    S = np.zeros(60_000, dtype=oracle_mod.TUPLE)
    S["key"] = rng.integers(-2**31, 2**31, S.shape[0], dtype=np.int64).astype(np.int32)
    S["key"][:5000] = R["key"][:5000]
    for variant, m, k, B in _random_filter_configs(48, 11):
        a = oracle_mod.bloom_build(R, variant, m, k, B)
        b = oracle_mod.ref_bloom_build(R, variant, m, k, B)
        assert (a == b).all(), (variant, m, k, B)
        assert oracle_mod.bloom_filter(a, S, variant, m, k, B) == oracle_mod.ref_bloom_count(b, S, variant, m, k, B), \
            (variant, m, k, B)


@pytest.mark.parametrize("cfg", _random_filter_configs(8, 5), ids=lambda c: f"v{c[0]}-m{c[1]}-k{c[2]}-B{c[3]}")
def test_random_join_configs_vs_compiled_reference(oracle_mod, cfg):
    """the whole join (matches, filtered, pair checksums) at seeded random filter settings, duplicate build keys included"""
    _need_ref(oracle_mod, mat=True)
    variant, m, k, B = cfg
    rng = np.random.default_rng(m + k)
    R = np.zeros(40_000, dtype=oracle_mod.TUPLE)
    R["key"] = rng.integers(1, 30_000, R.shape[0]).astype(np.int32)  # duplicates: several pairs per probe tuple
    R["payload"] = np.arange(R.shape[0])
    S = np.zeros(250_000, dtype=oracle_mod.TUPLE)
    S["key"] = rng.integers(1, 300_000, S.shape[0]).astype(np.int32)
    S["payload"] = np.arange(S.shape[0]) + 1_000_000
    ref = oracle_mod.ref_join(R, S, "PRO", 4, True, variant, m, k, B, mat=True)
    orc = oracle_mod.join(R, S, True, variant, m, k, B)
    for f in ("matches", "filtered", "checksum_pair", "checksum_rpay", "checksum_spay"):
        assert ref[f] == orc[f], (cfg, f)


def test_zipf_generator_vs_reference(oracle_mod):
    _need_ref(oracle_mod)
    a = oracle_mod.gen_zipf(200_000, 50_000, 1.0, seed=54321)
    b = oracle_mod.ref_generate(2, 200_000, 50_000, zipf=1.0, seed=54321)
    assert (a["key"] == b["key"]).all()
    assert a["key"].min() >= 1 and a["key"].max() <= 50_000


def test_reference_generator_multiset(oracle_mod):
    """the reference's own (time-shuffled) generator yields the closed-form multiset our generators restate"""
    _need_ref(oracle_mod)
    r, s, q = 60_000, 400_000, 0.01
    R = oracle_mod.ref_generate(0, r, r, nthreads=4)
    S = oracle_mod.ref_generate(1, s, r, q=q, seed=54321, nthreads=4)
    assert (np.sort(R["key"]) == np.sort(oracle_mod.gen_R(r, nthreads=4)["key"])).all()
    assert (np.sort(S["key"]) == np.sort(oracle_mod.gen_S(s, r, q, nthreads=4)["key"])).all()


def test_glibc_rand_restatement(oracle_mod):
    """the lock-free copy of glibc's rand() used for the FPR samples is checked against libc itself"""
    import ctypes as C
    libc = C.CDLL(None)
    for seed in (817263, 817264, 1, 0, 54321):
        out = np.empty(5000, dtype=np.int32)
        oracle_mod.lib().orc_glibc_rand(seed, out.ctypes.data_as(C.c_void_p), out.shape[0])
        libc.srand(seed)
        assert out.tolist() == [libc.rand() for _ in range(out.shape[0])]
