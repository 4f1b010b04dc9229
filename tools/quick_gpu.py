"""Quick GPU sanity + timing run (development helper; the real checks live in tests/)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import hwbloomradixjoin_b200 as H

H.set_quiet(True)
rng = np.random.default_rng(1)
keys = np.concatenate([rng.integers(-2**31, 2**31, 100000, dtype=np.int64).astype(np.int32),
                       np.array([0, 1, -1, 2**31 - 1, -2**31, 1000], dtype=np.int32)])
for w in range(10):
    a = H.hash_many(w, 42, keys); b = oracle.hash_many(w, 42, keys)
    assert (a == b).all(), ("hash", w)
print("hash ok")
R = oracle.gen_R(250000); S = oracle.gen_S(2000000, 250000, 0.01)
for variant, k, B, m in [(0, 1, 512, 1 << 21), (0, 3, 512, 1 << 21), (1, 1, 512, 1 << 21), (1, 4, 64, 1 << 21), (1, 3, 1024, 1 << 21), (1, 2, 8, 1 << 21)]:
    args = H.BloomFilterArgs(variant, m, k, B)
    bm = H.bloom_build(R, args); bo = oracle.bloom_build(R, variant, m, k, B)
    assert (bm == bo).all(), ("bitmap", variant, k, B)
    n, surv = H.bloom_probe(bo, S, args, want_survivors=True)
    no, so = oracle.bloom_filter(bo, S, variant, m, k, B, want_survivors=True)
    assert n == no, ("probe", n, no)
    assert (np.sort(surv, order=["key", "payload"]) == np.sort(so, order=["key", "payload"])).all()
print("bloom ok")
for bits in [0, 3, 7, 8, 11, 14]:
    out, off = H.radix_partition(S, bits)
    assert off[-1] == S.shape[0]
    pid = out["key"].astype(np.uint32) & ((1 << bits) - 1)
    assert (np.diff(pid.astype(np.int64)) >= 0).all(), ("partition order", bits)
    cnt = np.bincount(S["key"].astype(np.uint32) & ((1 << bits) - 1), minlength=1 << bits)
    assert (np.diff(off.astype(np.int64)) == cnt).all()
    assert (np.sort(out, order=["key", "payload"]) == np.sort(S, order=["key", "payload"])).all()
print("partition ok")
for variant, k, B in [(0, 1, 512), (0, 2, 512), (1, 3, 512)]:
    args = H.BloomFilterArgs(variant, 1 << 21, k, B)
    r = H.BPRO(R, S, 4, args); o = oracle.join(R, S, True, variant, 1 << 21, k, B)
    print(variant, k, r.totalresults, r.filtered, o["matches"], o["filtered"])
    assert (r.totalresults, r.filtered, r.checksum_pair, r.checksum_rpay, r.checksum_spay, r.checksum_key) == \
           (o["matches"], o["filtered"], o["checksum_pair"], o["checksum_rpay"], o["checksum_spay"], o["checksum_key"])
r = H.PRO(R, S, 4); o = oracle.join(R, S, False)
assert (r.totalresults, r.checksum_pair) == (o["matches"], o["checksum_pair"])
print("join ok")

def timed(name, r, s, q, args, reps=3):
    dR = H.DeviceRelation.generate(0, r, r, 1.0, 1); dS = H.DeviceRelation.generate(1, s, r, q, 2)
    for i in range(reps):
        res = H.join_device(dR, dS, args)
    st = res.stats
    print(f"{name}: matches={res.totalresults} filtered={res.filtered} total={st['ms_total']:.3f} ms "
          f"[build {st['ms_build']:.3f} partR {st['ms_part_r']:.3f} probe {st['ms_probe']:.3f} partS {st['ms_part_s']:.3f} "
          f"join {st['ms_join']:.3f}] memset {st['ms_memset']:.3f} bits={st['radix_bits']} ranges={st['range_passes']} "
          f"-> {(r + s) / st['ms_total'] / 1e3:.1f} Mtuples/s", flush=True)
    dR.free(); dS.free()
    return res

timed("16M/128M basic k1", 16_000_000, 128_000_000, 0.01, H.BloomFilterArgs(0, 1 << 27, 1, 512))
timed("C0 16M/256M basic k1", 16_000_000, 256_000_000, 0.01, H.BloomFilterArgs(0, 1 << 27, 1, 512))
timed("C3 plain 128M/128M", 128_000_000, 128_000_000, 1.0, None)
timed("C1 128M/1024M basic k1 m=2^30", 128_000_000, 1_024_000_000, 0.01, H.BloomFilterArgs(0, 1 << 30, 1, 512))
H.set_range_passes(1)
timed("C1 (1 range pass)", 128_000_000, 1_024_000_000, 0.01, H.BloomFilterArgs(0, 1 << 30, 1, 512))
H.set_range_passes(0)
timed("C1 blocked k3 B512", 128_000_000, 1_024_000_000, 0.01, H.BloomFilterArgs(1, 1 << 30, 3, 512))
timed("C1 blocked k4 B256", 128_000_000, 1_024_000_000, 0.01, H.BloomFilterArgs(1, 1 << 30, 4, 256))
timed("C1 basic k2", 128_000_000, 1_024_000_000, 0.01, H.BloomFilterArgs(0, 1 << 30, 2, 512))
