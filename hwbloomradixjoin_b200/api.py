"""Host-side mirror of the reference's join interface (main.c:277-282,331-339,473-478) on top of the C ABI.

The same names, argument meaning and error behaviour as the reference's operators:
``BPRO/BRJ/BPRH/BPRHO(relR, relS, nthreads, bloom_filter_args)`` and ``PRO/RJ/PRH/PRHO(relR, relS, nthreads)``.
Relations are numpy arrays of dtype ``TUPLE`` ({int32 key; int32 payload}, types.h:37-40). Every call runs the
CUDA kernels of libhwbrj_cuda.so; nothing here computes on the CPU and nothing falls back to it.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _native as N

TUPLE = np.dtype([("key", "<i4"), ("payload", "<i4")])
BASIC, BLOCKED = N.BASIC, N.BLOCKED
HASH_NAMES = ["crc", "FNV", "crapwow", "Coffin", "MurmurOAAT_32", "JenkinsOAAT_32", "Spooky", "KR_v2", "DJB2", "x17"]


@dataclass
class BloomFilterArgs:
    """bloom_filter_args_t (bloom_filter.h:50-55); defaults are the CLI's (main.c:389-393)."""
    variant: int = BASIC
    m: int = 256 << 20
    k: int = 8
    B: int = 1024

    def to_c(self) -> N.BloomFilterArgsT:
        return N.BloomFilterArgsT(int(self.variant), int(self.m), int(self.k), int(self.B))

    def check(self) -> None:
        """assert_args (bloom_filter.c:26-34); raises instead of exiting."""
        m, B = int(self.m), int(self.B)
        if m <= 0 or m & (m - 1):
            raise ValueError("m must be a power of 2")
        if m > 1 << 32:
            raise ValueError("m must be at most 2^32")
        if self.variant != BASIC:
            if B <= 0 or B & (B - 1):
                raise ValueError("B must be a power 2")
            if B < 8 or m % B:
                raise ValueError("m must be a multiple of B")


@dataclass
class JoinResult:
    """result_t (types.h:59-63) plus what the reference only prints (filtered) and the checksums."""
    totalresults: int
    nthreads: int
    filtered: int = -1
    checksum_pair: int = 0
    checksum_rpay: int = 0
    checksum_spay: int = 0
    checksum_key: int = 0
    stats: dict = field(default_factory=dict)


def as_relation(a) -> np.ndarray:
    a = np.ascontiguousarray(a)
    if a.dtype != TUPLE:
        if a.dtype == np.int32 and a.ndim == 2 and a.shape[1] == 2:
            a = a.view(TUPLE).reshape(-1)
        else:
            raise TypeError("a relation is an array of {int32 key; int32 payload} tuples")
    return a


def _rel(a: np.ndarray) -> N.RelationT:
    return N.RelationT(a.ctypes.data_as(C.c_void_p), a.shape[0])


def _finish(L, res_ptr) -> JoinResult:
    st = N.StatsT()
    L.hwbrj_last_stats(C.byref(st))
    out = JoinResult(res_ptr.contents.totalresults, res_ptr.contents.nthreads, st.filtered, st.checksum_pair,
                     st.checksum_rpay, st.checksum_spay, st.checksum_key, st.as_dict())
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(C.cast(res_ptr, C.c_void_p))  # the caller frees result_t (main.c:490)
    return out


def _bloom_call(name: str, relR, relS, nthreads: int, args: BloomFilterArgs) -> JoinResult:
    if args is None:
        raise ValueError(f"{name} needs bloom_filter_args")
    args.check()
    L = N.load()
    R, S = as_relation(relR), as_relation(relS)
    cargs = args.to_c()
    rr, rs = _rel(R), _rel(S)
    return _finish(L, getattr(L, name)(C.byref(rr), C.byref(rs), int(nthreads), C.byref(cargs)))


def _plain_call(name: str, relR, relS, nthreads: int) -> JoinResult:
    L = N.load()
    R, S = as_relation(relR), as_relation(relS)
    rr, rs = _rel(R), _rel(S)
    return _finish(L, getattr(L, name)(C.byref(rr), C.byref(rs), int(nthreads)))


def BPRO(relR, relS, nthreads, bloom_filter_args):  # parallel_radix_join_bloom.c:1782
    return _bloom_call("BPRO", relR, relS, nthreads, bloom_filter_args)


def BRJ(relR, relS, nthreads, bloom_filter_args):  # :1808
    return _bloom_call("BRJ", relR, relS, nthreads, bloom_filter_args)


def BPRH(relR, relS, nthreads, bloom_filter_args):  # :1791
    return _bloom_call("BPRH", relR, relS, nthreads, bloom_filter_args)


def BPRHO(relR, relS, nthreads, bloom_filter_args):  # :1799
    return _bloom_call("BPRHO", relR, relS, nthreads, bloom_filter_args)


def PRO(relR, relS, nthreads):  # parallel_radix_join.c:1697
    return _plain_call("PRO", relR, relS, nthreads)


def RJ(relR, relS, nthreads):  # :1718
    return _plain_call("RJ", relR, relS, nthreads)


def PRH(relR, relS, nthreads):
    return _plain_call("PRH", relR, relS, nthreads)


def PRHO(relR, relS, nthreads):
    return _plain_call("PRHO", relR, relS, nthreads)


# the dispatch table of main.c:331-339 (name -> (joinAlgo, joinAlgoBloom)); NPO is out of scope
ALGOS = {"PRO": (PRO, BPRO), "RJ": (RJ, BRJ), "PRH": (PRH, BPRH), "PRHO": (PRHO, BPRHO)}


def run(algo: str, relR, relS, nthreads: int = 2, bloom: BloomFilterArgs | None = None) -> JoinResult:
    """main.c:473-478: joinAlgoBloom when the filter is enabled, else joinAlgo."""
    if algo not in ALGOS:
        raise KeyError(f"Join algorithm named `{algo}' does not exist!")  # main.c:625-629
    plain, withbloom = ALGOS[algo]
    return withbloom(relR, relS, nthreads, bloom) if bloom is not None else plain(relR, relS, nthreads)


def last_filter(m: int) -> np.ndarray:
    """bitmap (m/8 bytes) of the filter the most recent Bloom join built on the device"""
    out = np.empty(m // 8, dtype=np.uint8)
    if N.load().hwbrj_last_filter(out.ctypes.data_as(C.c_void_p), out.shape[0]) != 0:
        raise RuntimeError("no filter of that size")
    return out


def set_quiet(quiet: bool = True) -> None:
    N.load().hwbrj_set_quiet(int(quiet))


def set_radix_bits(bits: int) -> None:
    N.load().hwbrj_set_radix_bits(int(bits))


def set_range_passes(passes: int) -> None:
    N.load().hwbrj_set_range_passes(int(passes))


def set_num_passes(passes: int) -> None:
    """NUM_PASSES of the reference (prj_params.h:20-22) at run time: 1 or 2 scatter passes, 0 = automatic"""
    N.load().hwbrj_set_num_passes(int(passes))


def set_gpus(n: int) -> None:
    """host-buffer joins (BPRO, PRO, ...) shard over the first n GPUs of this process"""
    if N.load().hwbrj_set_gpus(int(n)) != 0:
        raise ValueError(f"cannot use {n} GPUs (power of two, at most the device count)")


def set_hash_partition(mode: int) -> None:
    """0 never, 1 automatic (filters > 32 MiB), 2 whenever possible"""
    N.load().hwbrj_set_hash_partition(int(mode))


def device_count() -> int:
    return N.load().hwbrj_device_count()


# ---- device-resident relations ------------------------------------------------------------------------------
class DeviceRelation:
    """A relation resident in HBM (hwbrj_rel_t)."""

    def __init__(self, handle: int):
        if not handle:
            raise RuntimeError("device relation allocation failed")
        self._h = handle

    @classmethod
    def upload(cls, rel) -> "DeviceRelation":
        a = as_relation(rel)
        return cls(N.load().hwbrj_rel_upload(a.ctypes.data_as(C.c_void_p), a.shape[0]))

    @classmethod
    def generate(cls, kind: int, n: int, r: int, q: float = 1.0, seed: int = 1) -> "DeviceRelation":
        """kind 0: R (keys 1..n shuffled); kind 1: S (FK over r with selectivity q) -- generator.c closed form;
        kind 2: S = n Zipf foreign keys over 1..r with exponent q (create_relation_zipf, generator.c:659-676)."""
        return cls(N.load().hwbrj_rel_generate(kind, n, r, q, seed))

    def __len__(self) -> int:
        return int(N.load().hwbrj_rel_size(self._h))

    def download(self) -> np.ndarray:
        out = np.empty(len(self), dtype=TUPLE)
        N.load().hwbrj_rel_download(self._h, out.ctypes.data_as(C.c_void_p))
        return out

    def free(self) -> None:
        if self._h:
            N.load().hwbrj_rel_free(self._h)
            self._h = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def join_device(R: DeviceRelation, S: DeviceRelation, bloom: BloomFilterArgs | None = None) -> JoinResult:
    if bloom is not None:
        bloom.check()
    L = N.load()
    st = N.StatsT()
    cargs = bloom.to_c() if bloom is not None else None
    rc = L.hwbrj_join_device(R._h, S._h, C.byref(cargs) if cargs is not None else None, C.byref(st))
    if rc != 0:
        raise RuntimeError(f"hwbrj_join_device rc={rc}")
    return JoinResult(st.matches, 1, st.filtered, st.checksum_pair, st.checksum_rpay, st.checksum_spay,
                      st.checksum_key, st.as_dict())


def materialize_last(expected: int) -> np.ndarray:
    """Output pairs {key = R.payload, payload = S.payload} of the most recent join (JOIN_RESULT_MATERIALIZE, :307-312)"""
    out = np.empty(max(int(expected), 0), dtype=TUPLE)
    n = N.load().hwbrj_materialize_last(out.ctypes.data_as(C.c_void_p), out.shape[0])
    if n < 0:
        raise RuntimeError("no join to materialise")
    if n > out.shape[0]:
        return materialize_last(n)
    return out[:n]


def fpr_count(R: DeviceRelation, S: DeviceRelation, bloom: BloomFilterArgs, seed: int) -> int:
    """test_bloom_fpr (unit_tests.c:191-241) on the device: filter built with `seed` from R, number of S keys passing"""
    bloom.check()
    cargs = bloom.to_c()
    n = N.load().hwbrj_fpr_count(R._h, S._h, C.byref(cargs), seed & 0xFFFFFFFF)
    if n < 0:
        raise RuntimeError("hwbrj_fpr_count failed")
    return int(n)


# ---- building blocks (parity tests) --------------------------------------------------------------------------
def hash_many(which: int, seed: int, keys) -> np.ndarray:
    keys = np.ascontiguousarray(keys, dtype=np.int32)
    out = np.empty(keys.shape[0], dtype=np.uint32)
    rc = N.load().hwbrj_hash_many(which, seed & 0xFFFFFFFF, keys.ctypes.data_as(C.c_void_p), keys.shape[0],
                                  out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise ValueError("unknown hash function")
    return out


def bloom_build(relR, args: BloomFilterArgs, seed: int = 42) -> np.ndarray:
    args.check()
    R = as_relation(relR)
    bitmap = np.empty(args.m // 8, dtype=np.uint8)
    cargs = args.to_c()
    rc = N.load().hwbrj_bloom_build(R.ctypes.data_as(C.c_void_p), R.shape[0], C.byref(cargs), seed,
                                    bitmap.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("hwbrj_bloom_build failed")
    return bitmap


def bloom_probe(bitmap: np.ndarray, relS, args: BloomFilterArgs, seed: int = 42, want_survivors: bool = False):
    args.check()
    S = as_relation(relS)
    bitmap = np.ascontiguousarray(bitmap, dtype=np.uint8)
    assert bitmap.shape[0] == args.m // 8
    surv = np.empty(S.shape[0], dtype=TUPLE) if want_survivors else None
    cargs = args.to_c()
    n = N.load().hwbrj_bloom_probe(bitmap.ctypes.data_as(C.c_void_p), S.ctypes.data_as(C.c_void_p), S.shape[0],
                                   C.byref(cargs), seed, surv.ctypes.data_as(C.c_void_p) if want_survivors else None)
    if n < 0:
        raise RuntimeError("hwbrj_bloom_probe failed")
    return (n, surv[:n]) if want_survivors else n


def radix_partition(rel, bits: int):
    a = as_relation(rel)
    out = np.empty(a.shape[0], dtype=TUPLE)
    offsets = np.empty((1 << bits) + 1, dtype=np.uint64)
    rc = N.load().hwbrj_radix_partition(a.ctypes.data_as(C.c_void_p), a.shape[0], bits,
                                        out.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise ValueError("radix_partition: unsupported arguments")
    return out, offsets
