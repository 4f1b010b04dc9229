"""oracle -- TEST INFRASTRUCTURE (checker only, never the product path).

ctypes loaders for
  * ``liboracle.so``        our single-threaded C restatement of the hot path (oracle/oracle.c)
  * ``_ref/libref.so``      the UNMODIFIED reference compiled from /root/reference/src (when it was
                            built in the authoring container; it travels to the GPU box prebuilt)
  * ``_ref/libref_mat.so``  same with -DJOIN_RESULT_MATERIALIZE (pair checksum)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package. The product (hwbloomradixjoin_b200, libhwbrj_cuda.so) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TUPLE = np.dtype([("key", "<i4"), ("payload", "<i4")])  # types.h:37-40

HASH_NAMES = ["crc", "fnv", "crapwow", "coffin", "murmur_oaat", "jenkins_oaat", "spooky", "kr_v2", "djb2", "x17"]


class OrcResult(C.Structure):
    _fields_ = [("matches", C.c_int64), ("filtered", C.c_int64), ("checksum_pair", C.c_uint64),
                ("checksum_rpay", C.c_uint64), ("checksum_spay", C.c_uint64), ("checksum_key", C.c_uint64)]


class RefResult(C.Structure):
    _fields_ = [("matches", C.c_int64), ("filtered", C.c_int64), ("total_usecs", C.c_double),
                ("part_usecs", C.c_double), ("join_usecs", C.c_double), ("checksum_pair", C.c_uint64),
                ("checksum_rpay", C.c_uint64), ("checksum_spay", C.c_uint64), ("materialized", C.c_int32),
                ("radix_bits", C.c_int32)]


def build(verbose: bool = False) -> None:
    """Compile liboracle.so and, when /root/reference is present, oracle/_ref (building the checker
    is not using it)."""
    cmd = ["make", "-s", "-f", os.path.join(HERE, "Makefile"), "all"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)


_lib = None
_ref = {}


def _tp(a):
    assert a.dtype == TUPLE and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_hash.restype = C.c_uint32
        L.orc_hash.argtypes = [C.c_int, C.c_uint32, C.c_int32]
        L.orc_hash_many.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_bloom_build.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64,
                                      C.c_uint32, C.c_void_p]
        L.orc_bloom_filter.restype = C.c_int64
        L.orc_bloom_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64,
                                       C.c_uint64, C.c_uint32, C.c_void_p]
        L.orc_join.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint64,
                               C.c_uint64, C.c_uint64, C.c_int, C.POINTER(OrcResult)]
        L.orc_join_pairs.restype = C.c_int64
        L.orc_join_pairs.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint64,
                                     C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64]
        L.orc_gen_relation.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_double,
                                       C.c_uint64]
        L.orc_gen_zipf.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_double, C.c_uint]
        L.orc_glibc_rand.argtypes = [C.c_uint, C.c_void_p, C.c_uint32]
        L.orc_fpr_samples.restype = C.c_uint32
        L.orc_fpr_samples.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def ref_available(mat: bool = False) -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_mat.so" if mat else "libref.so"))


def ref(mat: bool = False):
    """The compiled unmodified reference (RTLD_LOCAL: it exports BPRO/PRO/... like the product)."""
    if mat not in _ref:
        path = os.path.join(HERE, "_ref", "libref_mat.so" if mat else "libref.so")
        L = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        L.refshim_join.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_int,
                                   C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(RefResult)]
        L.refshim_bloom_build.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64,
                                          C.c_uint32, C.c_void_p]
        L.refshim_bloom_count.restype = C.c_int64
        L.refshim_bloom_count.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64,
                                          C.c_uint64, C.c_uint32, C.c_void_p]
        L.refshim_hash.restype = C.c_uint32
        L.refshim_hash.argtypes = [C.c_int, C.c_uint32, C.c_int32]
        L.refshim_generate.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_double, C.c_double,
                                       C.c_uint, C.c_int]
        _ref[mat] = L
    return _ref[mat]


# ----------------------------------------------------------------------------------------------
# oracle (restatement) API
# ----------------------------------------------------------------------------------------------
def hash_one(which: int, seed: int, key: int) -> int:
    return lib().orc_hash(which, seed & 0xFFFFFFFF, int(np.int32(np.uint32(key & 0xFFFFFFFF))))


def hash_many(which: int, seed: int, keys: np.ndarray) -> np.ndarray:
    keys = np.ascontiguousarray(keys, dtype=np.int32)
    out = np.empty(keys.shape[0], dtype=np.uint32)
    lib().orc_hash_many(which, seed, keys.ctypes.data_as(C.c_void_p), keys.shape[0], out.ctypes.data_as(C.c_void_p))
    return out


def bloom_build(R: np.ndarray, variant: int, m: int, k: int, B: int, seed: int = 42) -> np.ndarray:
    bitmap = np.zeros(m // 8, dtype=np.uint8)
    lib().orc_bloom_build(_tp(R), R.shape[0], variant, m, k, B, seed, bitmap.ctypes.data_as(C.c_void_p))
    return bitmap


def bloom_filter(bitmap: np.ndarray, S: np.ndarray, variant: int, m: int, k: int, B: int, seed: int = 42,
                 want_survivors: bool = False):
    surv = np.empty(S.shape[0], dtype=TUPLE) if want_survivors else None
    n = lib().orc_bloom_filter(bitmap.ctypes.data_as(C.c_void_p), _tp(S), S.shape[0], variant, m, k, B, seed,
                               _tp(surv) if want_survivors else None)
    return (n, surv[:n]) if want_survivors else n


def join(R: np.ndarray, S: np.ndarray, bloom: bool, variant: int = 0, m: int = 1 << 28, k: int = 8, B: int = 1024,
         radix_bits: int = 10) -> dict:
    res = OrcResult()
    rc = lib().orc_join(_tp(R), R.shape[0], _tp(S), S.shape[0], int(bloom), variant, m, k, B, radix_bits,
                        C.byref(res))
    if rc != 0:
        raise MemoryError("orc_join failed")
    return {f: getattr(res, f) for f, _ in OrcResult._fields_}


def join_pairs(R: np.ndarray, S: np.ndarray, bloom: bool, variant: int = 0, m: int = 1 << 28, k: int = 8, B: int = 1024,
               capacity: int = 1 << 22) -> np.ndarray:
    """materialised output {key = R.payload, payload = S.payload} (:307-312)"""
    pairs = np.empty(capacity, dtype=TUPLE)
    n = lib().orc_join_pairs(_tp(R), R.shape[0], _tp(S), S.shape[0], int(bloom), variant, m, k, B, 10, _tp(pairs), capacity)
    if n < 0:
        raise MemoryError("orc_join_pairs failed")
    if n > capacity:
        return join_pairs(R, S, bloom, variant, m, k, B, n)
    return pairs[:n]


def gen_relation(n: int, nthreads: int, maxid: int, threshold: int, selectivity: float, shuffle_seed: int) -> np.ndarray:
    rel = np.empty(n, dtype=TUPLE)
    lib().orc_gen_relation(_tp(rel), n, nthreads, maxid, threshold, selectivity, shuffle_seed)
    return rel


def gen_R(r: int, nthreads: int = 8, seed: int = 12345) -> np.ndarray:
    """main.c:430-431: parallel_create_relation(&relR, r, nthreads, r, r, 1.0)"""
    return gen_relation(r, nthreads, r, r, 1.0, seed)


def gen_S(s: int, r: int, q: float, nthreads: int = 8, seed: int = 54321) -> np.ndarray:
    """main.c:464-465: parallel_create_relation(&relS, s, nthreads, INT_MAX, r, q)"""
    return gen_relation(s, nthreads, 2147483647, r, q, seed)


def gen_zipf(s: int, r: int, theta: float, seed: int = 54321) -> np.ndarray:
    rel = np.empty(s, dtype=TUPLE)
    lib().orc_gen_zipf(_tp(rel), s, r, theta, seed)
    return rel


def fpr_samples(seed: int, n_samples: int, n_insertions: int):
    """inputs of `unittests 2 seed n_samples n_insertions m k_max` (unit_tests.c:243-297) and the filter seed"""
    R = np.empty(n_insertions, dtype=TUPLE)
    S = np.empty(n_samples, dtype=TUPLE)
    fseed = lib().orc_fpr_samples(seed, n_samples, n_insertions, _tp(R), _tp(S))
    return R, S, fseed


# ----------------------------------------------------------------------------------------------
# compiled-reference API
# ----------------------------------------------------------------------------------------------
def ref_join(R: np.ndarray, S: np.ndarray, algo: str = "PRO", nthreads: int = 8, bloom: bool = True, variant: int = 0,
             m: int = 1 << 28, k: int = 8, B: int = 1024, mat: bool = False) -> dict:
    res = RefResult()
    rc = ref(mat).refshim_join(algo.encode(), _tp(R), R.shape[0], _tp(S), S.shape[0], nthreads, int(bloom), variant,
                               m, k, B, C.byref(res))
    if rc != 0:
        raise RuntimeError(f"refshim_join rc={rc}")
    return {f: getattr(res, f) for f, _ in RefResult._fields_}


def ref_bloom_build(R: np.ndarray, variant: int, m: int, k: int, B: int, seed: int = 42) -> np.ndarray:
    bitmap = np.zeros(m // 8, dtype=np.uint8)
    ref().refshim_bloom_build(_tp(R), R.shape[0], variant, m, k, B, seed, bitmap.ctypes.data_as(C.c_void_p))
    return bitmap


def ref_bloom_count(bitmap: np.ndarray, S: np.ndarray, variant: int, m: int, k: int, B: int, seed: int = 42) -> int:
    return ref().refshim_bloom_count(bitmap.ctypes.data_as(C.c_void_p), _tp(S), S.shape[0], variant, m, k, B, seed,
                                     None)


def ref_hash(which: int, seed: int, key: int) -> int:
    return ref().refshim_hash(which, seed & 0xFFFFFFFF, int(np.int32(np.uint32(key & 0xFFFFFFFF))))


def ref_generate(kind: int, n: int, r: int, q: float = 1.0, zipf: float = 0.0, seed: int = 12345, nthreads: int = 8) -> np.ndarray:
    rel = np.empty(n, dtype=TUPLE)
    ref().refshim_generate(kind, _tp(rel), n, r, q, zipf, seed, nthreads)
    return rel
