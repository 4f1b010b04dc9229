"""ctypes binding of libhwbrj_cuda.so (the C ABI declared in include/hwbrj.h). No fallback: a missing library or
a missing GPU is an error."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HWBRJ_LIB") or os.path.join(HERE, "libhwbrj_cuda.so")  # HWBRJ_LIB: tuning variants

BASIC, BLOCKED = 0, 1


class TupleT(C.Structure):  # types.h:37-40
    _fields_ = [("key", C.c_int32), ("payload", C.c_int32)]


class RelationT(C.Structure):  # types.h:46-49
    _fields_ = [("tuples", C.c_void_p), ("num_tuples", C.c_uint64)]


class ThreadResultT(C.Structure):  # types.h:52-56
    _fields_ = [("nresults", C.c_int64), ("results", C.c_void_p), ("threadid", C.c_uint32)]


class ResultT(C.Structure):  # types.h:59-63
    _fields_ = [("totalresults", C.c_int64), ("resultlist", C.POINTER(ThreadResultT)), ("nthreads", C.c_int)]


class BloomFilterArgsT(C.Structure):  # bloom_filter.h:50-55
    _fields_ = [("variant", C.c_int), ("m", C.c_uint64), ("k", C.c_uint64), ("B", C.c_uint64)]


class StatsT(C.Structure):  # hwbrj_stats_t
    _fields_ = [("matches", C.c_int64), ("filtered", C.c_int64), ("checksum_pair", C.c_uint64),
                ("checksum_rpay", C.c_uint64), ("checksum_spay", C.c_uint64), ("checksum_key", C.c_uint64),
                ("ms_total", C.c_float), ("ms_memset", C.c_float), ("ms_build", C.c_float), ("ms_part_r", C.c_float),
                ("ms_probe", C.c_float), ("ms_part_s", C.c_float), ("ms_join", C.c_float), ("ms_h2d", C.c_float),
                ("ms_e2e", C.c_float), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("kernel_launches", C.c_int32), ("radix_bits", C.c_int32), ("range_passes", C.c_int32),
                ("n_gpus", C.c_int32), ("ms_comm", C.c_float), ("phase_split", C.c_int32), ("reserved", C.c_float * 2),
                ("owned_r", C.c_uint64), ("owned_s", C.c_uint64)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


# every symbol include/hwbrj.h declares (tests check that the library exports all of them)
EXPORTS = ["BPRO", "BRJ", "BPRH", "BPRHO", "PRO", "RJ", "PRH", "PRHO", "hwbrj_last_stats", "hwbrj_last_filtered",
           "hwbrj_last_checksum", "hwbrj_last_filter", "hwbrj_set_quiet", "hwbrj_set_radix_bits", "hwbrj_set_num_passes",
           "hwbrj_set_range_passes", "hwbrj_set_gpus", "hwbrj_set_overlap_h2d", "hwbrj_set_hash_partition", "hwbrj_version",
           "hwbrj_device_count", "hwbrj_check_args", "hwbrj_rel_upload", "hwbrj_rel_generate", "hwbrj_rel_download",
           "hwbrj_rel_size", "hwbrj_rel_free", "hwbrj_join_device", "hwbrj_join_device_async", "hwbrj_host_alloc",
           "hwbrj_host_free", "hwbrj_hash_many", "hwbrj_bloom_build", "hwbrj_bloom_probe", "hwbrj_fpr_count",
           "hwbrj_materialize_last", "hwbrj_materialize_last_device", "hwbrj_radix_partition",
           "hwbrj_dist_create", "hwbrj_dist_connect", "hwbrj_dist_join", "hwbrj_dist_join_async", "hwbrj_dist_filter",
           "hwbrj_dist_destroy",
           "hwbrj_set_stream", "hwbrj_reset_stream", "hwbrj_sync", "hwbrj_set_device", "hwbrj_rel_wrap", "hwbrj_rel_ptr",
           "hwbrj_rel_generate_shard", "hwbrj_owner_partition", "hwbrj_filter_build", "hwbrj_filter_or", "hwbrj_filter_probe"]
DIST_HANDLE_BYTES = 128  # HWBRJ_DIST_HANDLE_BYTES

_lib = None


def load():
    """Load the shared library (building nothing: use hwbloomradixjoin_b200.build / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -m hwbloomradixjoin_b200.build` (needs nvcc). "
                           "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    rp = C.POINTER(ResultT)
    relp = C.POINTER(RelationT)
    argp = C.POINTER(BloomFilterArgsT)
    for name in ("BPRO", "BRJ", "BPRH", "BPRHO"):
        f = getattr(L, name)
        f.restype = rp
        f.argtypes = [relp, relp, C.c_int, argp]
    for name in ("PRO", "RJ", "PRH", "PRHO"):
        f = getattr(L, name)
        f.restype = rp
        f.argtypes = [relp, relp, C.c_int]
    L.hwbrj_last_stats.argtypes = [C.POINTER(StatsT)]
    L.hwbrj_last_filtered.restype = C.c_int64
    L.hwbrj_last_checksum.restype = C.c_uint64
    L.hwbrj_last_filter.argtypes = [C.c_void_p, C.c_uint64]
    L.hwbrj_set_quiet.argtypes = [C.c_int]
    L.hwbrj_set_radix_bits.argtypes = [C.c_int]
    L.hwbrj_set_range_passes.argtypes = [C.c_int]
    L.hwbrj_set_num_passes.argtypes = [C.c_int]
    L.hwbrj_set_gpus.argtypes = [C.c_int]
    L.hwbrj_set_hash_partition.argtypes = [C.c_int]
    L.hwbrj_set_overlap_h2d.argtypes = [C.c_int]
    L.hwbrj_version.restype = C.c_char_p
    L.hwbrj_check_args.argtypes = [argp]
    L.hwbrj_rel_upload.restype = C.c_void_p
    L.hwbrj_rel_upload.argtypes = [C.c_void_p, C.c_uint64]
    L.hwbrj_rel_generate.restype = C.c_void_p
    L.hwbrj_rel_generate.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_double, C.c_uint64]
    L.hwbrj_rel_download.argtypes = [C.c_void_p, C.c_void_p]
    L.hwbrj_rel_size.restype = C.c_uint64
    L.hwbrj_rel_size.argtypes = [C.c_void_p]
    L.hwbrj_rel_free.argtypes = [C.c_void_p]
    L.hwbrj_join_device.argtypes = [C.c_void_p, C.c_void_p, argp, C.POINTER(StatsT)]
    L.hwbrj_join_device_async.argtypes = [C.c_void_p, C.c_void_p, argp, C.c_void_p]
    L.hwbrj_dist_create.restype = C.c_void_p
    L.hwbrj_dist_create.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
    L.hwbrj_dist_connect.argtypes = [C.c_void_p, C.c_char_p]
    L.hwbrj_dist_join.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, argp, C.c_uint64, C.POINTER(StatsT)]
    L.hwbrj_dist_join_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, argp, C.c_uint64, C.c_void_p]
    L.hwbrj_dist_filter.restype = C.c_void_p
    L.hwbrj_dist_filter.argtypes = [C.c_void_p]
    L.hwbrj_dist_destroy.argtypes = [C.c_void_p]
    L.hwbrj_host_alloc.restype = C.c_void_p
    L.hwbrj_host_alloc.argtypes = [C.c_uint64]
    L.hwbrj_host_free.argtypes = [C.c_void_p]
    L.hwbrj_hash_many.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
    L.hwbrj_bloom_build.argtypes = [C.c_void_p, C.c_uint64, argp, C.c_uint32, C.c_void_p]
    L.hwbrj_bloom_probe.restype = C.c_int64
    L.hwbrj_bloom_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, argp, C.c_uint32, C.c_void_p]
    L.hwbrj_materialize_last.restype = C.c_int64
    L.hwbrj_materialize_last.argtypes = [C.c_void_p, C.c_uint64]
    L.hwbrj_materialize_last_device.restype = C.c_int64
    L.hwbrj_materialize_last_device.argtypes = [C.c_void_p, C.c_uint64]
    L.hwbrj_fpr_count.restype = C.c_int64
    L.hwbrj_fpr_count.argtypes = [C.c_void_p, C.c_void_p, argp, C.c_uint32]
    L.hwbrj_radix_partition.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    L.hwbrj_set_stream.argtypes = [C.c_void_p]
    L.hwbrj_set_device.argtypes = [C.c_int]
    L.hwbrj_rel_wrap.restype = C.c_void_p
    L.hwbrj_rel_wrap.argtypes = [C.c_void_p, C.c_uint64]
    L.hwbrj_rel_ptr.restype = C.c_void_p
    L.hwbrj_rel_ptr.argtypes = [C.c_void_p]
    L.hwbrj_rel_generate_shard.restype = C.c_void_p
    L.hwbrj_rel_generate_shard.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_double, C.c_uint64, C.c_uint64,
                                           C.c_uint64]
    L.hwbrj_owner_partition.argtypes = [C.c_void_p, C.c_int, argp, C.c_void_p, C.c_void_p]
    L.hwbrj_filter_build.argtypes = [C.c_void_p, argp, C.c_void_p, C.c_int]
    L.hwbrj_filter_or.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.hwbrj_filter_probe.restype = C.c_int64
    L.hwbrj_filter_probe.argtypes = [C.c_void_p, C.c_void_p, argp, C.c_void_p]
    _lib = L
    return L
