"""hwbloomradixjoin_b200 -- B200 (sm_100a) drop-in for the Bloom-filter radix hash join path of
Briimbo/HwBloomRadixJoin (src/parallel_radix_join_bloom.c). See DESIGN.md and include/hwbrj.h."""
from .api import (ALGOS, BASIC, BLOCKED, BPRH, BPRHO, BPRO, BRJ, PRH, PRHO, PRO, RJ, TUPLE, BloomFilterArgs,  # noqa: F401
                  DeviceRelation, JoinResult, bloom_build, bloom_probe, device_count, fpr_count, hash_many, join_device, last_filter, materialize_last,
                  radix_partition, run, set_gpus, set_quiet, set_hash_partition, set_num_passes, set_radix_bits, set_range_passes)

__version__ = "0.2.0"
