/*
 * hwbrj.h -- C ABI of libhwbrj_cuda.so: the B200 (sm_100a) drop-in for the reference's
 * Bloom-filter radix hash join path.
 *
 * Part 1 is exactly what the reference's call site binds (main.c:277-282,331-339,473-478):
 * the same type layouts (types.h:22-63, bloom_filter.h:10,50-55) and the same entry points
 * (parallel_radix_join_bloom.h:34-86, parallel_radix_join.h:33-82). Part 2 are extensions the
 * reference has no field for (filtered count, checksums, device timings, device-resident
 * relations, multi-GPU bootstrap). Everything is extern "C" with plain pointers and sizes.
 *
 * There is no CPU fallback: every entry point runs CUDA kernels on the current device or fails
 * loudly ([ERROR] ... + exit(EXIT_FAILURE), the reference's own error style,
 * parallel_radix_join_bloom.c:64-71).
 */
#ifndef HWBRJ_H
#define HWBRJ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------- */
/* Part 1: the reference's ABI                                                                  */
/* ------------------------------------------------------------------------------------------- */

#ifndef TYPES_H /* do not clash when a host also includes the reference's types.h */
#define TYPES_H
typedef int32_t intkey_t; /* types.h:26-27 (KEY_8B off: 8-byte tuples) */
typedef int32_t value_t;

typedef struct tuple_t        tuple_t;
typedef struct relation_t     relation_t;
typedef struct result_t       result_t;
typedef struct threadresult_t threadresult_t;

struct tuple_t { /* types.h:37-40 */
    intkey_t key;
    value_t  payload;
};

struct relation_t { /* types.h:46-49 */
    tuple_t * tuples;
    uint64_t  num_tuples;
};

struct threadresult_t { /* types.h:52-56 */
    int64_t  nresults;
    void *   results;
    uint32_t threadid;
};

struct result_t { /* types.h:59-63 */
    int64_t          totalresults;
    threadresult_t * resultlist;
    int              nthreads;
};
#endif /* TYPES_H */

#ifndef BLOOM_FILTER_H
#define BLOOM_FILTER_H
typedef enum { BASIC, BLOCKED } bloom_filter_variant_t; /* bloom_filter.h:10 */

typedef struct bloom_filter_args_t { /* bloom_filter.h:50-55 */
    bloom_filter_variant_t variant;
    uint64_t               m; /* filter size in bits, power of two */
    uint64_t               k; /* bits set per key */
    uint64_t               B; /* block size in bits (BLOCKED), power of two, m % B == 0 */
} bloom_filter_args_t;
#endif /* BLOOM_FILTER_H */

/*
 * Join entry points. Same signatures and semantics as the reference: blocking; the caller owns
 * relR/relS; the returned result_t is malloc()ed and the caller free()s it (main.c:490);
 * totalresults = number of (r,s) pairs with equal keys. The Bloom variants print
 * "S-tuples after filter: N" (BRJ does not, as in the reference) and all print the
 * print_timing() block (parallel_radix_join_bloom.c:1510-1547) unless hwbrj_set_quiet(1).
 * Unlike the reference the inputs are NOT modified and need no RELATION_PADDING slack.
 * `nthreads` is accepted for signature compatibility; it is echoed in result_t.nthreads.
 * BPRH/BPRHO/PRH/PRHO give the same results as BPRO/PRO and alias the same pipeline.
 */
result_t * BPRO(relation_t * relR, relation_t * relS, int nthreads, bloom_filter_args_t * args);  /* parallel_radix_join_bloom.c:1782 */
result_t * BRJ(relation_t * relR, relation_t * relS, int nthreads, bloom_filter_args_t * args);   /* :1808 */
result_t * BPRH(relation_t * relR, relation_t * relS, int nthreads, bloom_filter_args_t * args);  /* :1791 */
result_t * BPRHO(relation_t * relR, relation_t * relS, int nthreads, bloom_filter_args_t * args); /* :1799 */
result_t * PRO(relation_t * relR, relation_t * relS, int nthreads);                               /* parallel_radix_join.c:1697 */
result_t * RJ(relation_t * relR, relation_t * relS, int nthreads);                                /* :1718 */
result_t * PRH(relation_t * relR, relation_t * relS, int nthreads);                               /* :1704 */
result_t * PRHO(relation_t * relR, relation_t * relS, int nthreads);                              /* :1711 */

/* ------------------------------------------------------------------------------------------- */
/* Part 2: extensions                                                                           */
/* ------------------------------------------------------------------------------------------- */

/* pair mixer used by checksum_pair: splitmix64 finaliser over (R.payload << 32 | S.payload) */
static inline uint64_t
hwbrj_mix64(uint32_t rpay, uint32_t spay)
{
    uint64_t z = ((uint64_t) rpay << 32) | (uint64_t) spay;
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

typedef struct hwbrj_stats_t {
    int64_t  matches;       /* = result_t.totalresults */
    int64_t  filtered;      /* S tuples passing the filter; -1 when no filter was used */
    uint64_t checksum_pair; /* sum over output pairs of hwbrj_mix64(R.payload,S.payload) mod 2^64 */
    uint64_t checksum_rpay; /* sum over output pairs of (uint32)R.payload */
    uint64_t checksum_spay; /* sum over output pairs of (uint32)S.payload */
    uint64_t checksum_key;  /* sum over output pairs of (uint32)S.key */
    /* device timings, CUDA events on the library's stream, milliseconds */
    float    ms_total;      /* whole join, inputs resident, filter zeroing excluded (reference timing region) */
    float    ms_memset;     /* zero-fill of the filter and scratch (excluded from ms_total, as :1583 is) */
    float    ms_build;      /* R: Bloom insert + histogram */
    float    ms_part_r;     /* R: scatter passes */
    float    ms_probe;      /* S: Bloom probe + compaction (the K2 launches) */
    float    ms_part_s;     /* S: histogram + scan + scatter passes */
    float    ms_join;       /* per-partition build + probe */
    float    ms_h2d;        /* host->device copies (host-buffer entry points only) */
    float    ms_e2e;        /* wall clock of the whole host-buffer call, incl. copies */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    int32_t  kernel_launches; /* kernels launched inside the timed region */
    int32_t  radix_bits;      /* total radix bits used */
    int32_t  range_passes;    /* filter range passes used by insert/probe */
    int32_t  n_gpus;
    float    ms_comm;         /* multi-GPU: time in collectives */
    float    reserved[3];
} hwbrj_stats_t;

/* results of the most recent join in this process */
int      hwbrj_last_stats(hwbrj_stats_t * out);
int64_t  hwbrj_last_filtered(void);
uint64_t hwbrj_last_checksum(void); /* checksum_pair */
/* the Bloom filter bitmap the most recent Bloom join built on the device (first nbytes = m/8) */
int      hwbrj_last_filter(unsigned char * bitmap_out, uint64_t nbytes);

void hwbrj_set_quiet(int quiet);    /* 1: suppress the reference-style stdout lines */
/* tuning knobs (also read from env HWBRJ_RADIX_BITS / HWBRJ_RANGE_PASSES); 0 = automatic */
void hwbrj_set_radix_bits(int bits);
void hwbrj_set_range_passes(int passes);
/* host-buffer Bloom joins: upload S in chunks on a copy stream and probe each chunk as it lands (hides the join
 * under the PCIe copy; TOTAL-TIME-USECS then includes waiting for the copies). Off by default. */
void hwbrj_set_overlap_h2d(int on);
/* partition the join on the filter-slice index and build the filter in shared memory (BASIC k<=1):
 * 0 never, 1 when the filter exceeds 32 MiB (default), 2 whenever the slices fit */
void hwbrj_set_hash_partition(int mode);
const char * hwbrj_version(void);
int  hwbrj_device_count(void);

/* argument validation with the reference's rules (bloom_filter.c:26-34); returns 0 when valid,
 * else prints the reference's message and returns non-zero (the reference exits) */
int hwbrj_check_args(const bloom_filter_args_t * args);

/* ---- device-resident relations (inputs already in HBM: the timing region of `value`) ------- */
typedef struct hwbrj_rel hwbrj_rel_t; /* opaque device relation */

hwbrj_rel_t * hwbrj_rel_upload(const tuple_t * tuples, uint64_t n);
/* on-device generator with the reference generator's key multiset (generator.c:162-195,341-387):
 * kind 0: R = keys 1..n (payload = position); kind 1: S = FK relation over threshold r with selectivity q.
 * Positions are permuted by a seeded bijection (the reference's own shuffle is time-seeded).
 * kind 2: S = n Zipf-distributed foreign keys over the alphabet 1..r with exponent q (create_relation_zipf,
 * generator.c:659-676 / genzipf.c: permuted alphabet, cumulated-density table, binary search per tuple; the uniform
 * numbers come from a counter-based generator instead of glibc rand(), so the key ARRAY differs from a reference run). */
hwbrj_rel_t * hwbrj_rel_generate(int kind, uint64_t n, uint64_t r, double q, uint64_t seed);
int      hwbrj_rel_download(const hwbrj_rel_t * rel, tuple_t * out);
uint64_t hwbrj_rel_size(const hwbrj_rel_t * rel);
void     hwbrj_rel_free(hwbrj_rel_t * rel);

/* join of device-resident relations; args == NULL -> plain radix join (PRO). Returns 0 on success. */
int hwbrj_join_device(const hwbrj_rel_t * R, const hwbrj_rel_t * S, const bloom_filter_args_t * args,
                      hwbrj_stats_t * out);

/* enqueue-only variant (no events, no host synchronisation: can be captured in a CUDA graph together with the
 * collectives around it). Leaves {matches, checksum_pair, checksum_rpay, checksum_spay, checksum_key, filtered}
 * as six uint64 in d_out6 (device). Returns the number of kernels enqueued, < 0 on error. */
int hwbrj_join_device_async(const hwbrj_rel_t * R, const hwbrj_rel_t * S, const bloom_filter_args_t * args, void * d_out6);
/* Partition the build side of a FILTER-LESS join ahead of time: the histogram and scatter passes of R (the R half of
 * :808-849) are enqueued on a library-owned side stream forked from the current stream, so that they overlap whatever the
 * caller enqueues next (the multi-GPU join: filter all-gather and S probe). The next hwbrj_join_device[_async](R, S, NULL)
 * on the same relation joins the side stream and skips those passes. Returns the number of kernel launches, < 0 on
 * error. Experimental (HWBRJ_DIST_OVERLAP_R=1 in hwbloomradixjoin_b200.dist). */
int hwbrj_join_prepare_r(const hwbrj_rel_t * R);

/* pinned host buffers for callers that want full-speed PCIe copies (bench e2e leg) */
void * hwbrj_host_alloc(uint64_t bytes);
void   hwbrj_host_free(void * p);

/* ---- building blocks, exposed for parity tests --------------------------------------------- */
/* which: 0 crc,1 FNV,2 crapwow,3 Coffin,4 MurmurOAAT,5 JenkinsOAAT,6 Spooky,7 KR_v2,8 DJB2,9 x17 (hash.h:12-40) */
int hwbrj_hash_many(int which, uint32_t seed, const int32_t * keys, uint64_t n, uint32_t * out);
/* build the filter on the GPU from R's keys and copy the m/8-byte bitmap back (bloom_filter.c:74-89,126-132) */
int hwbrj_bloom_build(const tuple_t * R, uint64_t nR, const bloom_filter_args_t * args, uint32_t seed,
                      unsigned char * bitmap_out);
/* probe S against a given bitmap on the GPU; returns the pass count, optionally the survivors
 * (any order) and their number (bloom_filter.c:93-111,135-141) */
int64_t hwbrj_bloom_probe(const unsigned char * bitmap, const tuple_t * S, uint64_t nS,
                          const bloom_filter_args_t * args, uint32_t seed, tuple_t * survivors_out);
/* Result materialisation (the reference's -DJOIN_RESULT_MATERIALIZE, parallel_radix_join_bloom.c:307-312): writes the
 * output of the most recent join of this process as tuples {key = R.payload, payload = S.payload}, in any order.
 * Returns the number of output pairs; if it exceeds `capacity` only `capacity` pairs were written (retry larger). */
int64_t hwbrj_materialize_last(tuple_t * pairs_out /* host */, uint64_t capacity);
int64_t hwbrj_materialize_last_device(void * d_pairs /* device */, uint64_t capacity);
/* device analogue of the reference's FPR measurement (test_bloom_fpr, unit_tests.c:191-241): build a filter with
 * `seed` from the device-resident R, probe the device-resident S, return the number of passing S keys */
int64_t hwbrj_fpr_count(const hwbrj_rel_t * R, const hwbrj_rel_t * S, const bloom_filter_args_t * args, uint32_t seed);
/* radix-partition a relation on the GPU with the pipeline's own kernels: out receives the tuples
 * grouped by (key & (2^bits-1)) in increasing partition order, offsets (2^bits+1 entries) the
 * partition boundaries (parallel_radix_join_bloom.c:574-608,759-852 equivalent) */
int hwbrj_radix_partition(const tuple_t * in, uint64_t n, int bits, tuple_t * out, uint64_t * offsets);

/* ---- multi-GPU building blocks ------------------------------------------------------------
 * One process per GPU; the host side (hwbloomradixjoin_b200/dist.py) runs these between torch.distributed / NCCL
 * collectives. All take device pointers and run on the stream given to hwbrj_set_stream(). */
void hwbrj_set_stream(void * cuda_stream); /* run on the caller's stream (0 = the legacy default stream) */
void hwbrj_reset_stream(void);             /* back to the library's own stream */
int  hwbrj_sync(void);
int  hwbrj_set_device(int device);           /* before the first call: one process per GPU */
/* non-owning view of device memory; the pointer must be 16-byte aligned (NULL otherwise); an odd tuple count makes the
 * kernels read (not use) 8 bytes past the last tuple, which stay inside the allocation's last page */
hwbrj_rel_t * hwbrj_rel_wrap(void * device_tuples, uint64_t n);
/* same, but the real tuple count is a device-resident uint64 (written by an earlier kernel): `capacity` bounds it,
 * `expected` sizes the radix fan-out. Lets a pipeline run without host round trips. */
hwbrj_rel_t * hwbrj_rel_wrap_counted(void * device_tuples, uint64_t capacity, const void * d_count, uint64_t expected);
void *        hwbrj_rel_ptr(const hwbrj_rel_t * rel);
/* positions [begin, begin+count) of the global generated relation (a rank's contiguous input chunk, the GPU
 * analogue of the per-thread chunks of parallel_radix_join_bloom.c:1646-1672) */
hwbrj_rel_t * hwbrj_rel_generate_shard(int kind, uint64_t n, uint64_t r, double q, uint64_t seed, uint64_t begin,
                                       uint64_t count);
/* group the tuples by owner GPU into d_out (n tuples) and return the per-owner counts (host array of `world`).
 * Owner = the GPU holding the filter slice of the key's first bit (slice_args BASIC k<=1 or BLOCKED), else the
 * top bits of crapwow(42,key); equal keys always share an owner, so owners join independently. */
int hwbrj_owner_partition(const hwbrj_rel_t * in, int world, const bloom_filter_args_t * slice_args, void * d_out,
                          uint64_t * counts_out);
/* peer memory: buffers other ranks of the NVLink domain write into. The 64-byte handles travel through the host's
 * own channel (e.g. torch.distributed.all_gather) and are opened by every peer. */
#define HWBRJ_IPC_HANDLE_BYTES 64
void * hwbrj_symm_alloc(uint64_t bytes); /* zero-initialised device memory that can be exported */
void   hwbrj_symm_free(void * p);
int    hwbrj_ipc_export(void * p, void * handle_out /* HWBRJ_IPC_HANDLE_BYTES */);
void * hwbrj_ipc_open(const void * handle);
int    hwbrj_ipc_close(void * p);
/* fused partition-by-owner + all-to-all: every tuple of `in` is stored straight into its owner's receive buffer
 * (peer_bufs[g], capacity_tuples each) at a position claimed from the owner's cursor (peer_cursors[g], uint64) with
 * a system-scope atomic over NVLink. A claim that does not fit sets *d_overflow_flag (uint32) and is dropped.
 * The caller separates routing from consumption with a stream-ordered barrier across ranks. */
int hwbrj_route_peer(const hwbrj_rel_t * in, int world, const bloom_filter_args_t * slice_args, void * const * peer_bufs,
                     void * const * peer_cursors, uint64_t capacity_tuples, void * d_overflow_flag);
/* hwbrj_filter_probe without the host round trip: the survivor count is left in *d_count_out (device uint64) */
int hwbrj_filter_probe_async(const void * d_filter, const hwbrj_rel_t * S, const bloom_filter_args_t * args,
                             void * d_out, void * d_count_out);
/* insert R's keys into the full-size filter at d_filter (m/8 bytes, device) */
int hwbrj_filter_build(const hwbrj_rel_t * R, const bloom_filter_args_t * args, void * d_filter, int zero_first);
/* dst |= src over nbytes (multiple of 16): combines partial filters (NCCL has no bitwise-OR reduction) */
int hwbrj_filter_or(void * d_dst, const void * d_src, uint64_t nbytes);
/* probe S against the device filter; survivors (any order) go to d_out (capacity |S| tuples); returns the count */
int64_t hwbrj_filter_probe(const void * d_filter, const hwbrj_rel_t * S, const bloom_filter_args_t * args,
                           void * d_out);

#ifdef __cplusplus
}
#endif
#endif /* HWBRJ_H */
