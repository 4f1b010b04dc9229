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
    float    ms_build;      /* R: histogram (+ Bloom insert), filter slices built from the partitioned R */
    float    ms_part_r;     /* R: offsets + scatter passes */
    float    ms_probe;      /* S: Bloom probe + compaction (the K2 launches) */
    float    ms_part_s;     /* S: histogram + scan + scatter passes */
    float    ms_join;       /* work list + per-partition build + probe */
    float    ms_h2d;        /* host->device copies (host-buffer entry points only) */
    float    ms_e2e;        /* wall clock of the whole host-buffer call, incl. copies */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    int32_t  kernel_launches; /* kernels launched inside the timed region */
    int32_t  radix_bits;      /* total radix bits used */
    int32_t  range_passes;    /* filter range passes used by insert/probe */
    int32_t  n_gpus;
    float    ms_comm;         /* reserved */
    int32_t  phase_split;     /* 1: the five phase times are disjoint pieces of one stream (one GPU); 0: several GPUs, the
                                 level-2 pulls overlap the slice build / the join, see DESIGN.md section 7 */
    float    reserved[2];
    uint64_t owned_r;         /* R tuples this GPU owns after the routing (host-buffer multi-GPU calls: the maximum) */
    uint64_t owned_s;         /* probe tuples (filter survivors) this GPU owns: the per-GPU load of the join phase */
} hwbrj_stats_t;

/* results of the most recent join in this process */
int      hwbrj_last_stats(hwbrj_stats_t * out);
int64_t  hwbrj_last_filtered(void);
uint64_t hwbrj_last_checksum(void); /* checksum_pair */
/* the Bloom filter bitmap the most recent Bloom join built on the device (first nbytes = m/8) */
int      hwbrj_last_filter(unsigned char * bitmap_out, uint64_t nbytes);

void hwbrj_set_quiet(int quiet);    /* 1: suppress the reference-style stdout lines */
/* tuning knobs (also read from env HWBRJ_RADIX_BITS / HWBRJ_NUM_PASSES / HWBRJ_RANGE_PASSES); 0 = automatic.
 * radix_bits and num_passes are the runtime form of the reference's compile-time NUM_RADIX_BITS / NUM_PASSES
 * (prj_params.h:15-22, swept by measurements/run.py:205-269): total partition bits (<= 14) and 1 or 2 scatter passes.
 * One pass handles at most 7 bits, so num_passes = 1 caps the fan-out at 2^7 partitions; partitions that do not fit one
 * shared-memory table are joined in several table rounds (results are independent of both knobs). */
void hwbrj_set_radix_bits(int bits);
void hwbrj_set_num_passes(int passes);
void hwbrj_set_range_passes(int passes);
/* Host-buffer entry points (BPRO, PRO, ...): shard the relations over the first n GPUs of this process (contiguous
 * chunks, like the reference's worker threads, parallel_radix_join_bloom.c:1646-1672) and join them together over
 * NVLink peer memory. n must be a power of two <= the device count; the GPUs need peer access. Returns 0 on success. */
int  hwbrj_set_gpus(int n);
/* host-buffer Bloom joins: upload S in chunks on a copy stream and probe each chunk as it lands (hides the join
 * under the PCIe copy; TOTAL-TIME-USECS then includes waiting for the copies). Off by default. */
void hwbrj_set_overlap_h2d(int on);
/* partition the join on the filter-slice index and build the filter in shared memory (BASIC k<=1, BLOCKED any k):
 * 0 never, 1 when the filter exceeds 32 MiB or several GPUs join (default), 2 whenever the slices fit */
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

/* enqueue-only variant (no events, no host synchronisation: can be captured in a CUDA graph). Leaves {matches,
 * checksum_pair, checksum_rpay, checksum_spay, checksum_key, filtered, error flags, 0} as eight uint64 in d_out8 (device).
 * Returns the number of kernels enqueued, < 0 on error. */
int hwbrj_join_device_async(const hwbrj_rel_t * R, const hwbrj_rel_t * S, const bloom_filter_args_t * args, void * d_out8);

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

/* ---- the multi-GPU join: one rank per GPU -------------------------------------------------------------
 * The join shards on the HIGH bits of the partition id: rank g owns partitions [g*P/G, (g+1)*P/G) -- for a BASIC (k <= 1)
 * or BLOCKED filter these are the keys whose filter bits lie in the g-th 1/G slice of the filter. Per rank and join:
 *   histogram of the local R chunk -> rows all-gathered over NVLink -> every rank derives the layout of every rank's
 *   level-1 output -> level-1 scatter of the local chunk into a peer-mapped staging buffer (purely local) -> the level-2
 *   scatter of each owner PULLS the segments of its bins from all staging buffers with bulk loads over NVLink (fused
 *   all-to-all + partition pass: every tuple crosses the fabric once, no send/receive buffers, no remote atomics) ->
 *   filter-slice build in shared memory, each slice stored into EVERY rank's filter (fused build + all-gather) -> probe of
 *   the local S chunk against the replicated filter -> only the survivors are partitioned and pulled the same way ->
 *   per-partition build + probe on the owner -> result words all-gathered and summed.
 * All exchanges are peer-memory loads/stores issued by the kernels themselves; ranks meet at device-side barriers.
 * A rank is a process with one GPU (handles travel through the launcher's channel, e.g. torch.distributed.all_gather or
 * MPI) or one of several GPUs driven by one process (hwbrj_set_gpus does all of this internally). */
typedef struct hwbrj_dist hwbrj_dist_t;
#define HWBRJ_DIST_HANDLE_BYTES 128
/* allocate this rank's symmetric block on the CURRENT device: staging buffers of cap_r / cap_s tuples and room for a filter
 * of max_filter_bytes; writes the handle peers need (HWBRJ_DIST_HANDLE_BYTES) to handle_out. world: power of two <= 16.
 * cap_r (cap_s) bounds both a rank's input chunk of R (of S) and the tuples of R (the filter survivors) one rank may own;
 * a join that exceeds them fails with -2 on every rank (nothing is overrun). The same values on every rank. */
hwbrj_dist_t * hwbrj_dist_create(int rank, int world, uint64_t cap_r, uint64_t cap_s, uint64_t max_filter_bytes,
                                 void * handle_out);
/* map the peers; all_handles = the world handles in rank order. Returns 0 on success. */
int  hwbrj_dist_connect(hwbrj_dist_t * d, const void * all_handles);
/* collective: every rank calls it with its chunks of R and S. r_total = |R| over all ranks. The scalars of `out`
 * (matches, filtered, checksums) are the global ones, identical on every rank; timings are this rank's.
 * Returns 0, -2 if a receive buffer was too small or a peer did not arrive (results invalid). */
int  hwbrj_dist_join(hwbrj_dist_t * d, const hwbrj_rel_t * R, const hwbrj_rel_t * S, const bloom_filter_args_t * args,
                     uint64_t r_total, hwbrj_stats_t * out);
/* enqueue-only variant on the stream of hwbrj_set_stream() (capturable in a CUDA graph): d_out8 as in
 * hwbrj_join_device_async, already summed over the ranks */
int  hwbrj_dist_join_async(hwbrj_dist_t * d, const hwbrj_rel_t * R, const hwbrj_rel_t * S, const bloom_filter_args_t * args,
                           uint64_t r_total, void * d_out8);
void * hwbrj_dist_filter(hwbrj_dist_t * d); /* this rank's copy of the replicated filter (device pointer) */
void hwbrj_dist_destroy(hwbrj_dist_t * d);

/* ---- plumbing shared by the multi-GPU paths ---------------------------------------------------------------------- */
void hwbrj_set_stream(void * cuda_stream); /* run on the caller's stream (0 = the legacy default stream) */
void hwbrj_reset_stream(void);             /* back to the library's own stream */
int  hwbrj_sync(void);
int  hwbrj_set_device(int device);           /* cudaSetDevice: the current device selects the library's context */
/* non-owning view of device memory; the pointer must be 16-byte aligned (NULL otherwise); an odd tuple count makes the
 * kernels read (not use) 8 bytes past the last tuple, which stay inside the allocation's last page */
hwbrj_rel_t * hwbrj_rel_wrap(void * device_tuples, uint64_t n);
void *        hwbrj_rel_ptr(const hwbrj_rel_t * rel);
/* positions [begin, begin+count) of the global generated relation (a rank's contiguous input chunk, the GPU
 * analogue of the per-thread chunks of parallel_radix_join_bloom.c:1646-1672) */
hwbrj_rel_t * hwbrj_rel_generate_shard(int kind, uint64_t n, uint64_t r, double q, uint64_t seed, uint64_t begin,
                                       uint64_t count);

/* ---- building blocks of the NCCL reference path (hwbloomradixjoin_b200/dist.py: dist_join), which exchanges with
 * torch.distributed all-to-all / all-gather and is what the peer-memory join above is validated against ------------- */
/* group the tuples by owner GPU into d_out (n tuples) and return the per-owner counts (host array of `world`).
 * Owner = the GPU holding the filter slice of the key's first bit (slice_args BASIC k<=1 or BLOCKED), else the
 * top bits of crapwow(42,key); equal keys always share an owner, so owners join independently. */
int hwbrj_owner_partition(const hwbrj_rel_t * in, int world, const bloom_filter_args_t * slice_args, void * d_out,
                          uint64_t * counts_out);
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
