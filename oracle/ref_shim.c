/*
 * ref_shim.c -- TEST INFRASTRUCTURE (oracle side). Not part of the product path.
 *
 * Thin C entry points around the UNMODIFIED reference sources. This file is compiled together
 * with /root/reference/src/{parallel_radix_join_bloom,parallel_radix_join,bloom_filter,hash,
 * spooky,generator,genzipf,cpu_mapping,perf_counters}.c (see oracle/Makefile) into
 * oracle/_ref/libref.so (default flags) and oracle/_ref/libref_mat.so
 * (-DJOIN_RESULT_MATERIALIZE, so that the output pairs can be folded into a checksum).
 * No reference source is copied into this repository; only the built objects live in the
 * git-ignored oracle/_ref/.
 *
 * What it wraps (reference file:line):
 *   BPRO/BRJ/BPRH/BPRHO   parallel_radix_join_bloom.c:1782,1808,1791,1799
 *   PRO/RJ/PRH/PRHO       parallel_radix_join.c:1697,1718
 *   "S-tuples after filter" is only printed (parallel_radix_join_bloom.c:1253), so stdout is
 *   captured around the call and parsed, exactly like measurements/run.py:109-129 does.
 *   bloom_filter_create/add/contains   bloom_filter.c:144,74,93
 *   hash_*                hash.c:6-140
 *   parallel_create_relation / create_relation_zipf   generator.c:305,659
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "types.h"
#include "bloom_filter.h"
#include "generator.h"
#include "hash.h"
#include "parallel_radix_join.h"
#include "parallel_radix_join_bloom.h"
#include "prj_params.h"
#ifdef JOIN_RESULT_MATERIALIZE
#include "tuple_buffer.h"
#endif

void * alloc_aligned(size_t size); /* generator.c:52, not declared in generator.h */
extern int numalocalize; /* generator.c:46 */
extern int nthreads;     /* generator.c:47 */

typedef struct refshim_result_t {
    int64_t  matches;        /* result_t.totalresults */
    int64_t  filtered;       /* parsed "S-tuples after filter: N" (-1 if not printed) */
    double   total_usecs;    /* parsed TOTAL-TIME-USECS */
    double   part_usecs;     /* parsed PARTITION-TIME-USECS */
    double   join_usecs;     /* parsed JOIN-TIME-USECS */
    uint64_t checksum_pair;  /* sum mix64(R.payload,S.payload) over materialised pairs (MAT build) */
    uint64_t checksum_rpay;  /* sum (uint32)R.payload over materialised pairs (MAT build) */
    uint64_t checksum_spay;  /* sum (uint32)S.payload over materialised pairs (MAT build) */
    int32_t  materialized;   /* 1 when built with -DJOIN_RESULT_MATERIALIZE */
    int32_t  radix_bits;     /* NUM_RADIX_BITS the reference was compiled with */
} refshim_result_t;

/* the repo-wide pair mixer (defined in include/hwbrj.h as well): splitmix64 finaliser */
static inline uint64_t
mix64(uint32_t rpay, uint32_t spay)
{
    uint64_t z = ((uint64_t) rpay << 32) | (uint64_t) spay;
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

int
refshim_relation_padding(void)
{
    return (int) RELATION_PADDING;
}

int
refshim_is_materialized(void)
{
#ifdef JOIN_RESULT_MATERIALIZE
    return 1;
#else
    return 0;
#endif
}

static double
parse_after(const char * buf, const char * tag, int field)
{
    const char * p = strstr(buf, tag);
    if (!p) return -1.0;
    p = strchr(p, '\n');
    if (!p) return -1.0;
    p++;
    double v[3] = {-1, -1, -1};
    sscanf(p, "%lf %lf %lf", &v[0], &v[1], &v[2]);
    return v[field];
}

/**
 * Runs one of the reference joins on COPIES of the given arrays (the reference clobbers its
 * inputs and needs RELATION_PADDING slack, prj_params.h:88-92 / generator.c:27).
 * algo: "PRO","RJ","PRH","PRHO". bloom_enable selects the B* twin.
 */
int
refshim_join(const char * algo, const tuple_t * R, uint64_t nR, const tuple_t * S,
             uint64_t nS, int nthr, int bloom_enable, int variant, uint64_t m, uint64_t k,
             uint64_t B, refshim_result_t * out)
{
    relation_t relR, relS;
    relR.num_tuples = nR;
    relS.num_tuples = nS;
    numalocalize    = 0;
    nthreads        = nthr;
    relR.tuples = (tuple_t *) alloc_aligned(nR * sizeof(tuple_t) + RELATION_PADDING);
    relS.tuples = (tuple_t *) alloc_aligned(nS * sizeof(tuple_t) + RELATION_PADDING);
    if (!relR.tuples || !relS.tuples) return -1;
    memcpy(relR.tuples, R, nR * sizeof(tuple_t));
    memcpy(relS.tuples, S, nS * sizeof(tuple_t));

    bloom_filter_args_t args;
    args.variant = variant ? BLOCKED : BASIC;
    args.m       = m;
    args.k       = k;
    args.B       = B;

    /* capture stdout (the reference only prints `filtered` and its timings) */
    fflush(stdout);
    char tmpl[] = "/tmp/refshim_XXXXXX";
    int  tfd    = mkstemp(tmpl);
    int  saved  = dup(1);
    dup2(tfd, 1);

    result_t * res = NULL;
    if (bloom_enable) {
        if (!strcmp(algo, "PRO")) res = BPRO(&relR, &relS, nthr, &args);
        else if (!strcmp(algo, "RJ")) res = BRJ(&relR, &relS, nthr, &args);
        else if (!strcmp(algo, "PRH")) res = BPRH(&relR, &relS, nthr, &args);
        else if (!strcmp(algo, "PRHO")) res = BPRHO(&relR, &relS, nthr, &args);
    } else {
        if (!strcmp(algo, "PRO")) res = PRO(&relR, &relS, nthr);
        else if (!strcmp(algo, "RJ")) res = RJ(&relR, &relS, nthr);
        else if (!strcmp(algo, "PRH")) res = PRH(&relR, &relS, nthr);
        else if (!strcmp(algo, "PRHO")) res = PRHO(&relR, &relS, nthr);
    }

    fflush(stdout);
    dup2(saved, 1);
    close(saved);
    off_t len = lseek(tfd, 0, SEEK_END);
    char * buf = (char *) calloc((size_t) len + 1, 1);
    lseek(tfd, 0, SEEK_SET);
    if (len > 0 && read(tfd, buf, (size_t) len) < 0) buf[0] = 0;
    close(tfd);
    unlink(tmpl);

    memset(out, 0, sizeof(*out));
    out->radix_bits = NUM_RADIX_BITS;
    out->filtered   = -1;
    if (!res) {
        free(buf);
        free(relR.tuples);
        free(relS.tuples);
        return -2;
    }
    out->matches = res->totalresults;
    const char * f = strstr(buf, "S-tuples after filter:");
    if (f) out->filtered = atoll(f + strlen("S-tuples after filter:"));
    out->total_usecs = parse_after(buf, "TOTAL-TIME-USECS", 0);
    out->part_usecs  = parse_after(buf, "PARTITION-TIME-USECS", 0);
    out->join_usecs  = parse_after(buf, "PARTITION-TIME-USECS", 2);

#ifdef JOIN_RESULT_MATERIALIZE
    out->materialized = 1;
    for (int t = 0; t < res->nthreads; t++) {
        threadresult_t *       tr = &res->resultlist[t];
        chainedtuplebuffer_t * cb = (chainedtuplebuffer_t *) tr->results;
        if (!cb) continue;
        /* newest buffer first: it holds writepos tuples, every older one is full
           (tuple_buffer.h:92-108) */
        tuplebuffer_t * b   = cb->buf;
        uint32_t        cnt = cb->writepos;
        while (b) {
            for (uint32_t i = 0; i < cnt; i++) {
                uint32_t rp = (uint32_t) b->tuples[i].key;     /* R.payload, :310 */
                uint32_t sp = (uint32_t) b->tuples[i].payload; /* S.payload, :311 */
                out->checksum_pair += mix64(rp, sp);
                out->checksum_rpay += rp;
                out->checksum_spay += sp;
            }
            tuplebuffer_t * nx = b->next;
            free(b->tuples);
            free(b);
            b   = nx;
            cnt = CHAINEDBUFF_NUMTUPLESPERBUF;
        }
        free(cb);
    }
    free(res->resultlist);
#endif
    free(res);
    free(buf);
    free(relR.tuples);
    free(relS.tuples);
    return 0;
}

/** reference filter, built with the reference's own add(); returns the bitmap for byte equality */
int
refshim_bloom_build(const tuple_t * R, uint64_t nR, int variant, uint64_t m, uint64_t k,
                    uint64_t B, uint32_t seed, unsigned char * bitmap_out)
{
    bloom_filter_args_t args = {variant ? BLOCKED : BASIC, m, k, B};
    bloom_filter_strategy_t * s = bloom_filter_create(&args, seed);
    for (uint64_t i = 0; i < nR; i++) s->add(s->filter, R[i].key);
    memcpy(bitmap_out, s->filter->bitmap, m / 8);
    bloom_filter_destroy(s);
    return 0;
}

/** number of S tuples whose key passes contains() on the given bitmap (pass flags optional) */
int64_t
refshim_bloom_count(const unsigned char * bitmap, const tuple_t * S, uint64_t nS, int variant,
                    uint64_t m, uint64_t k, uint64_t B, uint32_t seed, unsigned char * pass_out)
{
    bloom_filter_args_t args = {variant ? BLOCKED : BASIC, m, k, B};
    bloom_filter_strategy_t * s = bloom_filter_create(&args, seed);
    memcpy(s->filter->bitmap, bitmap, m / 8);
    int64_t n = 0;
    for (uint64_t i = 0; i < nS; i++) {
        bool p = s->contains(s->filter, S[i].key);
        n += p;
        if (pass_out) pass_out[i] = (unsigned char) p;
    }
    bloom_filter_destroy(s);
    return n;
}

/** which: 0 crc,1 FNV,2 crapwow,3 Coffin,4 MurmurOAAT,5 JenkinsOAAT,6 Spooky,7 KR_v2,8 DJB2,9 x17 */
uint32_t
refshim_hash(int which, uint32_t seed, int32_t key)
{
    switch (which) {
        case 0: return hash_crc(seed, key);
        case 1: return hash_FNV(seed, key);
        case 2: return hash_crapwow(seed, key);
        case 3: return hash_Coffin(seed, key);
        case 4: return hash_MurmurOAAT_32(seed, key);
        case 5: return hash_JenkinsOAAT_32(seed, key);
        case 6: return hash_Spooky(seed, key);
        case 7: return hash_KR_v2(seed, key);
        case 8: return hash_DJB2(seed, key);
        case 9: return hash_x17(seed, key);
    }
    return 0;
}

/**
 * The reference's own generators (main.c:410-466): kind 0 = R (parallel_create_relation(n,n,n,1.0)),
 * kind 1 = S uniform FK with selectivity, kind 2 = S zipf. Writes n tuples into out.
 */
int
refshim_generate(int kind, tuple_t * out, uint64_t n, uint64_t r_size, double q, double zipf,
                 unsigned int seed, int nthr)
{
    relation_t rel;
    numalocalize = 0;
    nthreads     = nthr;
    seed_generator(seed);
    if (kind == 0) parallel_create_relation(&rel, n, nthr, n, n, 1.0);
    else if (kind == 1) parallel_create_relation(&rel, n, nthr, 2147483647, r_size, q);
    else {
        create_relation_zipf(&rel, n, r_size, zipf);
        /* genzipf.c:147-148 leaves payloads uninitialised: define them */
        for (uint64_t i = 0; i < n; i++) rel.tuples[i].payload = (value_t) i;
    }
    memcpy(out, rel.tuples, n * sizeof(tuple_t));
    delete_relation(&rel);
    return 0;
}
