/*
 * oracle.c -- TEST INFRASTRUCTURE. CPU restatement (plain C, single thread) of the reference's
 * Bloom-filter radix hash join hot path. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product path (libhwbrj_cuda.so)
 * never does and has no CPU fallback.
 *
 * PARITY STATUS: PINNED. tests/test_oracle.py checks this file against
 *   (1) the hash known-answer vectors of SURVEY.md Appendix C (generated from hash.c/spooky.c),
 *   (2) the filter bitmap vectors of Appendix C,
 *   (3) the golden `filtered` / `matches` values mined from the reference's own
 *       measurements/data/pkl pickles (tests/golden/golden_results.json),
 *   (4) the compiled, unmodified reference itself (oracle/_ref/libref*.so) on identical arrays.
 *
 * Every function cites the reference file:line (relative to /root/reference/src) it restates.
 * Nothing here is copied; the arithmetic is re-derived from SURVEY.md Appendix A.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t key;
    int32_t payload;
} orc_tuple_t; /* types.h:37-40 (8-byte tuples, KEY_8B off) */

typedef struct {
    int64_t  matches;       /* result_t.totalresults, types.h:59 */
    int64_t  filtered;      /* "S-tuples after filter", parallel_radix_join_bloom.c:1188-1193,1253 */
    uint64_t checksum_pair; /* sum mix64(R.payload,S.payload) over output pairs (:307-312) */
    uint64_t checksum_rpay; /* sum (uint32)R.payload over output pairs */
    uint64_t checksum_spay; /* sum (uint32)S.payload over output pairs */
    uint64_t checksum_key;  /* sum (uint32)S.key over output pairs */
} orc_result_t;

/* ------------------------------------------------------------------------------------------ */
/* hashes: hash.c:6-140, spooky.c:14-43, spooky.h:110-146,166                                  */
/* ------------------------------------------------------------------------------------------ */

/* hash.c:6-10 uses _mm_crc32_u32(seed,key): CRC-32C (Castagnoli, reflected poly 0x82F63B78),
 * register initialised with `seed`, the four key bytes fed LSB first, no final xor. */
static uint32_t
h_crc(uint32_t seed, int32_t key)
{
    uint32_t crc = seed ^ (uint32_t) key;
    for (int bit = 0; bit < 32; bit++) crc = (crc >> 1) ^ (0x82F63B78u & (0u - (crc & 1u)));
    return crc;
}

/* every byte-wise hash in hash.c stores the byte in a (signed) char before mixing it in
 * (hash.c:19,61,75,89,113,125,137), so bytes >= 0x80 are sign-extended. */
static inline int32_t
sbyte(int32_t key, int i)
{
    return (int32_t) (int8_t) ((key >> (8 * i)) & 0xFF);
}

static uint32_t
h_fnv(uint32_t seed, int32_t key) /* hash.c:12-25 */
{
    uint32_t h = seed ^ 2166136261u;
    for (int i = 0; i < 4; i++) {
        h ^= (uint32_t) sbyte(key, i);
        h *= 16777619u;
    }
    return h;
}

static uint32_t
h_crapwow(uint32_t seed, int32_t key) /* hash.c:27-47 */
{
    const uint32_t n  = 0x5052acdbu;
    uint32_t       lo = 4u;            /* h = sizeof(intkey_t) */
    uint32_t       hi = 4u + seed + n; /* k = h + seed + n */
    uint64_t       p  = (uint64_t) (uint32_t) key * n;
    lo ^= (uint32_t) p;
    hi ^= (uint32_t) (p >> 32);
    p = (uint64_t) (lo ^ (hi + n)) * n;
    lo ^= (uint32_t) p;
    hi ^= (uint32_t) (p >> 32);
    return hi ^ lo;
}

static inline uint32_t
rol32(uint32_t x, int b)
{
    return (x << b) | (x >> (32 - b));
}

static uint32_t
h_coffin(uint32_t seed, int32_t key) /* hash.c:55-66, seed unused */
{
    (void) seed;
    uint32_t r = 0x55555555u;
    for (int i = 0; i < 4; i++) {
        r ^= (uint32_t) sbyte(key, i);
        r = rol32(r, 5);
    }
    return r;
}

static uint32_t
h_murmur_oaat(uint32_t seed, int32_t key) /* hash.c:68-81 */
{
    uint32_t h = seed;
    for (int i = 0; i < 4; i++) {
        h ^= (uint32_t) sbyte(key, i);
        h *= 0x5bd1e995u;
        h ^= h >> 15;
    }
    return h;
}

static uint32_t
h_jenkins_oaat(uint32_t seed, int32_t key) /* hash.c:83-99 */
{
    uint32_t h = seed;
    for (int i = 0; i < 4; i++) {
        h += (uint32_t) sbyte(key, i);
        h += h << 10;
        h ^= h >> 6;
    }
    h += h << 3;
    h ^= h >> 11;
    h += h << 15;
    return h;
}

static inline uint64_t
rol64(uint64_t x, int b)
{
    return (x << b) | (x >> (64 - b));
}

/* spooky.c:14-20 Short(): c = sc_const + (sign-extended key), d = 4<<56, then ShortEnd
 * (spooky.h:110-146) on (h0=seed,h1=seed,c,d); hash_spooky32 returns the low 32 bits of h0
 * (spooky.c:37-43). hash_Spooky(seed,key) = hash_spooky32(key,seed), hash.c:101-105. */
static uint32_t
h_spooky(uint32_t seed, int32_t key)
{
    uint64_t h0 = seed, h1 = seed;
    uint64_t h2 = 0xdeadbeefdeadbeefULL + (uint64_t) (int64_t) key;
    uint64_t h3 = (uint64_t) 4 << 56;
    static const int rot[11] = {15, 52, 26, 51, 28, 9, 47, 54, 32, 25, 63};
    uint64_t *       v[4]    = {&h0, &h1, &h2, &h3};
    /* step s (0..10): a = v[(s+3)%4], b = v[(s+2)%4]:  a ^= b; b = rot(b); a += b  */
    for (int s = 0; s < 11; s++) {
        uint64_t * a = v[(s + 3) & 3];
        uint64_t * b = v[(s + 2) & 3];
        *a ^= *b;
        *b = rol64(*b, rot[s]);
        *a += *b;
    }
    return (uint32_t) h0;
}

static uint32_t
h_kr_v2(uint32_t seed, int32_t key) /* hash.c:107-117 */
{
    uint32_t h = seed;
    for (int i = 0; i < 4; i++) h = (uint32_t) sbyte(key, i) + 31u * h;
    return h;
}

static uint32_t
h_djb2(uint32_t seed, int32_t key) /* hash.c:119-129, seed unused */
{
    (void) seed;
    uint32_t h = 5381u;
    for (int i = 0; i < 4; i++) h = ((h << 5) + h) + (uint32_t) sbyte(key, i);
    return h;
}

static uint32_t
h_x17(uint32_t seed, int32_t key) /* hash.c:131-140 */
{
    uint32_t h = seed;
    for (int i = 0; i < 4; i++) h = 17u * h + (uint32_t) (sbyte(key, i) - ' ');
    return h ^ (h >> 16);
}

/** which: 0 crc,1 FNV,2 crapwow,3 Coffin,4 MurmurOAAT,5 JenkinsOAAT,6 Spooky,7 KR_v2,8 DJB2,9 x17
 *  (the order of hash.h:12-40) */
uint32_t
orc_hash(int which, uint32_t seed, int32_t key)
{
    switch (which) {
        case 0: return h_crc(seed, key);
        case 1: return h_fnv(seed, key);
        case 2: return h_crapwow(seed, key);
        case 3: return h_coffin(seed, key);
        case 4: return h_murmur_oaat(seed, key);
        case 5: return h_jenkins_oaat(seed, key);
        case 6: return h_spooky(seed, key);
        case 7: return h_kr_v2(seed, key);
        case 8: return h_djb2(seed, key);
        case 9: return h_x17(seed, key);
    }
    return 0;
}

void
orc_hash_many(int which, uint32_t seed, const int32_t * keys, uint64_t n, uint32_t * out)
{
    for (uint64_t i = 0; i < n; i++) out[i] = orc_hash(which, seed, keys[i]);
}

/* ------------------------------------------------------------------------------------------ */
/* Bloom filter: bloom_filter.c:74-141 (enhanced double hashing), variant 0 BASIC, 1 BLOCKED   */
/* ------------------------------------------------------------------------------------------ */

/* bit sequence of one key inside a (sub)filter of `size` bits: bloom_filter.c:77-88 / 96-109.
 * All arithmetic is uint32 and reduced with & (size-1) (mod_m, :60-63). */
static inline void
bloom_locate(int variant, uint64_t m, uint64_t B, uint32_t seed, int32_t key, uint64_t * base_bit,
             uint32_t * size)
{
    if (variant == 0) { /* add_basic/contains_basic :114-123 */
        *base_bit = 0;
        *size     = (uint32_t) m; /* uint32_t size parameter: m == 2^32 wraps to 0, as in the reference */
    } else {              /* add_blocked/contains_blocked :126-141 */
        uint64_t nblocks = m / B;
        uint32_t blk     = h_crc(seed, key) & (uint32_t) (nblocks - 1);
        *base_bit        = (uint64_t) blk * (B / 8) * 8;
        *size            = (uint32_t) B;
    }
}

void
orc_bloom_add(unsigned char * bitmap, int variant, uint64_t m, uint64_t k, uint64_t B,
              uint32_t seed, int32_t key)
{
    uint64_t base;
    uint32_t size;
    bloom_locate(variant, m, B, seed, key, &base, &size);
    uint32_t mask = size - 1u;
    uint32_t h    = h_crapwow(seed, key) & mask;
    uint32_t y    = ((uint32_t) key + seed) & mask;
    for (uint32_t i = 0; i < k; i++) {
        uint64_t bit = base + h;
        bitmap[bit >> 3] |= (unsigned char) (1u << (bit & 7));
        h = (h + y) & mask;
        y = (y + i + 1u) & mask;
    }
}

int
orc_bloom_contains(const unsigned char * bitmap, int variant, uint64_t m, uint64_t k, uint64_t B,
                   uint32_t seed, int32_t key)
{
    uint64_t base;
    uint32_t size;
    bloom_locate(variant, m, B, seed, key, &base, &size);
    uint32_t mask = size - 1u;
    uint32_t h    = h_crapwow(seed, key) & mask;
    uint32_t y    = ((uint32_t) key + seed) & mask;
    for (uint32_t i = 0; i < k; i++) {
        uint64_t bit = base + h;
        if (!(bitmap[bit >> 3] & (1u << (bit & 7)))) return 0;
        h = (h + y) & mask;
        y = (y + i + 1u) & mask;
    }
    return 1;
}

/** bloom_filter_create (:144-179: zeroed m/8 bytes) followed by add() over all R keys
 *  (parallel_radix_join_bloom.c:794-805, build branch). */
void
orc_bloom_build(const orc_tuple_t * R, uint64_t nR, int variant, uint64_t m, uint64_t k,
                uint64_t B, uint32_t seed, unsigned char * bitmap)
{
    memset(bitmap, 0, m / 8);
    for (uint64_t i = 0; i < nR; i++) orc_bloom_add(bitmap, variant, m, k, B, seed, R[i].key);
}

/** probe branch of the histogram loop (:794-805): returns the number of passing tuples and
 *  optionally compacts them (order preserved) into `survivors`. */
int64_t
orc_bloom_filter(const unsigned char * bitmap, const orc_tuple_t * S, uint64_t nS, int variant,
                 uint64_t m, uint64_t k, uint64_t B, uint32_t seed, orc_tuple_t * survivors)
{
    int64_t n = 0;
    for (uint64_t i = 0; i < nS; i++) {
        if (orc_bloom_contains(bitmap, variant, m, k, B, seed, S[i].key)) {
            if (survivors) survivors[n] = S[i];
            n++;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* radix clustering + bucket chaining                                                           */
/* ------------------------------------------------------------------------------------------ */

/* HASH_BIT_MODULO(K,MASK,NBITS) = (K & MASK) >> NBITS, parallel_radix_join_bloom.c:74 */
#define HBM(K, MASK, NBITS) ((((uint32_t) (K)) & (MASK)) >> (NBITS))

/** radix_cluster_nopadding without the filter hooks (:621-693): count, exclusive prefix,
 *  stable scatter on bits [R, R+D). hist must hold 1<<D entries and receives the counts. */
static void
radix_cluster(orc_tuple_t * out, const orc_tuple_t * in, uint64_t n, int R, int D, uint64_t * hist)
{
    uint32_t   fan  = 1u << D;
    uint32_t   M    = (fan - 1u) << R;
    uint64_t * dst  = (uint64_t *) malloc(sizeof(uint64_t) * fan);
    memset(hist, 0, sizeof(uint64_t) * fan);
    for (uint64_t i = 0; i < n; i++) hist[HBM(in[i].key, M, R)]++;
    uint64_t off = 0;
    for (uint32_t j = 0; j < fan; j++) {
        dst[j] = off;
        off += hist[j];
    }
    for (uint64_t i = 0; i < n; i++) out[dst[HBM(in[i].key, M, R)]++] = in[i];
    free(dst);
}

static inline uint64_t
mix64(uint32_t rpay, uint32_t spay)
{
    uint64_t z = ((uint64_t) rpay << 32) | (uint64_t) spay;
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

/** bucket_chaining_join (:260-329): N = next pow2 >= numR, MASK = (N-1) << radix_bits,
 *  next[i] = bucket[idx]; bucket[idx] = i+1; probe walks the chain and counts every equal key.
 *  The output pair is (R.payload, S.payload) (:307-312). */
/* optional materialisation target of bucket_chaining_join (the JOIN_RESULT_MATERIALIZE branch, :307-312) */
static orc_tuple_t * g_pairs     = NULL;
static uint64_t      g_pairs_cap = 0, g_pairs_n = 0;

static void
bucket_chaining_join(const orc_tuple_t * R, uint32_t numR, const orc_tuple_t * S, uint32_t numS,
                     int radix_bits, orc_result_t * acc)
{
    uint32_t N = numR;
    N--;
    N |= N >> 1;
    N |= N >> 2;
    N |= N >> 4;
    N |= N >> 8;
    N |= N >> 16;
    N++;
    if (N == 0) N = 1;
    const uint32_t MASK   = (N - 1u) << radix_bits;
    int32_t *      next   = (int32_t *) malloc(sizeof(int32_t) * (numR ? numR : 1));
    int32_t *      bucket = (int32_t *) calloc(N, sizeof(int32_t));
    for (uint32_t i = 0; i < numR;) {
        uint32_t idx = HBM(R[i].key, MASK, radix_bits);
        next[i]      = bucket[idx];
        bucket[idx]  = (int32_t) ++i;
    }
    for (uint32_t i = 0; i < numS; i++) {
        uint32_t idx = HBM(S[i].key, MASK, radix_bits);
        for (int32_t hit = bucket[idx]; hit > 0; hit = next[hit - 1]) {
            if (S[i].key == R[hit - 1].key) {
                if (g_pairs) { /* joinres->key = R-rid; joinres->payload = S-rid (:310-311) */
                    if (g_pairs_n < g_pairs_cap) {
                        g_pairs[g_pairs_n].key     = R[hit - 1].payload;
                        g_pairs[g_pairs_n].payload = S[i].payload;
                    }
                    g_pairs_n++;
                }
                acc->matches++;
                acc->checksum_pair += mix64((uint32_t) R[hit - 1].payload, (uint32_t) S[i].payload);
                acc->checksum_rpay += (uint32_t) R[hit - 1].payload;
                acc->checksum_spay += (uint32_t) S[i].payload;
                acc->checksum_key += (uint32_t) S[i].key;
            }
        }
    }
    free(bucket);
    free(next);
}

/** RJ/BRJ restated (parallel_radix_join_bloom.c:1808-1977; the parallel PRO/BPRO variants give
 *  the same result scalars): two LSD clustering passes over
 *  radix_bits/2 and radix_bits - radix_bits/2 bits (:1864-1879), then one bucket-chaining join
 *  per cluster (:1896-1939). R and S are left untouched (the reference clobbers them). */
static void
radix_join(const orc_tuple_t * R, uint64_t nR, const orc_tuple_t * S, uint64_t nS, int radix_bits,
           orc_result_t * acc)
{
    int           b1   = radix_bits / 2, b2 = radix_bits - b1;
    uint32_t      fan  = 1u << radix_bits;
    orc_tuple_t * tR   = (orc_tuple_t *) malloc(sizeof(orc_tuple_t) * (nR ? nR : 1));
    orc_tuple_t * tS   = (orc_tuple_t *) malloc(sizeof(orc_tuple_t) * (nS ? nS : 1));
    orc_tuple_t * pR   = (orc_tuple_t *) malloc(sizeof(orc_tuple_t) * (nR ? nR : 1));
    orc_tuple_t * pS   = (orc_tuple_t *) malloc(sizeof(orc_tuple_t) * (nS ? nS : 1));
    uint64_t *    hist = (uint64_t *) malloc(sizeof(uint64_t) * fan);
    radix_cluster(tR, R, nR, 0, b1, hist);
    radix_cluster(pR, tR, nR, b1, b2, hist);
    radix_cluster(tS, S, nS, 0, b1, hist);
    radix_cluster(pS, tS, nS, b1, b2, hist);
    /* recount per full cluster (:1896-1909); clusters are laid out in increasing order of the
       low radix_bits because the second pass is stable */
    uint64_t * cR = (uint64_t *) calloc(fan, sizeof(uint64_t));
    uint64_t * cS = (uint64_t *) calloc(fan, sizeof(uint64_t));
    for (uint64_t i = 0; i < nR; i++) cR[(uint32_t) pR[i].key & (fan - 1u)]++;
    for (uint64_t i = 0; i < nS; i++) cS[(uint32_t) pS[i].key & (fan - 1u)]++;
    uint64_t r = 0, s = 0;
    for (uint32_t c = 0; c < fan; c++) {
        if (cR[c] > 0 && cS[c] > 0)
            bucket_chaining_join(pR + r, (uint32_t) cR[c], pS + s, (uint32_t) cS[c], radix_bits, acc);
        r += cR[c];
        s += cS[c];
    }
    free(cR);
    free(cS);
    free(hist);
    free(tR);
    free(tS);
    free(pR);
    free(pS);
}

/** Full hot path: (optional) filter build over R, S pre-filter, radix join of R with survivors.
 *  bloom_enable=0 restates PRO/RJ (parallel_radix_join.c:1697,1718), else BPRO/BRJ with
 *  seed 42 (parallel_radix_join_bloom.c:1583,1823). */
int
orc_join(const orc_tuple_t * R, uint64_t nR, const orc_tuple_t * S, uint64_t nS, int bloom_enable,
         int variant, uint64_t m, uint64_t k, uint64_t B, int radix_bits, orc_result_t * out)
{
    memset(out, 0, sizeof(*out));
    if (!bloom_enable) {
        out->filtered = -1;
        radix_join(R, nR, S, nS, radix_bits, out);
        return 0;
    }
    unsigned char * bitmap = (unsigned char *) malloc(m / 8);
    orc_tuple_t *   surv   = (orc_tuple_t *) malloc(sizeof(orc_tuple_t) * (nS ? nS : 1));
    if (!bitmap || !surv) return -1;
    orc_bloom_build(R, nR, variant, m, k, B, 42u, bitmap);
    int64_t f = orc_bloom_filter(bitmap, S, nS, variant, m, k, B, 42u, surv);
    radix_join(R, nR, surv, (uint64_t) f, radix_bits, out);
    out->filtered = f;
    free(bitmap);
    free(surv);
    return 0;
}

/** orc_join that also materialises the output pairs {R.payload, S.payload}; returns the number of pairs */
int64_t
orc_join_pairs(const orc_tuple_t * R, uint64_t nR, const orc_tuple_t * S, uint64_t nS, int bloom_enable, int variant,
               uint64_t m, uint64_t k, uint64_t B, int radix_bits, orc_tuple_t * pairs, uint64_t cap)
{
    orc_result_t res;
    g_pairs     = pairs;
    g_pairs_cap = cap;
    g_pairs_n   = 0;
    int rc      = orc_join(R, nR, S, nS, bloom_enable, variant, m, k, B, radix_bits, &res);
    g_pairs     = NULL;
    return rc ? -1 : (int64_t) g_pairs_n;
}

/* ------------------------------------------------------------------------------------------ */
/* generators                                                                                   */
/* ------------------------------------------------------------------------------------------ */

static inline uint64_t
splitmix(uint64_t * s)
{
    uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
    z          = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z          = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

/** parallel_create_relation (generator.c:305-415) + random_unique_gen_thread (:162-221).
 *  Key fill follows the reference's per-thread arithmetic exactly (page-rounded chunking :341-354,
 *  firstkey/firstkey_above/ridstart :370-387, wrap rules :182-195), so the key MULTISET equals the
 *  reference's for the same (n, nthr, maxid, threshold, selectivity). The reference then shuffles key
 *  positions with a time-seeded nrand48 (:173-176,:204-217) -- not reproducible by design -- so the
 *  shuffle here is a seeded Fisher-Yates over keys only (payload stays = position, as in :179,:188). */
int
orc_gen_relation(orc_tuple_t * rel, uint64_t num_tuples, uint32_t nthr, uint64_t maxid,
                 uint64_t threshold, double selectivity, uint64_t shuffle_seed)
{
    const uint64_t pagesize        = 4096;
    uint64_t       npages          = (num_tuples * sizeof(orc_tuple_t)) / pagesize + 1;
    uint64_t       npages_perthr   = npages / nthr;
    uint64_t       ntuples_perthr  = npages_perthr * (pagesize / sizeof(orc_tuple_t));
    uint64_t       ntuples_above   = (uint64_t) (num_tuples * (1 - selectivity));
    if (npages_perthr == 0) ntuples_perthr = num_tuples / nthr;
    uint64_t ntuples_above_perthr  = (uint64_t) (ntuples_perthr * (1 - selectivity));
    uint64_t ntuples_lastthr       = num_tuples - ntuples_perthr * (nthr - 1);
    uint64_t ntuples_above_lastthr = ntuples_above - (nthr - 1) * ntuples_above_perthr;
    uint64_t offset = 0, offset_above = 0;
    for (uint32_t t = 0; t < nthr; t++) {
        int64_t  firstkey       = (int64_t) ((offset + 1) % threshold);
        uint64_t span           = maxid - threshold;
        int64_t  firstkey_above = (int64_t) (threshold + (offset_above + 1) % (span > 1 ? span : 1));
        uint64_t above          = (t == nthr - 1) ? ntuples_above_lastthr : ntuples_above_perthr;
        uint64_t cnt            = (t == nthr - 1) ? ntuples_lastthr : ntuples_perthr;
        uint64_t ridstart       = offset + offset_above;
        orc_tuple_t * p         = rel + ridstart;
        uint64_t below          = cnt - above;
        uint64_t i;
        for (i = 0; i < below; i++) {
            p[i].key     = (int32_t) firstkey;
            p[i].payload = (int32_t) (ridstart + i);
            if (firstkey == (int64_t) threshold) firstkey = 0;
            firstkey++;
        }
        for (; i < cnt; i++) {
            p[i].key     = (int32_t) firstkey_above;
            p[i].payload = (int32_t) (ridstart + i);
            if (firstkey_above == 2147483647) firstkey_above = (int64_t) threshold;
            firstkey_above++;
        }
        offset += ntuples_perthr - ntuples_above_perthr;
        offset_above += ntuples_above_perthr;
    }
    if (shuffle_seed) {
        uint64_t s = shuffle_seed;
        for (uint64_t i = num_tuples - 1; i > 0; i--) {
            uint64_t j   = (uint64_t) (((__uint128_t) splitmix(&s) * (i + 1)) >> 64);
            int32_t  tmp = rel[i].key;
            rel[i].key   = rel[j].key;
            rel[j].key   = tmp;
        }
    }
    return 0;
}

/** create_relation_zipf (generator.c:659-676) -> gen_zipf (genzipf.c:97-158) with gen_alphabet
 *  (:27-52) and gen_zipf_lut (:59-92). Uses glibc rand() after srand(seed) exactly like the
 *  reference (main.c:443 seed_generator(s_seed)), so with the same libc the key ARRAY is
 *  identical. The reference leaves payloads uninitialised (genzipf.c:147-148); they are defined
 *  here as the position index. */
int
orc_gen_zipf(orc_tuple_t * rel, uint64_t stream_size, uint32_t alphabet_size, double theta,
             unsigned int seed)
{
    srand(seed);
    uint32_t * alphabet = (uint32_t *) malloc(sizeof(uint32_t) * alphabet_size);
    double *   lut      = (double *) malloc(sizeof(double) * alphabet_size);
    if (!alphabet || !lut) return -1;
    for (uint32_t i = 0; i < alphabet_size; i++) alphabet[i] = i + 1;
    for (uint32_t i = alphabet_size - 1; i > 0; i--) {
        uint32_t k   = (uint32_t) ((unsigned long) i * (unsigned long) rand() / RAND_MAX);
        uint32_t tmp = alphabet[i];
        alphabet[i]  = alphabet[k];
        alphabet[k]  = tmp;
    }
    double scaling = 0.0;
    for (uint32_t i = 1; i <= alphabet_size; i++) scaling += 1.0 / pow(i, theta);
    double sum = 0.0;
    for (uint32_t i = 1; i <= alphabet_size; i++) {
        sum += 1.0 / pow(i, theta);
        lut[i - 1] = sum / scaling;
    }
    for (uint64_t i = 0; i < stream_size; i++) {
        double   r    = ((double) rand()) / RAND_MAX;
        uint32_t left = 0, right = alphabet_size - 1, pos;
        if (lut[0] >= r) pos = 0;
        else {
            while (right - left > 1) {
                uint32_t mid = (left + right) / 2;
                if (lut[mid] < r) left = mid;
                else right = mid;
            }
            pos = right;
        }
        rel[i].key     = (int32_t) alphabet[pos];
        rel[i].payload = (int32_t) i;
    }
    free(lut);
    free(alphabet);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* inputs of the reference's Bloom FPR measurement (`unittests 2 ...`)                          */
/* ------------------------------------------------------------------------------------------ */

/* glibc rand()/srand() restated (TYPE_3 additive feedback generator, x^31 + x^3 + 1): the reference draws ~2.1e9
 * numbers for one FPR run and libc's rand() takes a lock per call (~27 ns -> a minute); this lock-free copy of the
 * same recurrence is ~10x faster. tests/test_oracle.py checks it against libc's rand() itself. */
typedef struct {
    int32_t r[34];
    int     f, b; /* front / rear indices into r[3..33] style ring (implemented as a 31-entry ring) */
    int32_t ring[31];
} grand_t;

static void
grand_seed(grand_t * g, unsigned int seed)
{
    int32_t r[344 + 31];
    if (seed == 0) seed = 1;
    r[0] = (int32_t) seed;
    for (int i = 1; i < 31; i++) {
        int64_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        int64_t w  = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        r[i] = (int32_t) w;
    }
    for (int i = 31; i < 34; i++) r[i] = r[i - 31];
    for (int i = 34; i < 344; i++) r[i] = (int32_t) ((uint32_t) r[i - 31] + (uint32_t) r[i - 3]);
    /* keep the last 31 values as the ring; next output index is 344 */
    for (int i = 0; i < 31; i++) g->ring[i] = r[344 - 31 + i];
    g->f = 0; /* position of r[i-31] */
}

static inline int
grand_next(grand_t * g)
{
    /* r[i] = r[i-31] + r[i-3]; ring[f] holds r[i-31], ring[(f+28)%31] holds r[i-3] */
    int      b = g->f + 28;
    if (b >= 31) b -= 31;
    uint32_t v = (uint32_t) g->ring[g->f] + (uint32_t) g->ring[b];
    g->ring[g->f] = (int32_t) v;
    if (++g->f == 31) g->f = 0;
    return (int) (v >> 1);
}

/** first n outputs of the restated generator (for the test against libc) */
void
orc_glibc_rand(unsigned int seed, int * out, uint32_t n)
{
    grand_t g;
    grand_seed(&g, seed);
    for (uint32_t i = 0; i < n; i++) out[i] = grand_next(&g);
}

/* random_unique_gen_range (unit_tests.c:155-173): selection sampling (Knuth) of n sorted unique values from
 * [min, min + (max-min)) with rand(); key = payload = value. */
static void
unique_range(grand_t * g, orc_tuple_t * arr, uint64_t n, int32_t min, int32_t max)
{
    uint32_t inserted  = 0;
    int32_t  m_options = max - min;
    for (uint32_t i = 0; i < (uint32_t) m_options && inserted < n; ++i) {
        int rn = (int) (n - inserted);
        int rm = m_options - (int) i;
        if (grand_next(g) % rm < rn) {
            arr[inserted].key     = min + (int32_t) i;
            arr[inserted].payload = min + (int32_t) i;
            inserted++;
        }
    }
}

/** test_bloom_fpr_wrapper (unit_tests.c:243-297): srand(seed+1); R = n_insertions values below
 *  threshold = INT32_MAX * n_ins/(n_ins+n_samples), S = n_samples values above it; every filter of the table is
 *  then created with the seed `srand(seed); rand()` (test_bloom_fpr, unit_tests.c:195-201). Returns that seed. */
uint32_t
orc_fpr_samples(int seed, uint32_t n_samples, uint32_t n_insertions, orc_tuple_t * R, orc_tuple_t * S)
{
    grand_t g;
    grand_seed(&g, (unsigned) (seed + 1));
    int32_t threshold = (int32_t) (2147483647 * (n_insertions / (double) (n_insertions + n_samples)));
    unique_range(&g, R, n_insertions, 0, threshold);
    unique_range(&g, S, n_samples, threshold + 1, 2147483647);
    grand_seed(&g, (unsigned) seed);
    return (uint32_t) grand_next(&g);
}
