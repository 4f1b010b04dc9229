/*
 * mchashjoins_gpu.c -- C host driver: the reference's `mchashjoins` command line (main.c:351-731) on top of
 * libhwbrj_cuda.so. Same knobs (-a -n -r -s -x -y -q -z -b -m -k -B -R -S --non-unique --full-range
 * --basic-numa accepted), same dispatch through an algos[] table of function pointers (main.c:277-282,331-339),
 * same stdout contract (the lines measurements/run.py:100-156 parses). Only the join entry points differ: they
 * are the CUDA implementations exported by the C ABI in include/hwbrj.h.
 *
 * Relations are produced by the library's on-device generator (the reference generator's key multiset,
 * generator.c:162-195,341-387) and copied to HOST arrays, so the join call below is the genuine host-buffer
 * drop-in call `algo->joinAlgoBloom(&relR,&relS,nthreads,&bloom_filter_args)` of main.c:473-478.
 * -z (Zipf), --non-unique and --full-range restate the reference's serial glibc-rand() generators on the host
 * (genzipf.c:97-158, generator.c:271-279,531-651), so the same seeds give the same arrays as the reference binary.
 * --gpus N (not a reference knob) shards the join over N GPUs of the box (hwbrj_set_gpus).
 */
#include <getopt.h>
#include <limits.h>
#include <math.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hwbrj.h"

typedef struct algo_t {
    char name[128];
    result_t * (*joinAlgo)(relation_t *, relation_t *, int);
    result_t * (*joinAlgoBloom)(relation_t *, relation_t *, int, bloom_filter_args_t *);
} algo_t;

/* main.c:331-339 minus NPO/NPO_st (out of scope: no-partitioning join) */
static algo_t algos[] = {
    {"PRO", PRO, BPRO}, {"RJ", RJ, BRJ}, {"PRH", PRH, BPRH}, {"PRHO", PRHO, BPRHO}, {{0}, 0, 0}};

typedef struct param_t {
    algo_t * algo;
    uint32_t nthreads;
    uint64_t r_size, s_size;
    uint32_t r_seed, s_seed;
    double   skew, selectivity;
    char *   loadfileR;
    char *   loadfileS;
    bool                bloom_enable;
    bloom_filter_args_t bloom_filter_args;
    bool                nonunique_keys, fullrange_keys; /* main.c:386-387 */
    int                 gpus;
    char *              dump_prefix; /* --dump-relations: write the generated inputs as <prefix>R.tbl / <prefix>S.tbl and exit */
} param_t;

static void
print_help(const char * prog)
{
    printf("Usage: %s [options]\n", prog);
    printf("  -a --algo=<name>      PRO, RJ, PRH, PRHO [PRO]\n"
           "  -n --nthreads=<N>     accepted for compatibility (the join runs on the GPU) [2]\n"
           "  -r --r-size=<R>       tuples in build relation R [128000000]\n"
           "  -s --s-size=<S>       tuples in probe relation S [128000000]\n"
           "  -x --r-seed=<x>       seed for R [12345]      -y --s-seed=<y>  seed for S [54321]\n"
           "  -q --s-sel=<q>        fraction of S tuples with a join partner [1.0]\n"
           "  -z --skew=<z>         Zipf skew of S [0.0]\n"
           "  -R --r-file=<f>  -S --s-file=<f>   load relations from text files (key payload per line)\n"
           "  -b --bloom-filter=<b> no, basic, blocked   -m --bloom-size=<bits>  -k --bloom-hashes=<k>\n"
           "  -B --bloom-block-size=<bits>\n"
           "     --non-unique       R keys drawn at random from [0, min(R, INT_MAX*q)) (duplicates)\n"
           "     --full-range       R keys drawn at random from [0, INT_MAX*q)\n"
           "     --gpus=<N>         shard the join over N GPUs of this box [1]\n"
           "     --dump-relations=<prefix>  write the inputs as <prefix>R.tbl / <prefix>S.tbl (the -R/-S format) and exit\n");
}

/* generator.c:686-741 load_relation()/read_relation(): one header line, then "key payload", "key,payload" or "key" */
static int
load_relation(relation_t * rel, const char * fn, uint64_t n)
{
    FILE * fp = fopen(fn, "r");
    if (!fp) {
        perror(fn);
        return -1;
    }
    rel->num_tuples = n;
    rel->tuples     = (tuple_t *) malloc((n ? n : 1) * sizeof(tuple_t));
    char line[256];
    if (!fgets(line, sizeof line, fp)) n = 0; /* header */
    for (uint64_t i = 0; i < n; i++) {
        int key = 0, payload = 0;
        if (!fgets(line, sizeof line, fp)) {
            rel->num_tuples = i;
            break;
        }
        if (sscanf(line, "%d %d", &key, &payload) < 2 && sscanf(line, "%d,%d", &key, &payload) < 2)
            sscanf(line, "%d", &key);
        rel->tuples[i].key     = key;
        rel->tuples[i].payload = payload;
    }
    fclose(fp);
    return 0;
}

/* create_relation_zipf (generator.c:659-676) -> gen_zipf (genzipf.c:97-158): alphabet = rand()-permuted 1..r,
 * inverse-CDF lookup; payloads (uninitialised in the reference, genzipf.c:147-148) are set to the position */
static int
create_relation_zipf(relation_t * rel, uint64_t n, uint64_t maxid, double theta)
{
    uint32_t   asz      = (uint32_t) maxid;
    uint32_t * alphabet = (uint32_t *) malloc(sizeof(uint32_t) * asz);
    double *   lut      = (double *) malloc(sizeof(double) * asz);
    rel->num_tuples     = n;
    rel->tuples         = (tuple_t *) malloc((n ? n : 1) * sizeof(tuple_t));
    if (!alphabet || !lut || !rel->tuples) return -1;
    for (uint32_t i = 0; i < asz; i++) alphabet[i] = i + 1;
    for (uint32_t i = asz - 1; i > 0; i--) {
        uint32_t k   = (uint32_t) ((unsigned long) i * (unsigned long) rand() / RAND_MAX);
        uint32_t tmp = alphabet[i];
        alphabet[i]  = alphabet[k];
        alphabet[k]  = tmp;
    }
    double scaling = 0.0, sum = 0.0;
    for (uint32_t i = 1; i <= asz; i++) scaling += 1.0 / pow(i, theta);
    for (uint32_t i = 1; i <= asz; i++) {
        sum += 1.0 / pow(i, theta);
        lut[i - 1] = sum / scaling;
    }
    for (uint64_t i = 0; i < n; i++) {
        double   r    = ((double) rand()) / RAND_MAX;
        uint32_t left = 0, right = asz - 1, pos;
        if (lut[0] >= r) pos = 0;
        else {
            while (right - left > 1) {
                uint32_t mid = (left + right) / 2;
                if (lut[mid] < r) left = mid;
                else right = mid;
            }
            pos = right;
        }
        rel->tuples[i].key     = (intkey_t) alphabet[pos];
        rel->tuples[i].payload = (value_t) i;
    }
    free(lut);
    free(alphabet);
    return 0;
}

/* RAND_RANGE (generator.c:25): O + rand() / (RAND_MAX + 1) * (N - O), truncated by the assignment to an integer */
static int64_t
rand_range(int64_t lo, int64_t hi)
{
    return (int64_t) ((double) lo + ((double) rand() / ((double) RAND_MAX + 1) * (double) (hi - lo)));
}

/* random_gen (generator.c:271-279): n keys in [minid, maxid), payload = position inside this run */
static void
random_keys(tuple_t * t, uint64_t n, int64_t minid, int64_t maxid)
{
    for (uint64_t i = 0; i < n; i++) {
        t[i].key     = (intkey_t) rand_range(minid, maxid);
        t[i].payload = (value_t) i;
    }
}

/* knuth_shuffle (generator.c:99-110): keys only, payloads stay in place; the loop counter is an int there */
static void
shuffle_keys(relation_t * rel)
{
    for (int i = (int) rel->num_tuples - 1; i > 0; i--) {
        int64_t  j         = rand_range(0, i);
        intkey_t tmp       = rel->tuples[i].key;
        rel->tuples[i].key = rel->tuples[j].key;
        rel->tuples[j].key = tmp;
    }
}

/* create_relation_nonunique (generator.c:585-605): R for --non-unique / --full-range */
static int
create_relation_nonunique(relation_t * rel, uint64_t n, int64_t maxid)
{
    rel->num_tuples = n;
    rel->tuples     = (tuple_t *) malloc((n ? n : 1) * sizeof(tuple_t));
    if (!rel->tuples) return -1;
    random_keys(rel->tuples, n, 0, maxid);
    return 0;
}

/* create_relation_fk_from_pk (generator.c:531-582): S for --full-range. The first nb tuples are copies of R's tuples
 * (key and payload), the rest random keys above the threshold; then the keys are shuffled */
static int
create_relation_fk_from_pk(relation_t * fk, const relation_t * pk, uint64_t n, uint64_t threshold, double sel)
{
    uint64_t na = (uint64_t) ((double) n * (1 - sel)), nb = n - na;
    fk->num_tuples = n;
    fk->tuples     = (tuple_t *) malloc((n ? n : 1) * sizeof(tuple_t));
    if (!fk->tuples || (nb && !pk->num_tuples)) return -1;
    random_keys(fk->tuples + nb, na, (int64_t) threshold + 1, INT_MAX);
    for (uint64_t i = 0; i < nb; i++) fk->tuples[i] = pk->tuples[i % pk->num_tuples];
    shuffle_keys(fk);
    return 0;
}

/* create_relation_nonunique_from_pk (generator.c:608-645): S for --non-unique. na random keys above the threshold, then
 * keys sampled from R with replacement, payload = position; then the keys are shuffled */
static int
create_relation_nonunique_from_pk(relation_t * rel, const relation_t * pk, uint64_t n, int64_t threshold, double sel)
{
    uint64_t na = (uint64_t) ((double) n * (1 - sel));
    rel->num_tuples = n;
    rel->tuples     = (tuple_t *) malloc((n ? n : 1) * sizeof(tuple_t));
    if (!rel->tuples || (na < n && !pk->num_tuples)) return -1;
    random_keys(rel->tuples, na, threshold + 1, INT_MAX);
    for (uint64_t i = na; i < n; i++) { /* the reference's loop counters are ints */
        int j                  = (int) rand_range(0, (int64_t) pk->num_tuples);
        rel->tuples[i].key     = pk->tuples[j].key;
        rel->tuples[i].payload = (value_t) (int) i;
    }
    shuffle_keys(rel);
    return 0;
}

/* the text format load_relation()/read_relation() read (generator.c:686-741): a header line, then "key payload" */
static int
dump_relation(const relation_t * rel, const char * prefix, const char * name)
{
    char fn[1024];
    snprintf(fn, sizeof fn, "%s%s.tbl", prefix, name);
    FILE * fp = fopen(fn, "w");
    if (!fp) {
        perror(fn);
        return -1;
    }
    fprintf(fp, "#KEY, VAL\n");
    for (uint64_t i = 0; i < rel->num_tuples; i++) fprintf(fp, "%d %d\n", rel->tuples[i].key, rel->tuples[i].payload);
    fclose(fp);
    return 0;
}

static int
device_generate(relation_t * rel, int kind, uint64_t n, uint64_t r, double q, uint64_t seed)
{
    hwbrj_rel_t * d = hwbrj_rel_generate(kind, n, r, q, seed);
    rel->num_tuples = n;
    rel->tuples     = (tuple_t *) malloc((n ? n : 1) * sizeof(tuple_t));
    if (!d || !rel->tuples) return -1;
    hwbrj_rel_download(d, rel->tuples);
    hwbrj_rel_free(d);
    return 0;
}

static void
parse_args(int argc, char ** argv, param_t * p)
{
    static struct option long_options[] = {{"algo", required_argument, 0, 'a'},
                                           {"nthreads", required_argument, 0, 'n'},
                                           {"r-size", required_argument, 0, 'r'},
                                           {"s-size", required_argument, 0, 's'},
                                           {"r-seed", required_argument, 0, 'x'},
                                           {"s-seed", required_argument, 0, 'y'},
                                           {"s-sel", required_argument, 0, 'q'},
                                           {"skew", required_argument, 0, 'z'},
                                           {"r-file", required_argument, 0, 'R'},
                                           {"s-file", required_argument, 0, 'S'},
                                           {"bloom-filter", required_argument, 0, 'b'},
                                           {"bloom-size", required_argument, 0, 'm'},
                                           {"bloom-hashes", required_argument, 0, 'k'},
                                           {"bloom-block-size", required_argument, 0, 'B'},
                                           {"non-unique", no_argument, 0, 1},
                                           {"full-range", no_argument, 0, 2},
                                           {"basic-numa", no_argument, 0, 3},
                                           {"verbose", no_argument, 0, 4},
                                           {"gpus", required_argument, 0, 5},
                                           {"dump-relations", required_argument, 0, 6},
                                           {"help", no_argument, 0, 'h'},
                                           {0, 0, 0, 0}};
    int c, idx = 0;
    /* same getopt string as main.c:601 */
    while ((c = getopt_long(argc, argv, "a:n:p:q:r:s:o:x:y:z:R:S:b:m:k:B:Z:A:hv", long_options, &idx)) != -1) {
        switch (c) {
            case 'a': {
                int found = 0;
                for (int i = 0; algos[i].joinAlgo; i++)
                    if (strcmp(optarg, algos[i].name) == 0) {
                        p->algo = &algos[i];
                        found   = 1;
                    }
                if (!found) {
                    printf("[ERROR] Join algorithm named `%s' does not exist!\n", optarg); /* main.c:625-629 */
                    print_help(argv[0]);
                    exit(EXIT_SUCCESS);
                }
                break;
            }
            case 'h':
            case '?': print_help(argv[0]); exit(EXIT_SUCCESS);
            case 'n': p->nthreads = (uint32_t) atoi(optarg); break;
            case 'q': p->selectivity = atof(optarg); break;
            case 'r': p->r_size = (uint64_t) atol(optarg); break;
            case 's': p->s_size = (uint64_t) atol(optarg); break;
            case 'x': p->r_seed = (uint32_t) atoi(optarg); break;
            case 'y': p->s_seed = (uint32_t) atoi(optarg); break;
            case 'z': p->skew = atof(optarg); break;
            case 'R': p->loadfileR = strdup(optarg); break;
            case 'S': p->loadfileS = strdup(optarg); break;
            case 'b': /* main.c:692-698: anything but "no" enables the filter; "blocked" selects BLOCKED */
                p->bloom_enable = strcmp(optarg, "no") != 0;
                if (strcmp(optarg, "basic") == 0) p->bloom_filter_args.variant = BASIC;
                else if (strcmp(optarg, "blocked") == 0) p->bloom_filter_args.variant = BLOCKED;
                break;
            case 'm': p->bloom_filter_args.m = (uint64_t) atoll(optarg); break;
            case 'k': p->bloom_filter_args.k = (uint64_t) atoi(optarg); break;
            case 'B': p->bloom_filter_args.B = (uint64_t) atoi(optarg); break;
            case 1: p->nonunique_keys = true; break; /* main.c:386-387: flags set through long_options */
            case 2: p->fullrange_keys = true; break;
            case 5: p->gpus = atoi(optarg); break;
            case 6: p->dump_prefix = strdup(optarg); break;
            default: break;
        }
    }
    if (p->bloom_enable && hwbrj_check_args(&p->bloom_filter_args)) exit(1); /* assert_args, main.c:730 */
}

int
main(int argc, char ** argv)
{
    relation_t relR, relS;
    param_t    p;
    memset(&p, 0, sizeof p);
    /* defaults: main.c:370-393 */
    p.algo                      = &algos[0];
    p.nthreads                  = 2;
    p.r_size                    = 128000000;
    p.s_size                    = 128000000;
    p.r_seed                    = 12345;
    p.s_seed                    = 54321;
    p.skew                      = 0.0;
    p.selectivity               = 1.0;
    p.bloom_enable              = false;
    p.bloom_filter_args.variant = BASIC;
    p.bloom_filter_args.m       = 256 << 20;
    p.bloom_filter_args.k       = 8;
    p.bloom_filter_args.B       = 1024;
    p.gpus                      = 1;
    parse_args(argc, argv, &p);

    const bool host_generated_r = p.loadfileR || p.fullrange_keys || p.nonunique_keys;
    const bool host_generated_s = p.loadfileS || p.fullrange_keys || p.nonunique_keys || p.skew > 0;
    if (hwbrj_device_count() < 1 && !(p.dump_prefix && host_generated_r && host_generated_s)) {
        printf("[ERROR] no CUDA device: this driver has no CPU fallback\n");
        return EXIT_FAILURE;
    }
    if (p.gpus != 1 && hwbrj_set_gpus(p.gpus) != 0) {
        printf("[ERROR] --gpus %d: need a power of two that does not exceed the %d visible GPUs\n", p.gpus,
               hwbrj_device_count());
        return EXIT_FAILURE;
    }

    fprintf(stdout, "[INFO ] %s relation R with size = %.3lf MiB, #tuples = %llu : ",
            p.loadfileR ? "Loading" : "Creating", (double) sizeof(tuple_t) * p.r_size / 1024.0 / 1024.0,
            (unsigned long long) p.r_size);
    fflush(stdout);
    srand(p.r_seed); /* seed_generator(r_seed), main.c:411 */
    uint64_t threshold = 0;
    if (p.loadfileR) {
        if (load_relation(&relR, p.loadfileR, p.r_size)) return EXIT_FAILURE;
    } else if (p.fullrange_keys) { /* main.c:421-423 */
        threshold = (uint64_t) ceil(INT_MAX * p.selectivity);
        if (create_relation_nonunique(&relR, p.r_size, (int64_t) threshold)) return EXIT_FAILURE;
    } else if (p.nonunique_keys) { /* main.c:424-427 */
        double cap = ceil(INT_MAX * p.selectivity);
        threshold  = (double) p.r_size < cap ? p.r_size : (uint64_t) cap;
        if (create_relation_nonunique(&relR, p.r_size, (int64_t) threshold)) return EXIT_FAILURE;
    } else if (device_generate(&relR, 0, p.r_size, p.r_size, 1.0, p.r_seed))
        return EXIT_FAILURE;
    printf("OK \n");

    fprintf(stdout, "[INFO ] %s relation S with size = %.3lf MiB, #tuples = %lld : ",
            p.loadfileS ? "Loading" : "Creating", (double) sizeof(tuple_t) * p.s_size / 1024.0 / 1024.0,
            (long long) p.s_size);
    fflush(stdout);
    srand(p.s_seed); /* seed_generator(s_seed), main.c:443 */
    if (p.loadfileS) {
        if (load_relation(&relS, p.loadfileS, p.s_size)) return EXIT_FAILURE;
    } else if (p.fullrange_keys) { /* main.c:448-450 */
        if (create_relation_fk_from_pk(&relS, &relR, p.s_size, threshold, p.selectivity)) return EXIT_FAILURE;
    } else if (p.nonunique_keys) { /* main.c:451-453 */
        if (create_relation_nonunique_from_pk(&relS, &relR, p.s_size, (int64_t) threshold, p.selectivity))
            return EXIT_FAILURE;
    } else if (p.skew > 0) {
        if (create_relation_zipf(&relS, p.s_size, p.r_size, p.skew)) return EXIT_FAILURE;
    } else if (device_generate(&relS, 1, p.s_size, p.r_size, p.selectivity, p.s_seed))
        return EXIT_FAILURE;
    printf("OK \n");

    if (p.dump_prefix) { /* inputs only: what the reference's PERSIST_RELATIONS build writes (generator.c:43,574) */
        if (dump_relation(&relR, p.dump_prefix, "R") || dump_relation(&relS, p.dump_prefix, "S")) return EXIT_FAILURE;
        printf("[INFO ] relations written to %sR.tbl and %sS.tbl\n", p.dump_prefix, p.dump_prefix);
        return 0;
    }

    printf("[INFO ] Running join algorithm %s ...\n", p.algo->name);
    result_t * results;
    if (p.bloom_enable) results = p.algo->joinAlgoBloom(&relR, &relS, (int) p.nthreads, &p.bloom_filter_args);
    else results = p.algo->joinAlgo(&relR, &relS, (int) p.nthreads);
    printf("[INFO ] Results = %llu. DONE.\n", (unsigned long long) results->totalresults);

    hwbrj_stats_t st;
    hwbrj_last_stats(&st);
    printf("[INFO ] checksum(pair) = %llu, checksum(key) = %llu, device ms = %.3f, radix bits = %d, filter range passes = %d\n",
           (unsigned long long) st.checksum_pair, (unsigned long long) st.checksum_key, st.ms_total, st.radix_bits,
           st.range_passes);

    free(relR.tuples);
    free(relS.tuples);
    free(results);
    return 0;
}
