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
 * -z (Zipf) restates gen_zipf (genzipf.c:97-158) on the host with glibc rand(), like the reference.
 */
#include <getopt.h>
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
           "  -B --bloom-block-size=<bits>\n");
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
    parse_args(argc, argv, &p);

    if (hwbrj_device_count() < 1) {
        printf("[ERROR] no CUDA device: this driver has no CPU fallback\n");
        return EXIT_FAILURE;
    }

    fprintf(stdout, "[INFO ] %s relation R with size = %.3lf MiB, #tuples = %llu : ",
            p.loadfileR ? "Loading" : "Creating", (double) sizeof(tuple_t) * p.r_size / 1024.0 / 1024.0,
            (unsigned long long) p.r_size);
    fflush(stdout);
    if (p.loadfileR) {
        if (load_relation(&relR, p.loadfileR, p.r_size)) return EXIT_FAILURE;
    } else if (device_generate(&relR, 0, p.r_size, p.r_size, 1.0, p.r_seed))
        return EXIT_FAILURE;
    printf("OK \n");

    fprintf(stdout, "[INFO ] %s relation S with size = %.3lf MiB, #tuples = %lld : ",
            p.loadfileS ? "Loading" : "Creating", (double) sizeof(tuple_t) * p.s_size / 1024.0 / 1024.0,
            (long long) p.s_size);
    fflush(stdout);
    if (p.loadfileS) {
        if (load_relation(&relS, p.loadfileS, p.s_size)) return EXIT_FAILURE;
    } else if (p.skew > 0) {
        srand(p.s_seed); /* seed_generator(s_seed), main.c:443 */
        if (create_relation_zipf(&relS, p.s_size, p.r_size, p.skew)) return EXIT_FAILURE;
    } else if (device_generate(&relS, 1, p.s_size, p.r_size, p.selectivity, p.s_seed))
        return EXIT_FAILURE;
    printf("OK \n");

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
