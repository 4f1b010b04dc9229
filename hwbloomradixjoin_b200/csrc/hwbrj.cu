// hwbrj.cu -- host side of libhwbrj_cuda.so: workspace, the join pipeline (one GPU, or G GPUs of one NVLink domain)
// and the C ABI of include/hwbrj.h. Mirrors the reference's join_init_run()/prj_thread() orchestration
// (parallel_radix_join_bloom.c:1060-1506,1561-1778) as one CUDA stream of kernels per GPU with no host round trip
// between phases; the pthread barriers of the reference become kernel boundaries (one GPU) or peer-memory barriers
// (k_barrier) between the GPUs, which play the role of the reference's worker threads.
#include <cuda_runtime.h>
#include <unistd.h>
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/hwbrj.h"
#include "kernels.cuh"

namespace hwbrj {

[[noreturn]] void die(const char* fmt, ...) {
    // the reference's error style: print and exit (parallel_radix_join_bloom.c:64-71, bloom_filter.c:16-23)
    va_list ap;
    va_start(ap, fmt);
    fprintf(stdout, "[ERROR] hwbrj: ");
    vfprintf(stdout, fmt, ap);
    fprintf(stdout, "\n");
    fflush(stdout);
    va_end(ap);
    exit(EXIT_FAILURE);
}

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) die("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

static void invalidate_last();

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        invalidate_last();  // a re-allocation may free partitions a later hwbrj_materialize_last would read
        if (p) CK(cudaFree(p));
        size_t want = bytes + (bytes >> 4) + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        if (e != cudaSuccess) die("cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        cap = want;
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// small control block living in one allocation (zeroed with one memset per join)
struct Control {
    unsigned long long survivors;  // K2 output cursor == filtered (this rank's S chunk)
    unsigned long long n_own_r, n_own_s;  // tuples this rank owns after the level-1 routing
    JoinAccum acc;
    uint32_t item_counter[4];  // one per group of partitions (see `parts` in run_join)
    uint32_t abort;  // a buffer would overflow: the scatter kernels do nothing
    uint32_t err;    // barrier time-out
    uint32_t skew;   // k_probe_sample: bit 0 the probe relation repeats keys, bit 1 most of it passes the filter
    uint32_t pad[1];
    unsigned long long pair_cursor;  // materialised output pairs
    unsigned long long row[8];       // this rank's result words {matches, cpair, crpay, cspay, ckey, survivors, flags, 0}
    unsigned long long out[8];       // summed over the ranks
};

// what the pipeline needs to know about the GPUs that take part: for one GPU it points into the plain workspace
struct Fab {
    int world = 1, rank = 0, gbits = 0;
    PeerPtrs flags, histR, histS, rows, filter, partial;
    PeerBufs stageR, stageS;  // level-1 outputs of all ranks: the level-2 pass reads its input segments from them
    uint32_t* epoch = nullptr;
    uint64_t cap_r = 0, cap_s = 0, filter_bytes = 0;
};

struct Ctx {
    bool inited = false;
    int dev = 0;
    int sms = 0;
    int clock_khz = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev[14];
    uint32_t* d_crc = nullptr;
    DevBuf filter, histR, histS, offR, offS, cur1, cur2, tiles, segs, work, work_part, ctrl, rt1, rp, sc, st1, s2, inR, inS, scratch;
    // state of the most recent join's partitions (inputs of a materialising k_join pass)
    const uint2* last_Rp = nullptr;
    const uint2* last_Sp = nullptr;
    uint32_t last_P = 0, last_bits = 0;
    bool last_hash = false;
    DevBuf pairs;
    // host-buffer calls: upload S in chunks on a copy stream and probe each chunk as soon as it has landed
    bool overlap_h2d = false;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t side_stream = nullptr;  // several GPUs: every other group of owned bins runs here (see run_join)
    cudaEvent_t ev_ov[4];
    int dist_parts = 0;                  // groups of owned level-1 bins per relation (HWBRJ_DIST_PARTS; 0 = by size, 1 = no overlap)
    cudaEvent_t ev_copy[2];
    cudaEvent_t ev_chunk[66];
    int hash_partition = 1;  // 0 never, 1 automatic, 2 whenever the slices fit (see pick_mode)
    hwbrj_stats_t last;
    int radix_bits_override = 0;
    int passes_override = 0;
    int range_passes_override = 0;
    int probe_ctas_per_sm = 0;  // 0 = min(occupancy, 4)
    int probe_carveout = -1;    // K2 shared-memory carve-out in percent (-1 = driver default)
    bool probe_adaptive = true; // K2's load flavour and shape follow a sample of S (k_probe_sample; HWBRJ_PROBE_ADAPTIVE)
    bool probe_staged = true;   // k >= 2: probes 2..k run on compacted candidates (k_probe_staged; c1_blocked 12.6 -> 9.4 ms)
    DevBuf zipf_lut, zipf_sums;  // cumulated Zipf density of the last (alphabet size, exponent) that was generated
    uint64_t zipf_r = 0;
    double zipf_theta = -1.0;
    int occ_scatter = 1, occ_join = 1, occ_probe[8][2] = {{0}}, occ_staged = 1;
    // HWBRJ_TRACE=1: one CUDA event after every launch of a (non-captured) join; the per-kernel times are printed to stderr
    bool trace = false;
    std::vector<cudaEvent_t> tr_ev;
    std::vector<const char*> tr_name;
    size_t tr_n = 0;
};

static Ctx g_ctx[kMaxPeers];  // one context per device of this process (one process per GPU uses exactly one)
static Ctx* g_cur = nullptr;
#define g (*g_cur)
static std::recursive_mutex g_mu;
static bool g_quiet = false;
static int g_gpus = 1;  // hwbrj_set_gpus: the host-buffer entry points shard over this many GPUs of the process

static void invalidate_last() {
    if (g_cur) g.last_Rp = g.last_Sp = nullptr;
}

template <typename K>
static void set_smem(K kernel, int bytes) {
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
}

constexpr int kHistSmem = ((1 << kMaxRadixBits) + kCrcSmemWords) * 4;
constexpr int kFilterSliceSmemMax = 192 * 1024;

template <int M, int SH>
static void init_probe_shape() {
    const int smem = kProbeWarps * ProbeShape<SH>::SMEM_PER_WARP;
    set_smem(k_probe_compact<M, SH>, smem);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.occ_probe[M][SH], k_probe_compact<M, SH>, kProbeWarps * 32, smem));
    g.occ_probe[M][SH] = std::max(g.occ_probe[M][SH], 1);
    if (g.probe_carveout >= 0)
        CK(cudaFuncSetAttribute(k_probe_compact<M, SH>, cudaFuncAttributePreferredSharedMemoryCarveout, g.probe_carveout));
}
template <int M>
static void init_probe_mode() {
    init_probe_shape<M, 0>();
    init_probe_shape<M, 1>();
}

static void init_ctx() {
    std::lock_guard<std::recursive_mutex> lock(g_mu);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        die("no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev >= kMaxPeers) die("device ordinal %d not supported", dev);
    g_cur = &g_ctx[dev];  // the current CUDA device selects the context
    if (g.inited) return;
    g.dev = dev;
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, g.dev));
    g.sms = pr.multiProcessorCount;
    g.clock_khz = pr.clockRate;
    CK(cudaStreamCreateWithFlags(&g.own_stream, cudaStreamNonBlocking));
    g.stream = g.own_stream;
    for (auto& ev : g.ev) CK(cudaEventCreate(&ev));
    CK(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&g.side_stream, cudaStreamNonBlocking));
    for (auto& ev : g.ev_ov) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    if (const char* s = getenv("HWBRJ_DIST_PARTS")) g.dist_parts = std::max(0, std::min(4, atoi(s)));
    for (auto& ev : g.ev_copy) CK(cudaEventCreate(&ev));
    for (auto& ev : g.ev_chunk) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CrcTables T;
    crc_tables_fill(T);
    CK(cudaMalloc(&g.d_crc, sizeof(T)));
    CK(cudaMemcpy(g.d_crc, &T, sizeof(T), cudaMemcpyHostToDevice));
    if (const char* s = getenv("HWBRJ_OVERLAP_H2D")) g.overlap_h2d = atoi(s) != 0;
    if (const char* s = getenv("HWBRJ_RADIX_BITS")) g.radix_bits_override = atoi(s);
    if (const char* s = getenv("HWBRJ_NUM_PASSES")) g.passes_override = atoi(s);
    if (const char* s = getenv("HWBRJ_RANGE_PASSES")) g.range_passes_override = atoi(s);
    if (const char* s = getenv("HWBRJ_QUIET")) g_quiet = atoi(s) != 0;
    if (const char* s = getenv("HWBRJ_PROBE_CTAS")) g.probe_ctas_per_sm = std::max(0, atoi(s));
    if (const char* s = getenv("HWBRJ_PROBE_CARVEOUT")) g.probe_carveout = std::min(100, atoi(s));
    if (const char* s = getenv("HWBRJ_PROBE_STAGED")) g.probe_staged = atoi(s) != 0;
    if (const char* s = getenv("HWBRJ_PROBE_ADAPTIVE")) g.probe_adaptive = atoi(s) != 0;
    if (const char* s = getenv("HWBRJ_HASH_PARTITION")) g.hash_partition = std::max(0, std::min(2, atoi(s)));
    if (const char* s = getenv("HWBRJ_TRACE")) g.trace = atoi(s) != 0;
    // kernel attributes are per device: every instantiation the pipeline can launch is prepared here
    set_smem(k_build_hist<false, 0>, kHistSmem); set_smem(k_build_hist<false, 1>, kHistSmem);
    set_smem(k_build_hist<false, 2>, kHistSmem); set_smem(k_build_hist<true, 0>, kHistSmem);
    set_smem(k_scatter<1, 0>, kScatterSmem); set_smem(k_scatter<1, 1>, kScatterSmem); set_smem(k_scatter<1, 2>, kScatterSmem);
    set_smem(k_scatter<2, 0>, kScatterSmem); set_smem(k_scatter<2, 1>, kScatterSmem); set_smem(k_scatter<2, 2>, kScatterSmem);
    set_smem(k_filter_from_parts<false, false>, kFilterSliceSmemMax); set_smem(k_filter_from_parts<false, true>, kFilterSliceSmemMax);
    set_smem(k_filter_from_parts<true, false>, kFilterSliceSmemMax);  set_smem(k_filter_from_parts<true, true>, kFilterSliceSmemMax);
    set_smem(k_join<false>, kJoinSmemBytes); set_smem(k_join<true>, kJoinSmemBytes);
    set_smem(k_join<false, true>, kJoinSmemBytes); set_smem(k_join<true, true>, kJoinSmemBytes);
    init_probe_mode<0>(); init_probe_mode<1>(); init_probe_mode<2>(); init_probe_mode<3>();
    init_probe_mode<4>(); init_probe_mode<5>(); init_probe_mode<6>(); init_probe_mode<7>();
    set_smem(k_probe_staged<0>, kProbeWarps * kStagedSmemPerWarp); set_smem(k_probe_staged<1>, kProbeWarps * kStagedSmemPerWarp);
    set_smem(k_probe_staged<4>, kProbeWarps * kStagedSmemPerWarp); set_smem(k_probe_staged<5>, kProbeWarps * kStagedSmemPerWarp);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.occ_scatter, k_scatter<1, 1>, kScatterThreads, kScatterSmem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.occ_join, k_join<false>, kJoinThreads, kJoinSmemBytes));
    g.occ_scatter = std::max(g.occ_scatter, 1);
    g.occ_join = std::max(g.occ_join, 1);
    memset(&g.last, 0, sizeof(g.last));
    g.inited = true;
}

static int ilog2_u64(uint64_t v) {
    int l = 0;
    while ((1ull << (l + 1)) <= v) l++;
    return l;
}

static int check_args_impl(const bloom_filter_args_t* a, bool print) {
    // bloom_filter.c:26-34
    if (a->m == 0 || (a->m & (a->m - 1)) != 0) {
        if (print) printf("m must be a power of 2");
        return 1;
    }
    if (a->m > (1ull << 32)) {  // mod_m()/size are uint32_t in the reference (bloom_filter.c:60-63,75)
        if (print) printf("m must be at most 2^32");
        return 4;
    }
    if (a->variant != BASIC) {
        if (a->B == 0 || (a->B & (a->B - 1)) != 0) {
            if (print) printf("B must be a power 2");
            return 2;
        }
        if (a->B < 8 || a->m % a->B != 0) {  // B/8 bytes per block (bloom_filter.c:129)
            if (print) printf("m must be a multiple of B");
            return 3;
        }
    }
    return 0;
}

static BloomParams make_bloom(const bloom_filter_args_t* a, uint32_t seed, uint32_t* filter) {
    BloomParams bp;
    memset(&bp, 0, sizeof(bp));
    bp.filter = filter;
    bp.k = (uint32_t)a->k;
    bp.seed = seed;
    bp.blocked = a->variant == BLOCKED ? 1u : 0u;
    if (bp.blocked) {
        bp.size_mask = (uint32_t)(a->B - 1);
        bp.nblocks_mask = (uint32_t)(a->m / a->B - 1);
        bp.log2B = (uint32_t)ilog2_u64(a->B);
    } else {
        bp.size_mask = (uint32_t)(a->m - 1);  // m == 2^32 -> 0xFFFFFFFF
    }
    bp.nranges = 1;
    bp.range_shift = 0;
    bp.range_id = 0;
    bp.skew = nullptr;
    return bp;
}

// number of filter range passes: keep the actively probed part of the filter L2-resident (<= 64 MiB).
// Only valid when all k bits of a key fall into one range: k <= 1 or BLOCKED.
static int pick_ranges(const bloom_filter_args_t* a) {
    if (a->variant == BASIC && a->k > 1) return 1;
    uint64_t bytes = a->m / 8;
    int nr = 1;
    if (g.range_passes_override > 0) nr = std::min(g.range_passes_override, 64);
    else
        while ((bytes / nr) > (64ull << 20)) nr <<= 1;
    // must be a power of two and leave ranges >= one block / one word
    while (nr & (nr - 1)) nr &= nr - 1;
    uint64_t min_range_bits = a->variant == BLOCKED ? std::max<uint64_t>(a->B, 32) : 32;
    while (nr > 1 && a->m / nr < min_range_bits) nr >>= 1;
    return std::max(nr, 1);
}
static void set_ranges(BloomParams& bp, const bloom_filter_args_t* a, int nranges) {
    bp.nranges = (uint32_t)nranges;
    bp.range_shift = (uint32_t)(ilog2_u64(a->m) - ilog2_u64((uint64_t)nranges));
}

// total partition bits: the smallest fan-out whose R partitions fit one shared-memory table (the reference's
// NUM_RADIX_BITS, prj_params.h:15-17, is a compile-time constant; here HWBRJ_RADIX_BITS / hwbrj_set_radix_bits override)
static int pick_bits(uint64_t nR, int gbits) {
    int b = 0;
    if (g.radix_bits_override > 0) b = std::min(g.radix_bits_override, (int)kMaxRadixBits);
    else
        while (b < kMaxRadixBits && (nR >> b) > (uint64_t)(kTableCap * 3 / 4)) b++;
    // one scatter pass sorts a tile into at most 2^7 bins: NUM_PASSES = 1 caps the fan-out there (bigger partitions are
    // joined in several table rounds), as the reference's single pass has to handle all NUM_RADIX_BITS at once
    if (g.passes_override == 1) b = std::min(b, (int)kMaxLevelBits);
    return std::max(b, gbits);
}
// level-2 bits (the reference's NUM_PASSES, prj_params.h:20-22; HWBRJ_NUM_PASSES / hwbrj_set_num_passes override):
// one pass while the fan-out fits one scatter pass, else two passes of about equal width
static int pick_b2(int bits, int gbits) {
    int passes = g.passes_override;
    if (passes <= 0 || bits > kMaxLevelBits) passes = bits > kMaxLevelBits ? 2 : (passes <= 0 ? 1 : passes);
    if (passes == 1 || bits < 2) return 0;
    int b2 = std::min(bits / 2, bits - gbits);  // the owner is a prefix of the level-1 bin
    if (bits - b2 > kMaxLevelBits) b2 = bits - kMaxLevelBits;
    if (b2 > kMaxLevelBits) b2 = kMaxLevelBits;  // (then the fan-out shrinks: bits are capped by the caller's override)
    return std::max(b2, 0);
}

// ---- per-launch trace (HWBRJ_TRACE=1) ------------------------------------------------------------------------------
static bool g_tracing = false;  // set by run_join for the duration of a traced, non-captured join
static void TR(const char* name) {
    if (!g_tracing) return;
    if (g.tr_n == g.tr_ev.size()) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        g.tr_ev.push_back(e);
        g.tr_name.push_back(name);
    }
    g.tr_name[g.tr_n] = name;
    CK(cudaEventRecord(g.tr_ev[g.tr_n++], g.stream));
}
static void trace_print(int rank) {
    for (size_t i = 1; i < g.tr_n; i++) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, g.tr_ev[i - 1], g.tr_ev[i]));
        fprintf(stderr, "[hwbrj trace] rank %d  %-28s %8.3f ms\n", rank, g.tr_name[i], ms);
    }
    g.tr_n = 0;
}

// ---- kernel dispatch on the compile-time specialisations ---------------------------------------------------------------
// K2 launch: (blocked, k == 1, ranged)
static int launch_probe_mode(int mode, const uint2* in, uint64_t n, const unsigned long long* n_ptr, const BloomParams& bp,
                             uint2* out, unsigned long long* cursor) {
    if (g.probe_staged && bp.k >= 2u && !(mode & 2)) {  // staged probe for k >= 2 (HWBRJ_PROBE_STAGED=0 switches it off)
        const int smem = kProbeWarps * kStagedSmemPerWarp;
        const int grid = g.sms * (g.probe_ctas_per_sm ? g.probe_ctas_per_sm : 4);
        switch (mode) {
            case 0: k_probe_staged<0><<<grid, kProbeWarps * 32, smem, g.stream>>>(in, n, n_ptr, bp, g.d_crc, out, cursor); return 1;
            case 1: k_probe_staged<1><<<grid, kProbeWarps * 32, smem, g.stream>>>(in, n, n_ptr, bp, g.d_crc, out, cursor); return 1;
            case 4: k_probe_staged<4><<<grid, kProbeWarps * 32, smem, g.stream>>>(in, n, n_ptr, bp, g.d_crc, out, cursor); return 1;
            case 5: k_probe_staged<5><<<grid, kProbeWarps * 32, smem, g.stream>>>(in, n, n_ptr, bp, g.d_crc, out, cursor); return 1;
            default: break;
        }
    }
    // measured on B200: ProbeShape::CTAS CTAs/SM beat the occupancy maximum (more L1 left for the loads in flight).
    // With a sample of S (bp.skew) both shapes are launched and the device-side verdict lets one of them return at once;
    // without one only the shape for dense survivors runs.
#define HWBRJ_PROBE_LAUNCH(M, SH)                                                                                      \
    {                                                                                                                  \
        const int ctas = g.probe_ctas_per_sm ? g.probe_ctas_per_sm : std::min(g.occ_probe[M][SH], ProbeShape<SH>::CTAS); \
        k_probe_compact<M, SH><<<g.sms * ctas, kProbeWarps * 32, kProbeWarps * ProbeShape<SH>::SMEM_PER_WARP, g.stream>>>( \
            in, n, n_ptr, bp, g.d_crc, out, cursor);                                                                   \
    }
#define HWBRJ_PROBE_CASE(M)                                \
    case M:                                                \
        if (bp.skew != nullptr) HWBRJ_PROBE_LAUNCH(M, 0)   \
        HWBRJ_PROBE_LAUNCH(M, 1)                           \
        break;
    switch (mode) {
        HWBRJ_PROBE_CASE(0) HWBRJ_PROBE_CASE(1) HWBRJ_PROBE_CASE(2) HWBRJ_PROBE_CASE(3)
        HWBRJ_PROBE_CASE(4) HWBRJ_PROBE_CASE(5) HWBRJ_PROBE_CASE(6) HWBRJ_PROBE_CASE(7)
        default: die("bad probe mode %d", mode);
    }
#undef HWBRJ_PROBE_LAUNCH
#undef HWBRJ_PROBE_CASE
    return bp.skew != nullptr ? 2 : 1;
}

// all range passes of the S-side probe; returns the number of kernel launches
static int run_probe(const uint2* dS, uint64_t nS, const unsigned long long* n_ptr, BloomParams bp, int nranges, uint2* out,
                     unsigned long long* cursor) {
    const int base_mode = (bp.blocked ? 1 : 0) | (bp.k == 1u ? 2 : 0);
    bp.nranges = (uint32_t)nranges;
    if (nranges == 1) {
        const int l = launch_probe_mode(base_mode, dS, nS, n_ptr, bp, out, cursor);
        TR("K2 probe");
        return l;
    }
    int launches = 0;
    for (int r = 0; r < nranges; r++) {
        bp.range_id = (uint32_t)r;
        launches += launch_probe_mode(base_mode | 4, dS, nS, n_ptr, bp, out, cursor);
        TR("K2 probe (range pass)");
    }
    return launches;
}

static void launch_hist(int pmode, const uint2* in, uint64_t n, const unsigned long long* n_ptr, const BloomParams& bp,
                        uint32_t* hist, const PartFn& pf) {
    const int smem = (int)(((1u << pf.bits) + kCrcSmemWords) * 4);
    const int grid = g.sms * 2;
    switch (pmode) {
        case 0: k_build_hist<false, 0><<<grid, 1024, smem, g.stream>>>(in, n, n_ptr, bp, g.d_crc, hist, pf); break;
        case 1: k_build_hist<false, 1><<<grid, 1024, smem, g.stream>>>(in, n, n_ptr, bp, g.d_crc, hist, pf); break;
        default: k_build_hist<false, 2><<<grid, 1024, smem, g.stream>>>(in, n, n_ptr, bp, g.d_crc, hist, pf); break;
    }
}

template <int LEVEL>
static void launch_scatter_l(int pmode, const uint2* in, const PeerBufs& stages, uint2* out, const unsigned long long* n_ptr,
                             uint64_t n, uint32_t* cursor, const PartFn& pf, uint32_t nbins, uint32_t seg_lo, uint32_t seg_hi,
                             uint32_t G, const uint32_t* abort_flag) {
    // several GPUs: the level-2 pull is NVLink-bound; two CTAs per SM leave room for the kernels it overlaps with
    const int grid = g.sms * ((LEVEL == 2 && G > 1) ? std::min(g.occ_scatter, 2) : g.occ_scatter);
    const uint32_t* tiles = g.tiles.as<uint32_t>();
    const uint32_t* seg_start = g.segs.as<uint32_t>();
    const uint32_t* seg_cnt = seg_start + (1u << kMaxLevelBits);
#define HWBRJ_SCATTER(PM)                                                                                              \
    k_scatter<LEVEL, PM><<<grid, kScatterThreads, kScatterSmem, g.stream>>>(in, stages, out, n_ptr, n, tiles, seg_start,  \
                                                                           seg_cnt, cursor, pf, g.d_crc, nbins, seg_lo,  \
                                                                           seg_hi, G, abort_flag)
    switch (pmode) {
        case 0: HWBRJ_SCATTER(0); break;
        case 1: HWBRJ_SCATTER(1); break;
        default: HWBRJ_SCATTER(2); break;
    }
#undef HWBRJ_SCATTER
}

static void launch_barrier(const Fab& f, Control* ctrl) {
    // 2 s at the SM clock: a peer that never arrives ends the wait instead of hanging the GPU
    k_barrier<<<1, 32, 0, g.stream>>>(f.flags, (uint32_t)f.world, (uint32_t)f.rank, f.epoch, &ctrl->err,
                                      2ll * g.clock_khz * 1000ll);
}

// Histogram rows of all ranks -> offsets; level-1 scatter of this rank's chunk into its staging buffer (purely local).
// Afterwards the level-2 pass (launch_level2) pulls the segments of the OWNED level-1 bins from the staging buffers of all
// ranks. `hist` is this rank's row (already computed), `rows` where every rank's row is gathered (world > 1), `stages` the
// staging buffers, `capacity` bounds a chunk and an owner's share. Returns true when a level-2 pass has to follow (always
// with several GPUs: with b2 == 0 it degenerates to gathering the owned bins from the peers).
static bool partition_front(const Fab& f, int pmode, const PartFn& pf, const uint2* in, uint64_t n,
                            const unsigned long long* n_ptr, uint32_t* hist, const PeerPtrs& rows, const PeerBufs& stages,
                            uint64_t capacity, uint32_t* off, unsigned long long* n_own, Control* ctrl, int& launches) {
    const uint32_t P = 1u << pf.bits;
    const uint32_t b1 = pf.bits - pf.b2;
    const bool dist = f.world > 1;
    const uint32_t* hist_all = hist;
    if (dist) {
        k_push_rows<<<dim3(4, f.world), 256, 0, g.stream>>>(rows, (uint32_t)f.world, (uint32_t)f.rank, hist, P);
        TR("push histogram rows");
        launch_barrier(f, ctrl);
        TR("barrier");
        hist_all = reinterpret_cast<const uint32_t*>(rows.p[f.rank]);
        launches += 2;
    }
    uint32_t* seg_start = g.segs.as<uint32_t>();
    k_scan_dist<<<1, 1024, 0, g.stream>>>(hist_all, (uint32_t)f.world, (uint32_t)f.rank, P, pf.b2, capacity, off,
                                          g.cur1.as<uint32_t>(), g.cur2.as<uint32_t>(), g.tiles.as<uint32_t>(), seg_start,
                                          seg_start + (1u << kMaxLevelBits), n_own, &ctrl->abort);
    TR("K3 scan");
    launches++;
    launch_scatter_l<1>(pmode, in, stages, stages.buf[f.rank], n_ptr, n, g.cur1.as<uint32_t>(), pf, 1u << b1, 0u, 0u, 1u,
                        dist ? &ctrl->abort : nullptr);
    TR("K4 scatter level 1");
    launches++;
    if (dist) {
        launch_barrier(f, ctrl);  // every rank's staging buffer is complete
        TR("barrier");
        launches++;
    }
    return dist || pf.b2 != 0;
}

// level-2 pass over the owned level-1 bins [lb0, lb1) into t2
static void launch_level2(const Fab& f, int pmode, const PartFn& pf, const PeerBufs& stages, uint2* t2, uint32_t lb0,
                          uint32_t lb1, Control* ctrl, int& launches) {
    const uint32_t G = (uint32_t)f.world;
    launch_scatter_l<2>(pmode, nullptr, stages, t2, nullptr, 0, g.cur2.as<uint32_t>(), pf, 1u << pf.b2, lb0 * G, lb1 * G, G,
                        G > 1 ? &ctrl->abort : nullptr);
    TR(G > 1 ? "K4 scatter level 2 (pull)" : "K4 scatter level 2");
    launches++;
}

static void ensure_workspace(uint64_t capR, uint64_t nS, uint64_t capS, const bloom_filter_args_t* args, bool dist) {
    const size_t P = 1u << kMaxRadixBits;
    if (args && !dist) g.filter.ensure(std::max<uint64_t>(args->m / 8, 16));
    g.histR.ensure(P * 4);
    g.histS.ensure(P * 4);
    g.offR.ensure((P + 1) * 4);
    g.offS.ensure((P + 1) * 4);
    g.cur1.ensure(((size_t)1 << kMaxLevelBits) * 4);
    g.cur2.ensure(P * 4);
    g.tiles.ensure((((size_t)1 << kMaxLevelBits) + 1) * 4);
    g.segs.ensure(((size_t)2 << kMaxLevelBits) * 4);
    g.work.ensure(4 * (P + 1) * 4);                        // up to 4 groups of partitions, each with its own work list
    g.work_part.ensure(4 * (P + capS / kSChunk + 2) * 4);  // one entry per join work item
    g.ctrl.ensure(sizeof(Control));
    if (!dist) g.rt1.ensure(std::max<uint64_t>(capR, 1) * 8 + 64);
    g.rp.ensure(std::max<uint64_t>(capR, 1) * 8 + 64);
    g.sc.ensure(std::max<uint64_t>(nS, 1) * 8 + 64);
    if (!dist) g.st1.ensure(std::max<uint64_t>(nS, 1) * 8 + 64);
    else g.s2.ensure(std::max<uint64_t>(capS, 1) * 8 + 64);
}

// how the join partitions: 0 radix bits of the key (the reference's clustering), 1 / 2 the filter-slice index of a
// BASIC (k <= 1) / BLOCKED filter, which lets K1' build the filter in shared memory and -- across GPUs -- makes every rank
// the builder of exactly its slice of the replicated filter
static int pick_mode(const bloom_filter_args_t* a, int bits, int world) {
    if (!a || !g.hash_partition || bits < 1) return 0;
    const bool basic = a->variant == BASIC;
    if (basic && a->k > 1) return 0;  // a key's bits spread over the whole filter: not sliceable
    const uint64_t units = basic ? a->m : a->m / a->B;  // what is sliced: bits, or whole blocks
    if (ilog2_u64(units) < bits) return 0;
    const uint64_t slice_bits = a->m >> bits;
    if (slice_bits < 32 || slice_bits / 8 > (uint64_t)kFilterSliceSmemMax) return 0;
    // one GPU: for small filters the plain global atomics are as fast (C0: 1.71 vs 1.80 ms) and the radix table index has
    // shorter chains, so the slice build is used for filters beyond L2-friendly sizes unless forced
    if (world == 1 && g.hash_partition == 1 && a->m / 8 <= (32ull << 20)) return 0;
    return basic ? 1 : 2;
}

__global__ void k_export_row(Control* c, int has_filter) {
    if (threadIdx.x == 0) {
        c->row[0] = c->acc.matches;
        c->row[1] = c->acc.cpair;
        c->row[2] = c->acc.crpay;
        c->row[3] = c->acc.cspay;
        c->row[4] = c->acc.ckey;
        c->row[5] = has_filter ? c->survivors : 0ull;
        c->row[6] = (unsigned long long)c->abort | ((unsigned long long)c->err << 8);
        c->row[7] = 0ull;
        for (int i = 0; i < 8; i++) c->out[i] = c->row[i];  // one GPU: the sum is the row itself
    }
}
__global__ void k_copy8(const unsigned long long* src, unsigned long long* dst) {
    if (threadIdx.x < 8) dst[threadIdx.x] = src[threadIdx.x];
}

// S arriving from the host in chunks: chunk c (chunk_tuples tuples, the last one shorter) is complete when ev[c] fires
struct SFeed {
    uint64_t chunk_tuples;
    int nchunks;
    cudaEvent_t* ev;
};

// The join of this rank's chunks dR / dS (one GPU: the whole relations). args == nullptr: plain radix join.
// f describes the participating GPUs (f.world == 1: the plain workspace). r_total = |R| over all ranks (sizes the fan-out).
// Modes: kSync enqueues, waits and fills st; kEnqueue records the timing events but does not wait (collect_join() does:
// one host thread drives several GPUs); kCapture records no events and leaves the eight result words {matches, cpair,
// crpay, cspay, ckey, filtered, flags, 0}, summed over the ranks, in d_async_out (capturable in a CUDA graph).
// Returns 0, or -2 when a receive buffer was too small / a peer timed out (results invalid).
enum JoinMode { kSync, kEnqueue, kCapture };
static int collect_join(hwbrj_stats_t& st, bool has_filter);

static int run_join(const Fab& f, const uint2* dR, uint64_t nR, const uint2* dS, uint64_t nS,
                    const bloom_filter_args_t* args, uint64_t r_total, hwbrj_stats_t& st, const SFeed* feed = nullptr,
                    JoinMode mode = kSync, unsigned long long* d_async_out = nullptr) {
    const bool dist = f.world > 1;
    auto rec = [&](int e) {
        if (mode != kCapture) CK(cudaEventRecord(g.ev[e], g.stream));
    };
    // 32-bit tuple indices; the slack keeps "index + one batch of loads" from wrapping in the kernels
    if (nR >= (1ull << 32) - (1ull << 20) || nS >= (1ull << 32) - (1ull << 20) || f.cap_r >= (1ull << 32) - (1ull << 20) ||
        f.cap_s >= (1ull << 32) - (1ull << 20))
        die("relations of 2^32 - 2^20 or more tuples per GPU are not supported");
    if (args && check_args_impl(args, true)) die("invalid Bloom filter arguments");
    if (args && dist && args->m / 8 > f.filter_bytes) die("the filter exceeds the size the GPU group was created for");
    ensure_workspace(dist ? f.cap_r : nR, nS, dist ? f.cap_s : nS, args, dist);
    const int bits = pick_bits(r_total, f.gbits);
    const int b2 = pick_b2(bits, f.gbits);
    const int pmode = pick_mode(args, bits, f.world);
    const uint32_t P = 1u << bits, PL = P / (uint32_t)f.world;
    int launches = 0;
    Control* ctrl = g.ctrl.as<Control>();
    PartFn pf;
    memset(&pf, 0, sizeof(pf));
    pf.bits = (uint32_t)bits;
    pf.b2 = (uint32_t)b2;
    pf.seed = 42u;
    if (pmode == 1) {
        pf.size_mask = (uint32_t)(args->m - 1);
        pf.hshift = (uint32_t)(ilog2_u64(args->m) - bits);
    } else if (pmode == 2) {
        pf.size_mask = (uint32_t)(args->m / args->B - 1);
        pf.hshift = (uint32_t)(ilog2_u64(args->m / args->B) - bits);
    }
    // level-1 staging buffers and filters: the symmetric memory of the GPU group, or the plain workspace
    PeerBufs stageR = f.stageR, stageS = f.stageS;
    PeerPtrs filt = f.filter;
    uint32_t* my_filter = nullptr;
    if (!dist) {
        stageR.buf[0] = g.rt1.as<uint2>();
        stageS.buf[0] = g.st1.as<uint2>();
        filt.p[0] = g.filter.p;
    }
    if (args) my_filter = reinterpret_cast<uint32_t*>(filt.p[f.rank]);
    // non-sliceable filter on several GPUs: every rank inserts its R chunk into a full-size partial, then OR-combine
    uint32_t* insert_filter = (args && pmode == 0 && dist) ? reinterpret_cast<uint32_t*>(f.partial.p[f.rank]) : my_filter;

    // ---- untimed set-up (the reference allocates and zeroes its filter before the timed region, :1583) ----
    rec(0);
    if (args && pmode == 0) CK(cudaMemsetAsync(insert_filter, 0, std::max<uint64_t>(args->m / 8, 16), g.stream));
    CK(cudaMemsetAsync(g.histR.p, 0, P * 4, g.stream));
    CK(cudaMemsetAsync(g.histS.p, 0, P * 4, g.stream));
    CK(cudaMemsetAsync(g.ctrl.p, 0, sizeof(Control), g.stream));

    // ---- timed region ----------------------------------------------------------------------------------------
    rec(1);
    g_tracing = g.trace && mode != kCapture;
    g.tr_n = 0;
    TR("start");
    BloomParams bp;
    memset(&bp, 0, sizeof(bp));
    int nranges = 1;
    if (args) {
        bp = make_bloom(args, 42u, insert_filter);  // seed 42: parallel_radix_join_bloom.c:1583,1823
        nranges = pick_ranges(args);
        set_ranges(bp, args, nranges);
    }
    // K1: partition histogram of this rank's R chunk; the radix variant inserts into the filter on the way
    if (args && pmode == 0) {
        const int smem = (int)((P + kCrcSmemWords) * 4);
        for (int r = 0; r < nranges; r++) {
            bp.range_id = (uint32_t)r;
            k_build_hist<true, 0><<<g.sms * 2, 1024, smem, g.stream>>>(dR, nR, nullptr, bp, g.d_crc, g.histR.as<uint32_t>(), pf);
            TR("K1 insert + histogram R");
            launches++;
        }
    } else {
        launch_hist(pmode, dR, nR, nullptr, bp, g.histR.as<uint32_t>(), pf);
        TR("K1 histogram R");
        launches++;
    }
    rec(2);
    // Several GPUs: the owned level-1 bins are handled in `parts` groups on two streams, so that the NVLink pull of
    // group h+1 (inbound traffic) runs while group h is consumed -- on the R side by the filter-slice build, whose stores
    // into the peers' filters are outbound traffic; on the S side by the join. One GPU: one group, one stream.
    const uint32_t P1L = (1u << (bits - b2)) / (uint32_t)f.world;
    uint32_t parts = 1;
    if (dist) {
        // Pipelining pays while a group's kernels are bandwidth-bound. Measured at C1 (profiles/r2_bench_c1_8gpu*.json):
        // 8 GPUs own 16 M tuples each -- four groups 1.82 ms, two 1.71 ms, one 1.69 ms (a quarter of the slice build is a
        // latency-bound launch); 2 GPUs own 64 M each and gain from four groups.
        uint32_t want = (uint32_t)g.dist_parts;
        if (want == 0) {
            const uint64_t per_rank = r_total / (uint64_t)f.world;
            want = per_rank >= (48ull << 20) ? 4u : per_rank >= (24ull << 20) ? 2u : 1u;
        }
        parts = std::max(1u, std::min(want, P1L));
    }
    cudaStream_t main_stream = g.stream;
    auto stream_of = [&](uint32_t h) { return (parts > 1 && (h & 1u)) ? g.side_stream : main_stream; };
    auto part_begin = [&](uint32_t h) {  // group h starts on its stream after the pull of group h-1 has finished
        if (parts == 1) return;
        if (h == 0) {
            CK(cudaEventRecord(g.ev_ov[0], main_stream));  // fork: the side stream joins the work of this join
            CK(cudaStreamWaitEvent(g.side_stream, g.ev_ov[0], 0));
        } else {
            CK(cudaStreamWaitEvent(stream_of(h), g.ev_ov[1 + ((h - 1) & 1u)], 0));
        }
        g.stream = stream_of(h);
    };
    auto part_pulled = [&](uint32_t h) {
        if (parts > 1) CK(cudaEventRecord(g.ev_ov[1 + (h & 1u)], g.stream));
    };
    auto parts_end = [&]() {  // join: everything of both streams is ordered before what follows on the main stream
        if (parts == 1) return;
        CK(cudaEventRecord(g.ev_ov[3], g.side_stream));
        CK(cudaStreamWaitEvent(main_stream, g.ev_ov[3], 0));
        g.stream = main_stream;
    };
    auto launch_k1p = [&](const uint2* Rp, uint32_t p0, uint32_t np) {  // K1' over the local partitions [p0, p0 + np)
        const uint32_t slice_words = (uint32_t)((args->m >> bits) / 32);
        const uint32_t nbuf = 2u * slice_words * 4u <= 96u * 1024u ? 2u : 1u;
        const int smem = (int)(nbuf * slice_words * 4u);
        const int grid = g.sms * (smem <= 48 * 1024 ? 4 : 1);
        const uint32_t gbase = (uint32_t)f.rank * PL + p0;
        const uint32_t* roff = g.offR.as<uint32_t>() + p0;
#define HWBRJ_K1P(BL, PE)                                                                                              \
    k_filter_from_parts<BL, PE><<<grid, 512, smem, g.stream>>>(Rp, roff, np, gbase, slice_words, bp, g.d_crc, filt,      \
                                                               (uint32_t)f.world, nbuf)
        if (pmode == 2) {
            if (dist) HWBRJ_K1P(true, true); else HWBRJ_K1P(true, false);
        } else {
            if (dist) HWBRJ_K1P(false, true); else HWBRJ_K1P(false, false);
        }
#undef HWBRJ_K1P
        TR("K1' filter slices");
        launches++;
    };
    const uint2* Rp = stageR.buf[f.rank];
    const bool r_level2 = partition_front(f, pmode, pf, dR, nR, nullptr, g.histR.as<uint32_t>(), f.histR, stageR,
                                          dist ? f.cap_r : nR, g.offR.as<uint32_t>(), &ctrl->n_own_r, ctrl, launches);
    if (r_level2) Rp = g.rp.as<uint2>();
    rec(3);
    if (args) bp.filter = my_filter;
    const bool slice_build = args && pmode != 0;  // K1': the filter, slice by slice, from the partitioned R
    for (uint32_t h = 0; h < parts; h++) {
        const uint32_t lb0 = h * P1L / parts, lb1 = (h + 1) * P1L / parts;
        part_begin(h);
        if (r_level2) launch_level2(f, pmode, pf, stageR, g.rp.as<uint2>(), lb0, lb1, ctrl, launches);
        part_pulled(h);
        if (parts == 1) rec(11);  // one group: the level-2 pass ends here, the slice build follows
        if (slice_build) launch_k1p(Rp, lb0 << b2, (lb1 - lb0) << b2);
    }
    parts_end();
    if (args) {
        if (pmode == 0 && dist) {  // partial filters are complete on every rank (the barrier after the level-1 scatter)
            const uint64_t n16 = std::max<uint64_t>(args->m / 8, 16) / 16;
            const uint64_t per = n16 / (uint64_t)f.world;
            k_filter_or_bcast<<<g.sms * 4, 256, 0, g.stream>>>(f.partial, f.filter, (uint32_t)f.world, per * f.rank,
                                                               f.rank == f.world - 1 ? n16 - per * f.rank : per);
            TR("filter OR + broadcast");
            launches++;
        }
        if (dist) {
            launch_barrier(f, ctrl);  // the replicated filter is complete on every rank
            TR("barrier");
            launches++;
        }
    }
    rec(4);
    const uint2* Sin = dS;
    const unsigned long long* n_dev = nullptr;
    if (args) {
        const bool sample = g.probe_adaptive && nS > 0;
        if (sample) bp.skew = &ctrl->skew;
        if (sample && !feed) {
            k_probe_sample<<<1, 1024, 0, g.stream>>>(dS, nS, bp, g.d_crc, &ctrl->skew);
            TR("sample of S");
            launches++;
        }
        if (feed) {  // probe every chunk as soon as its host->device copy has completed
            for (int c = 0; c < feed->nchunks; c++) {
                const uint64_t off = (uint64_t)c * feed->chunk_tuples;
                const uint64_t cnt = std::min<uint64_t>(feed->chunk_tuples, nS - off);
                CK(cudaStreamWaitEvent(g.stream, feed->ev[c], 0));
                if (sample && c == 0) {  // the first chunk stands for the relation
                    k_probe_sample<<<1, 1024, 0, g.stream>>>(dS, cnt, bp, g.d_crc, &ctrl->skew);
                    launches++;
                }
                launches += run_probe(dS + off, cnt, nullptr, bp, nranges, g.sc.as<uint2>(), &ctrl->survivors);
            }
        } else {
            launches += run_probe(dS, nS, nullptr, bp, nranges, g.sc.as<uint2>(), &ctrl->survivors);
        }
        Sin = g.sc.as<uint2>();
        n_dev = &ctrl->survivors;
    }
    rec(5);  // ms_probe = the K2 launches only
    // partition histogram of the tuples that go on to the join (survivors, or all of S without a filter)
    launch_hist(pmode, Sin, nS, n_dev, bp, g.histS.as<uint32_t>(), pf);
    TR("histogram S side");
    launches++;
    // one GPU with a filter: sc -> st1 -> sc ; without: dS -> st1 -> sc ; several GPUs: sc -> own staging -> (pull) -> s2
    const uint2* Sp = stageS.buf[f.rank];
    const bool s_level2 = partition_front(f, pmode, pf, Sin, nS, n_dev, g.histS.as<uint32_t>(), f.histS, stageS,
                                          dist ? f.cap_s : nS, g.offS.as<uint32_t>(), &ctrl->n_own_s, ctrl, launches);
    uint2* s_out = dist ? g.s2.as<uint2>() : g.sc.as<uint2>();
    if (s_level2) Sp = s_out;
    rec(6);
    const size_t work_stride = (size_t)PL + (dist ? f.cap_s : nS) / kSChunk + 2;
    for (uint32_t h = 0; h < parts; h++) {
        const uint32_t lb0 = h * P1L / parts, lb1 = (h + 1) * P1L / parts;
        const uint32_t p0 = parts == 1 ? 0u : lb0 << b2, np = parts == 1 ? PL : (lb1 - lb0) << b2;
        part_begin(h);
        if (s_level2) launch_level2(f, pmode, pf, stageS, s_out, parts == 1 ? 0u : lb0, parts == 1 ? P1L : lb1, ctrl, launches);
        part_pulled(h);
        if (parts == 1) rec(12);
        // per-partition build + probe of the group's partitions (its own work list and item counter)
        uint32_t* work_off = g.work.as<uint32_t>() + (size_t)h * (PL + 1);
        uint32_t* work_part = g.work_part.as<uint32_t>() + (size_t)h * work_stride;
        const uint32_t* roff = g.offR.as<uint32_t>() + p0;
        const uint32_t* soff = g.offS.as<uint32_t>() + p0;
        k_worklist<<<1, 1024, 0, g.stream>>>(roff, soff, np, work_off, work_part);
        TR("work list");
        launches++;
        if (pmode != 0)
            k_join<true><<<g.sms * g.occ_join, kJoinThreads, kJoinSmemBytes, g.stream>>>(
                Rp, roff, Sp, soff, work_off, work_part, np, (uint32_t)bits, &ctrl->item_counter[h], &ctrl->acc);
        else
            k_join<false><<<g.sms * g.occ_join, kJoinThreads, kJoinSmemBytes, g.stream>>>(
                Rp, roff, Sp, soff, work_off, work_part, np, (uint32_t)bits, &ctrl->item_counter[h], &ctrl->acc);
        TR("K5 join");
        launches++;
    }
    parts_end();
    g.last_Rp = parts > 1 ? nullptr : Rp;
    g.last_Sp = parts > 1 ? nullptr : Sp;
    g.last_P = PL;
    g.last_bits = (uint32_t)bits;
    g.last_hash = pmode != 0;
    rec(7);
    // the result words; several GPUs: all-gather of the rows over NVLink, barrier, local sum
    k_export_row<<<1, 32, 0, g.stream>>>(ctrl, args ? 1 : 0);
    launches++;
    if (dist) {
        k_push_rows<<<dim3(1, f.world), 32, 0, g.stream>>>(f.rows, (uint32_t)f.world, (uint32_t)f.rank,
                                                           reinterpret_cast<const uint32_t*>(ctrl->row), 16u);
        launch_barrier(f, ctrl);  // also: every rank has finished reading its receive buffers -> the next join may start
        k_reduce_rows<<<1, 32, 0, g.stream>>>(reinterpret_cast<const unsigned long long*>(f.rows.p[f.rank]), (uint32_t)f.world,
                                              ctrl->out);
        TR("result rows: push + barrier + sum");
        launches += 3;
    }
    rec(8);
    g_tracing = false;
    st.kernel_launches = launches;
    st.phase_split = parts == 1 ? 1 : 0;
    st.radix_bits = bits;
    st.range_passes = nranges;
    st.n_gpus = f.world;
    if (mode == kCapture) {
        k_copy8<<<1, 32, 0, g.stream>>>(ctrl->out, d_async_out);
        st.kernel_launches++;
        return 0;
    }
    if (mode == kEnqueue) return 0;
    return collect_join(st, args != nullptr);
}

// waits for the join enqueued on this device's stream and fills the result and timing fields of st
static int collect_join(hwbrj_stats_t& st, bool has_filter) {
    Control h;
    CK(cudaMemcpyAsync(&h, g.ctrl.p, sizeof(Control), cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaGetLastError());
    if (g.trace) trace_print(st.n_gpus > 1 ? g.dev : 0);
    auto ms = [&](int a, int b) {
        float v = 0;
        CK(cudaEventElapsedTime(&v, g.ev[a], g.ev[b]));
        return v;
    };
    st.matches = (int64_t)h.out[0];
    st.checksum_pair = h.out[1];
    st.checksum_rpay = h.out[2];
    st.checksum_spay = h.out[3];
    st.checksum_key = h.out[4];
    st.filtered = has_filter ? (int64_t)h.out[5] : -1;
    st.ms_memset = ms(0, 1);
    st.ms_total = ms(1, 8);
    if (st.phase_split) {  // one group of partitions on one stream: every phase is a contiguous piece of the stream
        st.ms_build = ms(1, 2) + ms(11, 4);  // histogram of R (+ insert), filter slices from the partitioned R
        st.ms_part_r = ms(2, 11);            // offsets + level-1 + level-2 scatter of R
        st.ms_probe = ms(4, 5);
        st.ms_part_s = ms(5, 12);            // histogram + offsets + level-1 + level-2 scatter of the survivors
        st.ms_join = ms(12, 8);              // work list + per-partition build/probe + result words
    } else {  // several GPUs: the level-2 pulls overlap the slice build (R side) and the join (S side) on two streams
        st.ms_build = ms(1, 2);
        st.ms_part_r = ms(2, 4);   // routing + level-2 pull + slice build + filter exchange
        st.ms_probe = ms(4, 5);
        st.ms_part_s = ms(5, 6);   // histogram + offsets + level-1 scatter + barrier
        st.ms_join = ms(6, 8);     // level-2 pull + join, pipelined
    }
    st.owned_r = h.n_own_r;
    st.owned_s = h.n_own_s;
    st.d2h_bytes += sizeof(Control);
    return h.out[6] ? -2 : 0;
}

static void print_reference_lines(const hwbrj_stats_t& st, uint64_t nS, bool bloom_line) {
    if (g_quiet) return;
    // stdout contract of parallel_radix_join_bloom.c:1253 and print_timing (:1510-1547)
    if (bloom_line) fprintf(stdout, "S-tuples after filter: %d\n", (int)st.filtered);
    const double total_us = st.ms_total * 1000.0;
    const double part_us = (st.ms_build + st.ms_part_r + st.ms_probe + st.ms_part_s) * 1000.0;
    const double join_us = st.ms_join * 1000.0;
    const double mhz = g.clock_khz / 1000.0;
    unsigned long long cyc_total = (unsigned long long)(total_us * mhz);
    unsigned long long cyc_part = (unsigned long long)(part_us * mhz);
    unsigned long long cyc_build = (unsigned long long)(st.ms_build * 1000.0 * mhz);
    fprintf(stdout, "RUNTIME TOTAL, BUILD, PART (cycles): \n");
    fprintf(stdout, "%llu \t %llu \t %llu ", cyc_total, cyc_build, cyc_part);
    fprintf(stdout, "\n");
    fprintf(stdout, "TOTAL-TIME-USECS, TOTAL-TUPLES, NSEC-PER-TUPLE: \n");
    fprintf(stdout, "%.4lf \t %llu \t ", total_us, (unsigned long long)st.matches);
    fprintf(stdout, "%.4lf ", nS ? total_us * 1000.0 / (double)nS : 0.0);
    fprintf(stdout, "\n");
    fprintf(stdout, "PARTITION-TIME-USECS, PROBE-TIME-USECS, JOIN-TIME-USECS: \n");
    fprintf(stdout, "%.4lf \t %.4lf\t %.4lf \n", part_us, join_us, join_us);
    fprintf(stdout, "H2D-COPY-USECS, END-TO-END-USECS, GPUS: \n");
    fprintf(stdout, "%.4lf \t %.4lf\t %d \n", st.ms_h2d * 1000.0, st.ms_e2e * 1000.0, st.n_gpus);
    fflush(stdout);
}

static void h2d(void* dst, const void* src, size_t bytes) {
    if (bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g.stream));
}

static Fab single_fab() {
    Fab f;
    memset(&f, 0, sizeof(f));
    f.world = 1;
    return f;
}

}  // namespace hwbrj

// ---- a group of GPUs that join together ----------------------------------------------------------------
// Every rank allocates one "symmetric" block (same layout on every rank) that its peers map: barrier flags, the gathered
// histogram and result rows, the replicated filter (+ a partial filter for non-sliceable filters) and the two receive
// buffers of the level-1 routing. Ranks of other processes are mapped through CUDA IPC, ranks of this process through
// peer access.
struct hwbrj_dist {
    int rank = 0, world = 1, dev = 0;
    unsigned char* base = nullptr;
    size_t bytes = 0;
    size_t off_epoch = 0, off_histR = 0, off_histS = 0, off_rows = 0, off_filter = 0, off_partial = 0, off_recvR = 0, off_recvS = 0;
    uint64_t cap_r = 0, cap_s = 0, filter_bytes = 0;
    unsigned char* peer[hwbrj::kMaxPeers] = {nullptr};
    bool ipc_opened[hwbrj::kMaxPeers] = {false};
    bool connected = false;
    hwbrj::Fab fab;
    void* d_out8 = nullptr;  // result words of an enqueued join
};

namespace hwbrj {

struct DistHandle {  // what travels between the ranks (HWBRJ_DIST_HANDLE_BYTES)
    cudaIpcMemHandle_t ipc;
    int32_t pid, dev;
    uint64_t ptr, bytes, cap_r, cap_s, filter_bytes;
    unsigned char pad[HWBRJ_DIST_HANDLE_BYTES - 64 - 8 - 40];
};
static_assert(sizeof(DistHandle) == HWBRJ_DIST_HANDLE_BYTES, "handle size");

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static hwbrj_dist* dist_create(int rank, int world, uint64_t cap_r, uint64_t cap_s, uint64_t filter_bytes, void* handle_out) {
    if (world < 1 || world > kMaxPeers || (world & (world - 1)) || rank < 0 || rank >= world) return nullptr;
    hwbrj_dist* d = new hwbrj_dist;
    d->rank = rank;
    d->world = world;
    d->dev = g.dev;
    d->cap_r = cap_r;
    d->cap_s = cap_s;
    d->filter_bytes = align_up(std::max<uint64_t>(filter_bytes, 16), 256);
    const size_t P = (size_t)1 << kMaxRadixBits;
    size_t off = kMaxPeers * kFlagStride * 4;  // flags first
    d->off_epoch = off; off += 256;
    d->off_histR = off; off += (size_t)world * P * 4;
    d->off_histS = off; off += (size_t)world * P * 4;
    d->off_rows = off; off += align_up((size_t)world * 64, 256);
    d->off_filter = off; off += d->filter_bytes;
    d->off_partial = off; off += d->filter_bytes;
    d->off_recvR = off; off += align_up(cap_r * 8 + 64, 256);
    d->off_recvS = off; off += align_up(cap_s * 8 + 64, 256);
    d->bytes = off;
    cudaError_t e = cudaMalloc(&d->base, d->bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        delete d;
        return nullptr;
    }
    CK(cudaMemset(d->base, 0, d->off_filter));  // flags, epoch, rows
    CK(cudaMalloc(&d->d_out8, 64));
    DistHandle h;
    memset(&h, 0, sizeof(h));
    if (cudaIpcGetMemHandle(&h.ipc, d->base) != cudaSuccess) cudaGetLastError();  // same-process peers do not need it
    h.pid = (int32_t)getpid();
    h.dev = d->dev;
    h.ptr = (uint64_t)(uintptr_t)d->base;
    h.bytes = d->bytes;
    h.cap_r = cap_r;
    h.cap_s = cap_s;
    h.filter_bytes = d->filter_bytes;
    if (handle_out) memcpy(handle_out, &h, sizeof(h));
    return d;
}

static int dist_connect(hwbrj_dist* d, const void* all_handles) {
    const DistHandle* hs = reinterpret_cast<const DistHandle*>(all_handles);
    for (int r = 0; r < d->world; r++) {
        const DistHandle& h = hs[r];
        if (h.bytes != d->bytes || h.cap_r != d->cap_r || h.cap_s != d->cap_s || h.filter_bytes != d->filter_bytes) return -3;
        if (r == d->rank) {
            d->peer[r] = d->base;
        } else if (h.pid == (int32_t)getpid()) {  // a device of this process: peer access
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, d->dev, h.dev));
            if (!can) return -4;
            cudaError_t e = cudaDeviceEnablePeerAccess(h.dev, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return -4;
            cudaGetLastError();
            d->peer[r] = reinterpret_cast<unsigned char*>((uintptr_t)h.ptr);
        } else {
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                return -5;
            }
            d->peer[r] = reinterpret_cast<unsigned char*>(p);
            d->ipc_opened[r] = true;
        }
    }
    Fab& f = d->fab;
    memset(&f, 0, sizeof(f));
    f.world = d->world;
    f.rank = d->rank;
    f.gbits = ilog2_u64((uint64_t)d->world);
    for (int r = 0; r < d->world; r++) {
        f.flags.p[r] = d->peer[r];
        f.histR.p[r] = d->peer[r] + d->off_histR;
        f.histS.p[r] = d->peer[r] + d->off_histS;
        f.rows.p[r] = d->peer[r] + d->off_rows;
        f.filter.p[r] = d->peer[r] + d->off_filter;
        f.partial.p[r] = d->peer[r] + d->off_partial;
        f.stageR.buf[r] = reinterpret_cast<uint2*>(d->peer[r] + d->off_recvR);
        f.stageS.buf[r] = reinterpret_cast<uint2*>(d->peer[r] + d->off_recvS);
    }
    f.epoch = reinterpret_cast<uint32_t*>(d->base + d->off_epoch);
    f.cap_r = d->cap_r;
    f.cap_s = d->cap_s;
    f.filter_bytes = d->filter_bytes;
    d->connected = true;
    return 0;
}

static void dist_destroy(hwbrj_dist* d) {
    if (!d) return;
    invalidate_last();
    for (int r = 0; r < d->world; r++)
        if (d->ipc_opened[r]) cudaIpcCloseMemHandle(d->peer[r]);
    if (d->base) cudaFree(d->base);
    if (d->d_out8) cudaFree(d->d_out8);
    cudaGetLastError();
    delete d;
}

// ---- host-buffer entry: copy in, join, fill result_t (join_init_run, :1561-1778) ---------------------------------------
static std::vector<hwbrj_dist*> g_group;  // the GPUs of this process that join together (hwbrj_set_gpus)

static void group_release() {
    for (size_t i = 0; i < g_group.size(); i++) {
        CK(cudaSetDevice((int)i));
        init_ctx();
        CK(cudaDeviceSynchronize());
        dist_destroy(g_group[i]);
    }
    g_group.clear();
}

// (re)creates the in-process GPU group when it is missing or too small
static void group_ensure(int n, uint64_t cap_r, uint64_t cap_s, uint64_t filter_bytes) {
    if ((int)g_group.size() == n && g_group[0]->cap_r >= cap_r && g_group[0]->cap_s >= cap_s &&
        g_group[0]->filter_bytes >= filter_bytes)
        return;
    int home = 0;
    CK(cudaGetDevice(&home));
    group_release();
    std::vector<DistHandle> hs(n);
    for (int i = 0; i < n; i++) {
        CK(cudaSetDevice(i));
        init_ctx();
        hwbrj_dist* d = dist_create(i, n, cap_r, cap_s, filter_bytes, &hs[i]);
        if (!d) die("cannot allocate the receive buffers of GPU %d (%llu + %llu tuples)", i, (unsigned long long)cap_r,
                    (unsigned long long)cap_s);
        g_group.push_back(d);
    }
    for (int i = 0; i < n; i++) {
        CK(cudaSetDevice(i));
        init_ctx();
        int rc = dist_connect(g_group[i], hs.data());
        if (rc) die("GPU %d cannot map its peers (rc %d): the GPUs of a join need peer access (NVLink)", i, rc);
    }
    CK(cudaSetDevice(home));
    init_ctx();
}

// contiguous chunk of rank i (the reference's per-thread chunks, :1646-1672: the last one takes the remainder); chunk
// starts are kept even so that every chunk stays 16-byte aligned
static void chunk_of(uint64_t n, int world, int i, uint64_t& begin, uint64_t& count) {
    const uint64_t per = (n / (uint64_t)world) & ~1ull;
    begin = per * (uint64_t)i;
    count = i == world - 1 ? n - begin : per;
}

static result_t* host_join_multi(relation_t* relR, relation_t* relS, int nthreads, bloom_filter_args_t* args,
                                 bool print_filtered) {
    auto t0 = std::chrono::steady_clock::now();
    const int G = g_gpus;
    const uint64_t nR = relR->num_tuples, nS = relS->num_tuples;
    int home = 0;
    CK(cudaGetDevice(&home));
    // receive capacities: R keys are spread by a hash of the key (25 % slack), S worst case = everything survives and
    // lands on one owner (a hot key)
    group_ensure(G, nR / G + nR / (4 * G) + (1u << 16), nS + 2, args ? args->m / 8 : 16);
    std::vector<hwbrj_stats_t> sts(G);
    for (int i = 0; i < G; i++) {  // first all copies (a pageable source makes them block the host) ...
        CK(cudaSetDevice(i));
        init_ctx();
        memset(&sts[i], 0, sizeof(hwbrj_stats_t));
        uint64_t rb, rc, sb, sc_;
        chunk_of(nR, G, i, rb, rc);
        chunk_of(nS, G, i, sb, sc_);
        g.inR.ensure(std::max<uint64_t>(rc, 2) * 8 + 64);
        g.inS.ensure(std::max<uint64_t>(sc_, 2) * 8 + 64);
        CK(cudaEventRecord(g.ev[9], g.stream));
        h2d(g.inR.p, relR->tuples + rb, rc * 8);
        h2d(g.inS.p, relS->tuples + sb, sc_ * 8);
        CK(cudaEventRecord(g.ev[10], g.stream));
        sts[i].h2d_bytes = (rc + sc_) * 8;
    }
    for (int i = 0; i < G; i++) {  // ... then every GPU's join is enqueued; the barrier kernels meet on the devices
        CK(cudaSetDevice(i));
        init_ctx();
        uint64_t rb, rc, sb, sc_;
        chunk_of(nR, G, i, rb, rc);
        chunk_of(nS, G, i, sb, sc_);
        run_join(g_group[i]->fab, g.inR.as<uint2>(), rc, g.inS.as<uint2>(), sc_, args, nR, sts[i], nullptr, kEnqueue);
    }
    int rc_all = 0;
    hwbrj_stats_t st = sts[0];
    for (int i = 0; i < G; i++) {
        CK(cudaSetDevice(i));
        init_ctx();
        rc_all |= collect_join(sts[i], args != nullptr);
        CK(cudaEventElapsedTime(&sts[i].ms_h2d, g.ev[9], g.ev[10]));
        if (i == 0) st = sts[0];
        st.ms_total = std::max(st.ms_total, sts[i].ms_total);  // device time, max over the GPUs
        st.ms_h2d = std::max(st.ms_h2d, sts[i].ms_h2d);
        if (i) {
            st.h2d_bytes += sts[i].h2d_bytes;
            st.d2h_bytes += sts[i].d2h_bytes;
            st.kernel_launches += sts[i].kernel_launches;
            st.owned_r = std::max(st.owned_r, sts[i].owned_r);  // the most loaded GPU
            st.owned_s = std::max(st.owned_s, sts[i].owned_s);
        }
    }
    CK(cudaSetDevice(home));
    init_ctx();
    if (rc_all) die("multi-GPU join failed: a receive buffer overflowed or a GPU did not reach a barrier");
    st.ms_e2e = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    g.last = st;
    print_reference_lines(st, nS, args != nullptr && print_filtered);
    result_t* res = (result_t*)malloc(sizeof(result_t));
    if (!res) die("malloc(result_t) failed");
    res->totalresults = st.matches;
    res->resultlist = nullptr;
    res->nthreads = nthreads;
    return res;
}

static result_t* host_join(relation_t* relR, relation_t* relS, int nthreads, bloom_filter_args_t* args,
                           bool print_filtered) {
    std::lock_guard<std::recursive_mutex> lock(g_mu);
    init_ctx();
    if (!relR || !relS) die("NULL relation");
    if (args && check_args_impl(args, true)) die("invalid Bloom filter arguments");
    if (g_gpus > 1) return host_join_multi(relR, relS, nthreads, args, print_filtered);
    auto t0 = std::chrono::steady_clock::now();
    hwbrj_stats_t st;
    memset(&st, 0, sizeof(st));
    const uint64_t nR = relR->num_tuples, nS = relS->num_tuples;
    g.inR.ensure(std::max<uint64_t>(nR, 2) * 8 + 64);
    g.inS.ensure(std::max<uint64_t>(nS, 2) * 8 + 64);
    st.h2d_bytes = (nR + nS) * 8;
    const Fab f = single_fab();
    if (g.overlap_h2d && args && nS >= (1u << 22)) {
        // copies on their own stream; R first, then S in <= 64 chunks, each followed by an event the probe waits on
        SFeed feed;
        feed.nchunks = (int)std::min<uint64_t>(64, (nS + (1u << 22) - 1) >> 22);
        feed.chunk_tuples = (((nS + feed.nchunks - 1) / feed.nchunks) + 1) & ~1ull;  // even: chunks stay 16-byte aligned
        feed.nchunks = (int)((nS + feed.chunk_tuples - 1) / feed.chunk_tuples);
        feed.ev = g.ev_chunk + 1;
        CK(cudaEventRecord(g.ev_copy[0], g.copy_stream));
        if (nR) CK(cudaMemcpyAsync(g.inR.p, relR->tuples, nR * 8, cudaMemcpyHostToDevice, g.copy_stream));
        CK(cudaEventRecord(g.ev_chunk[0], g.copy_stream));
        for (int c = 0; c < feed.nchunks; c++) {
            const uint64_t off = (uint64_t)c * feed.chunk_tuples;
            const uint64_t cnt = std::min<uint64_t>(feed.chunk_tuples, nS - off);
            CK(cudaMemcpyAsync(g.inS.as<uint2>() + off, relS->tuples + off, cnt * 8, cudaMemcpyHostToDevice, g.copy_stream));
            CK(cudaEventRecord(feed.ev[c], g.copy_stream));
        }
        CK(cudaEventRecord(g.ev_copy[1], g.copy_stream));
        CK(cudaStreamWaitEvent(g.stream, g.ev_chunk[0], 0));  // the R phase needs all of R
        run_join(f, g.inR.as<uint2>(), nR, g.inS.as<uint2>(), nS, args, nR, st, &feed);
        CK(cudaEventSynchronize(g.ev_copy[1]));  // recorded on the copy stream: make sure it has completed before reading it
        CK(cudaEventElapsedTime(&st.ms_h2d, g.ev_copy[0], g.ev_copy[1]));
    } else {
        CK(cudaEventRecord(g.ev[9], g.stream));
        h2d(g.inR.p, relR->tuples, nR * 8);
        h2d(g.inS.p, relS->tuples, nS * 8);
        CK(cudaEventRecord(g.ev[10], g.stream));
        CK(cudaStreamSynchronize(g.stream));
        CK(cudaEventElapsedTime(&st.ms_h2d, g.ev[9], g.ev[10]));
        run_join(f, g.inR.as<uint2>(), nR, g.inS.as<uint2>(), nS, args, nR, st);
    }
    st.ms_e2e = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    g.last = st;
    print_reference_lines(st, nS, args != nullptr && print_filtered);
    result_t* res = (result_t*)malloc(sizeof(result_t));
    if (!res) die("malloc(result_t) failed");
    res->totalresults = st.matches;
    res->resultlist = nullptr;  // only allocated under JOIN_RESULT_MATERIALIZE in the reference (:1598-1601)
    res->nthreads = nthreads;
    return res;
}

}  // namespace hwbrj

using namespace hwbrj;

extern "C" {

// ---- Part 1: the reference's entry points ---------------------------------------------------------------------
result_t* BPRO(relation_t* relR, relation_t* relS, int nthreads, bloom_filter_args_t* args) {
    if (!args) die("BPRO: NULL bloom_filter_args");
    return host_join(relR, relS, nthreads, args, true);
}
result_t* BPRH(relation_t* relR, relation_t* relS, int nthreads, bloom_filter_args_t* args) {
    return BPRO(relR, relS, nthreads, args);
}
result_t* BPRHO(relation_t* relR, relation_t* relS, int nthreads, bloom_filter_args_t* args) {
    return BPRO(relR, relS, nthreads, args);
}
result_t* BRJ(relation_t* relR, relation_t* relS, int nthreads, bloom_filter_args_t* args) {
    if (!args) die("BRJ: NULL bloom_filter_args");
    result_t* r = host_join(relR, relS, nthreads, args, false);  // BRJ does not print the filtered count
    r->nthreads = 1;                                             // :1974
    return r;
}
result_t* PRO(relation_t* relR, relation_t* relS, int nthreads) { return host_join(relR, relS, nthreads, nullptr, false); }
result_t* PRH(relation_t* relR, relation_t* relS, int nthreads) { return PRO(relR, relS, nthreads); }
result_t* PRHO(relation_t* relR, relation_t* relS, int nthreads) { return PRO(relR, relS, nthreads); }
result_t* RJ(relation_t* relR, relation_t* relS, int nthreads) {
    result_t* r = host_join(relR, relS, nthreads, nullptr, false);
    r->nthreads = 1;
    return r;
}

// ---- Part 2: extensions ---------------------------------------------------------------------------------------
#define HWBRJ_ENTER() std::lock_guard<std::recursive_mutex> lock_(g_mu); init_ctx()

int hwbrj_last_stats(hwbrj_stats_t* out) {
    HWBRJ_ENTER();
    if (!out) return -1;
    *out = g.last;
    return 0;
}
int64_t hwbrj_last_filtered(void) {
    HWBRJ_ENTER();
    return g.last.filtered;
}
int hwbrj_last_filter(unsigned char* bitmap_out, uint64_t nbytes) {
    HWBRJ_ENTER();
    const void* src = g_group.empty() ? g.filter.p : (const void*)(g_group[g.dev < (int)g_group.size() ? g.dev : 0]->base +
                                                                   g_group[0]->off_filter);
    const uint64_t cap = g_group.empty() ? g.filter.cap : g_group[0]->filter_bytes;
    if (!bitmap_out || nbytes > cap || !src) return -1;
    CK(cudaMemcpy(bitmap_out, src, nbytes, cudaMemcpyDeviceToHost));
    return 0;
}
uint64_t hwbrj_last_checksum(void) {
    HWBRJ_ENTER();
    return g.last.checksum_pair;
}
void hwbrj_set_quiet(int quiet) { g_quiet = quiet != 0; }
void hwbrj_set_radix_bits(int bits) {
    HWBRJ_ENTER();
    g.radix_bits_override = bits;
}
void hwbrj_set_num_passes(int passes) {
    HWBRJ_ENTER();
    g.passes_override = passes;
}
void hwbrj_set_range_passes(int passes) {
    HWBRJ_ENTER();
    g.range_passes_override = passes;
}
void hwbrj_set_overlap_h2d(int on) {
    HWBRJ_ENTER();
    g.overlap_h2d = on != 0;
}
void hwbrj_set_hash_partition(int mode) {
    HWBRJ_ENTER();
    g.hash_partition = std::max(0, std::min(2, mode));
}
int hwbrj_set_gpus(int n) {
    std::lock_guard<std::recursive_mutex> lock(g_mu);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    if (n < 1 || n > ndev || n > kMaxPeers || (n & (n - 1))) return -1;
    if (n != g_gpus && !g_group.empty()) group_release();
    g_gpus = n;
    return 0;
}
const char* hwbrj_version(void) { return "hwbrj-b200 0.2 (sm_100a)"; }
int hwbrj_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
int hwbrj_check_args(const bloom_filter_args_t* args) { return args ? check_args_impl(args, true) : 1; }

struct hwbrj_rel {
    uint2* d;
    uint64_t n;
    bool owned;
};

hwbrj_rel_t* hwbrj_rel_upload(const tuple_t* tuples, uint64_t n) {
    HWBRJ_ENTER();
    hwbrj_rel_t* r = new hwbrj_rel;
    r->n = n;
    r->owned = true;
    CK(cudaMalloc(&r->d, std::max<uint64_t>(n, 2) * 8 + 64));
    if (n) CK(cudaMemcpy(r->d, tuples, n * 8, cudaMemcpyHostToDevice));
    return r;
}

hwbrj_rel_t* hwbrj_rel_generate(int kind, uint64_t n, uint64_t r, double q, uint64_t seed) {
    return hwbrj_rel_generate_shard(kind, n, r, q, seed, 0, n);
}

hwbrj_rel_t* hwbrj_rel_generate_shard(int kind, uint64_t n, uint64_t r, double q, uint64_t seed, uint64_t begin,
                                      uint64_t count) {
    HWBRJ_ENTER();
    if (begin > n) begin = n;
    if (count > n - begin) count = n - begin;
    hwbrj_rel_t* rel = new hwbrj_rel;
    rel->n = count;
    rel->owned = true;
    CK(cudaMalloc(&rel->d, std::max<uint64_t>(count, 2) * 8 + 64));
    if (count && kind == 2) {
        // Zipf foreign keys over the alphabet 1..r with exponent q (create_relation_zipf, generator.c:659-676)
        const uint64_t alpha = std::min<uint64_t>(r ? r : 1, 0xFFFFFFFFull);
        const uint32_t nchunks = (uint32_t)((alpha + kZipfChunk - 1) / kZipfChunk);
        if (g.zipf_r != alpha || g.zipf_theta != q) {  // the cumulated-density table is kept for the next shard / call
            g.zipf_lut.ensure(alpha * sizeof(double));
            g.zipf_sums.ensure(((size_t)nchunks + 1) * sizeof(double));
            k_zipf_scan_chunks<<<nchunks, 256, 0, g.stream>>>(g.zipf_lut.as<double>(), alpha, q, g.zipf_sums.as<double>());
            k_zipf_scan_sums<<<1, 1024, 0, g.stream>>>(g.zipf_sums.as<double>(), nchunks);
            k_zipf_finish<<<nchunks, 256, 0, g.stream>>>(g.zipf_lut.as<double>(), alpha, g.zipf_sums.as<double>(), nchunks);
            g.zipf_r = alpha;
            g.zipf_theta = q;
        }
        int bitsr = 1;
        while ((1ull << bitsr) < alpha) bitsr++;
        k_generate_zipf<<<g.sms * 8, 256, 0, g.stream>>>(rel->d, begin, count, g.zipf_lut.as<double>(), (uint32_t)alpha,
                                                         (uint32_t)((bitsr + 1) / 2), seed * 0x9e3779b97f4a7c15ULL + 12345);
        CK(cudaStreamSynchronize(g.stream));
        CK(cudaGetLastError());
    } else if (count) {
        // generator.c:344: ntuples_above = num_tuples * (1 - selectivity)
        uint64_t na = kind == 1 ? (uint64_t)((double)n * (1.0 - q)) : 0;
        uint64_t nb = n - na;
        int bitsn = 1;
        while ((1ull << bitsn) < n) bitsn++;
        uint32_t half = (uint32_t)((bitsn + 1) / 2);
        k_generate<<<g.sms * 8, 256, 0, g.stream>>>(rel->d, n, kind, r ? r : 1, nb, half, seed * 0x9e3779b97f4a7c15ULL + 12345,
                                                    begin, count);
        CK(cudaStreamSynchronize(g.stream));
        CK(cudaGetLastError());
    }
    return rel;
}

int hwbrj_rel_download(const hwbrj_rel_t* rel, tuple_t* out) {
    if (!rel || !out) return -1;
    if (rel->n) CK(cudaMemcpy(out, rel->d, rel->n * 8, cudaMemcpyDeviceToHost));
    return 0;
}
uint64_t hwbrj_rel_size(const hwbrj_rel_t* rel) { return rel ? rel->n : 0; }
void hwbrj_rel_free(hwbrj_rel_t* rel) {
    if (!rel) return;
    if (rel->owned) cudaFree(rel->d);
    delete rel;
}
hwbrj_rel_t* hwbrj_rel_wrap(void* device_tuples, uint64_t n) {
    if (reinterpret_cast<uintptr_t>(device_tuples) & 15) return nullptr;  // the kernels stream 128-bit / TMA bulk loads
    hwbrj_rel_t* r = new hwbrj_rel;
    r->d = reinterpret_cast<uint2*>(device_tuples);
    r->n = n;
    r->owned = false;
    return r;
}
void* hwbrj_rel_ptr(const hwbrj_rel_t* rel) { return rel ? rel->d : nullptr; }
void hwbrj_set_stream(void* cuda_stream) {
    HWBRJ_ENTER();
    g.stream = reinterpret_cast<cudaStream_t>(cuda_stream);  // 0 is a valid handle: the legacy default stream
}
void hwbrj_reset_stream(void) {
    HWBRJ_ENTER();
    g.stream = g.own_stream;
}
int hwbrj_set_device(int device) {
    // one process per GPU: call before the first library call; the current CUDA device selects the library's context
    if (cudaSetDevice(device) != cudaSuccess) {
        cudaGetLastError();
        return -2;
    }
    return 0;
}
int hwbrj_sync(void) {
    HWBRJ_ENTER();
    CK(cudaStreamSynchronize(g.stream));
    return 0;
}

int hwbrj_join_device(const hwbrj_rel_t* R, const hwbrj_rel_t* S, const bloom_filter_args_t* args, hwbrj_stats_t* out) {
    HWBRJ_ENTER();
    if (!R || !S) return -1;
    hwbrj_stats_t st;
    memset(&st, 0, sizeof(st));
    const int rc = run_join(single_fab(), R->d, R->n, S->d, S->n, args, R->n, st);
    g.last = st;
    if (out) *out = st;
    return rc;
}

int hwbrj_join_device_async(const hwbrj_rel_t* R, const hwbrj_rel_t* S, const bloom_filter_args_t* args, void* d_out8) {
    HWBRJ_ENTER();
    if (!R || !S || !d_out8) return -1;
    hwbrj_stats_t st;
    memset(&st, 0, sizeof(st));
    run_join(single_fab(), R->d, R->n, S->d, S->n, args, R->n, st, nullptr, kCapture,
             reinterpret_cast<unsigned long long*>(d_out8));
    return st.kernel_launches;
}

// ---- the multi-GPU join below the C ABI ---------------------------------------------------------------
hwbrj_dist_t* hwbrj_dist_create(int rank, int world, uint64_t cap_r, uint64_t cap_s, uint64_t max_filter_bytes,
                                void* handle_out) {
    HWBRJ_ENTER();
    return dist_create(rank, world, cap_r, cap_s, max_filter_bytes, handle_out);
}
int hwbrj_dist_connect(hwbrj_dist_t* d, const void* all_handles) {
    HWBRJ_ENTER();
    if (!d || !all_handles) return -1;
    return dist_connect(d, all_handles);
}
void hwbrj_dist_destroy(hwbrj_dist_t* d) {
    HWBRJ_ENTER();
    if (d) CK(cudaDeviceSynchronize());
    dist_destroy(d);
}
int hwbrj_dist_join(hwbrj_dist_t* d, const hwbrj_rel_t* R, const hwbrj_rel_t* S, const bloom_filter_args_t* args,
                    uint64_t r_total, hwbrj_stats_t* out) {
    HWBRJ_ENTER();
    if (!d || !d->connected || !R || !S) return -1;
    hwbrj_stats_t st;
    memset(&st, 0, sizeof(st));
    const int rc = run_join(d->fab, R->d, R->n, S->d, S->n, args, r_total, st);
    g.last = st;
    if (out) *out = st;
    return rc;
}
int hwbrj_dist_join_async(hwbrj_dist_t* d, const hwbrj_rel_t* R, const hwbrj_rel_t* S, const bloom_filter_args_t* args,
                          uint64_t r_total, void* d_out8) {
    HWBRJ_ENTER();
    if (!d || !d->connected || !R || !S || !d_out8) return -1;
    hwbrj_stats_t st;
    memset(&st, 0, sizeof(st));
    run_join(d->fab, R->d, R->n, S->d, S->n, args, r_total, st, nullptr, kCapture, reinterpret_cast<unsigned long long*>(d_out8));
    return st.kernel_launches;
}
void* hwbrj_dist_filter(hwbrj_dist_t* d) { return d ? d->base + d->off_filter : nullptr; }

void* hwbrj_host_alloc(uint64_t bytes) {
    HWBRJ_ENTER();
    void* p = nullptr;
    CK(cudaHostAlloc(&p, std::max<uint64_t>(bytes, 8), cudaHostAllocPortable));
    return p;
}
void hwbrj_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int hwbrj_hash_many(int which, uint32_t seed, const int32_t* keys, uint64_t n, uint32_t* out) {
    HWBRJ_ENTER();
    if (which < 0 || which > 9) return -1;
    if (!n) return 0;
    g.scratch.ensure(n * 8);
    int32_t* dk = g.scratch.as<int32_t>();
    uint32_t* dout = reinterpret_cast<uint32_t*>(dk + n);
    CK(cudaMemcpyAsync(dk, keys, n * 4, cudaMemcpyHostToDevice, g.stream));
    k_hash_many<<<g.sms * 4, 256, 0, g.stream>>>(which, seed, dk, n, dout);
    CK(cudaMemcpyAsync(out, dout, n * 4, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaGetLastError());
    return 0;
}

// insert R's keys into the filter at d_filter with global atomics (K1 without a histogram of its own: one bin)
static void enqueue_filter_build(const uint2* dR, uint64_t n, const bloom_filter_args_t* args, uint32_t seed, void* d_filter) {
    g.histR.ensure(((size_t)1 << kMaxRadixBits) * 4);
    CK(cudaMemsetAsync(g.histR.p, 0, 4, g.stream));
    BloomParams bp = make_bloom(args, seed, reinterpret_cast<uint32_t*>(d_filter));
    const int nranges = pick_ranges(args);
    set_ranges(bp, args, nranges);
    PartFn pf;
    memset(&pf, 0, sizeof(pf));
    for (int r = 0; r < nranges; r++) {
        bp.range_id = (uint32_t)r;
        k_build_hist<true, 0><<<g.sms * 2, 1024, (1 + kCrcSmemWords) * 4, g.stream>>>(dR, n, nullptr, bp, g.d_crc,
                                                                                  g.histR.as<uint32_t>(), pf);
    }
}
// probe S against the filter at d_filter; survivors to d_out, their number to *d_count (zeroed here)
static void enqueue_filter_probe(const uint2* dS, uint64_t n, const bloom_filter_args_t* args, uint32_t seed,
                                 const void* d_filter, uint2* d_out, unsigned long long* d_count) {
    CK(cudaMemsetAsync(d_count, 0, 8, g.stream));
    BloomParams bp = make_bloom(args, seed, reinterpret_cast<uint32_t*>(const_cast<void*>(d_filter)));
    const int nranges = pick_ranges(args);
    set_ranges(bp, args, nranges);
    run_probe(dS, n, nullptr, bp, nranges, d_out, d_count);
}

int hwbrj_bloom_build(const tuple_t* R, uint64_t nR, const bloom_filter_args_t* args, uint32_t seed,
                      unsigned char* bitmap_out) {
    HWBRJ_ENTER();
    if (!args || check_args_impl(args, true)) return -1;
    invalidate_last();
    g.inR.ensure(std::max<uint64_t>(nR, 2) * 8 + 64);
    g.filter.ensure(std::max<uint64_t>(args->m / 8, 16));
    h2d(g.inR.p, R, nR * 8);
    CK(cudaMemsetAsync(g.filter.p, 0, std::max<uint64_t>(args->m / 8, 16), g.stream));
    enqueue_filter_build(g.inR.as<uint2>(), nR, args, seed, g.filter.p);
    CK(cudaMemcpyAsync(bitmap_out, g.filter.p, args->m / 8, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaGetLastError());
    return 0;
}

int64_t hwbrj_bloom_probe(const unsigned char* bitmap, const tuple_t* S, uint64_t nS, const bloom_filter_args_t* args,
                          uint32_t seed, tuple_t* survivors_out) {
    HWBRJ_ENTER();
    if (!args || check_args_impl(args, true)) return -1;
    invalidate_last();
    g.inS.ensure(std::max<uint64_t>(nS, 2) * 8 + 64);
    g.sc.ensure(std::max<uint64_t>(nS, 1) * 8 + 64);
    g.filter.ensure(std::max<uint64_t>(args->m / 8, 16));
    g.ctrl.ensure(sizeof(Control));
    h2d(g.inS.p, S, nS * 8);
    h2d(g.filter.p, bitmap, args->m / 8);
    Control* ctrl = g.ctrl.as<Control>();
    enqueue_filter_probe(g.inS.as<uint2>(), nS, args, seed, g.filter.p, g.sc.as<uint2>(), &ctrl->survivors);
    unsigned long long cnt = 0;
    CK(cudaMemcpyAsync(&cnt, &ctrl->survivors, 8, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    if (survivors_out && cnt) CK(cudaMemcpy(survivors_out, g.sc.p, cnt * 8, cudaMemcpyDeviceToHost));
    CK(cudaGetLastError());
    return (int64_t)cnt;
}

// Materialise the output of the most recent join: re-runs only the per-partition build+probe (K5) over the
// partitions that join left in the workspace, writing one {R.payload, S.payload} tuple per match
// (bucket_chaining_join under JOIN_RESULT_MATERIALIZE, :307-312). Returns the number of pairs (which may exceed
// `capacity`: then only the first `capacity` slots were written and the caller retries with a larger buffer), or -1 when
// there is no join to materialise: any other call that touches the workspace invalidates it.
static int64_t materialize_last(uint2* d_pairs, uint64_t capacity) {
    if (!g.last_Rp || !g.last_Sp) return -1;
    Control* ctrl = g.ctrl.as<Control>();
    CK(cudaMemsetAsync(&ctrl->acc, 0, sizeof(JoinAccum), g.stream));
    CK(cudaMemsetAsync(&ctrl->item_counter[0], 0, sizeof(uint32_t), g.stream));
    CK(cudaMemsetAsync(&ctrl->pair_cursor, 0, sizeof(unsigned long long), g.stream));
    const int smem = kJoinSmemBytes;
    if (g.last_hash)
        k_join<true, true><<<g.sms * g.occ_join, kJoinThreads, smem, g.stream>>>(
            g.last_Rp, g.offR.as<uint32_t>(), g.last_Sp, g.offS.as<uint32_t>(), g.work.as<uint32_t>(), g.work_part.as<uint32_t>(), g.last_P,
            g.last_bits, &ctrl->item_counter[0], &ctrl->acc, d_pairs, &ctrl->pair_cursor, capacity);
    else
        k_join<false, true><<<g.sms * g.occ_join, kJoinThreads, smem, g.stream>>>(
            g.last_Rp, g.offR.as<uint32_t>(), g.last_Sp, g.offS.as<uint32_t>(), g.work.as<uint32_t>(), g.work_part.as<uint32_t>(), g.last_P,
            g.last_bits, &ctrl->item_counter[0], &ctrl->acc, d_pairs, &ctrl->pair_cursor, capacity);
    unsigned long long cnt = 0;
    CK(cudaMemcpyAsync(&cnt, &ctrl->pair_cursor, 8, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaGetLastError());
    return (int64_t)cnt;
}

int64_t hwbrj_materialize_last(tuple_t* pairs_out, uint64_t capacity) {
    HWBRJ_ENTER();
    if (!pairs_out && capacity) return -1;
    if (!g.last_Rp || !g.last_Sp) return -1;
    const uint2 *keepR = g.last_Rp, *keepS = g.last_Sp;
    g.pairs.ensure(std::max<uint64_t>(capacity, 1) * 8);  // its re-allocation does not touch the partitions
    g.last_Rp = keepR;
    g.last_Sp = keepS;
    int64_t n = materialize_last(g.pairs.as<uint2>(), capacity);
    if (n > 0) CK(cudaMemcpy(pairs_out, g.pairs.p, std::min<uint64_t>((uint64_t)n, capacity) * 8, cudaMemcpyDeviceToHost));
    return n;
}

int64_t hwbrj_materialize_last_device(void* d_pairs, uint64_t capacity) {
    HWBRJ_ENTER();
    return materialize_last(reinterpret_cast<uint2*>(d_pairs), capacity);
}

// device analogue of the reference's FPR measurement (test_bloom_fpr, unit_tests.c:191-241): build a filter with the
// given seed from R, probe S, return how many S keys pass
int64_t hwbrj_fpr_count(const hwbrj_rel_t* R, const hwbrj_rel_t* S, const bloom_filter_args_t* args, uint32_t seed) {
    HWBRJ_ENTER();
    if (!R || !S || !args || check_args_impl(args, true)) return -1;
    invalidate_last();
    g.filter.ensure(std::max<uint64_t>(args->m / 8, 16));
    g.sc.ensure(std::max<uint64_t>(S->n, 1) * 8 + 64);
    g.ctrl.ensure(sizeof(Control));
    CK(cudaMemsetAsync(g.filter.p, 0, std::max<uint64_t>(args->m / 8, 16), g.stream));
    Control* ctrl = g.ctrl.as<Control>();
    enqueue_filter_build(R->d, R->n, args, seed, g.filter.p);
    enqueue_filter_probe(S->d, S->n, args, seed, g.filter.p, g.sc.as<uint2>(), &ctrl->survivors);
    unsigned long long cnt = 0;
    CK(cudaMemcpyAsync(&cnt, &ctrl->survivors, 8, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaGetLastError());
    return (int64_t)cnt;
}

// partition `in` (device, n tuples) with the pipeline's own kernels into t-buffers of the workspace; returns the result
static const uint2* partition_local(int pmode, const PartFn& pf, const uint2* in, uint64_t n, int& launches) {
    ensure_workspace(n, 1, 1, nullptr, false);
    Control* ctrl = g.ctrl.as<Control>();
    const uint32_t P = 1u << pf.bits;
    CK(cudaMemsetAsync(g.histR.p, 0, P * 4, g.stream));
    CK(cudaMemsetAsync(g.ctrl.p, 0, sizeof(Control), g.stream));
    BloomParams bp;
    memset(&bp, 0, sizeof(bp));
    launch_hist(pmode, in, n, nullptr, bp, g.histR.as<uint32_t>(), pf);
    Fab f = single_fab();
    PeerBufs stage;
    memset(&stage, 0, sizeof(stage));
    stage.buf[0] = g.rt1.as<uint2>();
    if (!partition_front(f, pmode, pf, in, n, nullptr, g.histR.as<uint32_t>(), f.histR, stage, n, g.offR.as<uint32_t>(),
                         &ctrl->n_own_r, ctrl, launches))
        return stage.buf[0];
    launch_level2(f, pmode, pf, stage, g.rp.as<uint2>(), 0u, 1u << (pf.bits - pf.b2), ctrl, launches);
    return g.rp.as<uint2>();
}

int hwbrj_radix_partition(const tuple_t* in, uint64_t n, int bits, tuple_t* out, uint64_t* offsets) {
    HWBRJ_ENTER();
    if (bits < 0 || bits > kMaxRadixBits || n >= (1ull << 32) - (1ull << 20)) return -1;
    invalidate_last();
    g.inR.ensure(std::max<uint64_t>(n, 2) * 8 + 64);
    const uint32_t P = 1u << bits;
    h2d(g.inR.p, in, n * 8);
    PartFn pf;
    memset(&pf, 0, sizeof(pf));
    pf.bits = (uint32_t)bits;
    pf.b2 = (uint32_t)pick_b2(bits, 0);
    int launches = 0;
    const uint2* res = partition_local(0, pf, g.inR.as<uint2>(), n, launches);
    std::vector<uint32_t> off32(P + 1);
    if (n) CK(cudaMemcpyAsync(out, res, n * 8, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaMemcpyAsync(off32.data(), g.offR.p, (P + 1) * 4, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaGetLastError());
    for (uint32_t i = 0; i <= P; i++) offsets[i] = off32[i];
    return 0;
}

// ---- building blocks of the NCCL reference path (hwbloomradixjoin_b200/dist.py: dist_join) --------------------------
int hwbrj_owner_partition(const hwbrj_rel_t* in, int world, const bloom_filter_args_t* slice_args, void* d_out,
                          uint64_t* counts_out) {
    HWBRJ_ENTER();
    if (!in || world < 1 || world > 128 || (world & (world - 1)) || in->n >= (1ull << 32) - (1ull << 20)) return -1;
    if (slice_args && check_args_impl(slice_args, true)) return -1;
    if (slice_args && slice_args->variant == BASIC && slice_args->k > 1) slice_args = nullptr;  // not sliceable
    if (slice_args) {
        uint64_t units = slice_args->variant == BLOCKED ? slice_args->m / slice_args->B : slice_args->m;
        if (units < (uint64_t)world) return -1;
    }
    invalidate_last();
    // owner = the rank holding the filter slice of the key's first bit / block; without a sliceable filter the top bits of
    // crapwow(42, key). Equal keys share an owner.
    const int gbits = ilog2_u64((uint64_t)world);
    PartFn pf;
    memset(&pf, 0, sizeof(pf));
    pf.bits = (uint32_t)gbits;
    pf.seed = 42u;
    int pmode = 1;
    if (slice_args && slice_args->variant == BLOCKED) {
        pmode = 2;
        const uint64_t nblocks = slice_args->m / slice_args->B;
        pf.size_mask = (uint32_t)(nblocks - 1);
        pf.hshift = (uint32_t)(ilog2_u64(nblocks) - gbits);
    } else if (slice_args) {
        pf.size_mask = (uint32_t)(slice_args->m - 1);
        pf.hshift = (uint32_t)(ilog2_u64(slice_args->m) - gbits);
    } else {
        pf.size_mask = gbits ? 0xFFFFFFFFu : 0u;  // one rank: everything is partition 0 (no 32-bit shift)
        pf.hshift = gbits ? (uint32_t)(32 - gbits) : 0u;
    }
    if (pf.hshift > 31u) {  // world == 1 with a one-unit filter
        pf.size_mask = 0u;
        pf.hshift = 0u;
    }
    int launches = 0;
    const uint2* res = partition_local(pmode, pf, in->d, in->n, launches);
    std::vector<uint32_t> off((size_t)world + 1);
    if (in->n) CK(cudaMemcpyAsync(d_out, res, in->n * 8, cudaMemcpyDeviceToDevice, g.stream));
    CK(cudaMemcpyAsync(off.data(), g.offR.p, ((size_t)world + 1) * 4, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaGetLastError());
    for (int i = 0; i < world; i++) counts_out[i] = off[i + 1] - off[i];
    return 0;
}

int hwbrj_filter_build(const hwbrj_rel_t* R, const bloom_filter_args_t* args, void* d_filter, int zero_first) {
    HWBRJ_ENTER();
    if (!R || !args || !d_filter || check_args_impl(args, true)) return -1;
    invalidate_last();
    if (zero_first) CK(cudaMemsetAsync(d_filter, 0, std::max<uint64_t>(args->m / 8, 4), g.stream));
    enqueue_filter_build(R->d, R->n, args, 42u, d_filter);
    CK(cudaGetLastError());
    return 0;
}

int hwbrj_filter_or(void* d_dst, const void* d_src, uint64_t nbytes) {
    HWBRJ_ENTER();
    if (nbytes % 16) return -1;
    k_filter_or<<<g.sms * 8, 256, 0, g.stream>>>(reinterpret_cast<uint4*>(d_dst), reinterpret_cast<const uint4*>(d_src),
                                                 nbytes / 16);
    CK(cudaGetLastError());
    return 0;
}

int64_t hwbrj_filter_probe(const void* d_filter, const hwbrj_rel_t* S, const bloom_filter_args_t* args, void* d_out) {
    HWBRJ_ENTER();
    if (!S || !args || !d_filter || !d_out || check_args_impl(args, true)) return -1;
    invalidate_last();
    g.ctrl.ensure(sizeof(Control));
    Control* ctrl = g.ctrl.as<Control>();
    enqueue_filter_probe(S->d, S->n, args, 42u, d_filter, reinterpret_cast<uint2*>(d_out), &ctrl->survivors);
    unsigned long long cnt = 0;
    CK(cudaMemcpyAsync(&cnt, &ctrl->survivors, 8, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaGetLastError());
    return (int64_t)cnt;
}

}  // extern "C"
