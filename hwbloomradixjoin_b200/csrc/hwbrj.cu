// hwbrj.cu -- host side of libhwbrj_cuda.so: workspace, the single-GPU join pipeline and the C ABI of
// include/hwbrj.h. Mirrors the reference's join_init_run()/prj_thread() orchestration
// (parallel_radix_join_bloom.c:1060-1506,1561-1778) as one CUDA stream of kernels with no host round trip
// between phases; the pthread barriers of the reference become kernel boundaries.
#include <cuda_runtime.h>
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

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
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
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// small control block living in one allocation (zeroed with one memset per join)
struct Control {
    unsigned long long survivors;  // K2 output cursor == filtered
    unsigned long long defer[7];   // sizes of the deferred inputs of range passes 1..7
    JoinAccum acc;
    uint32_t item_counter;
    uint32_t pad[3];
    unsigned long long pair_cursor;  // materialised output pairs
    unsigned long long pad2;
};

struct Ctx {
    bool inited = false;
    int dev = 0;
    int sms = 0;
    int clock_khz = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev[8];
    uint32_t* d_crc = nullptr;
    DevBuf filter, histR, histS, offR, offS, cur1, cur2, tiles, work, work_part, ctrl, rt1, rp, sc, st1, inR, inS, scratch;
    DevBuf cur1b, cur2b, tilesb;          // second set of scatter cursors: R partitioning may overlap the S probe
    cudaStream_t side_stream = nullptr;   // R partitioning runs here while K2 runs on the main stream
    cudaEvent_t ev_side[4];
    // state of the most recent join's partitions (inputs of a materialising k_join pass)
    const uint2* last_Rp = nullptr;
    const uint2* last_Sp = nullptr;
    uint32_t last_P = 0, last_bits = 0;
    bool last_hash = false;
    DevBuf pairs;
    // host-buffer calls: upload S in chunks on a copy stream and probe each chunk as soon as it has landed
    bool overlap_h2d = false;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_chunk[66];
    bool hash_partition = true;           // BASIC k<=1: partition on the filter-slice index, build the filter in smem
    bool hash_partition_force = false;    // HWBRJ_HASH_PARTITION=2: also for small filters (tests)
    bool overlap_r_partition = false;     // measured: the scatter traffic evicts the probed filter range (C1: 11.1 vs 9.9 ms)
    hwbrj_stats_t last;
    bool quiet = false;
    int radix_bits_override = 0;
    int range_passes_override = 0;
    int probe_ctas_per_sm = 0;  // 0 = occupancy API
    int probe_carveout = -1;    // K2 shared-memory carve-out in percent (-1 = driver default)
    bool probe_staged = false;  // experimental: k >= 2 probes run on compacted candidates (k_probe_staged)
    // R side of a filter-less join partitioned ahead of time on the side stream (hwbrj_join_prepare_r)
    struct {
        bool valid = false;
        const uint2* d = nullptr;
        uint64_t n = 0;
        const unsigned long long* n_dev = nullptr;
        int bits = 0;
        const uint2* Rp = nullptr;
        int launches = 0;
    } prep;
    cudaEvent_t ev_prep_fork = nullptr, ev_prep_done = nullptr;
    bool route_precount = false;  // experimental: hwbrj_route_peer claims once per owner (k_route_claim)
    DevBuf route_hist, route_cur;
    DevBuf histF;  // one-bin histogram of filter-only builds while histR belongs to a prepared R partitioning
    bool defer_ranges = false;  // range passes with deferral: measured slower on B200 (deferred writes thrash L2), kept as an option
    DevBuf d1;                  // second deferral buffer (only for more than 2 range passes)
    DevBuf zipf_lut, zipf_sums;  // cumulated Zipf density of the last (alphabet size, exponent) that was generated
    uint64_t zipf_r = 0;
    double zipf_theta = -1.0;
    int occ_scatter1 = 1, occ_scatter2 = 1, occ_join = 1;
    std::mutex mu;
};

static Ctx g;

static void init_ctx() {
    if (g.inited) return;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        die("no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
    CK(cudaGetDevice(&g.dev));
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, g.dev));
    g.sms = pr.multiProcessorCount;
    g.clock_khz = pr.clockRate;
    CK(cudaStreamCreateWithFlags(&g.own_stream, cudaStreamNonBlocking));
    g.stream = g.own_stream;
    for (auto& ev : g.ev) CK(cudaEventCreate(&ev));
    CK(cudaStreamCreateWithFlags(&g.side_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
    for (auto& ev : g.ev_chunk) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    if (const char* s = getenv("HWBRJ_OVERLAP_H2D")) g.overlap_h2d = atoi(s) != 0;
    for (auto& ev : g.ev_side) CK(cudaEventCreate(&ev));
    CK(cudaEventCreateWithFlags(&g.ev_prep_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&g.ev_prep_done, cudaEventDisableTiming));
    if (const char* s = getenv("HWBRJ_OVERLAP")) g.overlap_r_partition = atoi(s) != 0;
    CrcTables T;
    crc_tables_fill(T);
    CK(cudaMalloc(&g.d_crc, sizeof(T)));
    CK(cudaMemcpy(g.d_crc, &T, sizeof(T), cudaMemcpyHostToDevice));
    const int hist_smem = ((1 << kMaxRadixBits) + kCrcSmemWords) * 4;
    CK(cudaFuncSetAttribute(k_build_hist<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, hist_smem));
    CK(cudaFuncSetAttribute(k_build_hist<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, hist_smem));
    CK(cudaFuncSetAttribute(k_join<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kJoinSmemBytes));
    CK(cudaFuncSetAttribute(k_join<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kJoinSmemBytes));
    CK(cudaFuncSetAttribute(k_join<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kJoinSmemBytes));
    CK(cudaFuncSetAttribute(k_join<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kJoinSmemBytes));
    if (const char* s = getenv("HWBRJ_RADIX_BITS")) g.radix_bits_override = atoi(s);
    if (const char* s = getenv("HWBRJ_RANGE_PASSES")) g.range_passes_override = atoi(s);
    if (const char* s = getenv("HWBRJ_QUIET")) g.quiet = atoi(s) != 0;
    if (const char* s = getenv("HWBRJ_PROBE_CTAS")) g.probe_ctas_per_sm = std::max(0, atoi(s));
    if (const char* s = getenv("HWBRJ_PROBE_CARVEOUT")) g.probe_carveout = std::min(100, atoi(s));
    if (const char* s = getenv("HWBRJ_PROBE_STAGED")) g.probe_staged = atoi(s) != 0;
    if (const char* s = getenv("HWBRJ_ROUTE_PRECOUNT")) g.route_precount = atoi(s) != 0;
    if (const char* s = getenv("HWBRJ_DEFER")) g.defer_ranges = atoi(s) != 0;
    CK(cudaFuncSetAttribute(k_scatter<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScatterSmem));
    CK(cudaFuncSetAttribute(k_scatter<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScatterSmem));
    CK(cudaFuncSetAttribute(k_scatter<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScatterSmem));
    CK(cudaFuncSetAttribute(k_scatter<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScatterSmem));
    CK(cudaFuncSetAttribute(k_scatter<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScatterSmem));
    CK(cudaFuncSetAttribute(k_filter_from_parts, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024));
    CK(cudaFuncSetAttribute(k_build_hist<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, hist_smem));
    if (const char* s = getenv("HWBRJ_HASH_PARTITION")) {
        g.hash_partition = atoi(s) != 0;
        g.hash_partition_force = atoi(s) == 2;
    }
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.occ_scatter1, k_scatter<1>, kScatterThreads, kScatterSmem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.occ_scatter2, k_scatter<2>, kScatterThreads, kScatterSmem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.occ_join, k_join<false>, kJoinThreads, kJoinSmemBytes));
    g.occ_scatter1 = std::max(g.occ_scatter1, 1);
    g.occ_scatter2 = std::max(g.occ_scatter2, 1);
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
    return bp;
}

// number of filter range passes: keep the actively probed part of the filter L2-resident (<= 64 MiB).
// Only valid when all k bits of a key fall into one range: k <= 1 or BLOCKED.
static int pick_ranges(const bloom_filter_args_t* a) {
    if (a->variant == BASIC && a->k > 1) return 1;
    uint64_t bytes = a->m / 8;
    int nr = 1;
    if (g.range_passes_override > 0) nr = g.range_passes_override;
    else
        while ((bytes / nr) > (64ull << 20)) nr <<= 1;
    // must be a power of two and leave ranges >= one block / one word
    while (nr & (nr - 1)) nr &= nr - 1;
    uint64_t min_range_bits = a->variant == BLOCKED ? std::max<uint64_t>(a->B, 32) : 32;
    while (nr > 1 && a->m / nr < min_range_bits) nr >>= 1;
    return std::max(nr, 1);
}

static int pick_bits(uint64_t nR) {
    if (g.radix_bits_override > 0) return std::min(g.radix_bits_override, (int)kMaxRadixBits);
    int b = 0;
    while (b < kMaxRadixBits && (nR >> b) > (uint64_t)(kTableCap * 3 / 4)) b++;
    return b;
}


// K2 launch with compile-time specialisation on (blocked, k == 1, ranged, defer)
static void launch_probe_mode(int mode, const uint2* in, uint64_t n, const unsigned long long* n_ptr, const BloomParams& bp,
                              uint2* out, unsigned long long* cursor, uint2* defer_out, unsigned long long* defer_cursor) {
    static int occ[16] = {0};
#define HWBRJ_PROBE_CASE(M)                                                                                         \
    case M: {                                                                                                       \
        const int smem = kProbeWarps * kProbeSmemPerWarp(M);                                                        \
        if (!occ[M]) {                                                                                              \
            CK(cudaFuncSetAttribute(k_probe_compact<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));        \
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[M], k_probe_compact<M>, kProbeWarps * 32, smem)); \
            occ[M] = std::max(occ[M], 1);                                                                           \
            if (g.probe_carveout >= 0)                                                                              \
                CK(cudaFuncSetAttribute(k_probe_compact<M>, cudaFuncAttributePreferredSharedMemoryCarveout,          \
                                        g.probe_carveout));                                                         \
        }                                                                                                           \
        /* measured on B200: 4 CTAs/SM beats the occupancy maximum (less L2 thrash of the filter range) */          \
        const int grid = g.sms * (g.probe_ctas_per_sm ? g.probe_ctas_per_sm : std::min(occ[M], 4));                 \
        k_probe_compact<M><<<grid, kProbeWarps * 32, smem, g.stream>>>(in, n, n_ptr, bp, g.d_crc, out, cursor,     \
                                                                       defer_out, defer_cursor);                    \
        break;                                                                                                      \
    }
    if (g.probe_staged && bp.k >= 2u && !(mode & (2 | 8))) {  // experimental staged probe for k >= 2 (HWBRJ_PROBE_STAGED=1)
        const int smem = kProbeWarps * kProbeSmemPerWarp(0);
        const int grid = g.sms * (g.probe_ctas_per_sm ? g.probe_ctas_per_sm : 4);
#define HWBRJ_STAGED_CASE(M)                                                                                        \
    case M: {                                                                                                       \
        static bool attr = false;                                                                                   \
        if (!attr) {                                                                                                \
            CK(cudaFuncSetAttribute(k_probe_staged<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));         \
            attr = true;                                                                                            \
        }                                                                                                           \
        k_probe_staged<M><<<grid, kProbeWarps * 32, smem, g.stream>>>(in, n, n_ptr, bp, g.d_crc, out, cursor);      \
        return;                                                                                                     \
    }
        switch (mode) {
            HWBRJ_STAGED_CASE(0) HWBRJ_STAGED_CASE(1) HWBRJ_STAGED_CASE(4) HWBRJ_STAGED_CASE(5)
            default: break;
        }
#undef HWBRJ_STAGED_CASE
    }
    switch (mode) {
        HWBRJ_PROBE_CASE(0) HWBRJ_PROBE_CASE(1) HWBRJ_PROBE_CASE(2) HWBRJ_PROBE_CASE(3)
        HWBRJ_PROBE_CASE(4) HWBRJ_PROBE_CASE(5) HWBRJ_PROBE_CASE(6) HWBRJ_PROBE_CASE(7)
        HWBRJ_PROBE_CASE(12) HWBRJ_PROBE_CASE(13) HWBRJ_PROBE_CASE(14) HWBRJ_PROBE_CASE(15)
        default: die("bad probe mode %d", mode);
    }
#undef HWBRJ_PROBE_CASE
}

// All range passes of the S-side probe. With deferral (default) pass i reads what pass i-1 deferred, so S itself is
// read once; buffers d0/d1 (|S| tuples each) ping-pong. Returns the number of kernel launches.
static int run_probe(const uint2* dS, uint64_t nS, BloomParams bp, int nranges, uint2* out, Control* ctrl, uint2* d0,
                     uint2* d1) {
    const int base_mode = (bp.blocked ? 1 : 0) | (bp.k == 1u ? 2 : 0);
    bp.nranges = (uint32_t)nranges;
    if (nranges == 1) {
        launch_probe_mode(base_mode, dS, nS, nullptr, bp, out, &ctrl->survivors, nullptr, nullptr);
        return 1;
    }
    const bool defer = g.defer_ranges && d0 && (nranges == 2 || d1);
    const uint2* in = dS;
    const unsigned long long* n_ptr = nullptr;
    for (int r = 0; r < nranges; r++) {
        bp.range_id = (uint32_t)r;
        if (!defer) {
            launch_probe_mode(base_mode | 4, dS, nS, nullptr, bp, out, &ctrl->survivors, nullptr, nullptr);
        } else if (r + 1 < nranges) {
            uint2* dout = (r & 1) ? d1 : d0;
            launch_probe_mode(base_mode | 4 | 8, in, nS, n_ptr, bp, out, &ctrl->survivors, dout, &ctrl->defer[r]);
            in = dout;
            n_ptr = &ctrl->defer[r];
        } else {
            BloomParams last = bp;  // everything left belongs to the last range: no range test needed
            last.nranges = 1;
            launch_probe_mode(base_mode, in, nS, n_ptr, last, out, &ctrl->survivors, nullptr, nullptr);
        }
    }
    return nranges;
}


// histogram already in `hist`; runs scan + 1 or 2 scatter passes. n_dev (optional) = device-side tuple count.
// Partition function of the join: key & (2^bits-1) (the reference's radix clustering), or -- for a BASIC k<=1 filter
// -- the filter-slice index (crapwow(42,key) & (m-1)) >> (log2 m - bits), which lets K1' build the filter in shared memory.
struct PartFn {
    bool hash = false;
    uint32_t seed = 42u, size_mask = 0u;
    int log2m = 0;
};

static const uint2* run_partition(const uint2* in, uint64_t n, const unsigned long long* n_dev, int bits, int b2,
                                  uint32_t* hist, uint32_t* off, uint2* t1, uint2* t2, int& launches,
                                  cudaStream_t stream = nullptr, bool second_set = false, PartFn pf = PartFn()) {
    if (!stream) stream = g.stream;
    uint32_t* cur1 = second_set ? g.cur1b.as<uint32_t>() : g.cur1.as<uint32_t>();
    uint32_t* cur2 = second_set ? g.cur2b.as<uint32_t>() : g.cur2.as<uint32_t>();
    uint32_t* tiles = second_set ? g.tilesb.as<uint32_t>() : g.tiles.as<uint32_t>();
    const uint32_t P = 1u << bits;
    const uint32_t pmask = P - 1u;
    const int b1 = bits - b2;
    k_scan<<<1, 1024, 0, stream>>>(hist, P, (uint32_t)b2, off, cur1, cur2, tiles);
    launches++;
    BinFn fn;
    memset(&fn, 0, sizeof(fn));
    fn.pmask = pmask;
    fn.b2 = (uint32_t)b2;
    fn.submask = (1u << b2) - 1u;
    if (pf.hash) {
        fn.seed = pf.seed;
        fn.size_mask = pf.size_mask;
        fn.oshift = (uint32_t)(pf.log2m - b1);  // level 1: the top b1 bits of the slice index
        fn.binmask = 0xFFFFFFFFu;
        k_scatter<3><<<g.sms * g.occ_scatter1, kScatterThreads, kScatterSmem, stream>>>(
            in, t1, reinterpret_cast<const uint64_t*>(n_dev), n, off, tiles, cur1, fn, g.d_crc, 1u << b1);
        launches++;
        if (b2 == 0) return t1;
        fn.oshift = (uint32_t)(pf.log2m - bits);  // level 2: the low b2 bits of the slice index
        fn.binmask = (1u << b2) - 1u;
        k_scatter<5><<<g.sms * g.occ_scatter2, kScatterThreads, kScatterSmem, stream>>>(t1, t2, nullptr, n, off, tiles,
                                                                                       cur2, fn, g.d_crc, 1u << b2);
        launches++;
        return t2;
    }
    k_scatter<1><<<g.sms * g.occ_scatter1, kScatterThreads, kScatterSmem, stream>>>(
        in, t1, reinterpret_cast<const uint64_t*>(n_dev), n, off, tiles, cur1, fn, g.d_crc, 1u << b1);
    launches++;
    if (b2 == 0) return t1;
    k_scatter<2><<<g.sms * g.occ_scatter2, kScatterThreads, kScatterSmem, stream>>>(t1, t2, nullptr, n, off, tiles, cur2,
                                                                                   fn, g.d_crc, 1u << b2);
    launches++;
    return t2;
}

static void ensure_workspace(uint64_t nR, uint64_t nS, const bloom_filter_args_t* args) {
    const size_t P = 1u << kMaxRadixBits;
    if (args) g.filter.ensure(std::max<uint64_t>(args->m / 8, 4));
    g.histR.ensure(P * 4);
    g.histS.ensure(P * 4);
    g.offR.ensure((P + 1) * 4);
    g.offS.ensure((P + 1) * 4);
    g.cur1.ensure(((size_t)1 << kMaxLevelBits) * 4);
    g.cur2.ensure(P * 4);
    g.tiles.ensure((((size_t)1 << kMaxLevelBits) + 1) * 4);
    g.cur1b.ensure(((size_t)1 << kMaxLevelBits) * 4);
    g.cur2b.ensure(P * 4);
    g.tilesb.ensure((((size_t)1 << kMaxLevelBits) + 1) * 4);
    g.work.ensure((P + 1) * 4);
    g.work_part.ensure((P + nS / kSChunk + 2) * 4);  // one entry per join work item
    g.ctrl.ensure(sizeof(Control));
    g.rt1.ensure(std::max<uint64_t>(nR, 1) * 8);
    g.rp.ensure(std::max<uint64_t>(nR, 1) * 8);
    g.sc.ensure(std::max<uint64_t>(nS, 1) * 8);
    g.st1.ensure(std::max<uint64_t>(nS, 1) * 8);
}

// The join on device-resident relations. args == nullptr: plain radix join.
// copies the result words of a join out of the control block (async / graph-captured joins)
__global__ void k_export_results(const Control* c, unsigned long long* out) {
    if (threadIdx.x == 0) {
        out[0] = c->acc.matches;
        out[1] = c->acc.cpair;
        out[2] = c->acc.crpay;
        out[3] = c->acc.cspay;
        out[4] = c->acc.ckey;
        out[5] = c->survivors;
    }
}

// S arriving from the host in chunks: chunk c (chunk_tuples tuples, the last one shorter) is complete when ev[c] fires
struct SFeed {
    uint64_t chunk_tuples;
    int nchunks;
    cudaEvent_t* ev;
};

static void run_join(const uint2* dR, uint64_t nR, const uint2* dS, uint64_t nS, const bloom_filter_args_t* args,
                     hwbrj_stats_t& st, const unsigned long long* nR_dev = nullptr, uint64_t nR_expect = 0,
                     const unsigned long long* nS_dev = nullptr, const SFeed* feed = nullptr,
                     unsigned long long* d_async_out = nullptr) {
    // d_async_out != nullptr: enqueue only (no events, no host synchronisation -- capturable in a CUDA graph); the six
    // result words {matches, cpair, crpay, cspay, ckey, survivors} are left in d_async_out (device)
    const bool async_mode = d_async_out != nullptr;
    auto rec = [&](cudaEvent_t e, cudaStream_t s) {
        if (!async_mode) CK(cudaEventRecord(e, s));
    };
    // nR/nS are exact counts, or capacities when the real counts live on the device (nR_dev/nS_dev)
    // 32-bit tuple indices; the slack keeps "index + one batch of loads" from wrapping in the kernels
    if (nR >= (1ull << 32) - (1ull << 20) || nS >= (1ull << 32) - (1ull << 20))
        die("relations of 2^32 - 2^20 or more tuples are not supported");
    if (args && check_args_impl(args, true)) die("invalid Bloom filter arguments");
    if (args && nS_dev) die("device-side S count is only supported for the filter-less join");
    ensure_workspace(nR, nS, args);
    // the R side may have been partitioned already (hwbrj_join_prepare_r, filter-less joins of exactly this relation)
    const bool prepared = !args && g.prep.valid && g.prep.d == dR && g.prep.n == nR && g.prep.n_dev == nR_dev;
    g.prep.valid = false;  // one use; a stale preparation of another relation is simply dropped
    const int bits = prepared ? g.prep.bits : pick_bits(nR_dev ? nR_expect : nR);
    const int b2 = bits > kMaxLevelBits ? bits / 2 : 0;
    const uint32_t P = 1u << bits;
    const uint32_t pmask = P - 1u;
    int launches = 0;
    Control* ctrl = g.ctrl.as<Control>();

    // ---- untimed set-up (the reference allocates and zeroes its filter before the timed region, :1583) ----
    rec(g.ev[0], g.stream);
    if (args) CK(cudaMemsetAsync(g.filter.p, 0, std::max<uint64_t>(args->m / 8, 4), g.stream));
    if (!prepared) CK(cudaMemsetAsync(g.histR.p, 0, P * 4, g.stream));  // else: in use on the side stream
    CK(cudaMemsetAsync(g.histS.p, 0, P * 4, g.stream));
    CK(cudaMemsetAsync(g.ctrl.p, 0, sizeof(Control), g.stream));

    // ---- timed region ----------------------------------------------------------------------------------------
    rec(g.ev[1], g.stream);
    BloomParams bp;
    memset(&bp, 0, sizeof(bp));
    int nranges = 1;
    const int hist_smem = (int)((P + kCrcSmemWords) * 4);
    const int grid_hist = g.sms * 2;
    // Hash-partitioned variant (BASIC, k <= 1): the join partitions on the filter-slice index, so each partition owns a
    // contiguous m/P-bit slice of the filter and K1' builds it in shared memory -- no global atomics, R read once less.
    PartFn pf;
    // Used when the filter is too big for its atomics to stay L2-resident (> 32 MiB: K1 1.25 -> 0.68 ms at C1); for
    // small filters the plain atomics are faster (C0: 1.71 vs 1.80 ms) and the radix table index has shorter chains.
    if (args && g.hash_partition && args->variant == BASIC && args->k <= 1 && bits >= 1 &&
        (args->m / 8 > (32ull << 20) || g.hash_partition_force) && ilog2_u64(args->m) >= bits + 5 &&
        (args->m >> bits) / 8 <= 96 * 1024 && !g.overlap_r_partition) {
        pf.hash = true;
        pf.size_mask = (uint32_t)(args->m - 1);
        pf.log2m = ilog2_u64(args->m);
    }
    const uint32_t hshift = pf.hash ? (uint32_t)(pf.log2m - bits) : 0u;
    if (args) {
        bp = make_bloom(args, 42u, g.filter.as<uint32_t>());  // seed 42: parallel_radix_join_bloom.c:1583,1823
        nranges = pick_ranges(args);
        bp.nranges = (uint32_t)nranges;
        bp.range_shift = (uint32_t)(ilog2_u64(args->m) - ilog2_u64((uint64_t)nranges));
        if (pf.hash) {
            k_build_hist<false, true><<<grid_hist, 1024, hist_smem, g.stream>>>(dR, nR, nR_dev, bp, g.d_crc,
                                                                               g.histR.as<uint32_t>(), pmask, hshift);
            launches++;
        } else {
            for (int r = 0; r < nranges; r++) {
                bp.range_id = (uint32_t)r;
                k_build_hist<true><<<grid_hist, 1024, hist_smem, g.stream>>>(dR, nR, nR_dev, bp, g.d_crc, g.histR.as<uint32_t>(), pmask);
                launches++;
            }
        }
    } else if (!prepared) {
        k_build_hist<false><<<grid_hist, 1024, hist_smem, g.stream>>>(dR, nR, nR_dev, bp, g.d_crc, g.histR.as<uint32_t>(), pmask);
        launches++;
    }
    rec(g.ev[2], g.stream);
    // R partitioning (HBM-bound) runs on a side stream underneath the S probe (L1TEX/issue-bound) when a filter is used
    const bool overlap = g.overlap_r_partition && args != nullptr;
    const uint2* Rp;
    if (prepared) {
        Rp = g.prep.Rp;  // partitioned on the side stream; joined below, before the work list
        launches += g.prep.launches;
    } else if (overlap) {
        CK(cudaStreamWaitEvent(g.side_stream, g.ev[2], 0));
        rec(g.ev_side[0], g.side_stream);
        Rp = run_partition(dR, nR, nR_dev, bits, b2, g.histR.as<uint32_t>(), g.offR.as<uint32_t>(), g.rt1.as<uint2>(),
                           g.rp.as<uint2>(), launches, g.side_stream, true);
        rec(g.ev_side[1], g.side_stream);
    } else {
        Rp = run_partition(dR, nR, nR_dev, bits, b2, g.histR.as<uint32_t>(), g.offR.as<uint32_t>(), g.rt1.as<uint2>(),
                           g.rp.as<uint2>(), launches, nullptr, false, pf);
    }
    rec(g.ev_side[2], g.stream);
    if (pf.hash && args->k >= 1) {  // K1': the filter, slice by slice, from the partitioned R (k = 0 sets no bit)
        const uint32_t slice_words = (uint32_t)((args->m >> bits) / 32);
        k_filter_from_parts<<<g.sms * 4, 512, 2 * slice_words * 4, g.stream>>>(Rp, g.offR.as<uint32_t>(), P, g.filter.as<uint32_t>(),
                                                                          slice_words, 42u, pf.size_mask);
        launches++;
    }
    rec(g.ev[3], g.stream);
    const uint2* Sin = dS;
    const unsigned long long* n_dev = nS_dev;
    if (args) {
        if (nranges > 2 && g.defer_ranges) g.d1.ensure(std::max<uint64_t>(nS, 1) * 8);
        if (feed) {  // probe every chunk as soon as its host->device copy has completed
            for (int c = 0; c < feed->nchunks; c++) {
                const uint64_t off = (uint64_t)c * feed->chunk_tuples;
                const uint64_t cnt = std::min<uint64_t>(feed->chunk_tuples, nS - off);
                CK(cudaStreamWaitEvent(g.stream, feed->ev[c], 0));
                launches += run_probe(dS + off, cnt, bp, nranges, g.sc.as<uint2>(), ctrl, g.st1.as<uint2>(),
                                      nranges > 2 ? g.d1.as<uint2>() : nullptr);
            }
        } else {
            launches += run_probe(dS, nS, bp, nranges, g.sc.as<uint2>(), ctrl, g.st1.as<uint2>(),
                                  nranges > 2 ? g.d1.as<uint2>() : nullptr);
        }
        Sin = g.sc.as<uint2>();
        n_dev = &ctrl->survivors;
    }
    rec(g.ev[4], g.stream);  // ms_probe = the K2 launches only
    // radix histogram of the tuples that go on to the join (survivors, or all of S without a filter)
    if (pf.hash)
        k_build_hist<false, true><<<grid_hist, 1024, hist_smem, g.stream>>>(Sin, nS, n_dev, bp, g.d_crc,
                                                                           g.histS.as<uint32_t>(), pmask, hshift);
    else
        k_build_hist<false><<<grid_hist, 1024, hist_smem, g.stream>>>(Sin, nS, n_dev, bp, g.d_crc, g.histS.as<uint32_t>(), pmask);
    launches++;
    // with a filter: sc -> st1 -> sc ; without: dS -> st1 -> sc
    const uint2* Sp = run_partition(Sin, nS, n_dev, bits, b2, g.histS.as<uint32_t>(), g.offS.as<uint32_t>(),
                                    g.st1.as<uint2>(), g.sc.as<uint2>(), launches, nullptr, false, pf);
    rec(g.ev[5], g.stream);
    if (overlap) CK(cudaStreamWaitEvent(g.stream, g.ev_side[1], 0));
    if (prepared) CK(cudaStreamWaitEvent(g.stream, g.ev_prep_done, 0));
    k_worklist<<<1, 1024, 0, g.stream>>>(g.offR.as<uint32_t>(), g.offS.as<uint32_t>(), P, g.work.as<uint32_t>(),
                                         g.work_part.as<uint32_t>());
    launches++;
    if (pf.hash)
        k_join<true><<<g.sms * g.occ_join, kJoinThreads, kJoinSmemBytes, g.stream>>>(
            Rp, g.offR.as<uint32_t>(), Sp, g.offS.as<uint32_t>(), g.work.as<uint32_t>(), g.work_part.as<uint32_t>(), P, (uint32_t)bits,
            &ctrl->item_counter, &ctrl->acc);
    else
        k_join<false><<<g.sms * g.occ_join, kJoinThreads, kJoinSmemBytes, g.stream>>>(
            Rp, g.offR.as<uint32_t>(), Sp, g.offS.as<uint32_t>(), g.work.as<uint32_t>(), g.work_part.as<uint32_t>(), P, (uint32_t)bits,
            &ctrl->item_counter, &ctrl->acc);
    launches++;
    g.last_Rp = Rp;
    g.last_Sp = Sp;
    g.last_P = P;
    g.last_bits = (uint32_t)bits;
    g.last_hash = pf.hash;
    rec(g.ev[6], g.stream);
    if (async_mode) {
        k_export_results<<<1, 32, 0, g.stream>>>(ctrl, d_async_out);
        st.kernel_launches = launches + 1;
        st.radix_bits = bits;
        st.range_passes = nranges;
        return;
    }
    Control h;
    CK(cudaMemcpyAsync(&h, ctrl, sizeof(Control), cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaGetLastError());

    auto ms = [&](int a, int b) {
        float v = 0;
        CK(cudaEventElapsedTime(&v, g.ev[a], g.ev[b]));
        return v;
    };
    st.matches = (int64_t)h.acc.matches;
    st.filtered = args ? (int64_t)h.survivors : -1;
    st.checksum_pair = h.acc.cpair;
    st.checksum_rpay = h.acc.crpay;
    st.checksum_spay = h.acc.cspay;
    st.checksum_key = h.acc.ckey;
    st.ms_memset = ms(0, 1);
    st.ms_total = ms(1, 6);
    st.ms_build = ms(1, 2);
    st.ms_part_r = ms(2, 3);
    if (pf.hash) {  // the filter is built after the scatter passes: book it under "build"
        float f = 0;
        CK(cudaEventElapsedTime(&f, g.ev_side[2], g.ev[3]));
        st.ms_build += f;
        st.ms_part_r -= f;
    }
    if (overlap) CK(cudaEventElapsedTime(&st.ms_part_r, g.ev_side[0], g.ev_side[1]));  // overlapped with ms_probe
    st.ms_probe = ms(3, 4);
    st.ms_part_s = ms(4, 5);
    st.ms_join = ms(5, 6);
    st.kernel_launches = launches;
    st.radix_bits = bits;
    st.range_passes = nranges;
    st.n_gpus = 1;
    st.d2h_bytes += sizeof(Control);
}

static void print_reference_lines(const hwbrj_stats_t& st, uint64_t nS, bool bloom_line) {
    if (g.quiet) return;
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

// host-buffer entry: copy in, join, fill result_t (join_init_run, :1561-1778)
static result_t* host_join(relation_t* relR, relation_t* relS, int nthreads, bloom_filter_args_t* args,
                           bool print_filtered) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    if (!relR || !relS) die("NULL relation");
    auto t0 = std::chrono::steady_clock::now();
    hwbrj_stats_t st;
    memset(&st, 0, sizeof(st));
    const uint64_t nR = relR->num_tuples, nS = relS->num_tuples;
    g.inR.ensure(std::max<uint64_t>(nR, 2) * 8);
    g.inS.ensure(std::max<uint64_t>(nS, 2) * 8);
    st.h2d_bytes = (nR + nS) * 8;
    if (g.overlap_h2d && args && nS >= (1u << 22)) {
        // copies on their own stream; R first, then S in <= 64 chunks, each followed by an event the probe waits on
        SFeed feed;
        feed.nchunks = (int)std::min<uint64_t>(64, (nS + (1u << 22) - 1) >> 22);
        feed.chunk_tuples = (((nS + feed.nchunks - 1) / feed.nchunks) + 1) & ~1ull;  // even: chunks stay 16-byte aligned
        feed.nchunks = (int)((nS + feed.chunk_tuples - 1) / feed.chunk_tuples);
        feed.ev = g.ev_chunk + 1;
        CK(cudaEventRecord(g.ev[7], g.copy_stream));
        if (nR) CK(cudaMemcpyAsync(g.inR.p, relR->tuples, nR * 8, cudaMemcpyHostToDevice, g.copy_stream));
        CK(cudaEventRecord(g.ev_chunk[0], g.copy_stream));
        for (int c = 0; c < feed.nchunks; c++) {
            const uint64_t off = (uint64_t)c * feed.chunk_tuples;
            const uint64_t cnt = std::min<uint64_t>(feed.chunk_tuples, nS - off);
            CK(cudaMemcpyAsync(g.inS.as<uint2>() + off, relS->tuples + off, cnt * 8, cudaMemcpyHostToDevice, g.copy_stream));
            CK(cudaEventRecord(feed.ev[c], g.copy_stream));
        }
        CK(cudaEventRecord(g.ev_side[3], g.copy_stream));
        CK(cudaStreamWaitEvent(g.stream, g.ev_chunk[0], 0));  // the R phase needs all of R
        run_join(g.inR.as<uint2>(), nR, g.inS.as<uint2>(), nS, args, st, nullptr, 0, nullptr, &feed);
        CK(cudaEventElapsedTime(&st.ms_h2d, g.ev[7], g.ev_side[3]));
    } else {
        CK(cudaEventRecord(g.ev[7], g.stream));
        h2d(g.inR.p, relR->tuples, nR * 8);
        h2d(g.inS.p, relS->tuples, nS * 8);
        CK(cudaEventRecord(g.ev[0], g.stream));
        CK(cudaStreamSynchronize(g.stream));
        CK(cudaEventElapsedTime(&st.ms_h2d, g.ev[7], g.ev[0]));
        run_join(g.inR.as<uint2>(), nR, g.inS.as<uint2>(), nS, args, st);
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
int hwbrj_last_stats(hwbrj_stats_t* out) {
    if (!out) return -1;
    *out = g.last;
    return 0;
}
int64_t hwbrj_last_filtered(void) { return g.last.filtered; }
int hwbrj_last_filter(unsigned char* bitmap_out, uint64_t nbytes) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    if (!bitmap_out || nbytes > g.filter.cap) return -1;
    CK(cudaMemcpy(bitmap_out, g.filter.p, nbytes, cudaMemcpyDeviceToHost));
    return 0;
}
uint64_t hwbrj_last_checksum(void) { return g.last.checksum_pair; }
void hwbrj_set_quiet(int quiet) { g.quiet = quiet != 0; }
void hwbrj_set_radix_bits(int bits) { g.radix_bits_override = bits; }
void hwbrj_set_range_passes(int passes) { g.range_passes_override = passes; }
void hwbrj_set_overlap_h2d(int on) { g.overlap_h2d = on != 0; }
void hwbrj_set_hash_partition(int mode) {
    g.hash_partition = mode != 0;
    g.hash_partition_force = mode == 2;
}
const char* hwbrj_version(void) { return "hwbrj-b200 0.1 (sm_100a)"; }
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
    uint64_t n;                        // tuple count, or capacity / upper bound when n_dev is set
    bool owned;
    const unsigned long long* n_dev;   // optional: the real count lives on the device (no host round trip)
    uint64_t n_expect;                 // sizing hint when n_dev is set
};

hwbrj_rel_t* hwbrj_rel_upload(const tuple_t* tuples, uint64_t n) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    hwbrj_rel_t* r = new hwbrj_rel;
    r->n = n;
    r->owned = true;
    r->n_dev = nullptr;
    r->n_expect = n;
    CK(cudaMalloc(&r->d, std::max<uint64_t>(n, 2) * 8 + 64));
    if (n) CK(cudaMemcpy(r->d, tuples, n * 8, cudaMemcpyHostToDevice));
    return r;
}

hwbrj_rel_t* hwbrj_rel_generate(int kind, uint64_t n, uint64_t r, double q, uint64_t seed) {
    return hwbrj_rel_generate_shard(kind, n, r, q, seed, 0, n);
}

hwbrj_rel_t* hwbrj_rel_generate_shard(int kind, uint64_t n, uint64_t r, double q, uint64_t seed, uint64_t begin,
                                      uint64_t count) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    if (begin > n) begin = n;
    if (count > n - begin) count = n - begin;
    hwbrj_rel_t* rel = new hwbrj_rel;
    rel->n = count;
    rel->owned = true;
    rel->n_dev = nullptr;
    rel->n_expect = count;
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
    r->n_dev = nullptr;
    r->n_expect = n;
    return r;
}
hwbrj_rel_t* hwbrj_rel_wrap_counted(void* device_tuples, uint64_t capacity, const void* d_count, uint64_t expected) {
    hwbrj_rel_t* r = hwbrj_rel_wrap(device_tuples, capacity);
    if (!r) return nullptr;
    r->n_dev = reinterpret_cast<const unsigned long long*>(d_count);
    r->n_expect = expected;
    return r;
}
void* hwbrj_rel_ptr(const hwbrj_rel_t* rel) { return rel ? rel->d : nullptr; }
void hwbrj_set_stream(void* cuda_stream) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    g.stream = reinterpret_cast<cudaStream_t>(cuda_stream);  // 0 is a valid handle: the legacy default stream
}
void hwbrj_reset_stream(void) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    g.stream = g.own_stream;
}
int hwbrj_set_device(int device) {
    // must precede the first library call of the process (one process per GPU); buffers live on that device
    if (g.inited && g.dev != device) return -1;
    if (cudaSetDevice(device) != cudaSuccess) {
        cudaGetLastError();
        return -2;
    }
    return 0;
}
int hwbrj_sync(void) {
    init_ctx();
    CK(cudaStreamSynchronize(g.stream));
    return 0;
}

int hwbrj_join_device(const hwbrj_rel_t* R, const hwbrj_rel_t* S, const bloom_filter_args_t* args, hwbrj_stats_t* out) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    if (!R || !S) return -1;
    hwbrj_stats_t st;
    memset(&st, 0, sizeof(st));
    run_join(R->d, R->n, S->d, S->n, args, st, R->n_dev, R->n_expect, S->n_dev);
    g.last = st;
    if (out) *out = st;
    return 0;
}

int hwbrj_join_device_async(const hwbrj_rel_t* R, const hwbrj_rel_t* S, const bloom_filter_args_t* args, void* d_out6) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    if (!R || !S || !d_out6) return -1;
    hwbrj_stats_t st;
    memset(&st, 0, sizeof(st));
    run_join(R->d, R->n, S->d, S->n, args, st, R->n_dev, R->n_expect, S->n_dev, nullptr,
             reinterpret_cast<unsigned long long*>(d_out6));
    return st.kernel_launches;
}

int hwbrj_join_prepare_r(const hwbrj_rel_t* R) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    if (!R || R->n >= (1ull << 32) - (1ull << 20)) return -1;
    ensure_workspace(R->n, 1, nullptr);
    const int bits = pick_bits(R->n_dev ? R->n_expect : R->n);
    const int b2 = bits > kMaxLevelBits ? bits / 2 : 0;
    const uint32_t P = 1u << bits;
    BloomParams bp;
    memset(&bp, 0, sizeof(bp));
    // fork from the caller's stream: everything R depends on has been enqueued there
    CK(cudaEventRecord(g.ev_prep_fork, g.stream));
    CK(cudaStreamWaitEvent(g.side_stream, g.ev_prep_fork, 0));
    CK(cudaMemsetAsync(g.histR.p, 0, P * 4, g.side_stream));
    k_build_hist<false><<<g.sms * 2, 1024, (int)((P + kCrcSmemWords) * 4), g.side_stream>>>(
        R->d, R->n, R->n_dev, bp, g.d_crc, g.histR.as<uint32_t>(), P - 1u);
    int launches = 1;
    g.prep.Rp = run_partition(R->d, R->n, R->n_dev, bits, b2, g.histR.as<uint32_t>(), g.offR.as<uint32_t>(),
                              g.rt1.as<uint2>(), g.rp.as<uint2>(), launches, g.side_stream, true);
    CK(cudaEventRecord(g.ev_prep_done, g.side_stream));
    CK(cudaGetLastError());
    g.prep.valid = true;
    g.prep.d = R->d;
    g.prep.n = R->n;
    g.prep.n_dev = R->n_dev;
    g.prep.bits = bits;
    g.prep.launches = launches;
    return launches;
}

void* hwbrj_host_alloc(uint64_t bytes) {
    init_ctx();
    void* p = nullptr;
    CK(cudaHostAlloc(&p, std::max<uint64_t>(bytes, 8), cudaHostAllocDefault));
    return p;
}
void hwbrj_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int hwbrj_hash_many(int which, uint32_t seed, const int32_t* keys, uint64_t n, uint32_t* out) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
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

int hwbrj_bloom_build(const tuple_t* R, uint64_t nR, const bloom_filter_args_t* args, uint32_t seed,
                      unsigned char* bitmap_out) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    if (!args || check_args_impl(args, true)) return -1;
    g.inR.ensure(std::max<uint64_t>(nR, 2) * 8);
    g.filter.ensure(std::max<uint64_t>(args->m / 8, 4));
    g.histR.ensure(((size_t)1 << kMaxRadixBits) * 4);
    h2d(g.inR.p, R, nR * 8);
    CK(cudaMemsetAsync(g.filter.p, 0, std::max<uint64_t>(args->m / 8, 4), g.stream));
    CK(cudaMemsetAsync(g.histR.p, 0, 4, g.stream));
    BloomParams bp = make_bloom(args, seed, g.filter.as<uint32_t>());
    int nranges = pick_ranges(args);
    bp.nranges = (uint32_t)nranges;
    bp.range_shift = (uint32_t)(ilog2_u64(args->m) - ilog2_u64((uint64_t)nranges));
    for (int r = 0; r < nranges; r++) {
        bp.range_id = (uint32_t)r;
        k_build_hist<true><<<g.sms * 2, 1024, (1 + kCrcSmemWords) * 4, g.stream>>>(g.inR.as<uint2>(), nR, nullptr, bp, g.d_crc,
                                                                          g.histR.as<uint32_t>(), 0u);
    }
    CK(cudaMemcpyAsync(bitmap_out, g.filter.p, args->m / 8, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaGetLastError());
    return 0;
}

int64_t hwbrj_bloom_probe(const unsigned char* bitmap, const tuple_t* S, uint64_t nS, const bloom_filter_args_t* args,
                          uint32_t seed, tuple_t* survivors_out) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    if (!args || check_args_impl(args, true)) return -1;
    g.inS.ensure(std::max<uint64_t>(nS, 2) * 8);
    g.sc.ensure(std::max<uint64_t>(nS, 1) * 8);
    g.filter.ensure(std::max<uint64_t>(args->m / 8, 4));
    g.histS.ensure(((size_t)1 << kMaxRadixBits) * 4);
    g.ctrl.ensure(sizeof(Control));
    h2d(g.inS.p, S, nS * 8);
    h2d(g.filter.p, bitmap, args->m / 8);
    CK(cudaMemsetAsync(g.histS.p, 0, 4, g.stream));
    CK(cudaMemsetAsync(g.ctrl.p, 0, sizeof(Control), g.stream));
    Control* ctrl = g.ctrl.as<Control>();
    BloomParams bp = make_bloom(args, seed, g.filter.as<uint32_t>());
    int nranges = pick_ranges(args);
    bp.nranges = (uint32_t)nranges;
    bp.range_shift = (uint32_t)(ilog2_u64(args->m) - ilog2_u64((uint64_t)nranges));
    g.st1.ensure(std::max<uint64_t>(nS, 1) * 8);
    if (nranges > 2 && g.defer_ranges) g.d1.ensure(std::max<uint64_t>(nS, 1) * 8);
    run_probe(g.inS.as<uint2>(), nS, bp, nranges, g.sc.as<uint2>(), ctrl, g.st1.as<uint2>(),
              nranges > 2 ? g.d1.as<uint2>() : nullptr);
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
// `capacity`: then only the first `capacity` slots were written and the caller retries with a larger buffer).
static int64_t materialize_last(uint2* d_pairs, uint64_t capacity) {
    if (!g.last_Rp || !g.last_Sp) return -1;
    Control* ctrl = g.ctrl.as<Control>();
    CK(cudaMemsetAsync(&ctrl->acc, 0, sizeof(JoinAccum), g.stream));
    CK(cudaMemsetAsync(&ctrl->item_counter, 0, sizeof(uint32_t), g.stream));
    CK(cudaMemsetAsync(&ctrl->pair_cursor, 0, sizeof(unsigned long long), g.stream));
    const int smem = kJoinSmemBytes;
    if (g.last_hash)
        k_join<true, true><<<g.sms * g.occ_join, kJoinThreads, smem, g.stream>>>(
            g.last_Rp, g.offR.as<uint32_t>(), g.last_Sp, g.offS.as<uint32_t>(), g.work.as<uint32_t>(), g.work_part.as<uint32_t>(), g.last_P,
            g.last_bits, &ctrl->item_counter, &ctrl->acc, d_pairs, &ctrl->pair_cursor, capacity);
    else
        k_join<false, true><<<g.sms * g.occ_join, kJoinThreads, smem, g.stream>>>(
            g.last_Rp, g.offR.as<uint32_t>(), g.last_Sp, g.offS.as<uint32_t>(), g.work.as<uint32_t>(), g.work_part.as<uint32_t>(), g.last_P,
            g.last_bits, &ctrl->item_counter, &ctrl->acc, d_pairs, &ctrl->pair_cursor, capacity);
    unsigned long long cnt = 0;
    CK(cudaMemcpyAsync(&cnt, &ctrl->pair_cursor, 8, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaGetLastError());
    return (int64_t)cnt;
}

int64_t hwbrj_materialize_last(tuple_t* pairs_out, uint64_t capacity) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    if (!pairs_out && capacity) return -1;
    g.pairs.ensure(std::max<uint64_t>(capacity, 1) * 8);
    int64_t n = materialize_last(g.pairs.as<uint2>(), capacity);
    if (n > 0) CK(cudaMemcpy(pairs_out, g.pairs.p, std::min<uint64_t>((uint64_t)n, capacity) * 8, cudaMemcpyDeviceToHost));
    return n;
}

int64_t hwbrj_materialize_last_device(void* d_pairs, uint64_t capacity) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    return materialize_last(reinterpret_cast<uint2*>(d_pairs), capacity);
}

// device analogue of the reference's FPR measurement (test_bloom_fpr, unit_tests.c:191-241): build a filter with the
// given seed from R, probe S, return how many S keys pass
int64_t hwbrj_fpr_count(const hwbrj_rel_t* R, const hwbrj_rel_t* S, const bloom_filter_args_t* args, uint32_t seed) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    if (!R || !S || !args || check_args_impl(args, true)) return -1;
    g.filter.ensure(std::max<uint64_t>(args->m / 8, 4));
    g.histR.ensure(((size_t)1 << kMaxRadixBits) * 4);
    g.sc.ensure(std::max<uint64_t>(S->n, 1) * 8);
    g.st1.ensure(std::max<uint64_t>(S->n, 1) * 8);
    g.ctrl.ensure(sizeof(Control));
    CK(cudaMemsetAsync(g.filter.p, 0, std::max<uint64_t>(args->m / 8, 4), g.stream));
    CK(cudaMemsetAsync(g.histR.p, 0, 4, g.stream));
    CK(cudaMemsetAsync(g.ctrl.p, 0, sizeof(Control), g.stream));
    Control* ctrl = g.ctrl.as<Control>();
    BloomParams bp = make_bloom(args, seed, g.filter.as<uint32_t>());
    int nranges = pick_ranges(args);
    bp.nranges = (uint32_t)nranges;
    bp.range_shift = (uint32_t)(ilog2_u64(args->m) - ilog2_u64((uint64_t)nranges));
    for (int r = 0; r < nranges; r++) {
        bp.range_id = (uint32_t)r;
        k_build_hist<true><<<g.sms * 2, 1024, (1 + kCrcSmemWords) * 4, g.stream>>>(R->d, R->n, R->n_dev, bp, g.d_crc,
                                                                                  g.histR.as<uint32_t>(), 0u);
    }
    if (nranges > 2 && g.defer_ranges) g.d1.ensure(std::max<uint64_t>(S->n, 1) * 8);
    run_probe(S->d, S->n, bp, nranges, g.sc.as<uint2>(), ctrl, g.st1.as<uint2>(), nranges > 2 ? g.d1.as<uint2>() : nullptr);
    unsigned long long cnt = 0;
    CK(cudaMemcpyAsync(&cnt, &ctrl->survivors, 8, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaGetLastError());
    return (int64_t)cnt;
}

int hwbrj_radix_partition(const tuple_t* in, uint64_t n, int bits, tuple_t* out, uint64_t* offsets) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    if (bits < 0 || bits > kMaxRadixBits || n >= (1ull << 32)) return -1;
    ensure_workspace(n, 1, nullptr);
    g.inR.ensure(std::max<uint64_t>(n, 2) * 8);
    const uint32_t P = 1u << bits;
    const int b2 = bits > kMaxLevelBits ? bits / 2 : 0;
    h2d(g.inR.p, in, n * 8);
    CK(cudaMemsetAsync(g.histR.p, 0, P * 4, g.stream));
    BloomParams bp;
    memset(&bp, 0, sizeof(bp));
    k_build_hist<false><<<g.sms * 2, 1024, (int)((P + kCrcSmemWords) * 4), g.stream>>>(g.inR.as<uint2>(), n, nullptr, bp, g.d_crc,
                                                                            g.histR.as<uint32_t>(), P - 1u);
    int launches = 0;
    const uint2* res = run_partition(g.inR.as<uint2>(), n, nullptr, bits, b2, g.histR.as<uint32_t>(),
                                     g.offR.as<uint32_t>(), g.rt1.as<uint2>(), g.rp.as<uint2>(), launches);
    std::vector<uint32_t> off32(P + 1);
    if (n) CK(cudaMemcpyAsync(out, res, n * 8, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaMemcpyAsync(off32.data(), g.offR.p, (P + 1) * 4, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaGetLastError());
    for (uint32_t i = 0; i <= P; i++) offsets[i] = off32[i];
    return 0;
}

}  // extern "C"

// ---- multi-GPU building blocks: the host (hwbloomradixjoin_b200/dist.py, one process per GPU)
// orchestrates these between torch.distributed collectives; all of them take raw device pointers and run on the
// stream given to hwbrj_set_stream() ------------------------------------------------------------------------------
namespace hwbrj {
static BinFn owner_fn(int world, const bloom_filter_args_t* slice_args, int& mode) {
    BinFn fn;
    memset(&fn, 0, sizeof(fn));
    const int gbits = ilog2_u64((uint64_t)world);
    fn.seed = 42u;
    fn.binmask = 0xFFFFFFFFu;
    if (slice_args && slice_args->variant == BLOCKED) {
        mode = 4;  // owner = top bits of the block index (all k bits of a key live in that block)
        uint64_t nblocks = slice_args->m / slice_args->B;
        fn.size_mask = (uint32_t)(nblocks - 1);
        fn.oshift = (uint32_t)std::max(0, ilog2_u64(nblocks) - gbits);
    } else if (slice_args) {
        mode = 3;  // owner = top bits of the first bit address crapwow & (m-1)
        fn.size_mask = (uint32_t)(slice_args->m - 1);
        fn.oshift = (uint32_t)std::max(0, ilog2_u64(slice_args->m) - gbits);
    } else {
        mode = 3;  // no sliceable filter: owner = top bits of crapwow(42,key)
        fn.size_mask = 0xFFFFFFFFu;
        fn.oshift = (uint32_t)(32 - gbits);
    }
    return fn;
}
}  // namespace hwbrj

extern "C" {

int hwbrj_owner_partition(const hwbrj_rel_t* in, int world, const bloom_filter_args_t* slice_args, void* d_out,
                          uint64_t* counts_out) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    if (!in || world < 1 || world > 128 || (world & (world - 1)) || in->n >= (1ull << 32)) return -1;
    if (slice_args && check_args_impl(slice_args, true)) return -1;
    if (slice_args && slice_args->variant == BASIC && slice_args->k > 1) slice_args = nullptr;  // not sliceable
    if (slice_args) {
        uint64_t units = slice_args->variant == BLOCKED ? slice_args->m / slice_args->B : slice_args->m;
        if (units < (uint64_t)world) return -1;
    }
    ensure_workspace(1, 1, nullptr);
    int mode = 3;
    BinFn fn = owner_fn(world, slice_args, mode);
    const uint32_t nb = (uint32_t)world;
    CK(cudaMemsetAsync(g.histR.p, 0, nb * 4, g.stream));
    if (mode == 4)
        k_owner_hist<4><<<g.sms * 4, 256, 0, g.stream>>>(in->d, in->n, nullptr, fn, g.d_crc, nb, g.histR.as<uint32_t>());
    else
        k_owner_hist<3><<<g.sms * 4, 256, 0, g.stream>>>(in->d, in->n, nullptr, fn, g.d_crc, nb, g.histR.as<uint32_t>());
    k_scan<<<1, 1024, 0, g.stream>>>(g.histR.as<uint32_t>(), nb, 0u, g.offR.as<uint32_t>(), g.cur1.as<uint32_t>(),
                                     g.cur2.as<uint32_t>(), g.tiles.as<uint32_t>());
    if (mode == 4)
        k_scatter<4><<<g.sms * g.occ_scatter1, kScatterThreads, kScatterSmem, g.stream>>>(
            in->d, reinterpret_cast<uint2*>(d_out), nullptr, in->n, g.offR.as<uint32_t>(), g.tiles.as<uint32_t>(),
            g.cur1.as<uint32_t>(), fn, g.d_crc, nb);
    else
        k_scatter<3><<<g.sms * g.occ_scatter1, kScatterThreads, kScatterSmem, g.stream>>>(
            in->d, reinterpret_cast<uint2*>(d_out), nullptr, in->n, g.offR.as<uint32_t>(), g.tiles.as<uint32_t>(),
            g.cur1.as<uint32_t>(), fn, g.d_crc, nb);
    std::vector<uint32_t> off(nb + 1);
    CK(cudaMemcpyAsync(off.data(), g.offR.p, (nb + 1) * 4, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaGetLastError());
    for (uint32_t i = 0; i < nb; i++) counts_out[i] = off[i + 1] - off[i];
    return 0;
}

// ---- peer memory (CUDA IPC) and the fused partition + all-to-all ---------------------------------------------------
void* hwbrj_symm_alloc(uint64_t bytes) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    void* p = nullptr;
    CK(cudaMalloc(&p, std::max<uint64_t>(bytes, 256)));
    CK(cudaMemset(p, 0, std::max<uint64_t>(bytes, 256)));
    return p;
}
void hwbrj_symm_free(void* p) {
    if (p) cudaFree(p);
}
int hwbrj_ipc_export(void* p, void* handle_out) {
    init_ctx();
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    static_assert(sizeof(h) == HWBRJ_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}
void* hwbrj_ipc_open(const void* handle) {
    init_ctx();
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
int hwbrj_ipc_close(void* p) {
    if (p && cudaIpcCloseMemHandle(p) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return 0;
}

int hwbrj_route_peer(const hwbrj_rel_t* in, int world, const bloom_filter_args_t* slice_args, void* const* peer_bufs,
                     void* const* peer_cursors, uint64_t capacity_tuples, void* d_overflow_flag) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    if (!in || world < 1 || world > kMaxPeers || (world & (world - 1)) || in->n >= (1ull << 32)) return -1;
    if (slice_args && check_args_impl(slice_args, true)) return -1;
    if (slice_args && slice_args->variant == BASIC && slice_args->k > 1) slice_args = nullptr;
    ensure_workspace(1, 1, nullptr);
    int mode = 3;
    BinFn fn = owner_fn(world, slice_args, mode);
    PeerTargets pt;
    memset(&pt, 0, sizeof(pt));
    for (int i = 0; i < world; i++) {
        pt.buf[i] = reinterpret_cast<uint2*>(peer_bufs[i]);
        pt.cursor[i] = reinterpret_cast<unsigned long long*>(peer_cursors[i]);
    }
    pt.capacity = capacity_tuples;
    pt.overflow = reinterpret_cast<unsigned int*>(d_overflow_flag);
    const uint64_t* n_ptr = reinterpret_cast<const uint64_t*>(in->n_dev);
    static int occ3 = 0, occ4 = 0;
    if (!occ3) {
        CK(cudaFuncSetAttribute(k_scatter<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScatterSmem));
        CK(cudaFuncSetAttribute(k_scatter<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScatterSmem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3, k_scatter<3, true>, kScatterThreads, kScatterSmem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ4, k_scatter<4, true>, kScatterThreads, kScatterSmem));
        occ3 = std::max(occ3, 1);
        occ4 = std::max(occ4, 1);
    }
    if (g.route_precount) {  // experimental: one remote claim per owner instead of one per (tile, owner)
        g.route_hist.ensure(kMaxPeers * sizeof(uint32_t));
        g.route_cur.ensure(kMaxPeers * sizeof(unsigned long long));
        CK(cudaMemsetAsync(g.route_hist.p, 0, kMaxPeers * sizeof(uint32_t), g.stream));
        if (mode == 4)
            k_owner_hist<4><<<g.sms * 4, 256, 0, g.stream>>>(in->d, in->n, in->n_dev, fn, g.d_crc, (uint32_t)world,
                                                             g.route_hist.as<uint32_t>());
        else
            k_owner_hist<3><<<g.sms * 4, 256, 0, g.stream>>>(in->d, in->n, in->n_dev, fn, g.d_crc, (uint32_t)world,
                                                             g.route_hist.as<uint32_t>());
        k_route_claim<<<1, 32, 0, g.stream>>>(g.route_hist.as<uint32_t>(), pt, (uint32_t)world,
                                              g.route_cur.as<unsigned long long>());
        for (int i = 0; i < world; i++) pt.cursor[i] = g.route_cur.as<unsigned long long>() + i;  // local sub-allocation
    }
    if (mode == 4)
        k_scatter<4, true><<<g.sms * occ4, kScatterThreads, kScatterSmem, g.stream>>>(
            in->d, nullptr, n_ptr, in->n, nullptr, nullptr, nullptr, fn, g.d_crc, (uint32_t)world, pt);
    else
        k_scatter<3, true><<<g.sms * occ3, kScatterThreads, kScatterSmem, g.stream>>>(
            in->d, nullptr, n_ptr, in->n, nullptr, nullptr, nullptr, fn, g.d_crc, (uint32_t)world, pt);
    CK(cudaGetLastError());
    return 0;
}

// probe that leaves the survivor count on the device (no host round trip): count_out is a device u64, zeroed here
int hwbrj_filter_probe_async(const void* d_filter, const hwbrj_rel_t* S, const bloom_filter_args_t* args, void* d_out,
                             void* d_count_out) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    if (!S || !args || !d_filter || !d_out || !d_count_out || check_args_impl(args, true)) return -1;
    g.ctrl.ensure(sizeof(Control));
    g.st1.ensure(std::max<uint64_t>(S->n, 1) * 8);
    CK(cudaMemsetAsync(d_count_out, 0, 8, g.stream));
    BloomParams bp = make_bloom(args, 42u, reinterpret_cast<uint32_t*>(const_cast<void*>(d_filter)));
    int nranges = pick_ranges(args);
    bp.range_shift = (uint32_t)(ilog2_u64(args->m) - ilog2_u64((uint64_t)nranges));
    const int base_mode = (bp.blocked ? 1 : 0) | (bp.k == 1u ? 2 : 0);
    bp.nranges = (uint32_t)nranges;
    for (int r = 0; r < nranges; r++) {
        bp.range_id = (uint32_t)r;
        launch_probe_mode(base_mode | (nranges > 1 ? 4 : 0), S->d, S->n, S->n_dev, bp, reinterpret_cast<uint2*>(d_out),
                          reinterpret_cast<unsigned long long*>(d_count_out), nullptr, nullptr);
    }
    CK(cudaGetLastError());
    return 0;
}

int hwbrj_filter_build(const hwbrj_rel_t* R, const bloom_filter_args_t* args, void* d_filter, int zero_first) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    if (!R || !args || !d_filter || check_args_impl(args, true)) return -1;
    g.histR.ensure(((size_t)1 << kMaxRadixBits) * 4);
    // the (unused) one-bin histogram of the insert kernel: histR, unless a prepared R partitioning owns histR right now
    uint32_t* hist = g.histR.as<uint32_t>();
    if (g.prep.valid) {
        g.histF.ensure(256);
        hist = g.histF.as<uint32_t>();
    }
    if (zero_first) CK(cudaMemsetAsync(d_filter, 0, std::max<uint64_t>(args->m / 8, 4), g.stream));
    CK(cudaMemsetAsync(hist, 0, 4, g.stream));
    BloomParams bp = make_bloom(args, 42u, reinterpret_cast<uint32_t*>(d_filter));
    int nranges = pick_ranges(args);
    bp.nranges = (uint32_t)nranges;
    bp.range_shift = (uint32_t)(ilog2_u64(args->m) - ilog2_u64((uint64_t)nranges));
    for (int r = 0; r < nranges; r++) {
        bp.range_id = (uint32_t)r;
        k_build_hist<true><<<g.sms * 2, 1024, (1 + kCrcSmemWords) * 4, g.stream>>>(R->d, R->n, R->n_dev, bp, g.d_crc, hist, 0u);
    }
    CK(cudaGetLastError());
    return 0;
}

int hwbrj_filter_or(void* d_dst, const void* d_src, uint64_t nbytes) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    if (nbytes % 16) return -1;
    k_filter_or<<<g.sms * 8, 256, 0, g.stream>>>(reinterpret_cast<uint4*>(d_dst), reinterpret_cast<const uint4*>(d_src),
                                                 nbytes / 16);
    CK(cudaGetLastError());
    return 0;
}

int64_t hwbrj_filter_probe(const void* d_filter, const hwbrj_rel_t* S, const bloom_filter_args_t* args, void* d_out) {
    std::lock_guard<std::mutex> lock(g.mu);
    init_ctx();
    if (!S || !args || !d_filter || !d_out || check_args_impl(args, true)) return -1;
    g.ctrl.ensure(sizeof(Control));
    CK(cudaMemsetAsync(g.ctrl.p, 0, sizeof(Control), g.stream));
    Control* ctrl = g.ctrl.as<Control>();
    BloomParams bp = make_bloom(args, 42u, reinterpret_cast<uint32_t*>(const_cast<void*>(d_filter)));
    int nranges = pick_ranges(args);
    bp.nranges = (uint32_t)nranges;
    bp.range_shift = (uint32_t)(ilog2_u64(args->m) - ilog2_u64((uint64_t)nranges));
    g.st1.ensure(std::max<uint64_t>(S->n, 1) * 8);
    if (nranges > 2 && g.defer_ranges) g.d1.ensure(std::max<uint64_t>(S->n, 1) * 8);
    run_probe(S->d, S->n, bp, nranges, reinterpret_cast<uint2*>(d_out), ctrl, g.st1.as<uint2>(),
              nranges > 2 ? g.d1.as<uint2>() : nullptr);
    unsigned long long cnt = 0;
    CK(cudaMemcpyAsync(&cnt, &ctrl->survivors, 8, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaGetLastError());
    return (int64_t)cnt;
}

}  // extern "C"
