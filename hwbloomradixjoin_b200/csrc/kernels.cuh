// kernels.cuh -- the sm_100a kernels of the Bloom-filter radix join (K1..K5 of SURVEY.md 2.1).
//
// Data layout in HBM: relations are arrays of 8-byte {int32 key; int32 payload} tuples (types.h:37-40)
// handled as uint2 (x = key, y = payload); the Bloom filter is m/32 uint32 words whose little-endian
// byte/bit order equals the reference's byte array (bit b -> word b>>5, mask 1<<(b&31) == byte b>>3,
// mask 1<<(b&7); bloom_filter.c:84,103).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <type_traits>
#include "hash.cuh"

namespace hwbrj {

constexpr int kMaxLevelBits = 7;                     // radix bits per scatter pass
constexpr int kMaxPeers = 16;                        // GPUs of one NVLink domain
constexpr int kMaxRadixBits = 2 * kMaxLevelBits;     // 2 passes
constexpr int kTableCap = 8192;                      // R tuples per shared-memory hash table
#ifndef HWBRJ_JOIN_THREADS
#define HWBRJ_JOIN_THREADS 512
#endif
#ifndef HWBRJ_JOIN_UNROLL
#define HWBRJ_JOIN_UNROLL 2
#endif
#ifndef HWBRJ_SCATTER_THREADS
#define HWBRJ_SCATTER_THREADS 256
#endif
#ifndef HWBRJ_SCATTER_TILE
#define HWBRJ_SCATTER_TILE 2048
#endif
#ifndef HWBRJ_SCATTER_STAGES
#define HWBRJ_SCATTER_STAGES 2
#endif
#ifndef HWBRJ_SCATTER_MINBLOCKS
#define HWBRJ_SCATTER_MINBLOCKS 4
#endif
constexpr int kJoinThreads = HWBRJ_JOIN_THREADS;
#ifndef HWBRJ_JOIN_SCHUNK
#define HWBRJ_JOIN_SCHUNK 32768
#endif
constexpr int kSChunk = HWBRJ_JOIN_SCHUNK;            // S tuples per join work item
constexpr uint32_t kBigItems = 128;                  // partitions with more work items are expanded cooperatively
constexpr int kScatterThreads = HWBRJ_SCATTER_THREADS;
constexpr int kScatterTile = HWBRJ_SCATTER_TILE;     // tuples per scatter tile
constexpr int kScatterStages = HWBRJ_SCATTER_STAGES; // TMA bulk-load ring depth
constexpr int kScatterStageTuples = kScatterTile + 2; // +1 misaligned head, +1 rounding to 16 bytes
constexpr int kScatterSmem = (kScatterStages * kScatterStageTuples + kScatterTile) * 8;
// K2 shapes (k_probe_compact<MODE, SHAPE>), measured at C1 / C5 (profiles/r2_kernel_experiments.log):
//  SHAPE 0, selective filters: 2 x 128-bit loads = 4 keys in flight per lane, 5 CTAs of 8 warps per SM (40 registers),
//           256-tuple survivor ring per warp (16 KB per CTA: ~150 KB of L1 left for the loads in flight) -- C1 5.17 ms
//           against 5.53 ms for shape 1; but 12.9 against 6.5 ms when every tuple survives (twice the drains and
//           cursor atomics);
//  SHAPE 1, most tuples survive: 8 keys per lane, 4 CTAs per SM, 512-tuple rings.
// A sample of S probed against the finished filter picks the shape on the device (k_probe_sample): both kernels are
// launched and the one that was not chosen returns at once.
#ifndef HWBRJ_PROBE_V
#define HWBRJ_PROBE_V 2
#endif
#ifndef HWBRJ_PROBE_MINBLOCKS
#define HWBRJ_PROBE_MINBLOCKS 5
#endif
#ifndef HWBRJ_PROBE_RING
#define HWBRJ_PROBE_RING 256  // survivor ring per warp (tuples, power of two >= 128)
#endif
template <int SHAPE>
struct ProbeShape {
    static constexpr int V = SHAPE == 0 ? HWBRJ_PROBE_V : 4;              // 128-bit loads in flight per lane
    static constexpr int RING = SHAPE == 0 ? HWBRJ_PROBE_RING : 512;      // survivor ring per warp, tuples
    static constexpr int CTAS = SHAPE == 0 ? HWBRJ_PROBE_MINBLOCKS : 4;   // CTAs per SM
    static constexpr int SMEM_PER_WARP = RING * 8;
};
constexpr int kStagedV = 4;                          // ... and in k_probe_staged (k >= 2), which keeps its measured shape

struct BloomParams {
    uint32_t* filter;      // m/32 words
    uint32_t size_mask;    // (m-1) for BASIC, (B-1) for BLOCKED: modulus of the in-filter arithmetic
    uint32_t nblocks_mask; // m/B - 1 (BLOCKED)
    uint32_t log2B;        // log2(B) (BLOCKED)
    uint32_t k;
    uint32_t seed;
    uint32_t blocked;
    // range passes: a key is handled in the pass whose id equals (first bit address >> range_shift)
    uint32_t nranges;
    uint32_t range_shift;
    uint32_t range_id;
    const uint32_t* skew;  // K2: verdict of k_probe_sample (nullptr = no sample: uniform keys, dense survivors assumed)
};

struct JoinAccum {
    unsigned long long matches, cpair, crpay, cspay, ckey;
};

__device__ __forceinline__ uint64_t mix64(uint32_t rpay, uint32_t spay) {
    uint64_t z = ((uint64_t)rpay << 32) | (uint64_t)spay;
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

// streaming 128-bit load: read-once data must not displace the filter in L2. In the micro-benchmark ld.global.cs
// + evict_last probes looked 6% better, in the real K2 it is 10% worse than this evict_first policy + plain __ldg
// probes (5.9 vs 5.3 ms at C1), so this is what ships.
__device__ __forceinline__ uint4 ld_stream_v4(const uint4* p, uint64_t pol) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ uint2 ld_stream_v2(const uint2* p, uint64_t pol) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;"
                 : "=r"(r.x), "=r"(r.y)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_stream_v2(uint2* p, uint2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.u32 [%0], {%1,%2}, %3;" ::"l"(p), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}


// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// global -> shared bulk copy; src and dst 16-byte aligned, bytes a multiple of 16; completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- Bloom index sequence (bloom_filter.c:74-111,126-141) ---------------------------------
// first bit address, and the (h, y) state of the enhanced double hashing inside the (sub)filter
__device__ __forceinline__ void bloom_start(const BloomParams& bp, const uint32_t* crc_tab, uint32_t key,
                                            uint32_t& base, uint32_t& h, uint32_t& y) {
    h = hash_crapwow(bp.seed, key) & bp.size_mask;
    y = (key + bp.seed) & bp.size_mask;
    base = bp.blocked ? ((crc32c_tab(crc_tab, bp.seed, key) & bp.nblocks_mask) << bp.log2B) : 0u;
}
__device__ __forceinline__ bool bloom_in_range(const BloomParams& bp, uint32_t addr0) {
    return bp.nranges == 1u || (addr0 >> bp.range_shift) == bp.range_id;
}

__device__ __forceinline__ void bloom_insert(const BloomParams& bp, uint32_t base, uint32_t h, uint32_t y) {
    for (uint32_t i = 0; i < bp.k; i++) {
        uint32_t a = base + h;
        atomicOr(bp.filter + (a >> 5), 1u << (a & 31u));  // RED.OR: relaxed, idempotent, order-free
        h = (h + y) & bp.size_mask;
        y = (y + i + 1u) & bp.size_mask;
    }
}

// filter probe load. BASIC: ld.global.cg (L2 only) -- measured in K2 at C1: __ldg 5.2 ms, L1::no_allocate 9.5 ms,
// .cg 4.96 ms. BLOCKED keeps the L1-allocating __ldg: probes 2..k of a key fall into the block (often the very sector)
// that its first probe has just brought into L1. HWBRJ_PROBE_LD forces one flavour for experiments (0 = __ldg,
// 1 = nc.L1::no_allocate, 2 = .cg, 3 = by filter variant).
// (K2 also slows down 3x when the L1 carve-out is minimal -- the in-flight loads of any flavour need L1 data space.)
#ifndef HWBRJ_PROBE_LD
#define HWBRJ_PROBE_LD 3
#endif
// `hot`: the probe relation is skewed (k_probe_sample found repeated keys in a sample of S, e.g. Zipf): then the
// L1-allocating load is used for BASIC filters as well, so that the few filter lines every SM keeps asking for are served by
// its own L1 instead of by one L2 slice (C5, theta = 1: K2 10.4 -> 6.6 ms). It is a compile-time constant at every call
// site: K2 holds two copies of its loop and picks one per launch. (Selecting the flavour per load inside one loop cost
// the uniform case 1.2 ms of 5.5 at C1: profiles/r2_kernel_experiments.log.)
__device__ __forceinline__ uint32_t ld_filter(const BloomParams& bp, const uint32_t* p, bool hot = false) {
#if HWBRJ_PROBE_LD == 1
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#elif HWBRJ_PROBE_LD == 2 || HWBRJ_PROBE_LD == 3
    if (HWBRJ_PROBE_LD == 3 && (bp.blocked || hot)) return __ldg(p);  // bp.blocked: compile-time constant inside K2
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

// probes 1..k-1 (the first word is passed in so that callers can batch the first probes)
__device__ __forceinline__ bool bloom_test_rest(const BloomParams& bp, uint32_t base, uint32_t h, uint32_t y,
                                                uint32_t w0) {
    uint32_t a = base + h;
    if (!((w0 >> (a & 31u)) & 1u)) return false;
    for (uint32_t i = 1; i < bp.k; i++) {
        h = (h + y) & bp.size_mask;
        y = (y + i) & bp.size_mask;
        a = base + h;
        if (!((ld_filter(bp, bp.filter + (a >> 5)) >> (a & 31u)) & 1u)) return false;
    }
    return true;
}

__device__ __forceinline__ void load_crc_tab(uint32_t* s_tab, const uint32_t* __restrict__ g_tab) {
    for (int i = threadIdx.x; i < kCrcSmemWords; i += blockDim.x) s_tab[i] = g_tab[i];
}

// ---- K0: hash library entry ---------------------------------------------------------------------------------
__global__ void k_hash_many(int which, uint32_t seed, const int32_t* __restrict__ keys, uint64_t n,
                            uint32_t* __restrict__ out) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = hash_dispatch(which, seed, (uint32_t)keys[i]);
}

// ---- partition function -------------------------------------------------------------------------------------------
// A tuple's partition id pid in [0, 2^bits) decides everything downstream: level-1 scatter bin = pid >> b2, level-2 bin =
// pid & (2^b2 - 1), owner GPU = pid >> (bits - log2 G), table index inside the partition = the key bits that are left.
// PMODE 0 (radix): pid = key & (2^bits - 1)                    -- HASH_BIT_MODULO, parallel_radix_join_bloom.c:74
// PMODE 1 (hash, BASIC filter): pid = (crapwow(seed,key) & (m-1)) >> (log2 m - bits), i.e. the index of the m/2^bits-bit
//         SLICE of the filter that holds the key's first bit: partition p owns bits [p*m/2^bits, (p+1)*m/2^bits)
// PMODE 2 (hash, BLOCKED filter): pid = (crc32c(seed,key) & (m/B-1)) >> (log2(m/B) - bits), the slice holding the key's block
struct PartFn {
    uint32_t bits, b2;
    uint32_t seed, size_mask, hshift;  // hash modes; size_mask == 0 (bits == 0) maps everything to partition 0
};
template <int PMODE>
__device__ __forceinline__ uint32_t pid_of(const PartFn& f, const uint32_t* crc_tab, uint32_t key) {
    if (PMODE == 0) return key & ((1u << f.bits) - 1u);
    if (PMODE == 1) return (hash_crapwow(f.seed, key) & f.size_mask) >> f.hshift;
    return (crc32c_tab(crc_tab, f.seed, key) & f.size_mask) >> f.hshift;
}

// ---- K1 (+K3 histogram): Bloom insert fused with the partition histogram ---------------------------------------
// replaces the build branch of the histogram loop, parallel_radix_join_bloom.c:794-805 + add_generic
// (bloom_filter.c:74-89). Also used without a filter (plain PRO histogram, parallel_radix_join.c:770-775) and for the
// survivors of the probe. dynamic smem: hist[2^bits] then crc table[kCrcSmemWords]
template <bool BLOOM, int PMODE>
__global__ void __launch_bounds__(1024, 2) k_build_hist(const uint2* __restrict__ rel, uint64_t n_static,
                                                    const unsigned long long* __restrict__ n_ptr, BloomParams bp,
                                                    const uint32_t* __restrict__ g_crc, uint32_t* __restrict__ ghist,
                                                    PartFn pf) {
    extern __shared__ uint32_t smem[];
    const uint64_t n = n_ptr ? min((uint64_t)*n_ptr, n_static) : n_static;  // n_static bounds a device-side count
    const uint32_t P = 1u << pf.bits;
    uint32_t* hist = smem;
    uint32_t* crc_tab = smem + P;
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) hist[i] = 0u;
    if ((BLOOM && bp.blocked) || PMODE == 2) load_crc_tab(crc_tab, g_crc);
    __syncthreads();
    const uint64_t pol = policy_evict_first();
    const uint64_t npairs = n >> 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint4* rel4 = reinterpret_cast<const uint4*>(rel);
    auto one = [&](uint32_t key) {
        if (BLOOM) {
            uint32_t base, h, y;
            bloom_start(bp, crc_tab, key, base, h, y);
            if (!bloom_in_range(bp, base + h)) return;
            bloom_insert(bp, base, h, y);
        }
        atomicAdd(&hist[pid_of<PMODE>(pf, crc_tab, key)], 1u);
    };
    constexpr int U = 4;  // 128-bit loads in flight per thread
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += stride * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t idx = i + (uint64_t)u * stride;
            v[u] = idx < npairs ? ld_stream_v4(rel4 + idx, pol) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (i + (uint64_t)u * stride < npairs) {
                one(v[u].x);
                one(v[u].z);
            }
        }
    }
    if ((n & 1ull) && blockIdx.x == 0 && threadIdx.x == 0) one(rel[n - 1].x);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
        uint32_t c = hist[i];
        if (c) atomicAdd(&ghist[i], c);
    }
}

// ---- sample of the probe relation for K2 ---------------------------------------------------------------------------------
// One CTA looks at 4096 keys spread evenly over S, after the filter is complete, and answers two questions:
//  bit 0: do keys repeat? (open-addressing set in shared memory; foreign keys drawn from a large domain repeat almost
//         never, Zipf-distributed ones or a small key domain all the time; set when > 1/64 of the sample are repeats)
//         -> K2 probes with L1-allocating loads (ld_filter);
//  bit 1: do most tuples pass the filter? (the sample is probed exactly as K2 probes; set when > 35 % pass)
//         -> K2 runs in its shape for dense survivors (ProbeShape<1>).
// A heuristic that only selects a load flavour and a launch shape -- results never depend on it. 0.013 ms.
constexpr int kSkewSample = 4096, kSkewSlots = 8192;
__global__ void __launch_bounds__(1024) k_probe_sample(const uint2* __restrict__ S, uint64_t n, BloomParams bp,
                                                      const uint32_t* __restrict__ g_crc, uint32_t* __restrict__ flag) {
    __shared__ uint32_t slot[kSkewSlots];
    __shared__ uint32_t crc_tab[kCrcSmemWords];
    __shared__ uint32_t repeats, passes;
    for (int i = threadIdx.x; i < kSkewSlots; i += 1024) slot[i] = 0xFFFFFFFFu;  // (a key equal to the marker counts once less)
    if (bp.filter != nullptr && bp.blocked) load_crc_tab(crc_tab, g_crc);
    if (threadIdx.x == 0) repeats = passes = 0u;
    __syncthreads();
    const uint64_t nsamp = n < (uint64_t)kSkewSample ? n : (uint64_t)kSkewSample;
    const uint64_t stride = nsamp ? n / nsamp : 1ull;
    uint32_t mine = 0u, pass = 0u;
    for (uint64_t i = threadIdx.x; i < nsamp; i += 1024u) {
        const uint32_t key = S[i * stride].x;
        uint32_t h = (key * 0x9E3779B1u) >> 19;  // 13 bits
        for (int probe = 0; probe < kSkewSlots; probe++) {
            const uint32_t old = atomicCAS(&slot[h], 0xFFFFFFFFu, key);
            if (old == 0xFFFFFFFFu) break;
            if (old == key) {
                mine++;
                break;
            }
            h = (h + 1u) & (kSkewSlots - 1);
        }
        if (bp.filter != nullptr) {  // contains_generic (bloom_filter.c:93-111) on the sampled key
            uint32_t b0, h0, y0;
            bloom_start(bp, crc_tab, key, b0, h0, y0);
            const uint32_t a = b0 + h0;
            if (bp.k == 0u || bloom_test_rest(bp, b0, h0, y0, __ldg(bp.filter + (a >> 5)))) pass++;
        }
    }
    if (mine) atomicAdd(&repeats, mine);
    if (pass) atomicAdd(&passes, pass);
    __syncthreads();
    if (threadIdx.x == 0)
        *flag = ((nsamp >= 64u && repeats * 64u > (uint32_t)nsamp) ? 1u : 0u) | ((passes * 100u > (uint32_t)nsamp * 35u) ? 2u : 0u);
}

// ---- K2: Bloom probe + ballot/prefix compaction of survivors ----------------------------------------------------
// replaces the probe branch of the histogram loop (:794-805), contains_generic (bloom_filter.c:93-111), the
// contains_cache bitmap (:788,:801,:843) and the `filtered` sum (:1188-1193).
// Warp-autonomous: no block barrier in the streaming loop, so loads of different warps stay in flight while
// others compact. Each warp streams 32 x kProbeV 128-bit loads (2 tuples each), issues all first probes
// together, and appends survivors to its private shared-memory ring; whenever the ring holds kWarpFlush
// tuples the warp claims kWarpFlush slots of the output with ONE global atomic and writes them as whole
// 128-byte lines (claims are multiples of kWarpFlush, so every flush but the last is line-aligned).
#ifndef HWBRJ_PROBE_WARPS
#define HWBRJ_PROBE_WARPS 8
#endif
constexpr int kProbeWarps = HWBRJ_PROBE_WARPS;  // warps per CTA (4 and 2 are 10 % slower, 16 the same: r2_kernel_experiments.log)

// per-warp shared-memory ring: tuples are appended with ballot/popc ranks and drained in line-aligned pieces.
// (Round-2 measurements, profiles/r2_k2_ablation_and_tuning.log: a private output region per warp instead of the shared
// cursor changes nothing (5.07 vs 5.00 ms), claiming the output space one iteration ahead is slower (5.18 ms), smaller
// rings are slower (6.2 ms at 128 tuples) -- the cursor atomic is not what K2 waits for.)
template <int CAP>  // power of two; drained CAP/2 tuples at a time
struct WarpRing {
    uint2* buf;
    uint32_t head, count;  // warp-uniform
    __device__ __forceinline__ void init(uint2* b) { buf = b; head = 0u; count = 0u; }
    __device__ __forceinline__ void append2(bool fa, uint2 a, bool fb, uint2 b, uint32_t lt) {
        const uint32_t ma = __ballot_sync(0xffffffffu, fa);
        const uint32_t mb = __ballot_sync(0xffffffffu, fb);
        const uint32_t tail = head + count;
        if (fa) buf[(tail + __popc(ma & lt)) & (CAP - 1)] = a;
        if (fb) buf[(tail + __popc(ma) + __popc(mb & lt)) & (CAP - 1)] = b;
        count += __popc(ma) + __popc(mb);
    }
    __device__ __forceinline__ void append1(bool f, uint2 a, uint32_t lt) {
        const uint32_t m = __ballot_sync(0xffffffffu, f);
        if (f) buf[(head + count + __popc(m & lt)) & (CAP - 1)] = a;
        count += __popc(m);
    }
    __device__ __forceinline__ void drain(uint32_t cnt, uint2* __restrict__ out, unsigned long long* cursor, uint64_t pol,
                                          uint32_t lane) {
        __syncwarp();
        unsigned long long gb = 0ull;
        if (lane == 0) gb = atomicAdd(cursor, (unsigned long long)cnt);
        gb = __shfl_sync(0xffffffffu, gb, 0);
        for (uint32_t i = lane; i < cnt; i += 32u) st_stream_v2(out + gb + i, buf[(head + i) & (CAP - 1)], pol);
        head = (head + cnt) & (CAP - 1);
        count -= cnt;
        __syncwarp();
    }
    // at most 64 tuples are appended between two calls, so one drain of CAP/2 keeps the ring from overflowing
    __device__ __forceinline__ void drain_if_full(uint2* __restrict__ out, unsigned long long* cursor, uint64_t pol, uint32_t lane) {
        if (count >= (uint32_t)(CAP / 2)) drain(CAP / 2, out, cursor, pol, lane);
    }
};

// MODE bit0: BLOCKED, bit1: single probe (k == 1), bit2: range passes active -- compile-time specialisation keeps the
// per-tuple instruction count down (the kernel is L1TEX-wavefront-bound: one divergent access per SM per clock).
// Range passes: pass i probes the keys whose first filter bit lies in range i, so that the probed part of the filter
// stays L2-resident. (Deferring the keys of later ranges to a buffer instead of re-reading S was measured slower in
// round 1 -- the deferred writes evict the probed range -- and has been removed.)
constexpr int kStagedCandCap = 256, kStagedSurvCap = 256;  // k_probe_staged: candidate ring + survivor ring per warp
constexpr int kStagedSmemPerWarp = (kStagedCandCap + kStagedSurvCap) * 8;
template <int MODE, int SHAPE>
__global__ void __launch_bounds__(kProbeWarps * 32, ProbeShape<SHAPE>::CTAS)
k_probe_compact(const uint2* __restrict__ S, uint64_t n_static, const unsigned long long* __restrict__ n_ptr,
                BloomParams bp_in, const uint32_t* __restrict__ g_crc, uint2* __restrict__ out,
                unsigned long long* __restrict__ out_cursor) {
    constexpr bool kBlocked = (MODE & 1) != 0, kSingle = (MODE & 2) != 0, kRanged = (MODE & 4) != 0;
    constexpr int kProbeV = ProbeShape<SHAPE>::V, kRing = ProbeShape<SHAPE>::RING;
    static_assert(kRing >= 128, "append2 adds up to 64 tuples between two drain checks");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint32_t crc_tab[kBlocked ? kCrcSmemWords : 1];
    // k_probe_sample's verdict: bit 0 = keys repeat (probe-load flavour), bit 1 = most tuples survive (shape 1). Without a
    // sample (bp.skew == nullptr) only shape 1 is launched.
    const uint32_t verdict = bp_in.skew != nullptr ? *bp_in.skew : 2u;
    if (((verdict >> 1) & 1u) != (uint32_t)SHAPE) return;
    BloomParams bp = bp_in;
    bp.blocked = kBlocked ? 1u : 0u;
    if (kSingle) bp.k = 1u;
    if (!kRanged) bp.nranges = 1u;
    if (kBlocked) {
        load_crc_tab(crc_tab, g_crc);
        __syncthreads();
    }
    const uint64_t n = n_ptr ? min((uint64_t)*n_ptr, n_static) : n_static;  // n_static bounds a device-side count
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    WarpRing<kRing> surv;
    surv.init(reinterpret_cast<uint2*>(smem_raw) + wid * kRing);
    const uint64_t pol = policy_evict_first();
    const uint64_t npairs = n >> 1;
    const uint4* S4 = reinterpret_cast<const uint4*>(S);
    const uint64_t warp_global = (uint64_t)blockIdx.x * kProbeWarps + wid;
    const uint64_t nwarps = (uint64_t)gridDim.x * kProbeWarps;
    constexpr uint64_t kPerIter = 32ull * kProbeV;  // pairs per warp iteration

    auto stream_loop = [&](auto hot_tag) {
    constexpr bool hot = decltype(hot_tag)::value;
    for (uint64_t it = warp_global; it * kPerIter < npairs; it += nwarps) {
        const uint64_t p0 = it * kPerIter + lane;
        uint4 t[kProbeV];
#pragma unroll
        for (int j = 0; j < kProbeV; j++) {
            const uint64_t idx = p0 + (uint64_t)j * 32u;
            t[j] = (idx < npairs) ? ld_stream_v4(S4 + idx, pol) : make_uint4(0u, 0u, 0u, 0u);
        }
        uint32_t base[2 * kProbeV], h[2 * kProbeV], y[2 * kProbeV], w[2 * kProbeV];
        bool act[2 * kProbeV];
#pragma unroll
        for (int j = 0; j < kProbeV; j++) {
            const bool valid = (p0 + (uint64_t)j * 32u) < npairs;
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int q = 2 * j + e;
                bloom_start(bp, crc_tab, e ? t[j].z : t[j].x, base[q], h[q], y[q]);
                const uint32_t a = base[q] + h[q];
                act[q] = valid && bloom_in_range(bp, a);
                w[q] = act[q] ? ld_filter(bp, bp.filter + (a >> 5), hot) : 0u;  // all first probes in flight together
            }
        }
#pragma unroll
        for (int j = 0; j < kProbeV; j++) {
            const bool fa = act[2 * j] && (bp.k == 0u || bloom_test_rest(bp, base[2 * j], h[2 * j], y[2 * j], w[2 * j]));
            const bool fb = act[2 * j + 1] &&
                            (bp.k == 0u || bloom_test_rest(bp, base[2 * j + 1], h[2 * j + 1], y[2 * j + 1], w[2 * j + 1]));
            surv.append2(fa, make_uint2(t[j].x, t[j].y), fb, make_uint2(t[j].z, t[j].w), lt);
            surv.drain_if_full(out, out_cursor, pol, lane);
        }
    }
    };
    // two copies of the loop, one per probe-load flavour; the choice is uniform over the launch (see ld_filter)
    if (!kBlocked && (verdict & 1u)) stream_loop(std::true_type{});
    else stream_loop(std::false_type{});
    if (surv.count) surv.drain(surv.count, out, out_cursor, pol, lane);
    // odd tail tuple
    if ((n & 1ull) && blockIdx.x == 0 && threadIdx.x == 0) {
        const uint2 tt = S[n - 1];
        uint32_t b0, h0, y0;
        bloom_start(bp, crc_tab, tt.x, b0, h0, y0);
        const uint32_t a = b0 + h0;
        if (bloom_in_range(bp, a) &&
            (bp.k == 0u || bloom_test_rest(bp, b0, h0, y0, ld_filter(bp, bp.filter + (a >> 5))))) {
            const unsigned long long pos = atomicAdd(out_cursor, 1ull);
            out[pos] = tt;
        }
    }
}

// ---- K2 for k >= 2, staged (default for k >= 2; HWBRJ_PROBE_STAGED=0 selects k_probe_compact; both parity-checked) ----
// With several probes per key the plain kernel walks probes 2..k under divergence: after the first probe 38 % of the
// lanes are still alive at C1-blocked (k = 4), then 14 %, then 5 %, but the warp pays for every round. Here the keys that
// pass their FIRST probe are compacted into a per-warp candidate ring (the tuple only), and probes 2..k run on full
// batches of 32 candidates, one per lane, re-deriving the index sequence from the key; a batch stops as soon as no
// lane is alive. Survivors go to a second, smaller ring (few tuples survive k probes).
template <int MODE>  // bit0 BLOCKED, bit2 ranged
__global__ void __launch_bounds__(kProbeWarps * 32) k_probe_staged(const uint2* __restrict__ S, uint64_t n_static,
                                                                  const unsigned long long* __restrict__ n_ptr,
                                                                  BloomParams bp_in, const uint32_t* __restrict__ g_crc,
                                                                  uint2* __restrict__ out,
                                                                  unsigned long long* __restrict__ out_cursor) {
    constexpr bool kBlocked = (MODE & 1) != 0, kRanged = (MODE & 4) != 0;
    constexpr int kCandCap = kStagedCandCap, kSurvCap = kStagedSurvCap;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint32_t crc_tab[kBlocked ? kCrcSmemWords : 1];
    BloomParams bp = bp_in;
    bp.blocked = kBlocked ? 1u : 0u;
    if (!kRanged) bp.nranges = 1u;
    if (kBlocked) {
        load_crc_tab(crc_tab, g_crc);
        __syncthreads();
    }
    const uint64_t n = n_ptr ? min((uint64_t)*n_ptr, n_static) : n_static;
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    uint2* wsm = reinterpret_cast<uint2*>(smem_raw) + wid * (kCandCap + kSurvCap);
    WarpRing<kCandCap> cand;
    cand.init(wsm);
    WarpRing<kSurvCap> surv;
    surv.init(wsm + kCandCap);
    const uint64_t pol = policy_evict_first();
    const uint64_t npairs = n >> 1;
    const uint4* S4 = reinterpret_cast<const uint4*>(S);
    const uint64_t warp_global = (uint64_t)blockIdx.x * kProbeWarps + wid;
    const uint64_t nwarps = (uint64_t)gridDim.x * kProbeWarps;
    constexpr uint64_t kPerIter = 32ull * kStagedV;

    // probes 2..k of up to 32 candidates (one per lane); the survivors move to the output ring
    auto finish_batch = [&](uint32_t nb) {
        __syncwarp();  // the candidates were stored by other lanes
        const uint2 tt = cand.buf[(cand.head + lane) & (kCandCap - 1)];
        cand.head = (cand.head + nb) & (kCandCap - 1);
        cand.count -= nb;
        uint32_t base, h, y;
        bloom_start(bp, crc_tab, tt.x, base, h, y);
        bool alive = lane < nb;
        for (uint32_t i = 1; i < bp.k; i++) {
            if (!__any_sync(0xffffffffu, alive)) break;
            h = (h + y) & bp.size_mask;
            y = (y + i) & bp.size_mask;
            const uint32_t a = base + h;
            if (alive) alive = ((ld_filter(bp, bp.filter + (a >> 5)) >> (a & 31u)) & 1u) != 0u;
        }
        surv.append1(alive, tt, lt);
        surv.drain_if_full(out, out_cursor, pol, lane);
        __syncwarp();  // all lanes have read their candidate before the ring is appended to again
    };

    for (uint64_t it = warp_global; it * kPerIter < npairs; it += nwarps) {
        const uint64_t p0 = it * kPerIter + lane;
        uint4 t[kStagedV];
#pragma unroll
        for (int j = 0; j < kStagedV; j++) {
            const uint64_t idx = p0 + (uint64_t)j * 32u;
            t[j] = (idx < npairs) ? ld_stream_v4(S4 + idx, pol) : make_uint4(0u, 0u, 0u, 0u);
        }
        uint32_t a0[2 * kStagedV], w[2 * kStagedV];
        bool act[2 * kStagedV];
#pragma unroll
        for (int j = 0; j < kStagedV; j++) {
            const bool valid = (p0 + (uint64_t)j * 32u) < npairs;
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int q = 2 * j + e;
                uint32_t base, h, y;
                bloom_start(bp, crc_tab, e ? t[j].z : t[j].x, base, h, y);
                a0[q] = base + h;
                act[q] = valid && bloom_in_range(bp, a0[q]);
                w[q] = act[q] ? ld_filter(bp, bp.filter + (a0[q] >> 5)) : 0u;  // all first probes in flight together
            }
        }
#pragma unroll
        for (int j = 0; j < kStagedV; j++) {
            const bool fa = act[2 * j] && ((w[2 * j] >> (a0[2 * j] & 31u)) & 1u);
            const bool fb = act[2 * j + 1] && ((w[2 * j + 1] >> (a0[2 * j + 1] & 31u)) & 1u);
            cand.append2(fa, make_uint2(t[j].x, t[j].y), fb, make_uint2(t[j].z, t[j].w), lt);
            while (cand.count >= 32u) finish_batch(32u);  // leaves < 32, at most 64 are appended per step: never full
        }
    }
    if (cand.count) finish_batch(cand.count);
    if (surv.count) surv.drain(surv.count, out, out_cursor, pol, lane);
    if ((n & 1ull) && blockIdx.x == 0 && threadIdx.x == 0) {  // odd tail tuple
        const uint2 tt = S[n - 1];
        uint32_t b0, h0, y0;
        bloom_start(bp, crc_tab, tt.x, b0, h0, y0);
        const uint32_t a = b0 + h0;
        if (bloom_in_range(bp, a) && bloom_test_rest(bp, b0, h0, y0, ld_filter(bp, bp.filter + (a >> 5)))) {
            const unsigned long long pos = atomicAdd(out_cursor, 1ull);
            out[pos] = tt;
        }
    }
}

// ---- K3: offsets from the partition histograms of all ranks ---------------------------------------------------------
// replaces the local prefix (:808-811) and the cross-thread offset computation (:819-837); the "threads" of the reference
// are the G ranks here. One CTA of 1024 threads; P <= 2^14 partitions, of which rank g owns the contiguous range
// [g*P/G, (g+1)*P/G). hist_all[src][p] = number of tuples of partition p in rank src's input (every rank holds all G
// rows: they are pushed over NVLink by k_push_rows; G == 1: the row is the local histogram). From these the kernel derives
//   * cursor1[B] for every GLOBAL level-1 bin B: where this rank's level-1 scatter puts bin B in its OWN staging buffer
//     (an exclusive scan of its own row: the level-1 pass is purely local);
//   * the level-2 input segments of the OWNED level-1 bins: segment (lb, src) = the tuples of owned bin lb that sit in
//     rank src's staging buffer, seg_start[lb*G+src] tuples into it (src's cursor1, which every rank can compute).
//     The level-2 pass PULLS these segments over NVLink with bulk loads, so every tuple crosses the fabric exactly once,
//     in large sequential reads, and nothing is ever written into a peer's partitions;
//   * the tile schedule of that pass (tile_off[P1+1] over the P1 = P1L*G segments), the fine boundaries fine_off[PL+1] of
//     the owned partitions (in increasing pid order -- the reference's cluster order, :1896-1939) and the level-2 cursors;
//   * n_own = tuples this rank owns; *abort = 1 if a rank's chunk or an owner's share exceeds `capacity`: then every
//     partition is declared empty and the scatter kernels do nothing (the join reports the failure).
constexpr int kBinsPerThread = (1 << kMaxRadixBits) / 1024;

// exclusive block scan of one value per thread (1024 threads); returns the exclusive prefix, total via warp_sums[32]
__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t local, uint32_t* warp_sums /* [33] */) {
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint32_t inc = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (uint32_t)d) inc += v;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = warp_sums[lane];
        uint32_t winc = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= (uint32_t)d) winc += v;
        }
        warp_sums[lane] = winc - ws;
        if (lane == 31) warp_sums[32] = winc;
    }
    __syncthreads();
    return warp_sums[wid] + inc - local;
}

__global__ void __launch_bounds__(1024) k_scan_dist(const uint32_t* __restrict__ hist_all, uint32_t G, uint32_t rank,
                                                   uint32_t P, uint32_t b2, unsigned long long capacity,
                                                   uint32_t* __restrict__ fine_off, uint32_t* __restrict__ cursor1,
                                                   uint32_t* __restrict__ cursor2, uint32_t* __restrict__ tile_off,
                                                   uint32_t* __restrict__ seg_start, uint32_t* __restrict__ seg_cnt,
                                                   unsigned long long* __restrict__ n_own, uint32_t* __restrict__ abort) {
    constexpr uint32_t NB = 1u << kMaxLevelBits;
    __shared__ uint32_t warp_sums[33];
    __shared__ unsigned long long s_bucket[kMaxPeers][NB];  // tuples of level-1 bin B in rank src's chunk
    __shared__ unsigned long long s_lstart[kMaxPeers][NB];  // start of bin B in src's staging buffer (src-local scan)
    __shared__ uint32_t s_abort;
    const uint32_t P1 = P >> b2, PL = P / G, P1L = P1 / G;
    for (uint32_t i = threadIdx.x; i < G * NB; i += 1024u) (&s_bucket[0][0])[i] = 0ull;
    if (threadIdx.x == 0) s_abort = *abort;  // the other relation of this join may already have overflowed
    __syncthreads();
    const uint32_t per = (P + 1023u) / 1024u;  // a power of two <= 2^b2 whenever P > 1024, else 1
    const uint32_t lo = threadIdx.x * per;
    uint32_t v[kBinsPerThread];
#pragma unroll
    for (int i = 0; i < kBinsPerThread; i++) v[i] = 0u;
    const bool mine = lo < P && lo / PL == rank;
    if (lo < P) {
        // the rows of two ranks per step, all loads of a step independent (L2 loads: the rows were written by peers)
        for (uint32_t src = 0; src < G; src += 2u) {
            const bool two = src + 1u < G;
            const uint32_t* row0 = hist_all + (size_t)src * P + lo;
            const uint32_t* row1 = row0 + (two ? P : 0u);
            uint32_t c0[kBinsPerThread], c1[kBinsPerThread];
            if (per == (uint32_t)kBinsPerThread) {  // 16 bins per thread: four 128-bit loads per row
#pragma unroll
                for (int q = 0; q < kBinsPerThread / 4; q++) {
                    const uint4 a = __ldcg(reinterpret_cast<const uint4*>(row0) + q);
                    const uint4 b = __ldcg(reinterpret_cast<const uint4*>(row1) + q);
                    c0[4 * q] = a.x; c0[4 * q + 1] = a.y; c0[4 * q + 2] = a.z; c0[4 * q + 3] = a.w;
                    c1[4 * q] = b.x; c1[4 * q + 1] = b.y; c1[4 * q + 2] = b.z; c1[4 * q + 3] = b.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < kBinsPerThread; i++) {
                    c0[i] = (uint32_t)i < per ? __ldcg(row0 + i) : 0u;
                    c1[i] = (uint32_t)i < per ? __ldcg(row1 + i) : 0u;
                }
            }
            uint32_t sum0 = 0, sum1 = 0;
#pragma unroll
            for (int i = 0; i < kBinsPerThread; i++) {
                sum0 += c0[i];
                sum1 += two ? c1[i] : 0u;
                if (mine) v[i] += c0[i] + (two ? c1[i] : 0u);
            }
            if (sum0) atomicAdd(&s_bucket[src][lo >> b2], (unsigned long long)sum0);
            if (sum1) atomicAdd(&s_bucket[src + 1u][lo >> b2], (unsigned long long)sum1);
        }
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    if (wid < G) {  // warp w: exclusive scan of rank w's row over the <= 128 level-1 bins (4 per lane)
        unsigned long long c[4], sum = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t B = lane * 4u + q;
            c[q] = B < P1 ? s_bucket[wid][B] : 0ull;
            sum += c[q];
        }
        unsigned long long inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (uint32_t)d) inc += u;
        }
        unsigned long long run = inc - sum;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t B = lane * 4u + q;
            if (B < P1) s_lstart[wid][B] = run;
            run += c[q];
        }
        if (lane == 31 && inc > capacity) s_abort = 1u;  // its chunk does not fit its staging buffer
    } else if (wid >= 16 && wid < 16 + G) {  // warp 16+o: what owner o will own
        const uint32_t o = wid - 16u;
        unsigned long long own = 0;
        for (uint32_t e = lane; e < P1L * G; e += 32u) own += s_bucket[e % G][o * P1L + e / G];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) own += __shfl_down_sync(0xffffffffu, own, d);
        if (lane == 0) {
            if (own > capacity) s_abort = 1u;
            if (o == rank) *n_own = own;
        }
    }
    __syncthreads();
    // every rank sees the same rows, so all ranks agree on an overflow
    const bool dead = s_abort != 0u;
    if (dead && threadIdx.x == 0) {
        *abort = 1u;
        *n_own = 0ull;
    }
    if (threadIdx.x < P1) cursor1[threadIdx.x] = (uint32_t)s_lstart[rank][threadIdx.x];
    // fine offsets of the owned partitions (contiguous thread range: other threads contribute 0)
    uint32_t local = 0;
#pragma unroll
    for (int i = 0; i < kBinsPerThread; i++) {
        if (dead) v[i] = 0u;
        local += v[i];
    }
    uint32_t run = block_exclusive_scan_1024(local, warp_sums);
    if (mine) {
        const uint32_t pl = lo - rank * PL;
#pragma unroll
        for (int i = 0; i < kBinsPerThread; i++) {
            if ((uint32_t)i < per && pl + i < PL) {
                fine_off[pl + i] = run;
                cursor2[pl + i] = run;
                run += v[i];
            }
        }
    }
    if (threadIdx.x == 0) fine_off[PL] = warp_sums[32];
    __syncthreads();
    // level-2 input segments (owned bin lb, source rank src) and their tile schedule, scanned by the same block scan
    uint32_t tiles = 0;
    if (threadIdx.x < P1) {
        const uint32_t lb = threadIdx.x / G, src = threadIdx.x % G, B = rank * P1L + lb;
        const uint32_t cnt = dead ? 0u : (uint32_t)s_bucket[src][B];
        seg_start[threadIdx.x] = (uint32_t)s_lstart[src][B];
        seg_cnt[threadIdx.x] = cnt;
        tiles = (cnt + kScatterTile - 1) / kScatterTile;
    }
    const uint32_t tex = block_exclusive_scan_1024(tiles, warp_sums);
    if (threadIdx.x <= P1) tile_off[threadIdx.x] = tex;  // thread P1 has tiles == 0: its prefix is the grand total
}

// ---- small peer-memory collectives (NVLink / NVSwitch, no host involvement) --------------------------------------------
// Every rank maps the symmetric control blocks of its peers (CUDA IPC or peer access). All three kernels are stream
// ordered like any other kernel and can be captured in a CUDA graph.
struct PeerPtrs {
    void* p[kMaxPeers];
};
constexpr uint32_t kFlagStride = 32;  // one 128-byte line per flag

// all-gather of one row per rank: dst row `rank` on every peer <- src (words); used for the partition histograms and the
// result words
__global__ void k_push_rows(PeerPtrs dst, uint32_t world, uint32_t rank, const uint32_t* __restrict__ src, uint32_t words) {
    for (uint32_t g = blockIdx.y; g < world; g += gridDim.y) {
        uint32_t* d = reinterpret_cast<uint32_t*>(dst.p[g]) + (size_t)rank * words;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) d[i] = src[i];
    }
}

// barrier across the ranks: rank r publishes the next epoch in slot r of every peer's flag array (release, system scope;
// everything earlier kernels of this stream stored -- locally or into peer memory -- is performed before it) and waits
// until all slots of its own array have reached that epoch. The epoch lives on the device, so a replayed CUDA graph keeps
// counting. A peer that never arrives trips the time-out instead of hanging the GPU.
__global__ void k_barrier(PeerPtrs flags, uint32_t world, uint32_t rank, uint32_t* __restrict__ epoch,
                          uint32_t* __restrict__ err, long long timeout_cycles) {
    const uint32_t t = threadIdx.x;
    const uint32_t e = *epoch + 1u;
    __syncwarp();
    if (t < world) {
        __threadfence_system();
        uint32_t* remote = reinterpret_cast<uint32_t*>(flags.p[t]) + rank * kFlagStride;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(e) : "memory");
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(flags.p[rank]) + t * kFlagStride;
        const long long t0 = clock64();
        for (;;) {
            uint32_t seen;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
            if ((int32_t)(seen - e) >= 0) break;
            if (clock64() - t0 > timeout_cycles) {
                *err = 2u;
                break;
            }
        }
    }
    __syncwarp();
    if (t == 0) *epoch = e;
}

// sum of the ranks' result rows (8 x u64 each, all-gathered by k_push_rows + barrier) -> out[8]; sums wrap mod 2^64
__global__ void k_reduce_rows(const unsigned long long* __restrict__ rows, uint32_t world, unsigned long long* __restrict__ out) {
    if (threadIdx.x < 8) {
        unsigned long long s = 0;
        for (uint32_t g = 0; g < world; g++) s += rows[g * 8 + threadIdx.x];
        out[threadIdx.x] = s;
    }
}

// Non-sliceable filters (BASIC k > 1: a key's bits spread over the whole filter): every rank has inserted its local R into
// a full-size partial filter; rank r ORs slice r of all partials (peer loads) and stores the result into slice r of every
// rank's final filter (peer stores) -- reduce-scatter + all-gather of a bitwise OR, which NCCL has no operator for.
__global__ void k_filter_or_bcast(PeerPtrs partials, PeerPtrs finals, uint32_t world, uint64_t first16, uint64_t n16) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        uint4 acc = make_uint4(0u, 0u, 0u, 0u);
        for (uint32_t g = 0; g < world; g++) {
            const uint4 b = reinterpret_cast<const uint4*>(partials.p[g])[first16 + i];
            acc.x |= b.x; acc.y |= b.y; acc.z |= b.z; acc.w |= b.w;
        }
        for (uint32_t g = 0; g < world; g++) reinterpret_cast<uint4*>(finals.p[g])[first16 + i] = acc;
    }
}

// bitwise OR of a partial filter into the destination (building block of the NCCL reference path)
__global__ void k_filter_or(uint4* __restrict__ dst, const uint4* __restrict__ src, uint64_t n16) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        uint4 a = dst[i], b = src[i];
        dst[i] = make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w);
    }
}

// ---- K4: scatter with shared-memory staging ---------------------------------------------------------------------
// replaces the scatter loop (:842-849) and pass-2 radix_cluster (:574-608).
// Persistent CTAs; every CTA walks its tiles of kScatterTile tuples through a kScatterStages-deep ring of TMA
// bulk loads (cp.async.bulk -> mbarrier), so the read of the next tiles is in flight while the current tile
// is sorted by destination bin in shared memory (block-level shared atomics -> ranks), claims one contiguous range per
// non-empty bin from the cursors and writes every bin's run with coalesced stores.
// LEVEL 1: input = a whole relation (this rank's chunk), bin = pid >> b2 (a GLOBAL level-1 bin), cursor index = bin,
//          output = this rank's staging buffer (peer-mapped memory when several GPUs join).
// LEVEL 2: input = the segments of the OWNED level-1 bins in the staging buffers of ALL ranks (k_scan_dist): the bulk
//          loads of a tile read peer memory over NVLink, so this kernel is a fused all-to-all + partition pass -- no send
//          or receive buffers, no remote stores or atomics, no collective call; every tuple crosses NVLink once, as part
//          of a large sequential read. Work item = (segment, tile) from tile_off, bin = pid & (2^b2-1), cursor index =
//          local pid, output = this rank's partitions.
struct ScatterItem {
    uint64_t src_al;   // first tuple index of the bulk load (even: 16-byte aligned)
    uint32_t skip;     // 0/1 tuples to skip at the head of the staged tile
    uint32_t cnt;      // tuples of this tile
    uint32_t cbase;    // cursor base index
    uint32_t bytes;    // bulk-load size
    uint32_t src;      // LEVEL 2: rank whose staging buffer holds the tile
};

struct PeerBufs {
    uint2* buf[kMaxPeers];  // staging buffers of all ranks (peer device pointers mapped into this process)
};

// tile_off / seg_start / seg_cnt: the CTA's shared-memory copies of the segment tables (LEVEL 2)
template <int LEVEL>
__device__ __forceinline__ ScatterItem scatter_item(uint64_t item, uint64_t n, const uint32_t* tile_off,
                                                    const uint32_t* seg_start, const uint32_t* seg_cnt, uint32_t seg_lo,
                                                    uint32_t seg_hi, uint32_t G, uint32_t b2) {
    ScatterItem it;
    uint64_t src0;
    if (LEVEL == 1) {
        src0 = item * kScatterTile;
        it.cnt = (uint32_t)min((uint64_t)kScatterTile, n - src0);
        it.cbase = 0u;
        it.src = 0u;
    } else {
        uint32_t lo = seg_lo, hi = seg_hi;  // segment j with tile_off[j] <= item < tile_off[j+1]
        while (hi - lo > 1u) {
            uint32_t mid = (lo + hi) >> 1;
            if (tile_off[mid] <= (uint32_t)item) lo = mid; else hi = mid;
        }
        const uint32_t first = ((uint32_t)item - tile_off[lo]) * kScatterTile;
        src0 = (uint64_t)seg_start[lo] + first;
        it.cnt = min((uint32_t)kScatterTile, seg_cnt[lo] - first);
        it.cbase = (lo / G) << b2;
        it.src = lo % G;
    }
    it.skip = (uint32_t)(src0 & 1ull);
    it.src_al = src0 - it.skip;
    it.bytes = ((it.cnt + it.skip + 1u) & ~1u) * 8u;
    return it;
}

template <int LEVEL, int PMODE>
__global__ void __launch_bounds__(kScatterThreads, HWBRJ_SCATTER_MINBLOCKS)
k_scatter(const uint2* __restrict__ in, PeerBufs stages, uint2* __restrict__ out,
          const unsigned long long* __restrict__ n_ptr, uint64_t n_static, const uint32_t* __restrict__ tile_off,
          const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ seg_cnt, uint32_t* __restrict__ cursor,
          PartFn pf, const uint32_t* __restrict__ g_crc, uint32_t nbins, uint32_t seg_lo, uint32_t seg_hi, uint32_t G,
          const uint32_t* __restrict__ abort_flag) {
    constexpr int PER = kScatterTile / kScatterThreads;
    constexpr int NB = 1 << kMaxLevelBits;
    static_assert(kScatterThreads >= 32 + NB, "claims run on threads 32.. beside the scan warp");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint2* raw = reinterpret_cast<uint2*>(smem_raw);                       // [stages][kScatterStageTuples]
    uint2* sorted = raw + kScatterStages * kScatterStageTuples;            // [kScatterTile]
    __shared__ uint32_t hist[NB];       // tuples per bin of this tile (block-level shared atomics give the ranks)
    __shared__ uint32_t binstart[NB];   // exclusive scan of hist
    __shared__ uint32_t gclaim[NB];
    __shared__ __align__(8) uint64_t mbar[kScatterStages];
    __shared__ ScatterItem desc[kScatterStages];
    __shared__ uint8_t sorted_bin[kScatterTile];
    __shared__ uint32_t crc_tab[PMODE == 2 ? kCrcSmemWords : 1];
    // LEVEL 2: the segment tables, so that finding a tile's segment is a binary search in shared memory
    __shared__ uint32_t s_tile_off[LEVEL == 2 ? NB + 1 : 1], s_seg_start[LEVEL == 2 ? NB : 1], s_seg_cnt[LEVEL == 2 ? NB : 1];
    if (abort_flag && *abort_flag) return;  // a buffer would overflow (k_scan_dist): nothing is read or written
    const uint64_t n = n_ptr ? min((uint64_t)*n_ptr, n_static) : n_static;
    const uint32_t b2 = pf.b2;
    const uint32_t submask = (1u << b2) - 1u;
    // LEVEL 2 handles the segments [seg_lo, seg_hi): a launch may cover only a part of the owned bins, so that the pull of
    // the next part overlaps whatever consumes this one (items are numbered over all segments)
    const uint64_t item0 = LEVEL == 1 ? 0ull : (uint64_t)tile_off[seg_lo];
    const uint64_t nitems = LEVEL == 1 ? (n + kScatterTile - 1) / kScatterTile : (uint64_t)tile_off[seg_hi];
    if (PMODE == 2) load_crc_tab(crc_tab, g_crc);
    if (LEVEL == 2) {
        for (uint32_t i = seg_lo + threadIdx.x; i <= seg_hi && i <= (uint32_t)NB; i += kScatterThreads) {
            s_tile_off[i] = tile_off[i];
            if (i < seg_hi) {
                s_seg_start[i] = seg_start[i];
                s_seg_cnt[i] = seg_cnt[i];
            }
        }
    }
    if (threadIdx.x == 0) {
        for (int st = 0; st < kScatterStages; st++) mbar_init(&mbar[st], 1u);
        mbar_fence_init();
    }
    __syncthreads();
    // The bulk loads are issued by a thread that has nothing else to do between the barriers (A) and (C): warp 0 scans
    // the bin counts and threads 32..159 wait for their cursor atomics there, and all of them would wait for the issue.
    constexpr uint32_t kIssuer = kScatterThreads - 32;
    static_assert(kIssuer >= 32u + NB, "the issuing thread must not be a scanning or claiming thread");
    auto issue = [&](uint64_t item, int st) {  // thread kIssuer only
        ScatterItem it = scatter_item<LEVEL>(item, n, s_tile_off, s_seg_start, s_seg_cnt, seg_lo, seg_hi, G, b2);
        desc[st] = it;
        const uint2* base = LEVEL == 1 ? in : stages.buf[it.src];  // LEVEL 2: a bulk read over NVLink when src is a peer
        mbar_arrive_expect_tx(&mbar[st], it.bytes);
        bulk_g2s(raw + st * kScatterStageTuples, base + it.src_al, it.bytes, &mbar[st]);
    };
    if (threadIdx.x == kIssuer)
        for (int st = 0; st < kScatterStages; st++) {
            uint64_t item = item0 + (uint64_t)blockIdx.x + (uint64_t)st * gridDim.x;
            if (item < nitems) issue(item, st);
        }
    uint32_t it_local = 0;
    for (uint64_t item = item0 + blockIdx.x; item < nitems; item += gridDim.x, it_local++) {
        const int st = it_local % kScatterStages;
        const uint32_t parity = (it_local / kScatterStages) & 1u;
        if (threadIdx.x < NB) hist[threadIdx.x] = 0u;
        mbar_wait(&mbar[st], parity);
        const ScatterItem d = desc[st];
        const uint2* tile = raw + st * kScatterStageTuples + d.skip;
        const uint32_t cnt = d.cnt;
        __syncthreads();  // hist zeroed, tile landed
        uint2 t[PER];
        uint32_t rank[PER];  // bin in the high 8 bits, rank within the bin in the low 24
#pragma unroll
        for (int j = 0; j < PER; j++) {
            uint32_t idx = threadIdx.x + j * kScatterThreads;
            if (idx < cnt) {
                t[j] = tile[idx];
                const uint32_t pid = pid_of<PMODE>(pf, crc_tab, t[j].x);
                const uint32_t bin = LEVEL == 1 ? pid >> b2 : pid & submask;
                rank[j] = (bin << 24) | atomicAdd(&hist[bin], 1u);
            }
        }
        __syncthreads();  // (A) every thread holds its tuples in registers: the stage can be refilled
        if (threadIdx.x == kIssuer) {
            uint64_t nxt = item + (uint64_t)kScatterStages * gridDim.x;
            if (nxt < nitems) issue(nxt, st);
        }
        if (threadIdx.x < 32) {  // warp 0: exclusive scan over the (<=128) bin counts, 4 per lane
            uint32_t v[4], sum = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t bb = threadIdx.x * 4 + q;
                v[q] = (bb < nbins) ? hist[bb] : 0u;
                sum += v[q];
            }
            uint32_t inc = sum;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                uint32_t u = __shfl_up_sync(0xffffffffu, inc, dd);
                if ((int)threadIdx.x >= dd) inc += u;
            }
            uint32_t run = inc - sum;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t bb = threadIdx.x * 4 + q;
                if (bb < nbins) binstart[bb] = run;
                run += v[q];
            }
        } else if (threadIdx.x - 32u < nbins) {  // meanwhile: one output range per non-empty bin
            const uint32_t bb = threadIdx.x - 32u;
            const uint32_t tot = hist[bb];
            gclaim[bb] = tot ? atomicAdd(&cursor[d.cbase + bb], tot) : 0u;
        }
        __syncthreads();  // (C)
        // sort the tile by bin in shared memory (the bin travels with the tuple: recomputing a hash bin costs more than a
        // byte of shared memory; a 4-byte output position per tuple instead was measured slower -- it costs a CTA per SM)
#pragma unroll
        for (int j = 0; j < PER; j++) {
            uint32_t idx = threadIdx.x + j * kScatterThreads;
            if (idx < cnt) {
                const uint32_t bin = rank[j] >> 24;
                const uint32_t pos = binstart[bin] + (rank[j] & 0xFFFFFFu);
                sorted[pos] = t[j];
                sorted_bin[pos] = (uint8_t)bin;
            }
        }
        __syncthreads();  // (D)
        for (uint32_t i = threadIdx.x; i < cnt; i += kScatterThreads) {
            const uint32_t bin = sorted_bin[i];
            out[gclaim[bin] + (i - binstart[bin])] = sorted[i];
        }
        // next iteration: hist is rewritten before its first barrier, binstart/gclaim after (A'), sorted after (C'):
        // no thread can pass that first barrier before every thread has finished this write-out loop
    }
}

// ---- K1': Bloom filter built from hash partitions, in shared memory, without global atomics ------------------------
// When the join partitions on the filter-slice index (PMODE 1/2), every partition owns one contiguous slice of
// m / 2^bits bits: a CTA zeroes the slice in shared memory, ORs in the bits of its partition's keys with shared-memory
// atomics and writes the slice out with coalesced stores. This replaces add_generic's atomic OR per key
// (bloom_filter.c:74-89) for BASIC k <= 1 and for BLOCKED with any k (all k bits of a key lie in its block, and a slice is
// a whole number of blocks); the bitmap is byte-identical. Every word of the filter is written, so no zero-fill is needed.
// PEER: the slice is stored into the filter of EVERY rank (peer stores over NVLink): the slice build and the all-gather
// that replicates the filter are one kernel.
template <bool BLOCKED, bool PEER>
__global__ void __launch_bounds__(512) k_filter_from_parts(const uint2* __restrict__ Rp, const uint32_t* __restrict__ r_off,
                                                          uint32_t PL, uint32_t gbase, uint32_t slice_words, BloomParams bp,
                                                          const uint32_t* __restrict__ g_crc, PeerPtrs filters,
                                                          uint32_t world, uint32_t nbuf) {
    extern __shared__ uint32_t s_slice[];  // nbuf (1 or 2) slices: with 2 only one barrier separates partitions
    __shared__ uint32_t crc_tab[BLOCKED ? kCrcSmemWords : 1];
    const uint32_t slice_mask = slice_words * 32u - 1u;
    const uint64_t pol = policy_evict_first();
    constexpr int U = 4;  // independent loads in flight per thread
    if (BLOCKED) load_crc_tab(crc_tab, g_crc);
    for (uint32_t i = threadIdx.x; i < nbuf * slice_words; i += blockDim.x) s_slice[i] = 0u;
    __syncthreads();
    uint32_t par = 0;
    for (uint32_t p = blockIdx.x; p < PL; p += gridDim.x, par = nbuf == 2 ? par ^ 1u : 0u) {
        uint32_t* sl = s_slice + par * slice_words;
        const uint32_t lo = r_off[p], hi = r_off[p + 1];
        for (uint32_t i0 = lo + threadIdx.x; i0 < hi; i0 += U * blockDim.x) {
            uint32_t key[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint32_t i = i0 + u * blockDim.x;
                key[u] = i < hi ? ld_stream_v2(Rp + i, pol).x : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (i0 + u * blockDim.x < hi) {
                    uint32_t base, h, y;
                    bp.blocked = BLOCKED ? 1u : 0u;
                    bloom_start(bp, crc_tab, key[u], base, h, y);
                    for (uint32_t i = 0; i < bp.k; i++) {  // the index sequence of add_generic, inside the slice
                        const uint32_t a = (base + h) & slice_mask;
                        atomicOr(&sl[a >> 5], 1u << (a & 31u));
                        h = (h + y) & bp.size_mask;
                        y = (y + i + 1u) & bp.size_mask;
                    }
                }
            }
        }
        __syncthreads();  // slice complete
        const size_t first = (size_t)(gbase + p) * slice_words;
        if ((slice_words & 3u) == 0u) {  // 16 bytes per thread and store: slices and filters are 16-byte aligned
            uint4* sl4 = reinterpret_cast<uint4*>(sl);
            for (uint32_t i = threadIdx.x; i < slice_words / 4u; i += blockDim.x) {
                const uint4 w = sl4[i];
                sl4[i] = make_uint4(0u, 0u, 0u, 0u);  // ready for a later partition
                if (PEER) {
                    for (uint32_t g = 0; g < world; g++) reinterpret_cast<uint4*>(filters.p[g])[first / 4 + i] = w;
                } else {
                    reinterpret_cast<uint4*>(filters.p[0])[first / 4 + i] = w;
                }
            }
        } else {
            for (uint32_t i = threadIdx.x; i < slice_words; i += blockDim.x) {
                const uint32_t w = sl[i];
                if (PEER) {
                    for (uint32_t g = 0; g < world; g++) reinterpret_cast<uint32_t*>(filters.p[g])[first + i] = w;
                } else {
                    reinterpret_cast<uint32_t*>(filters.p[0])[first + i] = w;
                }
                sl[i] = 0u;  // with two buffers the next partition uses the other one meanwhile
            }
        }
        if (nbuf != 2) __syncthreads();
    }
}

// ---- join work list: one item per (partition, S chunk) --------------------------------------------------------
__global__ void __launch_bounds__(1024) k_worklist(const uint32_t* __restrict__ r_off, const uint32_t* __restrict__ s_off,
                                                  uint32_t P, uint32_t* __restrict__ work_off,
                                                  uint32_t* __restrict__ work_part) {
    // single CTA (P <= 16384). work_part[item] = partition of the item, so that the join kernel finds its partition
    // with one load instead of a binary search over work_off.
    __shared__ uint32_t warp_sums[33];
    __shared__ uint32_t s_big[1024][3], s_nbig;  // partitions with many items are expanded by the whole CTA
    if (threadIdx.x == 0) s_nbig = 0u;
    const uint32_t per = (P + 1023u) / 1024u;
    const uint32_t lo = threadIdx.x * per;
    uint32_t items[kBinsPerThread];
    uint32_t local = 0;
#pragma unroll
    for (int i = 0; i < kBinsPerThread; i++) {
        const uint32_t p = lo + i;
        items[i] = 0u;
        if ((uint32_t)i < per && p < P) {
            const uint32_t nr = r_off[p + 1] - r_off[p], ns = s_off[p + 1] - s_off[p];
            items[i] = (nr && ns) ? (ns + kSChunk - 1) / kSChunk : 0u;
        }
        local += items[i];
    }
    uint32_t run = block_exclusive_scan_1024(local, warp_sums);
#pragma unroll
    for (int i = 0; i < kBinsPerThread; i++) {
        const uint32_t p = lo + i;
        if ((uint32_t)i < per && p < P) {
            work_off[p] = run;
            // > kBigItems chunks = > 4 M tuples: fewer than 2^32 / 2^22 = 1024 such partitions can exist
            if (items[i] > kBigItems) {
                const uint32_t slot = atomicAdd(&s_nbig, 1u);
                s_big[slot][0] = p;
                s_big[slot][1] = run;
                s_big[slot][2] = items[i];
            } else {
                for (uint32_t c = 0; c < items[i]; c++) work_part[run + c] = p;
            }
            run += items[i];
        }
    }
    if (threadIdx.x == 0) work_off[P] = warp_sums[32];
    __syncthreads();
    const uint32_t nbig = s_nbig;
    for (uint32_t b = 0; b < nbig; b++) {
        const uint32_t p = s_big[b][0], first = s_big[b][1], n_items = s_big[b][2];
        for (uint32_t c = threadIdx.x; c < n_items; c += 1024u) work_part[first + c] = p;
    }
}

// ---- K5: per-partition build + probe with the table in shared memory -----------------------------------------------
// replaces bucket_chaining_join (:260-329): bucket heads + next links over the R partition held in shared
// memory, idx = (key >> radix_bits) & (N-1) (HASH_BIT_MODULO with MASK=(N-1)<<bits), every equal key on the
// chain counts. R partitions larger than kTableRound are processed in rounds; S partitions larger than kSChunk are
// split over several work items (each rebuilds the table) so that skewed S does not serialise on one SM.
// HASHPART: partitions come from the hash (filter-slice) partitioning, so the table index is key & (N-1).
// PAIRS: materialise the output like -DJOIN_RESULT_MATERIALIZE does (:307-312): one {R.payload, S.payload} tuple per
// match, appended to pairs_out through a warp-aggregated atomic cursor; pairs beyond pair_capacity are dropped (the
// count stays exact, the caller retries with a larger buffer).
//
// The kernel is latency-bound (a partition is ~60 KB of R and ~60 KB of S), so the round trips are overlapped:
//  * the R partition arrives as ONE TMA bulk copy straight into the table (cp.async.bulk -> mbarrier); the <= 2 tuples
//    outside its 16-byte-aligned interior are loaded by two other threads;
//  * while it is in flight, thread 0 claims the NEXT work item (atomic + one load of work_part), every thread clears
//    its part of head[] and issues its first batch of S loads;
//  * in the probe loop the next batch of S is requested before the current one is processed.
// Tuple j of the round lives in slot j + off, off = parity of its global address / 8 (keeps the bulk copy aligned).
constexpr uint32_t kTableRound = kTableCap - 2;  // tuples per round: leaves room for the alignment shift
constexpr int kJoinSmemBytes = kTableCap * (8 + 4 + 2);

template <bool HASHPART, bool PAIRS = false>
__global__ void __launch_bounds__(kJoinThreads, 2) k_join(const uint2* __restrict__ Rp, const uint32_t* __restrict__ r_off,
                                                      const uint2* __restrict__ Sp, const uint32_t* __restrict__ s_off,
                                                      const uint32_t* __restrict__ work_off,
                                                      const uint32_t* __restrict__ work_part, uint32_t P, uint32_t bits,
                                                      uint32_t* __restrict__ item_counter, JoinAccum* __restrict__ acc_out,
                                                      uint2* __restrict__ pairs_out = nullptr,
                                                      unsigned long long* __restrict__ pair_cursor = nullptr,
                                                      unsigned long long pair_capacity = 0) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint2* tab = reinterpret_cast<uint2*>(smem_raw);                  // kTableCap slots
    uint32_t* head = reinterpret_cast<uint32_t*>(tab + kTableCap);    // kTableCap heads (slot+1, 0 = empty)
    uint16_t* next = reinterpret_cast<uint16_t*>(head + kTableCap);   // kTableCap links
    __shared__ uint32_t s_item[2], s_part[2];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ unsigned long long s_red[5][kJoinThreads / 32];
    unsigned long long matches = 0, cpair = 0, crpay = 0, cspay = 0, ckey = 0;
    const uint32_t total = work_off[P];
    const uint64_t pol = policy_evict_first();
    constexpr int U = HWBRJ_JOIN_UNROLL;  // independent 8-byte S loads per thread and batch
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1u);
        mbar_fence_init();
        const uint32_t it = atomicAdd(item_counter, 1u);
        s_item[0] = it;
        if (it < total) s_part[0] = work_part[it];
    }
    auto emit = [&](uint32_t rpay, const uint2 s) {  // one output pair {R.payload, S.payload} (:303-315)
        if (PAIRS) {
            const uint32_t am = __activemask();
            const uint32_t lane = threadIdx.x & 31u;
            const int leader = __ffs(am) - 1;
            unsigned long long pos = 0ull;
            if ((int)lane == leader) pos = atomicAdd(pair_cursor, (unsigned long long)__popc(am));
            pos = __shfl_sync(am, pos, leader) + __popc(am & ((1u << lane) - 1u));
            if (pos < pair_capacity) pairs_out[pos] = make_uint2(rpay, s.y);
        }
        matches++;
        cpair += mix64(rpay, s.y);
        crpay += rpay;
        cspay += s.y;
        ckey += s.x;
    };
    uint32_t par = 0u, phase = 0u;
    for (;; par ^= 1u) {
        __syncthreads();  // previous item completely done: table free, s_item[par] visible
        const uint32_t item = s_item[par];
        if (item >= total) break;
        const uint32_t p = s_part[par];
        const uint32_t chunk = item - work_off[p];
        const uint32_t r0 = r_off[p], nr = r_off[p + 1] - r0;
        const uint32_t sbeg = s_off[p] + chunk * kSChunk;
        const uint32_t send = min(s_off[p + 1], sbeg + (uint32_t)kSChunk);
        for (uint32_t rb = 0; rb < nr; rb += kTableRound) {
            const uint32_t cnt = min(kTableRound, nr - rb);
            uint32_t N = 1u;
            while (N < cnt) N <<= 1;
            const uint32_t nmask = N - 1u;
            if (rb) __syncthreads();  // probe of the previous round done
            const uint2* Rbase = Rp + (uint64_t)r0 + rb;
            const uint32_t off = (uint32_t)((reinterpret_cast<uintptr_t>(Rbase) >> 3) & 1u);
            const uint32_t nbulk = (cnt - off) & ~1u;  // tuples [off, off + nbulk) are 16-byte aligned pairs (cnt >= 1)
            uint32_t nxt = 0u;
            if (threadIdx.x == 0) {
                if (nbulk) {
                    mbar_arrive_expect_tx(&s_bar, nbulk * 8u);
                    bulk_g2s(tab + 2u * off, Rbase + off, nbulk * 8u, &s_bar);
                }
                if (rb == 0) nxt = atomicAdd(item_counter, 1u);  // next item: its latency hides under this one
            } else if (threadIdx.x == 32) {
                if (off) tab[1] = ld_stream_v2(Rbase, pol);
            } else if (threadIdx.x == 64) {
                if (off + nbulk < cnt) tab[cnt - 1u + off] = ld_stream_v2(Rbase + cnt - 1u, pol);
            }
            for (uint32_t i = threadIdx.x; i < N; i += kJoinThreads) head[i] = 0u;
            // first batch of S: in flight during the build
            uint32_t i0 = sbeg + threadIdx.x;
            uint2 sv[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint32_t i = i0 + u * kJoinThreads;
                sv[u] = i < send ? ld_stream_v2(Sp + i, pol) : make_uint2(0u, 0u);
            }
            __syncthreads();  // head[] cleared, edge tuples stored
            if (threadIdx.x == 0 && rb == 0) {
                s_item[par ^ 1u] = nxt;  // read by the CTA after the barrier at the top of the next item
                if (nxt < total) s_part[par ^ 1u] = work_part[nxt];
            }
            if (nbulk) {
                mbar_wait(&s_bar, phase);
                phase ^= 1u;
            }
            constexpr int UB = 4;  // build: independent table inserts per thread and step
            for (uint32_t j0 = threadIdx.x; j0 < cnt; j0 += UB * kJoinThreads) {
                uint32_t key[UB], old[UB];
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    const uint32_t j = j0 + u * kJoinThreads;
                    key[u] = j < cnt ? tab[j + off].x : 0u;
                }
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    const uint32_t j = j0 + u * kJoinThreads;
                    if (j < cnt) old[u] = atomicExch(&head[(HASHPART ? key[u] : (key[u] >> bits)) & nmask], j + off + 1u);
                }
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    const uint32_t j = j0 + u * kJoinThreads;
                    if (j < cnt) next[j + off] = (uint16_t)old[u];
                }
            }
            __syncthreads();  // table complete
            for (;;) {
                const uint32_t i1 = i0 + U * kJoinThreads;
                uint2 sn[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const uint32_t i = i1 + u * kJoinThreads;
                    sn[u] = i < send ? ld_stream_v2(Sp + i, pol) : make_uint2(0u, 0u);
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (i0 + u * kJoinThreads < send) {
                        const uint2 s = sv[u];
                        // The chain walk only compares keys; the (expensive) bookkeeping of a match runs once per S
                        // tuple after the walk, where the warp has reconverged. A second match of the same S tuple
                        // (duplicate build keys) flushes the first one inside the loop.
                        uint32_t m_rpay = 0u;
                        bool have = false;
                        for (uint32_t hit = head[(HASHPART ? s.x : (s.x >> bits)) & nmask]; hit; hit = next[hit - 1u]) {
                            const uint2 r = tab[hit - 1u];
                            if (r.x == s.x) {
                                if (have) emit(m_rpay, s);
                                m_rpay = r.y;
                                have = true;
                            }
                        }
                        if (have) emit(m_rpay, s);
                    }
                }
                if (i1 >= send) break;
                i0 = i1;
#pragma unroll
                for (int u = 0; u < U; u++) sv[u] = sn[u];
            }
        }
    }
    // block reduction
    unsigned long long v[5] = {matches, cpair, crpay, cspay, ckey};
#pragma unroll
    for (int q = 0; q < 5; q++) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], d);
        if ((threadIdx.x & 31) == 0) s_red[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        unsigned long long s = 0;
        for (int w = 0; w < kJoinThreads / 32; w++) s += s_red[threadIdx.x][w];
        if (s) atomicAdd(reinterpret_cast<unsigned long long*>(acc_out) + threadIdx.x, s);
    }
}

// ---- on-device generator (SURVEY.md A.4 closed form of generator.c:162-195,341-387) ------------------------------
// position permutation: 4-round Feistel over ceil(log2 n) bits with cycle walking (a bijection on [0,n))
__device__ __forceinline__ uint64_t feistel_perm(uint64_t x, uint64_t n, uint32_t half_bits, uint64_t seed) {
    const uint64_t hmask = (1ull << half_bits) - 1ull;
    do {
        uint64_t l = x >> half_bits, r = x & hmask;
#pragma unroll
        for (int round = 0; round < 4; round++) {
            uint64_t f = (r + seed + (uint64_t)round * 0x9e3779b97f4a7c15ULL) * 0xbf58476d1ce4e5b9ULL;
            f ^= f >> 29;
            f *= 0x94d049bb133111ebULL;
            f ^= f >> 32;
            uint64_t nl = r;
            r = l ^ (f & hmask);
            l = nl;
        }
        x = (l << half_bits) | r;
    } while (x >= n);
    return x;
}

// kind 0: R = permutation of 1..n, payload = position (main.c:430-431)
// kind 1: S = nb keys ((e mod r)+1) and na keys r+1+e', payload = position (main.c:464-465)
// The kernel fills positions [begin, begin+count) of the global relation of n tuples (a shard).
__global__ void k_generate(uint2* __restrict__ out, uint64_t n, int kind, uint64_t r, uint64_t nb, uint32_t half_bits,
                           uint64_t seed, uint64_t begin, uint64_t count) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t li = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; li < count; li += stride) {
        const uint64_t i = begin + li;
        uint64_t e = feistel_perm(i, n, half_bits, seed);
        uint32_t key;
        if (kind == 0) key = (uint32_t)(e + 1ull);
        else if (e < nb) key = (uint32_t)(e % r + 1ull);
        else {
            uint64_t span = 2147483647ull - r;  // keys wrap back to r+1 after INT_MAX (generator.c:191-193)
            key = (uint32_t)(r + 1ull + (e - nb) % (span ? span : 1ull));
        }
        out[li] = make_uint2(key, (uint32_t)i);
    }
}

// ---- kind 2: Zipf-distributed foreign keys (create_relation_zipf, generator.c:659-676 -> genzipf.c) ------------------
// Same construction as the reference: an alphabet = random permutation of 1..r (genzipf.c:27-52; here the Feistel
// permutation), a lookup table of the cumulated density sum_{j<=i} j^-theta / sum_j j^-theta (:59-92), and per tuple a
// uniform number with RAND_MAX resolution and the same binary search (:121-145). The reference draws from glibc rand()
// serially; here tuple i draws from a counter-based generator of (i, seed), so a shard can be generated on its own.
// Payload = position (the reference leaves it uninitialised, genzipf.c:147-148).
constexpr int kZipfChunk = 4096;  // table entries per CTA of the blocked scan (256 threads x 16)

// phase 1: terms j^-theta of one chunk, inclusive scan inside the chunk, chunk total to sums[chunk]
__global__ void __launch_bounds__(256) k_zipf_scan_chunks(double* __restrict__ lut, uint64_t r, double theta,
                                                         double* __restrict__ sums) {
    __shared__ double warp_tot[8];
    const uint64_t base = (uint64_t)blockIdx.x * kZipfChunk + (uint64_t)threadIdx.x * 16u;
    double v[16], run = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint64_t j = base + i;  // table index, value index j+1
        run += j < r ? 1.0 / pow((double)(j + 1), theta) : 0.0;
        v[i] = run;
    }
    // exclusive scan of the per-thread totals over the CTA
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    double inc = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (uint32_t)d) inc += u;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    double off = inc - run;
    for (uint32_t w = 0; w < wid; w++) off += warp_tot[w];
#pragma unroll
    for (int i = 0; i < 16; i++)
        if (base + i < r) lut[base + i] = off + v[i];
    if (threadIdx.x == 255) sums[blockIdx.x] = off + run;
}
// phase 2 (one CTA): exclusive scan of the chunk totals in place, grand total to sums[nchunks]
__global__ void __launch_bounds__(1024) k_zipf_scan_sums(double* __restrict__ sums, uint32_t nchunks) {
    __shared__ double part[1024];
    const uint32_t per = (nchunks + 1023u) / 1024u;
    const uint32_t lo = threadIdx.x * per;
    double local = 0.0;
    for (uint32_t i = 0; i < per; i++)
        if (lo + i < nchunks) local += sums[lo + i];
    part[threadIdx.x] = local;
    __syncthreads();
    if (threadIdx.x == 0) {  // 1024 additions: not worth a parallel scan
        double run = 0.0;
        for (uint32_t t = 0; t < 1024u; t++) {
            const double x = part[t];
            part[t] = run;
            run += x;
        }
        sums[nchunks] = run;
    }
    __syncthreads();
    double run = part[threadIdx.x];
    for (uint32_t i = 0; i < per; i++)
        if (lo + i < nchunks) {
            const double x = sums[lo + i];
            sums[lo + i] = run;
            run += x;
        }
}
// phase 3: add the chunk offsets and normalise (lut[i] = sum / scaling_factor, genzipf.c:86-89)
__global__ void __launch_bounds__(256) k_zipf_finish(double* __restrict__ lut, uint64_t r, const double* __restrict__ sums,
                                                    uint32_t nchunks) {
    const double total = sums[nchunks];
    const double off = sums[blockIdx.x];
    const uint64_t base = (uint64_t)blockIdx.x * kZipfChunk;
    for (uint32_t i = threadIdx.x; i < (uint32_t)kZipfChunk; i += 256u)
        if (base + i < r) lut[base + i] = (lut[base + i] + off) / total;
}

__global__ void k_generate_zipf(uint2* __restrict__ out, uint64_t begin, uint64_t count, const double* __restrict__ lut,
                                uint32_t r, uint32_t half_bits_r, uint64_t seed) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t li = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; li < count; li += stride) {
        const uint64_t i = begin + li;
        // uniform in [0,1] with the resolution of rand()/RAND_MAX (genzipf.c:124)
        uint64_t z = i * 0x9e3779b97f4a7c15ULL + seed;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        z ^= z >> 31;
        const double u = (double)(uint32_t)(z >> 33) / 2147483647.0;
        uint32_t left = 0u, right = r - 1u, pos;
        if (lut[0] >= u) pos = 0u;
        else {
            while (right - left > 1u) {
                const uint32_t mid = (left + right) / 2u;
                if (lut[mid] < u) left = mid;
                else right = mid;
            }
            pos = right;
        }
        const uint32_t key = (uint32_t)feistel_perm(pos, r, half_bits_r, seed ^ 0x5bf03635ULL) + 1u;  // alphabet[pos]
        out[li] = make_uint2(key, (uint32_t)i);
    }
}

}  // namespace hwbrj
