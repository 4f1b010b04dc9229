// Micro-benchmarks that size the design of the Bloom probe / insert / scatter kernels on B200.
// Not part of the product path; results are summarised in profiles/ and DESIGN.md.
//
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/probe_micro tools/micro/probe_micro.cu
//   run  : build/probe_micro [log2_probes=28]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t crapwow42(uint32_t key) {
    const uint32_t n = 0x5052acdbu;
    uint32_t h = 4u, k = 4u + 42u + n;
    uint64_t p = (uint64_t)key * n;
    h ^= (uint32_t)p; k ^= (uint32_t)(p >> 32);
    p = (uint64_t)(h ^ (k + n)) * n;
    h ^= (uint32_t)p; k ^= (uint32_t)(p >> 32);
    return k ^ h;
}

enum { LD_NC = 0, LD_NOALLOC = 1, LD_TEX = 2, LD_U8 = 3, LD_ATOM = 4, LD_CG = 5 };

template <int KIND>
__device__ __forceinline__ uint32_t probe_word(const uint32_t* __restrict__ f, cudaTextureObject_t tex, uint32_t h) {
    if (KIND == LD_NC) {
        return __ldg(f + (h >> 5));
    } else if (KIND == LD_NOALLOC) {
        uint32_t v;
        asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(f + (h >> 5)));
        return v;
    } else if (KIND == LD_CG) {
        return __ldcg(f + (h >> 5));
    } else if (KIND == LD_TEX) {
        return tex1Dfetch<uint32_t>(tex, (int)(h >> 5));
    } else if (KIND == LD_U8) {
        uint32_t b = __ldg(reinterpret_cast<const unsigned char*>(f) + (h >> 3));
        return b << (h & 24);
    } else {
        return atomicOr(const_cast<uint32_t*>(f) + (h >> 5), 0u);
    }
}

// pure probe rate: keys synthesised from the index, ILP probes in flight per thread
template <int KIND, int ILP>
__global__ void __launch_bounds__(256) k_probe(const uint32_t* __restrict__ f, cudaTextureObject_t tex,
                                               uint32_t mask, uint64_t n, unsigned long long* out) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t cnt = 0;
    for (uint64_t i = tid; i < n; i += stride * ILP) {
        uint32_t w[ILP], h[ILP];
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            uint64_t idx = i + (uint64_t)j * stride;
            h[j] = crapwow42((uint32_t)idx + 128000001u) & mask;
            w[j] = (idx < n) ? probe_word<KIND>(f, tex, h[j]) : 0u;
        }
#pragma unroll
        for (int j = 0; j < ILP; j++) cnt += (w[j] >> (h[j] & 31)) & 1u;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, (unsigned long long)cnt);
}

__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    uint64_t pol = policy_evict_first();
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    return r;
}

// stream S (2 tuples per 128-bit load, V loads in flight per thread) + probe + count
template <int KIND, int V>
__global__ void __launch_bounds__(256) k_stream_probe(const uint4* __restrict__ S, uint64_t npairs,
                                                      const uint32_t* __restrict__ f, cudaTextureObject_t tex,
                                                      uint32_t mask, unsigned long long* out) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t cnt = 0;
    for (uint64_t i = tid; i < npairs; i += stride * V) {
        uint4 t[V];
#pragma unroll
        for (int j = 0; j < V; j++) {
            uint64_t idx = i + (uint64_t)j * stride;
            t[j] = (idx < npairs) ? ld_stream(S + idx) : make_uint4(0, 0, 0, 0);
        }
        uint32_t h[2 * V], w[2 * V];
#pragma unroll
        for (int j = 0; j < V; j++) {
            h[2 * j] = crapwow42(t[j].x) & mask;
            h[2 * j + 1] = crapwow42(t[j].z) & mask;
            w[2 * j] = probe_word<KIND>(f, tex, h[2 * j]);
            w[2 * j + 1] = probe_word<KIND>(f, tex, h[2 * j + 1]);
        }
#pragma unroll
        for (int j = 0; j < 2 * V; j++) cnt += (w[j] >> (h[j] & 31)) & 1u;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, (unsigned long long)cnt);
}

// stream only (roofline of the S read)
template <int V>
__global__ void __launch_bounds__(256) k_stream(const uint4* __restrict__ S, uint64_t npairs, unsigned long long* out) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t cnt = 0;
    for (uint64_t i = tid; i < npairs; i += stride * V) {
#pragma unroll
        for (int j = 0; j < V; j++) {
            uint64_t idx = i + (uint64_t)j * stride;
            if (idx < npairs) { uint4 t = ld_stream(S + idx); cnt += t.x ^ t.z; }
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt == 0x12345678u) atomicAdd(out, 1ull);
}

// random RED.OR insert
__global__ void __launch_bounds__(256) k_insert(uint32_t* f, uint32_t mask, uint64_t n) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = tid; i < n; i += stride) {
        uint32_t h = crapwow42((uint32_t)i + 1u) & mask;
        atomicOr(f + (h >> 5), 1u << (h & 31));
    }
}

// probes against a shared-memory resident slice (2^20 bits = 128 KB)
__global__ void __launch_bounds__(512) k_probe_smem(const uint32_t* __restrict__ f, uint64_t n, unsigned long long* out) {
    extern __shared__ uint32_t sf[];
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) sf[i] = f[i];
    __syncthreads();
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t cnt = 0;
    for (uint64_t i = tid; i < n; i += stride) {
        uint32_t h = crapwow42((uint32_t)i + 128000001u) & 0xFFFFFu;
        cnt += (sf[h >> 5] >> (h & 31)) & 1u;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, (unsigned long long)cnt);
}

// scatter A: one global atomic per tuple + 8-byte store into slab
__global__ void __launch_bounds__(256) k_scatter_atomic(const uint2* __restrict__ in, uint64_t n, uint2* out,
                                                        uint32_t* cursor, uint32_t pmask, uint32_t cap) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = tid; i < n; i += stride) {
        uint2 t = in[i];
        uint32_t p = t.x & pmask;
        uint32_t pos = atomicAdd(cursor + p, 1u);
        out[(uint64_t)p * cap + pos] = t;
    }
}

// scatter B: CTA tile, shared histogram, one global atomic per (tile, non-empty partition), 8-byte stores
template <int TILE>
__global__ void __launch_bounds__(512) k_scatter_tile(const uint2* __restrict__ in, uint64_t n, uint2* out,
                                                      uint32_t* cursor, uint32_t P, uint32_t cap) {
    extern __shared__ uint32_t sh[];  // P counters -> then P bases
    const uint32_t pmask = P - 1;
    uint64_t ntiles = (n + TILE - 1) / TILE;
    constexpr int PER = TILE / 512;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (uint32_t i = threadIdx.x; i < P; i += 512) sh[i] = 0;
        __syncthreads();
        uint2 t[PER]; uint32_t r[PER];
        uint64_t base = tile * TILE;
#pragma unroll
        for (int j = 0; j < PER; j++) {
            uint64_t idx = base + threadIdx.x + j * 512;
            if (idx < n) { t[j] = in[idx]; r[j] = atomicAdd(&sh[t[j].x & pmask], 1u); }
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < P; i += 512) {
            uint32_t c = sh[i];
            if (c) sh[i] = atomicAdd(cursor + i, c);
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER; j++) {
            uint64_t idx = base + threadIdx.x + j * 512;
            if (idx < n) { uint32_t p = t[j].x & pmask; out[(uint64_t)p * cap + sh[p] + r[j]] = t[j]; }
        }
        __syncthreads();
    }
}

__global__ void k_fill_S(uint2* S, uint64_t n, uint32_t base) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = tid; i < n; i += stride) {
        // pseudo-shuffled distinct keys: odd-multiplier bijection on 2^32, then offset
        uint32_t k = (uint32_t)i * 2654435761u;
        S[i] = make_uint2(base + (k >> 2), (uint32_t)i);
    }
}

struct Timer {
    cudaEvent_t a, b;
    Timer() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
    void start() { CK(cudaEventRecord(a)); }
    float stop() { CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }
};

template <typename F>
float best_of(int reps, F f) {
    Timer t; float best = 1e30f;
    for (int i = 0; i < reps; i++) { t.start(); f(); float ms = t.stop(); best = std::min(best, ms); }
    return best;
}

int main(int argc, char** argv) {
    int lg = argc > 1 ? atoi(argv[1]) : 28;
    uint64_t n = 1ull << lg;
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    int maxPersist = 0, maxWin = 0;
    CK(cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, 0));
    CK(cudaDeviceGetAttribute(&maxWin, cudaDevAttrMaxAccessPolicyWindowSize, 0));
    printf("device %s SMs=%d L2=%d MB persistMax=%d MB winMax=%d MB smemOptin=%zu clock=%d kHz mem=%d kHz bus=%d\n", pr.name,
           pr.multiProcessorCount, pr.l2CacheSize >> 20, maxPersist >> 20, maxWin >> 20, pr.sharedMemPerBlockOptin,
           pr.clockRate, pr.memoryClockRate, pr.memoryBusWidth);
    const int SM = pr.multiProcessorCount;
    unsigned long long* d_out; CK(cudaMalloc(&d_out, 8)); CK(cudaMemset(d_out, 0, 8));

    // filters: fill ~11% via the insert kernel itself
    for (int lgm : {27, 29, 30, 31}) {
        uint64_t mbits = 1ull << lgm; size_t bytes = mbits / 8;
        uint32_t* f; CK(cudaMalloc(&f, bytes)); CK(cudaMemset(f, 0, bytes));
        uint32_t mask = (uint32_t)(mbits - 1);
        uint64_t nins = mbits / 8;  // same load factor as C1 (128M keys into 2^30 bits)
        float ms = best_of(3, [&] { k_insert<<<SM * 8, 256>>>(f, mask, nins); });
        printf("[insert ] m=2^%d  n=%llu  %.3f ms  %.1f Gins/s\n", lgm, (unsigned long long)nins, ms, nins / ms * 1e-6);
        cudaTextureObject_t tex = 0;
        {
            cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = f;
            rd.res.linear.desc = cudaCreateChannelDesc<uint32_t>(); rd.res.linear.sizeInBytes = bytes;
            cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
            if (cudaCreateTextureObject(&tex, &rd, &td, nullptr) != cudaSuccess) { tex = 0; cudaGetLastError(); printf("  (no texture at this size)\n"); }
        }
#define RUNP(KIND, ILP, G, name) { CK(cudaMemset(d_out, 0, 8)); \
        float ms_ = best_of(3, [&] { k_probe<KIND, ILP><<<SM * G, 256>>>(f, tex, mask, n, d_out); }); CK(cudaGetLastError()); \
        printf("[probe  ] m=2^%d %-10s ilp=%d ctas/sm=%d  %.3f ms  %.1f Gprobes/s\n", lgm, name, ILP, G, ms_, n / ms_ * 1e-6); }
        RUNP(LD_NC, 1, 8, "ldg.nc");
        RUNP(LD_NC, 4, 8, "ldg.nc");
        RUNP(LD_NC, 8, 8, "ldg.nc");
        RUNP(LD_NC, 8, 4, "ldg.nc");
        RUNP(LD_NOALLOC, 4, 8, "noalloc");
        RUNP(LD_NOALLOC, 8, 8, "noalloc");
        RUNP(LD_CG, 8, 8, "ld.cg");
        RUNP(LD_U8, 8, 8, "ldg.u8");
        if (tex) { RUNP(LD_TEX, 4, 8, "tex"); RUNP(LD_TEX, 8, 8, "tex"); }
        RUNP(LD_ATOM, 4, 8, "atom.or0");
        if (lgm == 30) {
            // persisting-L2 window over the filter
            CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxPersist));
            cudaStream_t st; CK(cudaStreamCreate(&st));
            for (float hr : {1.0f, 0.7f, 0.5f}) {
                cudaStreamAttrValue av = {};
                av.accessPolicyWindow.base_ptr = f;
                av.accessPolicyWindow.num_bytes = std::min(bytes, (size_t)maxWin);
                av.accessPolicyWindow.hitRatio = hr;
                av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                av.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
                CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
                Timer t; float best = 1e30f;
                for (int r = 0; r < 3; r++) {
                    CK(cudaEventRecord(t.a, st));
                    k_probe<LD_NC, 8><<<SM * 8, 256, 0, st>>>(f, tex, mask, n, d_out);
                    CK(cudaEventRecord(t.b, st)); CK(cudaEventSynchronize(t.b));
                    float ms2; CK(cudaEventElapsedTime(&ms2, t.a, t.b)); best = std::min(best, ms2);
                }
                printf("[probe  ] m=2^%d persist hitRatio=%.1f  %.3f ms  %.1f Gprobes/s\n", lgm, hr, best, n / best * 1e-6);
            }
            cudaStreamAttrValue av = {}; av.accessPolicyWindow.num_bytes = 0;
            CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
            CK(cudaCtxResetPersistingL2Cache());
            CK(cudaStreamDestroy(st));
        }
        if (lgm == 27 || lgm == 30) {
            // stream + probe
            uint64_t ns = n;  // tuples
            uint2* S; CK(cudaMalloc(&S, ns * 8));
            k_fill_S<<<SM * 8, 256>>>(S, ns, 128000001u); CK(cudaDeviceSynchronize());
            float ms0 = best_of(3, [&] { k_stream<4><<<SM * 8, 256>>>((const uint4*)S, ns / 2, d_out); });
            printf("[stream ] n=%llu  %.3f ms  %.1f GB/s\n", (unsigned long long)ns, ms0, ns * 8 / ms0 * 1e-6);
#define RUNS(KIND, V, G, name) { float ms_ = best_of(3, [&] { k_stream_probe<KIND, V><<<SM * G, 256>>>((const uint4*)S, ns / 2, f, tex, mask, d_out); }); CK(cudaGetLastError()); \
            printf("[s+probe] m=2^%d %-8s v=%d ctas/sm=%d  %.3f ms  %.1f Gtuples/s  (S %.0f GB/s)\n", lgm, name, V, G, ms_, ns / ms_ * 1e-6, ns * 8 / ms_ * 1e-6); }
            RUNS(LD_NC, 1, 8, "ldg.nc");
            RUNS(LD_NC, 2, 8, "ldg.nc");
            RUNS(LD_NC, 4, 8, "ldg.nc");
            RUNS(LD_NC, 4, 4, "ldg.nc");
            RUNS(LD_NOALLOC, 4, 8, "noalloc");
            if (tex) RUNS(LD_TEX, 4, 8, "tex");
            CK(cudaFree(S));
        }
        if (tex) cudaDestroyTextureObject(tex);
        CK(cudaFree(f));
    }
    {
        uint32_t* f; CK(cudaMalloc(&f, 131072)); CK(cudaMemset(f, 0x11, 131072));
        CK(cudaFuncSetAttribute(k_probe_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
        float ms = best_of(3, [&] { k_probe_smem<<<SM, 512, 131072>>>(f, n * 4, d_out); });
        printf("[smem   ] 128KB slice  %.3f ms  %.1f Gprobes/s\n", ms, n * 4 / ms * 1e-6);
        CK(cudaFree(f));
    }
    {
        // scatter tests: 2^27 tuples
        uint64_t ns = 1ull << 27;
        uint2 *in, *out; uint32_t* cursor;
        CK(cudaMalloc(&in, ns * 8));
        k_fill_S<<<SM * 8, 256>>>(in, ns, 1u); CK(cudaDeviceSynchronize());
        for (uint32_t P : {2048u, 8192u, 16384u, 32768u}) {
            uint32_t cap = (uint32_t)(ns / P * 5 / 4 + 64);
            CK(cudaMalloc(&out, (uint64_t)P * cap * 8)); CK(cudaMalloc(&cursor, P * 4));
            float msA = best_of(3, [&] { cudaMemsetAsync(cursor, 0, P * 4); k_scatter_atomic<<<SM * 8, 256>>>(in, ns, out, cursor, P - 1, cap); });
            CK(cudaGetLastError());
            printf("[scatter] P=%u atomic-per-tuple      %.3f ms  %.1f Gtuples/s (%.0f GB/s rd+wr)\n", P, msA, ns / msA * 1e-6, ns * 16 / msA * 1e-6);
            CK(cudaFuncSetAttribute(k_scatter_tile<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4));
            CK(cudaFuncSetAttribute(k_scatter_tile<8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4));
            float msB = best_of(3, [&] { cudaMemsetAsync(cursor, 0, P * 4); k_scatter_tile<4096><<<SM * 2, 512, P * 4>>>(in, ns, out, cursor, P, cap); });
            CK(cudaGetLastError());
            printf("[scatter] P=%u tile4096 smem-hist     %.3f ms  %.1f Gtuples/s (%.0f GB/s rd+wr)\n", P, msB, ns / msB * 1e-6, ns * 16 / msB * 1e-6);
            float msC = best_of(3, [&] { cudaMemsetAsync(cursor, 0, P * 4); k_scatter_tile<8192><<<SM * 2, 512, P * 4>>>(in, ns, out, cursor, P, cap); });
            CK(cudaGetLastError());
            printf("[scatter] P=%u tile8192 smem-hist     %.3f ms  %.1f Gtuples/s (%.0f GB/s rd+wr)\n", P, msC, ns / msC * 1e-6, ns * 16 / msC * 1e-6);
            CK(cudaFree(out)); CK(cudaFree(cursor));
        }
        CK(cudaFree(in));
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
