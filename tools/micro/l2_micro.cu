// L2 residency micro-benchmark: stream S (evict_first / .cs / default) while probing a filter range of
// 16..128 MiB with different cache policies. Decides how many filter range passes K1/K2 need.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t crapwow42(uint32_t key) {
    const uint32_t n = 0x5052acdbu; uint32_t h = 4u, k = 4u + 42u + n;
    uint64_t p = (uint64_t)key * n; h ^= (uint32_t)p; k ^= (uint32_t)(p >> 32);
    p = (uint64_t)(h ^ (k + n)) * n; h ^= (uint32_t)p; k ^= (uint32_t)(p >> 32);
    return k ^ h;
}
enum { S_EVICT_FIRST = 0, S_CS = 1, S_DEFAULT = 2, S_NOALLOC_ONLY = 3, S_DISCARD = 4, S_LU = 5 };
enum { P_DEFAULT = 0, P_EVICT_LAST = 1 };

template <int SM>
__device__ __forceinline__ uint4 lds(const uint4* p, uint64_t pol) {
    uint4 r;
    if (SM == S_EVICT_FIRST)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    else if (SM == S_CS)
        asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (SM == S_NOALLOC_ONLY)
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (SM == S_DISCARD) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    } else if (SM == S_LU)
        asm volatile("ld.global.lu.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else r = *p;
    return r;
}
template <int PM>
__device__ __forceinline__ uint32_t ldp(const uint32_t* p, uint64_t pol) {
    uint32_t v;
    if (PM == P_EVICT_LAST) asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    else v = __ldg(p);
    return v;
}

// probes only keys whose hash falls into [0, range_mask] after masking with full_mask >> shift == 0
template <int SM, int PM, int V>
__global__ void __launch_bounds__(256) k_sp(const uint4* __restrict__ S, uint64_t npairs, const uint32_t* __restrict__ f,
                                           uint32_t full_mask, uint32_t range_shift, unsigned long long* out) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t pol_s, pol_p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_s));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_p));
    uint32_t cnt = 0;
    for (uint64_t i = tid; i < npairs; i += stride * V) {
        uint4 t[V];
#pragma unroll
        for (int j = 0; j < V; j++) { uint64_t idx = i + (uint64_t)j * stride; t[j] = idx < npairs ? lds<SM>(S + idx, pol_s) : make_uint4(0,0,0,0); }
        uint32_t h[2 * V], w[2 * V];
#pragma unroll
        for (int j = 0; j < V; j++) {
            h[2*j] = crapwow42(t[j].x) & full_mask; h[2*j+1] = crapwow42(t[j].z) & full_mask;
            w[2*j] = (h[2*j] >> range_shift) == 0 ? ldp<PM>(f + (h[2*j] >> 5), pol_p) : 0u;
            w[2*j+1] = (h[2*j+1] >> range_shift) == 0 ? ldp<PM>(f + (h[2*j+1] >> 5), pol_p) : 0u;
        }
#pragma unroll
        for (int j = 0; j < 2 * V; j++) cnt += (w[j] >> (h[j] & 31)) & 1u;
        if (SM == S_DISCARD) {
#pragma unroll
            for (int j = 0; j < V; j++) {
                uint64_t idx = i + (uint64_t)j * stride;
                if (idx < npairs && (threadIdx.x & 7) == 0) {
                    const void* line = (const void*)((uintptr_t)(S + idx) & ~(uintptr_t)127);
                    asm volatile("discard.global.L2 [%0], 128;" :: "l"(line) : "memory");
                }
            }
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, (unsigned long long)cnt);
}
__global__ void k_fill(uint2* S, uint64_t n) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = tid; i < n; i += stride) S[i] = make_uint2(128000001u + ((uint32_t)i * 2654435761u >> 2), (uint32_t)i);
}
__global__ void k_fillf(uint32_t* f, uint64_t nw) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = tid; i < nw; i += stride) f[i] = (uint32_t)(i * 0x9e3779b1u) & 0x11111111u;
}
template <typename F> float best_of(int reps, F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float best = 1e30f;
    for (int i = 0; i < reps; i++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = std::min(best, ms); }
    return best;
}
int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0)); const int SMS = pr.multiProcessorCount;
    uint64_t ns = 1ull << 29;  // 4 GiB of tuples
    uint2* S; CK(cudaMalloc(&S, ns * 8)); k_fill<<<SMS * 8, 256>>>(S, ns);
    uint32_t* f; CK(cudaMalloc(&f, 128u << 20)); k_fillf<<<SMS * 8, 256>>>(f, (128u << 20) / 4);
    unsigned long long* out; CK(cudaMalloc(&out, 8)); CK(cudaMemset(out, 0, 8)); CK(cudaDeviceSynchronize());
    // filter is 2^30 bits; active range = 2^lgr bits (probe only keys hashing below it). probes = ns * 2^lgr / 2^30
    for (int lgr : {29, 30}) {
        uint32_t shift = lgr; double frac = double(1ull << lgr) / double(1ull << 30);
#define RUN(SM, PM, V, G, name) { float ms = best_of(3, [&] { k_sp<SM, PM, V><<<SMS * G, 256>>>((const uint4*)S, ns / 2, f, 0x3FFFFFFFu, shift, out); }); CK(cudaGetLastError()); \
        printf("active %3d MiB  %-28s v=%d g=%d  %.3f ms  stream %.0f GB/s  probes %.1f G/s\n", (1 << (lgr - 23)), name, V, G, ms, ns * 8 / ms * 1e-6, ns * frac / ms * 1e-6); }
        RUN(S_EVICT_FIRST, P_DEFAULT, 4, 8, "S evict_first, P default");
        RUN(S_EVICT_FIRST, P_EVICT_LAST, 4, 8, "S evict_first, P evict_last");
        RUN(S_CS, P_DEFAULT, 4, 8, "S .cs, P default");
        RUN(S_CS, P_EVICT_LAST, 4, 8, "S .cs, P evict_last");
        RUN(S_NOALLOC_ONLY, P_EVICT_LAST, 4, 8, "S L1 noalloc, P evict_last");
        RUN(S_DEFAULT, P_DEFAULT, 4, 8, "S default, P default");
        RUN(S_DISCARD, P_DEFAULT, 4, 8, "S evict_first+discard.L2");
        RUN(S_LU, P_DEFAULT, 4, 8, "S ld.lu (last use)");
        RUN(S_EVICT_FIRST, P_EVICT_LAST, 2, 8, "S evict_first, P evict_last");
        RUN(S_EVICT_FIRST, P_EVICT_LAST, 4, 4, "S evict_first, P evict_last");
    }
    // persisting window on the active range + evict_first stream
    int maxPersist = 0; CK(cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, 0));
    CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxPersist));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    for (int lgr : {28, 29}) {
        for (float hr : {1.0f, 0.6f}) {
            cudaStreamAttrValue av = {}; av.accessPolicyWindow.base_ptr = f; av.accessPolicyWindow.num_bytes = (size_t)1 << (lgr - 3);
            av.accessPolicyWindow.hitRatio = hr; av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float best = 1e30f;
            for (int r = 0; r < 4; r++) { cudaEventRecord(a, st); k_sp<S_EVICT_FIRST, P_DEFAULT, 4><<<SMS * 8, 256, 0, st>>>((const uint4*)S, ns / 2, f, 0x3FFFFFFFu, lgr, out);
                cudaEventRecord(b, st); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = std::min(best, ms); }
            double frac = double(1ull << lgr) / double(1ull << 30);
            printf("active %3d MiB  persisting window hr=%.1f            %.3f ms  stream %.0f GB/s  probes %.1f G/s\n", 1 << (lgr - 23), hr, best, ns * 8 / best * 1e-6, ns * frac / best * 1e-6);
        }
    }
    CK(cudaCtxResetPersistingL2Cache());
    printf("done\n");
    return 0;
}
