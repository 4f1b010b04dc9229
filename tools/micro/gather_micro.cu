// gather_micro.cu -- can random filter probes go faster than one divergent LSU access per SM per clock?
// Measures, on an L2-resident table (default 64 MiB), the rate of random 4-byte probes issued through
//   lsu      : per-lane ld.global.cg (the K2 probe path)                      -> the 1 access / clk / SM baseline
//   lsu_l1   : same loads on a table that fits L1 (64 KB): is the bound the tag stage or the miss path?
//   ldgsts   : cp.async.ca.shared.global 4 B (LDGSTS): global -> shared without a register round trip
//   tma      : per-lane cp.async.bulk of 16 B (UBLKCP) completing on one mbarrier per warp batch
//   mixed    : half of every batch through the LSU, half through TMA (do the two paths add up?)
// Not part of the product path. build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/gather_micro
//   tools/micro/gather_micro.cu ; run: build/gather_micro [log2_table_bytes=26] [log2_probes=29]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t mixh(uint32_t key) {
    const uint32_t n = 0x5052acdbu;
    uint32_t h = 4u, k = 4u + 42u + n;
    uint64_t p = (uint64_t)key * n;
    h ^= (uint32_t)p; k ^= (uint32_t)(p >> 32);
    p = (uint64_t)(h ^ (k + n)) * n;
    h ^= (uint32_t)p; k ^= (uint32_t)(p >> 32);
    return k ^ h;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk16(void* dst, const void* src, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(smem_u32(bar)) : "memory");
}

constexpr int WARPS = 8;

// MODE 0 lsu, 1 ldgsts, 2 tma, 3 mixed (ILP/2 lsu + ILP/2 tma)
template <int MODE, int ILP>
__global__ void __launch_bounds__(WARPS * 32) k_gather(const uint32_t* __restrict__ f, uint32_t wmask, uint64_t n,
                                                       unsigned long long* out) {
    __shared__ __align__(16) uint32_t stage[WARPS][ILP][32][4];  // 16 B per lane and slot
    __shared__ __align__(8) uint64_t bars[WARPS];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    if (lane == 0) mbar_init(&bars[wid], 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t cnt = 0, phase = 0;
    for (uint64_t i = tid; i < n; i += stride * ILP) {  // n is a multiple of stride * ILP: no tail
        uint32_t h[ILP], w[ILP];
#pragma unroll
        for (int j = 0; j < ILP; j++) h[j] = mixh((uint32_t)(i + (uint64_t)j * stride) + 128000001u);
        constexpr int NT = MODE == 2 ? ILP : (MODE == 3 ? ILP / 2 : 0);  // probes through TMA
        if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < ILP; j++) {
                const uint32_t* src = f + ((h[j] >> 5) & wmask);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&stage[wid][j][lane][0])), "l"(src)
                             : "memory");
            }
            asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int j = 0; j < ILP; j++) w[j] = stage[wid][j][lane][0];
        } else {
            if (NT > 0) {
                if (lane == 0) mbar_expect(&bars[wid], 32u * NT * 16u);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < NT; j++) {
                    const uint32_t widx = (h[j] >> 5) & wmask;
                    bulk16(&stage[wid][j][lane][0], f + (widx & ~3u), &bars[wid]);
                }
            }
#pragma unroll
            for (int j = NT; j < ILP; j++) {
                const uint32_t* src = f + ((h[j] >> 5) & wmask);
                asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(w[j]) : "l"(src));
            }
            if (NT > 0) {
                mbar_wait(&bars[wid], phase);
                phase ^= 1u;
#pragma unroll
                for (int j = 0; j < NT; j++) w[j] = stage[wid][j][lane][(h[j] >> 5) & 3u];
                __syncwarp();  // every lane has read its slots before the next batch overwrites them
            }
        }
#pragma unroll
        for (int j = 0; j < ILP; j++) cnt += (w[j] >> (h[j] & 31)) & 1u;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) atomicAdd(out, (unsigned long long)cnt);
}

template <int MODE, int ILP>
static void run(const char* name, const uint32_t* f, uint64_t table_bytes, uint64_t n, int ctas_per_sm, int sms,
                unsigned long long* d_out) {
    const int grid = sms * ctas_per_sm;
    const uint64_t quantum = (uint64_t)grid * WARPS * 32 * ILP;
    n = n / quantum * quantum;
    const uint32_t wmask = (uint32_t)(table_bytes / 4 - 1);
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float best = 1e30f;
    unsigned long long h_out = 0;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaMemset(d_out, 0, 8));
        CK(cudaEventRecord(a));
        k_gather<MODE, ILP><<<grid, WARPS * 32>>>(f, wmask, n, d_out);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
        CK(cudaMemcpy(&h_out, d_out, 8, cudaMemcpyDeviceToHost));
    }
    printf("%-8s ilp=%d ctas/sm=%d table=%5.1f MiB: %8.3f ms  %7.1f G probes/s  (%.3f per clk per SM at 1.965 GHz)  hits=%llu\n", name,
           ILP, ctas_per_sm, table_bytes / 1048576.0, best, n / best / 1e6, n / (best * 1e-3) / (sms * 1.965e9), h_out);
}

int main(int argc, char** argv) {
    const int lt = argc > 1 ? atoi(argv[1]) : 26;
    const int lp = argc > 2 ? atoi(argv[2]) : 29;
    const uint64_t table_bytes = 1ull << lt, n = 1ull << lp;
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    uint32_t* f;
    CK(cudaMalloc(&f, table_bytes));
    CK(cudaMemset(f, 0x5a, table_bytes));
    unsigned long long* d_out;
    CK(cudaMalloc(&d_out, 8));
    printf("%s, %d SMs, table %llu bytes, %llu probes per run\n", pr.name, sms, (unsigned long long)table_bytes,
           (unsigned long long)n);
    for (int c : {2, 4, 8}) {
        if (c == 2) run<0, 8>("lsu", f, table_bytes, n, c, sms, d_out);
        if (c == 4) run<0, 8>("lsu", f, table_bytes, n, c, sms, d_out);
        if (c == 8) run<0, 8>("lsu", f, table_bytes, n, c, sms, d_out);
    }
    run<0, 8>("lsu_l1", f, 64 << 10, n, 4, sms, d_out);
    run<0, 8>("lsu_1MB", f, 1 << 20, n, 4, sms, d_out);
    run<1, 8>("ldgsts", f, table_bytes, n, 4, sms, d_out);
    run<2, 4>("tma", f, table_bytes, n, 4, sms, d_out);
    run<2, 8>("tma", f, table_bytes, n, 2, sms, d_out);
    run<2, 8>("tma", f, table_bytes, n, 4, sms, d_out);
    run<2, 8>("tma", f, table_bytes, n, 6, sms, d_out);
    run<3, 8>("mixed", f, table_bytes, n, 4, sms, d_out);
    run<3, 8>("mixed", f, table_bytes, n, 6, sms, d_out);
    run<3, 8>("mixed", f, table_bytes, n, 3, sms, d_out);
    return 0;
}
