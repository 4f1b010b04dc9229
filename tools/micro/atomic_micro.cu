// Same-address atomic micro-benchmark: what does K2's single output cursor cost?
// Every warp claims `chunk` slots from ONE global 64-bit cursor `iters` times (lane 0 issues the atomicAdd), like
// WarpRing::drain. Variants: wait for the result immediately (as K2 does today), consume it one iteration later
// (claim-ahead), or spread the claims over 2..64 cursors. Reports claims/s and ns per claim per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/atomic_micro tools/micro/atomic_micro.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

// MODE 0: result consumed at once; MODE 1: result consumed one iteration later; `work` FMAs between claims stand in for
// the ~2.5 us of streaming/hashing/probing a K2 warp does per drain
template <int MODE>
__global__ void __launch_bounds__(256) k_claims(unsigned long long* cursors, int ncursors, int iters, int work,
                                               unsigned long long* sink) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned long long* cur = cursors + (warp % (uint32_t)ncursors) * 16;  // one cursor per 128-byte line
    unsigned long long acc = 0, pend = 0;
    float f = (float)lane;
    for (int i = 0; i < iters; i++) {
        unsigned long long gb = 0;
        if (lane == 0) gb = atomicAdd(cur, 256ull);
        if (MODE == 0) {
            gb = __shfl_sync(0xffffffffu, gb, 0);
            acc += gb;
        } else {
            acc += __shfl_sync(0xffffffffu, pend, 0);  // last iteration's claim
            pend = gb;
        }
        for (int w = 0; w < work; w++) f = f * 1.0001f + 0.5f;
    }
    if (MODE == 1) acc += __shfl_sync(0xffffffffu, pend, 0);
    if (acc == 0x1234567ull || f == 12345.f) sink[0] = acc;
}

template <int MODE>
static void run(const char* name, unsigned long long* d_cur, unsigned long long* d_sink, int ctas, int ncursors, int iters,
                int work) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaMemset(d_cur, 0, 64 * 128));
    k_claims<MODE><<<ctas, 256>>>(d_cur, ncursors, 10, work, d_sink);
    CK(cudaEventRecord(e0));
    k_claims<MODE><<<ctas, 256>>>(d_cur, ncursors, iters, work, d_sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double claims = (double)ctas * 8 * iters;
    printf("%-28s ctas=%4d cursors=%2d work=%5d: %8.3f ms  %8.1f M claims/s  %7.1f ns per claim per warp\n", name, ctas,
           ncursors, work, ms, claims / ms / 1e3, ms * 1e6 / iters);
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    unsigned long long *d_cur, *d_sink;
    CK(cudaMalloc(&d_cur, 64 * 128));
    CK(cudaMalloc(&d_sink, 8));
    const int iters = 2000;
    for (int per_sm : {1, 4, 8}) {
        for (int work : {0, 1000, 4000}) {
            run<0>("wait for the result", d_cur, d_sink, sms * per_sm, 1, iters, work);
            run<1>("consume one iteration later", d_cur, d_sink, sms * per_sm, 1, iters, work);
        }
    }
    for (int nc : {2, 8, 64}) run<0>("wait, several cursors", d_cur, d_sink, sms * 4, nc, iters, 0);
    return 0;
}
