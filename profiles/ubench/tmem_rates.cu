// TMEM read / write throughput probe (sm_100a): bytes / clk / SM of tcgen05.ld / tcgen05.st as seen by 4 or 8 warps
// (one or two warps per TMEM lane quarter), the access pattern of the attention softmax groups and the GEMM epilogue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I powerpaint_b200/csrc -o profiles/ubench/tmem_rates profiles/ubench/tmem_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "common.cuh"

#define ITERS 512
using namespace pp;

// KIND 0: ld x32 back to back (wait every load); 1: two loads in flight; 2: st x16; 3: ld x32 + 32 MUFU.EX2 per thread
template <int KIND>
__global__ void __launch_bounds__(320, 1) probe(float* out, long long* cycles, int warps) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot;
    float acc = 0.f;
    long long t0 = 0, t1 = 0;
    if (warp < warps) {
        const uint32_t taddr = base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
        uint32_t a[32], b[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { a[i] = i + threadIdx.x; b[i] = 0; }
        tmem_st16(taddr, reinterpret_cast<uint32_t(&)[16]>(a[0]));
        tmem_wait_st();
        __syncwarp();
        t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
            if (KIND == 0) {
                tmem_ld32(taddr + (it & 3) * 32, a);
                tmem_wait_ld();
                acc += __uint_as_float(a[it & 31]);
            } else if (KIND == 1) {
                tmem_ld32(taddr + (it & 1) * 64, a);
                tmem_ld32(taddr + (it & 1) * 64 + 32, b);
                tmem_wait_ld();
                acc += __uint_as_float(a[it & 31]) + __uint_as_float(b[it & 31]);
            } else if (KIND == 2) {
                tmem_st16(taddr + (it & 7) * 16, reinterpret_cast<uint32_t(&)[16]>(a[0]));
                tmem_wait_st();
            } else {
                tmem_ld32(taddr + (it & 3) * 32, a);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(a[i]))); acc += y; }
            }
        }
        t1 = clock64();
    }
    if (acc == 1234.5f) out[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(base, 512); }
}

template <int KIND>
void run(const char* name, int warps, double bytes_per_iter_per_warp) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* cyc;
    cudaMalloc(&out, 4); cudaMalloc(&cyc, sms * 8);
    probe<KIND><<<sms, 320>>>(out, cyc, warps);
    cudaDeviceSynchronize();
    probe<KIND><<<sms, 320>>>(out, cyc, warps);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double c = (double)h[0];
    printf("%-44s warps=%d  %8.1f cyc/iter  %7.1f B/clk/SM  (%s)\n", name, warps, c / ITERS,
           bytes_per_iter_per_warp * warps * ITERS / c, cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {1, 4, 8}) {
        run<0>("tcgen05.ld 32x32b.x32, wait each", w, 4096);
        run<1>("tcgen05.ld 32x32b.x32 x2 in flight", w, 8192);
        run<2>("tcgen05.st 32x32b.x16, wait each", w, 2048);
        run<3>("tcgen05.ld x32 + 32 ex2 per thread", w, 4096);
    }
    return 0;
}
