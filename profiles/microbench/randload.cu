// randload.cu -- microbenchmark: random 8-byte loads over a multi-GB array with different L2 prefetch-size hints
// and cudaLimitMaxL2FetchGranularity settings.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o randload randload.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
template <int MODE> __device__ __forceinline__ u64 ld(const u64* p) {
    u64 v;
    if (MODE == 0) asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 1) asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 2) asm volatile("ld.global.nc.L2::128B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 3) asm volatile("ld.global.nc.L2::256B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 4) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 5) asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 6) asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
template <int MODE> __global__ void k(const u64* a, u64 nmask, u64 per, u64* out) {
    u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    u64 acc = 0;
    for (u64 i = 0; i < per; i++) acc += ld<MODE>(a + (mix(t * per + i + 12345) & nmask));
    if (acc == 42) out[0] = acc;
}
template <int MODE> float run(const u64* a, u64 nmask, u64* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256; const u64 per = 256;
    k<MODE><<<blocks, threads>>>(a, nmask, 16, out);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(a, nmask, per, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double loads = (double)blocks * threads * per;
    printf("mode %d: %.3f ms, %.2f G loads/s\n", MODE, ms, loads / ms / 1e6);
    return ms;
}
int main(int argc, char** argv) {
    int gran = argc > 1 ? atoi(argv[1]) : -1;
    if (gran >= 0) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); printf("set limit %d: %s\n", gran, cudaGetErrorString(e)); }
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit: %zu\n", g);
    const u64 n = 1ULL << 28;  // 2 GiB of u64
    u64 *a, *out; cudaMalloc(&a, n * 8); cudaMalloc(&out, 8); cudaMemset(a, 1, n * 8);
    run<0>(a, n - 1, out); run<1>(a, n - 1, out); run<2>(a, n - 1, out); run<3>(a, n - 1, out);
    run<4>(a, n - 1, out); run<5>(a, n - 1, out); run<6>(a, n - 1, out); run<7>(a, n - 1, out);
    return 0;
}
