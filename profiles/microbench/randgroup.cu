// randgroup.cu -- is the divergent-request cap counted per 32-byte SECTOR or per 128-byte LINE request?
// G adjacent lanes share one random 128-byte line of a 2 GiB table and issue ONE 256-bit load each:
//  mode 0: all G lanes read the same sector of the line           (reference: randsize part 2)
//  mode 1: lane i reads sector i % 4 of the line                  (4 distinct sectors of one line per instruction for G >= 4)
//  mode 2: lane i reads a pseudo-random sector of the line        (what a minimizer-addressed line with hashed sector choice does)
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
template <int MODE> __global__ void k(const u64* a, u64 lineMask, u64 per, int G, u64* out) {
    const u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    const u64 grp = t / G;
    u64 acc = 0;
    for (u64 i = 0; i < per; i++) {
        const u64 r = mix(grp * per + i + 999);
        u64 sect = 0;
        if (MODE == 1) sect = t & 3;
        if (MODE == 2) sect = mix(t * per + i + 12345) & 3;
        const u64* p = a + (r & lineMask) * 16 + sect * 4;
        u64 x0, x1, x2, x3;
        asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x0), "=l"(x1), "=l"(x2), "=l"(x3) : "l"(p));
        acc += x0 ^ x1 ^ x2 ^ x3;
    }
    if (acc == 42) out[0] = acc;
}
template <int MODE> void run(const u64* a, u64 lineMask, int G, u64* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 4, threads = 256; const u64 per = 256;
    k<MODE><<<blocks, threads>>>(a, lineMask, 16, G, out);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(a, lineMask, per, G, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double lanes = (double)blocks * threads * per;
    printf("mode %d G=%2d: %.3f ms, %7.2f G lane-probes/s, %6.2f G lines/s\n", MODE, G, ms, lanes / ms / 1e6, lanes / G / ms / 1e6);
}
int main() {
    const u64 lines = 1ULL << 24;  // 2 GiB
    u64 *a, *out; cudaMalloc(&a, lines * 128); cudaMalloc(&out, 8); cudaMemset(a, 1, lines * 128);
    for (int G : {1, 2, 4, 8, 16}) { run<0>(a, lines - 1, G, out); run<1>(a, lines - 1, G, out); run<2>(a, lines - 1, G, out); }
    return 0;
}
