// randwide.cu -- request-rate of fully divergent loads by width: 8 B, 16 B (LDG.128), 32 B (LDG.256, sm_100+), and
// 32 B followed by a dependent 8 B load from the same 128-byte line (40 % of the probes).  2 GiB table.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
template <int MODE> __global__ void k(const u64* a, u64 sectMask, u64 per, u64* out) {
    u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    u64 acc = 0;
    for (u64 i = 0; i < per; i++) {
        u64 r = mix(t * per + i + 999);
        const u64* p = a + (r & sectMask) * 4;  // random 32-byte sector
        u64 x0 = 0, x1 = 0, x2 = 0, x3 = 0;
        if (MODE == 0) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(x0) : "l"(p));
        if (MODE == 1) asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(x0), "=l"(x1) : "l"(p));
        if (MODE >= 2) asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x0), "=l"(x1), "=l"(x2), "=l"(x3) : "l"(p));
        acc += x0 ^ x1 ^ x2 ^ x3;
        if (MODE == 3 && ((r >> 40) % 10) < 4) { u64 y; asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(y) : "l"((const u64*)((u64)p ^ 64))); acc += y; }
    }
    if (acc == 42) out[0] = acc;
}
template <int MODE> void run(const u64* a, u64 sectMask, u64* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 4, threads = 256; const u64 per = 256;
    k<MODE><<<blocks, threads>>>(a, sectMask, 16, out);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(a, sectMask, per, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("mode %d: %.3f ms, %.2f G probes/s\n", MODE, ms, (double)blocks * threads * per / ms / 1e6);
}
int main() {
    const u64 sects = 1ULL << 26;  // 2 GiB
    u64 *a, *out; cudaMalloc(&a, sects * 32); cudaMalloc(&out, 8); cudaMemset(a, 1, sects * 32);
    run<0>(a, sects - 1, out); run<1>(a, sects - 1, out); run<2>(a, sects - 1, out); run<3>(a, sects - 1, out);
    return 0;
}
