// randline.cu -- what does one probe-table lookup cost?  Random 128-byte lines of a 2 GiB table:
//  mode 0: one 8-byte load            mode 1: one 16-byte load (header)
//  mode 2: header + dependent 8-byte load from a random OTHER sector of the same line (40 % of the probes)
//  mode 3: header + dependent 8-byte load from the SAME sector (40 %)
//  mode 4: two 16-byte loads of one 32-byte sector, issued together
//  mode 5: like 2 but dependent load for 100 % of the probes
// blocks per SM and loads in flight per thread are parameters.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
template <int MODE, int ILP> __global__ void k(const uint4* a, u64 lineMask, u64 per, u64* out) {
    u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    u64 acc = 0;
    for (u64 i = 0; i < per; i += ILP) {
        uint4 h[ILP]; u64 r[ILP];
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            r[j] = mix(t * per + i + j + 999);
            const uint4* line = a + (r[j] & lineMask) * 8;
            if (MODE == 0) { h[j].x = (unsigned)__ldcg((const u64*)line); h[j].y = h[j].z = h[j].w = 0; }
            else h[j] = __ldcg(line);
        }
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            const uint4* line = a + (r[j] & lineMask) * 8;
            acc += h[j].x + h[j].w;
            if (MODE == 2 || MODE == 5) { if (MODE == 5 || ((r[j] >> 40) % 10) < 4) acc += __ldg((const u64*)line + 2 + ((r[j] >> 50) % 14)); }
            if (MODE == 3) { if (((r[j] >> 40) % 10) < 4) acc += __ldg((const u64*)line + 2 + ((r[j] >> 50) & 1)); }
            if (MODE == 4) { uint4 h2 = __ldcg(line + 1); acc += h2.y; }
        }
    }
    if (acc == 42) out[0] = acc;
}
template <int MODE, int ILP> void run(const uint4* a, u64 lineMask, u64* out, int bps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * bps, threads = 256; const u64 per = 256;
    k<MODE, ILP><<<blocks, threads>>>(a, lineMask, 16, out);
    cudaEventRecord(e0);
    k<MODE, ILP><<<blocks, threads>>>(a, lineMask, per, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("mode %d ilp %d blocks/SM %d: %.3f ms, %.2f G probes/s\n", MODE, ILP, bps, ms, (double)blocks * threads * per / ms / 1e6);
}
int main() {
    const u64 lines = 1ULL << 24;  // 2 GiB
    uint4* a; u64* out; cudaMalloc(&a, lines * 128); cudaMalloc(&out, 8); cudaMemset(a, 1, lines * 128);
    for (int bps : {2, 4, 8}) {
        run<0, 1>(a, lines - 1, out, bps); run<1, 1>(a, lines - 1, out, bps); run<1, 4>(a, lines - 1, out, bps);
        run<2, 1>(a, lines - 1, out, bps); run<2, 4>(a, lines - 1, out, bps); run<3, 4>(a, lines - 1, out, bps);
        run<4, 4>(a, lines - 1, out, bps); run<5, 4>(a, lines - 1, out, bps);
    }
    return 0;
}
