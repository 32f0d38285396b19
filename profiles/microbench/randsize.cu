// randsize.cu -- request rate of fully divergent 32-byte loads (LDG.E.256) as a function of the table footprint
// (4 MiB .. 16 GiB): separates the L2-resident regime (< ~100 MB), the TLB-reach regime (< 256 MB per the microarch
// notes) and the DRAM regime.  Second part: G consecutive lanes share one sector (G = 1, 2, 4, 8, 32) at 2 GiB, i.e.
// what a layout that keeps neighbouring k-mers of a read in one bucket would buy.  Third part: two-level probe -- every
// lane probes a small table (8 B load), a fraction F of the lanes then probes the 2 GiB table (32 B).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

__device__ __forceinline__ u64 ld256(const u64* p) {
    u64 x0, x1, x2, x3;
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x0), "=l"(x1), "=l"(x2), "=l"(x3) : "l"(p));
    return x0 ^ x1 ^ x2 ^ x3;
}
__device__ __forceinline__ u64 ld64(const u64* p) { u64 x; asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(x) : "l"(p)); return x; }

// G = lanes per shared sector
__global__ void k_size(const u64* a, u64 sectMask, u64 per, int gshift, u64* out) {
    u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    u64 acc = 0;
    const u64 tg = t >> gshift;
#pragma unroll 4
    for (u64 i = 0; i < per; i++) {
        u64 r = mix(tg * per + i + 999);
        acc += ld256(a + (r & sectMask) * 4);
    }
    if (acc == 42) out[0] = acc;
}

// two-level: small table probe by everybody (8 B), big table probe (32 B) by a fraction num/256
__global__ void k_two(const u64* small, u64 smallMask, const u64* big, u64 bigMask, u64 per, unsigned num, u64* out) {
    u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    u64 acc = 0;
#pragma unroll 4
    for (u64 i = 0; i < per; i++) {
        u64 r = mix(t * per + i + 999);
        u64 w = ld64(small + (r & smallMask));
        acc += w;
        if (((r >> 50) & 255) < num + (w & 1)) acc += ld256(big + ((r >> 8) & bigMask) * 4);
    }
    if (acc == 42) out[0] = acc;
}

static float timeit(void (*launch)(void*), void* arg) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(arg);
    cudaEventRecord(e0);
    launch(arg);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms;
}

int main() {
    const int blocks = 148 * 8, threads = 256; const u64 per = 128;
    const u64 maxBytes = 16ULL << 30;
    u64 *a, *out;
    if (cudaMalloc(&a, maxBytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&out, 8); cudaMemset(a, 0, maxBytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("# part 1: footprint sweep, fully divergent 32-byte loads\n");
    for (u64 bytes = 4ULL << 20; bytes <= maxBytes; bytes <<= 1) {
        const u64 mask = bytes / 32 - 1;
        k_size<<<blocks, threads>>>(a, mask, per, 0, out);
        k_size<<<blocks, threads>>>(a, mask, per, 0, out);
        cudaEventRecord(e0);
        k_size<<<blocks, threads>>>(a, mask, per, 0, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("footprint %6llu MiB: %.3f ms, %.2f G probes/s\n", bytes >> 20, ms, (double)blocks * threads * per / ms / 1e6);
    }
    printf("# part 2: G lanes share a sector, 2 GiB footprint\n");
    for (int g = 0; g <= 5; g++) {
        const u64 mask = (2ULL << 30) / 32 - 1;
        k_size<<<blocks, threads>>>(a, mask, per, g, out);
        cudaEventRecord(e0);
        k_size<<<blocks, threads>>>(a, mask, per, g, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("G=%2d: %.3f ms, %.2f G lane-probes/s, %.2f G distinct sectors/s\n", 1 << g, ms, (double)blocks * threads * per / ms / 1e6,
               (double)blocks * threads * per / ms / 1e6 / (1 << g));
    }
    printf("# part 3: small table (8 B probe, everybody) + 2 GiB table (32 B probe, fraction)\n");
    for (u64 sb = 16ULL << 20; sb <= (256ULL << 20); sb <<= 1) {
        for (unsigned num : {0u, 26u, 64u, 128u}) {
            const u64 smask = sb / 8 - 1, bmask = (2ULL << 30) / 32 - 1;
            const u64* big = a + (4ULL << 30) / 8;
            k_two<<<blocks, threads>>>(a, smask, big, bmask, per, num, out);
            k_two<<<blocks, threads>>>(a, smask, big, bmask, per, num, out);
            cudaEventRecord(e0);
            k_two<<<blocks, threads>>>(a, smask, big, bmask, per, num, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("small %4llu MiB, big fraction %.2f: %.3f ms, %.2f G k-mers/s\n", sb >> 20, num / 256.0, ms, (double)blocks * threads * per / ms / 1e6);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
