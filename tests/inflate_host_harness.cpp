// Test infrastructure (not part of the product): compiles the device code of genestrip_b200/csrc/gs_inflate.cu as plain
// C++ -- the CUDA keywords are defined away, threadIdx / blockIdx are globals -- and runs the kernel body thread by thread,
// so that the deflate decoder's logic is checked against zlib on every CPU run of the suite, with ASan/UBSan if wanted.
// The test (tests/test_host_cpu.py) passes -DGS_INFLATE_BODY="<path of the extracted kernel source>".
#include <cstddef>
#include <cstdint>
#include <cstring>
typedef uint32_t u32;
typedef uint64_t u64;
struct dim3_ { unsigned x, y, z; };
static dim3_ threadIdx, blockIdx, blockDim = {128, 1, 1};
#define __device__
#define __global__
#define __shared__ static
#define __forceinline__ inline
#define __launch_bounds__(x)
#define __restrict__
static void __syncthreads() {}
static inline uint32_t __brev(uint32_t v) {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}
struct gs_deflate_block { uint64_t in_off, out_off; uint32_t in_len, out_len, crc32, status; };
#define GS_INFLATE_HOST_HARNESS 1
#include GS_INFLATE_BODY

extern "C" int gs_inflate_harness_run(const uint8_t* comp, uint8_t* text, gs_deflate_block* blocks, uint32_t n) {
    // the prologue fills the shared-memory tables in two phases separated by a barrier: run it twice for all threads
    for (int pass = 0; pass < 2; pass++)
        for (unsigned t = 0; t < blockDim.x; t++) { threadIdx.x = t; blockIdx.x = 0; gs_inflate_blocks_kernel(comp, text, blocks, 0); }
    const unsigned warps = blockDim.x / 32;
    for (uint32_t b = 0; b < n; b++) { blockIdx.x = b / warps; threadIdx.x = (b % warps) * 32; gs_inflate_blocks_kernel(comp, text, blocks, n); }
    return 0;
}
