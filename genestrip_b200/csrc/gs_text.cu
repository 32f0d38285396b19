// gs_text.cu -- FASTQ record splitting on the GPU (SURVEY.md §8f "parallel feeder": raw-text parse on the device).
//
// A chunk of raw FASTQ text that consists of whole records is copied to the device as it is; three streaming kernels find
// the line ends, and a fourth turns lines 4i..4i+3 into record i (sequence start / end) while validating that the chunk
// is strict 4-line FASTQ under the reference parser's rules (AbstractFastqReader.doReadFastq,
// C/fastq/AbstractFastqReader.java:288-368; BufferedLineReader splits on '\n' only, B/io/BufferedLineReader.java:160-182):
//   * no NUL byte (the reference drops them),
//   * the third line of every record starts with '+' (otherwise the reference treats it as another sequence line),
//   * the quality line is at least as long as the sequence (otherwise the reference keeps reading quality lines),
//   * the number of lines is a multiple of 4.
// Any violation sets an error flag and the host re-parses with the sequential CPU parser, so results never depend on this
// fast path.  '\r' stays part of the sequence exactly as in the reference (CRLF quirk).
#include "gs_kernels.cuh"

#define GS_TEXT_SEG 16384          // bytes per block segment
#define GS_TEXT_THREADS 256        // 64 bytes per thread

__device__ __forceinline__ u32 gs_eq_mask16(uint4 v, u32 pat) {  // bit i = byte i of the 16-byte vector equals the pattern byte
    u32 m = (((__vcmpeq4(v.x, pat) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu;
    m |= ((((__vcmpeq4(v.y, pat) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu) << 4;
    m |= ((((__vcmpeq4(v.z, pat) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu) << 8;
    m |= ((((__vcmpeq4(v.w, pat) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu) << 12;
    return m;
}

// newline / NUL masks of the 64 bytes a thread owns (text is 16-byte aligned and padded; bytes >= n are masked off)
__device__ __forceinline__ void gs_thread_masks(const uint8_t* text, u64 n, u64 base, u64& nl, u64& nul) {
    nl = 0; nul = 0;
    if (base >= n) return;
    const uint4* p = (const uint4*)(text + base);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint4 v = __ldg(p + q);
        nl |= (u64)gs_eq_mask16(v, 0x0A0A0A0Au) << (16 * q);
        nul |= (u64)gs_eq_mask16(v, 0u) << (16 * q);
    }
    if (n - base < 64) { const u64 keep = (1ULL << (n - base)) - 1; nl &= keep; nul &= keep; }
}

__global__ void __launch_bounds__(GS_TEXT_THREADS) gs_text_count_kernel(const uint8_t* __restrict__ text, u64 n, u32* blockCounts, u32* meta) {
    __shared__ u32 s_cnt[GS_TEXT_THREADS / 32];
    const u64 base = (u64)blockIdx.x * GS_TEXT_SEG + (u64)threadIdx.x * 64;
    u64 nl, nul;
    gs_thread_masks(text, n, base, nl, nul);
    u32 c = __popcll(nl);
    if (nul) atomicOr(meta + 2, 1u);  // GS_TEXT_ERR_NUL
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 t = 0;
        for (int i = 0; i < GS_TEXT_THREADS / 32; i++) t += s_cnt[i];
        blockCounts[blockIdx.x] = t;
    }
}

// exclusive scan of the block counts (one block; the array has a few thousand entries per 100 MB of text)
__global__ void gs_text_scan_kernel(u32* blockCounts, u32 nBlocks, u32* meta, u64 n, const uint8_t* __restrict__ text, u32 lineCap) {
    __shared__ u32 s_part[1024];
    const u32 per = (nBlocks + blockDim.x - 1) / blockDim.x;
    const u32 b0 = threadIdx.x * per, b1 = min(nBlocks, b0 + per);
    u32 sum = 0;
    for (u32 b = b0; b < b1; b++) sum += blockCounts[b];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 acc = 0;
        for (u32 i = 0; i < blockDim.x; i++) { const u32 v = s_part[i]; s_part[i] = acc; acc += v; }
        // the last line may end at the end of the text instead of at a '\n'
        const u32 lines = acc + ((n > 0 && text[n - 1] != '\n') ? 1u : 0u);
        meta[0] = lines;          // number of lines
        meta[1] = lines / 4;      // number of records
        if (lines % 4) atomicOr(meta + 2, 2u);       // GS_TEXT_ERR_LINES
        if (lines > lineCap) atomicOr(meta + 2, 4u); // GS_TEXT_ERR_CAP
    }
    __syncthreads();
    u32 acc = s_part[threadIdx.x];
    for (u32 b = b0; b < b1; b++) { const u32 v = blockCounts[b]; blockCounts[b] = acc; acc += v; }
}

__global__ void __launch_bounds__(GS_TEXT_THREADS) gs_text_fill_kernel(const uint8_t* __restrict__ text, u64 n, const u32* __restrict__ blockOffsets,
                                                                        u32* lineEnd, u32 lineCap, const u32* __restrict__ meta) {
    __shared__ u32 s_warp[GS_TEXT_THREADS / 32];
    if (meta[2] & 4u) return;  // capacity exceeded: nothing may be written
    const u64 base = (u64)blockIdx.x * GS_TEXT_SEG + (u64)threadIdx.x * 64;
    u64 nl, nul;
    gs_thread_masks(text, n, base, nl, nul);
    const u32 c = __popcll(nl);
    // exclusive scan of c over the block
    u32 incl = c;
    for (int d = 1; d < 32; d <<= 1) { const u32 v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((threadIdx.x & 31) >= d) incl += v; }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    u32 warpBase = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) warpBase += s_warp[w];
    u32 slot = blockOffsets[blockIdx.x] + warpBase + incl - c;
    while (nl) {
        const int b = __ffsll((long long)nl) - 1;
        nl &= nl - 1;
        if (slot < lineCap) lineEnd[slot] = (u32)(base + (u64)b);
        slot++;
    }
    // a last line without '\n' ends at n
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0 && n > 0 && text[n - 1] != '\n') {
        const u32 last = meta[0] - 1;
        if (last < lineCap) lineEnd[last] = (u32)n;
    }
}

// record i = lines 4i .. 4i+3 -> starts[i], ends[i] of the sequence; validation as described in the file header
__global__ void gs_text_records_kernel(const uint8_t* __restrict__ text, u64 n, const u32* __restrict__ lineEnd, u32* meta, u64* starts, u64* ends,
                                       u64 textBase, int k, unsigned long long* totals) {
    if (meta[2]) return;
    const u32 nRec = meta[1];
    unsigned long long kmers = 0, bps = 0;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nRec; i += gridDim.x * blockDim.x) {
        const u32 l0 = 4 * i;
        const u32 hdrEnd = lineEnd[l0], seqEnd = lineEnd[l0 + 1], plusEnd = lineEnd[l0 + 2], qualEnd = lineEnd[l0 + 3];
        const u32 seqStart = hdrEnd + 1, plusStart = seqEnd + 1, qualStart = plusEnd + 1;
        const u32 L = seqEnd - seqStart;
        bool ok = plusStart < plusEnd && text[plusStart] == '+';   // the '+' line must start with '+' (and hold it)
        ok = ok && (qualEnd - qualStart) >= L;                      // enough quality characters on one line
        if (!ok) atomicOr(meta + 2, 8u);                            // GS_TEXT_ERR_RECORD
        starts[i] = textBase + seqStart;
        ends[i] = textBase + seqEnd;
        bps += L;
        if ((int)L >= k) kmers += L - k + 1;
    }
    kmers = __reduce_add_sync(0xFFFFFFFFu, (u32)kmers) ;  // per-thread sums stay far below 2^32 (grid-stride over <= 2^30 records)
    bps = __reduce_add_sync(0xFFFFFFFFu, (u32)bps);
    if ((threadIdx.x & 31) == 0) { atomicAdd(totals, kmers); atomicAdd(totals + 1, bps); }
}

void gs_launch_text_split(const uint8_t* text, u64 n, u32* blockCounts, u32* lineEnd, u32 lineCap, u32* meta, u64* starts, u64* ends,
                          u64 textBase, int k, unsigned long long* totals, cudaStream_t st) {
    const u32 nBlocks = (u32)((n + GS_TEXT_SEG - 1) / GS_TEXT_SEG);
    if (nBlocks == 0) return;
    gs_text_count_kernel<<<nBlocks, GS_TEXT_THREADS, 0, st>>>(text, n, blockCounts, meta);
    gs_text_scan_kernel<<<1, 1024, 0, st>>>(blockCounts, nBlocks, meta, n, text, lineCap);
    gs_text_fill_kernel<<<nBlocks, GS_TEXT_THREADS, 0, st>>>(text, n, blockCounts, lineEnd, lineCap, meta);
    gs_text_records_kernel<<<148 * 4, 256, 0, st>>>(text, n, lineEnd, meta, starts, ends, textBase, k, totals);
}
