// gs_text.cu -- FASTQ record splitting on the GPU (SURVEY.md §8f "parallel feeder": raw-text parse on the device).
//
// A chunk of raw FASTQ text that consists of whole records is copied to the device as it is; three streaming kernels find
// the line ends, and a fourth turns lines 4i..4i+3 into record i (sequence start / end) while validating that the chunk
// is strict 4-line FASTQ under the reference parser's rules (AbstractFastqReader.doReadFastq,
// C/fastq/AbstractFastqReader.java:288-368; BufferedLineReader splits on '\n' only, B/io/BufferedLineReader.java:160-182):
//   * no NUL byte (the reference drops them),
//   * the third line of every record starts with '+' (otherwise the reference treats it as another sequence line),
//   * the quality line is at least as long as the sequence (otherwise the reference keeps reading quality lines),
//   * the number of lines is a multiple of 4 and the chunk ends with '\n'.
// Any violation sets an error flag and the host re-parses with the sequential CPU parser, so results never depend on this
// fast path.  '\r' stays part of the sequence exactly as in the reference (CRLF quirk).
#include "gs_kernels.cuh"

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
        // a last line without '\n' loses its final byte in the reference (BufferedLineReader.nextLine returns the count without
        // the missing terminator and the parser subtracts one): such a chunk is left to the CPU parser
        const u32 lines = acc + ((n > 0 && text[n - 1] != '\n') ? 1u : 0u);
        if (n > 0 && text[n - 1] != '\n') atomicOr(meta + 2, 16u);  // GS_TEXT_ERR_TAIL
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

// record i = lines 4i .. 4i+3 -> recs[i] (header start, sequence start / length, quality start) and lens[i]; validation as
// described in the file header.  recs[nRec].hdr_start = offset just behind the last line (its '\n' included if present).
__global__ void gs_text_records_kernel(const uint8_t* __restrict__ text, u64 n, const u32* __restrict__ lineEnd, u32* meta, gs_fastq_rec* recs, u32* lens,
                                       int k, unsigned long long* totals) {
    if (meta[2]) return;
    const u32 nRec = meta[1];
    unsigned long long kmers = 0, bps = 0;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i <= nRec; i += gridDim.x * blockDim.x) {
        const u32 l0 = 4 * i;
        const u32 hdrStart = i == 0 ? 0u : lineEnd[l0 - 1] + 1;
        if (i == nRec) { recs[i] = gs_fastq_rec{hdrStart, hdrStart, 0u, hdrStart}; break; }
        const u32 hdrEnd = lineEnd[l0], seqEnd = lineEnd[l0 + 1], plusEnd = lineEnd[l0 + 2], qualEnd = lineEnd[l0 + 3];
        const u32 seqStart = hdrEnd + 1, plusStart = seqEnd + 1, qualStart = plusEnd + 1;
        const u32 L = seqEnd - seqStart;
        bool ok = plusStart < plusEnd && text[plusStart] == '+';   // the '+' line must start with '+' (and hold it)
        ok = ok && (qualEnd - qualStart) >= L;                      // enough quality characters on one line
        if (!ok) atomicOr(meta + 2, 8u);                            // GS_TEXT_ERR_RECORD
        recs[i] = gs_fastq_rec{hdrStart, seqStart, L, qualStart};
        lens[i] = L;
        bps += L;
        if ((int)L >= k) kmers += L - k + 1;
    }
    kmers = __reduce_add_sync(0xFFFFFFFFu, (u32)kmers);  // per-thread sums stay far below 2^32 (text chunks are < 4 GiB)
    bps = __reduce_add_sync(0xFFFFFFFFu, (u32)bps);
    if ((threadIdx.x & 31) == 0) { atomicAdd(totals, kmers); atomicAdd(totals + 1, bps); }
}

void gs_launch_text_split(const uint8_t* text, u64 n, u32* blockCounts, u32* lineEnd, u32 lineCap, u32* meta, gs_fastq_rec* recs, u32* lens,
                          int k, unsigned long long* totals, cudaStream_t st) {
    const u32 nBlocks = (u32)((n + GS_TEXT_SEG - 1) / GS_TEXT_SEG);
    if (nBlocks == 0) return;
    gs_text_count_kernel<<<nBlocks, GS_TEXT_THREADS, 0, st>>>(text, n, blockCounts, meta);
    gs_text_scan_kernel<<<1, 1024, 0, st>>>(blockCounts, nBlocks, meta, n, text, lineCap);
    gs_text_fill_kernel<<<nBlocks, GS_TEXT_THREADS, 0, st>>>(text, n, blockCounts, lineEnd, lineCap, meta);
    gs_text_records_kernel<<<148 * 4, 256, 0, st>>>(text, n, lineEnd, meta, recs, lens, k, totals);
}

// ---- read offsets = exclusive prefix sums of the sequence lengths (three passes over blocks of 1024 reads)
#define GS_SCAN_TILE 1024
__global__ void __launch_bounds__(256) gs_scan_tile_sums_kernel(const u32* __restrict__ lens, u32 n, u64* tileSums) {
    __shared__ u64 s_w[8];
    const u32 base = blockIdx.x * GS_SCAN_TILE;
    u64 v = 0;
    for (u32 i = threadIdx.x; i < GS_SCAN_TILE; i += 256) if (base + i < n) v += lens[base + i];
    for (int d = 16; d; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { u64 t = 0; for (int i = 0; i < 8; i++) t += s_w[i]; tileSums[blockIdx.x] = t; }
}
__global__ void gs_scan_top_kernel(u64* tileSums, u32 nTiles) {  // one block: exclusive scan in place (<= a few thousand tiles per chunk)
    __shared__ u64 s_part[1024];
    const u32 per = (nTiles + blockDim.x - 1) / blockDim.x;
    const u32 b0 = threadIdx.x * per, b1 = min(nTiles, b0 + per);
    u64 sum = 0;
    for (u32 b = b0; b < b1; b++) sum += tileSums[b];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) { u64 acc = 0; for (u32 i = 0; i < blockDim.x; i++) { const u64 v = s_part[i]; s_part[i] = acc; acc += v; } }
    __syncthreads();
    u64 acc = s_part[threadIdx.x];
    for (u32 b = b0; b < b1; b++) { const u64 v = tileSums[b]; tileSums[b] = acc; acc += v; }
}
__global__ void __launch_bounds__(256) gs_scan_apply_kernel(const u32* __restrict__ lens, u32 n, const u64* __restrict__ tileSums, u64* offsets) {
    // thread t owns 4 consecutive reads of the tile
    __shared__ u64 s_w[8];
    const u32 base = blockIdx.x * GS_SCAN_TILE + threadIdx.x * 4;
    u32 l[4];
#pragma unroll
    for (int j = 0; j < 4; j++) l[j] = base + j < n ? lens[base + j] : 0u;
    const u64 mine = (u64)l[0] + l[1] + l[2] + l[3];
    u64 incl = mine;
    for (int d = 1; d < 32; d <<= 1) { const u64 v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((threadIdx.x & 31) >= d) incl += v; }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    u64 acc = tileSums[blockIdx.x] + incl - mine;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) acc += s_w[w];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (base + j < n) offsets[base + j] = acc;
        acc += l[j];
        if (base + j + 1 == n) offsets[n] = acc;
    }
}
// ---- bases of all reads back to back (what the label kernel expects): one warp per read, 32 bytes per step
__global__ void __launch_bounds__(256) gs_text_gather_kernel(const uint8_t* __restrict__ text, const gs_fastq_rec* __restrict__ recs, const u64* __restrict__ offsets,
                                                             u32 n, uint8_t* bases) {
    const int lane = threadIdx.x & 31;
    const u32 warpsTotal = (gridDim.x * blockDim.x) >> 5;
    for (u32 r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warpsTotal) {
        const gs_fastq_rec rc = recs[r];
        const uint8_t* src = text + rc.seq_start;
        uint8_t* dst = bases + offsets[r];
        for (u32 i = lane; i < rc.seq_len; i += 32) dst[i] = __ldg(src + i);
    }
}
void gs_launch_text_compact(const uint8_t* text, const gs_fastq_rec* recs, const u32* lens, u32 n, u64* tileSums, u64* offsets, uint8_t* bases, cudaStream_t st) {
    if (n == 0) return;
    const u32 nTiles = (n + GS_SCAN_TILE - 1) / GS_SCAN_TILE;
    gs_scan_tile_sums_kernel<<<nTiles, 256, 0, st>>>(lens, n, tileSums);
    gs_scan_top_kernel<<<1, 1024, 0, st>>>(tileSums, nTiles);
    gs_scan_apply_kernel<<<nTiles, 256, 0, st>>>(lens, n, tileSums, offsets);
    gs_text_gather_kernel<<<148 * 8, 256, 0, st>>>(text, recs, offsets, n, bases);
}
// k-mer offsets of the reads (exclusive prefix sums of max(0, L - k + 1)): where each read's contig runs go (want_runs)
__global__ void gs_text_klens_kernel(const u32* __restrict__ lens, u32 n, int k, u32* klens) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) klens[i] = lens[i] >= (u32)k ? lens[i] - (u32)k + 1u : 0u;
}
void gs_launch_text_kmer_offsets(const u32* lens, u32 n, int k, u32* klens, u64* tileSums, u64* kmerOff, cudaStream_t st) {
    if (n == 0) { cudaMemsetAsync(kmerOff, 0, sizeof(u64), st); return; }
    const u32 nTiles = (n + GS_SCAN_TILE - 1) / GS_SCAN_TILE;
    gs_text_klens_kernel<<<148 * 4, 256, 0, st>>>(lens, n, k, klens);
    gs_scan_tile_sums_kernel<<<nTiles, 256, 0, st>>>(klens, n, tileSums);
    gs_scan_top_kernel<<<1, 1024, 0, st>>>(tileSums, nTiles);
    gs_scan_apply_kernel<<<nTiles, 256, 0, st>>>(klens, n, tileSums, kmerOff);
}
// header offsets of the reads named by the max-contig events of a batch (the host copies the descriptor from its text)
__global__ void gs_text_event_headers_kernel(const gs_maxcontig_event* __restrict__ ev, const u32* __restrict__ nEv, u32 evCap, const gs_fastq_rec* __restrict__ recs,
                                             u64 firstReadNo, u32 n, u32* hdr) {
    const u32 m = min(*nEv, evCap);
    for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < m; e += gridDim.x * blockDim.x) {
        const u64 r = ev[e].read_no - firstReadNo;
        hdr[e] = r < n ? recs[r].hdr_start : 0xFFFFFFFFu;
    }
}
void gs_launch_text_event_headers(const gs_maxcontig_event* ev, const u32* nEv, u32 evCap, const gs_fastq_rec* recs, u64 firstReadNo, u32 n, u32* hdr, cudaStream_t st) {
    gs_text_event_headers_kernel<<<32, 256, 0, st>>>(ev, nEv, evCap, recs, firstReadNo, n, hdr);
}
