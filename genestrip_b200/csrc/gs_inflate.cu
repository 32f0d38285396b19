// gs_inflate.cu -- raw DEFLATE (RFC 1951) on the device, for block-gzip (BGZF) FASTQ input.
//
// The reference inflates on its single producer thread (java.util.zip.GZIPInputStream in front of
// C/fastq/AbstractFastqReader.java:224).  A BGZF file is a multi-member gzip file whose members are
// independent deflate streams of at most 64 KB, and a batch of reads holds thousands of them, so the device gives every
// block its own thread: the 32 lanes of a warp decode 32 blocks side by side (canonical Huffman decoding by code length,
// the tables in thread-local memory), and the parallelism that hides the latency of the serial bit stream comes from the
// number of blocks in flight, not from inside a block.  Every block is checked like gzread / GZIPInputStream check it:
// the stream must end exactly at ISIZE bytes and the CRC-32 of the output must match the member's trailer.
// No loop runs longer than the block's input bits plus its output bytes, whatever the input holds.
#include "gs_kernels.cuh"

#include <cuda_runtime.h>

typedef unsigned char u8;
typedef unsigned short u16;

#define GS_INF_THREADS 32
#define GS_INF_MAXBITS 15
#define GS_INF_MAXL 288
#define GS_INF_MAXD 30

namespace {

struct BitReader {
    const u8* in;
    u32 pos, end;
    u64 buf;
    int cnt;  // valid bits in buf; negative = the stream was read past its end
    __device__ __forceinline__ void refill() {
        while (cnt <= 56 && pos < end) { buf |= (u64)in[pos++] << cnt; cnt += 8; }
    }
    __device__ __forceinline__ u32 take(int n) {
        const u32 v = (u32)buf & ((1u << n) - 1u);
        buf >>= n; cnt -= n;
        return v;
    }
};

// canonical Huffman code of n symbols from their code lengths (count[len] = symbols of that length, symbol[] = symbols
// ordered by code): 0 = complete, > 0 = incomplete, < 0 = over-subscribed
__device__ int gs_inf_construct(u16* count, u16* symbol, const u8* length, int n) {
    u16 offs[GS_INF_MAXBITS + 1];
    for (int l = 0; l <= GS_INF_MAXBITS; l++) count[l] = 0;
    for (int s = 0; s < n; s++) count[length[s]]++;
    if (count[0] == n) return 0;
    int left = 1;
    for (int l = 1; l <= GS_INF_MAXBITS; l++) {
        left <<= 1;
        left -= (int)count[l];
        if (left < 0) return left;
    }
    offs[1] = 0;
    for (int l = 1; l < GS_INF_MAXBITS; l++) offs[l + 1] = offs[l] + count[l];
    for (int s = 0; s < n; s++)
        if (length[s] != 0) symbol[offs[length[s]]++] = (u16)s;
    return left;
}

// one symbol: walk the code lengths, one bit each (the caller refilled: at least 48 bits are there unless the input ends)
__device__ __forceinline__ int gs_inf_decode(BitReader& br, const u16* count, const u16* symbol) {
    int code = 0, first = 0, index = 0;
    u64 b = br.buf;
#pragma unroll 1
    for (int len = 1; len <= GS_INF_MAXBITS; len++) {
        code |= (int)(b & 1u);
        b >>= 1;
        const int c = (int)count[len];
        if (code - c < first) {
            br.buf = b; br.cnt -= len;
            return (int)symbol[index + (code - first)];
        }
        index += c; first += c;
        first <<= 1; code <<= 1;
    }
    return -1;
}

}  // namespace

// status written per block: 0 = ok, else the first reason the block is not what its trailer says
#define GS_INF_ERR_STREAM 1u   // malformed deflate stream / input ends early
#define GS_INF_ERR_SIZE 2u     // output does not have exactly ISIZE bytes
#define GS_INF_ERR_CRC 3u      // CRC-32 mismatch

__global__ void __launch_bounds__(GS_INF_THREADS) gs_inflate_blocks_kernel(const u8* __restrict__ comp, u8* __restrict__ text, gs_deflate_block* blocks,
                                                                          u32 nBlocks) {
    __shared__ u16 s_lbase[29], s_dbase[30];
    __shared__ u8 s_lext[29], s_dext[30], s_order[19];
    __shared__ u32 s_crc[256];
    {   // RFC 1951 3.2.5 / 3.2.7 tables and the CRC-32 table (polynomial 0xEDB88320, RFC 1952 8), built once per CTA
        const u16 lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        const u8 lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        const u16 dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        const u8 dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        const u8 order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        if (threadIdx.x < 29) { s_lbase[threadIdx.x] = lbase[threadIdx.x]; s_lext[threadIdx.x] = lext[threadIdx.x]; }
        if (threadIdx.x < 30) { s_dbase[threadIdx.x] = dbase[threadIdx.x]; s_dext[threadIdx.x] = dext[threadIdx.x]; }
        if (threadIdx.x < 19) s_order[threadIdx.x] = order[threadIdx.x];
        for (u32 i = threadIdx.x; i < 256; i += blockDim.x) {
            u32 c = i;
            for (int j = 0; j < 8; j++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            s_crc[i] = c;
        }
    }
    __syncthreads();
    const u32 bi = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= nBlocks) return;
    const gs_deflate_block B = blocks[bi];
    u8* out = text + B.out_off;
    const u32 outLen = B.out_len;
    u32 o = 0;
    BitReader br{comp + B.in_off, 0u, B.in_len, 0ULL, 0};
    u16 lcount[GS_INF_MAXBITS + 1], lsym[GS_INF_MAXL], dcount[GS_INF_MAXBITS + 1], dsym[GS_INF_MAXD];
    u8 lengths[GS_INF_MAXL + GS_INF_MAXD + 2];
    u32 err = 0;
    bool last = false;
    while (!last && !err) {
        br.refill();
        last = br.take(1) != 0;
        const u32 type = br.take(2);
        if (br.cnt < 0) { err = GS_INF_ERR_STREAM; break; }
        if (type == 0) {
            // stored block: to the next byte boundary, LEN / NLEN, then LEN bytes as they are
            br.take(br.cnt & 7);
            br.refill();
            const u32 len = br.take(16), nlen = br.take(16);
            if (br.cnt < 0 || len != (~nlen & 0xFFFFu)) { err = GS_INF_ERR_STREAM; break; }
            br.pos -= (u32)br.cnt >> 3;   // bytes that sit unread in the bit buffer
            br.buf = 0; br.cnt = 0;
            if (br.pos + len > br.end) { err = GS_INF_ERR_STREAM; break; }
            if (o + len > outLen) { err = GS_INF_ERR_SIZE; break; }
            for (u32 i = 0; i < len; i++) out[o + i] = br.in[br.pos + i];
            o += len; br.pos += len;
            continue;
        }
        if (type == 3) { err = GS_INF_ERR_STREAM; break; }
        if (type == 1) {   // fixed codes (RFC 1951 3.2.6)
            for (int s = 0; s < 144; s++) lengths[s] = 8;
            for (int s = 144; s < 256; s++) lengths[s] = 9;
            for (int s = 256; s < 280; s++) lengths[s] = 7;
            for (int s = 280; s < GS_INF_MAXL; s++) lengths[s] = 8;
            gs_inf_construct(lcount, lsym, lengths, GS_INF_MAXL);
            for (int s = 0; s < GS_INF_MAXD; s++) lengths[s] = 5;
            gs_inf_construct(dcount, dsym, lengths, GS_INF_MAXD);
        } else {           // dynamic codes (RFC 1951 3.2.7)
            const int nlen = (int)br.take(5) + 257, ndist = (int)br.take(5) + 1, ncode = (int)br.take(4) + 4;
            if (br.cnt < 0 || nlen > 286 || ndist > GS_INF_MAXD) { err = GS_INF_ERR_STREAM; break; }
            for (int i = 0; i < 19; i++) lengths[i] = 0;
            for (int i = 0; i < ncode; i++) { br.refill(); lengths[s_order[i]] = (u8)br.take(3); }
            if (br.cnt < 0 || gs_inf_construct(lcount, lsym, lengths, 19) != 0) { err = GS_INF_ERR_STREAM; break; }
            int idx = 0;
            while (idx < nlen + ndist) {
                br.refill();
                int sym = gs_inf_decode(br, lcount, lsym);
                if (sym < 0 || br.cnt < 0) { err = GS_INF_ERR_STREAM; break; }
                if (sym < 16) { lengths[idx++] = (u8)sym; continue; }
                int rep, val = 0;
                if (sym == 16) {
                    if (idx == 0) { err = GS_INF_ERR_STREAM; break; }
                    val = lengths[idx - 1];
                    rep = 3 + (int)br.take(2);
                } else if (sym == 17) rep = 3 + (int)br.take(3);
                else rep = 11 + (int)br.take(7);
                if (br.cnt < 0 || idx + rep > nlen + ndist) { err = GS_INF_ERR_STREAM; break; }
                while (rep--) lengths[idx++] = (u8)val;
            }
            if (err) break;
            if (lengths[256] == 0) { err = GS_INF_ERR_STREAM; break; }
            int e = gs_inf_construct(lcount, lsym, lengths, nlen);
            if (e < 0 || (e > 0 && nlen - (int)lcount[0] != 1)) { err = GS_INF_ERR_STREAM; break; }
            e = gs_inf_construct(dcount, dsym, lengths + nlen, ndist);
            if (e < 0 || (e > 0 && ndist - (int)dcount[0] != 1)) { err = GS_INF_ERR_STREAM; break; }
        }
        // literals and matches until the end-of-block symbol
        for (;;) {
            br.refill();
            int sym = gs_inf_decode(br, lcount, lsym);
            if (sym < 0 || br.cnt < 0) { err = GS_INF_ERR_STREAM; break; }
            if (sym < 256) {
                if (o >= outLen) { err = GS_INF_ERR_SIZE; break; }
                out[o++] = (u8)sym;
                continue;
            }
            if (sym == 256) break;
            sym -= 257;
            if (sym >= 29) { err = GS_INF_ERR_STREAM; break; }
            const u32 len = (u32)s_lbase[sym] + br.take(s_lext[sym]);
            const int ds = gs_inf_decode(br, dcount, dsym);
            if (ds < 0 || br.cnt < 0) { err = GS_INF_ERR_STREAM; break; }
            const u32 dist = (u32)s_dbase[ds] + br.take(s_dext[ds]);
            if (br.cnt < 0 || dist > o) { err = GS_INF_ERR_STREAM; break; }
            if (o + len > outLen) { err = GS_INF_ERR_SIZE; break; }
            const u8* from = out + o - dist;
            for (u32 i = 0; i < len; i++) out[o + i] = from[i];   // byte by byte: the ranges overlap when dist < len
            o += len;
        }
    }
    if (!err && o != outLen) err = GS_INF_ERR_SIZE;
    if (!err) {
        u32 crc = 0xFFFFFFFFu;
        for (u32 i = 0; i < outLen; i++) crc = s_crc[(crc ^ out[i]) & 0xFFu] ^ (crc >> 8);
        if ((crc ^ 0xFFFFFFFFu) != B.crc32) err = GS_INF_ERR_CRC;
    }
    blocks[bi].status = err;
}

void gs_launch_inflate_blocks(const uint8_t* comp, uint8_t* text, gs_deflate_block* blocks, uint32_t nBlocks, cudaStream_t st) {
    if (nBlocks == 0) return;
    gs_inflate_blocks_kernel<<<(nBlocks + GS_INF_THREADS - 1) / GS_INF_THREADS, GS_INF_THREADS, 0, st>>>(comp, text, blocks, nBlocks);
}
