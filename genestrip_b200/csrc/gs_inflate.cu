// gs_inflate.cu -- raw DEFLATE (RFC 1951) on the device, for block-gzip (BGZF) FASTQ input.
//
// The reference inflates on its single producer thread (java.util.zip.GZIPInputStream in front of
// C/fastq/AbstractFastqReader.java:224).  A BGZF file is a multi-member gzip file whose members are
// independent deflate streams of at most 64 KB, and a batch of reads holds thousands of them.  A deflate stream is serial
// (every code starts where the previous one ended), so the parallelism comes from the number of blocks in flight: every
// block gets its own warp, whose first lane walks the bit stream (lanes that decode different blocks side by side diverge
// at every symbol and end up serialized -- measured: 32 blocks per warp took 32 x as long as one).  What keeps the walk
// short: a symbol whose code has at most 10 bits (8 for distances) is one shared-memory load from a per-warp look-up
// table indexed by the next bits of the stream, two literals whose codes fit those bits together are one load as well; longer codes need no dependent loads to find their length either -- the
// canonical code's left-justified first-code limits of the 15 code lengths sit in registers and the code length is the
// number of limits the next 15 bits reach (15 independent compares).  Matches are copied eight bytes at a time, runs
// (distance below 8) from a register that holds the period.
// Every block is checked like gzread / GZIPInputStream check it: the stream must end exactly at ISIZE bytes and the
// CRC-32 (all 32 lanes: a slice each by slicing-by-4 from shared tables, combined like zlib's crc32_combine) must match
// the member's trailer.  No loop runs longer than the block's
// input bits plus its output bytes, whatever the input holds.
#include "gs_kernels.cuh"

#include <cuda_runtime.h>

typedef unsigned char u8;
typedef unsigned short u16;

#define GS_INF_WARPS 4        // blocks per CTA: one warp each
#define GS_INF_MAXBITS 15
#define GS_INF_MAXL 288
#define GS_INF_MAXD 30
#define GS_INF_LBITS 10       // look-up tables: literal/length codes up to 10 bits, distance codes up to 8 bits in one load
#define GS_INF_DBITS 8

namespace {

struct BitReader {
    const u8* in;
    u32 pos, end;
    u64 buf;
    int cnt;  // valid bits in buf; negative = the stream was read past its end
    __device__ __forceinline__ void refill() {
        if (pos + 8 <= end) {
            // the next eight bytes from three aligned 32-bit loads (the stream starts at any byte of the buffer); the bits
            // above cnt are the stream's next bits and are ORed in again, unchanged, by the next refill
            const size_t a = (size_t)(in + pos);
            const u32* w = (const u32*)(a & ~(size_t)3);
            const u32 sh = (u32)(a & 3u) * 8u;
            const u32 w0 = w[0], w1 = w[1], w2 = sh ? w[2] : 0u;
            const u64 lo = ((u64)w1 << 32) | w0;
            const u64 v = sh ? (lo >> sh) | ((u64)w2 << (64u - sh)) : lo;
            buf |= v << cnt;
            const int adv = (63 - cnt) >> 3;
            pos += (u32)adv; cnt += adv * 8;
        } else {
            // tail of the input: the speculative bits above cnt must go, the bytes below are added exactly
            buf &= cnt > 0 ? (~0ULL >> (64 - cnt)) : 0ULL;
            while (cnt <= 56 && pos < end) { buf |= (u64)in[pos++] << cnt; cnt += 8; }
        }
    }
    __device__ __forceinline__ u32 take(int n) {
        const u32 v = (u32)buf & ((1u << n) - 1u);
        buf >>= n; cnt -= n;
        return v;
    }
};

// Canonical Huffman code of n symbols from their code lengths: symbol[] = the symbols ordered by code, limit[l] = the
// left-justified (15-bit) first code beyond length l, base[l] = index of length l's first symbol minus its first code.
// 0 = complete, > 0 = incomplete, < 0 = over-subscribed; coded = symbols that have a code.  limit[] is indexed by
// constants only (registers).
// tab (or nullptr): direct look-up by the next tabBits bits of the stream: symbol | code length << 9 | (symbol < 256) << 15,
// 0 = the code is longer; pairs (or nullptr): see below.
__device__ __forceinline__ int gs_inf_construct(u32 (&limit)[GS_INF_MAXBITS + 1], int* base, u16* symbol, const u8* length, int n, int& coded,
                                                u16* tab = nullptr, int tabBits = 0, u32* pairs = nullptr) {
    u16 count[GS_INF_MAXBITS + 1], offs[GS_INF_MAXBITS + 1];
    for (int l = 0; l <= GS_INF_MAXBITS; l++) count[l] = 0;
    for (int s = 0; s < n; s++) count[length[s]]++;
    int left = 1;
    u32 first = 0, index = 0;
    bool over = false;
#pragma unroll
    for (int l = 1; l <= GS_INF_MAXBITS; l++) {
        const u32 c = count[l];
        left = (left << 1) - (int)c;
        over |= left < 0;
        offs[l] = (u16)index;
        base[l] = (int)index - (int)first;
        index += c;
        first += c;
        limit[l] = over ? 0u : first << (GS_INF_MAXBITS - l);
        first <<= 1;
    }
    coded = n - (int)count[0];
    if (over) return -1;
    for (int s = 0; s < n; s++)
        if (length[s] != 0) symbol[offs[length[s]]++] = (u16)s;
    if (tab) {
        // codes are packed first bit first into a stream that is read lowest bit first: entry index = the code's bits
        // reversed, repeated for every value of the bits that follow it
        for (int i = 0; i < (1 << tabBits); i++) tab[i] = 0;
        u32 code = 0;
        int si = 0;
        for (int l = 1; l <= tabBits; l++) {
            for (u32 c = 0; c < count[l]; c++, si++, code++) {
                const u16 e = (u16)(symbol[si] | (l << 9) | (symbol[si] < 256 ? 0x8000 : 0));   // bit 15: a literal (or a distance symbol: not looked at)
                for (u32 k = __brev(code) >> (32 - l); k < (1u << tabBits); k += 1u << l) tab[k] = e;
            }
            code <<= 1;
        }
        if (pairs) {
            // two literals in one look-up where the table's bits hold both codes (bases are 2-bit codes in FASTQ text):
            // first | second << 8 | bits of both << 16, 0 = no such pair
            for (u32 i = 0; i < (1u << tabBits); i++) {
                const u32 e1 = tab[i];
                u32 p = 0;
                if (e1 & 0x8000u) {
                    const u32 l1 = (e1 >> 9) & 15u;
                    const u32 e2 = tab[i >> l1];
                    const u32 l2 = (e2 >> 9) & 15u;
                    if ((e2 & 0x8000u) && l1 + l2 <= (u32)tabBits) p = (e1 & 0xFFu) | ((e2 & 0xFFu) << 8) | ((l1 + l2) << 16);
                }
                pairs[i] = p;
            }
        }
    }
    return coded == 0 ? 0 : left;
}

// one symbol (the caller refilled: at least 48 bits are there unless the input ends): -1 = no such code
__device__ __forceinline__ int gs_inf_decode(BitReader& br, const u32 (&limit)[GS_INF_MAXBITS + 1], const int* base, const u16* symbol) {
    const u32 code = __brev((u32)br.buf) >> 17;   // the next 15 bits, first bit on top
    int len = 1;
#pragma unroll
    for (int l = 1; l <= GS_INF_MAXBITS; l++) len += code >= limit[l];
    if (len > GS_INF_MAXBITS) return -1;
    br.buf >>= len; br.cnt -= len;
    return (int)symbol[base[len] + (int)(code >> (GS_INF_MAXBITS - len))];
}

// the same through the look-up table: one shared-memory load for codes of at most tabBits bits
__device__ __forceinline__ int gs_inf_decode_tab(BitReader& br, const u16* tab, u32 tabMask, const u32 (&limit)[GS_INF_MAXBITS + 1], const int* base,
                                                 const u16* symbol) {
    const u32 e = tab[(u32)br.buf & tabMask];
    if (e == 0) return gs_inf_decode(br, limit, base, symbol);
    const int len = (int)(e >> 9) & 15;
    br.buf >>= len; br.cnt -= len;
    return (int)(e & 0x1FFu);
}

// ---- CRC-32 (RFC 1952 8) of a block by the whole warp: every lane takes a slice, the slices' CRCs are combined with
// crc(A || B) = crc(A) * x^(8 |B|) mod P  xor  crc(B)  (polynomials in the reflected representation, as zlib's crc32_combine)
__device__ __forceinline__ u32 gs_crc_bytes(const u8* p, u32 n, const u32 (*T)[256]) {
    u32 crc = 0xFFFFFFFFu, i = 0;
    for (; i < n && ((size_t)(p + i) & 3u); i++) crc = T[0][(crc ^ p[i]) & 0xFFu] ^ (crc >> 8);
    for (; i + 4 <= n; i += 4) {   // slicing by four
        crc ^= *(const u32*)(p + i);
        crc = T[3][crc & 0xFFu] ^ T[2][(crc >> 8) & 0xFFu] ^ T[1][(crc >> 16) & 0xFFu] ^ T[0][crc >> 24];
    }
    for (; i < n; i++) crc = T[0][(crc ^ p[i]) & 0xFFu] ^ (crc >> 8);
    return crc ^ 0xFFFFFFFFu;
}
__device__ __forceinline__ u32 gs_crc_mul(u32 a, u32 b) {   // a * b mod P
    u32 p = 0;
    for (u32 m = 1u << 31; m; m >>= 1) {
        if (a & m) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}
__device__ __forceinline__ u32 gs_crc_xpow8(u32 n) {        // x^(8 n) mod P
    u32 r = 1u << 31, base = 1u << 23;                       // x^0, x^8
    for (; n; n >>= 1) {
        if (n & 1u) r = gs_crc_mul(r, base);
        base = gs_crc_mul(base, base);
    }
    return r;
}

}  // namespace

// status written per block: 0 = ok, else the first reason the block is not what its trailer says
#define GS_INF_ERR_STREAM 1u   // malformed deflate stream / input ends early
#define GS_INF_ERR_SIZE 2u     // output does not have exactly ISIZE bytes
#define GS_INF_ERR_CRC 3u      // CRC-32 mismatch

__global__ void __launch_bounds__(GS_INF_WARPS * 32) gs_inflate_blocks_kernel(const u8* __restrict__ comp, u8* __restrict__ text, gs_deflate_block* blocks,
                                                                             u32 nBlocks) {
    __shared__ u32 s_ltab2[29], s_dtab2[30];   // base | extra bits << 16
    __shared__ u8 s_order[19];
    __shared__ u32 s_crc[4][256];
    __shared__ u16 s_ltab[GS_INF_WARPS][1 << GS_INF_LBITS], s_dtab[GS_INF_WARPS][1 << GS_INF_DBITS];
    __shared__ u32 s_lpair[GS_INF_WARPS][1 << GS_INF_LBITS];
    {   // RFC 1951 3.2.5 / 3.2.7 tables and the CRC-32 tables (polynomial 0xEDB88320, RFC 1952 8; slicing by 4), once per CTA
        const u16 lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        const u8 lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        const u16 dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        const u8 dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        const u8 order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        if (threadIdx.x < 29) s_ltab2[threadIdx.x] = (u32)lbase[threadIdx.x] | ((u32)lext[threadIdx.x] << 16);
        if (threadIdx.x < 30) s_dtab2[threadIdx.x] = (u32)dbase[threadIdx.x] | ((u32)dext[threadIdx.x] << 16);
        if (threadIdx.x < 19) s_order[threadIdx.x] = order[threadIdx.x];
        for (u32 i = threadIdx.x; i < 256; i += blockDim.x) {
            u32 c = i;
            for (int j = 0; j < 8; j++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            s_crc[0][i] = c;
        }
        __syncthreads();
        for (u32 i = threadIdx.x; i < 256; i += blockDim.x) {
            u32 c = s_crc[0][i];
            for (int t = 1; t < 4; t++) { c = (c >> 8) ^ s_crc[0][c & 0xFFu]; s_crc[t][i] = c; }
        }
    }
    __syncthreads();
    const u32 bi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // one block per warp
    if (bi >= nBlocks) return;
    const u32 lane = threadIdx.x & 31u;
    const gs_deflate_block B = blocks[bi];
    u8* out = text + B.out_off;
    const u32 outLen = B.out_len;
    u16* ltab = s_ltab[threadIdx.x >> 5];
    u16* dtab = s_dtab[threadIdx.x >> 5];
    u32* lpair = s_lpair[threadIdx.x >> 5];
    u32 o = 0;
    BitReader br{comp + B.in_off, 0u, B.in_len, 0ULL, 0};
    u32 llimit[GS_INF_MAXBITS + 1], dlimit[GS_INF_MAXBITS + 1];   // registers
    int lidx[GS_INF_MAXBITS + 1], didx[GS_INF_MAXBITS + 1];
    u16 lsym[GS_INF_MAXL], dsym[GS_INF_MAXD];
    u8 lengths[GS_INF_MAXL + GS_INF_MAXD + 2];
    u32 err = 0;
    bool last = lane != 0;   // the bit stream is walked by the warp's first lane; the others join again for the CRC
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
        int coded = 0;
        if (type == 1) {   // fixed codes (RFC 1951 3.2.6)
            for (int s = 0; s < 144; s++) lengths[s] = 8;
            for (int s = 144; s < 256; s++) lengths[s] = 9;
            for (int s = 256; s < 280; s++) lengths[s] = 7;
            for (int s = 280; s < GS_INF_MAXL; s++) lengths[s] = 8;
            gs_inf_construct(llimit, lidx, lsym, lengths, GS_INF_MAXL, coded, ltab, GS_INF_LBITS, lpair);
            for (int s = 0; s < GS_INF_MAXD; s++) lengths[s] = 5;
            gs_inf_construct(dlimit, didx, dsym, lengths, GS_INF_MAXD, coded, dtab, GS_INF_DBITS);
        } else {           // dynamic codes (RFC 1951 3.2.7)
            const int nlen = (int)br.take(5) + 257, ndist = (int)br.take(5) + 1, ncode = (int)br.take(4) + 4;
            if (br.cnt < 0 || nlen > 286 || ndist > GS_INF_MAXD) { err = GS_INF_ERR_STREAM; break; }
            for (int i = 0; i < 19; i++) lengths[i] = 0;
            for (int i = 0; i < ncode; i++) { br.refill(); lengths[s_order[i]] = (u8)br.take(3); }
            // the code-length code lives in the distance code's tables until those are built
            if (br.cnt < 0 || gs_inf_construct(dlimit, didx, dsym, lengths, 19, coded) != 0) { err = GS_INF_ERR_STREAM; break; }
            int idx = 0;
            while (idx < nlen + ndist) {
                br.refill();
                const int sym = gs_inf_decode(br, dlimit, didx, dsym);
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
            int e = gs_inf_construct(llimit, lidx, lsym, lengths, nlen, coded, ltab, GS_INF_LBITS, lpair);
            if (e < 0 || (e > 0 && coded != 1)) { err = GS_INF_ERR_STREAM; break; }
            e = gs_inf_construct(dlimit, didx, dsym, lengths + nlen, ndist, coded, dtab, GS_INF_DBITS);
            if (e < 0 || (e > 0 && coded != 1)) { err = GS_INF_ERR_STREAM; break; }
        }
        // literals and matches until the end-of-block symbol
        for (;;) {
            if (br.cnt < 48) br.refill();   // a length/distance pair takes at most 15 + 5 + 15 + 13 = 48 bits
            // literals whose code the table holds take the short way round: one load, one store, one shift.  (Bits read
            // past the end of the input are zeros and are noticed below or at the end of the block: cnt < 0.)
            u32 pe = lpair[(u32)br.buf & ((1u << GS_INF_LBITS) - 1u)];
            while (pe != 0 && o + 2 <= outLen) {   // two literals at once
                out[o] = (u8)pe; out[o + 1] = (u8)(pe >> 8);
                o += 2;
                const int l = (int)(pe >> 16);
                br.buf >>= l; br.cnt -= l;
                if (br.cnt < 48) br.refill();
                pe = lpair[(u32)br.buf & ((1u << GS_INF_LBITS) - 1u)];
            }
            u32 e = ltab[(u32)br.buf & ((1u << GS_INF_LBITS) - 1u)];
            if ((e & 0x8000u) && o < outLen) {     // one literal, then pairs again
                out[o++] = (u8)e;
                const int l = (int)(e >> 9) & 15;
                br.buf >>= l; br.cnt -= l;
                continue;
            }
            int sym = gs_inf_decode_tab(br, ltab, (1u << GS_INF_LBITS) - 1u, llimit, lidx, lsym);
            if (sym < 0 || br.cnt < 0) { err = GS_INF_ERR_STREAM; break; }
            if (sym < 256) {
                if (o >= outLen) { err = GS_INF_ERR_SIZE; break; }
                out[o++] = (u8)sym;
                continue;
            }
            if (sym == 256) break;
            sym -= 257;
            if (sym >= 29) { err = GS_INF_ERR_STREAM; break; }
            const u32 lt = s_ltab2[sym];
            const u32 len = (lt & 0xFFFFu) + br.take((int)(lt >> 16));
            const int ds = gs_inf_decode_tab(br, dtab, (1u << GS_INF_DBITS) - 1u, dlimit, didx, dsym);
            if (ds < 0 || br.cnt < 0) { err = GS_INF_ERR_STREAM; break; }
            const u32 dt = s_dtab2[ds];
            const u32 dist = (dt & 0xFFFFu) + br.take((int)(dt >> 16));
            if (br.cnt < 0 || dist > o) { err = GS_INF_ERR_STREAM; break; }
            if (o + len > outLen) { err = GS_INF_ERR_SIZE; break; }
            const u8* from = out + o - dist;
            if (dist >= 8) {
                // eight bytes at a time: source and destination of a group do not overlap, so the loads are independent of
                // the stores (a byte-by-byte copy waits for a full load latency per byte: the compiler must keep the order).
                // While there is room, whole groups are written: the bytes beyond the match are overwritten by what follows.
                const u32 groups = (len + 7u) & ~7u;
                if (o + groups <= outLen) {
                    for (u32 i = 0; i < groups; i += 8) {
                        u8 t[8];
#pragma unroll
                        for (u32 j = 0; j < 8; j++) t[j] = from[i + j];
#pragma unroll
                        for (u32 j = 0; j < 8; j++) out[o + i + j] = t[j];
                    }
                } else {
                    for (u32 i = 0; i < len; i += 8) {
                        const u32 n = len - i;
                        u64 t = 0;
#pragma unroll
                        for (u32 j = 0; j < 8; j++) if (j < n) t |= (u64)from[i + j] << (8 * j);
#pragma unroll
                        for (u32 j = 0; j < 8; j++) if (j < n) out[o + i + j] = (u8)(t >> (8 * j));
                    }
                }
            } else {
                // the match overlaps its own output (runs, short periods): the period goes into a register once
                u64 pat = 0;
#pragma unroll
                for (u32 j = 0; j < 7; j++) if (j < dist) pat |= (u64)from[j] << (8 * j);
                u32 j = 0;
                for (u32 i = 0; i < len; i++) {
                    out[o + i] = (u8)(pat >> (8 * j));
                    j = j + 1 == dist ? 0 : j + 1;
                }
            }
            o += len;
        }
    }
    if (lane == 0 && !err && o != outLen) err = GS_INF_ERR_SIZE;
    // ---- CRC-32 of the output: a slice per lane, combined in slice order
    const u32 slice = (((outLen + 31u) / 32u) + 3u) & ~3u;
    u32 acc = 0;
#ifdef GS_INFLATE_HOST_HARNESS
    if (!err) {   // (the host harness runs one lane: it walks the 32 slices itself)
        const u32 mSlice = gs_crc_xpow8(slice);
        for (u32 i = 0; i < 32; i++) {
            const u32 start = i * slice < outLen ? i * slice : outLen, n = outLen - start < slice ? outLen - start : slice;
            const u32 ci = gs_crc_bytes(out + start, n, s_crc);
            if (n) acc = i == 0 ? ci : gs_crc_mul(n == slice ? mSlice : gs_crc_xpow8(n), acc) ^ ci;
        }
    }
#else
    __syncwarp();
    err = __shfl_sync(0xFFFFFFFFu, err, 0);
    if (!err) {
        const u32 myStart = lane * slice < outLen ? lane * slice : outLen, myN = outLen - myStart < slice ? outLen - myStart : slice;
        const u32 c = gs_crc_bytes(out + myStart, myN, s_crc);
        const u32 mSlice = gs_crc_xpow8(slice);
        for (u32 i = 0; i < 32; i++) {
            const u32 start = i * slice < outLen ? i * slice : outLen, n = outLen - start < slice ? outLen - start : slice;
            const u32 ci = __shfl_sync(0xFFFFFFFFu, c, (int)i);
            if (n) acc = i == 0 ? ci : gs_crc_mul(n == slice ? mSlice : gs_crc_xpow8(n), acc) ^ ci;
        }
    }
    if (lane != 0) return;
#endif
    if (!err && acc != B.crc32) err = GS_INF_ERR_CRC;
    blocks[bi].status = err;
}

void gs_launch_inflate_blocks(const uint8_t* comp, uint8_t* text, gs_deflate_block* blocks, uint32_t nBlocks, cudaStream_t st) {
    if (nBlocks == 0) return;
    gs_inflate_blocks_kernel<<<(nBlocks + GS_INF_WARPS - 1) / GS_INF_WARPS, GS_INF_WARPS * 32, 0, st>>>(comp, text, blocks, nBlocks);
}
