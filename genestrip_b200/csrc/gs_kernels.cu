// gs_kernels.cu -- hand-written sm_100a kernels of the Genestrip read-matching path.
//
// One warp owns one read.  A read is staged tile-wise into shared memory as a packed 2-bit code stream plus a
// validity bit per base (128-bit global loads, ASCII -> 2 bit with SIMD-in-register arithmetic); then every lane
// takes one k-mer position of a 32-position chunk: window validity, forward/reverse-complement k-mer from the
// packed stream (registers only), blocked Bloom probe, bucketed binary search of the sorted k-mer array.  The 32
// labels of a chunk sit in the 32 lanes, so the reference's sequential contig logic
// (C/match/FastqKMerMatcher.java:327-535) becomes ballot/shuffle run-length detection; per-taxon statistics go
// out as atomics at run ends, unique k-mers as test-then-atomicOr on the position bitset, and the per-read
// vote table (distinct taxa in first-occurrence order + counts) feeds the classification at the end of the read.
//
// Why the decomposition is exact (derivation in DESIGN.md "matchRead as a per-position labelling"):
//  * a position is INVALID iff its window holds a non-CGAT byte; consecutive loop iterations of the reference
//    cover exactly these positions, so contigs are maximal runs of equal per-position labels;
//  * the tax-error counter adds one per miss position and one per INVALID *iteration* = one per bad base b with
//    b <= max-1, plus one if there is a bad base in [max, L-1] and base max-1 is fine;
//  * the error gate is monotone, so "gate closed at any time" == "final count exceeds the bound", and a closed
//    gate discards all votes (:474), hence votes are only needed for reads whose gate stays open;
//  * mergeReadTaxidPath (:568-586) is idempotent per node, so only first occurrences of distinct taxa matter.
#include "gs_kernels.cuh"

#include <algorithm>
#include <cstring>
#include <type_traits>

#define FULL 0xFFFFFFFFu

// ---------------------------------------------------------------------------------------------------------
// match
// ---------------------------------------------------------------------------------------------------------
struct WarpTable {
    u32* vi;
    u32* cnt;
    int cap;
};

// add `len` votes for taxon `v`; new taxa are appended in first-occurrence order.  Returns false on overflow.
// `countR1From`: entries with index >= countR1From bump reads1KMer (FastqKMerMatcher.java:434-439).
__device__ __forceinline__ bool gs_table_add(const WarpTable& T, int& nTab, u32 v, u32 len, int lane, long long* r1k, int countR1From) {
    int hit = -1;
    for (int j = lane; j < nTab; j += 32) if (T.vi[j] == v) hit = j;
    bool ok = true;
    if (__ballot_sync(FULL, hit >= 0)) {
        if (hit >= 0) T.cnt[hit] += len;
    } else if (nTab < T.cap) {
        if (lane == 0) {
            T.vi[nTab] = v;
            T.cnt[nTab] = len;
            if (nTab >= countR1From) atomicAdd((u64*)(r1k + v), 1ULL);
        }
        nTab++;
    } else {
        ok = false;
    }
    __syncwarp();
    return ok;
}

__device__ __forceinline__ int gs_table_count(const WarpTable& T, int nTab, int node, int lane) {
    int c = 0;
    for (int j = lane; j < nTab; j += 32) if ((int)T.vi[j] == node) c = (int)T.cnt[j];
    return __reduce_add_sync(FULL, c);
}

// sumCounts (C/tax/SmallTaxTree.java:184-193): votes of node and all its ancestors
__device__ __forceinline__ int gs_sum_counts(const GsDbView& db, const WarpTable& T, int nTab, int node, int lane) {
    int pc = __ldg(db.pre + node);
    int part = 0;
    for (int j = lane; j < nTab; j += 32) {
        int e = (int)T.vi[j];
        if (__ldg(db.pre + e) <= pc && pc <= __ldg(db.last + e)) part += (int)T.cnt[j];
    }
    return __reduce_add_sync(FULL, part);
}

// getLowestCommonAncestor (C/tax/SmallTaxTree.java:263-289); -1 = null
__device__ __forceinline__ int gs_lca(const GsDbView& db, int a, int b) {
    if (a == b) return a;
    if (a < 0 || b < 0) return -1;
    int da = __ldg(db.depth + a), dbb = __ldg(db.depth + b);
    while (da > dbb) { a = __ldg(db.parent + a); da--; }
    while (dbb > da) { b = __ldg(db.parent + b); dbb--; }
    while (a != b) {
        a = __ldg(db.parent + a); b = __ldg(db.parent + b);
        if (a < 0 || b < 0) return -1;
    }
    return a;
}

// Java short ++ with wrap-around, two counters per 32-bit word (KMerUniqueCounterBits.putInlined :134-140)
__device__ __forceinline__ void gs_hit_count_inc(uint16_t* hitCounts, u64 pos) {
    u32* hp = (u32*)hitCounts + (pos >> 1);
    const int sh = (int)(pos & 1) * 16;
    u32 old = *hp, assumed;
    do {
        assumed = old;
        u32 nv = (assumed & ~(0xFFFFu << sh)) | ((((assumed >> sh) + 1u) & 0xFFFFu) << sh);
        old = atomicCAS(hp, assumed, nv);
    } while (old != assumed);
}

// ---- K0: read-start bitmap over the flat base array (bit f = a read starts at flat position f)
__global__ void gs_mark_starts_kernel(const u64* __restrict__ offsets, u32 nReads, u64 off0, u32 lead, u64 flatLen, u32* startBits) {
    const u32 stride = gridDim.x * blockDim.x;
    for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < nReads; r += stride) {
        const u64 f = offsets[r] - off0 + lead;
        if (offsets[r] >= off0 && f < flatLen) atomicOr(startBits + (f >> 5), 1u << (f & 31));
    }
}

// ---- K1: label kernel.  The batch's bases are one flat array (reads back to back); position f gets the label of the k-mer
// that starts there: value index, MISS, INVALID (window holds a non-CGAT byte) or END (window crosses a read boundary or the
// end of the batch: never read by the reduce kernel, and no side effects).  A warp claims segments of GS_SEG_POS positions:
// 64 aligned 128-bit loads stage 1024 bases as a packed 2-bit stream + validity bits; then chunk by chunk every lane takes
// one position: forward / reverse-complement k-mer in registers, minimizer prefilter (L2-resident bit filter, shared by
// neighbouring lanes), one 256-bit probe-table load for the k-mers that pass, seen bit / hit counter for the hits.
// The k-mer and m-mer hash of chunk c+1 are computed while chunk c is finished (the sliding minimum needs them anyway).
// KT = k as a compile-time constant (31, the reference's default: masks and shift counts become immediates) or 0 = run time.
// Instruction diet of round 2 (same results; ncu: 373 -> see profiles/r02 warp instructions per 32 k-mers): the three staged
// streams of a warp sit in ONE shared-memory struct (one base register, immediate offsets), the code stream is kept as 32-bit
// big-endian words (no index swizzle per load), label / mask stores run on incremented pointers, the first-bucket match is
// computed once (the chain into following buckets is a cold call) in both lookup shapes.
struct __align__(16) GsSegStage {
    u32 code[2 * GS_SEG_WORDS];   // 2-bit codes, word j = bases 16j .. 16j+15 of the segment, first base in the top pair
    u32 valid[GS_SEG_WORDS];      // bit b of word w = base 32w + b is one of CGAT
    u32 start[GS_SEG_WORDS];      // bit b of word w = a read starts at base 32w + b
};

// forward k-mer of the window whose first base sits in word cp[0] at bit offset sh = 2 * (position & 15) from the top
__device__ __forceinline__ u64 gs_extract_w(const u32* cp, u32 sh, int k) {
    const u32 w0 = cp[0], w1 = cp[1], w2 = cp[2];
    const u32 x1 = __funnelshift_l(w1, w0, sh), x0 = __funnelshift_l(w2, w1, sh);
    return (((u64)x1 << 32) | x0) >> (64 - 2 * k);
}

template <int LAYOUT, bool DUMP, bool WIDE, int KT>
__global__ void __launch_bounds__(GS_WARPS_PER_BLOCK * 32, WIDE ? GS_LABEL_MIN_BLOCKS_WIDE : GS_LABEL_MIN_BLOCKS) gs_label_kernel(const GsMatchParams P) {
    typedef typename std::conditional<WIDE, u64, u32>::type MzT;  // the minimizer order: 32-bit hash, or 64 bits for large stores
    __shared__ GsSegStage s_stage[GS_WARPS_PER_BLOCK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const GsDbView& db = P.db;
    const int k = KT > 0 ? KT : db.k;
    const u32 kmask = (k >= 32) ? 0xFFFFFFFFu : ((1u << k) - 1u);
    const u32 k1mask = (1u << (k - 1)) - 1u;   // read starts inside (f, f + k - 1] = the window crosses a read boundary
    const bool useBloom = P.useBloom && db.hasBloom;
    const bool mz = LAYOUT == GS_LAYOUT_TABLE && db.mzFilter != nullptr;
    const int m = k - GS_MZ_S;
    const u64 mmask = (1ULL << (2 * (m > 0 ? m : 1))) - 1;
    GsSegStage& S = s_stage[warp];
    const u32* cp = S.code + (lane >> 4);      // this lane's window of chunk c starts in word cp[2c]
    const u32 csh = (u32)(lane & 15) * 2;
    const uint8_t* fb = P.bases + P.off0 - P.lead;  // 16-byte aligned start of the flat array
    const u64 nSeg = (P.flatLen + GS_SEG_POS - 1) / GS_SEG_POS;
    for (;;) {
        u32 seg = 0;
        if (lane == 0) seg = atomicAdd(P.segCounter, 1u);
        seg = __shfl_sync(FULL, seg, 0);
        if (seg >= nSeg) break;
        const u64 f0 = (u64)seg * GS_SEG_POS;
        const int nb = (int)min((u64)GS_SEG_BASES, P.flatLen - f0);
        const u64 w0 = (u64)seg * GS_SEG_CHUNKS;   // first mask word of the segment (validBits / startBits / bmask / packed words)
        __syncwarp();
        // ---- stage: ASCII -> packed 2-bit codes + validity bits (C/util/CGAT.java:60-69)
        if (seg < P.packSegs) {  // the host packed this part of the batch: the two streams arrive ready-made, one coalesced load each
            const u64 cwd = __ldg(P.packCodes + w0 + lane);
            ((uint2*)S.code)[lane] = make_uint2((u32)(cwd >> 32), (u32)cwd);
            S.valid[lane] = __ldg(P.packValid + w0 + lane);
        } else {
            const uint4* ap = (const uint4*)(fb + f0);
#pragma unroll
            for (int j = lane; j < GS_SEG_BASES / 16; j += 32) {
                u32 code = 0, valid = 0;
                const int rem = nb - j * 16;
                if (rem > 0) {
                    const uint4 A = __ldg(ap + j);
                    u32 c0, c1, c2, c3, v0, v1, v2, v3;
                    gs_conv4(A.x, c0, v0); gs_conv4(A.y, c1, v1); gs_conv4(A.z, c2, v2); gs_conv4(A.w, c3, v3);
                    code = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
                    valid = v0 | (v1 << 4) | (v2 << 8) | (v3 << 12);
                    if (rem < 16) valid &= (1u << rem) - 1u;
                    if (seg == 0) {  // alignment bytes in front of the first read are not bases
                        const int lo = (int)P.lead - j * 16;
                        if (lo > 0) valid &= lo >= 16 ? 0u : ~((1u << lo) - 1u);
                    }
                }
                S.code[j] = code;
                ((uint16_t*)S.valid)[j] = (uint16_t)valid;
            }
        }
        S.start[lane] = __ldg(P.startBits + w0 + lane);
        if (lane < 2) { ((uint2*)S.code)[32 + lane] = make_uint2(0u, 0u); S.valid[32 + lane] = 0; S.start[32 + lane] = lane == 0 ? __ldg(P.startBits + w0 + 32) : 0u; }
        __syncwarp();
        if (lane < GS_SEG_CHUNKS) P.validBits[w0 + lane] = S.valid[lane];  // for the reduce kernel's INVALID count

        // ---- chunks
        // only the forward k-mer is carried from chunk to chunk: the reverse complement is needed in full by the lanes that reach
        // the table (computed there) and otherwise only for the m-mer hash (its low 2m bits)
        u64 fwdN = gs_extract_w(cp, csh, k);
        u32 carryLab = 0;  // label of the last position of the previous chunk (run-boundary masks)
        MzT hN = 0, preN = 0;
        if (mz) { const u64 rcm = gs_revcomp(fwdN, k) & mmask; hN = WIDE ? (MzT)gs_mmer_hash2w(fwdN >> (2 * GS_MZ_S), rcm) : (MzT)gs_mmer_hash2(fwdN >> (2 * GS_MZ_S), rcm); preN = gs_seg_prefix_min(hN, lane); }
        u32* labP = P.labels + f0 + lane;                       // this lane's label of chunk c goes to labP[32c]
        int left = (int)min((u64)GS_SEG_POS, P.flatLen - f0) - lane;   // positions of the segment from this lane's on: store while > 0
#pragma unroll 1
        for (int c = 0; c < GS_SEG_CHUNKS; c++) {
            const u64 fwd = fwdN;
            const MzT hC = hN, preC = preN;
            fwdN = gs_extract_w(cp + 2 * c + 2, csh, k);
            const u32 vbits = __funnelshift_r(S.valid[c], S.valid[c + 1], lane);
            const u32 sbits = __funnelshift_rc(S.start[c], S.start[c + 1], lane + 1);
            u32 lab = (sbits & k1mask) ? GS_LABEL_END : ((vbits & kmask) != kmask ? GS_LABEL_INVALID : GS_LABEL_PENDING);
#ifdef GS_EXP_FAKE_LOCAL
            u32 expLine = 0;   // timing experiment only (wrong results): address the table by minimizer, as a line-addressed table would
#endif
            if (mz) {
                const u64 rcm = gs_revcomp(fwdN, k) & mmask;
                hN = WIDE ? (MzT)gs_mmer_hash2w(fwdN >> (2 * GS_MZ_S), rcm) : (MzT)gs_mmer_hash2(fwdN >> (2 * GS_MZ_S), rcm);
                preN = gs_seg_prefix_min(hN, lane);
                const u32 mzv = gs_mz_index((u32)gs_window_min(gs_seg_suffix_min(hC, lane), preC, preN, lane), db.mzMask);
                if (lab == GS_LABEL_PENDING && !((__ldg(db.mzFilter + (mzv >> 6)) >> (mzv & 63)) & 1ULL)) lab = GS_LABEL_MISS;
#ifdef GS_EXP_FAKE_LOCAL
                expLine = mzv;
#endif
            }
            {
                // The lookup.  Two shapes of the same probe (template parameter, same results):
                //  * large stores (WIDE): EVERY lane runs the first-bucket probe -- lanes without a pending k-mer read bucket 0, one
                //    shared L2-resident sector -- so that the warp does not split into a few lanes that wait for DRAM and many
                //    that run ahead into the next chunk (measured on the 2e9-k-mer store, 10 % of the reads from the database:
                //    31 % more warp instructions and 7.70 instead of 6.58 ms with the divergent shape);
                //  * otherwise only the pending lanes probe (fewer instructions when most chunks are either all pending or all
                //    rejected by the prefilter: 6.53 instead of 6.69 ms on the viral workload).
                // Either way the first bucket is matched in straight-line code; the rare walk into following buckets is a call.
                const bool pend = lab == GS_LABEL_PENDING;
                u64 pos = 0;
                bool seen = false;
                if (LAYOUT == GS_LAYOUT_TABLE) {
                    if (WIDE ? __any_sync(FULL, pend) : pend) {
                        const u64 h = gs_mix62(gs_canonical(fwd, k));  // standardKMer (CGAT.java:145-147)
#ifdef GS_EXP_FAKE_LOCAL
                        const u64 b0 = (WIDE && !pend) ? 0ULL : ((((u64)(expLine * 0x85EBCA77u) & ((1ULL << (db.tbits - 2)) - 1)) << 2) | (h & 3));
#else
                        const u64 b0 = (WIDE && !pend) ? 0ULL : (h >> db.rbits);
#endif
                        const GsBucket bk = gs_load_bucket(db.tab, b0);
                        u64 e;
                        const int j = gs_table_match(db.rbits, h, bk, e);
                        const u32 v = (u32)(e >> GS_TAB_VAL_SHIFT) & 0xFFFFu;
                        const bool more = pend && j < 0 && (bk.e[0] & GS_TAB_SPILL);
                        if (pend) lab = (j >= 0 && v != GS_VAL_NONODE) ? v : GS_LABEL_MISS;
                        pos = b0 * GS_TAB_SLOT_STRIDE + (u64)(j >= 0 ? j : 0);
                        seen = e & GS_TAB_SEEN;
                        if (more) {  // rare: the key was pushed to a following bucket
                            const ulonglong2 r = gs_table_chain(db.tab, db.rbits, h, b0);
                            pos = r.x; lab = (u32)r.y; seen = (r.y >> 32) != 0;
                        }
                    }
                } else if (pend) {
                    lab = gs_lookup(db, gs_canonical(fwd, k), useBloom, pos);
                }
                if (pend && lab < GS_LABEL_INVALID) {
                    // unique k-mer bit / hit counter (KMerUniqueCounterBits.putInlined, C/store/KMerUniqueCounterBits.java:117-143)
                    if (P.seenTab) {   // the seen bit came with the bucket
                        if (!seen) {
                            atomicOr(P.seenTab + pos * 2, (u32)GS_TAB_SEEN);
                            if (P.bitset) atomicOr(P.bitset + (pos >> 6), 1ULL << (pos & 63));   // dual mode: compact bitset kept current for the merge
                        }
                    }
                    else if (P.bitset) {
                        const u64 bit = 1ULL << (pos & 63);
                        if (!(*(volatile u64*)(P.bitset + (pos >> 6)) & bit)) atomicOr(P.bitset + (pos >> 6), bit);
                    }
                    if (P.hitCounts) gs_hit_count_inc(P.hitCounts, pos);
                    if (DUMP) P.flatPos[f0 + (u64)(c * 32 + lane)] = (long long)pos;
                }
            }
            __syncwarp();  // reconverge here, not at the compiler's leisure: the shuffles of the next chunk need the full warp
            if (left > 0) *labP = lab;
            labP += 32;
            left -= 32;
            if (P.bmask) {
                // run-boundary mask for the warp-per-read reduce kernel: bit = this position's label differs from its
                // predecessor's.  The predecessor of a segment's first position belongs to another warp: that bit is set
                // unconditionally and verified by the reader.
                u32 pv = __shfl_up_sync(FULL, lab, 1);
                if (lane == 0) pv = carryLab;
                const u32 bm = __ballot_sync(FULL, lab != pv || (c == 0 && lane == 0));
                if (lane == 0) P.bmask[w0 + c] = bm;
                carryLab = __shfl_sync(FULL, lab, 31);
            }
        }
    }
}

// ---- K2: reduce kernel, one warp per read: the reference's sequential contig logic (C/match/FastqKMerMatcher.java:327-535)
// over the read's labels, 32 positions per step.
// MODE 0: fast path (vote table in shared memory).  MODE 1: slow path for reads that overflowed the fast table
// (table in global scratch sized nValues; contig statistics were already applied by the fast path, only reads1KMer
// beyond the first GS_TABLE_CAP taxa and the classification are done here).
// DUMP: additionally write the per-position labels / positions in k-mer order (parity tests).
template <int MODE, bool DUMP, bool MASKED>
__global__ void __launch_bounds__(GS_WARPS_PER_BLOCK * 32, GS_MIN_BLOCKS) gs_reduce_kernel(const GsMatchParams P) {
    __shared__ uint16_t s_bpos[MASKED ? GS_WARPS_PER_BLOCK : 1][MASKED ? 1024 : 1];  // run boundaries of 1024 positions, per warp
    __shared__ u32 s_tabVi[MODE == 0 ? GS_WARPS_PER_BLOCK : 1][MODE == 0 ? GS_TABLE_CAP : 1];
    __shared__ u32 s_tabCnt[MODE == 0 ? GS_WARPS_PER_BLOCK : 1][MODE == 0 ? GS_TABLE_CAP : 1];
    __shared__ u32 s_cand[GS_WARPS_PER_BLOCK][GS_MAX_PATHS];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 gw = blockIdx.x * GS_WARPS_PER_BLOCK + warp, nw = gridDim.x * GS_WARPS_PER_BLOCK;
    const GsDbView& db = P.db;
    const int k = db.k;
    const int V = db.nValues;
    u32* cand = s_cand[warp];
    WarpTable T;
    if (MODE == 0) { T.vi = s_tabVi[warp]; T.cnt = s_tabCnt[warp]; T.cap = GS_TABLE_CAP; }
    else { T.vi = P.slowTable + (size_t)gw * 2 * (size_t)V; T.cnt = T.vi + V; T.cap = V; }
    const u32 nItems = MODE == 0 ? (P.redoList ? *P.redoCount : P.nReads) : *P.overflowCount;

    // Dynamic distribution: a warp claims GS_CLAIM consecutive reads at a time.
    u32 claimBase = 0, claimPos = GS_CLAIM;
    for (;;) {
        u32 item;
        if (MODE == 0) {
            if (claimPos == GS_CLAIM) {
                if (lane == 0) claimBase = atomicAdd(P.workCounter, (u32)GS_CLAIM);
                claimBase = __shfl_sync(FULL, claimBase, 0);
                claimPos = 0;
            }
            item = claimBase + claimPos++;
            if (item >= nItems) { if (claimBase >= nItems) break; claimPos = GS_CLAIM; continue; }
        } else {
            item = gw + claimBase * nw;
            claimBase++;
            if (item >= nItems) break;
        }
        const u32 r = MODE == 0 ? (P.redoList ? P.redoList[item] : item) : P.overflowList[item];
        const u64 start = P.offsets[r];
        const u64 end = P.offsets[r + 1];
        int L = (int)(end - start);
        if (end < start || end - start > 0x7FFFFFF0ULL || start < P.off0 || end - P.off0 + P.lead > P.flatLen) {  // malformed offsets: reported by gs_match_collect
            if (lane == 0 && P.errFlag) atomicOr(P.errFlag, 1u);
            L = 0;
        }
        const int max = L - k + 1;
        const u64 ordinal = P.firstReadNo + r;
        int classV = -1;
        u32 readKmers = 0, flags = 0, taxErr = P.classify ? 0u : 0xFFFFFFFFu;
        if (max <= 0) {
            if (lane == 0 && MODE == 0) P.out[r] = gs_read_result{classV, readKmers, taxErr, flags};
            continue;
        }
        const u64 fs = start - P.off0 + P.lead;  // flat position of the read's first base
        const u32* lp = P.labels + fs;
        int nTab = 0, misses = 0, carryLen = 0;
        bool overflow = false, sawInvalid = false;
        u32 carryLabel = GS_LABEL_MISS;  // lastTaxid = null (FastqKMerMatcher.java:336)
        u64 runCursor = 0;               // want_runs: next free slot of this read's run list

        if (MASKED) {
            // ---- long reads: walk the run BOUNDARIES the label kernel marked (one bit per position) instead of the labels.
            // 32 mask words = 1024 positions per round; their set bits are spread out into a per-warp list, then 32 boundaries
            // at a time are handled like the starts of the label loop below: the run that ends at boundary p has label
            // labels[p - 1] and length p - (previous boundary).
            const u64 fEnd = fs + (u64)max;      // terminator position: flushes the last run
            u64 prevB = fs;
            uint16_t* bp = s_bpos[warp];
            for (u64 w0 = fs >> 5; w0 * 32 <= fEnd; w0 += 32) {
                const u64 w = w0 + lane, p0 = w * 32;
                u32 mw = 0;
                if (p0 <= fEnd && p0 + 32 > fs) {
                    mw = P.bmask[w];
                    if ((w % GS_SEG_CHUNKS) == 0 && (mw & 1u) && p0 > fs && p0 < fEnd && P.labels[p0] == P.labels[p0 - 1]) mw &= ~1u;  // segment start: verify
                    const int lo = fs > p0 ? (int)(fs - p0) : 0;
                    const int hi = fEnd - p0 >= 32 ? 32 : (int)(fEnd - p0);
                    mw &= (hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
                    if (fs >= p0) mw |= 1u << lo;                                  // the read's first position starts a run
                    if (fEnd >= p0 && fEnd - p0 < 32) mw |= 1u << (int)(fEnd - p0);  // terminator
                }
                const int cnt = __popc(mw);
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += v; }
                const int total = __shfl_sync(FULL, incl, 31);
                int slot = incl - cnt;
                while (mw) { const int b = __ffs(mw) - 1; mw &= mw - 1; bp[slot++] = (uint16_t)(lane * 32 + b); }
                __syncwarp();
                for (int g = 0; g < total; g += 32) {
                    const int idx = g + lane;
                    const bool act = idx < total;
                    const u64 p = w0 * 32 + (u64)bp[act ? idx : total - 1];
                    u64 pp = __shfl_up_sync(FULL, p, 1);
                    if (lane == 0) pp = prevB;
                    const int runLen = act ? (int)(p - pp) : 0;
                    const u32 prev = runLen > 0 ? __ldg(P.labels + p - 1) : GS_LABEL_END;  // label of the run that ends at p
                    if (P.classify) misses += __reduce_add_sync(FULL, prev == GS_LABEL_MISS ? runLen : 0);
                    sawInvalid |= __any_sync(FULL, prev == GS_LABEL_INVALID);
                    const bool flushTax = prev < GS_LABEL_INVALID && runLen > 0;
                    if (flushTax) {  // :396-410 (contig boundary) and :458-471 (final contig)
                        atomicAdd((u64*)(P.counters + 0 * (size_t)V + prev), (u64)runLen);
                        atomicAdd((u64*)(P.counters + 1 * (size_t)V + prev), 1ULL);
                        atomicAdd((u64*)(P.counters + 2 * (size_t)V + prev), (u64)runLen * (u64)runLen);
                        atomicMax(P.maxcontig + prev, ((u64)runLen << GS_MAXCONTIG_SHIFT) | (GS_ORDINAL_MASK - (ordinal & GS_ORDINAL_MASK)));
                    }
                    if (P.runs) {  // printKrakenStyleOut (:597-611): every finished run incl. '0' and 'A'
                        const bool flushAny = runLen > 0;
                        const u32 FA = __ballot_sync(FULL, flushAny);
                        if (flushAny) {
                            u64 slotR = P.runOffsets[r] + runCursor + (u64)__popc(FA & ((1u << lane) - 1u));
                            if (slotR < P.runsCap) P.runs[slotR] = gs_run{prev, (u32)runLen};
                        }
                        runCursor += (u64)__popc(FA);
                    }
                    u32 F = __ballot_sync(FULL, flushTax);
                    while (F) {
                        const int src = __ffs(F) - 1;
                        F &= F - 1;
                        const u32 v = __shfl_sync(FULL, prev, src);
                        const u32 n = (u32)__shfl_sync(FULL, runLen, src);
                        if (!overflow) {
                            if (!gs_table_add(T, nTab, v, n, lane, P.counters + 3 * (size_t)V, 0)) overflow = true;
                        }
                    }
                    prevB = __shfl_sync(FULL, p, min(31, total - g - 1));
                }
                __syncwarp();
            }
        } else {
#pragma unroll 1
        for (int c0 = 0; c0 <= max; c0 += 32) {  // position `max` is the terminator that flushes the last run
            const int p = c0 + lane;
            const u32 lab = p < max ? __ldcs(lp + p) : GS_LABEL_END;
            if (DUMP && p < max) {
                const u64 o = P.kmerOffsets[r] + (u64)p;
                P.dumpLabels[o] = lab == GS_LABEL_INVALID ? -2 : (lab == GS_LABEL_MISS ? -1 : (int)lab);
                P.dumpPos[o] = lab < GS_LABEL_INVALID ? P.flatPos[fs + p] : -1LL;
            }
            if (P.classify) misses += __popc(__ballot_sync(FULL, lab == GS_LABEL_MISS));
            // ---- contigs = maximal runs of equal labels (FastqKMerMatcher.java:370, 390-421)
            u32 prev = __shfl_up_sync(FULL, lab, 1);
            if (lane == 0) prev = carryLabel;
            const bool isStart = lab != prev;
            const u32 S = __ballot_sync(FULL, isStart);
            if (S == 0) { carryLen += 32; continue; }  // the current run covers the whole chunk
            sawInvalid |= __any_sync(FULL, lab == GS_LABEL_INVALID);
            const u32 below = S & ((1u << lane) - 1u);
            int runLen;  // length of the run that ends right before this lane (meaningful if isStart)
            if (lane == 0) runLen = carryLen;
            else if (below) runLen = lane - (31 - __clz(below));
            else runLen = carryLen + lane;
            const bool flushTax = isStart && prev < GS_LABEL_INVALID && runLen > 0;
            if (MODE == 0 && flushTax) {  // :396-410 (contig boundary) and :458-471 (final contig)
                atomicAdd((u64*)(P.counters + 0 * (size_t)V + prev), (u64)runLen);
                atomicAdd((u64*)(P.counters + 1 * (size_t)V + prev), 1ULL);
                atomicAdd((u64*)(P.counters + 2 * (size_t)V + prev), (u64)runLen * (u64)runLen);
                atomicMax(P.maxcontig + prev, ((u64)runLen << GS_MAXCONTIG_SHIFT) | (GS_ORDINAL_MASK - (ordinal & GS_ORDINAL_MASK)));
            }
            if (MODE == 0 && P.runs) {  // printKrakenStyleOut (:597-611): every finished run incl. '0' and 'A'
                const bool flushAny = isStart && runLen > 0;
                const u32 FA = __ballot_sync(FULL, flushAny);
                if (flushAny) {
                    u64 slot = P.runOffsets[r] + runCursor + (u64)__popc(FA & ((1u << lane) - 1u));
                    if (slot < P.runsCap) P.runs[slot] = gs_run{prev, (u32)runLen};
                }
                runCursor += (u64)__popc(FA);
            }
            u32 F = __ballot_sync(FULL, flushTax);
            while (F) {
                const int src = __ffs(F) - 1;
                F &= F - 1;
                const u32 v = __shfl_sync(FULL, prev, src);
                const u32 n = (u32)__shfl_sync(FULL, runLen, src);
                if (!overflow) {
                    if (!gs_table_add(T, nTab, v, n, lane, P.counters + 3 * (size_t)V, MODE == 0 ? 0 : GS_TABLE_CAP)) overflow = true;
                }
            }
            if (S) carryLen = 32 - (31 - __clz(S)); else carryLen += 32;
            carryLabel = __shfl_sync(FULL, lab, 31);
        }
        }
        if (MODE == 0 && P.runs && lane == 0) P.runCounts[r] = (u32)runCursor;

        bool found = nTab > 0;
        if (MODE == 0 && overflow) {
            flags |= GS_READ_SLOWPATH;
            if (lane == 0) { u32 slot = atomicAdd(P.overflowCount, 1u); P.overflowList[slot] = r; }
        }
        if (P.classify) {
            // INVALID iterations (:346-363, 372-373), see file header: one per bad base b <= max-1, plus one if there is a
            // bad base in [max, L-1] and base max-1 is fine.  Validity bits of the batch come from the label kernel.
            int badLow = 0, badTail = 0;
            const u64 bLow = fs + (u64)max, bEnd = fs + (u64)L;  // [fs, bLow) and [bLow, bEnd)
            // (a bad base makes at least one window INVALID, so reads without an INVALID label skip this)
            for (u64 w = (fs >> 5) + lane; sawInvalid && w * 32 < bEnd; w += 32) {
                const u32 inval = ~P.validBits[w];
                const u64 w0 = w * 32;
                // bits of this word inside [a, b): a_rel = clamp(a - w0), b_rel = clamp(b - w0)
                const int a0 = fs > w0 ? (int)(fs - w0) : 0;
                const int lo1 = bLow > w0 ? (int)min((u64)32, bLow - w0) : 0;
                const int e1 = (int)min((u64)32, bEnd - w0);
                const u32 mLow = lo1 > a0 ? (u32)((((1ULL << lo1) - 1) >> a0) << a0) : 0u;
                const int t0 = lo1 > a0 ? lo1 : a0;
                const u32 mTail = e1 > t0 ? (u32)((((1ULL << e1) - 1) >> t0) << t0) : 0u;
                badLow += __popc(inval & mLow);
                badTail |= (inval & mTail) != 0;
            }
            const int bl = __reduce_add_sync(FULL, badLow);
            const bool bt = __any_sync(FULL, badTail);
            const u64 fLast = fs + (u64)max - 1;
            const bool lastOk = !sawInvalid || ((P.validBits[fLast >> 5] >> (fLast & 31)) & 1u);
            const int inv = bl + ((bt && lastOk) ? 1 : 0);
            const int E = inv + misses;
            const double mte = P.maxTaxErr;
            const bool closed = mte >= 0 && ((mte >= 1 && (double)E > mte) || ((double)E > mte * (double)max));  // :374-379
            taxErr = closed ? 0xFFFFFFFFu : (u32)E;
            if (found && !closed && !overflow) {
                int best = 0, ties = 0, node;
                if (nTab == 1 && P.threshold <= 1) {
                    // one taxon in the read: it is the only candidate, its score is its vote count
                    node = (int)T.vi[0]; best = (int)T.cnt[0];
                    if (lane == 0) cand[0] = (u32)node;
                    __syncwarp();
                } else {
                // ---- mergeReadTaxidPath over the distinct taxa in first-occurrence order (:568-586)
                int used = 0;
                for (int j = 0; j < nTab; j++) {
                    const int n = (int)T.vi[j];
                    const int pn = __ldg(db.pre + n), ln = __ldg(db.last + n);
                    int hit = -1;
                    bool repl = false;
                    for (int base = 0; base < used && hit < 0; base += 32) {
                        const int i = base + lane;
                        bool r1 = false, r2 = false;
                        if (i < used) {
                            const int cnode = (int)cand[i];
                            const int pc = __ldg(db.pre + cnode), lc = __ldg(db.last + cnode);
                            r1 = pc <= pn && pn <= lc;  // candidate is ancestor-or-self of node -> replace by node
                            r2 = pn <= pc && pc <= ln;  // node is ancestor-or-self of candidate  -> keep
                        }
                        const u32 bm = __ballot_sync(FULL, r1 || r2);
                        if (bm) { const int f = __ffs(bm) - 1; hit = base + f; repl = __shfl_sync(FULL, (int)r1, f) != 0; }
                    }
                    if (hit >= 0) { if (repl && lane == 0) cand[hit] = (u32)n; }
                    else if (used < P.maxPaths) { if (lane == 0) cand[used] = (u32)n; used++; }
                    __syncwarp();
                }
                // ---- score candidates, keep maxima and ties in order (:474-487)
                for (int i = 0; i < used; i++) {
                    const int cnode = (int)cand[i];
                    const int sum = gs_sum_counts(db, T, nTab, cnode, lane);
                    __syncwarp();
                    if (sum > best) { best = sum; if (lane == 0) cand[0] = (u32)cnode; ties = 0; }
                    else if (sum == best) { ties++; if (lane == 0) cand[ties] = (u32)cnode; }
                    __syncwarp();
                }
                // ---- lowestNodeWhereSumAboveThreshold (:488-492, C/tax/SmallTaxTree.java:208-221)
                if (P.threshold > 1) {
                    for (int i = 0; i <= ties; i++) {
                        int node = (int)cand[i], res = 0, outn = -1;
                        while (node >= 0) {
                            const int cnt = gs_table_count(T, nTab, node, lane);
                            if (cnt > 0) { res += cnt; if (res >= P.threshold) { outn = node; break; } }
                            node = __ldg(db.parent + node);
                        }
                        __syncwarp();
                        if (lane == 0) cand[i] = (u32)outn;
                        __syncwarp();
                    }
                }
                // ---- LCA of the ties (:493-497)
                node = (int)cand[0];
                for (int i = 1; i <= ties; i++) node = gs_lca(db, node, (int)cand[i]);
                }
                classV = node;
                if (node < 0) {
                    found = false;  // `return false` (:498-500)
                } else {
                    const int rk = (ties > 0 || P.threshold > 1) ? gs_sum_counts(db, T, nTab, (int)cand[0], lane) : best;  // :506-507
                    readKmers = (u32)rk;
                    const int classErrC = max - rk;
                    const double mce = P.maxClassErr;
                    if (mce < 0 || (mce >= 1 && (double)classErrC <= mce) || ((double)classErrC <= mce * (double)max)) {  // :509-510
                        flags |= GS_READ_ACCEPTED;
                        if (lane == 0) {  // :518-520 (the four double sums are done by the host in read order)
                            atomicAdd((u64*)(P.counters + 4 * (size_t)V + node), 1ULL);
                            atomicAdd((u64*)(P.counters + 5 * (size_t)V + node), (u64)rk);
                            atomicAdd((u64*)(P.counters + 6 * (size_t)V + node), (u64)L);
                        }
                    }
                }
            }
        }
        if (found) flags |= GS_READ_FOUND;
        if (lane == 0 && !(MODE == 0 && overflow)) {
            if (MODE == 1) flags |= GS_READ_SLOWPATH;
            P.out[r] = gs_read_result{classV, readKmers, taxErr, flags};
        } else if (lane == 0) {
            P.out[r] = gs_read_result{-1, 0u, taxErr, flags | GS_READ_FOUND};
        }
    }
}

// ---- K2T: reduce kernel for short reads, one THREAD per read.  A 150-base read has 120 labels: a warp per read spends
// most of its instructions on per-read set-up and warp-wide bookkeeping, a thread per read just walks its labels.  A warp
// takes 32 consecutive reads; their labels are staged GS_T_POS positions at a time through a [32][GS_T_POS + 1] shared-memory
// tile (row i = read i: coalesced global reads, conflict-free transposed reads), so global traffic is one pass over the labels.
// The contig statistics are accumulated per (read, taxon) in a small per-thread table and applied once at the end of the read
// (sums and maxima: same result as the reference's per-contig updates, FastqKMerMatcher.java:396-410), so a read that turns
// out to need more than GS_T_CAP table entries has had no side effects yet and is handed to the warp-per-read kernel.
#define GS_T_CAP 8
#define GS_T_THREADS 128
#define GS_T_MAX_LEN 2047  // cnt (11 bits) | contigs (10 bits) | maxlen (11 bits) share one word
#ifndef GS_T_POS
#define GS_T_POS 16        // label positions per staging round: tile = [32 reads][GS_T_POS + 1] words, two of them per warp
#endif
#define GS_T_ROW (GS_T_POS + 1)
#define GS_T_BLOCKS (GS_T_POS == 16 ? 7 : 4)   // resident CTAs per SM the shared memory allows (31.2 KB / 47.6 KB per CTA)
#define GS_T_MAXB 16       // MASKED: run boundaries per read handled by the thread (more: warp-per-read kernel)
// MASKED = the label kernel wrote run-boundary masks (P.bmask): a thread then reads the few mask words of its read, collects
// the boundary positions, fetches the labels of the runs that end there with independent loads and never touches the other
// labels -- no staging tile at all.
template <bool MASKED>
__global__ void __launch_bounds__(GS_T_THREADS, MASKED ? 8 : GS_T_BLOCKS) gs_reduce_thread_kernel(const GsMatchParams P) {
    __shared__ u32 s_vi[GS_T_CAP][GS_T_THREADS];
    __shared__ u32 s_pk[GS_T_CAP][GS_T_THREADS];   // cnt << 21 | contigs << 11 | maxlen
    __shared__ u32 s_sq[GS_T_CAP][GS_T_THREADS];
    __shared__ u32 s_tile[GS_T_THREADS / 32][MASKED ? GS_T_CAP * 32 : 2 * 32 * GS_T_ROW];
    __shared__ u64 s_fs[MASKED ? 1 : GS_T_THREADS / 32][MASKED ? 1 : 32];
    __shared__ int s_max[MASKED ? 1 : GS_T_THREADS / 32][MASKED ? 1 : 32];
    __shared__ uint16_t s_bp[MASKED ? GS_T_MAXB : 1][MASKED ? GS_T_THREADS : 1];  // boundary positions relative to the read's first position
    __shared__ u32 s_bl[MASKED ? GS_T_MAXB : 1][MASKED ? GS_T_THREADS : 1];       // label of the run that ends at the boundary
    const GsDbView& db = P.db;
    const int k = db.k, V = db.nValues, t = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32* tile = s_tile[warp];
    int* cand = (int*)tile;  // classification scratch: the tile is idle once the labels are walked
    const u32 nGroups = (P.nReads + 31) / 32;
    for (;;) {
        u32 g = 0;
        if (lane == 0) g = atomicAdd(P.groupCounter, 1u);
        g = __shfl_sync(FULL, g, 0);
        if (g >= nGroups) break;
        const u32 r = g * 32 + lane;
        const bool have = r < P.nReads;
        u64 start = 0, end = 0;
        if (have) { start = P.offsets[r]; end = P.offsets[r + 1]; }
        int L = (int)(end - start);
        if (have && (end < start || end - start > 0x7FFFFFF0ULL || start < P.off0 || end - P.off0 + P.lead > P.flatLen)) {
            if (P.errFlag) atomicOr(P.errFlag, 1u);
            L = 0;
        }
        const int max = L - k + 1;
        int classV = -1;
        u32 readKmers = 0, flags = 0, taxErr = P.classify ? 0u : 0xFFFFFFFFu;
        const u64 fs = start - P.off0 + P.lead;
        bool walk = have && max > 0;
        if (have && max <= 0) P.out[r] = gs_read_result{classV, readKmers, taxErr, flags};
        if (walk && L > GS_T_MAX_LEN) { P.redoList[atomicAdd(P.redoCount, 1u)] = r; walk = false; }
        int nTab = 0, misses = 0, len = 0;
        bool overflow = false, sawInvalid = false;
        u32 prev = GS_LABEL_MISS;  // lastTaxid = null (FastqKMerMatcher.java:336)
        // one contig / miss run / invalid run ends: the body of the reference's run handling (:390-421, 455-473)
        auto endRun = [&](u32 lab, int rl) {
            if (lab < GS_LABEL_INVALID && rl > 0) {
                int j = 0;
                while (j < nTab && s_vi[j][t] != lab) j++;
                if (j < nTab) {
                    const u32 pk = s_pk[j][t];
                    const u32 ml = pk & 0x7FFu;
                    s_pk[j][t] = (((pk >> 21) + (u32)rl) << 21) | ((((pk >> 11) & 0x3FFu) + 1u) << 11) | ((u32)rl > ml ? (u32)rl : ml);
                    s_sq[j][t] += (u32)rl * (u32)rl;
                } else if (nTab < GS_T_CAP) {
                    s_vi[nTab][t] = lab;
                    s_pk[nTab][t] = ((u32)rl << 21) | (1u << 11) | (u32)rl;
                    s_sq[nTab][t] = (u32)rl * (u32)rl;
                    nTab++;
                } else {
                    overflow = true;
                }
            } else if (lab == GS_LABEL_MISS) {
                misses += rl;
            } else if (lab == GS_LABEL_INVALID) {
                sawInvalid = true;
            }
        };
        if (MASKED) {
            if (walk) {
                // ---- boundaries of the read from its mask words (forced at the first position and at the terminator)
                const u64 fEnd = fs + (u64)max;
                int nb = 0;
                for (u64 w = fs >> 5; w * 32 <= fEnd && nb <= GS_T_MAXB; w++) {
                    const u64 p0 = w * 32;
                    u32 mw = __ldg(P.bmask + w);
                    if ((w % GS_SEG_CHUNKS) == 0 && (mw & 1u) && p0 > fs && p0 < fEnd && P.labels[p0] == P.labels[p0 - 1]) mw &= ~1u;  // segment start: verify
                    const int lo = fs > p0 ? (int)(fs - p0) : 0;
                    const int hi = fEnd - p0 >= 32 ? 32 : (int)(fEnd - p0);
                    mw &= (hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
                    if (fs >= p0) mw &= ~(1u << lo);                                  // the read's first position: no run ends there
                    if (fEnd - p0 < 32) mw |= 1u << (int)(fEnd - p0);                 // terminator
                    while (mw && nb <= GS_T_MAXB) {
                        const int bit = __ffs(mw) - 1;
                        mw &= mw - 1;
                        if (nb < GS_T_MAXB) s_bp[nb][t] = (uint16_t)(p0 + (u64)bit - fs);
                        nb++;
                    }
                }
                if (nb > GS_T_MAXB) { overflow = true; }
                else {
                    for (int j = 0; j < nb; j++) s_bl[j][t] = __ldg(P.labels + fs + (u64)s_bp[j][t] - 1);   // independent loads
                    int pb = 0;
                    for (int j = 0; j < nb && !overflow; j++) {
                        const int q = (int)s_bp[j][t];
                        endRun(s_bl[j][t], q - pb);
                        pb = q;
                    }
                }
                if (overflow) { P.redoList[atomicAdd(P.redoCount, 1u)] = r; }
            }
        } else {
        __syncwarp();
        s_fs[warp][lane] = fs;
        s_max[warp][lane] = walk ? max : 0;
        const int wmax = __reduce_max_sync(FULL, walk ? max + 1 : 0);  // + the terminator position that flushes the last run
        __syncwarp();
        // ---- stage positions [b, b + 32) of the warp's 32 reads with 4-byte cp.async (global -> shared without registers),
        // double-buffered: the copies of round b + 32 are in flight while round b is walked
        auto stage = [&](int b, u32* dst) {
            // a cp.async instruction moves GS_T_POS positions of 32 / GS_T_POS reads: lane = (read parity, position)
            const int sub = lane / GS_T_POS, pos = lane % GS_T_POS;
            const int p = b + pos;
#pragma unroll 8
            for (int i0 = 0; i0 < 32; i0 += 32 / GS_T_POS) {
                const int i = i0 + sub;
                if (p < s_max[warp][i]) {
                    const u32 sa = (u32)__cvta_generic_to_shared(dst + i * GS_T_ROW + pos);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sa), "l"(P.labels + s_fs[warp][i] + p) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (wmax > 0) stage(0, tile);
        for (int b = 0; b < wmax; b += GS_T_POS) {
            u32* cur = tile + ((b / GS_T_POS) & 1) * (32 * GS_T_ROW);
            if (b + GS_T_POS < wmax) {
                stage(b + GS_T_POS, tile + (((b / GS_T_POS) & 1) ^ 1) * (32 * GS_T_ROW));
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncwarp();
            if (walk) {
                const int n = min(GS_T_POS, max + 1 - b);  // position `max` is the terminator that flushes the last run
                // run boundaries of this round as a bit mask (all loads first, no branch per label) ...
                u32 m = 0, pl = prev;
#pragma unroll
                for (int i = 0; i < GS_T_POS; i++) {
                    const u32 l = b + i == max ? GS_LABEL_END : cur[lane * GS_T_ROW + i];
                    m |= (u32)(l != pl) << i;
                    pl = l;
                }
                m &= (1u << n) - 1u;
                // ... then one step per boundary
                int cursor = 0;
                while (m) {
                    const int i = __ffs(m) - 1;
                    m &= m - 1;
                    len += i - cursor;
                    cursor = i;
                    if (prev < GS_LABEL_INVALID && len > 0) {  // a contig of taxon `prev` ends (:396-410, 458-471)
                        int j = 0;
                        while (j < nTab && s_vi[j][t] != prev) j++;
                        if (j < nTab && ((s_pk[j][t] >> 11) & 0x3FFu) == 0x3FFu) {
                            // the taxon's contig count would leave its 10-bit field (a taxon alternating with misses in a read of
                            // ~2000 bases): hand the read to the warp kernel, like a table overflow
                            overflow = true;
                            break;
                        } else if (j < nTab) {
                            const u32 pk = s_pk[j][t];
                            const u32 ml = pk & 0x7FFu;
                            s_pk[j][t] = (((pk >> 21) + (u32)len) << 21) | ((((pk >> 11) & 0x3FFu) + 1u) << 11) | ((u32)len > ml ? (u32)len : ml);
                            s_sq[j][t] += (u32)len * (u32)len;
                        } else if (nTab < GS_T_CAP) {
                            s_vi[nTab][t] = prev;
                            s_pk[nTab][t] = ((u32)len << 21) | (1u << 11) | (u32)len;
                            s_sq[nTab][t] = (u32)len * (u32)len;
                            nTab++;
                        } else {
                            overflow = true;
                            break;
                        }
                    } else if (prev == GS_LABEL_MISS) {
                        misses += len;
                    } else if (prev == GS_LABEL_INVALID) {
                        sawInvalid = true;
                    }
                    prev = b + i == max ? GS_LABEL_END : cur[lane * GS_T_ROW + i];
                    len = 0;
                }
                len += n - cursor;
                if (overflow) { P.redoList[atomicAdd(P.redoCount, 1u)] = r; walk = false; }
                else if (b + GS_T_POS > max) walk = false;  // terminator consumed: the read is complete
            }
            __syncwarp();
        }
        }
        if (!have || max <= 0 || overflow || L > GS_T_MAX_LEN) continue;
        // ---- apply the read's contig statistics
        const u64 ordinal = P.firstReadNo + r;
        for (int j = 0; j < nTab; j++) {
            const u32 v = s_vi[j][t], pk = s_pk[j][t];
            atomicAdd((u64*)(P.counters + 0 * (size_t)V + v), (u64)(pk >> 21));
            atomicAdd((u64*)(P.counters + 1 * (size_t)V + v), (u64)((pk >> 11) & 0x3FFu));
            atomicAdd((u64*)(P.counters + 2 * (size_t)V + v), (u64)s_sq[j][t]);
            atomicAdd((u64*)(P.counters + 3 * (size_t)V + v), 1ULL);  // reads1KMer (:434-439)
            atomicMax(P.maxcontig + v, ((u64)(pk & 0x7FFu) << GS_MAXCONTIG_SHIFT) | (GS_ORDINAL_MASK - (ordinal & GS_ORDINAL_MASK)));
            s_pk[j][t] = pk >> 21;  // from here on: the taxon's votes
        }
        bool found = nTab > 0;
        if (P.classify) {
            int inv = 0;
            if (sawInvalid) {  // INVALID iterations (:346-363, 372-373), see the warp kernel
                int badLow = 0;
                bool badTail = false;
                const u64 bLow = fs + (u64)max, bEnd = fs + (u64)L;
                for (u64 w = fs >> 5; w * 32 < bEnd; w++) {
                    const u32 inval = ~P.validBits[w];
                    const u64 w0 = w * 32;
                    const int a0 = fs > w0 ? (int)(fs - w0) : 0;
                    const int lo1 = bLow > w0 ? (int)min((u64)32, bLow - w0) : 0;
                    const int e1 = (int)min((u64)32, bEnd - w0);
                    const u32 mLow = lo1 > a0 ? (u32)((((1ULL << lo1) - 1) >> a0) << a0) : 0u;
                    const int t0 = lo1 > a0 ? lo1 : a0;
                    const u32 mTail = e1 > t0 ? (u32)((((1ULL << e1) - 1) >> t0) << t0) : 0u;
                    badLow += __popc(inval & mLow);
                    badTail |= (inval & mTail) != 0;
                }
                const u64 fLast = fs + (u64)max - 1;
                inv = badLow + ((badTail && ((P.validBits[fLast >> 5] >> (fLast & 31)) & 1u)) ? 1 : 0);
            }
            const int E = inv + misses;
            const double mte = P.maxTaxErr;
            const bool closed = mte >= 0 && ((mte >= 1 && (double)E > mte) || ((double)E > mte * (double)max));  // :374-379
            taxErr = closed ? 0xFFFFFFFFu : (u32)E;
            if (found && !closed) {
                int best = 0, ties = 0, node;
                if (nTab == 1 && P.threshold <= 1) {
                    node = (int)s_vi[0][t]; best = (int)s_pk[0][t];
                    cand[(0) * 32 + lane] = node;
                } else {
                    // ---- mergeReadTaxidPath over the distinct taxa in first-occurrence order (:568-586)
                    int used = 0;
                    for (int j = 0; j < nTab; j++) {
                        const int n = (int)s_vi[j][t];
                        const int pn = __ldg(db.pre + n), ln = __ldg(db.last + n);
                        int hit = -1;
                        bool repl = false;
                        for (int i = 0; i < used; i++) {
                            const int cnode = cand[(i) * 32 + lane];
                            const int pc = __ldg(db.pre + cnode), lc = __ldg(db.last + cnode);
                            const bool r1 = pc <= pn && pn <= lc, r2 = pn <= pc && pc <= ln;
                            if (r1 || r2) { hit = i; repl = r1; break; }
                        }
                        if (hit >= 0) { if (repl) cand[(hit) * 32 + lane] = n; }
                        else if (used < P.maxPaths) { cand[(used) * 32 + lane] = n; used++; }
                    }
                    // ---- score candidates, keep maxima and ties in order (:474-487); sumCounts via the DFS intervals
                    for (int i = 0; i < used; i++) {
                        const int cnode = cand[(i) * 32 + lane];
                        const int pc = __ldg(db.pre + cnode);
                        int sum = 0;
                        for (int j = 0; j < nTab; j++) {
                            const int e = (int)s_vi[j][t];
                            if (__ldg(db.pre + e) <= pc && pc <= __ldg(db.last + e)) sum += (int)s_pk[j][t];
                        }
                        if (sum > best) { best = sum; cand[(0) * 32 + lane] = cnode; ties = 0; }
                        else if (sum == best) { ties++; cand[(ties) * 32 + lane] = cnode; }
                    }
                    // ---- lowestNodeWhereSumAboveThreshold (:488-492, C/tax/SmallTaxTree.java:208-221)
                    if (P.threshold > 1) {
                        for (int i = 0; i <= ties; i++) {
                            int nd = cand[(i) * 32 + lane], res = 0, outn = -1;
                            while (nd >= 0) {
                                int cnt = 0;
                                for (int j = 0; j < nTab; j++) if ((int)s_vi[j][t] == nd) cnt = (int)s_pk[j][t];
                                if (cnt > 0) { res += cnt; if (res >= P.threshold) { outn = nd; break; } }
                                nd = __ldg(db.parent + nd);
                            }
                            cand[(i) * 32 + lane] = outn;
                        }
                    }
                    // ---- LCA of the ties (:493-497)
                    node = cand[(0) * 32 + lane];
                    for (int i = 1; i <= ties; i++) node = gs_lca(db, node, cand[(i) * 32 + lane]);
                }
                classV = node;
                if (node < 0) {
                    found = false;  // `return false` (:498-500)
                } else {
                    int rk = best;
                    if (ties > 0 || P.threshold > 1) {  // :506-507
                        const int c0 = cand[(0) * 32 + lane];
                        const int pc = __ldg(db.pre + c0);
                        rk = 0;
                        for (int j = 0; j < nTab; j++) {
                            const int e = (int)s_vi[j][t];
                            if (__ldg(db.pre + e) <= pc && pc <= __ldg(db.last + e)) rk += (int)s_pk[j][t];
                        }
                    }
                    readKmers = (u32)rk;
                    const int classErrC = max - rk;
                    const double mce = P.maxClassErr;
                    if (mce < 0 || (mce >= 1 && (double)classErrC <= mce) || ((double)classErrC <= mce * (double)max)) {  // :509-510
                        flags |= GS_READ_ACCEPTED;
                        atomicAdd((u64*)(P.counters + 4 * (size_t)V + node), 1ULL);
                        atomicAdd((u64*)(P.counters + 5 * (size_t)V + node), (u64)rk);
                        atomicAdd((u64*)(P.counters + 6 * (size_t)V + node), (u64)L);
                    }
                }
            }
        }
        if (found) flags |= GS_READ_FOUND;
        P.out[r] = gs_read_result{classV, readKmers, taxErr, flags};
    }
}
void gs_launch_reduce_thread(const GsMatchParams& P, int blocks, cudaStream_t st) {
    if (P.bmask) gs_reduce_thread_kernel<true><<<blocks, GS_T_THREADS, 0, st>>>(P);
    else gs_reduce_thread_kernel<false><<<blocks, GS_T_THREADS, 0, st>>>(P);
}

void gs_launch_mark_starts(const GsMatchParams& P, cudaStream_t st) {
    if (P.nReads) gs_mark_starts_kernel<<<148 * 4, 256, 0, st>>>(P.offsets, P.nReads, P.off0, P.lead, P.flatLen, P.startBits);
}
void gs_launch_label(const GsMatchParams& P, bool dump, int blocks, cudaStream_t st) {
    const int threads = GS_WARPS_PER_BLOCK * 32;
    if (P.layout == GS_LAYOUT_TABLE) {
        const bool wide = P.db.mzFilter != nullptr && P.db.mzWide;
        if (dump) { if (wide) gs_label_kernel<GS_LAYOUT_TABLE, true, true, 0><<<blocks, threads, 0, st>>>(P); else gs_label_kernel<GS_LAYOUT_TABLE, true, false, 0><<<blocks, threads, 0, st>>>(P); }
        else if (P.db.k == 31) {  // the reference's default k (C/GSConfigKey.java KMER_SIZE): masks and shift counts compiled in
            if (wide) gs_label_kernel<GS_LAYOUT_TABLE, false, true, 31><<<blocks, threads, 0, st>>>(P); else gs_label_kernel<GS_LAYOUT_TABLE, false, false, 31><<<blocks, threads, 0, st>>>(P);
        } else { if (wide) gs_label_kernel<GS_LAYOUT_TABLE, false, true, 0><<<blocks, threads, 0, st>>>(P); else gs_label_kernel<GS_LAYOUT_TABLE, false, false, 0><<<blocks, threads, 0, st>>>(P); }
    } else {
        if (dump) gs_label_kernel<GS_LAYOUT_CLASSIC, true, false, 0><<<blocks, threads, 0, st>>>(P);
        else gs_label_kernel<GS_LAYOUT_CLASSIC, false, false, 0><<<blocks, threads, 0, st>>>(P);
    }
}
void gs_launch_reduce(const GsMatchParams& P, int mode, bool dump, int blocks, cudaStream_t st) {
    const int threads = GS_WARPS_PER_BLOCK * 32;
    if (mode == 0) {
        if (dump) gs_reduce_kernel<0, true, false><<<blocks, threads, 0, st>>>(P);
        else if (P.bmask) gs_reduce_kernel<0, false, true><<<blocks, threads, 0, st>>>(P);
        else gs_reduce_kernel<0, false, false><<<blocks, threads, 0, st>>>(P);
    } else {
        gs_reduce_kernel<1, false, false><<<blocks, threads, 0, st>>>(P);
    }
}

// New per-taxon maximum contig lengths set by reads of [firstReadNo, firstReadNo + nReads) (:402-409).
__global__ void gs_maxcontig_events_kernel(const u64* __restrict__ maxcontig, int V, u64 firstReadNo, u32 nReads,
                                           gs_maxcontig_event* ev, u32* nEv) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    u64 cur = maxcontig[v];
    if (!cur) return;
    u64 ord = GS_ORDINAL_MASK - (cur & GS_ORDINAL_MASK);
    if (ord >= firstReadNo && ord < firstReadNo + nReads) {
        u32 slot = atomicAdd(nEv, 1u);
        ev[slot] = gs_maxcontig_event{(u32)v, (u32)(cur >> GS_MAXCONTIG_SHIFT), ord};
    }
}
void gs_launch_maxcontig_events(const u64* maxcontig, int V, u64 firstReadNo, u32 nReads, gs_maxcontig_event* ev, u32* nEv, cudaStream_t st) {
    gs_maxcontig_events_kernel<<<(V + 255) / 256, 256, 0, st>>>(maxcontig, V, firstReadNo, nReads, ev, nEv);
}

// KMerUniqueCounterBits.getUniqueKmerCounts (C/store/KMerUniqueCounterBits.java:146-163): per value index, the number of
// set bits among its storage positions -- and, across GPUs, the OR-merge that makes ONE bitset out of the per-GPU ones first
// (KMerUniqueCounterBits is one object in the reference, :117-163; here every GPU has set bits for its share of the reads).
//
// A rank owns the words [wordBegin, wordEnd) of the bitset.  src.p[q][w] is rank q's word w -- either q's bitset itself, read
// over NVLink through a peer mapping (no staging copy: the transfer IS the kernel's loads), or the slice q sent into a local
// receive buffer (pointer pre-offset by -wordBegin).  nSrc == 0: no merge, count the own words.
// Shape: the storage positions of a word are 64 consecutive table slots (or sorted-array indices), so the values are streamed,
// not looked up: a warp ORs 32 words (one coalesced load per source), then walks its non-zero words with every lane owning two
// positions -- one 16-byte load of the two table entries where a bit is set.  Counts go to a per-CTA table in shared memory
// (value indices are spread evenly over the positions: global atomics on [V] counters were the whole cost of the first
// version, 44 ms for the 2e9-k-mer store on 2 GPUs); every CTA writes its table to `partial`, gs_sum_partials_kernel adds them up.
#define GS_POP_THREADS 1024
#define GS_POP_VCHUNK 57344u   // value indices counted per pass: 224 KB of u32 counters in shared memory
__global__ void __launch_bounds__(GS_POP_THREADS, 1) gs_merge_or_popcount_kernel(const GsPeerPtrs src, int nSrc, u64* own, u64 wordBegin, u64 wordEnd,
                                                                                const GsDbView db, int layout, u32 vLo, u32 vCnt, u32* __restrict__ partial) {
    extern __shared__ u32 s_cnt[];
    for (u32 i = threadIdx.x; i < vCnt; i += GS_POP_THREADS) s_cnt[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 stride = (u64)gridDim.x * (GS_POP_THREADS / 32) * 32;
    for (u64 g = wordBegin + ((u64)blockIdx.x * (GS_POP_THREADS / 32) + warp) * 32; g < wordEnd; g += stride) {
        const u64 w = g + lane;
        u64 m = 0;
        if (w < wordEnd) {
            if (nSrc > 0) {
                for (int q = 0; q < nSrc; q++) m |= src.p[q][w];
                own[w] = m;
            } else {
                m = own[w];
            }
        }
        u32 nz = __ballot_sync(FULL, m != 0);
        while (nz) {
            const int j = __ffs(nz) - 1;
            nz &= nz - 1;
            const u64 word = __shfl_sync(FULL, m, j);
            const u32 b2 = (u32)(word >> (2 * lane)) & 3u;
            if (b2) {
                const u64 pos = (g + (u64)j) * 64 + 2 * (u64)lane;   // two neighbouring storage positions per lane
                u32 v0 = GS_VAL_NONODE, v1 = GS_VAL_NONODE;
                if (layout == GS_LAYOUT_TABLE) {
                    if (pos + 1 < db.tabSlots) {
                        u64 e0, e1;
                        asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(e0), "=l"(e1) : "l"(db.tab + pos));
                        if (e0 & GS_TAB_OCC) v0 = (u32)(e0 >> GS_TAB_VAL_SHIFT) & 0xFFFFu;
                        if (e1 & GS_TAB_OCC) v1 = (u32)(e1 >> GS_TAB_VAL_SHIFT) & 0xFFFFu;
                    }
                } else {
                    if (pos < db.n) v0 = __ldg(db.vals + pos);
                    if (pos + 1 < db.n) v1 = __ldg(db.vals + pos + 1);
                }
                if ((b2 & 1u) && v0 != GS_VAL_NONODE && v0 - vLo < vCnt) atomicAdd(&s_cnt[v0 - vLo], 1u);
                if ((b2 & 2u) && v1 != GS_VAL_NONODE && v1 - vLo < vCnt) atomicAdd(&s_cnt[v1 - vLo], 1u);
            }
        }
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < vCnt; i += GS_POP_THREADS) partial[(u64)blockIdx.x * vCnt + i] = s_cnt[i];
}
__global__ void gs_sum_partials_kernel(const u32* __restrict__ partial, int nBlocks, u32 vCnt, long long* unique) {
    const u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= vCnt) return;
    long long acc = 0;
    for (int b = 0; b < nBlocks; b++) acc += partial[(u64)b * vCnt + v];
    unique[v] += acc;
}
// partial = scratch of gs_popcount_scratch_words(blocks, nValues) u32.  unique[v] += set bits of value index v in the words.
u64 gs_popcount_scratch_words(int blocks, int nValues) { return (u64)blocks * std::min<u64>((u64)std::max(nValues, 1), GS_POP_VCHUNK); }
void gs_launch_merge_or_popcount(const GsPeerPtrs& src, int nSrc, u64* own, u64 wordBegin, u64 wordEnd, const GsDbView& db, int layout,
                                 long long* unique, u32* partial, int blocks, cudaStream_t st) {
    if (wordEnd < wordBegin) return;   // an empty range still launches (one block): loads the function ahead of the merge
    const u64 groups = (wordEnd - wordBegin + GS_POP_THREADS - 1) / GS_POP_THREADS;
    blocks = (int)std::max<u64>(1, std::min<u64>((u64)blocks, groups));
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(gs_merge_or_popcount_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(GS_POP_VCHUNK * sizeof(u32))); attr = true; }
    const u32 V = (u32)std::max(db.nValues, 1);
    for (u32 vLo = 0; vLo < V; vLo += GS_POP_VCHUNK) {   // one pass unless there are more than GS_POP_VCHUNK value indices
        const u32 vCnt = std::min(GS_POP_VCHUNK, V - vLo);
        gs_merge_or_popcount_kernel<<<blocks, GS_POP_THREADS, vCnt * sizeof(u32), st>>>(src, vLo == 0 ? nSrc : 0, own, wordBegin, wordEnd, db, layout, vLo, vCnt, partial);
        gs_sum_partials_kernel<<<(vCnt + 255) / 256, 256, 0, st>>>(partial, blocks, vCnt, unique + vLo);
    }
}
void gs_launch_unique_popcount(const u64* bits, u64 wordBegin, u64 wordEnd, const GsDbView& db, int layout, long long* unique, u32* partial, int blocks, cudaStream_t st) {
    GsPeerPtrs none;
    memset(&none, 0, sizeof(none));
    gs_launch_merge_or_popcount(none, 0, const_cast<u64*>(bits), wordBegin, wordEnd, db, layout, unique, partial, blocks, st);
}

// (value index, hit counter) of every set position, for getMaxCountsCounts (C/store/KMerUniqueCounterBits.java:173-199)
__global__ void gs_collect_hits_kernel(const u64* __restrict__ bits, u64 nWords, const uint16_t* __restrict__ hitCounts, GsDbView db, int layout,
                                       u32* out, unsigned long long* nOut, u64 cap) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x; w < nWords; w += stride) {
        u64 word = bits[w];
        while (word) {
            const int b = __ffsll((long long)word) - 1;
            word &= word - 1;
            const u64 pos = w * 64 + (u64)b;
            const u32 vv = gs_value_at(db, layout, pos);
            if (vv == GS_VAL_NONODE) continue;
            const u64 slot = atomicAdd(nOut, 1ULL);
            if (slot < cap) out[slot] = (vv << 16) | (u32)hitCounts[pos];
        }
    }
}
void gs_launch_collect_hits(const u64* bits, u64 nWords, const uint16_t* hitCounts, const GsDbView& db, int layout, u32* out, unsigned long long* nOut, u64 cap, cudaStream_t st) {
    gs_collect_hits_kernel<<<148 * 8, 256, 0, st>>>(bits, nWords, hitCounts, db, layout, out, nOut, cap);
}

// ---- probe table build.  The layout is a pure function of the key set (no dependence on thread arrival order), so that
// every rank of a multi-GPU job -- each builds its own replica -- addresses the same k-mer by the same slot id:
//   1. count:  c_b = number of keys whose home bucket is b                                  (gs_table_count_kernel)
//   2. scan:   delta_0 = 0, delta_{b+1} = max(0, delta_b + c_b - 4)                         (gs_table_scan_kernel, three passes)
//              -- the keys of bucket b occupy the slots [4b + delta_b, 4b + delta_b + c_b): sorted linear probing
//   3. place:  every key takes a slot of its bucket's range (arrival order)                 (gs_table_place_kernel)
//   4. sort:   every range is sorted by entry = by remainder (distinct within a bucket), then every entry is stamped with
//              its displacement = bucket it sits in - home bucket                           (gs_table_sort_kernel)
//   5. flags:  bucket b is flagged `spill` iff delta_{b+1} > 0, i.e. a key whose home is <= b sits behind bucket b
// The recurrence of step 2 composes as functions d -> max(d + s, t): a block of buckets is summarised by (s, t), blocks are
// combined left to right (gs_lq_combine), and delta at a block's first bucket is its prefix applied to 0.
struct GsLq { long long s, t; };   // d -> max(d + s, t)
__device__ __forceinline__ GsLq gs_lq_combine(GsLq a, GsLq b) {  // first a, then b
    GsLq r;
    r.s = a.s + b.s;
    r.t = max(a.t + b.s, b.t);
    return r;
}
#define GS_TB_THREADS 256
#define GS_TB_PER_THREAD 8
#define GS_TB_BLOCK (GS_TB_THREADS * GS_TB_PER_THREAD)

__global__ void gs_table_count_kernel(const u64* __restrict__ keys, u64 n, u32* counts, int rbits) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(counts + (gs_mix62(keys[i]) >> rbits), 1u);
}
// PHASE 0: (s, t) of every block of GS_TB_BLOCK buckets -> agg[block].  PHASE 1: delta of every bucket from blockIn[block].
template <int PHASE>
__global__ void __launch_bounds__(GS_TB_THREADS) gs_table_scan_kernel(const u32* __restrict__ counts, u64 nBuckets, GsLq* agg, const u32* __restrict__ blockIn, u32* delta, u32* bad) {
    __shared__ GsLq sh[GS_TB_THREADS];
    const u64 b0 = (u64)blockIdx.x * GS_TB_BLOCK + (u64)threadIdx.x * GS_TB_PER_THREAD;
    GsLq mine = {0, -(1LL << 60)};  // identity
    u32 c[GS_TB_PER_THREAD];
#pragma unroll
    for (int i = 0; i < GS_TB_PER_THREAD; i++) {
        const u32 raw = b0 + i < nBuckets ? counts[b0 + i] : 4u;      // buckets past the end: neutral (x = 0)
        if (PHASE == 0 && raw > 0xFFFFu) atomicOr(bad, 1u);           // the place kernel keeps its fill cursor in the upper half
        c[i] = raw & 0xFFFFu;
        mine = gs_lq_combine(mine, GsLq{(long long)c[i] - 4, 0});
    }
    sh[threadIdx.x] = mine;
    __syncthreads();
    for (int d = 1; d < GS_TB_THREADS; d <<= 1) {   // inclusive scan, left operand first
        GsLq v = sh[threadIdx.x];
        if ((int)threadIdx.x >= d) v = gs_lq_combine(sh[threadIdx.x - d], v);
        __syncthreads();
        sh[threadIdx.x] = v;
        __syncthreads();
    }
    if (PHASE == 0) {
        if (threadIdx.x == GS_TB_THREADS - 1) agg[blockIdx.x] = sh[threadIdx.x];
    } else {
        long long d = (long long)blockIn[blockIdx.x];
        if (threadIdx.x > 0) { const GsLq p = sh[threadIdx.x - 1]; d = max(d + p.s, p.t); }
#pragma unroll
        for (int i = 0; i < GS_TB_PER_THREAD; i++) {
            if (b0 + i < nBuckets) delta[b0 + i] = (u32)d;
            d = max(0LL, d + (long long)c[i] - 4);
        }
        if (b0 + GS_TB_PER_THREAD >= nBuckets && b0 < nBuckets) delta[nBuckets] = (u32)d;  // written by the thread that owns the last bucket
    }
}
// one block: blockIn[j] = delta at the first bucket of block j (sequential over a thread's share, scan across the threads)
__global__ void __launch_bounds__(1024) gs_table_scan_blocks_kernel(const GsLq* __restrict__ agg, u64 nBlocks, u32* blockIn) {
    __shared__ GsLq sh[1024];
    const u64 per = (nBlocks + 1023) / 1024;
    const u64 j0 = (u64)threadIdx.x * per, j1 = min(nBlocks, j0 + per);
    GsLq mine = {0, -(1LL << 60)};
    for (u64 j = j0; j < j1; j++) mine = gs_lq_combine(mine, agg[j]);
    sh[threadIdx.x] = mine;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        GsLq v = sh[threadIdx.x];
        if ((int)threadIdx.x >= d) v = gs_lq_combine(sh[threadIdx.x - d], v);
        __syncthreads();
        sh[threadIdx.x] = v;
        __syncthreads();
    }
    long long d = 0;
    if (threadIdx.x > 0) { const GsLq p = sh[threadIdx.x - 1]; d = max(0LL + p.s, p.t); d = max(d, 0LL); }
    for (u64 j = j0; j < j1; j++) {
        blockIn[j] = (u32)d;
        const GsLq a = agg[j];
        d = max(d + a.s, a.t);
    }
}
__global__ void gs_table_place_kernel(const u64* __restrict__ keys, const uint16_t* __restrict__ vals, u64 n, u64* tab, u32* counts, const u32* __restrict__ delta, int rbits) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u64 h = gs_mix62(keys[i]);
        const u64 b = h >> rbits;
        const u64 entry = ((h & ((1ULL << rbits) - 1)) << GS_TAB_REM_SHIFT) | ((u64)vals[i] << GS_TAB_VAL_SHIFT) | GS_TAB_OCC;
        const u32 idx = atomicAdd(counts + b, 0x10000u) >> 16;   // fill cursor in the upper half, the count stays below
        tab[b * 4 + delta[b] + idx] = entry;
    }
}
__global__ void gs_table_sort_kernel(u64* tab, const u32* __restrict__ counts, const u32* __restrict__ delta, u64 nBuckets, u32* bad) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x; b < nBuckets; b += stride) {
        const u32 c = counts[b] & 0xFFFFu;
        if (c == 0) continue;
        const u32 d = delta[b];
        u64* g = tab + b * 4 + d;
        for (u32 i = 1; i < c; i++) {  // insertion sort: the ranges hold a handful of entries
            const u64 e = g[i];
            u32 j = i;
            while (j > 0 && g[j - 1] > e) { g[j] = g[j - 1]; j--; }
            g[j] = e;
        }
        if ((d + c - 1) / 4 > GS_TAB_DISP_MAX) { atomicOr(bad, 2u); continue; }   // further from home than the entry format can say
        for (u32 i = d >= 4 ? 0 : 4 - d; i < c; i++) g[i] |= (u64)((d + i) / 4) << GS_TAB_DISP_SHIFT;
    }
}
__global__ void gs_table_flags_kernel(u64* tab, const u32* __restrict__ delta, u64 nBuckets) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x; b < nBuckets; b += stride)
        if (delta[b + 1] > 0) tab[b * 4] |= GS_TAB_SPILL;
}
// seen bits of all entries -> 0 (a session leases them), and seen bits -> compact bitset (bit = slot id)
__global__ void gs_table_clear_seen_kernel(u64* tab, u64 nSlots) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nSlots; i += stride) {
        const u64 e = tab[i];
        if (e & GS_TAB_SEEN) tab[i] = e & ~GS_TAB_SEEN;
    }
}
__global__ void gs_table_extract_seen_kernel(const u64* __restrict__ tab, u64 nSlots, u64* out) {
    // one warp gathers 32 x 2 entries into one u64 word of the bitset with two ballots
    const u64 warpsTotal = ((u64)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (u64 w = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w * 64 < nSlots; w += warpsTotal) {
        const u64 i0 = w * 64 + lane, i1 = i0 + 32;
        const u32 lo = __ballot_sync(FULL, i0 < nSlots && (__ldcg(tab + i0) & GS_TAB_SEEN));
        const u32 hi = __ballot_sync(FULL, i1 < nSlots && (__ldcg(tab + i1) & GS_TAB_SEEN));
        if (lane == 0) out[w] = ((u64)hi << 32) | lo;
    }
}
// minimizer prefilter of the store: one bit per minimizer hash of every stored key (gs_device.cuh "minimizer prefilter")
__global__ void gs_mz_build_kernel(const u64* __restrict__ keys, u64 n, int k, u64* filter, u32 mask, int wide) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u32 idx = gs_mz_index(wide ? (u32)gs_mz_of_key_w(keys[i], k) : gs_mz_of_key(keys[i], k), mask);
        const u64 bit = 1ULL << (idx & 63);
        if (!(filter[idx >> 6] & bit)) atomicOr(filter + (idx >> 6), bit);
    }
}
void gs_launch_mz_build(const u64* keys, u64 n, int k, u64* filter, u32 mask, int wide, cudaStream_t st) { gs_mz_build_kernel<<<148 * 8, 256, 0, st>>>(keys, n, k, filter, mask, wide); }
void gs_launch_table_clear_seen(u64* tab, u64 nSlots, cudaStream_t st) { gs_table_clear_seen_kernel<<<148 * 8, 256, 0, st>>>(tab, nSlots); }
void gs_launch_table_extract_seen(const u64* tab, u64 nSlots, u64* out, cudaStream_t st) { gs_table_extract_seen_kernel<<<148 * 8, 256, 0, st>>>(tab, nSlots, out); }
// nBuckets = 2^tbits + GS_TAB_PAD_BUCKETS; counts[nBuckets], delta[nBuckets + 1], agg / blockIn[gs_table_scan_blocks(nBuckets)]
// are scratch (counts and *bad zeroed by the caller).  Afterwards delta[nBuckets] must be 0 (nothing ran past the pad buckets)
// and *bad must be 0 (bit 0: a bucket with 2^16 or more keys, bit 1: a key more than GS_TAB_DISP_MAX buckets from home).
u64 gs_table_scan_blocks(u64 nBuckets) { return (nBuckets + GS_TB_BLOCK - 1) / GS_TB_BLOCK; }
void gs_launch_table_build(const u64* keys, const uint16_t* vals, u64 n, u64* tab, u32* counts, u32* delta, void* agg, u32* blockIn,
                           u32* bad, u64 nBuckets, int rbits, cudaStream_t st) {
    const u64 nBlocks = gs_table_scan_blocks(nBuckets);
    gs_table_count_kernel<<<148 * 8, 256, 0, st>>>(keys, n, counts, rbits);
    gs_table_scan_kernel<0><<<(unsigned)nBlocks, GS_TB_THREADS, 0, st>>>(counts, nBuckets, (GsLq*)agg, nullptr, nullptr, bad);
    gs_table_scan_blocks_kernel<<<1, 1024, 0, st>>>((const GsLq*)agg, nBlocks, blockIn);
    gs_table_scan_kernel<1><<<(unsigned)nBlocks, GS_TB_THREADS, 0, st>>>(counts, nBuckets, nullptr, blockIn, delta, bad);
    gs_table_place_kernel<<<148 * 8, 256, 0, st>>>(keys, vals, n, tab, counts, delta, rbits);
    gs_table_sort_kernel<<<148 * 8, 256, 0, st>>>(tab, counts, delta, nBuckets, bad);
    gs_table_flags_kernel<<<148 * 8, 256, 0, st>>>(tab, delta, nBuckets);
}

// ---- end-of-run merge of the per-GPU unique-k-mer state: gs_merge_or_popcount_kernel above; here the hit counters
// per-position hit counters (maxKMerResCounts > 0): own[i] = sum over the ranks, Java short wrap-around (:134-140)
__global__ void gs_merge_add_u16_kernel(const GsPeerPtrs src, int nSrc, uint16_t* own, u64 begin, u64 end) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = begin + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += stride) {
        u32 acc = 0;
        for (int q = 0; q < nSrc; q++) acc += ((const uint16_t*)src.p[q])[i];
        own[i] = (uint16_t)acc;
    }
}
void gs_launch_merge_add_u16(const GsPeerPtrs& src, int nSrc, uint16_t* own, u64 begin, u64 end, cudaStream_t st) {
    if (end > begin) gs_merge_add_u16_kernel<<<148 * 8, 256, 0, st>>>(src, nSrc, own, begin, end);
}

// ---------------------------------------------------------------------------------------------------------
// database build helpers
// ---------------------------------------------------------------------------------------------------------
// bucket index: bstart[b] = first position whose key has top bits >= b; bstart[nb] = n
__global__ void gs_bucket_index_kernel(const u64* __restrict__ keys, u64 n, int bshift, u64 nb, u32* bstart) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) {
        const u64 cur = i < n ? (keys[i] >> bshift) : nb;
        const u64 from = i == 0 ? 0 : (keys[i - 1] >> bshift) + 1;
        for (u64 b = from; b <= cur && b <= nb; b++) bstart[b] = (u32)i;
    }
}
void gs_launch_bucket_index(const u64* keys, u64 n, int bshift, u64 nb, u32* bstart, cudaStream_t st) {
    gs_bucket_index_kernel<<<148 * 8, 256, 0, st>>>(keys, n, bshift, nb, bstart);
}

// BlockedKMerBloomFilter.putLong (C/bloom/BlockedKMerBloomFilter.java:108-124) for every stored key
__global__ void gs_bloom_build_kernel(const u64* __restrict__ keys, u64 n, u64* words, u64 buckets, u64 magic, long long seed) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        long long h = seed ^ (long long)keys[i];
        u64 start = gs_absmod(h, buckets, magic);
        u64 h2 = (u64)h ^ gs_rotl64((u64)h, 32);
        u64 m1 = (1ULL << (h2 & 63)) | (1ULL << ((h2 >> 6) & 63));
        u64 m2 = (1ULL << ((h2 >> 12) & 63)) | (1ULL << ((h2 >> 18) & 63));
        atomicOr(words + start, m1);
        atomicOr(words + start + 1 + (h2 >> 60), m2);
    }
}
void gs_launch_bloom_build(const u64* keys, u64 n, u64* words, u64 buckets, u64 magic, long long seed, cudaStream_t st) {
    gs_bloom_build_kernel<<<148 * 8, 256, 0, st>>>(keys, n, words, buckets, magic, seed);
}

__global__ void gs_convert_values_kernel(const int16_t* __restrict__ raw, const int* __restrict__ hasNode, u64 n, int V, uint16_t* vals, u32* bad) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int idx = (int)raw[i] + 32768;  // index - Short.MIN_VALUE (C/store/KMerSortedArray.java:348)
        if (idx >= V) { atomicAdd(bad, 1u); vals[i] = GS_VAL_NONODE; }
        else vals[i] = (hasNode && !hasNode[idx]) ? GS_VAL_NONODE : (uint16_t)idx;
    }
}
void gs_launch_convert_values(const int16_t* raw, const int* hasNode, u64 n, int V, uint16_t* vals, u32* bad, cudaStream_t st) {
    gs_convert_values_kernel<<<148 * 8, 256, 0, st>>>(raw, hasNode, n, V, vals, bad);
}

__global__ void gs_check_sorted_kernel(const u64* __restrict__ keys, u64 n, int k, u32* bad) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (keys[i] >> (2 * k)) atomicAdd(bad, 1u);
        if (i + 1 < n && keys[i] >= keys[i + 1]) atomicAdd(bad, 1u);
    }
}
void gs_launch_check_sorted(const u64* keys, u64 n, int k, u32* bad, cudaStream_t st) { gs_check_sorted_kernel<<<148 * 8, 256, 0, st>>>(keys, n, k, bad); }

// KMerStore.getLong probe (tests)
__global__ void gs_lookup_kernel(GsDbView db, const u64* __restrict__ kmers, u64 n, int useBloom, int* vidx, long long* pos) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        u64 p = 0;
        u32 lab = gs_lookup(db, kmers[i], useBloom && db.hasBloom, p);
        // the probe table must agree with the reference structures on every query
        u64 p2 = 0;
        if (db.tab && gs_lookup_table(db, kmers[i], p2) != lab) lab = 0x7FFFFFFFu;
        // ... and the minimizer prefilter must pass every stored key
        if (db.mzFilter && lab != GS_LABEL_MISS && !gs_mz_test(db.mzFilter, db.mzMask, db.mzWide ? (u32)gs_mz_of_key_w(kmers[i], db.k) : gs_mz_of_key(kmers[i], db.k))) lab = 0x7FFFFFFEu;
        vidx[i] = lab == GS_LABEL_MISS ? -1 : (int)lab;
        pos[i] = lab == GS_LABEL_MISS ? -1LL : (long long)p;
    }
}
void gs_launch_lookup(const GsDbView& db, const u64* kmers, u64 n, int useBloom, int* vidx, long long* pos, cudaStream_t st) {
    gs_lookup_kernel<<<148 * 4, 256, 0, st>>>(db, kmers, n, useBloom, vidx, pos);
}

// ---------------------------------------------------------------------------------------------------------
// filter
// ---------------------------------------------------------------------------------------------------------
// KMerProbFilter.containsLong for the three filter kinds
__device__ __forceinline__ bool gs_filter_contains(const GsFilterView& f, u64 key) {
    if (f.kind == GS_BLOOM_BLOCKED) return gs_bloom_blocked(f.words, (u64)f.p1, f.magic, f.p0, key);
    // AbstractKMerBloomFilter.containsLong (C/bloom/AbstractKMerBloomFilter.java:209-216): stop at the first 0 bit
    for (int i = 0; i < (int)f.p1; i++) {
        const u64 fac = (u64)__ldg(f.factors + i);
        const long long h = f.kind == GS_BLOOM_XOR ? (long long)(fac ^ key) : (long long)gs_murmur64(key, fac);
        const u64 idx = gs_absmod(h, (u64)f.p0, f.magic);
        if (!((__ldg(f.words + (idx >> 6)) >> (idx & 63)) & 1ULL)) return false;
    }
    return true;
}

__global__ void gs_filter_contains_kernel(GsFilterView f, const u64* __restrict__ kmers, u64 n, uint8_t* out) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = gs_filter_contains(f, kmers[i]) ? 1 : 0;
}
void gs_launch_filter_contains(const GsFilterView& f, const u64* kmers, u64 n, uint8_t* out, cudaStream_t st) {
    gs_filter_contains_kernel<<<148 * 4, 256, 0, st>>>(f, kmers, n, out);
}

// FastqBloomFilter.isAcceptRead (C/bloom/FastqBloomFilter.java:120-161).  The two early exits are mutually
// exclusive (hits + misses <= max), so the decision is  hits_total >= max(1, posThreshold).
//
// Shape (round 2; the warp-per-read kernel of round 1 sat on the DRAM line-request cap with 1.70 lines per k-mer): the same
// flat decomposition as the match path.  gs_filter_flat_kernel walks the batch's bases as ONE array (segments of GS_SEG_POS
// positions claimed by atomic counter, staged like the label kernel), every lane tests one k-mer and the warp writes one hit bit
// per position; gs_filter_accept_kernel counts a read's bits (one thread per read).
//
// What was measured on the way (profiles/experiments/README.md "filter kernel"): the flat kernel runs at the same speed as the
// warp-per-read one (18.6 vs 18.2 ms per 4 M reads) -- both sit on the DRAM line-request cap, an absent k-mer costs two
// probes = two DRAM lines of a filter that is half full and four times the L2 -- and a pass that first looked for a clear
// bit among the hash functions landing in an L2-pinned part of the filter cost more issue slots than it saved requests.
#ifndef GS_FILTER_MIN_BLOCKS
#define GS_FILTER_MIN_BLOCKS 6
#endif

template <int KIND>
__device__ __forceinline__ u64 gs_filter_index(const GsFilterView& f, long long fac, u64 key) {
    const long long h = KIND == GS_BLOOM_XOR ? (long long)((u64)fac ^ key) : (long long)gs_murmur64(key, (u64)fac);
    return gs_absmod(h, (u64)f.p0, f.magic);
}

// AbstractKMerBloomFilter.containsLong (C/bloom/AbstractKMerBloomFilter.java:209-216): the reference's order, stop at the first 0 bit
template <int KIND>
__device__ __forceinline__ bool gs_filter_contains_hashed(const GsFilterView& f, const long long* fac, u64 key) {
    const int H = (int)f.p1;
    for (int i = 0; i < H; i++) {
        const u64 idx = gs_filter_index<KIND>(f, fac[i], key);
        if (!((__ldg(f.words + (idx >> 6)) >> (idx & 63)) & 1ULL)) return false;
    }
    return true;
}

#define GS_FILTER_MAX_FACTORS 64   // hash factors staged in shared memory (27 at fpp 1e-8); filters with more use the global array

template <int KIND, int KT>
__global__ void __launch_bounds__(GS_WARPS_PER_BLOCK * 32, GS_FILTER_MIN_BLOCKS) gs_filter_flat_kernel(const GsFilterParams P) {
    __shared__ GsSegStage s_stage[GS_WARPS_PER_BLOCK];
    __shared__ long long s_fac[GS_FILTER_MAX_FACTORS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = KT > 0 ? KT : P.k;
    const u32 kmask = (k >= 32) ? 0xFFFFFFFFu : ((1u << k) - 1u);
    const u32 k1mask = (1u << (k - 1)) - 1u;
    const long long* fac = P.f.factors;
    if (KIND != GS_BLOOM_BLOCKED && (int)P.f.p1 <= GS_FILTER_MAX_FACTORS) {
        if ((int)threadIdx.x < (int)P.f.p1) s_fac[threadIdx.x] = __ldg(P.f.factors + threadIdx.x);
        fac = s_fac;
    }
    __syncthreads();
    GsSegStage& S = s_stage[warp];
    const u32* cp = S.code + (lane >> 4);
    const u32 csh = (u32)(lane & 15) * 2;
    const uint8_t* fb = P.bases + P.off0 - P.lead;  // 16-byte aligned start of the flat array
    const u64 nSeg = (P.flatLen + GS_SEG_POS - 1) / GS_SEG_POS;
    for (;;) {
        u32 seg = 0;
        if (lane == 0) seg = atomicAdd(P.segCounter, 1u);
        seg = __shfl_sync(FULL, seg, 0);
        if (seg >= nSeg) break;
        const u64 f0 = (u64)seg * GS_SEG_POS;
        const int nb = (int)min((u64)GS_SEG_BASES, P.flatLen - f0);
        const u64 w0 = (u64)seg * GS_SEG_CHUNKS;
        __syncwarp();
        const uint4* ap = (const uint4*)(fb + f0);
#pragma unroll
        for (int j = lane; j < GS_SEG_BASES / 16; j += 32) {
            u32 code = 0, valid = 0;
            const int rem = nb - j * 16;
            if (rem > 0) {
                const uint4 A = __ldg(ap + j);
                u32 c0, c1, c2, c3, v0, v1, v2, v3;
                gs_conv4(A.x, c0, v0); gs_conv4(A.y, c1, v1); gs_conv4(A.z, c2, v2); gs_conv4(A.w, c3, v3);
                code = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
                valid = v0 | (v1 << 4) | (v2 << 8) | (v3 << 12);
                if (rem < 16) valid &= (1u << rem) - 1u;
                if (seg == 0) {  // alignment bytes in front of the first read are not bases
                    const int lo = (int)P.lead - j * 16;
                    if (lo > 0) valid &= lo >= 16 ? 0u : ~((1u << lo) - 1u);
                }
            }
            S.code[j] = code;
            ((uint16_t*)S.valid)[j] = (uint16_t)valid;
        }
        S.start[lane] = __ldg(P.startBits + w0 + lane);
        if (lane < 2) { ((uint2*)S.code)[32 + lane] = make_uint2(0u, 0u); S.valid[32 + lane] = 0; S.start[32 + lane] = lane == 0 ? __ldg(P.startBits + w0 + 32) : 0u; }
        __syncwarp();
#pragma unroll 1
        for (int c = 0; c < GS_SEG_CHUNKS; c++) {
            const u32 vbits = __funnelshift_r(S.valid[c], S.valid[c + 1], lane);
            const u32 sbits = __funnelshift_rc(S.start[c], S.start[c + 1], lane + 1);
            // a window that crosses a read boundary is no k-mer of any read; one that holds a non-CGAT byte counts as a miss
            // (FastqBloomFilter.java:133-141)
            const bool valid = !(sbits & k1mask) && (vbits & kmask) == kmask;
            bool hit = false;
            if (valid) {
                const u64 key = gs_canonical(gs_extract_w(cp + 2 * c, csh, k), k);
                if (KIND == GS_BLOOM_BLOCKED) hit = gs_bloom_blocked(P.f.words, (u64)P.f.p1, P.f.magic, P.f.p0, key);
                else hit = gs_filter_contains_hashed<KIND>(P.f, fac, key);
            }
            __syncwarp();
            const u32 hb = __ballot_sync(FULL, hit);
            if (lane == 0) P.hitBits[w0 + c] = hb;
        }
    }
}

// one thread per read: hits = set bits of the read's k-mer positions, accept = hits >= max(1, posThreshold) (:122, :143-158)
__global__ void gs_filter_accept_kernel(const GsFilterParams P) {
    const u32 stride = gridDim.x * blockDim.x;
    for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < P.nReads; r += stride) {
        const u64 start = P.offsets[r], end = P.offsets[r + 1];
        long long L = (long long)(end - start);
        const u64 f = start - P.off0 + P.lead;
        if (end < start || end - start > 0x7FFFFFF0ULL || start < P.off0 || f + (u64)L > P.flatLen) {
            if (P.errFlag) atomicOr(P.errFlag, 1u);
            L = 0;
        }
        const int max = (int)L - P.k + 1;
        const int posThreshold = P.minPosCount > 0 ? P.minPosCount : (int)((double)max * P.posRatio);  // :122
        const int need = posThreshold < 1 ? 1 : posThreshold;
        int hits = 0;
        if (max > 0) {
            const u64 e = f + (u64)max;   // bits [f, e)
            for (u64 w = f >> 5; w <= ((e - 1) >> 5); w++) {
                u32 m = __ldg(P.hitBits + w);
                if (w == (f >> 5)) m &= 0xFFFFFFFFu << (f & 31);
                if (w == ((e - 1) >> 5)) m &= 0xFFFFFFFFu >> (31 - (u32)((e - 1) & 31));
                hits += __popc(m);
            }
        }
        P.accept[r] = hits >= need ? 1 : 0;
    }
}

void gs_launch_filter(const GsFilterParams& P, int blocks, cudaStream_t st) {
    const int threads = GS_WARPS_PER_BLOCK * 32;
    gs_mark_starts_kernel<<<148 * 4, 256, 0, st>>>(P.offsets, P.nReads, P.off0, P.lead, P.flatLen, P.startBits);
    if (P.f.kind == GS_BLOOM_XOR) { if (P.k == 31) gs_filter_flat_kernel<GS_BLOOM_XOR, 31><<<blocks, threads, 0, st>>>(P); else gs_filter_flat_kernel<GS_BLOOM_XOR, 0><<<blocks, threads, 0, st>>>(P); }
    else if (P.f.kind == GS_BLOOM_MURMUR) gs_filter_flat_kernel<GS_BLOOM_MURMUR, 0><<<blocks, threads, 0, st>>>(P);
    else gs_filter_flat_kernel<GS_BLOOM_BLOCKED, 0><<<blocks, threads, 0, st>>>(P);
    gs_filter_accept_kernel<<<148 * 4, 256, 0, st>>>(P);
}

int gs_match_kernel_occupancy(int mode) {
    int nb = 0;
    if (mode == 0) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gs_reduce_kernel<0, false, false>, GS_WARPS_PER_BLOCK * 32, 0);
    else if (mode == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gs_reduce_kernel<1, false, false>, GS_WARPS_PER_BLOCK * 32, 0);
    else if (mode == 3) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gs_label_kernel<GS_LAYOUT_TABLE, false, false, 31>, GS_WARPS_PER_BLOCK * 32, 0);
    else if (mode == 4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gs_label_kernel<GS_LAYOUT_CLASSIC, false, false, 0>, GS_WARPS_PER_BLOCK * 32, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gs_filter_flat_kernel<GS_BLOOM_XOR, 31>, GS_WARPS_PER_BLOCK * 32, 0);
    return nb;
}

// ---------------------------------------------------------------------------------------------------------
// database update phase (C/goals/refseq/DBGoal.java:234-311): value = LCA(value, node of the region) for every stored
// k-mer of a genome region.  The label kernel (classic layout, dump mode) has left the storage position of every hit in
// flatPos; LCA only moves values towards the root, so a compare-and-swap loop on the 16-bit value makes the result
// independent of the order in which regions and positions are applied (C/goals/refseq/FastaReaderGoal.java:104-108).
// ---------------------------------------------------------------------------------------------------------
__global__ void gs_db_update_kernel(const GsDbView db, const u32* __restrict__ labels, const long long* __restrict__ flatPos, u64 flatLen,
                                    const u64* __restrict__ offsets, const int* __restrict__ regionNode, u32 nRegions, uint16_t* vals,
                                    unsigned long long* nChanged) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 f = (u64)blockIdx.x * blockDim.x + threadIdx.x; f < flatLen; f += stride) {
        if (labels[f] >= GS_LABEL_INVALID) continue;
        u32 lo = 0, hi = nRegions;  // region r with offsets[r] <= f < offsets[r + 1]
        while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (__ldg(offsets + mid) <= f) lo = mid; else hi = mid; }
        const int node = __ldg(regionNode + lo);
        if (node < 0 || node >= db.nValues) continue;
        const u64 pos = (u64)flatPos[f];
        u32* wp = (u32*)vals + (pos >> 1);
        const int sh = (int)(pos & 1) * 16;
        u32 old = *(volatile u32*)wp;
        for (;;) {
            const u32 cur = (old >> sh) & 0xFFFFu;
            if (cur == GS_VAL_NONODE) break;                       // getNodeByTaxId(oldValue) == null: unchanged
            const int lca = gs_lca(db, (int)cur, node);
            if (lca < 0 || (u32)lca == cur) break;                 // lcaNode == null keeps the old value (:252)
            const u32 nv = (old & ~(0xFFFFu << sh)) | ((u32)lca << sh);
            const u32 seen = atomicCAS(wp, old, nv);
            if (seen == old) { atomicAdd(nChanged, 1ULL); break; }
            old = seen;
        }
    }
}
void gs_launch_db_update(const GsDbView& db, const u32* labels, const long long* flatPos, u64 flatLen, const u64* offsets, const int* regionNode,
                         u32 nRegions, uint16_t* vals, unsigned long long* nChanged, cudaStream_t st) {
    gs_db_update_kernel<<<148 * 8, 256, 0, st>>>(db, labels, flatPos, flatLen, offsets, regionNode, nRegions, vals, nChanged);
}
// cgatToUpperCase (C/util/CGAT.java:91-99) over a byte buffer, 16 bytes per thread step
__global__ void gs_cgat_upper_kernel(uint8_t* buf, u64 n) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i * 4 < n; i += stride) {
        u32 w = ((u32*)buf)[i];
        const u32 m = __vcmpeq4(w, 0x61616161u) | __vcmpeq4(w, 0x63636363u) | __vcmpeq4(w, 0x67676767u) | __vcmpeq4(w, 0x74747474u);
        ((u32*)buf)[i] = w ^ (m & 0x20202020u);
    }
}
void gs_launch_cgat_upper(uint8_t* buf, u64 n, cudaStream_t st) { gs_cgat_upper_kernel<<<148 * 8, 256, 0, st>>>(buf, n); }
// `kept` = the shorts as they were uploaded (only there when some value index has no tree node): a position whose value was
// collapsed to GS_VAL_NONODE gets its original short back, so save -> load round-trips such databases.
__global__ void gs_values_to_raw_kernel(const uint16_t* __restrict__ vals, const int16_t* __restrict__ kept, u64 n, int16_t* raw) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u32 v = vals[i];
        raw[i] = v == GS_VAL_NONODE ? (kept ? kept[i] : (int16_t)-1) : (int16_t)((int)v - 32768);  // value index + Short.MIN_VALUE (KMerSortedArray.java:348)
    }
}
void gs_launch_values_to_raw(const uint16_t* vals, const int16_t* kept, u64 n, int16_t* raw, cudaStream_t st) { gs_values_to_raw_kernel<<<148 * 8, 256, 0, st>>>(vals, kept, n, raw); }
