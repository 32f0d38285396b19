// gs_device.cuh -- device-side primitives of the read-matching path (sm_100a).
// Reference semantics cited as C/ = core/src/main/java/org/metagene/genestrip/ (pfeiferd/genestrip v3.0).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#define GS_LABEL_END 0xFFFFFFFFu      // terminator lane (past the last k-mer position)
#define GS_LABEL_MISS 0xFFFFFFFEu     // taxIdNode == null
#define GS_LABEL_PENDING 0xFFFFFFFCu  // kernel-internal: table header load issued, not resolved yet
#ifndef GS_GROUP
#define GS_GROUP 4                    // chunks (of 32 positions) whose probe-table loads are in flight together
#endif
#ifndef GS_CLAIM
#define GS_CLAIM 8                    // reads a warp claims per atomic of the work counter
#endif
#ifndef GS_MIN_BLOCKS
#define GS_MIN_BLOCKS 3               // resident CTAs per SM the match kernel is compiled for (register budget)
#endif
#define GS_LABEL_INVALID 0xFFFFFFFDu  // INVALID_NODE (C/match/FastqKMerMatcher.java:63)
#define GS_LAYOUT_TABLE 0             // default device index: 128-byte probe table (one DRAM line touch per k-mer)
#define GS_LAYOUT_CLASSIC 1           // the reference's structures: blocked Bloom filter + (bucketed) binary search
#define GS_VAL_NONODE 0xFFFFu         // stored value whose tax id has no tree node -> null (C/store/Database.java:136-143)

// label kernel: a warp labels a segment of GS_SEG_POS consecutive positions of the batch's flat base array
#define GS_SEG_CHUNKS 31
#define GS_SEG_POS (GS_SEG_CHUNKS * 32)    // 992 positions ...
#define GS_SEG_BASES 1024                  // ... need 992 + k - 1 <= 1023 bases: 64 aligned 16-byte groups
#define GS_SEG_WORDS 34                    // 32 words of codes / validity / read starts + zero padding for the funnel shifts
// Resident CTAs per SM the label kernel is compiled for.  Measured (viral workload, 4 M reads): 4 CTAs 8.16 ms, 5 CTAs 7.25 ms,
// 6 CTAs 6.71 ms, 8 CTAs (32 registers, no spills) 6.51 ms -- the kernel hides DRAM latency with resident warps.
#ifndef GS_LABEL_MIN_BLOCKS
#define GS_LABEL_MIN_BLOCKS 8
#endif
#ifndef GS_LABEL_MIN_BLOCKS_WIDE
#define GS_LABEL_MIN_BLOCKS_WIDE 8
#endif
#define GS_WARPS_PER_BLOCK 8
#define GS_TILE_POS 1024                   // k-mer positions per tile
#define GS_TILE_BASES (GS_TILE_POS + 32)   // bases staged per tile (positions + k-1 <= +30, padded to 16)
#define GS_TILE_GROUPS (GS_TILE_BASES / 16)
#define GS_CODE_WORDS (GS_TILE_BASES / 32 + 1)
#define GS_VALID_WORDS (GS_TILE_BASES / 32 + 1)
#define GS_TABLE_CAP 128                   // distinct taxa per read tracked by the fast path
#define GS_MAX_PATHS 128                   // maxClassificationPaths upper bound (C/GSConfigKey.java:350)
#define GS_MAXCONTIG_SHIFT 40
#define GS_ORDINAL_MASK ((1ULL << GS_MAXCONTIG_SHIFT) - 1)

// Device view of the database (one per device).
struct GsDbView {
    const u64* keys;        // sorted ascending, storage position == index (KMerSortedArray)
    const uint16_t* vals;   // value index per position (Java short - Short.MIN_VALUE), GS_VAL_NONODE if no node
    u64 n;
    const u32* bstart;      // bucket index over the top `bbits` key bits: [bstart[b], bstart[b+1])
    int bshift;             // key >> bshift = bucket
    u64 nBuckets;
    int k;
    const u64* bloom;       // BlockedKMerBloomFilter words (buckets + 17)
    u64 bloomBuckets, bloomMagic;
    long long bloomSeed;
    int hasBloom;
    // probe table (GS_LAYOUT_TABLE): 2^tbits buckets of one 32-byte sector = four 8-byte entries; one 256-bit load
    // answers hit/miss and yields the value and the position's seen bit (see "probe table" below)
    const u64* tab;
    int tbits, rbits;       // bucket = mix62(key) >> rbits, remainder = low rbits bits, rbits = 62 - tbits
    u64 tabSlots;           // 4 * (2^tbits + GS_TAB_PAD_BUCKETS): slots incl. the landing zone behind the last home bucket
    // minimizer prefilter (see "minimizer prefilter" below): bit gs_mz_index(h, mzMask) of mzFilter is set for the hash h of the
    // minimizer of every stored k-mer; NULL = no prefilter (k too small)
    const u64* mzFilter;
    u32 mzMask;
    int mzWide;             // 1: minimizers are ordered by the 64-bit hash and the filter bit comes from its lower half (large stores)
    const int* parent;      // by value index, -1 root / none
    const int* depth;
    const int* pre;         // DFS interval labels: a is ancestor-or-self of b  <=>  pre[a] <= pre[b] <= last[a]
    const int* last;
    int nValues;
};

// |v| mod d for the Java idiom Math.abs(v % d) (C/bloom/BlockedKMerBloomFilter.java:248-249,
// C/bloom/AbstractKMerBloomFilter.java:265-267): abs(v % d) == |v| mod d, and |Long.MIN_VALUE| = 2^63 as unsigned.
// magic = floor((2^64-1)/d): q = mulhi(a, magic) is floor(a/d) or one less, so one conditional subtract suffices.
__device__ __forceinline__ u64 gs_absmod(long long v, u64 d, u64 magic) {
    u64 a = v < 0 ? (u64)0 - (u64)v : (u64)v;
    u64 q = __umul64hi(a, magic);
    u64 r = a - q * d;
    return r >= d ? r - d : r;
}

__device__ __forceinline__ u64 gs_rotl64(u64 x, int s) { return (x << s) | (x >> (64 - s)); }

// 4 ASCII bases (little-endian word, first base in the low byte) -> 8 bits of 2-bit codes, first base in the
// top pair, C=0 G=1 A=2 T=3 (C/util/CGAT.java:66-69), and a 4-bit validity mask (bit i = base i is one of the
// upper-case letters CGAT; everything else, incl. lower case and N, is invalid: CGAT.java:60-69).
__device__ __forceinline__ void gs_conv4(u32 w, u32& code8, u32& valid4) {
    u32 x = (w >> 1) & 0x03030303u;                                    // A=0 C=1 G=3 T=2
    u32 c = (((~x) & 0x01010101u) << 1) | ((x >> 1) & 0x01010101u);    // C=0 G=1 A=2 T=3
    code8 = (c * 0x40100401u) >> 24;
    u32 eq = __vcmpeq4(w, 0x41414141u) | __vcmpeq4(w, 0x43434343u) | __vcmpeq4(w, 0x47474747u) | __vcmpeq4(w, 0x54545454u);
    valid4 = (((eq & 0x01010101u) * 0x01020408u) >> 24) & 0xFu;
}

// forward k-mer of the window starting at tile-relative position p from the packed big-endian 2-bit stream
__device__ __forceinline__ u64 gs_extract(const u64* cw, int p, int k) {
    // 32-bit view of the big-endian stream: stream word q sits at c32[q ^ 1] (u64 word w = (c32[2w+1] << 32) | c32[2w]).
    // Two funnel shifts give the 64 bits that start at base p (branch-free: a shift of 0 needs no special case).
    const u32* c32 = (const u32*)cw;
    const int q = p >> 4;
    const u32 sh = (u32)(p & 15) * 2;
    const u32 w0 = c32[q ^ 1], w1 = c32[(q + 1) ^ 1], w2 = c32[(q + 2) ^ 1];
    const u32 x1 = __funnelshift_l(w1, w0, sh), x0 = __funnelshift_l(w2, w1, sh);
    return (((u64)x1 << 32) | x0) >> (64 - 2 * k);
}

// reverse complement in 2-bit space: complement = code ^ 1 (C<->G, A<->T; CGAT.java:71-74), order reversed
// (kMerToLongReverse, CGAT.java:245-265)
__device__ __forceinline__ u64 gs_revcomp(u64 fwd, int k) {
    u64 x = fwd ^ 0x5555555555555555ULL;
    u64 r = __brevll(x);
    r = ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);
    return r >> (64 - 2 * k);
}

// standardKMer (CGAT.java:145-147): the larger of the two encodings (both < 2^62, so unsigned == signed order)
__device__ __forceinline__ u64 gs_canonical(u64 fwd, int k) {
    u64 rc = gs_revcomp(fwd, k);
    return fwd > rc ? fwd : rc;
}

// BlockedKMerBloomFilter.containsLong (C/bloom/BlockedKMerBloomFilter.java:181-198).  The second word is only
// loaded when the first one passes (same answer; ~0.89 of the misses stop after one 8-byte load).
__device__ __forceinline__ bool gs_bloom_blocked(const u64* __restrict__ words, u64 buckets, u64 magic, long long seed, u64 key) {
    long long h = seed ^ (long long)key;
    u64 start = gs_absmod(h, buckets, magic);
    u64 h2 = (u64)h ^ gs_rotl64((u64)h, 32);
    u64 a = __ldg(words + start);
    u64 m1 = (1ULL << (h2 & 63)) | (1ULL << ((h2 >> 6) & 63));
    if ((a & m1) != m1) return false;
    u64 b = __ldg(words + start + 1 + (h2 >> 60));
    u64 m2 = (1ULL << ((h2 >> 12) & 63)) | (1ULL << ((h2 >> 18) & 63));
    return (b & m2) == m2;
}

// MurmurHash3DropIn.hash64 (C/util/MurmurHash3DropIn.java:60-87)
__device__ __forceinline__ u64 gs_murmur64(u64 data, u64 seed) {
    u64 hash = seed;
    u64 k = ((data & 0x00ff00ff00ff00ffULL) << 8) | ((data >> 8) & 0x00ff00ff00ff00ffULL);
    k = (k << 48) | ((k & 0xffff0000ULL) << 16) | ((k >> 16) & 0xffff0000ULL) | (k >> 48);
    k *= 0x87c37b91114253d5ULL;
    k = gs_rotl64(k, 31);
    k *= 0x4cf5ad432745937fULL;
    hash ^= k;
    hash = gs_rotl64(hash, 27) * 5 + 0x52dce729ULL;
    hash ^= 8;
    hash ^= hash >> 33;
    hash *= 0xff51afd7ed558ccdULL;
    hash ^= hash >> 33;
    hash *= 0xc4ceb9fe1a85ec53ULL;
    hash ^= hash >> 33;
    return hash ^ data;
}

// KMerSortedArray.getLong (C/store/KMerSortedArray.java:298-349): prefilter, then the position of `key` in the
// sorted array.  The binary search runs inside the bucket of the key's top bits instead of over [0, n): same
// position because keys are distinct and sorted.  Returns the label (value index / MISS) and the position.
__device__ __forceinline__ u32 gs_lookup(const GsDbView& db, u64 key, bool useBloom, u64& pos) {
    if (useBloom && !gs_bloom_blocked(db.bloom, db.bloomBuckets, db.bloomMagic, db.bloomSeed, key)) return GS_LABEL_MISS;
    u64 b = key >> db.bshift;
    if (b >= db.nBuckets) return GS_LABEL_MISS;  // not a 2k-bit value (only reachable through gs_db_lookup)
    u32 lo = __ldg(db.bstart + b), hi = __ldg(db.bstart + b + 1);
    while (lo < hi) {
        u32 mid = (lo + hi) >> 1;
        u64 kv = __ldg(db.keys + mid);
        if (kv < key) lo = mid + 1; else hi = mid;
    }
    if ((u64)lo >= db.n || __ldg(db.keys + lo) != key) return GS_LABEL_MISS;
    pos = lo;
    uint16_t v = __ldg(db.vals + lo);
    return v == GS_VAL_NONODE ? GS_LABEL_MISS : (u32)v;
}

// ---- probe table -------------------------------------------------------------------------------------------
// Measured on B200 (profiles/microbench/randwide.cu, randline.cu): fully divergent loads are capped at ~39.4 G requests/s
// for the whole GPU whatever their width (8, 16 or 32 bytes per lane) and whether DRAM moves 64 or 128 bytes for them.
// The reference's structures need Bloom word(s) + bucket index + keys + value = 3-5 requests per k-mer; this table
// needs one: a bucket is one 32-byte sector of four 8-byte entries, fetched with a single 256-bit load (LDG.E.256).
//   entry = remainder << 24 | displacement << 19 | value << 3 | spill << 2 | occupied << 1 | seen
// bucket = mix62(key) >> rbits, remainder = low rbits bits (mix62 is a bijection, so bucket + remainder identify the key);
// `displacement` = how many buckets behind its home bucket the entry sits (0 for ~98 % of the keys): an entry that was
// pushed into bucket b + t answers only to a lookup that has walked t buckets from ITS home, so a key of another home bucket
// with the same remainder can never be taken for it -- the match is exact, not exact-up-to-2^-rbits;
// `spill` (slot 0 only) = a key of this bucket was pushed to a following bucket; `seen` = unique-k-mer bit of the session
// that leases it (KMerUniqueCounterBits), so counting a k-mer as seen needs no second memory request.
// The placement is a pure function of the key set (gs_kernels.cu "probe table build"): the keys of home bucket b occupy the
// consecutive slots [4b + delta_b, 4b + delta_b + c_b) in ascending remainder order, delta_0 = 0,
// delta_{b+1} = max(0, delta_b + c_b - 4) -- sorted linear probing at slot granularity.  Every process that builds the table
// from the same store therefore gets the same slot ids, which is what lets per-GPU unique-k-mer bitsets (bit = slot id) be
// OR-merged across ranks.  Chains run forward without wrap-around into GS_TAB_PAD_BUCKETS spare buckets behind the last one.
#define GS_TAB_SLOTS 4
#define GS_TAB_SLOT_STRIDE 4    // slot id = bucket * 4 + j: the "storage position" used for unique k-mer counting
#define GS_TAB_PAD_BUCKETS 64   // landing zone behind bucket 2^tbits - 1 (the build fails if a chain would run past it)
#define GS_TAB_MIN_BITS 22      // rbits <= 40 so that remainder + displacement + value + 3 flag bits fit 64 bits
#define GS_TAB_SEEN 1ULL
#define GS_TAB_OCC 2ULL
#define GS_TAB_SPILL 4ULL
#define GS_TAB_VAL_SHIFT 3
#define GS_TAB_DISP_SHIFT 19
#define GS_TAB_DISP_MAX 31      // 5 bits; the build fails (and the caller doubles the table) if a key would sit further from home
#define GS_TAB_REM_SHIFT 24
#define GS_M62 ((1ULL << 62) - 1)

// bijection on [0, 2^62) (xor-shifts and odd multiplications mod 2^62): equal hashes <=> equal keys
__host__ __device__ __forceinline__ u64 gs_mix62(u64 x) {
    x ^= x >> 31;
    x = (x * 0x7FB5D329728EA185ULL) & GS_M62;
    x ^= x >> 27;
    x = (x * 0x81DADEF4BC2DD44DULL) & GS_M62;
    x ^= x >> 33;
    return x;
}

struct GsBucket { u64 e[4]; };

// one 256-bit load of a bucket; .cg: served by L2 so that seen bits set by other SMs are visible
__device__ __forceinline__ GsBucket gs_load_bucket(const u64* tab, u64 b) {
    GsBucket r;
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r.e[0]), "=l"(r.e[1]), "=l"(r.e[2]), "=l"(r.e[3]) : "l"(tab + b * 4));
    return r;
}

// Resolve a lookup whose home bucket is already loaded.  Returns the label (value index / MISS);
// pos = slot id of the match, seen = the slot's seen bit as loaded.
__device__ __forceinline__ u32 gs_table_resolve(const GsDbView& db, u64 h, GsBucket bk, u64& pos, bool& seen) {
    u64 b = h >> db.rbits;
    u64 want = ((h & ((1ULL << db.rbits) - 1)) << GS_TAB_REM_SHIFT) | GS_TAB_OCC;   // displacement 0: the home bucket
    const u64 cmpMask = ~((1ULL << GS_TAB_DISP_SHIFT) - 1) | GS_TAB_OCC;
    u32 lab = GS_LABEL_MISS;
    for (int t = 0;; t++) {  // single exit, slot match by selects: no branch per slot, one structured join for the warp
        int j = -1;
        u64 e = 0;
#pragma unroll
        for (int jj = GS_TAB_SLOTS - 1; jj >= 0; jj--)
            if ((bk.e[jj] & cmpMask) == want) { j = jj; e = bk.e[jj]; }
        if (j >= 0) {
            pos = b * GS_TAB_SLOT_STRIDE + (u64)j;
            seen = e & GS_TAB_SEEN;
            const u32 v = (u32)(e >> GS_TAB_VAL_SHIFT) & 0xFFFFu;
            lab = v == GS_VAL_NONODE ? GS_LABEL_MISS : v;
            break;
        }
        if (!(bk.e[0] & GS_TAB_SPILL) || t == GS_TAB_DISP_MAX) break;  // nothing spilled past this bucket
        b = b + 1;                             // no wrap-around: the pad buckets end every chain
        want += 1ULL << GS_TAB_DISP_SHIFT;     // an entry of this key there carries the distance walked
        bk = gs_load_bucket(db.tab, b);
    }
    return lab;
}

// The same split in two for the label kernel: the first bucket is matched without a branch (every lane of the warp runs it,
// lanes without a pending lookup on bucket 0), the rare continuation into the following buckets is a loop of its own.
// Returns the slot index of the match in bk (or -1) and the matching entry.
__device__ __forceinline__ int gs_table_match(int rbits, u64 h, const GsBucket& bk, u64& e, int t = 0) {
    const u64 want = ((h & ((1ULL << rbits) - 1)) << GS_TAB_REM_SHIFT) | ((u64)t << GS_TAB_DISP_SHIFT) | GS_TAB_OCC;
    const u64 cmpMask = ~((1ULL << GS_TAB_DISP_SHIFT) - 1) | GS_TAB_OCC;
    int j = -1;
    e = 0;
#pragma unroll
    for (int jj = GS_TAB_SLOTS - 1; jj >= 0; jj--)
        if ((bk.e[jj] & cmpMask) == want) { j = jj; e = bk.e[jj]; }
    return j;
}
// (scalars instead of the view: a reference to the kernel parameter block would force a local copy of it; the result comes back
// in registers -- x = slot id, y = label | seen bit << 32 -- so that the caller keeps nothing in local memory for the call)
static __device__ __noinline__ ulonglong2 gs_table_chain(const u64* tab, int rbits, u64 h, u64 b) {
    ulonglong2 r = make_ulonglong2(0ULL, (u64)GS_LABEL_MISS);
    for (int t = 1; t <= GS_TAB_DISP_MAX; t++) {
        b = b + 1;
        const GsBucket bk = gs_load_bucket(tab, b);
        u64 e;
        const int j = gs_table_match(rbits, h, bk, e, t);
        if (j >= 0) {
            const u32 v = (u32)(e >> GS_TAB_VAL_SHIFT) & 0xFFFFu;
            r.x = b * GS_TAB_SLOT_STRIDE + (u64)j;
            r.y = (u64)(v == GS_VAL_NONODE ? GS_LABEL_MISS : v) | ((e & GS_TAB_SEEN) << 32);
            break;
        }
        if (!(bk.e[0] & GS_TAB_SPILL)) break;
    }
    return r;
}

__device__ __forceinline__ u32 gs_lookup_table(const GsDbView& db, u64 key, u64& pos) {
    if (key > GS_M62) return GS_LABEL_MISS;
    const u64 h = gs_mix62(key);
    bool seen;
    return gs_table_resolve(db, h, gs_load_bucket(db.tab, h >> db.rbits), pos, seen);
}

// ---- minimizer prefilter -------------------------------------------------------------------------------------
// Measured (profiles/microbench/randsize.*): the request cap only counts DISTINCT sectors that miss the L2 -- lanes that
// share a sector are served together, and a footprint <= 64 MiB is served by the L2 at ~270 G requests/s.  A k-mer that is
// not in the store can therefore be answered without touching the probe table if a small, shared structure proves it absent:
// the minimizer of a k-mer = the smallest hash among its GS_MZ_W = 9 windows of m = k - 8 bases (hash of the CANONICAL m-mer,
// so both strands of a k-mer give the same value; any deterministic function of the canonical k-mer keeps the lookup exact).
// Neighbouring k-mers of a read share their minimizer (~5 in a row), so a warp's 32 lanes touch ~7 filter words instead of 32
// table sectors, and the set of minimizers of the store is ~5x smaller than the store.  One bit per minimizer hash; no false
// negatives by construction (built from the stored keys with gs_mz_of_key, same hash as the read path).
#define GS_MZ_S 8                  // windows per k-mer minus one (power of two: van Herk blocks = 8-lane shuffle segments)
#define GS_MZ_W (GS_MZ_S + 1)
#define GS_MZ_MIN_K 24             // m = k - 8 >= 16: below that the minimizer space is too small to filter anything

// hash of the canonical form of an m-mer given both strands (x = forward value, r = reverse complement, 2m bits each).
// In the label kernel both come for free: x = top 2m bits of the forward k-mer, r = low 2m bits of its reverse complement.
__device__ __forceinline__ u32 gs_mmer_hash2(u64 x, u64 r) {
    const u64 y = (x > r ? x : r) * 0x9E3779B97F4A7C15ULL;
    u32 h = (u32)(y >> 32);
    h ^= h >> 15; h *= 0x85EBCA77u; h ^= h >> 13;
    return h;
}
__device__ __forceinline__ u32 gs_mmer_hash(u64 x, int m) {
    u64 c = x ^ 0x5555555555555555ULL;
    u64 r = __brevll(c);
    r = ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);
    return gs_mmer_hash2(x, r >> (64 - 2 * m));
}

// Wide variant for stores with more than ~6e8 k-mers: the 32-bit hash space is too small there -- a minimizer hash is a
// minimum of nine, so the store's and the reads' minimizers crowd into the lowest ninth of it and unrelated m-mers collide on
// the hash VALUE itself (measured on the 2e9-k-mer database: 40 % of the absent k-mers passed the prefilter instead of 9 %).
// The order is then taken on 64 bits (the 32-bit hash on top, an independent 32-bit hash below) and the filter bit comes from
// the lower half, which is uniform.
__device__ __forceinline__ u64 gs_mmer_hash2w(u64 x, u64 r) {
    const u64 c = x > r ? x : r;
    const u64 y = c * 0x9E3779B97F4A7C15ULL;
    u32 h = (u32)(y >> 32);
    h ^= h >> 15; h *= 0x85EBCA77u; h ^= h >> 13;
    const u32 l = (u32)((c * 0xD6E8FEB86659FD93ULL) >> 32);
    return ((u64)h << 32) | l;
}
__device__ __forceinline__ u64 gs_mmer_hashw(u64 x, int m) {
    u64 c = x ^ 0x5555555555555555ULL;
    u64 r = __brevll(c);
    r = ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);
    return gs_mmer_hash2w(x, r >> (64 - 2 * m));
}
__device__ __forceinline__ u64 gs_mz_of_key_w(u64 key, int k) {
    const int m = k - GS_MZ_S;
    const u64 mmask = (m >= 32) ? ~0ULL : ((1ULL << (2 * m)) - 1);
    u64 best = ~0ULL;
#pragma unroll
    for (int j = 0; j <= GS_MZ_S; j++) best = min(best, gs_mmer_hashw((key >> (2 * (GS_MZ_S - j))) & mmask, m));
    return best;
}

// minimizer hash of a stored (canonical) k-mer: min over its GS_MZ_W windows (database build, gs_db_lookup self check)
__device__ __forceinline__ u32 gs_mz_of_key(u64 key, int k) {
    const int m = k - GS_MZ_S;
    const u64 mmask = (m >= 32) ? ~0ULL : ((1ULL << (2 * m)) - 1);
    u32 best = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j <= GS_MZ_S; j++) best = min(best, gs_mmer_hash((key >> (2 * (GS_MZ_S - j))) & mmask, m));
    return best;
}

// Bit index of a minimizer hash.  A minimizer hash is the MINIMUM of nine hashes, i.e. its distribution is concentrated near
// zero (density 9(1-x)^8): used as it is, a filter with close to 2^32 bits would be ~9x fuller where both the store's and the
// reads' minimizers fall (measured on the 2e9-k-mer database: 40 % of the absent k-mers passed instead of 9 %).  The odd
// multiplication is a bijection on 32 bits that spreads the small values over the whole range.
__device__ __forceinline__ u32 gs_mz_index(u32 h, u32 mask) { return (h * 0x9E3779B1u) & mask; }
__device__ __forceinline__ bool gs_mz_test(const u64* __restrict__ filter, u32 mask, u32 h) {
    const u32 idx = gs_mz_index(h, mask);
    return (__ldg(filter + (idx >> 6)) >> (idx & 63)) & 1ULL;
}

// Sliding minimum over GS_MZ_W consecutive positions held one per lane: cur = hashes of 32 positions, nxt = hashes of the
// following 32 (only lanes < GS_MZ_S matter).  van Herk: prefix/suffix minima inside 8-lane blocks, then
// min(suffix[i], prefix[i + 8]).  preNxt = block prefix minima of nxt (computed by the caller for the next chunk anyway).
template <typename T>
__device__ __forceinline__ T gs_seg_prefix_min(T v, int lane) {
#pragma unroll
    for (int d = 1; d < GS_MZ_S; d <<= 1) { const T t = __shfl_up_sync(0xFFFFFFFFu, v, d, GS_MZ_S); if ((lane & (GS_MZ_S - 1)) >= d) v = min(v, t); }
    return v;
}
template <typename T>
__device__ __forceinline__ T gs_seg_suffix_min(T v, int lane) {
#pragma unroll
    for (int d = 1; d < GS_MZ_S; d <<= 1) { const T t = __shfl_down_sync(0xFFFFFFFFu, v, d, GS_MZ_S); if ((lane & (GS_MZ_S - 1)) + d < GS_MZ_S) v = min(v, t); }
    return v;
}
template <typename T>
__device__ __forceinline__ T gs_window_min(T sufCur, T preCur, T preNxt, int lane) {
    // position lane + 8 lives in lane + 8 of the current chunk or in lane - 24 of the next one: one rotate of the merged register
    const T merged = lane < GS_MZ_S ? preNxt : preCur;
    return min(sufCur, (T)__shfl_sync(0xFFFFFFFFu, merged, (lane + GS_MZ_S) & 31));
}

// value stored at a "storage position" of the unique-k-mer bitset: sorted-array index (classic) or table slot id
__device__ __forceinline__ u32 gs_value_at(const GsDbView& db, int layout, u64 pos) {
    if (layout == GS_LAYOUT_TABLE) {
        if (pos >= db.tabSlots) return GS_VAL_NONODE;
        const u64 e = __ldcg(db.tab + pos);
        return (e & GS_TAB_OCC) ? (u32)(e >> GS_TAB_VAL_SHIFT) & 0xFFFFu : GS_VAL_NONODE;
    }
    return pos < db.n ? (u32)__ldg(db.vals + pos) : GS_VAL_NONODE;
}

__device__ __forceinline__ bool gs_anc_or_self(const GsDbView& db, int a, int b) {
    int pb = __ldg(db.pre + b);
    return __ldg(db.pre + a) <= pb && pb <= __ldg(db.last + a);
}

// Stage one tile of a read into shared memory: packed 2-bit codes (big-endian u64 stream) and a validity bit
// per base.  `src` = first base of the tile, nb = number of real bases in the tile (bases beyond are invalid).
// Adds this lane's number of invalid bases with tile-relative index < lowLimit to badLow, and sets badTail if it
// saw an invalid base with index >= lowLimit (for the INVALID-iteration count, see gs_match.cu); only bases with
// tile-relative index < countLimit are counted.
__device__ __forceinline__ void gs_stage_tile(const uint8_t* __restrict__ src, int nb, int lane, u32* code32, uint16_t* valid16,
                                              int lowLimit, int countLimit, int& badLow, int& badTail) {
    const int groups = (nb + 15) >> 4;
    const int sh = (int)((uintptr_t)src & 15);
    const uint4* ap = (const uint4*)(src - sh);
    for (int j = lane; j < GS_TILE_GROUPS; j += 32) {
        u32 code = 0, valid = 0;
        if (j < groups) {
            uint4 A = __ldg(ap + j);
            u32 r0, r1, r2, r3;
            if (sh == 0) { r0 = A.x; r1 = A.y; r2 = A.z; r3 = A.w; }
            else {
                uint4 B = __ldg(ap + j + 1);
                const int bs = (sh & 3) * 8;
                u32 w0, w1, w2, w3, w4;
                switch (sh >> 2) {
                    case 0: w0 = A.x; w1 = A.y; w2 = A.z; w3 = A.w; w4 = B.x; break;
                    case 1: w0 = A.y; w1 = A.z; w2 = A.w; w3 = B.x; w4 = B.y; break;
                    case 2: w0 = A.z; w1 = A.w; w2 = B.x; w3 = B.y; w4 = B.z; break;
                    default: w0 = A.w; w1 = B.x; w2 = B.y; w3 = B.z; w4 = B.w; break;
                }
                r0 = __funnelshift_r(w0, w1, bs); r1 = __funnelshift_r(w1, w2, bs);
                r2 = __funnelshift_r(w2, w3, bs); r3 = __funnelshift_r(w3, w4, bs);
            }
            u32 c0, c1, c2, c3, v0, v1, v2, v3;
            gs_conv4(r0, c0, v0); gs_conv4(r1, c1, v1); gs_conv4(r2, c2, v2); gs_conv4(r3, c3, v3);
            code = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
            valid = v0 | (v1 << 4) | (v2 << 8) | (v3 << 12);
            const int rem = nb - j * 16;                       // real bases in this group
            const u32 inRange = rem >= 16 ? 0xFFFFu : ((1u << rem) - 1u);
            valid &= inRange;
            const int cntRem = countLimit - j * 16;            // tiles overlap by 32 bases: count each base once
            const u32 cntMask = cntRem >= 16 ? 0xFFFFu : (cntRem <= 0 ? 0u : ((1u << cntRem) - 1u));
            u32 bad = (~valid) & inRange & cntMask;
            const int lowRem = lowLimit - j * 16;              // bases of this group below lowLimit
            const u32 lowMask = lowRem >= 16 ? 0xFFFFu : (lowRem <= 0 ? 0u : ((1u << lowRem) - 1u));
            badLow += __popc(bad & lowMask);
            badTail |= (bad & ~lowMask) != 0;
        }
        code32[j ^ 1] = code;      // u64 word w = (code32 group 2w << 32) | group 2w+1
        valid16[j] = (uint16_t)valid;
    }
}
