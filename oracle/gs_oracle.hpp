// gs_oracle.hpp -- CPU ORACLE for the Genestrip read-matching hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may build, load or call anything under oracle/.
// The product (genestrip_b200/) never links, imports or falls back to it.
//
// It is a plain, single-threaded C++17 restatement of the reference's (pfeiferd/genestrip v3.0, Java)
// algorithm for the path, statement by statement where the order of evaluation matters.  Every
// function cites the reference file:line it follows, with the abbreviations
//   C/ = core/src/main/java/org/metagene/genestrip/    B/ = base/src/main/java/org/metagene/genestrip/
//   T/ = core/src/test/java/org/metagene/genestrip/
//
// Parity pinning: the reference is 100 % Java and no JDK exists in the build container, so the
// reference itself cannot be executed here.  The oracle is pinned against the reference's own
// known-answer tests and fixtures restated in tests/test_oracle_*.py (see DESIGN.md "Oracle").
// Parity with KrakenUniq (T/goals/refseq/ComprehensiveMatchTest.java) is UNPINNED (needs network).
#pragma once
#include <algorithm>
#include <cassert>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace gso {

typedef int64_t jlong;
typedef int32_t jint;
typedef int16_t jshort;
typedef uint64_t ulong_t;

static inline jlong jshl(jlong v, int s) { return (jlong)((ulong_t)v << (s & 63)); }    // Java <<
static inline jlong jushr(jlong v, int s) { return (jlong)((ulong_t)v >> (s & 63)); }   // Java >>>
static inline jlong jshr(jlong v, int s) { return v >> (s & 63); }                      // Java >>
static inline jlong jrotl(jlong v, int s) { s &= 63; return s == 0 ? v : (jlong)(((ulong_t)v << s) | ((ulong_t)v >> (64 - s))); }
static inline jlong jmul(jlong a, jlong b) { return (jlong)((ulong_t)a * (ulong_t)b); }
static inline jlong jadd(jlong a, jlong b) { return (jlong)((ulong_t)a + (ulong_t)b); }
// Math.abs(v % m) with Java's truncating remainder (C++ % truncates as well).
static inline jlong jabsmod(jlong v, jlong m) { jlong r = v % m; return r < 0 ? -r : r; }

// ---------------------------------------------------------------------------------------------
// java.util.Random (JDK specification of the 48-bit LCG; SURVEY.md §8c check value:
// new Random(42).nextLong() == -5025562857975149833).
// ---------------------------------------------------------------------------------------------
struct JavaRandom {
    jlong seed;
    explicit JavaRandom(jlong s) { seed = (s ^ 0x5DEECE66DLL) & ((1LL << 48) - 1); }
    jint next(int bits) {
        seed = (jlong)(((ulong_t)seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1));
        return (jint)(seed >> (48 - bits));
    }
    jint nextInt() { return next(32); }
    jint nextInt(jint bound) {
        if (bound <= 0) throw std::invalid_argument("bound must be positive");
        jint r = next(31);
        jint m = bound - 1;
        if ((bound & m) == 0) {
            r = (jint)(((jlong)bound * (jlong)r) >> 31);
        } else {
            for (jint u = r; u - (r = u % bound) + m < 0; u = next(31)) {
            }
        }
        return r;
    }
    jlong nextLong() { jlong hi = (jlong)next(32); jlong lo = (jlong)next(32); return jadd(jshl(hi, 32), lo); }
    double nextDouble() { return (double)(((jlong)next(26) << 27) + next(27)) * (1.0 / (double)(1LL << 53)); }
};

// ---------------------------------------------------------------------------------------------
// C/util/CGAT.java:37-265
// ---------------------------------------------------------------------------------------------
namespace cgat {
static const char DECODE_TABLE[4] = {'C', 'G', 'A', 'T'};  // CGAT.java:44
// CGAT.java:60-74; the Java tables have 128 entries and would throw for bytes >= 0x80 -- the
// oracle treats those as invalid (-1) like any other non-CGAT byte.
static inline int jump(uint8_t c) {
    switch (c) { case 'C': return 0; case 'G': return 1; case 'A': return 2; case 'T': return 3; default: return -1; }
}
static inline int revjump(uint8_t c) {
    switch (c) { case 'C': return 1; case 'G': return 0; case 'A': return 3; case 'T': return 2; default: return -1; }
}
static inline int complement(uint8_t c) {  // CGAT.java:52-58
    switch (c) { case 'C': return 'G'; case 'G': return 'C'; case 'A': return 'T'; case 'T': return 'A'; default: return -1; }
}
// CGAT.java:76-82
static inline jlong shiftFilterStraight(int k) { return k == 32 ? -1LL : ~jshl(-1LL, k * 2); }
static inline int shiftFilterReverse(int k) { return (k - 1) * 2; }

// CGAT.java:159-180
static inline jlong kMerToLongStraight(const uint8_t* seq, int start, int k, int* badPos) {
    jlong res = 0;
    if (badPos) *badPos = -1;
    int max = start + k;
    for (int i = start; i < max; i++) {
        res = jrotl(res, 2);
        int c = jump(seq[i]);
        if (c == -1) { if (badPos) *badPos = i; return -1LL; }
        res += c;
    }
    return res;
}
// CGAT.java:245-265
static inline jlong kMerToLongReverse(const uint8_t* seq, int start, int k, int* badPos) {
    jlong res = 0;
    if (badPos) *badPos = -1;
    for (int i = start + k - 1; i >= start; i--) {
        res = jrotl(res, 2);
        int c = revjump(seq[i]);
        if (c == -1) { if (badPos) *badPos = i; return -1LL; }
        res += c;
    }
    return res;
}
// CGAT.java:145-147
static inline jlong standardKMer(jlong straight, jlong reverse) { return straight > reverse ? straight : reverse; }
// CGAT.java:132-136
static inline jlong kMerToLong(const uint8_t* seq, int start, int k, int* badPos) {
    jlong reverse = kMerToLongReverse(seq, start, k, badPos);
    jlong straight = kMerToLongStraight(seq, start, k, badPos);
    return standardKMer(reverse, straight);
}
// CGAT.java:208-214
static inline jlong nextKMerStraight(jlong kmer, uint8_t bp, int k) {
    int c = jump(bp);
    if (c == -1) return -1LL;
    return (jshl(kmer, 2) & shiftFilterStraight(k)) | (jlong)c;
}
// CGAT.java:226-232
static inline jlong nextKMerReverse(jlong kmer, uint8_t bp, int k) {
    int c = revjump(bp);
    if (c == -1) return -1LL;
    return jushr(kmer, 2) | jshl((jlong)c, shiftFilterReverse(k));
}
// CGAT.java:191-197
static inline void longToKMerStraight(jlong kmer, uint8_t* res, int start, int k) {
    for (int i = k - 1; i >= 0; i--) { res[start + i] = DECODE_TABLE[(int)(kmer & 3)]; kmer = jushr(kmer, 2); }
}
// CGAT.java:283-296
static inline void reverse(uint8_t* seq, int start, int k) {
    int end = start + k - 1;
    while (start < end) {
        uint8_t h = (uint8_t)complement(seq[start]);
        seq[start] = (uint8_t)complement(seq[end]);
        seq[end] = h;
        start++; end--;
    }
    if (start == end) seq[start] = (uint8_t)complement(seq[start]);
}
static inline uint8_t cgatToUpperCase(uint8_t c) {  // CGAT.java:91-99
    switch (c) { case 'a': return 'A'; case 'c': return 'C'; case 'g': return 'G'; case 't': return 'T'; default: return c; }
}
}  // namespace cgat

// ---------------------------------------------------------------------------------------------
// C/util/MurmurHash3DropIn.java:60-87
// ---------------------------------------------------------------------------------------------
static inline jlong murmurHash64(jlong data, jlong hashBase) {
    const jlong C1 = (jlong)0x87c37b91114253d5ULL, C2 = (jlong)0x4cf5ad432745937fULL;
    jlong hash = hashBase;
    jlong k = jshl(data & 0x00ff00ff00ff00ffLL, 8) | (jushr(data, 8) & 0x00ff00ff00ff00ffLL);
    k = jshl(k, 48) | jshl(k & 0xffff0000LL, 16) | (jushr(k, 16) & 0xffff0000LL) | jushr(k, 48);
    k = jmul(k, C1);
    k = jrotl(k, 31);
    k = jmul(k, C2);
    hash ^= k;
    hash = jadd(jmul(jrotl(hash, 27), 5), 0x52dce729LL);
    hash ^= 8;  // Long.BYTES
    hash ^= jushr(hash, 33);
    hash = jmul(hash, (jlong)0xff51afd7ed558ccdULL);
    hash ^= jushr(hash, 33);
    hash = jmul(hash, (jlong)0xc4ceb9fe1a85ec53ULL);
    hash ^= jushr(hash, 33);
    return hash ^ data;
}

// ---------------------------------------------------------------------------------------------
// C/bloom/KMerProbFilter.java:66 (interface)
// ---------------------------------------------------------------------------------------------
struct KMerProbFilter {
    virtual ~KMerProbFilter() {}
    virtual bool containsLong(jlong key) const = 0;
    virtual void putLong(jlong key) = 0;
    virtual jlong ensureExpectedSize(jlong n, bool enforceLarge) = 0;
    virtual void clear() = 0;
    virtual int kind() const = 0;  // 0 blocked, 1 xor, 2 murmur
};

// C/bloom/BlockedKMerBloomFilter.java:48-255
struct BlockedKMerBloomFilter : KMerProbFilter {
    int bitsPerKey;
    jlong seed;
    jlong buckets = 0;
    std::vector<jlong> data;  // buckets + 16 + 1 words (:215)
    jlong entries = 0;
    explicit BlockedKMerBloomFilter(int bpk = 10) : bitsPerKey(bpk) { JavaRandom r(42); seed = r.nextLong(); }  // :91-93
    BlockedKMerBloomFilter(int bpk, jlong s) : bitsPerKey(bpk), seed(s) {}
    int kind() const override { return 0; }
    jlong hash(jlong x) const { return seed ^ x; }               // :224-233
    jlong reduce(jlong v) const { return jabsmod(v, buckets); }  // :248-249
    void putLong(jlong key) override {                           // :108-124
        entries++;
        jlong h = hash(key);
        jlong start = reduce(h);
        h = h ^ jrotl(h, 32);
        jlong m1 = jshl(1, (int)h) | jshl(1, (int)jshr(h, 6));
        jlong m2 = jshl(1, (int)jshr(h, 12)) | jshl(1, (int)jshr(h, 18));
        data[(size_t)start] |= m1;
        data[(size_t)(start + 1 + jushr(h, 60))] |= m2;
    }
    // putLong from several threads at once (bench-scale database construction only): same bits, OR-ed atomically
    void putLongShared(jlong key) {
        jlong h = hash(key);
        jlong start = reduce(h);
        h = h ^ jrotl(h, 32);
        jlong m1 = jshl(1, (int)h) | jshl(1, (int)jshr(h, 6));
        jlong m2 = jshl(1, (int)jshr(h, 12)) | jshl(1, (int)jshr(h, 18));
        __atomic_fetch_or(&data[(size_t)start], m1, __ATOMIC_RELAXED);
        __atomic_fetch_or(&data[(size_t)(start + 1 + jushr(h, 60))], m2, __ATOMIC_RELAXED);
    }
    bool containsLong(jlong key) const override {  // :181-198
        jlong h = seed ^ key;
        jlong start = reduce(h);
        h = h ^ jrotl(h, 32);
        jlong a = data[(size_t)start];
        jlong b = data[(size_t)(start + 1 + jushr(h, 60))];
        jlong m1 = jshl(1, (int)h) | jshl(1, (int)jshr(h, 6));
        jlong m2 = jshl(1, (int)jshr(h, 12)) | jshl(1, (int)jshr(h, 18));
        return ((m1 & a) == m1) && ((m2 & b) == m2);
    }
    jlong ensureExpectedSize(jlong entryCount, bool) override {  // :201-219
        entryCount = std::max<jlong>(1, entryCount);
        jlong bits = entryCount * bitsPerKey;
        jlong newSize = (bits + 63) / 64;
        entries = 0;
        if (newSize > buckets) { buckets = newSize; data.assign((size_t)buckets + 16 + 1, 0); }
        return bits;
    }
    void clear() override { std::fill(data.begin(), data.end(), 0); }
};

// C/util/LargeBitVector.java:99-221 (small backing only; the large backing is the same bits in
// 2^27-word segments, fastutil BigArrays)
struct LargeBitVector {
    jlong size = -1;  // words
    std::vector<jlong> bits;
    explicit LargeBitVector(jlong initialSize = 0) { ensureCapacity(initialSize); }
    bool ensureCapacity(jlong newSize) {  // :99-121
        newSize = (newSize + 63) / 64;
        if (newSize > size) { size = newSize; bits.resize((size_t)size, 0); return true; }
        return false;
    }
    void clear() { std::fill(bits.begin(), bits.end(), 0); }
    void set(jlong index) { bits[(size_t)jushr(index, 6)] |= jshl(1, (int)(index & 63)); }           // :161-175
    bool get(jlong index) const { return ((jshr(bits[(size_t)jushr(index, 6)], (int)(index & 63))) & 1LL) == 1; }  // :207-221
    jlong getBitSize() const { return size * 64; }
};

// C/bloom/AbstractKMerBloomFilter.java:72-267, XORKMerBloomFilter.java:43-58, MurmurKMerBloomFilter.java:45-47
struct HashedKMerBloomFilter : KMerProbFilter {
    bool xorHash;
    double fpp;
    JavaRandom random;
    LargeBitVector bitVector;
    jlong bits = 0;
    jlong expectedInsertions = 0;
    int hashes = 0;
    std::vector<jlong> hashFactors;
    jlong entries = 0;
    HashedKMerBloomFilter(double p, bool xorh) : xorHash(xorh), fpp(p), random(42), bitVector(0) {
        if (p <= 0 || p >= 1) throw std::invalid_argument("fpp must be a probability");
    }
    int kind() const override { return xorHash ? 1 : 2; }
    static jlong optimalNumOfBits(jlong n, double p) {  // :178-180
        double v = -(double)n * std::log(p) / (std::log(2.0) * std::log(2.0));
        return std::max<jlong>(1LL, (jlong)v);
    }
    static int optimalNumOfHashFunctions(jlong n, jlong m) {  // :167-169 (Math.round = floor(x+0.5))
        double x = ((double)m) / (double)n * std::log(2.0);
        jlong r = (jlong)std::floor(x + 0.5);
        return (int)std::max<jlong>(1, r);
    }
    jlong ensureExpectedSize(jlong n, bool) override {  // :95-113
        if (n < 0) throw std::invalid_argument("expected insertions must be > 0");
        expectedInsertions = n;
        bits = optimalNumOfBits(n, fpp);
        if (bitVector.ensureCapacity(bits)) {
            hashes = optimalNumOfHashFunctions(n, bits);
            hashFactors.assign((size_t)hashes, 0);
            for (int i = 0; i < hashes; i++) hashFactors[(size_t)i] = random.nextLong();
        }
        return bits;
    }
    jlong hash(jlong data, int i) const { return xorHash ? (hashFactors[(size_t)i] ^ data) : murmurHash64(data, hashFactors[(size_t)i]); }
    jlong reduce(jlong v) const { return jabsmod(v, bits); }  // :265-267
    void putLong(jlong data) override {                       // :183-205
        entries++;
        for (int i = 0; i < hashes; i++) bitVector.set(reduce(hash(data, i)));
    }
    bool containsLong(jlong data) const override {  // :209-216
        for (int i = 0; i < hashes; i++) if (!bitVector.get(reduce(hash(data, i)))) return false;
        return true;
    }
    void clear() override { bitVector.clear(); entries = 0; }
};

// ---------------------------------------------------------------------------------------------
// C/store/KMerStore.java + AbstractKMerStore.java:271-356 + KMerSortedArray.java:168-423.
// Values are tax id strings; value index = order of first registration (getAddValueIndex).
// ---------------------------------------------------------------------------------------------
struct KMerStoreBase {
    int k = 31;
    double optimizedFpp = 0.01;
    bool useFilter = true;
    bool sorted = false;
    jlong entries = 0;
    std::unique_ptr<KMerProbFilter> filter;
    std::vector<std::string> indexMap;  // value index -> value
    std::unordered_map<std::string, int> valueMap;
    bool fillFilterXor = true;
    virtual ~KMerStoreBase() {}
    int getNValues() const { return (int)indexMap.size(); }
    int getIndexForValue(const std::string& v) const { auto it = valueMap.find(v); return it == valueMap.end() ? -1 : it->second; }
    int getAddValueIndex(const std::string& v) {  // AbstractKMerStore.java:231-243 of the stripped view
        auto it = valueMap.find(v);
        if (it != valueMap.end()) return it->second;
        int idx = (int)indexMap.size();
        if (idx >= maxValues()) throw std::runtime_error("Too many different values");
        indexMap.push_back(v);
        valueMap[v] = idx;
        return idx;
    }
    virtual int maxValues() const = 0;
    // returns value index or -1 (null); pos receives the storage position
    virtual int getLong(jlong kmer, jlong* pos) const = 0;
    virtual bool putLong(jlong kmer, const std::string& value) = 0;
    virtual void optimize() = 0;
    virtual void visit(const std::function<void(jlong kmer, int vidx, jlong pos)>& f) const = 0;
    virtual void setIndexAtPosition(jlong pos, int index) = 0;
    std::unique_ptr<KMerProbFilter> createOptimizedFilter() const {  // AbstractKMerStore.java:271-285
        if (optimizedFpp >= 1) return nullptr;
        std::unique_ptr<KMerProbFilter> f;
        if (optimizedFpp == 0.01) f.reset(new BlockedKMerBloomFilter());
        else f.reset(new HashedKMerBloomFilter(optimizedFpp, fillFilterXor));
        f->ensureExpectedSize(entries, false);
        return f;
    }
    // AbstractKMerStore.java:338-356 getNKmersPerTaxid: per value count; null key -> entries
    std::vector<jlong> nKmersPerValueIndex() const {
        std::vector<jlong> c((size_t)getNValues(), 0);
        visit([&](jlong, int vidx, jlong) { c[(size_t)vidx]++; });
        return c;
    }
};

struct KMerSortedArray : KMerStoreBase {
    static const int MAX_VALUES = 65535;  // KMerSortedArray.java:56 (Short range minus sentinel)
    jlong size = 0;
    std::vector<jlong> kmers;
    std::vector<jshort> valueIndexes;  // index + Short.MIN_VALUE
    KMerSortedArray(int k_, double entryFpp, double optFpp, bool xorh) {
        k = k_; optimizedFpp = optFpp; fillFilterXor = xorh;
        if (k < 1 || k > 31) throw std::invalid_argument("k must be in [1, 31]");
        filter.reset(new HashedKMerBloomFilter(entryFpp, xorh));
    }
    int maxValues() const override { return MAX_VALUES; }
    void initSize(jlong s) {  // KMerSortedArray.java:133-152
        filter->ensureExpectedSize(s, false);
        filter->clear();
        size = s;
        kmers.assign((size_t)s, 0);
        valueIndexes.assign((size_t)s, 0);
    }
    bool putLong(jlong kmer, const std::string& value) override {  // :168-202
        sorted = false;
        if (filter->containsLong(kmer)) return false;
        if (entries == size) return false;
        jlong pos = entries++;
        int sindex = getAddValueIndex(value);
        filter->putLong(kmer);
        kmers[(size_t)pos] = kmer;
        setIndexAtPosition(pos, sindex);
        return true;
    }
    void setIndexAtPosition(jlong pos, int index) override { valueIndexes[(size_t)pos] = (jshort)(index + INT16_MIN); }
    int indexAtPosition(jlong pos) const { return (int)valueIndexes[(size_t)pos] - INT16_MIN; }
    static jlong javaBinarySearch(const jlong* a, jlong from, jlong to, jlong key) {  // java.util.Arrays.binarySearch
        jlong low = from, high = to - 1;
        while (low <= high) {
            jlong mid = (jlong)(((ulong_t)low + (ulong_t)high) >> 1);
            jlong midVal = a[mid];
            if (midVal < key) low = mid + 1;
            else if (midVal > key) high = mid - 1;
            else return mid;
        }
        return -(low + 1);
    }
    int getLong(jlong kmer, jlong* posStore) const override {  // :298-349
        if (filter && useFilter && !filter->containsLong(kmer)) return -1;
        jlong pos;
        if (sorted) {
            pos = javaBinarySearch(kmers.data(), 0, entries, kmer);
            if (pos < 0) return -1;
        } else {
            pos = -1;
            for (jlong i = 0; i < entries; i++) if (kmers[(size_t)i] == kmer) { pos = i; break; }
            if (pos < 0) return -1;
        }
        if (posStore) *posStore = pos;
        return indexAtPosition(pos);
    }
    void optimize() override {  // :362-423
        if (sorted) return;
        std::vector<size_t> perm((size_t)entries);
        for (size_t i = 0; i < perm.size(); i++) perm[i] = i;
        std::sort(perm.begin(), perm.end(), [&](size_t a, size_t b) { return kmers[a] < kmers[b]; });
        std::vector<jlong> nk((size_t)size, 0);
        std::vector<jshort> nv((size_t)size, 0);
        for (size_t i = 0; i < perm.size(); i++) { nk[i] = kmers[perm[i]]; nv[i] = valueIndexes[perm[i]]; }
        kmers.swap(nk); valueIndexes.swap(nv);
        sorted = true;
        filter = createOptimizedFilter();
        if (filter) for (jlong i = 0; i < entries; i++) filter->putLong(kmers[(size_t)i]);
    }
    void visit(const std::function<void(jlong, int, jlong)>& f) const override {  // :425-437
        for (jlong i = 0; i < entries; i++) f(kmers[(size_t)i], indexAtPosition(i), i);
    }
};

// C/store/RadixKMerStore.java:102-174, 308-412, 633-730
struct RadixKMerStore : KMerStoreBase {
    int radixBits, remainingBits;
    jint radixMask;
    jlong remainingMask;
    std::vector<std::vector<jlong>> radixIndex;
    std::vector<char> hasBucket;
    std::vector<int> bucketFill;
    std::vector<jlong> bucketOffset;
    RadixKMerStore(int k_, int rbits, const std::vector<int>& bucketSizes, double entryFpp, double optFpp, bool xorh) {
        k = k_; optimizedFpp = optFpp; fillFilterXor = xorh;
        if (rbits < 16 || rbits > 30) throw std::invalid_argument("radixBits out of range");
        radixBits = rbits; radixMask = (1 << rbits) - 1;
        remainingBits = 62 - rbits; remainingMask = (1LL << remainingBits) - 1;
        size_t nb = (size_t)1 << rbits;
        radixIndex.resize(nb); hasBucket.assign(nb, 0); bucketFill.assign(nb, 0); bucketOffset.assign(nb, 0);
        jlong total = 0;
        for (size_t r = 0; r < nb; r++) {
            bucketOffset[r] = total;
            if (bucketSizes[r] > 0) { radixIndex[r].assign((size_t)bucketSizes[r], 0); hasBucket[r] = 1; total += bucketSizes[r]; }
        }
        filter.reset(new HashedKMerBloomFilter(entryFpp, xorh));
        filter->ensureExpectedSize(total, false);
    }
    int maxValues() const override { int vb = std::min(30, 64 - remainingBits); return 1 << vb; }
    static int radixOf(jlong kmer, int rbits) { return (int)(kmer & ((1 << rbits) - 1)); }
    jlong remainingOf(jlong kmer) const { return jushr(kmer, radixBits); }
    jlong entryOf(int vi, jlong remaining) const { return jshl((jlong)vi, remainingBits) | remaining; }
    bool putLong(jlong kmer, const std::string& value) override {  // :324-366
        sorted = false;
        int radix = (int)(kmer & radixMask);
        if (!hasBucket[(size_t)radix]) return false;
        if (filter && useFilter && filter->containsLong(kmer)) return false;
        int fill = bucketFill[(size_t)radix];
        if (fill >= (int)radixIndex[(size_t)radix].size()) return false;
        bucketFill[(size_t)radix] = fill + 1;
        int vi = getAddValueIndex(value);
        if (filter) filter->putLong(kmer);
        entries++;
        radixIndex[(size_t)radix][(size_t)fill] = entryOf(vi, remainingOf(kmer));
        return true;
    }
    int getLong(jlong kmer, jlong* posStore) const override {  // :369-412
        int radix = (int)(kmer & radixMask);
        if (!hasBucket[(size_t)radix]) return -1;  // before the filter (:372-375)
        if (filter && useFilter && !filter->containsLong(kmer)) return -1;
        const std::vector<jlong>& bucket = radixIndex[(size_t)radix];
        jlong remaining = remainingOf(kmer);
        int fill = bucketFill[(size_t)radix];
        int pos = -1;
        if (sorted) {
            int lo = 0, hi = fill - 1;
            while (lo <= hi) {
                int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1);
                jlong midRem = bucket[(size_t)mid] & remainingMask;
                if (midRem < remaining) lo = mid + 1;
                else if (midRem > remaining) hi = mid - 1;
                else { pos = mid; break; }
            }
        } else {
            for (int i = 0; i < fill; i++) if ((bucket[(size_t)i] & remainingMask) == remaining) { pos = i; break; }
        }
        if (pos < 0) return -1;
        if (posStore) *posStore = bucketOffset[(size_t)radix] + pos;
        return (int)jushr(bucket[(size_t)pos], remainingBits);
    }
    void optimize() override {  // :633-675
        if (sorted) return;
        for (size_t r = 0; r < radixIndex.size(); r++) {
            if (!hasBucket[r]) continue;
            int fill = bucketFill[r];
            if (fill > 1) std::sort(radixIndex[r].begin(), radixIndex[r].begin() + fill,
                                    [&](jlong a, jlong b) { return (a & remainingMask) < (b & remainingMask); });
        }
        sorted = true;
        jlong offset = 0;
        for (size_t r = 0; r < radixIndex.size(); r++) { bucketOffset[r] = offset; offset += bucketFill[r]; }
        filter = createOptimizedFilter();
        if (filter) visit([&](jlong kmer, int, jlong) { filter->putLong(kmer); });
    }
    void visit(const std::function<void(jlong, int, jlong)>& f) const override {  // :714-730
        for (size_t r = 0; r < radixIndex.size(); r++) {
            if (!hasBucket[r]) continue;
            int fill = bucketFill[r];
            jlong base = bucketOffset[r];
            for (int i = 0; i < fill; i++) {
                jlong entry = radixIndex[r][(size_t)i];
                jlong kmer = jshl(entry & remainingMask, radixBits) | (jlong)r;
                f(kmer, (int)jushr(entry, remainingBits), base + i);
            }
        }
    }
    void setIndexAtPosition(jlong pos, int index) override {  // :677-686
        size_t lo = 0, hi = bucketOffset.size() - 1, radix = 0;
        while (lo <= hi) {
            size_t mid = (lo + hi) >> 1;
            if (bucketOffset[mid] <= pos) { radix = mid; lo = mid + 1; }
            else { if (mid == 0) break; hi = mid - 1; }
        }
        size_t local = (size_t)(pos - bucketOffset[radix]);
        radixIndex[radix][local] = entryOf(index, radixIndex[radix][local] & remainingMask);
    }
};

// ---------------------------------------------------------------------------------------------
// C/tax/Rank.java:39-122 (names only; ordinal order preserved)
// ---------------------------------------------------------------------------------------------
static const char* const RANK_NAMES[] = {
    "cellular root", "acellular root", "superkingdom", "domain", "realm", "kingdom", "phylum", "subphylum",
    "superclass", "class", "subclass", "superorder", "order", "suborder", "superfamily", "family", "subfamily",
    "tribe", "genus", "subgenus", "species group", "species", "varietas", "subspecies", "serogroup", "biotype",
    "strain", "serotype", "genotype", "forma", "forma specialis", "isolate", "clade", "no rank", "subkingdom",
    "section", "REFINED", "DATA", "FILE", "ID"};
static const int N_RANKS = (int)(sizeof(RANK_NAMES) / sizeof(RANK_NAMES[0]));
static inline int rankFromName(const std::string& s) {
    for (int i = 0; i < N_RANKS; i++) if (s == RANK_NAMES[i]) return i;
    return -1;
}

// ---------------------------------------------------------------------------------------------
// C/tax/TaxTree.java:196-251, 491-502 (nodes.dmp / names.dmp parser, pre-order positions,
// markRequired :536-544) and C/tax/SmallTaxTree.java (compact tree of required nodes).
// ---------------------------------------------------------------------------------------------
struct TaxNode {
    std::string taxId, name;
    bool hasName = false;
    int rank = -1;
    TaxNode* parent = nullptr;
    std::vector<TaxNode*> subNodes;
    int position = 0, depth = 0;
    bool required = false, requested = false;
    int storeIndex = -1;
    // per-consumer vote slots (SmallTaxTree.java:406-407, 643-658)
    std::vector<jint> counts;
    std::vector<jlong> countsInitKeys;
    int getLevel() const { int l = 0; for (const TaxNode* c = parent; c; c = c->parent) l++; return l; }
    void markRequired() { if (required) return; required = true; if (parent) parent->markRequired(); }
    int initPositions(int counter, int d) {
        position = counter; depth = d;
        for (TaxNode* s : subNodes) counter = s->initPositions(counter + 1, d + 1);
        return counter;
    }
    void incCount(int index, jlong initKey, int size) {  // SmallTaxTree.java:643-658
        if (counts.empty()) { counts.assign((size_t)size, 0); countsInitKeys.assign((size_t)size, 0); }
        if (countsInitKeys[(size_t)index] == initKey) counts[(size_t)index]++;
        else { countsInitKeys[(size_t)index] = initKey; counts[(size_t)index] = 1; }
    }
    void resetCounts() {  // :660-671
        for (auto& kx : countsInitKeys) kx = -1;
        for (TaxNode* s : subNodes) s->resetCounts();
    }
};

struct TaxTree {
    std::vector<std::unique_ptr<TaxNode>> pool;
    std::unordered_map<std::string, TaxNode*> byId;
    TaxNode* root = nullptr;
    int countSize = 0;
    TaxNode* getNodeByTaxId(const std::string& id) const { auto it = byId.find(id); return it == byId.end() ? nullptr : it->second; }
    TaxNode* getOrCreate(const std::string& id) {
        TaxNode* n = getNodeByTaxId(id);
        if (!n) { pool.emplace_back(new TaxNode()); n = pool.back().get(); n->taxId = id; byId[id] = n; }
        return n;
    }
    static std::vector<std::string> splitLines(const std::string& text) {
        // BufferedLineReader semantics (B/io/BufferedLineReader.java:160-182): split on '\n' only, NUL bytes dropped.
        std::vector<std::string> lines;
        std::string cur;
        for (char c : text) {
            if (c == 0) continue;
            cur.push_back(c);
            if (c == '\n') { lines.push_back(cur); cur.clear(); }
        }
        if (!cur.empty()) lines.push_back(cur);
        return lines;
    }
    // TaxTree.java:223-251 readNodesFromStream
    void readNodes(const std::string& text) {
        for (const std::string& line : splitLines(text)) {
            size_t a = line.find('|');
            size_t b = a == std::string::npos ? a : line.find('|', a + 1);
            size_t c = b == std::string::npos ? b : line.find('|', b + 1);
            if (a == std::string::npos || b == std::string::npos) continue;
            std::string idA = line.substr(0, a - 1);
            std::string idB = line.substr(a + 2, (b - 1) - (a + 2));
            TaxNode* nodeA = getOrCreate(idA);
            TaxNode* nodeB = getOrCreate(idB);
            if (nodeA != nodeB) { nodeB->subNodes.push_back(nodeA); nodeA->parent = nodeB; }
            std::string rk = c == std::string::npos ? std::string() : line.substr(b + 2, (c - 1) - (b + 2));
            nodeA->rank = rankFromName(rk);
            if (nodeA == nodeB && idA == "1") root = nodeA;
        }
        if (root) root->initPositions(0, 0);
    }
    // TaxTree.java:196-221 readNamesFromStream
    void readNames(const std::string& text) {
        for (const std::string& line : splitLines(text)) {
            bool scientific = line.find("scientific name") != std::string::npos;
            size_t a = line.find('|');
            size_t b = a == std::string::npos ? a : line.find('|', a + 1);
            if (a == std::string::npos || b == std::string::npos) continue;
            TaxNode* node = getNodeByTaxId(line.substr(0, a - 1));
            if (node) {
                std::string name = line.substr(a + 2, b - a - 3);
                if (!node->hasName || scientific) { node->name = name; node->hasName = true; }
            }
        }
    }
    // SmallTaxTree(TaxTree) (SmallTaxTree.java:60-64, 427-454): keeps only required nodes, child order
    // and positions of the full tree; depth re-initialised by initTrie (:673-681).
    std::unique_ptr<TaxTree> toSmallTaxTree() const {
        std::unique_ptr<TaxTree> s(new TaxTree());
        if (!root) return s;
        std::function<TaxNode*(const TaxNode*, TaxNode*, int)> copy = [&](const TaxNode* n, TaxNode* parent, int depth) {
            s->pool.emplace_back(new TaxNode());
            TaxNode* m = s->pool.back().get();
            m->taxId = n->taxId; m->name = n->name; m->hasName = n->hasName; m->rank = n->rank;
            m->position = n->position; m->depth = depth; m->parent = parent; m->storeIndex = -1;
            s->byId[m->taxId] = m;
            for (const TaxNode* c : n->subNodes) if (c->required) m->subNodes.push_back(copy(c, m, depth + 1));
            return m;
        };
        s->root = copy(root, nullptr, 0);
        return s;
    }
    // SmallTaxTree.java:123-151
    void initCountSize(int n) { countSize = n; }
    void resetCounts() { if (root) root->resetCounts(); }
    void incCount(TaxNode* node, int index, jlong initKey) { node->incCount(index, initKey, countSize); }  // :170-172
    int sumCounts(TaxNode* node, int index, jlong initKey) const {  // :184-193
        int res = 0;
        while (node) {
            if (!node->counts.empty() && node->countsInitKeys[(size_t)index] == initKey) res += node->counts[(size_t)index];
            node = node->parent;
        }
        return res;
    }
    TaxNode* lowestNodeWhereSumAboveThreshold(TaxNode* node, int index, jlong initKey, int threshold) const {  // :208-221
        int res = 0;
        while (node) {
            if (!node->counts.empty() && node->countsInitKeys[(size_t)index] == initKey) {
                res += node->counts[(size_t)index];
                if (res >= threshold) return node;
            }
            node = node->parent;
        }
        return nullptr;
    }
    static bool isAncestorOf(TaxNode* node, const TaxNode* ancestor) {  // :242-252
        while (node) { if (node == ancestor) return true; node = node->parent; }
        return false;
    }
    static TaxNode* getLowestCommonAncestor(TaxNode* node1, TaxNode* node2) {  // :263-289
        if (node1 == node2) return node1;
        if (!node1 || !node2) return nullptr;
        TaxNode* a = node1; TaxNode* b = node2;
        while (a->depth > b->depth) a = a->parent;
        while (b->depth > a->depth) b = b->parent;
        while (a != b) { a = a->parent; b = b->parent; }
        return a;
    }
    // pre-order iteration (SmallTaxTree.java:335-373)
    void preorder(const std::function<void(TaxNode*)>& f) const {
        std::function<void(TaxNode*)> rec = [&](TaxNode* n) { f(n); for (TaxNode* c : n->subNodes) rec(c); };
        if (root) rec(root);
    }
};

// ---------------------------------------------------------------------------------------------
// C/store/Database.java:107-143
// ---------------------------------------------------------------------------------------------
struct Database {
    std::unique_ptr<KMerStoreBase> store;
    std::unique_ptr<TaxTree> taxTree;   // SmallTaxTree
    std::vector<TaxNode*> nodeByValueIndex;  // convertKMerStore(): value -> node or nullptr
    std::vector<jlong> dbKmersPerValueIndex;  // store.getFixedNKmersPerTaxid()
    void initStoreIndices() {  // :107-128
        std::function<void(TaxNode*)> rec = [&](TaxNode* n) {
            if (!n) return;
            n->storeIndex = store->getAddValueIndex(n->taxId);
            for (TaxNode* c : n->subNodes) rec(c);
        };
        rec(taxTree->root);
    }
    void convert() {  // :136-143
        nodeByValueIndex.assign((size_t)store->getNValues(), nullptr);
        for (int i = 0; i < store->getNValues(); i++) nodeByValueIndex[(size_t)i] = taxTree->getNodeByTaxId(store->indexMap[(size_t)i]);
    }
    void fix() { dbKmersPerValueIndex = store->nKmersPerValueIndex(); }
};

// ---------------------------------------------------------------------------------------------
// C/store/KMerUniqueCounterBits.java:60-199
// ---------------------------------------------------------------------------------------------
struct KMerUniqueCounterBits {
    const KMerStoreBase* store;
    LargeBitVector bitVector;
    std::vector<jshort> counts;
    bool withCounts;
    KMerUniqueCounterBits(const KMerStoreBase* s, bool wc) : store(s), bitVector(s->entries), withCounts(wc) {
        if (wc) counts.assign((size_t)s->entries, 0);
    }
    void clear() { bitVector.clear(); std::fill(counts.begin(), counts.end(), 0); }
    void putInlined(jlong index) {  // :117-143
        bitVector.set(index);
        if (withCounts) counts[(size_t)index] = (jshort)(counts[(size_t)index] + 1);
    }
    std::vector<jlong> getUniqueKmerCounts() const {  // :146-163 (by value index)
        std::vector<jlong> vc((size_t)store->getNValues(), 0);
        store->visit([&](jlong, int index, jlong i) { if (bitVector.get(i)) vc[(size_t)index]++; });
        return vc;
    }
    static void updateMaxCounts(jshort count, std::vector<jshort>& target) {  // :201-211
        for (size_t j = 0; j < target.size(); j++) {
            if (count > target[j]) {
                for (size_t k = target.size() - 1; k > j; k--) target[k] = target[k - 1];
                target[j] = count;
                return;
            }
        }
    }
    // :173-199; key -1 = total (the null key); only value indices that were hit get an entry
    std::map<int, std::vector<jshort>> getMaxCountsCounts(int n) const {
        std::map<int, std::vector<jshort>> res;
        res[-1] = std::vector<jshort>((size_t)n, 0);
        store->visit([&](jlong, int index, jlong i) {
            if (bitVector.get(i)) {
                auto it = res.find(index);
                if (it == res.end()) it = res.emplace(index, std::vector<jshort>((size_t)n, 0)).first;
                jshort c = counts[(size_t)i];
                updateMaxCounts(c, it->second);
                updateMaxCounts(c, res[-1]);
            }
        });
        return res;
    }
};

// ---------------------------------------------------------------------------------------------
// C/match/CountsPerTaxid.java:127-181, 593-622
// ---------------------------------------------------------------------------------------------
struct CountsPerTaxid {
    int level = 0;
    std::string taxid;
    bool nullTaxid = false;
    jlong reads = 0, reads1KMer = 0, readsBPs = 0, readsKmers = 0, uniqueKmers = 0, kmers = 0;
    jint contigs = 0;
    jlong contigLenSquaredSum = 0;
    jint maxContigLen = 0;
    std::string maxContigDescriptor;  // bytes up to the 0 terminator
    size_t maxContigDescriptorCap = 0;
    bool hasMaxKMerCounts = false;
    std::vector<jshort> maxKMerCounts;
    double errorSum = 0, errorSquaredSum = 0, classErrorSum = 0, classErrorSquaredSum = 0;
    // completed values
    int pos = 0;
    std::string name; bool hasNode = false;
    int rank = -1;
    jlong dbKMers = 0;
    std::string parentTaxId; bool hasParentTaxId = false;
    bool hasExtended = false;
    jlong acc[5] = {0, 0, 0, 0, 0};
    double accNorm[5] = {0, 0, 0, 0, 0};
    double accErrorSum = 0, accErrorSquaredSum = 0, accClassErrorSum = 0, accClassErrorSquaredSum = 0;
    // ValueType order: READS, KMERS, READS_BPS, READS_1KMER, READS_KMERS (CountsPerTaxid.java:43-53)
    jlong valueFor(int t) const {
        switch (t) { case 0: return reads; case 1: return kmers; case 2: return readsBPs; case 3: return reads1KMer; default: return readsKmers; }
    }
    void setDescriptorFromHeader(const uint8_t* desc, int descSize) {  // FastqKMerMatcher.java:404-408
        std::string d;
        int j = 1;
        for (; j < descSize && (size_t)j < maxContigDescriptorCap && desc[j] != ' '; j++) d.push_back((char)desc[j]);
        maxContigDescriptor = d;
    }
    void completeValues(int p, jlong dbk, const TaxNode* node) {  // :593-612
        pos = p; dbKMers = dbk;
        if (node) {
            hasNode = true;
            name = node->name; rank = node->rank;
            parentTaxId = node->parent ? node->parent->taxId : std::string(); hasParentTaxId = true;
            for (int i = 0; i < 5; i++) {
                jlong v = valueFor(i);
                acc[i] = v;
                accNorm[i] = dbk > 0 ? ((double)v) / (double)dbk : 0;
            }
            hasExtended = true;
            accErrorSum = errorSum; accErrorSquaredSum = errorSquaredSum;
            accClassErrorSum = classErrorSum; accClassErrorSquaredSum = classErrorSquaredSum;
        } else {
            name = "TOTAL";
        }
    }
    void accumulateFrom(const CountsPerTaxid& o) {  // :614-622
        for (int i = 0; i < 5; i++) { acc[i] += o.acc[i]; accNorm[i] += o.accNorm[i]; }
        accErrorSum += o.accErrorSum; accErrorSquaredSum += o.accErrorSquaredSum;
        accClassErrorSum += o.accClassErrorSum; accClassErrorSquaredSum += o.accClassErrorSquaredSum;
    }
};

// Java's Double.toString (JDK >= 19: shortest decimal that round-trips, Raffaello Giulietti's
// algorithm; formatting rules of java.lang.Double#toString).
static inline std::string javaDoubleToString(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "Infinity" : "-Infinity";
    if (v == 0) return std::signbit(v) ? "-0.0" : "0.0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
    std::string s(buf, r.ptr);  // [-]d[.ddd]e[+-]XX
    bool neg = s[0] == '-';
    if (neg) s = s.substr(1);
    size_t e = s.find('e');
    std::string mant = s.substr(0, e);
    int exp10 = std::stoi(s.substr(e + 1));
    std::string digits;
    for (char c : mant) if (c != '.') digits.push_back(c);
    std::string out;
    if (exp10 >= -3 && exp10 < 7) {
        if (exp10 >= 0) {
            std::string ip = digits.substr(0, std::min(digits.size(), (size_t)exp10 + 1));
            while (ip.size() < (size_t)exp10 + 1) ip.push_back('0');
            std::string fp = digits.size() > (size_t)exp10 + 1 ? digits.substr((size_t)exp10 + 1) : "0";
            out = ip + "." + fp;
        } else {
            out = "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
        }
    } else {
        std::string fp = digits.size() > 1 ? digits.substr(1) : "0";
        out = digits.substr(0, 1) + "." + fp + "E" + std::to_string(exp10);
    }
    return neg ? "-" + out : out;
}

// ---------------------------------------------------------------------------------------------
// Read entry (C/fastq/AbstractFastqReader.java:494-604, C/match/FastqKMerMatcher.java:618-757)
// ---------------------------------------------------------------------------------------------
struct ReadEntry {
    jlong readNo = 0;
    std::vector<uint8_t> readDescriptor; int readDescriptorSize = 0;
    std::vector<uint8_t> read; int readSize = 0;
    std::vector<uint8_t> readProbs; int readProbsSize = 0; bool hasProbs = false;
    // matcher part
    std::string buffer; bool bufferNull = true;
    int usedPaths = 0;
    std::vector<TaxNode*> readTaxIdNode;
    std::vector<jint> counts;
    jlong indexPos = 0;
    TaxNode* classNode = nullptr;
    // results kept for the parity tests (not in the reference)
    int outReadTaxErrorCount = 0, outReadKmers = 0, outClassErrC = 0; bool outAccepted = false;
    void init(int maxReadSizeBytes, bool withProbs, int paths) {
        readDescriptor.assign((size_t)maxReadSizeBytes, 0);
        read.assign((size_t)maxReadSizeBytes, 0);
        hasProbs = withProbs;
        if (withProbs) readProbs.assign((size_t)maxReadSizeBytes, 0);
        readTaxIdNode.assign((size_t)paths, nullptr);
        counts.assign((size_t)paths, 0);
    }
    void printChar(char c) { bufferNull = false; buffer.push_back(c); }
    void printString(const std::string& s) { bufferNull = false; buffer += s; }
    void printInt(jint v) { bufferNull = false; buffer += std::to_string(v); }  // ByteArrayUtil.intToByteArray (B/util/ByteArrayUtil.java:243-273)
    // AbstractFastqReader.java:570-584
    void write(std::string& out) const {
        out.append((const char*)readDescriptor.data(), (size_t)readDescriptorSize);
        out.push_back('\n');
        out.append((const char*)read.data(), (size_t)readSize);
        out.push_back('\n');
        out.append("+\n");
        if (hasProbs && readProbsSize >= 0) out.append((const char*)readProbs.data(), (size_t)readProbsSize);
        else out.append((size_t)readSize, '~');
        out.push_back('\n');
    }
    // FastqKMerMatcher.java:723-756
    void writeMatchDetails(std::string& out) const {
        if (bufferNull) return;
        out += classNode == nullptr ? "U\t" : "C\t";
        int index = -1;
        for (int i = 1; i < readDescriptorSize; i++) if (readDescriptor[(size_t)i] == ' ') { index = i; break; }
        int end = index == -1 ? readDescriptorSize : index;
        out.append((const char*)readDescriptor.data() + 1, (size_t)(end - 1));
        out.push_back('\t');
        if (classNode == nullptr) out.push_back('0'); else out += classNode->taxId;
        out.push_back('\t');
        out += std::to_string(readSize);
        out.push_back('\t');
        out += buffer;
        out.push_back('\n');
    }
};

struct MatchConfig {
    int k = 31;
    bool classify = true;             // taxTree != null
    int maxPaths = 10;                // maxClassificationPaths
    double maxReadTaxErrorCount = -1;
    double maxReadClassErrorCount = -1;
    int threshold = 1;                // minKMersForClass
    int maxKmerResCounts = 0;
    bool countUnique = true;
    bool writeAll = true;
    bool writeKraken = false;
    bool writeFiltered = false;
    bool withProbs = false;
    int initialReadSize = 4096;
};

// ---------------------------------------------------------------------------------------------
// C/match/FastqKMerMatcher.java:181-611 (threads = 0: one consumer slot, index 0)
// ---------------------------------------------------------------------------------------------
struct FastqKMerMatcher {
    Database* db;
    MatchConfig cfg;
    int k;
    TaxTree* taxTree;  // null when classification is off
    std::vector<std::unique_ptr<CountsPerTaxid>> statsIndex;
    std::vector<jlong> readNoRow;  // readNoPerCPerStat[0]
    KMerUniqueCounterBits* uniqueCounter = nullptr;
    uint64_t* sharedBits = nullptr;  // multi-threaded CPU baseline only: bitset shared by all consumers (atomic OR)
    std::string* krakenOut = nullptr;   // `out`
    std::string* filteredOut = nullptr; // `indexed`
    jlong totalReads = 0, totalKMers = 0, totalBPs = 0;
    // optional per-position dump for the parity tests: label per k-mer position
    // (>=0 value index, -1 miss, -2 invalid) and store position (or -1)
    std::vector<jint>* dumpLabels = nullptr;
    std::vector<jlong>* dumpPos = nullptr;
    static TaxNode* INVALID_NODE() { static TaxNode n; return &n; }

    FastqKMerMatcher(Database* d, const MatchConfig& c) : db(d), cfg(c), k(d->store->k) {
        taxTree = c.classify ? d->taxTree.get() : nullptr;
        statsIndex.resize((size_t)d->store->getNValues());
        readNoRow.assign((size_t)d->store->getNValues(), -1);
        if (taxTree) taxTree->initCountSize(1);
    }
    void initStats() { for (auto& s : statsIndex) s.reset(); }
    void beginFile() {  // readFastq (:238-252)
        if (taxTree) taxTree->resetCounts();
        std::fill(readNoRow.begin(), readNoRow.end(), -1);
    }
    CountsPerTaxid* getCountsPerTaxid(const TaxNode* node, int vi) {  // :545-556
        if (!statsIndex[(size_t)vi]) {
            statsIndex[(size_t)vi].reset(new CountsPerTaxid());
            statsIndex[(size_t)vi]->level = node->getLevel();
            statsIndex[(size_t)vi]->taxid = node->taxId;
            statsIndex[(size_t)vi]->maxContigDescriptorCap = (size_t)cfg.initialReadSize;
        }
        return statsIndex[(size_t)vi].get();
    }
    void mergeReadTaxidPath(TaxNode* node, ReadEntry& entry) {  // :568-586
        bool found = false;
        for (int i = 0; i < entry.usedPaths; i++) {
            if (TaxTree::isAncestorOf(node, entry.readTaxIdNode[(size_t)i])) { entry.readTaxIdNode[(size_t)i] = node; found = true; break; }
            else if (TaxTree::isAncestorOf(entry.readTaxIdNode[(size_t)i], node)) { found = true; break; }
        }
        if (!found && entry.usedPaths < cfg.maxPaths) { entry.readTaxIdNode[(size_t)entry.usedPaths] = node; entry.usedPaths++; }
    }
    void printKrakenStyleOut(ReadEntry& entry, const TaxNode* taxid, int contigLen, int state) {  // :597-611
        if (state != 0) entry.printChar(' ');
        if (taxid == INVALID_NODE()) entry.printChar('A');
        else if (taxid == nullptr) entry.printChar('0');
        else entry.printString(taxid->taxId);
        entry.printChar(':');
        entry.printInt(contigLen);
    }
    void flushContig(CountsPerTaxid* stats, int contigLen, const ReadEntry& entry) {  // :396-410 / :458-471
        stats->kmers += contigLen;
        stats->contigs++;
        stats->contigLenSquaredSum += ((jlong)contigLen) * contigLen;
        if (contigLen > stats->maxContigLen) {
            stats->maxContigLen = contigLen;
            stats->setDescriptorFromHeader(entry.readDescriptor.data(), entry.readDescriptorSize);
        }
    }
    // :269-285
    void nextEntry(ReadEntry& e) {
        e.buffer.clear();  // bufferPos = 0; `buffer` itself stays non-null once anything was printed into this entry
        e.usedPaths = 0; e.classNode = nullptr;
        for (int i = 0; i < cfg.maxPaths; i++) { e.readTaxIdNode[(size_t)i] = nullptr; e.counts[(size_t)i] = 0; }
        e.outReadTaxErrorCount = 0; e.outReadKmers = 0; e.outClassErrC = 0; e.outAccepted = false;
        bool found = matchRead(e, 0);
        // afterMatch (:304-315)
        if (found && filteredOut) e.write(*filteredOut);
        if (krakenOut) { if (cfg.writeAll || e.classNode != nullptr) e.writeMatchDetails(*krakenOut); }
    }
    // :327-535
    bool matchRead(ReadEntry& entry, int index) {
        bool found = false;
        int prints = 0;
        int readTaxErrorCount = taxTree == nullptr ? -1 : 0;
        TaxNode* taxIdNode;
        int max = entry.readSize - k + 1;
        double maxReadTaxErrorCount = cfg.maxReadTaxErrorCount;
        double maxReadTaxErrorCountTimesMax = maxReadTaxErrorCount * max;
        TaxNode* lastTaxid = nullptr;
        int contigLen = 0;
        CountsPerTaxid* stats = nullptr;
        const uint8_t* read = entry.read.data();
        jlong kmer = -1, reverseKmer = -1;
        int oldIndex = 0;
        int badPos = -1;
        for (int i = 0; i < max; i++) {
            int labelFrom = i;
            if (kmer == -1) {
                kmer = cgat::kMerToLongStraight(read, i, k, &badPos);
                if (kmer == -1) { oldIndex = i; i = badPos; }
                else reverseKmer = cgat::kMerToLongReverse(read, i, k, nullptr);
            } else {
                uint8_t lastBase = read[i + k - 1];
                kmer = cgat::nextKMerStraight(kmer, lastBase, k);
                if (kmer == -1) { oldIndex = i; i += k - 1; }
                else reverseKmer = cgat::nextKMerReverse(reverseKmer, lastBase, k);
            }
            if (kmer == -1) taxIdNode = INVALID_NODE();
            else {
                int vi = db->store->getLong(cgat::standardKMer(kmer, reverseKmer), &entry.indexPos);
                taxIdNode = vi < 0 ? nullptr : db->nodeByValueIndex[(size_t)vi];
            }
            if (dumpLabels) {
                int to = taxIdNode == INVALID_NODE() ? std::min(i, max - 1) : labelFrom;
                for (int p = labelFrom; p <= to; p++) {
                    dumpLabels->push_back(taxIdNode == INVALID_NODE() ? -2 : (taxIdNode == nullptr ? -1 : taxIdNode->storeIndex));
                    dumpPos->push_back((taxIdNode == INVALID_NODE() || taxIdNode == nullptr) ? -1 : entry.indexPos);
                }
            }
            const bool newContig = taxIdNode != lastTaxid;
            if (readTaxErrorCount != -1) {
                if (taxIdNode == nullptr || taxIdNode == INVALID_NODE()) {
                    readTaxErrorCount++;
                    if (maxReadTaxErrorCount >= 0) {
                        if ((maxReadTaxErrorCount >= 1 && readTaxErrorCount > maxReadTaxErrorCount)
                            || (readTaxErrorCount > maxReadTaxErrorCountTimesMax)) {
                            readTaxErrorCount = -1;
                        }
                    }
                } else {
                    taxTree->incCount(taxIdNode, index, entry.readNo);
                    if (newContig) mergeReadTaxidPath(taxIdNode, entry);
                }
            }
            if (taxIdNode != lastTaxid) {
                if (contigLen > 0) {
                    if (krakenOut) printKrakenStyleOut(entry, lastTaxid, contigLen, prints++);
                    if (stats != nullptr) flushContig(stats, contigLen, entry);
                    contigLen = 0;
                }
            }
            if (taxIdNode == INVALID_NODE()) contigLen += i >= max ? max - oldIndex : i - oldIndex + 1;
            else contigLen++;
            lastTaxid = taxIdNode;
            if (taxIdNode != nullptr && taxIdNode != INVALID_NODE()) {
                found = true;
                if (newContig) {
                    int vi = taxIdNode->storeIndex;
                    stats = getCountsPerTaxid(taxIdNode, vi);
                    if (readNoRow[(size_t)vi] != entry.readNo) { readNoRow[(size_t)vi] = entry.readNo; stats->reads1KMer++; }
                }
                if (uniqueCounter) uniqueCounter->putInlined(entry.indexPos);
                else if (sharedBits) __atomic_fetch_or(&sharedBits[(size_t)(entry.indexPos >> 6)], 1ULL << (entry.indexPos & 63), __ATOMIC_RELAXED);
            } else {
                stats = nullptr;
            }
        }
        if (contigLen > 0 && krakenOut) printKrakenStyleOut(entry, lastTaxid, contigLen, prints);
        entry.outReadTaxErrorCount = readTaxErrorCount;
        if (found) {
            if (contigLen > 0 && stats != nullptr) flushContig(stats, contigLen, entry);
            if (readTaxErrorCount != -1) {
                int ties = 0;
                for (int i = 0; i < entry.usedPaths; i++) {
                    int sum = taxTree->sumCounts(entry.readTaxIdNode[(size_t)i], index, entry.readNo);
                    if (sum > entry.counts[0]) { entry.counts[0] = sum; entry.readTaxIdNode[0] = entry.readTaxIdNode[(size_t)i]; ties = 0; }
                    else if (sum == entry.counts[0]) { ties++; entry.counts[(size_t)ties] = sum; entry.readTaxIdNode[(size_t)ties] = entry.readTaxIdNode[(size_t)i]; }
                }
                if (cfg.threshold > 1) {
                    for (int i = 0; i <= ties; i++)
                        entry.readTaxIdNode[(size_t)i] = taxTree->lowestNodeWhereSumAboveThreshold(entry.readTaxIdNode[(size_t)i], index, entry.readNo, cfg.threshold);
                }
                TaxNode* node = entry.readTaxIdNode[0];
                for (int i = 1; i <= ties; i++) node = TaxTree::getLowestCommonAncestor(node, entry.readTaxIdNode[(size_t)i]);
                entry.classNode = node;
                if (node == nullptr) return false;
                int readKmers = (ties > 0 || cfg.threshold > 1) ? taxTree->sumCounts(entry.readTaxIdNode[0], index, entry.readNo) : entry.counts[0];
                int classErrC = max - readKmers;
                entry.outReadKmers = readKmers; entry.outClassErrC = classErrC;
                double mrc = cfg.maxReadClassErrorCount;
                if (mrc < 0 || (mrc >= 1 && classErrC <= mrc) || (classErrC <= mrc * max)) {
                    double err = ((double)readTaxErrorCount) / max;
                    double classErr = ((double)classErrC) / max;
                    entry.classNode = node;
                    entry.outAccepted = true;
                    int vi = node->storeIndex;
                    if (vi >= 0) {
                        stats = getCountsPerTaxid(node, vi);
                        stats->reads++;
                        stats->readsKmers += readKmers;
                        stats->readsBPs += entry.readSize;
                        stats->errorSum += err;
                        stats->errorSquaredSum += err * err;
                        stats->classErrorSum += classErr;
                        stats->classErrorSquaredSum += classErr * classErr;
                    }
                }
            }
        }
        return found;
    }
};

// ---------------------------------------------------------------------------------------------
// B/io/BufferedLineReader.java:114-182 over an in-memory byte string
// ---------------------------------------------------------------------------------------------
struct BufferedLineReader {
    const uint8_t* data; size_t n; size_t pos = 0; bool eof = false;
    BufferedLineReader(const uint8_t* d, size_t len) : data(d), n(len) {}
    // returns bytes written incl. the '\n'; target.size()+1 if target filled up before end of line
    int nextLine(std::vector<uint8_t>& target, int startPos = 0) {
        int size = startPos;
        if (eof) return size;
        uint8_t c = 0xFF;
        for (; size < (int)target.size() && pos < n && c != '\n'; pos++) {
            target[(size_t)size] = c = data[pos];
            if (c != 0) size++;
        }
        if (c == '\n') return size;
        if (size == (int)target.size()) return size + 1;
        eof = true;  // stream.read returned -1
        return size;
    }
    int skipLine() {
        int size = 0;
        if (eof) return size;
        uint8_t c = 0xFF;
        for (; pos < n && c != '\n'; pos++) { c = data[pos]; if (c != 0) size++; }
        if (c == '\n') return size;
        eof = true;
        return size;
    }
};

// ---------------------------------------------------------------------------------------------
// C/fastq/AbstractFastqReader.java:288-438 (threads = 0: nextEntry inline), :593-604 growReadBuffer
// ---------------------------------------------------------------------------------------------
struct FastqReader {
    int k;
    jlong reads = 0, kMers = 0, readBPs = 0;
    std::function<void(ReadEntry&)> nextEntry;
    static void growReadBuffer(ReadEntry& e, BufferedLineReader& lr) {
        while (e.readSize == (int)e.read.size()) {
            size_t oldLen = e.read.size();
            e.read.resize(oldLen * 2, 0);
            e.readSize = lr.nextLine(e.read, (int)oldLen) - 1;
        }
        if (e.hasProbs) e.readProbs.assign(e.read.size(), 0);
    }
    // one reusable entry is enough for threads=0; a second one is used for FASTA look-ahead
    void readFastq(const uint8_t* bytes, size_t len, ReadEntry& e) {  // :288-368
        reads = 0; kMers = 0; readBPs = 0;
        BufferedLineReader lr(bytes, len);
        std::vector<uint8_t> plusLine(64);  // :48-49 scratch for the rest of an over-long '+' line
        for (e.readDescriptorSize = lr.nextLine(e.readDescriptor) - 1; e.readDescriptorSize >= 0;
             e.readDescriptorSize = lr.nextLine(e.readDescriptor) - 1) {
            e.readDescriptor[(size_t)e.readDescriptorSize] = 0;
            e.readSize = lr.nextLine(e.read) - 1;
            if (e.readSize == (int)e.read.size()) growReadBuffer(e, lr);
            int newSize = lr.nextLine(e.read, e.readSize) - 1;
            while (e.read[(size_t)e.readSize] != '+') {
                e.readSize = newSize;
                if (newSize == (int)e.read.size()) growReadBuffer(e, lr);
                newSize = lr.nextLine(e.read, e.readSize) - 1;
            }
            e.read[(size_t)e.readSize] = 0;
            if (newSize == (int)e.read.size()) { while (lr.nextLine(plusLine) == (int)plusLine.size() + 1) {} }
            if (e.hasProbs) {
                int readProbsSize = lr.nextLine(e.readProbs) - 1;
                while (readProbsSize < e.readSize) {
                    int oldSize = readProbsSize;
                    readProbsSize = lr.nextLine(e.readProbs, readProbsSize) - 1;
                    if (readProbsSize == oldSize - 1) break;
                }
                e.readProbsSize = readProbsSize;
                e.readProbs[(size_t)readProbsSize] = 0;
            } else {
                int readProbsSize = lr.skipLine() - 1;
                while (readProbsSize < e.readSize) {
                    int oldSize = readProbsSize;
                    readProbsSize = readProbsSize + lr.skipLine() - 1;
                    if (readProbsSize == oldSize - 1) break;
                }
                e.readProbsSize = readProbsSize;
            }
            e.readNo = reads;
            reads++;
            if (e.readSize >= k) kMers += e.readSize - k + 1;
            readBPs += e.readSize;
            nextEntry(e);
        }
    }
    void readFasta(const uint8_t* bytes, size_t len, ReadEntry& e1, ReadEntry& e2) {  // :375-438
        reads = 0; kMers = 0; readBPs = 0;
        BufferedLineReader lr(bytes, len);
        ReadEntry* rs = &e1;
        ReadEntry* other = &e2;
        rs->readDescriptorSize = lr.nextLine(rs->readDescriptor) - 1;
        while (rs != nullptr && rs->readDescriptorSize >= 0) {
            rs->readDescriptor[(size_t)rs->readDescriptorSize] = 0;
            rs->readDescriptor[0] = '@';
            rs->readSize = 0;
            int newSize = lr.nextLine(rs->read, rs->readSize) - 1;
            while (rs->read[(size_t)rs->readSize] != '>' && newSize > rs->readSize - 1) {
                rs->readSize = newSize;
                if (newSize == (int)rs->read.size()) growReadBuffer(*rs, lr);
                newSize = lr.nextLine(rs->read, rs->readSize) - 1;
            }
            ReadEntry* rs2 = nullptr;
            if (newSize != rs->readSize - 1) {
                rs2 = other;
                int len2 = newSize - rs->readSize;
                if ((int)rs2->readDescriptor.size() < len2) rs2->readDescriptor.resize((size_t)len2);
                std::memcpy(rs2->readDescriptor.data(), rs->read.data() + rs->readSize, (size_t)len2);
                if (newSize == (int)rs->read.size()) rs2->readDescriptorSize = lr.nextLine(rs2->readDescriptor, len2);
                else rs2->readDescriptorSize = len2;
            }
            rs->read[(size_t)rs->readSize] = 0;
            rs->readNo = reads;
            reads++;
            if (rs->readSize >= k) kMers += rs->readSize - k + 1;
            readBPs += rs->readSize;
            rs->readProbsSize = -1;
            nextEntry(*rs);
            other = rs;
            rs = rs2;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// C/match/MatchingResult.java:65-118 + C/match/ResultReporter.java:190-279
// ---------------------------------------------------------------------------------------------
struct MatchingResult {
    int k;
    CountsPerTaxid globalStats;
    std::vector<CountsPerTaxid> rows;  // after completeResults: sorted by pos
    bool withMaxKMerCounts = false;
};

static inline MatchingResult completeResults(FastqKMerMatcher& m, jlong totalReads, jlong totalKMers, jlong totalBPs,
                                             const std::vector<jlong>* uniqueCounts,
                                             const std::map<int, std::vector<jshort>>* maxCounts) {
    Database* db = m.db;
    MatchingResult res;
    res.k = m.k;
    // runMatcher tail (FastqKMerMatcher.java:199-234)
    std::map<std::string, CountsPerTaxid> taxid2Stats;  // keyed by taxid; TOTAL handled separately
    for (size_t vi = 0; vi < m.statsIndex.size(); vi++) {
        if (!m.statsIndex[vi]) continue;
        CountsPerTaxid s = *m.statsIndex[vi];
        if (uniqueCounts) s.uniqueKmers = (*uniqueCounts)[vi]; else s.uniqueKmers = -1;
        if (maxCounts) {
            auto it = maxCounts->find((int)vi);
            if (it != maxCounts->end()) { s.hasMaxKMerCounts = true; s.maxKMerCounts = it->second; }
        }
        taxid2Stats[s.taxid] = s;
    }
    CountsPerTaxid& g = res.globalStats;  // MatchingResult.java:65-72, CountsPerTaxid.java:170-181
    g.level = 0; g.nullTaxid = true; g.reads = totalReads; g.kmers = totalKMers; g.readsBPs = totalBPs;
    if (maxCounts) { g.hasMaxKMerCounts = true; g.maxKMerCounts = maxCounts->at(-1); res.withMaxKMerCounts = true; }
    // completeResults (MatchingResult.java:84-118)
    TaxTree* tree = db->taxTree.get();
    std::vector<std::string> keys0;
    for (auto& kv : taxid2Stats) keys0.push_back(kv.first);
    for (const std::string& key : keys0) {
        TaxNode* node = tree->getNodeByTaxId(key);
        if (node) for (node = node->parent; node; node = node->parent) {
            if (!taxid2Stats.count(node->taxId)) {
                CountsPerTaxid c; c.level = node->getLevel(); c.taxid = node->taxId; c.maxContigDescriptorCap = 0;
                taxid2Stats[node->taxId] = c;
            }
        }
    }
    // sortTaxidsViaTree (SmallTaxTree.java:298-321): null key (TOTAL) has no node -> first; nodes by position
    std::vector<CountsPerTaxid*> order;
    order.push_back(&g);
    std::vector<CountsPerTaxid*> withNode, withoutNode;
    for (auto& kv : taxid2Stats) (tree->getNodeByTaxId(kv.first) ? withNode : withoutNode).push_back(&kv.second);
    std::sort(withoutNode.begin(), withoutNode.end(), [](CountsPerTaxid* a, CountsPerTaxid* b) { return a->taxid < b->taxid; });
    std::sort(withNode.begin(), withNode.end(), [&](CountsPerTaxid* a, CountsPerTaxid* b) {
        return tree->getNodeByTaxId(a->taxid)->position < tree->getNodeByTaxId(b->taxid)->position; });
    // (a null-node taxid sorts before any node; TOTAL's null key compares via o1.compareTo -> only one such key in practice)
    for (auto* p : withoutNode) order.push_back(p);
    for (auto* p : withNode) order.push_back(p);
    int pos = 0;
    for (CountsPerTaxid* stats : order) {
        jlong dbKMers;
        if (stats->nullTaxid) dbKMers = db->store->entries;  // getStats().getLong(null) -> entries
        else { int vi = db->store->getIndexForValue(stats->taxid); dbKMers = vi >= 0 ? db->dbKmersPerValueIndex[(size_t)vi] : 0; }
        TaxNode* node = stats->nullTaxid ? nullptr : tree->getNodeByTaxId(stats->taxid);
        stats->completeValues(pos++, dbKMers, node);
        if (node) for (node = node->parent; node; node = node->parent) {
            auto it = taxid2Stats.find(node->taxId);
            if (it != taxid2Stats.end()) it->second.accumulateFrom(*stats);
        }
    }
    for (CountsPerTaxid* p : order) res.rows.push_back(*p);
    std::sort(res.rows.begin(), res.rows.end(), [](const CountsPerTaxid& a, const CountsPerTaxid& b) { return a.pos < b.pos; });
    return res;
}

static const char* const VALUE_TYPE_NAMES[5] = {"reads", "kmers", "reads bps", "read >=1 kmer", "reads kmers"};

// ResultReporter.java:190-279. Column order = @MDCDescription.pos (CountsPerTaxid.java).
static inline std::string printMatchResult(const MatchingResult& res) {
    std::string o;
    auto pd = [&](double v, bool allow) { if (!std::isnan(v) && !std::isinf(v) && allow) o += javaDoubleToString(v); o.push_back(';'); };
    static const char* const head1[] = {"pos", "level", "name", "rank", "taxid", "reads", "kmers from reads", "kmers", "unique kmers",
        "contigs", "average contig length", "max contig length", "reads >=1 kmer", "reads bps", "avg. read length", "db coverage",
        "exp. unique kmers", "unique kmers / exp.", "db kmers", "parent taxid", "mean error", "kmer error std. dev.",
        "mean class error", "class error std. dev.", "contig len std. dev."};
    for (const char* h : head1) { o += h; o.push_back(';'); }
    for (int t = 0; t < 5; t++) { o += "norm. "; o += VALUE_TYPE_NAMES[t]; o.push_back(';'); }
    for (int t = 0; t < 5; t++) { o += "acc. "; o += VALUE_TYPE_NAMES[t]; o.push_back(';'); o += "acc. norm. "; o += VALUE_TYPE_NAMES[t]; o.push_back(';'); }
    static const char* const head2[] = {"max contig desc.", "acc. mean error", "acc. error std. dev.", "acc. mean class error", "acc. class error std. dev."};
    for (const char* h : head2) { o += h; o.push_back(';'); }
    if (res.withMaxKMerCounts) o += "max kmer counts;";
    o.push_back('\n');
    for (const CountsPerTaxid& c : res.rows) {
        bool nz = c.pos != 0;
        o += std::to_string(c.pos); o.push_back(';');
        o += std::to_string(c.level); o.push_back(';');
        o += c.name; o.push_back(';');                                       // getName (null -> "")
        if (c.rank >= 0) o += RANK_NAMES[c.rank]; o.push_back(';');         // getRank (null -> "")
        if (!c.nullTaxid) o += c.taxid; o.push_back(';');
        o += std::to_string(c.reads); o.push_back(';');
        o += std::to_string(c.readsKmers); o.push_back(';');
        o += std::to_string(c.kmers); o.push_back(';');
        o += std::to_string(c.uniqueKmers); o.push_back(';');
        o += std::to_string(c.contigs); o.push_back(';');
        pd(((double)c.kmers) / c.contigs, nz);
        o += std::to_string(c.maxContigLen); o.push_back(';');
        o += std::to_string(c.reads1KMer); o.push_back(';');
        o += std::to_string(c.readsBPs); o.push_back(';');
        pd(((double)c.readsBPs) / (double)c.reads, true);  // pos 13 is printed on the TOTAL row too
        pd(((double)c.uniqueKmers) / (double)c.dbKMers, nz);
        double expU = (1 - std::pow(1 - 1.0 / (double)c.dbKMers, (double)c.kmers)) * (double)c.dbKMers;
        pd(expU, nz);
        pd((double)c.uniqueKmers / expU, nz);
        o += std::to_string(c.dbKMers); o.push_back(';');
        if (c.hasParentTaxId) o += c.parentTaxId; o.push_back(';');
        pd(c.errorSum / (double)c.reads, nz);
        pd(std::sqrt((c.errorSquaredSum - c.errorSum * c.errorSum / (double)c.reads) / (double)(c.reads - 1)), nz);
        pd(c.classErrorSum / (double)c.reads, nz);
        pd(std::sqrt((c.classErrorSquaredSum - c.classErrorSum * c.classErrorSum / (double)c.reads) / (double)(c.reads - 1)), nz);
        pd(std::sqrt(((double)c.contigLenSquaredSum - ((double)c.kmers * (double)c.kmers) / c.contigs) / (c.contigs - 1)), nz);
        for (int t = 0; t < 5; t++) pd(((double)c.valueFor(t)) / (double)c.dbKMers, nz);
        for (int t = 0; t < 5; t++) {
            if (c.hasExtended) o += std::to_string(c.acc[t]); o.push_back(';');
            if (c.hasExtended) o += javaDoubleToString(c.accNorm[t]); o.push_back(';');
        }
        for (char ch : c.maxContigDescriptor) { if (ch == 0) break; o.push_back(ch); } o.push_back(';');  // ByteArrayUtil.print stops at 0
        jlong accReads = c.hasExtended ? c.acc[0] : 0;
        pd(c.accErrorSum / (double)accReads, nz);
        pd(std::sqrt((c.accErrorSquaredSum - (c.accErrorSum * c.accErrorSum) / (double)accReads) / (double)(accReads - 1)), nz);
        pd(c.accClassErrorSum / (double)accReads, nz);
        pd(std::sqrt((c.accClassErrorSquaredSum - (c.accClassErrorSum * c.accClassErrorSum) / (double)accReads) / (double)(accReads - 1)), nz);
        if (res.withMaxKMerCounts) {
            if (c.hasMaxKMerCounts) for (size_t i = 0; i < c.maxKMerCounts.size(); i++) { if (i) o.push_back(';'); o += std::to_string(c.maxKMerCounts[i]); }
            o.push_back(';');
        }
        o.push_back('\n');
    }
    return o;
}

// ---------------------------------------------------------------------------------------------
// C/bloom/FastqBloomFilter.java:120-161
// ---------------------------------------------------------------------------------------------
static inline bool isAcceptRead(const KMerProbFilter& filter, int k, int minPosCount, double positiveRatio,
                                const uint8_t* read, int readSize) {
    int max = readSize - k + 1;
    int posThreshold = (minPosCount > 0) ? minPosCount : (int)(max * positiveRatio);
    int negThreshold = max - posThreshold;
    jlong kmer = -1, reverseKmer = -1;
    int counter = 0, negCounter = 0, badPos = -1;
    for (int i = 0; i < max; i++) {
        if (kmer == -1) {
            kmer = cgat::kMerToLongStraight(read, i, k, &badPos);
            if (kmer == -1) i = badPos;
            else reverseKmer = cgat::kMerToLongReverse(read, i, k, nullptr);
        } else {
            kmer = cgat::nextKMerStraight(kmer, read[i + k - 1], k);
            if (kmer == -1) i += k - 1;
            else reverseKmer = cgat::nextKMerReverse(reverseKmer, read[i + k - 1], k);
        }
        if (kmer != -1) {
            if (filter.containsLong(cgat::standardKMer(kmer, reverseKmer))) {
                counter++;
                if (counter >= posThreshold) return true;
            } else {
                negCounter++;
                if (negCounter > negThreshold) return false;
            }
        }
    }
    return false;
}

}  // namespace gso
