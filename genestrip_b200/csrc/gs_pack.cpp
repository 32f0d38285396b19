// gs_pack.cpp -- see gs_pack.hpp.  Compiled by the host compiler (not nvcc): the AVX2 body uses a function-level target
// attribute and is chosen at run time, so the library still loads on a CPU without AVX2.
#include "gs_pack.hpp"

#include <immintrin.h>
#include <sched.h>

#include <algorithm>
#include <cstdlib>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace gsp {

// C=0 G=1 A=2 T=3 (C/util/CGAT.java:66-69); -1 = not a base (lower case, N, anything else: CGAT.java:60-69)
static inline int code_of(uint8_t c) {
    switch (c) {
        case 'C': return 0;
        case 'G': return 1;
        case 'A': return 2;
        case 'T': return 3;
        default: return -1;
    }
}

static void pack_words_scalar(const uint8_t* b, uint64_t n, uint64_t w0, uint64_t w1, uint64_t* codes, uint32_t* valid) {
    for (uint64_t w = w0; w < w1; w++) {
        uint64_t c = 0;
        uint32_t v = 0;
        const uint64_t base = w * 32;
        const int lim = (int)std::min<uint64_t>(32, n > base ? n - base : 0);
        for (int i = 0; i < lim; i++) {
            const int x = code_of(b[base + i]);
            if (x >= 0) { c |= (uint64_t)x << (62 - 2 * i); v |= 1u << i; }
        }
        codes[w] = c;
        valid[w] = v;
    }
}

// 32 bases per step.  The low nibble of the four letters is distinct (A 1, C 3, G 7, T 4): one byte shuffle looks up the code,
// a second one the letter that nibble would have to be, and a byte compare against the input gives the validity mask.  The
// codes are folded 2 -> 4 -> 8 bits by two multiply-adds and gathered big-endian (first base on top) by a last shuffle.
__attribute__((target("avx2"))) static void pack_words_avx2(const uint8_t* b, uint64_t n, uint64_t w0, uint64_t w1, uint64_t* codes, uint32_t* valid) {
    const uint64_t full = std::min(w1, n / 32);  // words whose 32 bases all exist
    // entries of nibbles that belong to no letter: a byte with a different low nibble, so the compare can never succeed
    const __m256i lutChar = _mm256_setr_epi8(1, 'A', 3, 'C', 'T', 4, 7, 'G', 9, 8, 11, 10, 13, 12, 15, 14,
                                             1, 'A', 3, 'C', 'T', 4, 7, 'G', 9, 8, 11, 10, 13, 12, 15, 14);
    const __m256i lutCode = _mm256_setr_epi8(0, 2, 0, 0, 3, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0,
                                             0, 2, 0, 0, 3, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i low4 = _mm256_set1_epi8(0x0F);
    const __m256i mul1 = _mm256_set1_epi16(0x0104);      // bytes (4, 1): c0 * 4 + c1
    const __m256i mul2 = _mm256_set1_epi32(0x00010010);  // words (16, 1): t0 * 16 + t1
    const __m256i gather = _mm256_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                            12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    uint64_t w = w0;
    for (; w < full; w++) {
        const __m256i v = _mm256_loadu_si256((const __m256i*)(b + w * 32));
        const __m256i idx = _mm256_and_si256(v, low4);
        const __m256i want = _mm256_shuffle_epi8(lutChar, idx);
        const __m256i ok = _mm256_cmpeq_epi8(want, v);
        const __m256i code = _mm256_and_si256(_mm256_shuffle_epi8(lutCode, idx), ok);
        const __m256i t = _mm256_maddubs_epi16(code, mul1);
        const __m256i u = _mm256_madd_epi16(t, mul2);
        const __m256i g = _mm256_shuffle_epi8(u, gather);
        const uint32_t hi = (uint32_t)_mm256_cvtsi256_si32(g);        // bases 0..15
        const uint32_t lo = (uint32_t)_mm256_extract_epi32(g, 4);     // bases 16..31
        codes[w] = ((uint64_t)hi << 32) | lo;
        valid[w] = (uint32_t)_mm256_movemask_epi8(ok);
    }
    if (w < w1) pack_words_scalar(b, n, w, w1, codes, valid);
}

// 64 bases per step: the byte compare yields both validity words as one mask register, a down-convert gathers the code bytes.
__attribute__((target("avx512f,avx512bw"))) static void pack_words_avx512(const uint8_t* b, uint64_t n, uint64_t w0, uint64_t w1, uint64_t* codes, uint32_t* valid) {
    const uint64_t full = std::min(w1, n / 32);
    const __m512i lutChar = _mm512_broadcast_i32x4(_mm_setr_epi8(1, 'A', 3, 'C', 'T', 4, 7, 'G', 9, 8, 11, 10, 13, 12, 15, 14));
    const __m512i lutCode = _mm512_broadcast_i32x4(_mm_setr_epi8(0, 2, 0, 0, 3, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0));
    const __m512i low4 = _mm512_set1_epi8(0x0F);
    const __m512i mul1 = _mm512_set1_epi16(0x0104);
    const __m512i mul2 = _mm512_set1_epi32(0x00010010);
    const __m128i swap8 = _mm_setr_epi8(7, 6, 5, 4, 3, 2, 1, 0, 15, 14, 13, 12, 11, 10, 9, 8);
    uint64_t w = w0;
    // One core streams ~10 GB/s through its L1 fill buffers; a software prefetch into the L2 a few KB ahead goes through the
    // L2's deeper queue instead: 12.2 GB/s per core, 100.7 instead of 85-91 GB/s on 16 (GS_PACK_PREFETCH = distance in bytes,
    // default 4096, 0 = off; profiles/microbench/pack_bw.txt).
    static const long pfDist = [] { const char* e = getenv("GS_PACK_PREFETCH"); return e ? atol(e) : 4096L; }();
    if (pfDist > 0) {
        for (; w + 2 <= full; w += 2) {
            _mm_prefetch((const char*)(b + w * 32 + pfDist), _MM_HINT_T1);
            const __m512i v = _mm512_loadu_si512((const void*)(b + w * 32));
            const __m512i idx = _mm512_and_si512(v, low4);
            const __mmask64 ok = _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(lutChar, idx), v);
            const __m512i code = _mm512_maskz_shuffle_epi8(ok, lutCode, idx);
            const __m512i u = _mm512_madd_epi16(_mm512_maddubs_epi16(code, mul1), mul2);
            const __m128i g = _mm_shuffle_epi8(_mm512_cvtepi32_epi8(u), swap8);
            _mm_storeu_si128((__m128i*)(codes + w), g);
            const uint64_t m = (uint64_t)ok;
            valid[w] = (uint32_t)m;
            valid[w + 1] = (uint32_t)(m >> 32);
        }
    }
    for (; w + 2 <= full; w += 2) {
        const __m512i v = _mm512_loadu_si512((const void*)(b + w * 32));
        const __m512i idx = _mm512_and_si512(v, low4);
        const __mmask64 ok = _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(lutChar, idx), v);
        const __m512i code = _mm512_maskz_shuffle_epi8(ok, lutCode, idx);
        const __m512i u = _mm512_madd_epi16(_mm512_maddubs_epi16(code, mul1), mul2);
        const __m128i g = _mm_shuffle_epi8(_mm512_cvtepi32_epi8(u), swap8);   // byte j = bases 4j .. 4j+3 -> two big-endian words
        _mm_storeu_si128((__m128i*)(codes + w), g);
        const uint64_t m = (uint64_t)ok;
        valid[w] = (uint32_t)m;
        valid[w + 1] = (uint32_t)(m >> 32);
    }
    if (w < w1) pack_words_avx2(b, n, w, w1, codes, valid);
}

static bool have_avx512() {
    static const bool v = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
    return v;
}
static bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}
const char* pack_isa() { return have_avx512() ? "avx512" : have_avx2() ? "avx2" : "scalar"; }

static inline void pack_words(const uint8_t* b, uint64_t n, uint64_t w0, uint64_t w1, uint64_t* codes, uint32_t* valid) {
    if (have_avx512()) pack_words_avx512(b, n, w0, w1, codes, valid);
    else if (have_avx2()) pack_words_avx2(b, n, w0, w1, codes, valid);
    else pack_words_scalar(b, n, w0, w1, codes, valid);
}

void pack_range(const uint8_t* bases, uint64_t n, uint64_t* codes, uint32_t* valid) { pack_words(bases, n, 0, (n + 31) / 32, codes, valid); }

// ---- pool
struct Packer::Impl {
    static constexpr uint64_t PIECE_WORDS = 1u << 13;  // 256 KiB of bases per claim
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cvStart, cvDone;
    uint64_t generation = 0;
    int running = 0;
    bool quit = false;
    // the current job
    const uint8_t* bases = nullptr;
    uint64_t n = 0, words = 0;
    uint64_t* codes = nullptr;
    uint32_t* valid = nullptr;
    std::atomic<uint64_t> next{0};

    void work() {
        for (;;) {
            const uint64_t w0 = next.fetch_add(PIECE_WORDS, std::memory_order_relaxed);
            if (w0 >= words) break;
            pack_words(bases, n, w0, std::min(words, w0 + PIECE_WORDS), codes, valid);
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m);
                cvStart.wait(lk, [&] { return quit || generation != seen; });
                if (quit) return;
                seen = generation;
            }
            work();
            {
                std::lock_guard<std::mutex> lk(m);
                if (--running == 0) cvDone.notify_all();
            }
        }
    }
};

Packer::Packer(int threads) : p(new Impl()) {
    if (threads <= 0) {
        cpu_set_t set;
        CPU_ZERO(&set);
        threads = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
        if (threads >= 8) threads -= 2;   // every pack() is a barrier over the pool: leave room for the caller's and the driver's other threads
        threads = std::max(1, std::min(threads, 32));
    }
    for (int i = 1; i < threads; i++) p->workers.emplace_back([this] { p->loop(); });
}
Packer::~Packer() {
    {
        std::lock_guard<std::mutex> lk(p->m);
        p->quit = true;
    }
    p->cvStart.notify_all();
    for (std::thread& t : p->workers) t.join();
    delete p;
}
int Packer::threads() const { return (int)p->workers.size() + 1; }

void Packer::pack(const uint8_t* bases, uint64_t n, uint64_t* codes, uint32_t* valid) {
    Impl& I = *p;
    I.bases = bases; I.n = n; I.words = (n + 31) / 32; I.codes = codes; I.valid = valid;
    I.next.store(0, std::memory_order_relaxed);
    const bool fanOut = !I.workers.empty() && I.words > Impl::PIECE_WORDS;
    if (fanOut) {
        {
            std::lock_guard<std::mutex> lk(I.m);
            I.running = (int)I.workers.size();
            I.generation++;
        }
        I.cvStart.notify_all();
    }
    I.work();
    if (fanOut) {
        std::unique_lock<std::mutex> lk(I.m);
        I.cvDone.wait(lk, [&] { return I.running == 0; });
    }
}

}  // namespace gsp
