// gs_host.cpp -- see gs_host.hpp.  Host-side mirror of the reference's goal drivers; the per-read work is done by the GPU
// through the C ABI (gs_match_* / gs_filter_*).  No CPU implementation of the hot path lives here.
#include "gs_host.hpp"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <stdexcept>
#include <thread>

#include <fcntl.h>
#include <unistd.h>

namespace gsh {

static void fail(const std::string& what) { throw std::runtime_error(what); }
static void check(int rc, const char* what) {
    if (rc != GS_OK) fail(std::string(what) + ": " + gs_last_error());
}

void DbMeta::resize(int n) {
    nValues = n;
    taxid.assign(n, ""); name.assign(n, ""); rank.assign(n, -1); parent.assign(n, -1); position.assign(n, -1); level.assign(n, 0);
    hasNode.assign(n, 0); dbKmers.assign(n, 0);
}

// C/tax/Rank.java:39-122 (Rank.toString() == the NCBI rank name)
static const char* const kRankNames[] = {
    "cellular root", "acellular root", "superkingdom", "domain", "realm", "kingdom", "phylum", "subphylum", "superclass", "class",
    "subclass", "superorder", "order", "suborder", "superfamily", "family", "subfamily", "tribe", "genus", "subgenus", "species group",
    "species", "varietas", "subspecies", "serogroup", "biotype", "strain", "serotype", "genotype", "forma", "forma specialis",
    "isolate", "clade", "no rank", "subkingdom", "section", "REFINED", "DATA", "FILE", "ID"};
const char* rankName(int ordinal) {
    const int n = (int)(sizeof(kRankNames) / sizeof(kRankNames[0]));
    return (ordinal >= 0 && ordinal < n) ? kRankNames[ordinal] : "";
}

// java.lang.Double.toString (JDK >= 19 specification: the shortest decimal that rounds to the double; computerized
// scientific notation outside [1e-3, 1e7)).
std::string javaDoubleToString(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "Infinity" : "-Infinity";
    if (v == 0) return std::signbit(v) ? "-0.0" : "0.0";
    char buf[48];
    auto res = std::to_chars(buf, buf + sizeof(buf), std::fabs(v), std::chars_format::scientific);
    std::string sci(buf, res.ptr);
    const size_t ePos = sci.find('e');
    const int exp10 = std::atoi(sci.c_str() + ePos + 1);
    std::string digits;
    for (size_t i = 0; i < ePos; i++)
        if (sci[i] != '.') digits.push_back(sci[i]);
    std::string out = std::signbit(v) ? "-" : "";
    if (exp10 >= -3 && exp10 < 7) {
        if (exp10 >= 0) {
            const size_t intLen = (size_t)exp10 + 1;
            for (size_t i = 0; i < intLen; i++) out.push_back(i < digits.size() ? digits[i] : '0');
            out.push_back('.');
            if (digits.size() > intLen) out.append(digits, intLen, std::string::npos); else out.push_back('0');
        } else {
            out += "0.";
            out.append((size_t)(-exp10 - 1), '0');
            out += digits;
        }
    } else {
        out.push_back(digits[0]);
        out.push_back('.');
        if (digits.size() > 1) out.append(digits, 1, std::string::npos); else out.push_back('0');
        out += "E" + std::to_string(exp10);
    }
    return out;
}

// ---------------------------------------------------------------------------------------------------------
// output sinks
// ---------------------------------------------------------------------------------------------------------
static bool endsWith(const std::string& s, const char* suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}
bool OutputSink::open() {
    if (mem || path.empty()) return true;
    if (endsWith(path, ".gz") || endsWith(path, ".gzip")) { gz = gzopen(path.c_str(), "wb"); return gz != nullptr; }
    fp = fopen(path.c_str(), "wb");
    return fp != nullptr;
}
void OutputSink::write(const char* p, size_t n) {
    if (mem) mem->append(p, n);
    else if (gz) gzwrite((gzFile)gz, p, (unsigned)n);
    else if (fp) fwrite(p, 1, n, fp);
}
void OutputSink::close() {
    if (gz) { gzclose((gzFile)gz); gz = nullptr; }
    if (fp) { fclose(fp); fp = nullptr; }
}

// ---------------------------------------------------------------------------------------------------------
// BufferedLineReader (B/io/BufferedLineReader.java:114-182): lines end at '\n' only, NUL bytes are dropped.  next()
// returns the reference's nextLine() count (bytes incl. the '\n', 0 at end of input); callers use count - 1 as the
// content length exactly like the reference does -- a last line without '\n' therefore loses its final byte.
// ---------------------------------------------------------------------------------------------------------
class LineReader {
   public:
    explicit LineReader(const Input& in) {
        if (in.path.empty()) { cur_ = in.data; end_ = in.data + in.len; memory_ = true; }
        else {
            gz_ = gzopen(in.path.c_str(), "rb");  // transparently reads plain files too
            if (!gz_) fail("cannot open " + in.path);
            gzbuffer(gz_, 1 << 20);
            buf_.resize(8 << 20);
            cur_ = end_ = buf_.data();
        }
    }
    // continue an open stream whose first bytes were already consumed into memory (GPU feeder falling back to this parser)
    LineReader(gzFile gz, const uint8_t* prefix, size_t prefixLen) {
        gz_ = gz; ownGz_ = false;
        buf_.resize(std::max<size_t>(8 << 20, prefixLen * 2));
        if (prefixLen) memcpy(buf_.data(), prefix, prefixLen);
        cur_ = buf_.data(); end_ = cur_ + prefixLen;
        if (!gz_) { memory_ = true; }
    }
    ~LineReader() { if (gz_ && ownGz_) gzclose(gz_); }
    // line = content without the '\n'; returns the count as defined above
    size_t next(const uint8_t*& line, size_t& len) {
        for (;;) {
            const uint8_t* nl = (const uint8_t*)memchr(cur_, '\n', (size_t)(end_ - cur_));
            if (nl) { line = cur_; len = (size_t)(nl - cur_); cur_ = nl + 1; return dropNul(line, len) + 1; }
            if (memory_ || eof_) {  // last line without '\n'
                line = cur_; len = (size_t)(end_ - cur_); cur_ = end_;
                return dropNul(line, len);
            }
            refill();
        }
    }

   private:
    size_t dropNul(const uint8_t*& line, size_t& len) {
        if (len && memchr(line, 0, len)) {
            scratch_.clear();
            for (size_t i = 0; i < len; i++) if (line[i]) scratch_.push_back(line[i]);
            line = scratch_.data(); len = scratch_.size();
        }
        return len;
    }
    void refill() {
        const size_t keep = (size_t)(end_ - cur_);
        if (keep == buf_.size()) buf_.resize(buf_.size() * 2);  // one line longer than the buffer
        uint8_t* base = buf_.data();
        if (keep) memmove(base, cur_, keep);
        const int got = gzread(gz_, base + keep, (unsigned)std::min<size_t>(buf_.size() - keep, 1u << 30));
        if (got < 0) fail("read error");
        if (got == 0) eof_ = true;
        cur_ = base; end_ = base + keep + (size_t)got;
    }
    const uint8_t *cur_ = nullptr, *end_ = nullptr;
    bool memory_ = false, eof_ = false, ownGz_ = true;
    gzFile gz_ = nullptr;
    std::vector<uint8_t> buf_, scratch_;
};

// One parsed read, as AbstractFastqReader hands it to nextEntry.
struct Record {
    std::vector<uint8_t> descriptor, read, probs;
    bool hasProbs = false;   // readProbsSize >= 0 (FASTQ with withProbs)
    int entry = 0;           // which of the two pooled ReadEntry objects carried it (threads = 0)
};

// AbstractFastqReader.doReadFastq / doReadFasta (C/fastq/AbstractFastqReader.java:288-438) for threads = 0.
class FastqReader {
   public:
    FastqReader(int k, bool withProbs) : k_(k), withProbs_(withProbs) {}
    int64_t reads = 0, kMers = 0, readBPs = 0;
    template <typename F>
    void readFastq(const Input& in, F&& nextEntry) {
        reads = kMers = readBPs = 0;
        LineReader lr(in);
        if (in.fasta) doReadFasta(lr, nextEntry); else doReadFastq(lr, nextEntry);
    }
    template <typename F>
    void readFastqFrom(LineReader& lr, F&& nextEntry) {
        reads = kMers = readBPs = 0;
        doReadFastq(lr, nextEntry);
    }

   private:
    template <typename F>
    void doReadFastq(LineReader& lr, F&& nextEntry) {
        Record rec;
        rec.entry = 0;  // nextFreeReadStruct() always returns pool[0] when threads = 0 (:447-455)
        const uint8_t* l; size_t n;
        for (;;) {
            size_t c = lr.next(l, n);
            if (c == 0) break;                                   // readDescriptorSize = -1
            rec.descriptor.assign(l, l + (c - 1));
            c = lr.next(l, n);
            if (c == 0) break;                                   // truncated record
            rec.read.assign(l, l + (c - 1));
            bool truncated = false;
            for (;;) {                                           // lines up to the one that starts with '+' (:299-307)
                c = lr.next(l, n);
                if (c == 0) { truncated = true; break; }
                if (n > 0 && l[0] == '+') break;
                rec.read.insert(rec.read.end(), l, l + (c - 1));
            }
            if (truncated) break;
            const int64_t readSize = (int64_t)rec.read.size();
            rec.probs.clear();
            rec.hasProbs = withProbs_;
            int64_t probsSize;
            c = lr.next(l, n);                                   // quality: lines until readSize characters (:318-341)
            probsSize = (int64_t)c - 1;
            if (withProbs_ && c) rec.probs.assign(l, l + (c - 1));
            while (probsSize < readSize) {
                const int64_t old = probsSize;
                c = lr.next(l, n);
                probsSize = probsSize + (int64_t)c - 1;
                if (probsSize == old - 1) break;                 // end of input
                if (withProbs_) rec.probs.insert(rec.probs.end(), l, l + (c - 1));
            }
            if (withProbs_ && probsSize < (int64_t)rec.probs.size()) rec.probs.resize((size_t)std::max<int64_t>(probsSize, 0));
            emit(rec, nextEntry);
        }
    }
    template <typename F>
    void doReadFasta(LineReader& lr, F&& nextEntry) {
        Record rec;
        int entry = 0;
        const uint8_t* l; size_t n;
        size_t c = lr.next(l, n);
        bool have = c != 0;
        if (have) rec.descriptor.assign(l, l + (c - 1));
        while (have) {
            if (!rec.descriptor.empty()) rec.descriptor[0] = '@';  // :380
            rec.read.clear();
            rec.hasProbs = false;
            rec.entry = entry;
            have = false;
            for (;;) {
                c = lr.next(l, n);
                if (c == 0) break;                                 // end of input
                if (n > 0 && l[0] == '>') { have = true; break; }  // next record's descriptor
                rec.read.insert(rec.read.end(), l, l + (c - 1));
            }
            std::vector<uint8_t> nextDesc;
            if (have) nextDesc.assign(l, l + (c - 1));
            emit(rec, nextEntry);
            rec.descriptor.swap(nextDesc);
            entry ^= 1;                                            // the two pooled entries alternate (:395-437)
        }
    }
    template <typename F>
    void emit(Record& rec, F&& nextEntry) {
        const int64_t L = (int64_t)rec.read.size();
        reads++;
        if (L >= k_) kMers += L - k_ + 1;                          // :346-349
        readBPs += L;
        nextEntry(rec, reads - 1);
    }
    int k_;
    bool withProbs_;
};

// ReadEntry.write (C/fastq/AbstractFastqReader.java:570-584)
static void writeRead(OutputSink& out, const uint8_t* desc, size_t descLen, const uint8_t* read, size_t readLen, const uint8_t* probs,
                      size_t probsLen, bool hasProbs, std::string& scratch) {
    scratch.clear();
    scratch.append((const char*)desc, descLen);
    scratch.push_back('\n');
    scratch.append((const char*)read, readLen);
    scratch += "\n+\n";
    if (hasProbs) scratch.append((const char*)probs, probsLen); else scratch.append(readLen, '~');
    scratch.push_back('\n');
    out.write(scratch.data(), scratch.size());
}

// ---------------------------------------------------------------------------------------------------------
// pinned, double-buffered batches
// ---------------------------------------------------------------------------------------------------------
struct HostBatch {
    uint8_t* bases = nullptr; size_t basesCap = 0;       // pinned
    uint64_t* offsets = nullptr; size_t offsetsCap = 0;  // pinned, [n+1]
    uint32_t n = 0;
    uint64_t firstOrdinal = 0, totalKmers = 0;
    std::string meta;                                    // descriptors (and qualities) back to back
    std::vector<uint64_t> descOff, probsOff;             // [n+1] / [n+1] offsets into meta
    std::vector<uint8_t> hasProbs, entry;
    gs_ticket ticket = 0;
    // raw FASTQ text batches (gs_match_submit_fastq)
    bool isText = false;
    uint8_t* text = nullptr; size_t textCap = 0, textLen = 0;  // pinned
    void ensureText(size_t bytes) {
        if (bytes <= textCap) return;
        uint8_t* nt = (uint8_t*)gs_alloc_pinned(bytes);
        if (!nt) fail(std::string("pinned allocation failed: ") + gs_last_error());
        if (text) { memcpy(nt, text, textLen); gs_free_pinned(text); }
        text = nt; textCap = bytes;
    }
    ~HostBatch() { gs_free_pinned(bases); gs_free_pinned(offsets); gs_free_pinned(text); }
    void ensure(size_t bytes, size_t reads) {
        if (bytes + 64 > basesCap) {
            const size_t ncap = std::max(bytes + 64, basesCap * 2);
            uint8_t* nb = (uint8_t*)gs_alloc_pinned(ncap);
            if (!nb) fail(std::string("pinned allocation failed: ") + gs_last_error());
            if (bases) { memcpy(nb, bases, used()); gs_free_pinned(bases); }
            bases = nb; basesCap = ncap;
        }
        if (reads + 1 > offsetsCap) {
            const size_t ncap = std::max(reads + 1, offsetsCap * 2);
            uint64_t* no = (uint64_t*)gs_alloc_pinned(ncap * sizeof(uint64_t));
            if (!no) fail(std::string("pinned allocation failed: ") + gs_last_error());
            if (offsets) { memcpy(no, offsets, ((size_t)n + 1) * sizeof(uint64_t)); gs_free_pinned(offsets); }
            else no[0] = 0;
            offsets = no; offsetsCap = ncap;
        }
    }
    size_t used() const { return offsets ? (size_t)offsets[n] : 0; }
    void reset(uint64_t ordinal) {
        n = 0; firstOrdinal = ordinal; totalKmers = 0; isText = false; textLen = 0; meta.clear(); descOff.assign(1, 0); probsOff.assign(1, 0); hasProbs.clear(); entry.clear();
        if (offsets) offsets[0] = 0;
    }
    void add(const Record& r, int k, bool keepProbs) {
        ensure(used() + r.read.size(), (size_t)n + 1);
        if (!r.read.empty()) memcpy(bases + offsets[n], r.read.data(), r.read.size());
        offsets[n + 1] = offsets[n] + r.read.size();
        meta.append((const char*)r.descriptor.data(), r.descriptor.size());
        descOff.push_back(meta.size());
        const bool hp = keepProbs && r.hasProbs;
        if (hp) meta.append((const char*)r.probs.data(), r.probs.size());
        probsOff.push_back(meta.size());
        hasProbs.push_back(hp ? 1 : 0);
        entry.push_back((uint8_t)r.entry);
        if ((int64_t)r.read.size() >= k) totalKmers += r.read.size() - k + 1;
        n++;
    }
    const uint8_t* desc(uint32_t i, size_t& len) const {
        const uint64_t a = i == 0 ? 0 : probsOff[i];
        len = (size_t)(descOff[i + 1] - a);
        return (const uint8_t*)meta.data() + a;
    }
    const uint8_t* probs(uint32_t i, size_t& len) const {
        len = (size_t)(probsOff[i + 1] - descOff[i + 1]);
        return (const uint8_t*)meta.data() + descOff[i + 1];
    }
};

int64_t CountsPerTaxid::valueFor(int type) const {
    switch (type) { case 0: return reads; case 1: return kmers; case 2: return readsBPs; case 3: return reads1KMer; default: return readsKmers; }
}

// Last byte offset in text[0, len) at which a record starts such that everything before it consists of whole records: a
// line that starts with '@' whose second-next line starts with '+' (a quality line may start with '@', but then the line
// two further down is a sequence).  Looks at the last few dozen lines only; 0 = none found.
static size_t lastRecordStart(const uint8_t* text, size_t len) {
    size_t starts[64];
    int ns = 0;
    size_t p = len;
    while (ns < 64 && p > 0) {  // line starts from the back: position after each '\n'
        const void* q = memrchr(text, '\n', p - 1);
        const size_t st = q ? (size_t)((const uint8_t*)q - text) + 1 : 0;
        starts[ns++] = st;
        if (!q) break;
        p = st;
    }
    // starts[0] = start of the last (possibly incomplete) line, starts[i] ascending towards the front
    for (int i = 2; i < ns; i++)
        if (text[starts[i]] == '@' && starts[i - 2] < len && text[starts[i - 2]] == '+') return starts[i];
    return 0;
}


// Block gzip (BGZF, the `bgzip` format of htslib): an ordinary multi-member gzip file -- java.util.zip.GZIPInputStream and
// zlib's gzread inflate it member by member like any other -- whose members are independent deflate streams of at most
// 64 KB and carry their own compressed size in a 'BC' extra subfield (RFC 1952 2.3.1.1).  The block boundaries are
// therefore known without inflating anything, and the blocks a text chunk needs are inflated by several threads straight
// into the pinned chunk (the reference inflates on its single producer thread).  Every block is checked
// like gzread checks it (CRC-32 and ISIZE); a member without the subfield ends the fast path at that byte (`foreignAt`),
// where the caller goes on with zlib's sequential reader, so the text never depends on which path produced it.
static std::atomic<uint64_t> g_deviceInflateNanos{0}, g_deviceInflateCalls{0}, g_bgzfReadNanos{0};
static std::atomic<uint64_t> g_deviceInflatedBlocks{0};   // block-gzip members inflated on the device by this process (tests, benchmarks)

class BgzfReader {
public:
    // true if the file's first member is a BGZF block
    static bool probe(int fd) {
        uint8_t h[18];
        if (pread(fd, h, sizeof(h), 0) != (ssize_t)sizeof(h)) return false;
        size_t hdr = 0, total = 0;
        return parseHeader(h, sizeof(h), hdr, total) == 1;
    }
    // ctx != nullptr: the blocks are inflated on the device (gs_inflate_blocks; GS_GPU_INFLATE=0 keeps the host threads)
    explicit BgzfReader(int fd, gs_ctx* ctx = nullptr) : fd_(fd) {
        const char* e = getenv("GS_INFLATE_THREADS");
        const long v = e ? atol(e) : 0;
        threads_ = (unsigned)(v > 0 ? v : std::max(1u, std::min(32u, std::thread::hardware_concurrency())));
        const char* g = getenv("GS_GPU_INFLATE");
        if (ctx && !(g && g[0] == '0')) ctx_ = ctx;
    }
    ~BgzfReader() { if (cpin_) gs_free_pinned(cpin_); }
    BgzfReader(const BgzfReader&) = delete;
    BgzfReader& operator=(const BgzfReader&) = delete;
    bool eof() const { return eof_ && carryPos_ == carry_.size(); }
    bool foreign() const { return foreign_ && carryPos_ == carry_.size(); }   // a non-BGZF member follows at foreignAt()
    size_t foreignAt() const { return cpos_; }   // compressed offset of the first block that has not been inflated yet
    void drainCarry(std::vector<uint8_t>& out) { out.insert(out.end(), carry_.begin() + (long)carryPos_, carry_.end()); carryPos_ = carry_.size(); }
    // up to `want` bytes of inflated text into dst; fewer only at the end of the file or in front of a foreign member
    size_t read(uint8_t* dst, size_t want) {
        struct Timer { std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
                       ~Timer() { g_bgzfReadNanos += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); } } timer;
        size_t done = 0;
        while (done < want) {
            if (carryPos_ < carry_.size()) {
                const size_t take = std::min(want - done, carry_.size() - carryPos_);
                memcpy(dst + done, carry_.data() + carryPos_, take);
                carryPos_ += take; done += take;
                continue;
            }
            if (eof_ || foreign_) break;
            done += round(dst + done, want - done);
        }
        return done;
    }

private:
    struct Block { size_t off, hdr, total; uint32_t isize; size_t out; };
    // 1 = BGZF block header (hdr = header bytes, total = member bytes), 0 = some other gzip member / not gzip, -1 = need more bytes
    static int parseHeader(const uint8_t* h, size_t n, size_t& hdr, size_t& total) {
        if (n < 12) return -1;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || h[3] != 4) return 0;   // BGZF: deflate, FLG = FEXTRA only
        const size_t xlen = (size_t)h[10] | ((size_t)h[11] << 8);
        if (n < 12 + xlen) return -1;
        for (size_t p = 12; p + 4 <= 12 + xlen;) {
            const size_t slen = (size_t)h[p + 2] | ((size_t)h[p + 3] << 8);
            if (h[p] == 'B' && h[p + 1] == 'C' && slen == 2 && p + 6 <= 12 + xlen) {
                hdr = 12 + xlen;
                total = ((size_t)h[p + 4] | ((size_t)h[p + 5] << 8)) + 1;
                return total >= hdr + 8 ? 1 : 0;
            }
            p += 4 + slen;
        }
        return 0;
    }
    // one window of compressed bytes: list its whole blocks, inflate those that fit into dst in parallel (the first one
    // that does not fit goes through the carry buffer)
    size_t round(uint8_t* dst, size_t room) {
        // compressed bytes this round will need, from the ratio the previous rounds saw (a short window only costs another round)
        const size_t window = std::max<size_t>((size_t)1 << 16, std::min<size_t>((size_t)((double)room / ratio_ * 1.1) + ((size_t)1 << 17), (size_t)64 << 20));
        uint8_t* cbuf = compBuffer(window);
        size_t got = 0;
        while (got < window) {
            const ssize_t r = pread(fd_, cbuf + got, window - got, (off_t)(cpos_ + got));
            if (r < 0) fail("read error");
            if (r == 0) break;
            got += (size_t)r;
        }
        if (got == 0) { eof_ = true; return 0; }
        std::vector<Block> blocks;
        size_t off = 0, out = 0;
        bool toCarry = false;
        while (off < got) {
            size_t hdr = 0, total = 0;
            const uint8_t* h = cbuf + off;
            if ((got - off >= 1 && h[0] != 0x1f) || (got - off >= 2 && h[1] != 0x8b)) {
                // bytes behind the last member that are no gzip header: ignored, as by zlib's gzread and GZIPInputStream
                if (blocks.empty()) { eof_ = true; return 0; }
                break;
            }
            const int rc = parseHeader(h, got - off, hdr, total);
            if (rc == 0) { if (blocks.empty()) { foreign_ = true; return 0; } break; }
            if (rc < 0 || off + total > got) {
                if (blocks.empty() && got < window) fail("read error");   // the file ends inside a block (gzread: unexpected end of file)
                break;
            }
            const uint8_t* t = cbuf + off + total - 4;
            const uint32_t isize = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
            if (isize > (1u << 16)) { if (blocks.empty()) { foreign_ = true; return 0; } break; }   // not a BGZF block after all
            if (out + isize > room) {
                if (!blocks.empty()) break;
                toCarry = true;                                           // the only block of this round: through the carry buffer
            }
            blocks.push_back(Block{off, hdr, total, isize, out});
            off += total; out += isize;
            if (toCarry) break;
        }
        if (blocks.empty()) fail("read error");
        uint8_t* target = dst;
        if (toCarry) { carry_.resize(blocks[0].isize); carryPos_ = 0; target = carry_.data(); }
        if (ctx_ && !toCarry && blocks.size() >= 32) {
            // on the device: one thread per block, size and CRC-32 checked there (gs_inflate.cu)
            std::vector<gs_deflate_block> tab(blocks.size());
            for (size_t i = 0; i < blocks.size(); i++) {
                const Block& b = blocks[i];
                const uint8_t* c = cbuf + b.off + b.total - 8;
                tab[i] = gs_deflate_block{b.off + b.hdr, b.out, (uint32_t)(b.total - b.hdr - 8), b.isize,
                                          (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16) | ((uint32_t)c[3] << 24), 0u};
            }
            const auto t0 = std::chrono::steady_clock::now();
            const int rc = gs_inflate_blocks(ctx_, cbuf, off, tab.data(), (uint32_t)tab.size(), target, out);
            g_deviceInflateNanos += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
            g_deviceInflateCalls += 1;
            if (rc == GS_ERR_DATA) fail("read error");   // corrupt block (gzread: data error / incorrect data check)
            if (rc != GS_OK) fail(std::string("gs_inflate_blocks: ") + gs_last_error());
            g_deviceInflatedBlocks += blocks.size();
            cpos_ += off;
            if (off > 0 && out > 0) ratio_ = std::max(1.0, (double)out / (double)off);
            return out;
        }
        const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(threads_, blocks.size() / 8 + 1));
        std::vector<int> bad(nt, 0);
        auto work = [&](unsigned t) {
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            if (inflateInit2(&zs, -15) != Z_OK) { bad[t] = 1; return; }
            for (size_t i = blocks.size() * t / nt; i < blocks.size() * (t + 1) / nt; i++) {
                const Block& b = blocks[i];
                const uint8_t* src = cbuf + b.off;
                zs.next_in = (Bytef*)(src + b.hdr); zs.avail_in = (uInt)(b.total - b.hdr - 8);
                zs.next_out = (Bytef*)(target + b.out); zs.avail_out = (uInt)b.isize;
                uint8_t none;
                if (b.isize == 0) { zs.next_out = &none; zs.avail_out = 0; }
                const int rc = inflate(&zs, Z_FINISH);
                const uint8_t* c = src + b.total - 8;
                const uint32_t crc = (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16) | ((uint32_t)c[3] << 24);
                if (rc != Z_STREAM_END || zs.avail_out != 0 || (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef*)(target + b.out), (uInt)b.isize) != crc) { bad[t] = 1; break; }
                if (inflateReset(&zs) != Z_OK) { bad[t] = 1; break; }
            }
            inflateEnd(&zs);
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; t++) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
        for (unsigned t = 0; t < nt; t++)
            if (bad[t]) fail("read error");   // corrupt block (gzread: data error / incorrect data check)
        cpos_ += off;
        if (off > 0 && out > 0) ratio_ = std::max(1.0, (double)out / (double)off);
        return toCarry ? 0 : out;
    }
    // the window of compressed bytes: pinned when it goes to the device
    uint8_t* compBuffer(size_t bytes) {
        if (!ctx_) { cbuf_.resize(bytes); return cbuf_.data(); }
        if (bytes > cpinCap_) {
            if (cpin_) gs_free_pinned(cpin_);
            cpinCap_ = bytes + bytes / 4;
            cpin_ = (uint8_t*)gs_alloc_pinned(cpinCap_);
            if (!cpin_) { cpinCap_ = 0; fail(std::string("pinned allocation failed: ") + gs_last_error()); }
        }
        return cpin_;
    }
    gs_ctx* ctx_ = nullptr;
    uint8_t* cpin_ = nullptr; size_t cpinCap_ = 0;
    int fd_;
    unsigned threads_ = 1;
    size_t cpos_ = 0;
    bool eof_ = false, foreign_ = false;
    std::vector<uint8_t> cbuf_, carry_;
    size_t carryPos_ = 0;
    double ratio_ = 3.0;   // inflated / compressed bytes of the last round
};

// The GPU feeder's host side, shared by the match and filter drivers: stream one FASTQ input as pinned text chunks that end
// at a record boundary.  cur() = the batch to fill, next() = the batch cur() will yield after the next successful submit (the
// read-ahead target); submit(batch, cut) hands text[0, cut) to the device and returns true if the device took it (the driver
// then parks the batch and cur() yields the next one), false if it refused (not strict 4-line
// FASTQ); from the first refusal on, the rest of the input -- the refused chunk included -- goes through the sequential
// parser (`sequential(LineReader&)`), so the results never depend on this fast path.
// GS_HOST_TIMING=1: where the goal driver's wall time goes (seconds per phase on stderr at the end of runMatcher)
struct HostTimers {
    double submit = 0, collectWait = 0, process = 0, readerJoin = 0, open = 0, finish = 0;
    bool on = getenv("GS_HOST_TIMING") != nullptr;
    static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
};
static HostTimers g_timers;
template <typename CurFn, typename NextFn, typename SubmitFn, typename SeqFn>
static void feedFastqText(gs_ctx* ctx, const Input& in, size_t chunkBytes, CurFn&& cur, NextFn&& next, SubmitFn&& submit, SeqFn&& sequential) {
    gzFile gz = nullptr;
    int fd = -1;        // uncompressed files are read with parallel pread() straight into the pinned chunk
    int zfd = -1;       // block-gzip files: the descriptor the BgzfReader reads from
    std::unique_ptr<BgzfReader> bgzf;
    size_t filePos = 0;
    if (!in.path.empty()) {
        fd = open(in.path.c_str(), O_RDONLY);
        if (fd < 0) fail("cannot open " + in.path);
        unsigned char magic[2] = {0, 0};
        const ssize_t m = pread(fd, magic, 2, 0);
        if (m == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
            if (BgzfReader::probe(fd) && !getenv("GS_NO_BGZF")) {   // block gzip: the blocks are inflated in parallel
                bgzf.reset(new BgzfReader(fd, ctx));
                zfd = fd; fd = -1;
            } else {                                                // gzip: zlib inflates on this thread
                close(fd); fd = -1;
                gz = gzopen(in.path.c_str(), "rb");
                if (!gz) fail("cannot open " + in.path);
                gzbuffer(gz, 1 << 20);
            }
        }
    }
    struct Closer { gzFile& g; int& f; int& z; ~Closer() { if (g) gzclose(g); if (f >= 0) close(f); if (z >= 0) close(z); } } closer{gz, fd, zfd};
    auto readPlain = [&](uint8_t* dst, size_t want) -> size_t {
        const size_t slice = (size_t)4 << 20;
        static const size_t maxThreads = [] { const char* e = getenv("GS_FEEDER_THREADS"); const long v = e ? atol(e) : 0; return (size_t)(v > 0 ? v : 8); }();
        const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>({maxThreads, (size_t)std::max(1u, std::thread::hardware_concurrency()), (want + slice - 1) / slice}));
        std::vector<size_t> got(nt, 0);
        std::vector<int> err(nt, 0);
        auto work = [&](unsigned t) {
            const size_t a = want * t / nt, b = want * (t + 1) / nt;
            size_t done = 0;
            while (a + done < b) {
                const ssize_t r = pread(fd, dst + a + done, b - a - done, (off_t)(filePos + a + done));
                if (r < 0) { err[t] = 1; break; }
                if (r == 0) break;
                done += (size_t)r;
            }
            got[t] = done;
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; t++) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
        size_t total = 0;
        for (unsigned t = 0; t < nt; t++) {
            if (err[t]) fail("read error");
            total += got[t];
            if (got[t] < want * (t + 1) / nt - want * t / nt) break;   // end of file inside this slice
        }
        filePos += total;
        return total;
    };
    const size_t chunk = std::max<size_t>(chunkBytes, 1 << 12);
    size_t memPos = 0;
    bool eof = false;
    // bytes [len, target) of a chunk buffer from the stream / file / caller's memory; sets eof at the end of the input
    auto fill = [&](uint8_t* text, size_t& len, size_t target) {
        while (!eof && len < target) {
            if (bgzf) {
                len += bgzf->read(text + len, target - len);
                if (bgzf->eof()) eof = true;
                else if (bgzf->foreign()) {   // an ordinary gzip member follows: zlib's sequential reader takes over from there
                    if (lseek(zfd, (off_t)bgzf->foreignAt(), SEEK_SET) < 0) fail("seek error");
                    gz = gzdopen(zfd, "rb");
                    if (!gz) fail("cannot open " + in.path);
                    zfd = -1;                 // owned by gz now
                    gzbuffer(gz, 1 << 20);
                    bgzf.reset();
                }
            } else if (gz) {
                const int got = gzread(gz, text + len, (unsigned)std::min<size_t>(target - len, 1u << 30));
                if (got < 0) fail("read error");
                if (got == 0) eof = true;
                len += (size_t)got;
            } else if (fd >= 0) {
                const size_t want = target - len;
                const size_t got = readPlain(text + len, want);
                if (got < want) eof = true;
                len += got;
            } else {
                const size_t take = std::min(target - len, in.len - memPos);
                if (take) memcpy(text + len, in.data + memPos, take);
                memPos += take; len += take;
                if (memPos == in.len) eof = true;
            }
        }
    };
    HostBatch* b = cur();
    b->isText = true;
    b->textLen = 0;
    b->ensureText(chunk + 64);
    size_t len = 0;
    fill(b->text, len, chunk);
    for (;;) {
        // ---- cut at the last record boundary (widen the chunk if a single record is longer than it)
        size_t cut = 0, target = std::max(chunk, len);
        bool refused = false;
        for (;;) {
            cut = eof ? len : lastRecordStart(b->text, len);
            if (eof || cut > 0) break;
            if (target >= ((size_t)1 << 31)) { refused = true; break; }  // no record boundary in 2 GiB: not 4-line FASTQ
            target *= 2;
            b->textLen = len;
            b->ensureText(target + 64);
            fill(b->text, len, target);
        }
        if (!refused && cut == 0) return;   // end of the input, nothing left
        // ---- read ahead: the next chunk (tail of this one first) is read by a helper thread into the batch that cur() will
        // yield next, while this thread hands the current chunk to the device and collects the oldest batch
        HostBatch* nb = nullptr;
        size_t nlen = 0;
        std::thread reader;
        std::string readErr;
        const bool more = !refused && (!eof || cut < len);
        if (more) {
            nb = next();
            nb->isText = true;
            nb->textLen = 0;
            nlen = len - cut;
            nb->ensureText(std::max(chunk, nlen + 1) + 64);
            if (nlen) memcpy(nb->text, b->text + cut, nlen);
            if (!eof) {
                const size_t ntarget = std::max(chunk, nlen + 1);
                reader = std::thread([&, ntarget] { try { fill(nb->text, nlen, ntarget); } catch (const std::exception& e) { readErr = e.what(); } });
            }
        }
        bool taken = false;
        std::string submitErr;
        if (!refused) {
            b->textLen = cut;
            try { taken = submit(b, cut); } catch (const std::exception& e) { submitErr = e.what(); }
        }
        { const double t0 = HostTimers::now(); if (reader.joinable()) reader.join(); g_timers.readerJoin += HostTimers::now() - t0; }
        if (!submitErr.empty()) fail(submitErr);
        if (!readErr.empty()) fail(readErr);
        if (!taken) {
            // sequential parser over everything not yet consumed: this chunk, what was read ahead, then the rest of the input
            std::vector<uint8_t> pending(b->text, b->text + len);
            if (nb && nlen > len - cut) pending.insert(pending.end(), nb->text + (len - cut), nb->text + nlen);
            if (fd >= 0) {   // plain file: hand the rest of it to zlib's transparent reader, positioned behind what was consumed
                gz = gzopen(in.path.c_str(), "rb");
                if (!gz) fail("cannot open " + in.path);
                gzbuffer(gz, 1 << 20);
                if (gzseek(gz, (z_off_t)filePos, SEEK_SET) < 0) fail("seek error");
            } else if (bgzf) {   // block gzip: what the reader still holds, then zlib's sequential reader from the next block on
                bgzf->drainCarry(pending);
                if (!bgzf->eof()) {
                    if (lseek(zfd, (off_t)bgzf->foreignAt(), SEEK_SET) < 0) fail("seek error");
                    gz = gzdopen(zfd, "rb");
                    if (!gz) fail("cannot open " + in.path);
                    zfd = -1;
                    gzbuffer(gz, 1 << 20);
                }
                bgzf.reset();
            } else if (!gz && in.path.empty()) {
                pending.insert(pending.end(), in.data + memPos, in.data + in.len);   // in-memory input: the rest follows in memory
            }
            LineReader lr(gz, pending.data(), pending.size());
            sequential(lr);
            return;
        }
        if (!more) return;
        b = cur();          // the batch the read-ahead went into
        if (b != nb) fail("feeder: batch rotation out of step");
        b->isText = true;   // (the driver's reset() of the fresh batch cleared the flag; the text is already in place)
        len = nlen;
    }
}

// ---------------------------------------------------------------------------------------------------------
// FastqKMerMatcher
// ---------------------------------------------------------------------------------------------------------
struct FastqKMerMatcher::Batch : HostBatch {};

FastqKMerMatcher::FastqKMerMatcher(gs_db* db, const DbMeta& meta, const MatchConfig& cfg) : db_(db), meta_(meta), cfg_(cfg) {}
FastqKMerMatcher::~FastqKMerMatcher() {}

static void appendInt(std::string& s, long long v) { s += std::to_string(v); }

// writeMatchDetails + printKrakenStyleOut (C/match/FastqKMerMatcher.java:597-611, 723-756) for one read
void FastqKMerMatcher::writeKrakenLine(OutputSink& krakenOut, const gs_read_result& r, const uint8_t* d, size_t dl, int64_t L, const gs_run* runs,
                                       size_t nRuns, int entry, std::string& line) {
    if (nRuns > 0) entryBufferUsed_[entry] = true;  // printKrakenStyleOut allocated the pooled entry's buffer
    if (!((cfg_.writeAll || r.class_vidx >= 0) && entryBufferUsed_[entry])) return;
    line.clear();
    line += r.class_vidx >= 0 ? "C\t" : "U\t";
    size_t sp = dl;
    for (size_t j = 1; j < dl; j++) if (d[j] == ' ') { sp = j; break; }
    if (sp > 1) line.append((const char*)d + 1, sp - 1);
    line.push_back('\t');
    if (r.class_vidx >= 0) line += meta_.taxid[(size_t)r.class_vidx]; else line.push_back('0');
    line.push_back('\t');
    appendInt(line, L);
    line.push_back('\t');
    for (size_t j = 0; j < nRuns; j++) {
        if (j > 0) line.push_back(' ');
        if (runs[j].label == GS_RUN_INVALID) line.push_back('A');
        else if (runs[j].label == GS_RUN_MISS) line.push_back('0');
        else line += meta_.taxid[runs[j].label];
        line.push_back(':');
        appendInt(line, runs[j].len);
    }
    line.push_back('\n');
    krakenOut.write(line.data(), line.size());
}

// ---- the four double sums of the classified reads (FastqKMerMatcher.java:511-526) --------------------------------------------
// errorSum, errorSquaredSum, classErrorSum and classErrorSquaredSum are floating-point sums, i.e. order dependent: they are
// added on this thread in read order.  (Measured, GS_HOST_TIMING=1, 32 M reads: 0.19 s = 6 ns per read; a pool of threads that
// each own a residue class of taxa -- every taxon still sees its reads in order, so the doubles stay bit-identical -- was tried
// and was slower, 0.27 s: eight threads each stream the whole result table of the batch out of pinned memory.)
// maxOf(i) = L - k + 1 of read i of the batch
template <typename MaxFn>
static void classifiedReadSums(const gs_read_result* res, uint32_t n, std::vector<CountsPerTaxid>& stats, MaxFn&& maxOf) {
    for (uint32_t i = 0; i < n; i++) {
        const gs_read_result& r = res[i];
        if (!(r.flags & GS_READ_ACCEPTED)) continue;
        const int64_t max = maxOf(i);
        CountsPerTaxid& st = stats[(size_t)r.class_vidx];
        const double err = ((double)(int32_t)r.tax_err) / (double)max;
        const double classErr = ((double)(int32_t)(max - (int64_t)r.read_kmers)) / (double)max;
        st.errorSum += err;
        st.errorSquaredSum += err * err;
        st.classErrorSum += classErr;
        st.classErrorSquaredSum += classErr * classErr;
    }
}

void FastqKMerMatcher::processBatch(gs_sess* s, Batch& b, OutputSink* filtered, OutputSink* krakenOut, std::vector<CountsPerTaxid>& stats,
                                    std::vector<uint64_t>& bestKey) {
    const int V = meta_.nValues;
    std::vector<gs_read_result> res(b.n);
    std::vector<gs_maxcontig_event> ev((size_t)std::max(V, 1));
    std::vector<uint64_t> runOff;
    std::vector<gs_run> runs;
    uint32_t nEv = 0;
    if (krakenOut) { runOff.assign((size_t)b.n + 1, 0); runs.resize((size_t)std::max<uint64_t>(b.totalKmers, 1)); }
    check(gs_match_collect(s, b.ticket, res.data(), ev.data(), (uint32_t)ev.size(), &nEv, krakenOut ? runOff.data() : nullptr,
                           krakenOut ? runs.data() : nullptr, b.totalKmers), "gs_match_collect");
    // maxContigDescriptor (FastqKMerMatcher.java:402-409): header[1 .. first ' '), at most initialReadSize-1 bytes, of the
    // first read (lowest ordinal) that reached the maximum contig length of its taxon
    for (uint32_t e = 0; e < nEv; e++) {
        const gs_maxcontig_event& x = ev[e];
        const uint64_t key = ((uint64_t)x.contig_len << 40) | ((((uint64_t)1 << 40) - 1) - x.read_no);
        if (key <= bestKey[x.vidx]) continue;
        bestKey[x.vidx] = key;
        size_t dl;
        const uint8_t* d = b.desc((uint32_t)(x.read_no - b.firstOrdinal), dl);
        std::string& dst = stats[x.vidx].maxContigDescriptor;
        dst.clear();
        for (size_t j = 1; j < dl && j < (size_t)cfg_.initialReadSizeBytes && d[j] != ' '; j++) dst.push_back((char)d[j]);
    }
    std::string scratch, line;
    const int k = meta_.k;
    for (uint32_t i = 0; i < b.n; i++) {
        const gs_read_result& r = res[i];
        const uint64_t a = b.offsets[i], e = b.offsets[i + 1];
        const int64_t L = (int64_t)(e - a);
        size_t dl, pl;
        const uint8_t* d = b.desc(i, dl);
        // afterMatch (:304-315)
        if ((r.flags & GS_READ_FOUND) && filtered) {
            const uint8_t* p = b.probs(i, pl);
            writeRead(*filtered, d, dl, b.bases + a, (size_t)L, p, pl, b.hasProbs[i] != 0, scratch);
        }
        if (krakenOut) writeKrakenLine(*krakenOut, r, d, dl, L, runs.data() + runOff[i], (size_t)(runOff[i + 1] - runOff[i]), b.entry[i], line);
    }
    // classified-read statistics: the four double sums, per taxon in read order (:511-526)
    classifiedReadSums(res.data(), b.n, stats, [&](uint32_t i) { return (int64_t)(b.offsets[i + 1] - b.offsets[i]) - k + 1; });
}

// A batch that went to the device as raw FASTQ text: descriptors, bases and qualities are read from the (pinned) text through
// the record table the device returns; same bookkeeping as processBatch.
void FastqKMerMatcher::processTextBatch(gs_sess* s, Batch& b, OutputSink* filtered, OutputSink* krakenOut, std::vector<CountsPerTaxid>& stats,
                                        std::vector<uint64_t>& bestKey) {
    const gs_read_result* res = nullptr; const gs_maxcontig_event* ev = nullptr; const uint32_t* evHdr = nullptr; const gs_fastq_rec* recs = nullptr;
    uint32_t n = 0, nEv = 0;
    std::vector<uint64_t> runOff;
    std::vector<gs_run> runs;
    if (krakenOut) { runOff.assign((size_t)b.n + 1, 0); runs.resize((size_t)std::max<uint64_t>(b.totalKmers, 1)); }
    const double tc0 = HostTimers::now();
    check(gs_match_collect_fastq(s, b.ticket, &res, &n, &ev, &evHdr, &nEv, &recs, krakenOut ? runOff.data() : nullptr,
                                 krakenOut ? runs.data() : nullptr, b.totalKmers), "gs_match_collect_fastq");
    const double tc1 = HostTimers::now();
    g_timers.collectWait += tc1 - tc0;
    struct ProcT { double t0; ~ProcT() { g_timers.process += HostTimers::now() - t0; } } procT{tc1};
    for (uint32_t e = 0; e < nEv; e++) {  // maxContigDescriptor (FastqKMerMatcher.java:402-409)
        const gs_maxcontig_event& x = ev[e];
        const uint64_t key = ((uint64_t)x.contig_len << 40) | ((((uint64_t)1 << 40) - 1) - x.read_no);
        if (key <= bestKey[x.vidx]) continue;
        bestKey[x.vidx] = key;
        const gs_fastq_rec& rc = recs[(size_t)(x.read_no - b.firstOrdinal)];
        const uint8_t* d = b.text + rc.hdr_start;
        const size_t dl = (size_t)(rc.seq_start - 1 - rc.hdr_start);
        std::string& dst = stats[x.vidx].maxContigDescriptor;
        dst.clear();
        for (size_t j = 1; j < dl && j < (size_t)cfg_.initialReadSizeBytes && d[j] != ' '; j++) dst.push_back((char)d[j]);
    }
    std::string scratch, line;
    const int k = meta_.k;
    if (filtered || krakenOut)   // the outputs are written in read order by this thread
        for (uint32_t i = 0; i < n; i++) {
            const gs_read_result& r = res[i];
            const gs_fastq_rec& rc = recs[i];
            const int64_t L = (int64_t)rc.seq_len;
            if ((r.flags & GS_READ_FOUND) && filtered)  // afterMatch (:304-315)
                writeRead(*filtered, b.text + rc.hdr_start, (size_t)(rc.seq_start - 1 - rc.hdr_start), b.text + rc.seq_start, (size_t)L,
                          b.text + rc.qual_start, (size_t)(recs[i + 1].hdr_start - 1 - rc.qual_start), cfg_.withProbs, scratch);
            if (krakenOut)  // FASTQ records always travel in the first pooled ReadEntry at threads = 0 (AbstractFastqReader.java:447-455)
                writeKrakenLine(*krakenOut, r, b.text + rc.hdr_start, (size_t)(rc.seq_start - 1 - rc.hdr_start), L, runs.data() + runOff[i],
                                (size_t)(runOff[i + 1] - runOff[i]), 0, line);
        }
    // the four double sums, per taxon in read order (:511-526)
    classifiedReadSums(res, n, stats, [&](uint32_t i) { return (int64_t)recs[i].seq_len - k + 1; });
}

MatchingResult FastqKMerMatcher::runMatcher(const std::vector<Input>& fastqs, OutputSink* filtered, OutputSink* krakenOut) {
    const int V = meta_.nValues;
    gs_match_cfg c;
    gs_match_cfg_default(&c);
    c.classify_reads = cfg_.classifyReads; c.count_unique_kmers = cfg_.countUniqueKMers; c.max_kmer_res_counts = cfg_.maxKMerResCounts;
    c.use_bloom_filter = cfg_.useBloomFilterForMatch; c.max_classification_paths = cfg_.maxClassificationPaths;
    c.min_kmers_for_class = cfg_.minKMersForClass; c.max_read_tax_error_count = cfg_.maxReadTaxErrorCount;
    c.max_read_class_error_count = cfg_.maxReadClassErrorCount; c.want_runs = krakenOut ? 1 : 0; c.layout = cfg_.layout;
    const double tOpen0 = HostTimers::now();
    g_timers = HostTimers();
    gs_sess* s = gs_match_open(db_, &c);
    if (!s) fail(std::string("gs_match_open: ") + gs_last_error());
    g_timers.open = HostTimers::now() - tOpen0;
    struct Closer { gs_sess* s; ~Closer() { gs_match_close(s); } } closer{s};
    if (filtered && !filtered->open()) fail("cannot open filtered output");
    if (krakenOut && !krakenOut->open()) fail("cannot open kraken output");

    std::vector<CountsPerTaxid> stats((size_t)V);           // statsIndex (initStats)
    std::vector<uint64_t> bestKey((size_t)V, 0);
    const size_t nSlots = (size_t)GS_MAX_INFLIGHT * 8 + 1;  // enough host batches for every device's in-flight slots
    std::vector<std::unique_ptr<Batch>> pool;
    std::deque<Batch*> inflight, freeList;
    size_t maxInflight = 0;
    {   // in-flight limit = GS_MAX_INFLIGHT per device
        maxInflight = (size_t)GS_MAX_INFLIGHT * (size_t)std::max(1, gs_db_n_devices(db_));
        for (size_t i = 0; i < std::min(nSlots, maxInflight + 1); i++) { pool.emplace_back(new Batch()); freeList.push_back(pool.back().get()); }
    }
    uint64_t ordinal = 0;
    int64_t totalReads = 0, totalKMers = 0, totalBPs = 0;
    Batch* cur = freeList.front(); freeList.pop_front();
    cur->reset(ordinal);
    auto collectOldest = [&]() {
        Batch* b = inflight.front(); inflight.pop_front();
        if (b->isText) processTextBatch(s, *b, filtered, krakenOut, stats, bestKey);
        else processBatch(s, *b, filtered, krakenOut, stats, bestKey);
        freeList.push_back(b);
    };
    auto flush = [&]() {
        if (cur->n == 0) return;
        cur->ensure(cur->used(), cur->n);
        check(gs_match_submit(s, cur->bases, cur->offsets, cur->n, cur->firstOrdinal, &cur->ticket), "gs_match_submit");
        inflight.push_back(cur);
        if (inflight.size() >= maxInflight) collectOldest();
        cur = freeList.front(); freeList.pop_front();
        cur->reset(ordinal);
    };
    const bool keepProbs = filtered != nullptr && cfg_.withProbs;
    FastqReader reader(meta_.k, cfg_.withProbs);
    auto onRecord = [&](const Record& r, int64_t) {
        if (cur->n >= cfg_.batchReads || (cur->n > 0 && cur->used() + r.read.size() > cfg_.batchBytes)) flush();
        cur->add(r, meta_.k, keepProbs);
        ordinal++;
    };
    for (const Input& in : fastqs) {   // processFastqStreams (C/fastq/AbstractLoggingFastqStreamer.java:95-131)
        if (!cfg_.gpuParse || in.fasta) {
            reader.readFastq(in, onRecord);
            totalReads += reader.reads; totalKMers += reader.kMers; totalBPs += reader.readBPs;
            continue;
        }
        // ---- GPU feeder: text chunks cut at record boundaries; `cur` is the batch being filled
        flush();  // host-parsed reads of an earlier input keep their place in the order
        feedFastqText(gs_db_context(db_), in, cfg_.textChunkBytes,
            [&]() -> HostBatch* { return cur; },
            [&]() -> HostBatch* { return freeList.front(); },
            [&](HostBatch* b, size_t cut) -> bool {
                gs_fastq_info info;
                gs_ticket t = 0;
                { const double t0 = HostTimers::now(); check(gs_match_submit_fastq(s, b->text, cut, ordinal, &info, &t), "gs_match_submit_fastq"); g_timers.submit += HostTimers::now() - t0; }
                textChunks++;
                if (info.status) return false;
                b->ticket = t; b->n = info.n_reads; b->firstOrdinal = ordinal; b->totalKmers = info.total_kmers;
                ordinal += info.n_reads;
                totalReads += info.n_reads; totalKMers += (int64_t)info.total_kmers; totalBPs += (int64_t)info.total_bps;
                inflight.push_back(cur);
                if (inflight.size() >= maxInflight) collectOldest();
                cur = freeList.front(); freeList.pop_front();
                cur->reset(ordinal);
                return true;
            },
            [&](LineReader& lr) {
                textChunksRefused++;
                cur->reset(ordinal);
                reader.readFastqFrom(lr, onRecord);
                totalReads += reader.reads; totalKMers += reader.kMers; totalBPs += reader.readBPs;
            });
        if (cur->isText) cur->reset(ordinal);   // a chunk buffer that was filled but never submitted (empty input)
    }
    flush();
    while (!inflight.empty()) collectOldest();
    if (filtered) filtered->close();
    if (krakenOut) krakenOut->close();

    std::vector<gs_taxon_counts> counts((size_t)std::max(V, 1));
    std::vector<int16_t> top;
    const bool withCounts = cfg_.countUniqueKMers && cfg_.maxKMerResCounts > 0;
    if (withCounts) top.assign((size_t)(V + 1) * (size_t)cfg_.maxKMerResCounts, 0);
    { const double t0 = HostTimers::now(); check(gs_match_finish(s, counts.data(), withCounts ? top.data() : nullptr), "gs_match_finish"); g_timers.finish = HostTimers::now() - t0; }
    launches_ += gs_match_kernel_launches(s);
    if (g_timers.on)
        fprintf(stderr, "gs_host timing: total %.3f s since open | open %.3f submit %.3f collect-wait %.3f process %.3f reader-join %.3f finish %.3f (text chunks %llu)\n",
                HostTimers::now() - tOpen0, g_timers.open, g_timers.submit, g_timers.collectWait, g_timers.process, g_timers.readerJoin, g_timers.finish,
                (unsigned long long)textChunks);

    // runMatcher tail (:199-234)
    MatchingResult res;
    res.k = meta_.k;
    res.totalReads = totalReads; res.totalKMers = totalKMers; res.totalBPs = totalBPs;
    for (int v = 0; v < V; v++) {
        const gs_taxon_counts& g = counts[(size_t)v];
        if (!g.touched) continue;
        CountsPerTaxid st = stats[(size_t)v];
        st.vidx = v; st.level = meta_.level[(size_t)v];
        st.kmers = g.kmers; st.contigs = (int32_t)g.contigs; st.contigLenSquaredSum = g.contig_len_squared_sum;
        st.maxContigLen = g.max_contig_len; st.reads1KMer = g.reads_1kmer; st.reads = g.reads; st.readsKmers = g.reads_kmers;
        st.readsBPs = g.reads_bps;
        st.uniqueKmers = cfg_.countUniqueKMers ? g.unique_kmers : -1;
        if (withCounts) {  // countMap.get(taxid): only taxa with at least one hit position have a row
            if (g.unique_kmers > 0) {
                st.hasMaxKMerCounts = true;
                st.maxKMerCounts.assign(top.begin() + (size_t)v * cfg_.maxKMerResCounts, top.begin() + (size_t)(v + 1) * cfg_.maxKMerResCounts);
            }
        }
        res.taxid2Stats[v] = st;
    }
    CountsPerTaxid& g = res.globalStats;  // new CountsPerTaxid(0, null, totalReads, totalKMers, totalBPs, totalMaxCounts)
    g.level = 0; g.vidx = -1; g.reads = totalReads; g.kmers = totalKMers; g.readsBPs = totalBPs;
    if (withCounts) {
        res.withMaxKMerCounts = true;
        g.hasMaxKMerCounts = true;
        g.maxKMerCounts.assign(top.begin() + (size_t)V * cfg_.maxKMerResCounts, top.end());
    }
    return res;
}

// MatchingResult.completeResults (C/match/MatchingResult.java:84-118)
void MatchingResult::completeResults(const DbMeta& meta) {
    std::vector<int> keys;
    for (auto& kv : taxid2Stats) keys.push_back(kv.first);
    for (int v : keys) {  // add missing ancestors
        if (!meta.hasNode[(size_t)v]) continue;
        for (int p = meta.parent[(size_t)v]; p >= 0; p = meta.parent[(size_t)p]) {
            if (!taxid2Stats.count(p)) { CountsPerTaxid c; c.vidx = p; c.level = meta.level[(size_t)p]; taxid2Stats[p] = c; }
        }
    }
    // sortTaxidsViaTree: the null key (TOTAL) first, then by tree position; tax ids without node lexicographically before nodes
    std::vector<CountsPerTaxid*> order;
    order.push_back(&globalStats);
    std::vector<CountsPerTaxid*> withNode, withoutNode;
    for (auto& kv : taxid2Stats) (meta.hasNode[(size_t)kv.first] ? withNode : withoutNode).push_back(&kv.second);
    std::sort(withoutNode.begin(), withoutNode.end(), [&](CountsPerTaxid* a, CountsPerTaxid* b) { return meta.taxid[(size_t)a->vidx] < meta.taxid[(size_t)b->vidx]; });
    std::sort(withNode.begin(), withNode.end(), [&](CountsPerTaxid* a, CountsPerTaxid* b) { return meta.position[(size_t)a->vidx] < meta.position[(size_t)b->vidx]; });
    order.insert(order.end(), withoutNode.begin(), withoutNode.end());
    order.insert(order.end(), withNode.begin(), withNode.end());
    int pos = 0;
    for (CountsPerTaxid* st : order) {
        st->pos = pos++;
        st->dbKMers = st->vidx < 0 ? meta.totalKmers : meta.dbKmers[(size_t)st->vidx];
        st->hasNode = st->vidx >= 0 && meta.hasNode[(size_t)st->vidx];
        if (!st->hasNode) continue;
        for (int t = 0; t < 5; t++) {  // new AccValues(value, dbKMers)
            const int64_t v = st->valueFor(t);
            st->acc[t] = v;
            st->accNorm[t] = st->dbKMers > 0 ? ((double)v) / (double)st->dbKMers : 0;
        }
        st->accErrorSum = st->errorSum; st->accErrorSquaredSum = st->errorSquaredSum;
        st->accClassErrorSum = st->classErrorSum; st->accClassErrorSquaredSum = st->classErrorSquaredSum;
        for (int p = meta.parent[(size_t)st->vidx]; p >= 0; p = meta.parent[(size_t)p]) {  // accumulateFrom
            auto it = taxid2Stats.find(p);
            if (it == taxid2Stats.end()) continue;
            CountsPerTaxid& up = it->second;
            for (int t = 0; t < 5; t++) { up.acc[t] += st->acc[t]; up.accNorm[t] += st->accNorm[t]; }
            up.accErrorSum += st->accErrorSum; up.accErrorSquaredSum += st->accErrorSquaredSum;
            up.accClassErrorSum += st->accClassErrorSum; up.accClassErrorSquaredSum += st->accClassErrorSquaredSum;
        }
    }
    rows.assign(order.begin(), order.end());
}

// ResultReporter.printMatchResult (C/match/ResultReporter.java:190-279); columns in @MDCDescription.pos order
std::string MatchingResult::printMatchResult(const DbMeta& meta) const {
    static const char* const valueTypes[5] = {"reads", "kmers", "reads bps", "read >=1 kmer", "reads kmers"};
    std::string o;
    auto col = [&](const std::string& s) { o += s; o.push_back(';'); };
    for (const char* h : {"pos", "level", "name", "rank", "taxid", "reads", "kmers from reads", "kmers", "unique kmers", "contigs",
                          "average contig length", "max contig length", "reads >=1 kmer", "reads bps", "avg. read length", "db coverage",
                          "exp. unique kmers", "unique kmers / exp.", "db kmers", "parent taxid", "mean error", "kmer error std. dev.",
                          "mean class error", "class error std. dev.", "contig len std. dev."}) col(h);
    for (const char* t : valueTypes) col(std::string("norm. ") + t);
    for (const char* t : valueTypes) { col(std::string("acc. ") + t); col(std::string("acc. norm. ") + t); }
    for (const char* h : {"max contig desc.", "acc. mean error", "acc. error std. dev.", "acc. mean class error", "acc. class error std. dev."}) col(h);
    if (withMaxKMerCounts) col("max kmer counts");
    o.push_back('\n');
    for (const CountsPerTaxid* cp : rows) {
        const CountsPerTaxid& c = *cp;
        const bool notTotal = c.pos != 0;
        auto num = [&](double v, bool allowed) { if (!std::isnan(v) && !std::isinf(v) && allowed) o += javaDoubleToString(v); o.push_back(';'); };
        const double kmers = (double)c.kmers, reads = (double)c.reads, dbk = (double)c.dbKMers;
        col(std::to_string(c.pos));
        col(std::to_string(c.level));
        col(c.hasNode ? meta.name[(size_t)c.vidx] : std::string("TOTAL"));  // completeValues: no node -> "TOTAL"
        col(c.hasNode ? std::string(rankName(meta.rank[(size_t)c.vidx])) : std::string());
        col(c.vidx < 0 ? std::string() : meta.taxid[(size_t)c.vidx]);
        col(std::to_string(c.reads));
        col(std::to_string(c.readsKmers));
        col(std::to_string(c.kmers));
        col(std::to_string(c.uniqueKmers));
        col(std::to_string(c.contigs));
        num(kmers / c.contigs, notTotal);
        col(std::to_string(c.maxContigLen));
        col(std::to_string(c.reads1KMer));
        col(std::to_string(c.readsBPs));
        num(((double)c.readsBPs) / reads, true);                                      // pos 13: printed on the TOTAL row too
        num(((double)c.uniqueKmers) / dbk, notTotal);
        const double expected = (1 - std::pow(1 - 1.0 / dbk, (double)c.kmers)) * dbk;
        num(expected, notTotal);
        num((double)c.uniqueKmers / expected, notTotal);
        col(std::to_string(c.dbKMers));
        col(c.hasNode ? (meta.parent[(size_t)c.vidx] >= 0 ? meta.taxid[(size_t)meta.parent[(size_t)c.vidx]] : std::string()) : std::string());
        num(c.errorSum / reads, notTotal);
        num(std::sqrt((c.errorSquaredSum - c.errorSum * c.errorSum / reads) / (double)(c.reads - 1)), notTotal);
        num(c.classErrorSum / reads, notTotal);
        num(std::sqrt((c.classErrorSquaredSum - c.classErrorSum * c.classErrorSum / reads) / (double)(c.reads - 1)), notTotal);
        num(std::sqrt(((double)c.contigLenSquaredSum - (kmers * kmers) / c.contigs) / (c.contigs - 1)), notTotal);
        for (int t = 0; t < 5; t++) num(((double)c.valueFor(t)) / dbk, notTotal);
        for (int t = 0; t < 5; t++) {
            if (c.hasNode) o += std::to_string(c.acc[t]);
            o.push_back(';');
            if (c.hasNode) o += javaDoubleToString(c.accNorm[t]);
            o.push_back(';');
        }
        col(c.maxContigDescriptor);
        const double accReads = c.hasNode ? (double)c.acc[0] : 0.0;
        const int64_t accReadsL = c.hasNode ? c.acc[0] : 0;
        num(c.accErrorSum / accReads, notTotal);
        num(std::sqrt((c.accErrorSquaredSum - (c.accErrorSum * c.accErrorSum) / accReads) / (double)(accReadsL - 1)), notTotal);
        num(c.accClassErrorSum / accReads, notTotal);
        num(std::sqrt((c.accClassErrorSquaredSum - (c.accClassErrorSum * c.accClassErrorSum) / accReads) / (double)(accReadsL - 1)), notTotal);
        if (withMaxKMerCounts) {
            if (c.hasMaxKMerCounts)
                for (size_t i = 0; i < c.maxKMerCounts.size(); i++) { if (i) o.push_back(';'); o += std::to_string(c.maxKMerCounts[i]); }
            o.push_back(';');
        }
        o.push_back('\n');
    }
    return o;
}

// ---------------------------------------------------------------------------------------------------------
// FastqBloomFilter
// ---------------------------------------------------------------------------------------------------------
FastqBloomFilter::FastqBloomFilter(gs_filter* f, int k, int minPosCount, double positiveRatio, bool withProbs, uint32_t batchReads)
    : f_(f), k_(k), minPosCount_(minPosCount), positiveRatio_(positiveRatio), withProbs_(withProbs), batchReads_(batchReads) {}

void FastqBloomFilter::runFilter(const std::vector<Input>& fastqs, OutputSink* filtered, OutputSink* rest) {
    gs_fsess* s = gs_filter_open(f_, k_, minPosCount_, positiveRatio_);
    if (!s) fail(std::string("gs_filter_open: ") + gs_last_error());
    struct Closer { gs_fsess* s; ~Closer() { gs_filter_close(s); } } closer{s};
    if (filtered && !filtered->open()) fail("cannot open filtered output");
    if (rest && !rest->open()) fail("cannot open rest output");
    const size_t maxInflight = (size_t)GS_MAX_INFLIGHT * (size_t)std::max(1, gs_filter_n_devices(f_));
    std::vector<std::unique_ptr<HostBatch>> pool;
    std::deque<HostBatch*> inflight, freeList;
    for (size_t i = 0; i < maxInflight + 1; i++) { pool.emplace_back(new HostBatch()); freeList.push_back(pool.back().get()); }
    HostBatch* cur = freeList.front(); freeList.pop_front();
    cur->reset(0);
    std::string scratch;
    std::vector<uint8_t> acc;
    auto collectOldest = [&]() {
        HostBatch* b = inflight.front(); inflight.pop_front();
        if (b->isText) {   // records are read from the pinned text through the device's record table
            const uint8_t* a = nullptr; const gs_fastq_rec* recs = nullptr; uint32_t n = 0;
            check(gs_filter_collect_fastq(s, b->ticket, &a, &n, &recs), "gs_filter_collect_fastq");
            for (uint32_t i = 0; i < n; i++) {
                accept.push_back(a[i]);
                acceptedReads += a[i];
                OutputSink* out = a[i] ? filtered : rest;
                if (!out) continue;
                const gs_fastq_rec& rc = recs[i];
                writeRead(*out, b->text + rc.hdr_start, (size_t)(rc.seq_start - 1 - rc.hdr_start), b->text + rc.seq_start, (size_t)rc.seq_len,
                          b->text + rc.qual_start, (size_t)(recs[i + 1].hdr_start - 1 - rc.qual_start), withProbs_, scratch);
            }
            freeList.push_back(b);
            return;
        }
        acc.resize(b->n);
        check(gs_filter_collect(s, b->ticket, acc.data()), "gs_filter_collect");
        for (uint32_t i = 0; i < b->n; i++) {   // nextEntry (:92-105): rewriteInput to `indexed` or `notIndexed`
            accept.push_back(acc[i]);
            acceptedReads += acc[i];
            OutputSink* out = acc[i] ? filtered : rest;
            if (!out) continue;
            size_t dl, pl;
            const uint8_t* d = b->desc(i, dl);
            const uint8_t* p = b->probs(i, pl);
            writeRead(*out, d, dl, b->bases + b->offsets[i], (size_t)(b->offsets[i + 1] - b->offsets[i]), p, pl, b->hasProbs[i] != 0, scratch);
        }
        freeList.push_back(b);
    };
    auto flush = [&]() {
        if (cur->n == 0) return;
        check(gs_filter_submit(s, cur->bases, cur->offsets, cur->n, &cur->ticket), "gs_filter_submit");
        inflight.push_back(cur);
        if (inflight.size() >= maxInflight) collectOldest();
        cur = freeList.front(); freeList.pop_front();
        cur->reset(0);
    };
    FastqReader reader(k_, withProbs_);
    auto onRecord = [&](const Record& r, int64_t) {
        if (cur->n >= batchReads_) flush();
        cur->add(r, k_, withProbs_);
    };
    for (const Input& in : fastqs) {
        if (!gpuParse || in.fasta) {
            reader.readFastq(in, onRecord);
            totalReads += reader.reads; totalKMers += reader.kMers; totalBPs += reader.readBPs;
            continue;
        }
        flush();
        feedFastqText(gs_filter_context(f_), in, textChunkBytes,
            [&]() -> HostBatch* { return cur; },
            [&]() -> HostBatch* { return freeList.front(); },
            [&](HostBatch* b, size_t cut) -> bool {
                gs_fastq_info info;
                gs_ticket t = 0;
                check(gs_filter_submit_fastq(s, b->text, cut, &info, &t), "gs_filter_submit_fastq");
                textChunks++;
                if (info.status) return false;
                b->ticket = t; b->n = info.n_reads;
                totalReads += info.n_reads; totalKMers += (int64_t)info.total_kmers; totalBPs += (int64_t)info.total_bps;
                inflight.push_back(cur);
                if (inflight.size() >= maxInflight) collectOldest();
                cur = freeList.front(); freeList.pop_front();
                cur->reset(0);
                return true;
            },
            [&](LineReader& lr) {
                textChunksRefused++;
                cur->reset(0);
                reader.readFastqFrom(lr, onRecord);
                totalReads += reader.reads; totalKMers += reader.kMers; totalBPs += reader.readBPs;
            });
        if (cur->isText) cur->reset(0);
    }
    flush();
    while (!inflight.empty()) collectOldest();
    if (filtered) filtered->close();
    if (rest) rest->close();
}

}  // namespace gsh

// ---------------------------------------------------------------------------------------------------------
// C entry points of the host layer (what a JNI shim or a test harness calls when it wants the whole goal, not batches)
// ---------------------------------------------------------------------------------------------------------
using namespace gsh;

struct gsh_result {
    std::string csv, filtered, kraken, rest, error;
    int64_t totals[3] = {0, 0, 0};
    std::vector<uint8_t> accept;
    std::vector<double> dsums;  // [4][V]: errorSum, errorSquaredSum, classErrorSum, classErrorSquaredSum
    uint64_t launches = 0;
    uint64_t feeder[2] = {0, 0};  // text chunks split on the GPU / chunks handed back to the sequential parser
};

extern "C" {

DbMeta* gsh_meta_new(int k, int n_values, int64_t total_kmers) {
    DbMeta* m = new DbMeta();
    m->k = k; m->totalKmers = total_kmers; m->resize(n_values);
    return m;
}
void gsh_meta_free(DbMeta* m) { delete m; }
void gsh_meta_set_node(DbMeta* m, int vidx, const char* taxid, const char* name, int rank, int parent, int position, int level, int has_node,
                       int64_t db_kmers) {
    m->taxid[(size_t)vidx] = taxid ? taxid : ""; m->name[(size_t)vidx] = name ? name : ""; m->rank[(size_t)vidx] = rank;
    m->parent[(size_t)vidx] = parent; m->position[(size_t)vidx] = position; m->level[(size_t)vidx] = level; m->hasNode[(size_t)vidx] = has_node;
    m->dbKmers[(size_t)vidx] = db_kmers;
}

typedef struct gsh_match_cfg {
    int classify_reads, count_unique_kmers, use_bloom_filter, max_kmer_res_counts, max_classification_paths, min_kmers_for_class;
    double max_read_tax_error_count, max_read_class_error_count;
    int write_all, with_probs, initial_read_size_bytes, layout, write_filtered, write_kraken;
    uint32_t batch_reads;
    uint32_t text_chunk_bytes;  // GPU FASTQ feeder: 0 = default chunk size, 0xFFFFFFFF = parse on the host only
} gsh_match_cfg;

static std::vector<Input> toInputs(const uint8_t* const* data, const size_t* lens, const char* const* paths, const int* is_fasta, int n) {
    std::vector<Input> in((size_t)n);
    for (int i = 0; i < n; i++) {
        if (paths && paths[i] && paths[i][0]) in[(size_t)i].path = paths[i];
        else { in[(size_t)i].data = data[i]; in[(size_t)i].len = lens[i]; }
        in[(size_t)i].fasta = is_fasta && is_fasta[i];
    }
    return in;
}

// The `match` goal for one key: MatchResultGoal.doMakeThis + MatchGoal.makeFile (C/goals/MatchResultGoal.java:91-164,
// C/goals/MatchGoal.java:84-92).  Inputs are in-memory buffers or file paths; outputs are returned as strings (or written
// to filtered_path / kraken_path when given).
gsh_result* gsh_match_goal(gs_db* db, const DbMeta* meta, const gsh_match_cfg* c, const uint8_t* const* data, const size_t* lens,
                           const char* const* paths, const int* is_fasta, int n_inputs, const char* filtered_path, const char* kraken_path) {
    gsh_result* r = new gsh_result();
    try {
        MatchConfig cfg;
        cfg.classifyReads = c->classify_reads; cfg.countUniqueKMers = c->count_unique_kmers; cfg.useBloomFilterForMatch = c->use_bloom_filter;
        cfg.maxKMerResCounts = c->max_kmer_res_counts; cfg.maxClassificationPaths = c->max_classification_paths;
        cfg.minKMersForClass = c->min_kmers_for_class; cfg.maxReadTaxErrorCount = c->max_read_tax_error_count;
        cfg.maxReadClassErrorCount = c->max_read_class_error_count; cfg.writeAll = c->write_all; cfg.withProbs = c->with_probs;
        cfg.initialReadSizeBytes = c->initial_read_size_bytes; cfg.layout = c->layout;
        if (c->batch_reads) cfg.batchReads = c->batch_reads;
        if (c->text_chunk_bytes == 0xFFFFFFFFu) cfg.gpuParse = false;
        else if (c->text_chunk_bytes) cfg.textChunkBytes = c->text_chunk_bytes;
        FastqKMerMatcher matcher(db, *meta, cfg);
        OutputSink filtered, kraken;
        if (filtered_path && filtered_path[0]) filtered.path = filtered_path; else filtered.mem = &r->filtered;
        if (kraken_path && kraken_path[0]) kraken.path = kraken_path; else kraken.mem = &r->kraken;
        MatchingResult res = matcher.runMatcher(toInputs(data, lens, paths, is_fasta, n_inputs), c->write_filtered ? &filtered : nullptr,
                                                c->write_kraken ? &kraken : nullptr);
        res.completeResults(*meta);
        r->csv = res.printMatchResult(*meta);
        r->totals[0] = res.totalReads; r->totals[1] = res.totalKMers; r->totals[2] = res.totalBPs;
        r->launches = matcher.kernelLaunches();
        r->feeder[0] = matcher.textChunks; r->feeder[1] = matcher.textChunksRefused;
        const int V = meta->nValues;
        r->dsums.assign((size_t)4 * V, 0.0);
        for (auto& kv : res.taxid2Stats) {
            r->dsums[(size_t)0 * V + kv.first] = kv.second.errorSum; r->dsums[(size_t)1 * V + kv.first] = kv.second.errorSquaredSum;
            r->dsums[(size_t)2 * V + kv.first] = kv.second.classErrorSum; r->dsums[(size_t)3 * V + kv.first] = kv.second.classErrorSquaredSum;
        }
    } catch (const std::exception& e) { r->error = e.what(); }
    return r;
}

// The `filter` goal: FilterGoal.makeFile (C/goals/FilterGoal.java:80-108)
gsh_result* gsh_filter_goal(gs_filter* f, int k, int min_pos_count, double pos_ratio, int with_probs, uint32_t batch_reads,
                            const uint8_t* const* data, const size_t* lens, const char* const* paths, const int* is_fasta, int n_inputs,
                            const char* filtered_path, const char* rest_path, int want_rest, uint32_t text_chunk_bytes) {
    gsh_result* r = new gsh_result();
    try {
        FastqBloomFilter flt(f, k, min_pos_count, pos_ratio, with_probs != 0, batch_reads ? batch_reads : (1u << 20));
        if (text_chunk_bytes == 0xFFFFFFFFu) flt.gpuParse = false;
        else if (text_chunk_bytes) flt.textChunkBytes = text_chunk_bytes;
        OutputSink filtered, rest;
        if (filtered_path && filtered_path[0]) filtered.path = filtered_path; else filtered.mem = &r->filtered;
        if (rest_path && rest_path[0]) rest.path = rest_path; else rest.mem = &r->rest;
        flt.runFilter(toInputs(data, lens, paths, is_fasta, n_inputs), &filtered, want_rest ? &rest : nullptr);
        r->totals[0] = flt.totalReads; r->totals[1] = flt.totalKMers; r->totals[2] = flt.totalBPs;
        r->feeder[0] = flt.textChunks; r->feeder[1] = flt.textChunksRefused;
        r->accept.swap(flt.accept);
    } catch (const std::exception& e) { r->error = e.what(); }
    return r;
}

// Parser only (no GPU): every record of the inputs rewritten by ReadEntry.write, plus the reader's totals.  Lets the
// CPU test-suite check the host-side FASTQ/FASTA parsing against the oracle without a device.
gsh_result* gsh_parse_only(int k, int with_probs, const uint8_t* const* data, const size_t* lens, const char* const* paths,
                           const int* is_fasta, int n_inputs) {
    gsh_result* r = new gsh_result();
    try {
        OutputSink out;
        out.mem = &r->rest;
        std::string scratch;
        FastqReader reader(k, with_probs != 0);
        for (const Input& in : toInputs(data, lens, paths, is_fasta, n_inputs)) {
            reader.readFastq(in, [&](const Record& rec, int64_t) {
                writeRead(out, rec.descriptor.data(), rec.descriptor.size(), rec.read.data(), rec.read.size(), rec.probs.data(),
                          rec.probs.size(), with_probs != 0 && rec.hasProbs, scratch);
                r->accept.push_back((uint8_t)rec.entry);
            });
            r->totals[0] += reader.reads; r->totals[1] += reader.kMers; r->totals[2] += reader.readBPs;
        }
    } catch (const std::exception& e) { r->error = e.what(); }
    return r;
}

// The feeder's block-gzip reader alone (no GPU): the inflated text of a BGZF file, read in requests of `request` bytes;
// an ordinary gzip member in the middle ends the text there (*foreign_at = its compressed offset, else -1).
gsh_result* gsh_bgzf_read_all(const char* path, size_t request, int64_t* foreign_at) {
    gsh_result* r = new gsh_result();
    *foreign_at = -1;
    int fd = -1;
    try {
        fd = open(path, O_RDONLY);
        if (fd < 0) fail(std::string("cannot open ") + path);
        if (!BgzfReader::probe(fd)) fail("not a block-gzip file");
        BgzfReader rd(fd);
        std::vector<uint8_t> buf(std::max<size_t>(request, 1));
        for (;;) {
            const size_t got = rd.read(buf.data(), buf.size());
            r->rest.append((const char*)buf.data(), got);
            if (rd.foreign()) { *foreign_at = (int64_t)rd.foreignAt(); break; }
            if (got < buf.size()) break;
        }
    } catch (const std::exception& e) { r->error = e.what(); }
    if (fd >= 0) close(fd);
    return r;
}

uint64_t gsh_device_inflated_blocks(void) { return g_deviceInflatedBlocks.load(); }
// [0] seconds inside gs_inflate_blocks, [1] calls, [2] seconds inside BgzfReader::read (either inflater), since the process started
void gsh_bgzf_timers(double* out) { out[0] = g_deviceInflateNanos.load() * 1e-9; out[1] = (double)g_deviceInflateCalls.load(); out[2] = g_bgzfReadNanos.load() * 1e-9; }

// The feeder's record-boundary search alone (no GPU): offset of the last record start in text[0, len), 0 = none in sight.
size_t gsh_last_record_start(const uint8_t* text, size_t len) { return lastRecordStart(text, len); }

void gsh_result_free(gsh_result* r) { delete r; }
const char* gsh_result_error(const gsh_result* r) { return r->error.c_str(); }
const char* gsh_result_text(const gsh_result* r, int which, size_t* len) {
    const std::string& s = which == 0 ? r->csv : which == 1 ? r->filtered : which == 2 ? r->kraken : r->rest;
    *len = s.size();
    return s.data();
}
void gsh_result_totals(const gsh_result* r, int64_t* out) { out[0] = r->totals[0]; out[1] = r->totals[1]; out[2] = r->totals[2]; }
const uint8_t* gsh_result_accept(const gsh_result* r, size_t* n) { *n = r->accept.size(); return r->accept.data(); }
const double* gsh_result_dsums(const gsh_result* r, size_t* n) { *n = r->dsums.size(); return r->dsums.data(); }
uint64_t gsh_result_launches(const gsh_result* r) { return r->launches; }
void gsh_result_feeder(const gsh_result* r, uint64_t* out) { out[0] = r->feeder[0]; out[1] = r->feeder[1]; }
int gsh_java_double_to_string(double v, char* buf, int cap) {
    const std::string s = javaDoubleToString(v);
    if ((int)s.size() + 1 > cap) return -1;
    memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}

}  // extern "C"
