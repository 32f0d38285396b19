// gs_host.hpp -- C++ host side above the C ABI: the mirror of the reference's goal drivers for this path.
//
// The reference's host is Java (no JDK in this image), so the host side is written in C++ with the reference's class
// names, argument meaning and error behaviour:
//   FastqReader           <- AbstractFastqReader.doReadFastq / doReadFasta  (C/fastq/AbstractFastqReader.java:288-438)
//                            + BufferedLineReader                            (B/io/BufferedLineReader.java:114-182)
//   FastqKMerMatcher      <- FastqKMerMatcher.runMatcher / afterMatch        (C/match/FastqKMerMatcher.java:181-315)
//   CountsPerTaxid        <- CountsPerTaxid                                   (C/match/CountsPerTaxid.java:127-181, 593-622)
//   MatchingResult        <- MatchingResult.completeResults                   (C/match/MatchingResult.java:84-118)
//   printMatchResult      <- ResultReporter.printMatchResult                  (C/match/ResultReporter.java:190-279)
//   FastqBloomFilter      <- FastqBloomFilter.runFilter / nextEntry           (C/bloom/FastqBloomFilter.java:80-105)
// The per-read kernels (matchRead, isAcceptRead) run on the GPU behind include/genestrip_b200.h; the parser streams reads
// into pinned, double-buffered batches.  Outputs follow the reference at threads=0 (input order).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/genestrip_b200.h"

namespace gsh {

// What Database / SmallTaxTree / the store's statistics give the Java host, flattened by value index.
struct DbMeta {
    int k = 31;
    int nValues = 0;
    int64_t totalKmers = 0;                 // store.getEntries() == getStats().getLong(null)
    std::vector<std::string> taxid, name;   // indexMap value / SmallTaxIdNode.getName() ("" = null)
    std::vector<int> rank;                  // Rank ordinal or -1 (C/tax/Rank.java:39-122)
    std::vector<int> parent, position, level, hasNode;
    std::vector<int64_t> dbKmers;           // getNKmersPerTaxid per value
    void resize(int n);
};
const char* rankName(int ordinal);

struct Input {
    const uint8_t* data = nullptr;  // in-memory input, or
    size_t len = 0;
    std::string path;               // a file (gzip by suffix .gz / .gzip, B/io/StreamProvider.java:93-150)
    bool fasta = false;
};

struct MatchConfig {  // the config keys the path reads (C/GSConfigKey.java:302-364)
    bool classifyReads = true, countUniqueKMers = true, useBloomFilterForMatch = true;
    int maxKMerResCounts = 0, maxClassificationPaths = 10, minKMersForClass = 1;
    double maxReadTaxErrorCount = -1, maxReadClassErrorCount = -1;
    bool writeAll = true, withProbs = false;
    int initialReadSizeBytes = 4096;
    int layout = GS_LAYOUT_TABLE;
    uint32_t batchReads = 1u << 20;       // reads per pinned batch
    size_t batchBytes = (size_t)256 << 20;  // bases per pinned batch
    // FASTQ inputs (not FASTA): raw text chunks go to the GPU, which splits the records
    // (gs_match_submit_fastq); the first chunk the device refuses switches the rest of that input to the sequential parser
    bool gpuParse = true;
    size_t textChunkBytes = (size_t)64 << 20;
};

struct CountsPerTaxid {
    int level = 0;
    int vidx = -1;  // -1 = TOTAL row (taxid null)
    int64_t reads = 0, reads1KMer = 0, readsBPs = 0, readsKmers = 0, uniqueKmers = 0, kmers = 0;
    int32_t contigs = 0;
    int64_t contigLenSquaredSum = 0;
    int32_t maxContigLen = 0;
    std::string maxContigDescriptor;
    bool hasMaxKMerCounts = false;
    std::vector<int16_t> maxKMerCounts;
    double errorSum = 0, errorSquaredSum = 0, classErrorSum = 0, classErrorSquaredSum = 0;
    // completeValues
    int pos = 0;
    int64_t dbKMers = 0;
    bool hasNode = false;
    int64_t acc[5] = {0, 0, 0, 0, 0};
    double accNorm[5] = {0, 0, 0, 0, 0};
    double accErrorSum = 0, accErrorSquaredSum = 0, accClassErrorSum = 0, accClassErrorSquaredSum = 0;
    int64_t valueFor(int type) const;  // ValueType order: reads, kmers, reads bps, read >=1 kmer, reads kmers
};

struct MatchingResult {
    int k = 31;
    CountsPerTaxid globalStats;
    std::map<int, CountsPerTaxid> taxid2Stats;  // by value index
    bool withMaxKMerCounts = false;
    std::vector<const CountsPerTaxid*> rows;    // after completeResults: TOTAL first, then tree pre-order
    int64_t totalReads = 0, totalKMers = 0, totalBPs = 0;
    void completeResults(const DbMeta& meta);
    std::string printMatchResult(const DbMeta& meta) const;
};

std::string javaDoubleToString(double v);

struct OutputSink {  // filtered FASTQ / kraken-style out: in-memory string or file (gzip if path ends in .gz)
    std::string* mem = nullptr;
    std::string path;
    void* gz = nullptr;
    FILE* fp = nullptr;
    bool open();
    void write(const char* p, size_t n);
    void close();
};

class FastqKMerMatcher {
   public:
    FastqKMerMatcher(gs_db* db, const DbMeta& meta, const MatchConfig& cfg);
    ~FastqKMerMatcher();
    // FastqKMerMatcher.runMatcher: all inputs of one key; filtered / krakenOut may be null.  Throws std::runtime_error
    // (the JNI shim rethrows it as RuntimeException like consumer-thread failures, AbstractFastqReader.java:124-143).
    MatchingResult runMatcher(const std::vector<Input>& fastqs, OutputSink* filtered, OutputSink* krakenOut);
    uint64_t kernelLaunches() const { return launches_; }

   private:
    struct Batch;
    void processBatch(gs_sess* s, Batch& b, OutputSink* filtered, OutputSink* krakenOut, std::vector<CountsPerTaxid>& stats,
                      std::vector<uint64_t>& bestKey);
    void processTextBatch(gs_sess* s, Batch& b, OutputSink* filtered, OutputSink* krakenOut, std::vector<CountsPerTaxid>& stats,
                          std::vector<uint64_t>& bestKey);
    void writeKrakenLine(OutputSink& krakenOut, const gs_read_result& r, const uint8_t* desc, size_t descLen, int64_t L, const gs_run* runs,
                         size_t nRuns, int entry, std::string& line);

   public:
    uint64_t textChunks = 0, textChunksRefused = 0;  // chunks split on the GPU / handed back to the sequential parser

   private:
    gs_db* db_;
    const DbMeta& meta_;
    MatchConfig cfg_;
    bool entryBufferUsed_[2] = {false, false};  // MatcherReadEntry.buffer != null of the two pooled entries (threads=0)
    uint64_t launches_ = 0;
};

class FastqBloomFilter {
   public:
    FastqBloomFilter(gs_filter* f, int k, int minPosCount, double positiveRatio, bool withProbs, uint32_t batchReads = 1u << 20);
    // FastqBloomFilter.runFilter: accepted reads -> filtered, the others -> rest (either may be null)
    void runFilter(const std::vector<Input>& fastqs, OutputSink* filtered, OutputSink* rest);
    int64_t totalReads = 0, totalKMers = 0, totalBPs = 0, acceptedReads = 0;
    std::vector<uint8_t> accept;  // per read, input order (kept for the parity tests)
    bool gpuParse = true;         // FASTQ inputs as raw text chunks, records split on the GPU (see MatchConfig::gpuParse)
    size_t textChunkBytes = (size_t)64 << 20;
    uint64_t textChunks = 0, textChunksRefused = 0;

   private:
    gs_filter* f_;
    int k_, minPosCount_;
    double positiveRatio_;
    bool withProbs_;
    uint32_t batchReads_;
};

}  // namespace gsh
