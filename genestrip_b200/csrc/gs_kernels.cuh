// gs_kernels.cuh -- kernel parameter blocks and launchers shared by gs_kernels.cu and gs_capi.cu
#pragma once
#include "../../include/genestrip_b200.h"
#include "gs_device.cuh"

struct GsMatchParams {
    GsDbView db;
    const uint8_t* bases;   // reads back to back, ASCII
    const u64* packCodes;   // or (gs_pack.hpp) the same bases packed by the host: 2-bit codes, 32 per word, first base on top ...
    const u32* packValid;   // ... and a validity bit per base; flat position 0 = the batch's first base (lead = 0)
    u32 packSegs;           // segments [0, packSegs) of the label kernel are staged from the packed words, the others from `bases`
    const u64* offsets;     // [nReads + 1]
    u32 nReads;
    u64 firstReadNo;
    gs_read_result* out;    // [nReads]
    long long* counters;    // [7][nValues]: kmers, contigs, contigLenSquaredSum, reads1KMer, reads, readsKmers, readsBPs
    u64* maxcontig;         // [nValues] (maxContigLen << 40) | (2^40-1 - read ordinal)
    u64* bitset;            // unique k-mer bits by storage position, or NULL
    u32* seenTab;           // probe table as u32 words when the session leases its in-entry seen bits (then bitset == NULL)
    uint16_t* hitCounts;    // per-position hit counters (maxKMerResCounts > 0), or NULL
    int classify, useBloom, maxPaths, threshold;
    int layout;             // GS_LAYOUT_TABLE / GS_LAYOUT_CLASSIC
    double maxTaxErr, maxClassErr;
    u32* overflowList;      // reads with more than GS_TABLE_CAP distinct taxa
    u32* overflowCount;
    u32* workCounter;       // next unclaimed read of the batch (dynamic distribution over the persistent warps)
    u32* errFlag;           // set when a read's offsets are malformed (descending / longer than 2^31)
    u32* slowTable;         // MODE 1: per-warp vote tables in global memory, 2 * nValues u32 each
    // flat geometry of the batch (label kernel): flat position f = byte offset - off0 + lead, lead = alignment bytes in front
    u64 off0, flatLen;
    u32 lead;
    u32* labels;            // [flatLen] label of the k-mer starting at each flat position
    u32* validBits;         // [segments * 31 (+pad)] bit f = base f is one of CGAT
    u32* startBits;         // same size: bit f = a read starts at f
    u32* bmask;             // same size, or NULL: bit f = the label at f differs from the label at f - 1 (long-read reduce)
    u32* segCounter;        // next unclaimed segment of the label kernel
    u32* redoList;          // reads the thread-per-read reduce kernel hands to the warp-per-read kernel (NULL: warp kernel takes all reads)
    u32* redoCount;
    u32* groupCounter;      // next unclaimed group of 32 reads of the thread-per-read kernel
    long long* flatPos;     // label dump only: storage position per flat position
    // kraken-style runs (want_runs)
    gs_run* runs; const u64* runOffsets; u64 runsCap; u32* runCounts;
    // label dump (parity tests)
    const u64* kmerOffsets; int* dumpLabels; long long* dumpPos;
};

struct GsFilterView {
    int kind;
    long long p0, p1;        // blocked: seed, buckets; hashed: bits, hashes
    u64 magic;               // floor((2^64-1)/modulus)
    const long long* factors;
    const u64* words;
};

struct GsFilterParams {
    GsFilterView f;
    const uint8_t* bases;
    const u64* offsets;
    u32 nReads;
    int k, minPosCount;
    double posRatio;
    uint8_t* accept;
    u32* errFlag;
    // flat geometry of the batch, as in GsMatchParams: flat position f = byte offset - off0 + lead
    u64 off0, flatLen;
    u32 lead;
    u32* startBits;         // [segments * 31 (+pad)] bit f = a read starts at f (zeroed by the caller, set by gs_mark_starts_kernel)
    u32* hitBits;           // same size: bit f = the k-mer that starts at f is a k-mer of a read and is contained in the filter
    u32* segCounter;        // next unclaimed segment (zeroed by the caller)
};

void gs_launch_mark_starts(const GsMatchParams& P, cudaStream_t st);
void gs_launch_label(const GsMatchParams& P, bool dump, int blocks, cudaStream_t st);
void gs_launch_reduce(const GsMatchParams& P, int mode, bool dump, int blocks, cudaStream_t st);
void gs_launch_reduce_thread(const GsMatchParams& P, int blocks, cudaStream_t st);
void gs_launch_maxcontig_events(const u64* maxcontig, int V, u64 firstReadNo, u32 nReads, gs_maxcontig_event* ev, u32* nEv, cudaStream_t st);
u64 gs_popcount_scratch_words(int blocks, int nValues);
void gs_launch_unique_popcount(const u64* bits, u64 wordBegin, u64 wordEnd, const GsDbView& db, int layout, long long* unique, u32* partial, int blocks, cudaStream_t st);
void gs_launch_collect_hits(const u64* bits, u64 nWords, const uint16_t* hitCounts, const GsDbView& db, int layout, u32* out, unsigned long long* nOut, u64 cap, cudaStream_t st);
void gs_launch_table_clear_seen(u64* tab, u64 nSlots, cudaStream_t st);
void gs_launch_table_extract_seen(const u64* tab, u64 nSlots, u64* out, cudaStream_t st);
u64 gs_table_scan_blocks(u64 nBuckets);
void gs_launch_table_build(const u64* keys, const uint16_t* vals, u64 n, u64* tab, u32* counts, u32* delta, void* agg, u32* blockIn,
                           u32* bad, u64 nBuckets, int rbits, cudaStream_t st);
void gs_launch_mz_build(const u64* keys, u64 n, int k, u64* filter, u32 mask, int wide, cudaStream_t st);
// end-of-run merge across GPUs: per-rank source pointers (peer mappings or receive buffers), see gs_merge_or_popcount_kernel
#define GS_MAX_RANKS 64
struct GsPeerPtrs { const u64* p[GS_MAX_RANKS]; };
void gs_launch_merge_or_popcount(const GsPeerPtrs& src, int nSrc, u64* own, u64 wordBegin, u64 wordEnd, const GsDbView& db, int layout,
                                 long long* unique, u32* partial, int blocks, cudaStream_t st);
void gs_launch_merge_add_u16(const GsPeerPtrs& src, int nSrc, uint16_t* own, u64 begin, u64 end, cudaStream_t st);
void gs_launch_bucket_index(const u64* keys, u64 n, int bshift, u64 nb, u32* bstart, cudaStream_t st);
void gs_launch_bloom_build(const u64* keys, u64 n, u64* words, u64 buckets, u64 magic, long long seed, cudaStream_t st);
void gs_launch_convert_values(const int16_t* raw, const int* hasNode, u64 n, int V, uint16_t* vals, u32* bad, cudaStream_t st);
void gs_launch_check_sorted(const u64* keys, u64 n, int k, u32* bad, cudaStream_t st);
void gs_launch_lookup(const GsDbView& db, const u64* kmers, u64 n, int useBloom, int* vidx, long long* pos, cudaStream_t st);
void gs_launch_filter_contains(const GsFilterView& f, const u64* kmers, u64 n, uint8_t* out, cudaStream_t st);
void gs_launch_filter(const GsFilterParams& P, int blocks, cudaStream_t st);
int gs_match_kernel_occupancy(int mode);

// gs_text.cu: FASTQ record splitting and base compaction on the device
#define GS_TEXT_SEG 16384          // bytes per block segment
void gs_launch_text_split(const uint8_t* text, u64 n, u32* blockCounts, u32* lineEnd, u32 lineCap, u32* meta, gs_fastq_rec* recs, u32* lens,
                          int k, unsigned long long* totals, cudaStream_t st);
void gs_launch_text_compact(const uint8_t* text, const gs_fastq_rec* recs, const u32* lens, u32 n, u64* tileSums, u64* offsets, uint8_t* bases, cudaStream_t st);
void gs_launch_text_kmer_offsets(const u32* lens, u32 n, int k, u32* klens, u64* tileSums, u64* kmerOff, cudaStream_t st);
void gs_launch_text_event_headers(const gs_maxcontig_event* ev, const u32* nEv, u32 evCap, const gs_fastq_rec* recs, u64 firstReadNo, u32 n, u32* hdr, cudaStream_t st);

// database update phase
void gs_launch_db_update(const GsDbView& db, const u32* labels, const long long* flatPos, u64 flatLen, const u64* offsets, const int* regionNode,
                         u32 nRegions, uint16_t* vals, unsigned long long* nChanged, cudaStream_t st);
void gs_launch_cgat_upper(uint8_t* buf, u64 n, cudaStream_t st);
void gs_launch_values_to_raw(const uint16_t* vals, const int16_t* kept, u64 n, int16_t* raw, cudaStream_t st);

// gs_inflate.cu: one thread per raw-deflate block (block-gzip input), status per block
void gs_launch_inflate_blocks(const uint8_t* comp, uint8_t* text, gs_deflate_block* blocks, uint32_t nBlocks, cudaStream_t st);
