/* genestrip_b200.h -- C ABI of the B200-native Genestrip read-matching hot path.
 *
 * This is the drop-in boundary (DESIGN.md "Boundary").  The reference (pfeiferd/genestrip v3.0) is pure
 * Java and has no FFI for this path; its seams are Java override points.  Each entry point below names
 * the reference interface it replaces, with
 *   C/ = core/src/main/java/org/metagene/genestrip/
 * A JNI shim (integration/jni/gs_jni.cpp, shown in INTEGRATION.md) binds exactly these symbols.
 *
 * Conventions: plain pointers and sizes only; every function returning int returns 0 on success and a
 * negative gs_status on failure (text via gs_last_error(), thread-local); nothing throws or calls back
 * into the host.  There is NO CPU fallback: without a CUDA device every compute entry point fails with
 * GS_ERR_CUDA.
 */
#ifndef GENESTRIP_B200_H
#define GENESTRIP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS_ABI_VERSION 8

typedef enum gs_status {
    GS_OK = 0,
    GS_ERR_ARG = -1,     /* bad argument / call order */
    GS_ERR_CUDA = -2,    /* CUDA runtime error (incl. no device) */
    GS_ERR_STATE = -3,   /* object not finalized / ticket not pending */
    GS_ERR_LIMIT = -4,   /* a documented size limit was exceeded */
    GS_ERR_DATA = -5     /* corrupt input data (block-gzip: malformed deflate stream, size or CRC-32 mismatch) */
} gs_status;

typedef struct gs_ctx gs_ctx;       /* one per process: the set of GPUs used */
typedef struct gs_db gs_db;         /* device-resident database: k-mer store + tax tree + Bloom prefilter */
typedef struct gs_sess gs_sess;     /* one `match` run over one fastq key (FastqKMerMatcher.runMatcher) */
typedef struct gs_filter gs_filter; /* device-resident `filter` goal index (KMerProbFilter) */
typedef struct gs_fsess gs_fsess;   /* one `filter` run (FastqBloomFilter.runFilter) */
typedef uint64_t gs_ticket;

int gs_abi_version(void);
const char* gs_last_error(void);

/* ---- context ------------------------------------------------------------------------------------
 * Replaces: ExecutionContext / thread pool sizing (C/DefaultExecutionContext.java:78) -- the unit of
 * parallelism is a GPU instead of a consumer thread.  device_ordinals == NULL, n == 0 => device 0. */
gs_ctx* gs_ctx_create(const int* device_ordinals, int n_devices);
void gs_ctx_destroy(gs_ctx*);
int gs_ctx_n_devices(const gs_ctx*);

/* Pinned host memory for read batches (replaces the ReadEntry pool, C/fastq/AbstractFastqReader.java:85-104). */
void* gs_alloc_pinned(size_t bytes);
void gs_free_pinned(void*);

/* ---- database -----------------------------------------------------------------------------------
 * Replaces: Database.load + convertKMerStore (C/store/Database.java:136-143, 265-314) as the data source,
 * KMerSortedArray's arrays (C/store/KMerSortedArray.java:63-66) as the layout that is uploaded.
 * keys: sorted ascending, distinct, < 2^62, storage position == array index (KMerSortedArray.getLong :298-349).
 * vidx_raw: the Java short as stored (value index + Short.MIN_VALUE).  Segments may be streamed in any order; the
 * source pointers may be host or device memory (cudaMemcpyDefault). */
gs_db* gs_db_create(gs_ctx*, int k, uint64_t n_kmers, int n_values);
int gs_db_put_keys(gs_db*, uint64_t offset, const int64_t* keys, uint64_t n);
int gs_db_put_values(gs_db*, uint64_t offset, const int16_t* vidx_raw, uint64_t n);
/* RadixKMerStore source (C/store/RadixKMerStore.java:369-412, 714-730): entries of one bucket, packed as
 * (valueIndex << (62 - radix_bits)) | (kmer >>> radix_bits), sorted by the remaining bits.  The library
 * rebuilds full keys and merges them into its own sorted layout (results are identical by
 * T/match/RadixKMerStoreBenchmarkTest.java:186-229).  Call once per non-empty bucket, then finalize. */
int gs_db_put_radix_bucket(gs_db*, int radix_bits, uint32_t radix, const int64_t* entries, uint32_t n);
/* SmallTaxTree flattened by value index (C/tax/SmallTaxTree.java; storeIndex == value index, Database.java:107-128).
 * parent_by_vidx[v] = value index of the parent node, -1 for the root.  has_node[v] == 0 marks a stored value
 * whose tax id has no tree node: such k-mers convert to null (Database.java:136-143) and count as misses.
 * has_node may be NULL (all present). */
int gs_db_set_tree(gs_db*, const int32_t* parent_by_vidx, const int32_t* has_node, int n_values);
/* BlockedKMerBloomFilter as deserialized from bloom.ser (C/bloom/BlockedKMerBloomFilter.java:181-198):
 * seed, buckets, data[buckets+17]. */
int gs_db_set_bloom_blocked(gs_db*, int64_t seed, uint64_t buckets, const int64_t* words, uint64_t n_words);
/* Build the store's optimized filter on the device exactly as KMerSortedArray.optimize does
 * (C/store/KMerSortedArray.java:409-422 with AbstractKMerStore.createOptimizedFilter :271-285, 10 bits/key,
 * seed = new Random(42).nextLong()).  Needs all keys uploaded.  words_out (may be NULL) receives the
 * buckets+17 words for comparison with a host-built filter. */
int gs_db_build_bloom_blocked(gs_db*, int64_t* words_out, uint64_t n_words_out);
int gs_db_finalize(gs_db*); /* builds the bucket index, replicates to every device of the context */
/* The update phase of the `db` goal (C/goals/refseq/DBGoal.java:234-311, C/store/KMerSortedArray.java update): for every
 * stored k-mer that occurs in a genome region, value = LCA(value, node of the region) (unchanged when there is no common
 * ancestor or the value has no tree node).  seq = the regions' sequence bytes back to back WITHOUT line terminators (the
 * FASTA reader strips them, C/refseq/AbstractStoreFastaReader.java:88-120), region r = [region_offsets[r],
 * region_offsets[r + 1]) with node value index region_vidx[r] (< 0: region skipped); the k-mer window restarts at every region
 * and at every byte that is not one of CGAT (after cgatToUpperCase when upper_case != 0, C/util/CGAT.java:91-99).
 * stepSize = 1 and no dust filter (the defaults).  Needs a finalized database with its tree and no open unique-counting
 * session; the order of calls and regions does not matter (LCA is associative and commutative).  n_changed (may be NULL)
 * counts value changes.  gs_db_get_values reads the Java shorts back in storage order (value index + Short.MIN_VALUE; -1 for
 * values without a tree node). */
int gs_db_update(gs_db*, const uint8_t* seq, uint64_t n_bytes, const uint64_t* region_offsets, const int32_t* region_vidx,
                 uint32_t n_regions, int upper_case, uint64_t* n_changed);
int gs_db_get_values(gs_db*, uint64_t offset, int16_t* vidx_raw, uint64_t n);
/* Flat little-endian database file "GSB1" (layout in gs_capi.cu): the arrays Database.save serializes with Java object streams
 * (C/store/Database.java:201-260: k-mers, value indices, tax tree by value index, blocked Bloom filter), written by
 * gs_db_save_file -- e.g. after gs_db_update -- or by the Java-side exporter (integration/java/.../GsbExporter.java), and
 * loaded without a JVM by gs_db_load_file (create + put_* + set_tree + set_bloom_blocked + finalize). */
int gs_db_info(const gs_db*, int* k, uint64_t* n_kmers, int* n_values);
int gs_db_save_file(gs_db*, const char* path);
gs_db* gs_db_load_file(gs_ctx*, const char* path);
void gs_db_destroy(gs_db*);
uint64_t gs_db_device_bytes(const gs_db*);
int gs_db_n_devices(const gs_db*);
/* Lookup probe for tests (KMerStore.getLong, C/store/KMerStore.java:157): vidx_out[i] = value index or -1
 * (miss / value without node), pos_out[i] = storage position or -1.  use_bloom follows useBloomFilterForMatch. */
int gs_db_lookup(gs_db*, const int64_t* kmers, uint64_t n, int use_bloom, int32_t* vidx_out, int64_t* pos_out);

/* ---- match --------------------------------------------------------------------------------------
 * Replaces: MatchResultGoal.createMatcher (C/goals/MatchResultGoal.java:174-197) -> FastqKMerMatcher
 * (C/match/FastqKMerMatcher.java:127-147) and its config keys (C/GSConfigKey.java:302-350). */
typedef struct gs_match_cfg {
    int classify_reads;                /* classifyReads && !matchlr  (taxTree != null)          */
    int count_unique_kmers;            /* countUniqueKMers                                       */
    int max_kmer_res_counts;           /* maxKMerResCounts (>0: per-k-mer hit counters)          */
    int use_bloom_filter;              /* useBloomFilterForMatch                                 */
    int max_classification_paths;      /* maxClassificationPaths, 1..128                         */
    int min_kmers_for_class;           /* minKMersForClass (threshold)                           */
    double max_read_tax_error_count;   /* maxReadTaxErrorCount                                   */
    double max_read_class_error_count; /* maxReadClassErrorCount                                 */
    int want_runs;                     /* writeKrakenStyleOut: return per-read contig runs       */
    int layout;                        /* device index: GS_LAYOUT_TABLE (default) or GS_LAYOUT_CLASSIC; same results  */
    int prefilter;                     /* 1 (default): skip the probe table for k-mers whose minimizer is in no stored k-mer
                                          (L2-resident bit filter built by gs_db_finalize; same results, k >= 24 only)   */
    int host_pack_threads;             /* how gs_match_submit moves the bases to the device.  0: as they are, 1 byte per base.
                                          n > 0: n host threads (the caller included) first pack them to 2-bit codes + a
                                          validity bit, 0.375 bytes per base on the link; -1 (default): n = the CPUs this
                                          process may run on (two fewer from eight on), at most 32.  Same results either way. */
    int host_pack_percent;             /* with host_pack_threads != 0: the share of every batch that is packed, the rest crosses
                                          the link as ASCII while the pool packs (link and cores work side by side).
                                          -1 (default): follows the measured cost of the two routes; 0..100: fixed.       */
} gs_match_cfg;
#define GS_LAYOUT_TABLE 0   /* 128-byte probe table built from the store's arrays: one DRAM line touch per k-mer      */
#define GS_LAYOUT_CLASSIC 1 /* the reference's own structures: blocked Bloom filter + binary search of the sorted array */
void gs_match_cfg_default(gs_match_cfg*);

/* Per-read result, 16 bytes (what FastqKMerMatcher.matchRead leaves in MatcherReadEntry, :327-535). */
typedef struct gs_read_result {
    int32_t class_vidx;   /* entry.classNode's value index, -1 = null                                  */
    uint32_t read_kmers;  /* readKmers (:506-507); 0 if unclassified                                   */
    uint32_t tax_err;     /* readTaxErrorCount at the end of the loop; 0xFFFFFFFF = -1 (gate closed/off) */
    uint32_t flags;       /* GS_READ_* */
} gs_read_result;
#define GS_READ_FOUND 1u      /* matchRead's return value (drives the filtered FASTQ, :304-307)        */
#define GS_READ_ACCEPTED 2u   /* the classified-read statistics were updated (:509-526)                 */
#define GS_READ_SLOWPATH 4u   /* more distinct taxa than the fast path tracks; resolved by the slow path */

/* One contig run of a read for the kraken-style line (printKrakenStyleOut, :597-611). */
typedef struct gs_run {
    uint32_t label; /* value index, GS_RUN_MISS ('0') or GS_RUN_INVALID ('A') */
    uint32_t len;
} gs_run;
#define GS_RUN_MISS 0xFFFFFFFEu
#define GS_RUN_INVALID 0xFFFFFFFDu

/* A new per-taxon maximum contig length was set by a read of this batch (:402-409); the host copies the
 * read's descriptor while it still has the batch. */
typedef struct gs_maxcontig_event {
    uint32_t vidx;
    uint32_t contig_len;
    uint64_t read_no;
} gs_maxcontig_event;

/* Per value index: the integer fields of CountsPerTaxid (C/match/CountsPerTaxid.java:127-159). */
typedef struct gs_taxon_counts {
    int64_t kmers, contigs, contig_len_squared_sum, reads_1kmer, reads, reads_kmers, reads_bps, unique_kmers;
    int32_t max_contig_len;
    int32_t touched;             /* statsIndex[vi] != null (a row exists even if all counters are 0) */
    uint64_t max_contig_read_no; /* first read (lowest ordinal) that reached max_contig_len */
} gs_taxon_counts;

gs_sess* gs_match_open(gs_db*, const gs_match_cfg*);
/* Submit one batch of reads: bases = the reads' sequence bytes back to back (ASCII, exactly as parsed:
 * no upper-casing), offsets[n_reads+1] byte offsets into bases, first_read_no =
 * global ordinal of read 0 (file order; used for the maxContigDescriptor tie-break).  Host buffers must stay
 * valid until the ticket is collected.  Batches are dealt round-robin to the devices of the context; up to
 * GS_MAX_INFLIGHT tickets may be pending per device (collect in submission order). */
#define GS_MAX_INFLIGHT 3
int gs_match_submit(gs_sess*, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads,
                    uint64_t first_read_no, gs_ticket* ticket);
/* ---- raw FASTQ text (the parallel feeder: record splitting on the GPU instead of AbstractFastqReader's single producer
 * thread, C/fastq/AbstractFastqReader.java:288-368 + B/io/BufferedLineReader.java:114-182).
 * text = n_bytes of FASTQ that consist of WHOLE records (the host cuts its input at a record boundary) and end with '\n'.
 * The device finds the line ends, takes lines 4i .. 4i+3 as record i and checks that the chunk is strict 4-line FASTQ
 * under the reference parser's rules (no NUL byte, third line starts with '+', quality at least as long as the sequence,
 * line count a multiple of 4).  If it is not, info->status holds GS_FASTQ_* bits, *ticket is 0, nothing is pending, and the
 * caller parses this chunk with the sequential parser (gs_match_submit) -- results never depend on the fast path.
 * Otherwise the batch runs exactly like gs_match_submit of the same reads.  n_bytes < 2^32 - 256. */
typedef struct gs_fastq_info {
    uint32_t n_reads;
    uint32_t status;       /* 0 = strict 4-line FASTQ, batch submitted */
    uint64_t total_kmers;  /* sum of max(0, L - k + 1): AbstractFastqReader.kMers (:346-349) */
    uint64_t total_bps;    /* sum of L: AbstractFastqReader.readBPs */
} gs_fastq_info;
#define GS_FASTQ_NUL 1u      /* a NUL byte (the reference drops them, BufferedLineReader.java:166-169) */
#define GS_FASTQ_LINES 2u    /* number of lines not a multiple of 4 */
#define GS_FASTQ_CAP 4u      /* lines shorter than 16 bytes on average */
#define GS_FASTQ_RECORD 8u   /* multi-line sequence / '+' line missing / quality shorter than the sequence */
#define GS_FASTQ_TAIL 16u    /* last line without '\n' */
/* Where record i sits in the text chunk: the header line starts at hdr_start (incl. '@'), the sequence at seq_start with
 * seq_len bytes, the quality line at qual_start and ends right before recs[i + 1].hdr_start - 1 (its '\n'). */
typedef struct gs_fastq_rec {
    uint32_t hdr_start, seq_start, seq_len, qual_start;
} gs_fastq_rec;
int gs_match_submit_fastq(gs_sess*, const uint8_t* text, uint64_t n_bytes, uint64_t first_read_no, gs_fastq_info* info,
                          gs_ticket* ticket);
/* Zero-copy views (valid like gs_match_collect_view's): out[n_reads], recs[n_reads + 1] (the last entry marks the end of
 * the text), events[n_events] with event_hdr_start[e] = header offset of the read that caused event e.  With want_runs:
 * run_offsets[n_reads + 1] / runs[runs_cap] as in gs_match_collect (runs_cap >= info.total_kmers always suffices); else NULL. */
int gs_match_collect_fastq(gs_sess*, gs_ticket, const gs_read_result** out, uint32_t* n_reads,
                           const gs_maxcontig_event** events, const uint32_t** event_hdr_start, uint32_t* n_events,
                           const gs_fastq_rec** recs, uint64_t* run_offsets, gs_run* runs, uint64_t runs_cap);
/* Wait for a ticket.  out[n_reads]; events[ev_cap] / n_events may be NULL.  If want_runs: run_offsets
 * [n_reads+1] and runs[runs_cap] receive the contig runs (GS_ERR_LIMIT if runs_cap is too small). */
int gs_match_collect(gs_sess*, gs_ticket, gs_read_result* out, gs_maxcontig_event* events, uint32_t ev_cap,
                     uint32_t* n_events, uint64_t* run_offsets, gs_run* runs, uint64_t runs_cap);
/* Zero-copy variant: *out / *events point into the session's pinned staging buffers (a JNI shim wraps them with
 * NewDirectByteBuffer); valid until GS_MAX_INFLIGHT further batches were submitted to the same device. */
int gs_match_collect_view(gs_sess*, gs_ticket, const gs_read_result** out, uint32_t* n_reads,
                          const gs_maxcontig_event** events, uint32_t* n_events);
/* End of run: merges all devices, runs the unique-k-mer count (KMerUniqueCounterBits.getUniqueKmerCounts,
 * C/store/KMerUniqueCounterBits.java:146-163) and returns per-value-index counts[n_values].
 * top_counts (may be NULL): (n_values+1) x max_kmer_res_counts Java shorts, row n_values = total
 * (getMaxCountsCounts :173-199). */
int gs_match_finish(gs_sess*, gs_taxon_counts* counts, int16_t* top_counts);
void gs_match_close(gs_sess*);

/* ---- several GPUs: the end-of-run merge ------------------------------------------------------------
 * Reads shard over GPUs with the database replicated; nothing is exchanged while matching.  At the end the per-GPU state has to
 * become what ONE FastqKMerMatcher would hold after all reads (the tail of runMatcher, C/match/FastqKMerMatcher.java:199-234;
 * KMerUniqueCounterBits.getUniqueKmerCounts, C/store/KMerUniqueCounterBits.java:146-163): counters by ncclAllReduce(sum),
 * max-contigs by ncclAllReduce(max) on (len << 40 | ~ordinal), unique-k-mer bitsets OR-merged slice-wise by a kernel that reads
 * the other GPUs' bitsets over NVLink (peer mappings; ncclSend/ncclRecv exchange where none can be made) and counts the merged
 * bits per value index, then ncclAllReduce(sum) of the counts.  NCCL is loaded at run time ("libnccl.so.2", or $GS_NCCL_LIB).
 *  - one process, several GPUs (gs_ctx_create with n devices): gs_match_finish does all of this by itself;
 *  - one process per GPU: rank 0 calls gs_comm_unique_id and hands the 128 bytes to the other processes by whatever channel the
 *    host has (the Java host: its own RPC; bench.py: torch.distributed broadcast); every process calls gs_comm_create and, at
 *    the end of the run, gs_match_finish_comm (collective: all ranks, same configuration).  Every rank receives the full result.
 * The database must have been built from the same store on every rank (the probe table's slot ids are a function of the key set). */
typedef struct gs_comm gs_comm;
#define GS_COMM_ID_BYTES 128
int gs_comm_unique_id(uint8_t* id /* [GS_COMM_ID_BYTES] */);
gs_comm* gs_comm_create(gs_ctx*, const uint8_t* id, int world, int rank); /* the context must hold exactly one device */
int gs_comm_world(const gs_comm*);
int gs_comm_rank(const gs_comm*);
void gs_comm_destroy(gs_comm*);
int gs_match_finish_comm(gs_sess*, gs_comm*, gs_taxon_counts* counts, int16_t* top_counts);
/* Optional, collective, any time before gs_match_finish_comm (best right after gs_match_open): exchanges the peer mappings of
 * the ranks' unique-k-mer bitsets and loads the merge kernel, so that the merge at the end of the run does not pay for them
 * (5 ms for a 512 MB bitset).  Every rank must call gs_match_finish_comm before any rank closes its session. */
int gs_match_prepare_merge(gs_sess*, gs_comm*);
/* The packer itself (host only, no device needed): n ASCII bases -> codes[ceil(n/32)] (2-bit codes C=0 G=1 A=2 T=3,
 * C/util/CGAT.java:66-69, 32 per word, first base in the top two bits) and valid[ceil(n/32)] (bit i = base 32w+i is one of the
 * upper-case letters CGAT, CGAT.java:60-69).  threads as in gs_match_cfg.host_pack_threads (1 = the calling thread alone). */
int gs_pack_bases(const uint8_t* bases, uint64_t n, uint64_t* codes, uint32_t* valid, int threads);
const char* gs_pack_isa(void); /* "avx512", "avx2" or "scalar": the body chosen on this CPU */
/* Host-side packing of this session so far (gs_match_cfg.host_pack_threads): threads of the pool (0 = none was needed), host
 * seconds spent packing inside gs_match_submit, bases packed, and the bytes of base data all submits put on the link. */
int gs_match_pack_stats(const gs_sess*, int* threads, double* pack_seconds, uint64_t* bases_packed, uint64_t* h2d_base_bytes);
double gs_match_pack_fraction(const gs_sess*); /* the share of a batch that is packed right now (host_pack_percent) */
/* Measurement of the last merge of this session: CUDA-event time of the whole merge and of its bitset part (ms, on the
 * session's compute stream), bitset bytes this rank read from the other ranks, path (1 = peer mappings, 2 = NCCL exchange). */
int gs_match_merge_stats(const gs_sess*, double* total_ms, double* bitset_ms, uint64_t* bytes_from_peers, int* path);
/* L2-persisting access window of this session's compute stream on its first device (north_star: "L2-persisting access
 * windows"): bytes of the window (the minimizer prefilter), bytes of the persisting carve-out, hit ratio; all 0 = no window
 * (classic layout, prefilter off, GS_L2_PERSIST=0, or refused by the device). */
int gs_match_l2_window(const gs_sess*, uint64_t* window_bytes, uint64_t* persisting_bytes, double* hit_ratio);

/* Device-resident variants (inputs already in HBM; used by bench.py's kernel-only number and by a host that
 * decodes on the GPU).  d_bases must be 16-byte aligned and readable 32 bytes past the last base; d_offsets[0] == 0 and
 * d_offsets[n_reads] == n_bases (checked on the device).  Runs on the session's device 0, asynchronously on the session's
 * stream; gs_match_sync waits. */
int gs_match_run_device(gs_sess*, const uint8_t* d_bases, const uint64_t* d_offsets, uint32_t n_reads,
                        uint64_t n_bases, uint64_t first_read_no, gs_read_result* d_out);
/* The kernels of a batch run on two streams of the session: the label kernel on the compute stream (gs_match_stream), the
 * reduce kernels on a second one, so that they overlap the next batch's label kernel.  gs_match_join orders the compute stream
 * behind everything submitted so far without a host wait -- call it before recording an event of your own on gs_match_stream
 * or before chaining your own device work there.  gs_match_sync / collect / finish need no join. */
int gs_match_join(gs_sess*);
int gs_match_sync(gs_sess*);
/* Raw device state (tests; before ABI 5 also the hook for a reduction outside the library -- see gs_match_finish_comm):
 * counters = int64[7][n_values] (kmers, contigs, sqsum, reads1, reads, readsKmers, readsBPs),
 * maxcontig = uint64[n_values] packed (len << 40 | ~ordinal), bitset = uint64[bitset_words] or NULL, one bit per
 * storage position of the session's layout (table slot id or sorted-array index). */
int gs_match_device_state(gs_sess*, int64_t** counters, uint64_t** maxcontig, uint64_t** bitset,
                          uint64_t* bitset_words);
/* Per-taxon popcount of bitset words [word_begin, word_end) into d_unique (int64[n_values], device, added to). */
int gs_match_unique_popcount(gs_sess*, const uint64_t* d_bitset, uint64_t word_begin, uint64_t word_end,
                             int64_t* d_unique);
/* CUDA stream (cudaStream_t) the device-resident calls are issued on; kernels launched since open. */
void* gs_match_stream(gs_sess*);
uint64_t gs_match_kernel_launches(const gs_sess*);
/* Measurement: with timing on, every batch records CUDA events on the compute stream around the label kernel (encode ->
 * prefilter -> store lookup -> unique bits) and around the reduce kernels (per-taxon counting, classification);
 * gs_match_kernel_times waits for the stream and returns the average duration per batch since timing was switched on. */
int gs_match_set_timing(gs_sess*, int on);
int gs_match_kernel_times(gs_sess*, double* label_ms, double* reduce_ms, uint64_t* n_batches);
/* Debug/parity: per-position labels of a device-resident batch: labels[i] = value index, -1 miss, -2 invalid;
 * pos[i] = storage position or -1; kmer_offsets[n_reads+1] = prefix sums of max(0, L-k+1). All device pointers. */
int gs_match_dump_labels(gs_sess*, const uint8_t* d_bases, const uint64_t* d_offsets, uint32_t n_reads,
                         const uint64_t* d_kmer_offsets, int32_t* d_labels, int64_t* d_pos);

/* ---- filter -------------------------------------------------------------------------------------
 * Replaces: LoadIndexGoal's KMerProbFilter (C/goals/LoadIndexGoal.java:92-104) and FastqBloomFilter
 * (C/bloom/FastqBloomFilter.java:62-161) built in FilterGoal.makeFile (C/goals/FilterGoal.java:80-108). */
#define GS_BLOOM_BLOCKED 0
#define GS_BLOOM_XOR 1
#define GS_BLOOM_MURMUR 2
/* blocked: p0 = seed, p1 = buckets, factors = NULL;  xor/murmur: p0 = bits (the modulus), p1 = hashes,
 * factors[hashes] = hashFactors (C/bloom/AbstractKMerBloomFilter.java:104-110). */
gs_filter* gs_filter_create(gs_ctx*, int kind, int64_t p0, int64_t p1, const int64_t* factors,
                            const int64_t* words, uint64_t n_words);
/* Flat little-endian filter index file "GSF1" (layout in gs_capi.cu): what KMerProbFilter.save / KMerProbFilter.load move as
 * a Java object stream (`*_index.ser.gz`, C/goals/LoadIndexGoal.java:92-104, C/goals/refseq/BloomIndexGoal.java:66-111), as plain
 * arrays: kind, bits / hashes (or seed / buckets), hash factors, words.  Written by gs_filter_save_file or by the Java-side
 * exporter (integration/java/.../bloom/GsfExporter.java); gs_filter_load_file = read + gs_filter_create, no JVM needed. */
int gs_filter_save_file(gs_filter*, const char* path);
gs_filter* gs_filter_load_file(gs_ctx*, const char* path);
void gs_filter_destroy(gs_filter*);
int gs_filter_n_devices(const gs_filter*);
/* KMerProbFilter.containsLong (C/bloom/KMerProbFilter.java:66) for tests. */
int gs_filter_contains(gs_filter*, const int64_t* kmers, uint64_t n, uint8_t* out);
gs_fsess* gs_filter_open(gs_filter*, int k, int min_pos_count, double pos_ratio);
int gs_filter_submit(gs_fsess*, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, gs_ticket*);
/* accept[n_reads]: 1 = isAcceptRead (C/bloom/FastqBloomFilter.java:120-161). */
int gs_filter_collect(gs_fsess*, gs_ticket, uint8_t* accept);
/* Raw FASTQ text (see gs_match_submit_fastq): accept[n_reads] and recs[n_reads + 1] are views of the session's pinned
 * staging, valid until GS_MAX_INFLIGHT further batches were submitted to the same device. */
int gs_filter_submit_fastq(gs_fsess*, const uint8_t* text, uint64_t n_bytes, gs_fastq_info* info, gs_ticket* ticket);
int gs_filter_collect_fastq(gs_fsess*, gs_ticket, const uint8_t** accept, uint32_t* n_reads, const gs_fastq_rec** recs);
/* Device-resident variant (inputs already in HBM): d_bases 16-byte aligned and readable 32 bytes past the last base,
 * d_offsets[0] == 0 and d_offsets[n_reads] == n_bases; asynchronous on the session's compute stream (gs_filter_stream). */
int gs_filter_run_device(gs_fsess*, const uint8_t* d_bases, const uint64_t* d_offsets, uint32_t n_reads, uint64_t n_bases,
                         uint8_t* d_accept);
uint64_t gs_filter_kernel_launches(const gs_fsess*); /* kernels this session has launched so far */
int gs_filter_sync(gs_fsess*);
void* gs_filter_stream(gs_fsess*);
void gs_filter_close(gs_fsess*);

/* ---- block-gzip input ---------------------------------------------------------------------------
 * Replaces: java.util.zip.GZIPInputStream in front of the FASTQ reader (B/io/StreamProvider via
 * C/fastq/AbstractFastqReader.java:224) for block-gzip (BGZF) files -- multi-member gzip files whose members are
 * independent raw-deflate streams of at most 64 KB.  The caller finds the members from their headers (the 'BC' extra
 * subfield holds the member size) and describes them here; the device inflates them, one thread per block, and checks
 * each against its trailer like gzread / GZIPInputStream do (exactly out_len bytes, CRC-32).
 * comp[comp_bytes] and out[out_bytes] are host buffers (pinned ones copy faster); blocks[i] = deflate data at
 * comp + in_off (in_len bytes, gzip header and trailer excluded) -> out + out_off (out_len = ISIZE bytes, crc32 = the
 * trailer's CRC-32).  Returns GS_ERR_DATA if a block is corrupt (blocks[i].status != 0 marks it); nothing of `out` is
 * valid then.  Calls on one context are serialized. */
typedef struct gs_deflate_block {
    uint64_t in_off, out_off;
    uint32_t in_len, out_len, crc32, status;
} gs_deflate_block;
int gs_inflate_blocks(gs_ctx*, const uint8_t* comp, uint64_t comp_bytes, gs_deflate_block* blocks, uint32_t n_blocks,
                      uint8_t* out, uint64_t out_bytes);
/* the context a database / filter index lives on (for callers that only hold the object) */
gs_ctx* gs_db_context(gs_db*);
gs_ctx* gs_filter_context(gs_filter*);

#ifdef __cplusplus
}
#endif
#endif /* GENESTRIP_B200_H */
