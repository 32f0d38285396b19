// GpuFastqKMerMatcher.java -- the reference-side subclass a Genestrip maintainer adds (NOT compiled here: no JDK in this
// image).  It lives in package org.metagene.genestrip.match so that it can fill the protected fields of CountsPerTaxid and
// use FastqKMerMatcher's package-private state, and it is installed through the reference's own override point
//   MatchResultGoal.createMatcher(...)            (core/.../goals/MatchResultGoal.java:174-197)
// by a GSMaker subclass that overrides createGoalChainForMatchResult (core/.../GSMaker.java:560-583), exactly how the
// reference's own ComprehensiveFilterTest swaps goals (core/src/test/.../goals/refseq/ComprehensiveFilterTest.java:91-152).
package org.metagene.genestrip.match;

import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;

import org.metagene.genestrip.ExecutionContext;
import org.metagene.genestrip.store.KMerSortedArray;
import org.metagene.genestrip.store.KMerStore;
import org.metagene.genestrip.tax.SmallTaxTree;
import org.metagene.genestrip.tax.SmallTaxTree.SmallTaxIdNode;

final class GsNative {
    static { System.loadLibrary("gsjni"); }
    static native long ctxCreate(int[] devices);
    static native void ctxDestroy(long ctx);
    static native long dbCreate(long ctx, int k, long nKmers, int nValues);
    static native void dbPutKeys(long db, long offset, long[] keys, int n);
    static native void dbPutValues(long db, long offset, short[] vals, int n);
    static native void dbPutRadixBucket(long db, int radixBits, int radix, long[] entries, int n);
    static native void dbSetTree(long db, int[] parentByValueIndex, int[] hasNode);
    static native void dbSetBloomBlocked(long db, long seed, long buckets, long[] words);
    static native void dbFinalize(long db);
    static native void dbDestroy(long db);
    static native long matchOpen(long db, boolean classify, boolean countUnique, int maxKmerResCounts, boolean useBloom, int maxPaths,
                                 int minKmersForClass, double maxTaxErr, double maxClassErr, boolean wantRuns);
    static native ByteBuffer allocPinned(long bytes);
    static native void freePinned(ByteBuffer b);
    static native long matchSubmit(long sess, ByteBuffer bases, ByteBuffer offsets, int nReads, long firstReadNo);
    static native ByteBuffer[] matchCollect(long sess, long ticket);
    static native void matchFinish(long sess, ByteBuffer counts, ByteBuffer topCounts);
    static native void matchClose(long sess);
    // raw FASTQ text chunks (records split on the GPU); {ticket, nReads, status, totalKmers, totalBps}
    static native long[] matchSubmitFastq(long sess, ByteBuffer text, long nBytes, long firstReadNo);
    static native ByteBuffer[] matchCollectFastq(long sess, long ticket);   // results, events, event header offsets, record table
    static native long[] filterSubmitFastq(long fsess, ByteBuffer text, long nBytes);
    static native ByteBuffer[] filterCollectFastq(long fsess, long ticket);
    // db goal, update phase (DBGoal.MyFastaReader): value = LCA(value, region node) for the stored k-mers of the regions
    static native long dbUpdate(long db, ByteBuffer seq, long nBytes, long[] regionOffsets, int[] regionValueIndex, boolean upperCase);
    static native void dbGetValues(long db, long offset, short[] vals, int n);
    // block-gzip input: blocks = nBlocks x {long inOff, long outOff, int inLen, int outLen, int crc32, int status} (little endian),
    // filled from the members' headers ('BC' subfield) and trailers; throws when a member is corrupt, like GZIPInputStream
    static native void inflateBlocks(long ctx, ByteBuffer comp, long compBytes, ByteBuffer blocks, int nBlocks, ByteBuffer out, long outBytes);
}

/**
 * Runs with zero Java consumer threads: nextEntry(...) (AbstractFastqReader.java:488) appends the parsed read to the current
 * pinned batch instead of matching it; full batches are submitted to the GPU and collected in order; afterMatch(...) and
 * the four double sums are then replayed on the host per read, in read order, from the 16-byte result records.
 */
public class GpuFastqKMerMatcher extends FastqKMerMatcher {
    private static final int BATCH_READS = 1 << 20, RESULT_BYTES = 16, COUNTS_BYTES = 80;
    private final long db;
    private long sess;
    private final ByteBuffer[] bases = new ByteBuffer[3], offsets = new ByteBuffer[3];
    private final long[] tickets = new long[3];
    private int cur, nInBatch, inflight;
    private long ordinal;

    public GpuFastqKMerMatcher(long db, KMerStore<SmallTaxIdNode> store, int initialReadSize, int maxQueueSize, ExecutionContext bundle,
            boolean withProbs, int maxKmerResCounts, SmallTaxTree taxTree, int maxPaths, double maxReadTaxErrorCount,
            double maxReadClassErrorCount, boolean writeAll, int threshold, String dbMD5) {
        super(store, initialReadSize, maxQueueSize, /* bundle with threads = 0 */ bundle, withProbs, maxKmerResCounts, taxTree, maxPaths,
                maxReadTaxErrorCount, maxReadClassErrorCount, writeAll, threshold, dbMD5);
        this.db = db;
        for (int i = 0; i < 3; i++) {
            bases[i] = GsNative.allocPinned(256L << 20).order(ByteOrder.LITTLE_ENDIAN);
            offsets[i] = GsNative.allocPinned(8L * (BATCH_READS + 1)).order(ByteOrder.LITTLE_ENDIAN);
        }
    }

    /** Upload of a KMerSortedArray-backed database: segments are pinned only while they are copied. */
    public static long upload(long ctx, KMerSortedArray<String> store, SmallTaxTree tree /* + Bloom filter words */) {
        long db = GsNative.dbCreate(ctx, store.getK(), store.getEntries(), store.getNValues());
        // store.visitSegments((off, long[] keys, short[] vals, n) -> { dbPutKeys(db, off, keys, n); dbPutValues(db, off, vals, n); });
        // dbSetTree(db, parent[storeIndex], hasNode[storeIndex]); dbSetBloomBlocked(db, seed, buckets, data); dbFinalize(db);
        return db;
    }

    @Override
    protected void nextEntry(ReadEntry entry, int threadIndex) throws IOException {
        ByteBuffer b = bases[cur], o = offsets[cur];
        if (nInBatch == BATCH_READS || b.remaining() < entry.readSize) flush();
        if (nInBatch == 0) { o.clear(); o.putLong(0); }
        bases[cur].put(entry.read, 0, entry.readSize);
        offsets[cur].putLong(bases[cur].position());
        // descriptor (and quality) are kept per batch on the Java side for afterMatch / maxContigDescriptor
        nInBatch++;
        ordinal++;
    }

    private void flush() {
        if (nInBatch == 0) return;
        tickets[cur] = GsNative.matchSubmit(sess, bases[cur], offsets[cur], nInBatch, ordinal - nInBatch);
        if (++inflight == 2) collectOldest();
        cur = (cur + 1) % 3;
        bases[cur].clear();
        nInBatch = 0;
    }

    private void collectOldest() {
        ByteBuffer[] r = GsNative.matchCollect(sess, tickets[(cur + 3 - (inflight - 1)) % 3]);
        ByteBuffer res = r[0].order(ByteOrder.LITTLE_ENDIAN);
        for (int i = 0; res.remaining() >= RESULT_BYTES; i++) {
            int classVidx = res.getInt(), readKmers = res.getInt(), taxErr = res.getInt(), flags = res.getInt();
            // found = (flags & 1) != 0 -> rewriteInput(entry, indexed); kraken line from the run list when enabled;
            // accepted = (flags & 2) != 0 -> stats = getCountsPerTaxid(node, classVidx); the four double sums exactly as in
            // FastqKMerMatcher.matchRead :511-526: err = ((double) taxErr) / max, classErr = ((double) (max - readKmers)) / max.
        }
        // r[1]: (vidx, contigLen, readNo) events -> copy the descriptor of read readNo into stats.maxContigDescriptor (:402-409)
        inflight--;
    }

    // runMatcher (FastqKMerMatcher.java:181-235): sess = matchOpen(...); processFastqStreams(fastqs); flush(); drain;
    // matchFinish(sess, counts, top) -> for every touched value index fill statsIndex[vi].{kmers, contigs, contigLenSquaredSum,
    // maxContigLen, reads1KMer, reads, readsKmers, readsBPs, uniqueKmers, maxKMerCounts}; return new MatchingResult(...).
    // dump() (called from MatchResultGoal.doMakeThis finally, :156-160) -> matchClose(sess), freePinned(...).
}
