// GpuFastqKMerMatcher.java -- the reference-side subclass a Genestrip maintainer adds (integration/build.sh compiles it against
// the reference's jar when a JDK is present; this image has none).  It lives in package org.metagene.genestrip.match so that it
// can fill the protected fields of CountsPerTaxid, and it is installed through the reference's own override point
//   MatchResultGoal.createMatcher(...)            (core/.../goals/MatchResultGoal.java:174-197)
// by GpuGSMaker (integration/java/org/metagene/genestrip/GpuGSMaker.java).
//
// FastqKMerMatcher.nextEntry is final (core/.../match/FastqKMerMatcher.java:277-296) but the two methods it calls are not:
//   matchRead(entry, index)  -> here it only RECORDS the read in the current pinned batch and answers false;
//   afterMatch(entry, found) -> suppressed while the parser runs.
// Full batches go to the GPU (gs_match_submit); when a batch comes back, every read is replayed in read order through the
// reference's own afterMatch (filtered FASTQ, kraken-style line) with the device's per-read result, and the four double sums of
// the classified-read statistics are accumulated exactly as matchRead does (:511-526).  The integer statistics come from the
// device at the end of the run (gs_match_finish).  The parser must run inline (threads = 0), which is also what makes the
// reference's own output order deterministic.
package org.metagene.genestrip.match;

import java.io.File;
import java.io.IOException;
import java.io.InputStream;
import java.io.OutputStream;
import java.io.PrintStream;
import java.io.UncheckedIOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.charset.StandardCharsets;
import java.util.Arrays;
import java.util.HashMap;
import java.util.Map;

import org.metagene.genestrip.ExecutionContext;
import org.metagene.genestrip.gpu.GsNative;
import org.metagene.genestrip.io.StreamProvider;
import org.metagene.genestrip.io.StreamingResourceStream;
import org.metagene.genestrip.store.KMerStore;
import org.metagene.genestrip.store.KMerUniqueCounterBits;
import org.metagene.genestrip.tax.SmallTaxTree;
import org.metagene.genestrip.tax.SmallTaxTree.SmallTaxIdNode;

public class GpuFastqKMerMatcher extends FastqKMerMatcher {
    private static final int SLOTS = 3, RESULT_BYTES = 16, EVENT_BYTES = 16, COUNTS_BYTES = 80;
    private static final int READ_FOUND = 1, READ_ACCEPTED = 2;
    private static final int RUN_MISS = 0xFFFFFFFE, RUN_INVALID = 0xFFFFFFFD;

    /** One pinned batch: the reads' bases back to back plus what the host needs again when the results arrive. */
    private static final class Batch {
        ByteBuffer bases, offsets;         // pinned (gs_alloc_pinned), little endian
        byte[] text = new byte[1 << 20];   // descriptors and, with withProbs, qualities of the batch's reads, back to back
        int textUsed;
        int[] descStart, descSize, probsStart, probsSize;
        long[] readNo;
        int n;
        long ticket, firstOrdinal, kmerRuns;
    }

    private final long db;
    private final boolean useBloom, withProbs;
    private final int batchReads;
    private final SmallTaxIdNode[] nodeByIndex;
    private final Batch[] batches = new Batch[SLOTS];
    private final MatcherReadEntry replay;
    private long sess;
    private int cur, inflight;
    private long ordinal;        // read ordinal over all files of the key: the tie-break of maxContigDescriptor across files
    private boolean parsing;
    private AfterMatchCallback callback;

    /**
     * @param db the device database made by {@link #upload(long, KMerStore)} (owned by the caller, shared by all keys)
     */
    public GpuFastqKMerMatcher(long db, boolean useBloom, int batchReads, long batchBytes, KMerStore<SmallTaxIdNode> store, int initialReadSize,
            int maxQueueSize, ExecutionContext bundle, boolean withProbs, int maxKmerResCounts, SmallTaxTree taxTree, int maxPaths,
            double maxReadTaxErrorCount, double maxReadClassErrorCount, boolean writeAll, int threshold, String dbMD5) {
        super(store, initialReadSize, maxQueueSize, bundle, withProbs, maxKmerResCounts, taxTree, maxPaths, maxReadTaxErrorCount,
                maxReadClassErrorCount, writeAll, threshold, dbMD5);
        if (bundle.getThreads() > 0) {
            throw new IllegalArgumentException("the GPU matcher replays reads in input order: run it with threads = 0");
        }
        this.db = db;
        this.useBloom = useBloom;
        this.withProbs = withProbs;
        this.batchReads = batchReads;
        nodeByIndex = new SmallTaxIdNode[store.getNValues()];
        for (int v = 0; v < nodeByIndex.length; v++) {
            nodeByIndex[v] = store.getValueForIndex(v);
        }
        for (int i = 0; i < SLOTS; i++) {
            Batch b = batches[i] = new Batch();
            b.bases = GsNative.allocPinned(batchBytes).order(ByteOrder.LITTLE_ENDIAN);
            b.offsets = GsNative.allocPinned(8L * (batchReads + 1)).order(ByteOrder.LITTLE_ENDIAN);
            b.descStart = new int[batchReads];
            b.descSize = new int[batchReads];
            b.probsStart = new int[batchReads];
            b.probsSize = new int[batchReads];
            b.readNo = new long[batchReads];
        }
        replay = new MatcherReadEntry(initialReadSize, withProbs, maxPaths);
    }

    /**
     * Upload of the converted store (Database.convertKMerStore, core/.../store/Database.java:136-143): keys and value indices
     * in storage order through the store's own visitor (KMerSortedArray.visit, core/.../store/KMerSortedArray.java:426-441 --
     * positions ascend, so segments can be streamed), the tax tree flattened by value index, the store's filter built on the
     * device.  For a RadixKMerStore use GsNative.dbPutRadixBucket per bucket instead of the visitor.
     */
    public static long upload(long ctx, KMerStore<SmallTaxIdNode> store) {
        final int nValues = store.getNValues();
        final long db = GsNative.dbCreate(ctx, store.getK(), store.getEntries(), nValues);
        final int seg = 1 << 22;
        final long[] keys = new long[seg];
        final short[] vals = new short[seg];
        final long[] base = { 0 };
        final int[] fill = { 0 };
        store.visit(new KMerStore.IndexedKMerStoreVisitor<SmallTaxIdNode>() {
            @Override
            public void nextValue(KMerStore<SmallTaxIdNode> s, long kmer, int index, long pos) {
                keys[fill[0]] = kmer;
                vals[fill[0]] = (short) (index + Short.MIN_VALUE);
                if (++fill[0] == seg) {
                    GsNative.dbPutKeys(db, base[0], keys, seg);
                    GsNative.dbPutValues(db, base[0], vals, seg);
                    base[0] += seg;
                    fill[0] = 0;
                }
            }
        });
        if (fill[0] > 0) {
            GsNative.dbPutKeys(db, base[0], keys, fill[0]);
            GsNative.dbPutValues(db, base[0], vals, fill[0]);
        }
        int[] parent = new int[nValues], hasNode = new int[nValues];
        for (int v = 0; v < nValues; v++) {
            SmallTaxIdNode node = store.getValueForIndex(v);
            hasNode[v] = node != null ? 1 : 0;
            parent[v] = node != null && node.getParent() != null ? node.getParent().getStoreIndex() : -1;
        }
        GsNative.dbSetTree(db, parent, hasNode);
        GsNative.dbBuildBloom(db);
        GsNative.dbFinalize(db);
        return db;
    }

    @Override
    public void setAfterMatchCallback(AfterMatchCallback afterMatchCallback) {
        callback = afterMatchCallback;   // called from the replay, with the real result (super would call it at parse time)
    }

    // ---- parse time: record the read, decide nothing
    @Override
    protected boolean matchRead(final MatcherReadEntry entry, final int index) {
        if (!parsing) {
            return super.matchRead(entry, index);
        }
        try {
            Batch b = batches[cur];
            if (b.n == batchReads || b.bases.remaining() < entry.readSize) {
                flush();
                b = batches[cur];
            }
            if (entry.readSize > b.bases.capacity()) {
                throw new IllegalStateException("read longer than a batch: " + entry.readSize + " bases");
            }
            if (b.n == 0) {
                b.bases.clear();
                b.offsets.clear();
                b.offsets.putLong(0);
                b.textUsed = 0;
                b.firstOrdinal = ordinal;
                b.kmerRuns = 0;
            }
            b.bases.put(entry.read, 0, entry.readSize);
            b.offsets.putLong(b.bases.position());
            int need = b.textUsed + entry.readDescriptorSize + Math.max(0, entry.readProbsSize);
            if (need > b.text.length) {
                b.text = Arrays.copyOf(b.text, Math.max(need, 2 * b.text.length));
            }
            b.descStart[b.n] = b.textUsed;
            b.descSize[b.n] = entry.readDescriptorSize;
            System.arraycopy(entry.readDescriptor, 0, b.text, b.textUsed, entry.readDescriptorSize);
            b.textUsed += entry.readDescriptorSize;
            b.probsStart[b.n] = b.textUsed;
            b.probsSize[b.n] = withProbs ? entry.readProbsSize : -1;
            if (withProbs && entry.readProbsSize > 0) {
                System.arraycopy(entry.readProbs, 0, b.text, b.textUsed, entry.readProbsSize);
                b.textUsed += entry.readProbsSize;
            }
            b.readNo[b.n] = entry.readNo;
            b.kmerRuns += Math.max(0, entry.readSize - k + 1);
            b.n++;
            ordinal++;
        } catch (IOException e) {
            throw new UncheckedIOException(e);
        }
        return false;
    }

    @Override
    protected void afterMatch(MatcherReadEntry myEntry, boolean found) throws IOException {
        if (parsing && myEntry != replay) {
            return;   // the read's result does not exist yet: afterMatch runs from the replay
        }
        super.afterMatch(myEntry, found);
    }

    @Override
    protected void readFastq(InputStream inputStream, boolean fasta) throws IOException {
        parsing = true;
        try {
            super.readFastq(inputStream, fasta);   // readNo restarts at 0 for every file (AbstractFastqReader.java:226)
            flush();
            while (inflight > 0) {
                collectOldest();
            }
        } finally {
            parsing = false;
        }
    }

    private void flush() throws IOException {
        Batch b = batches[cur];
        if (b.n == 0) {
            return;
        }
        b.ticket = GsNative.matchSubmit(sess, b.bases, b.offsets, b.n, b.firstOrdinal);
        inflight++;
        if (inflight == SLOTS) {
            collectOldest();
        }
        cur = (cur + 1) % SLOTS;
        batches[cur].n = 0;
    }

    // ---- a batch is back: replay its reads in read order
    private void collectOldest() throws IOException {
        Batch b = batches[(cur + SLOTS - (inflight - 1)) % SLOTS];
        ByteBuffer[] r = out != null ? GsNative.matchCollectRuns(sess, b.ticket, b.n, b.kmerRuns) : GsNative.matchCollect(sess, b.ticket);
        ByteBuffer res = r[0].order(ByteOrder.LITTLE_ENDIAN), ev = r[1].order(ByteOrder.LITTLE_ENDIAN);
        ByteBuffer runOff = out != null ? r[2].order(ByteOrder.LITTLE_ENDIAN) : null, runs = out != null ? r[3].order(ByteOrder.LITTLE_ENDIAN) : null;
        // new per-taxon maximum contig lengths set by reads of this batch: the descriptor of that read (:402-409)
        while (ev.remaining() >= EVENT_BYTES) {
            int vidx = ev.getInt(), contigLen = ev.getInt();
            int i = (int) (ev.getLong() - b.firstOrdinal);
            CountsPerTaxid stats = getCountsPerTaxid(nodeByIndex[vidx], vidx);
            stats.maxContigLen = Math.max(stats.maxContigLen, contigLen);
            int j = 1;
            for (; j < b.descSize[i] && j < stats.maxContigDescriptor.length && b.text[b.descStart[i] + j] != ' '; j++) {
                stats.maxContigDescriptor[j - 1] = b.text[b.descStart[i] + j];
            }
            stats.maxContigDescriptor[j - 1] = 0;
        }
        for (int i = 0; i < b.n; i++) {
            int classVidx = res.getInt(), readKmers = res.getInt(), taxErr = res.getInt(), flags = res.getInt();
            int start = (int) b.offsets.getLong(8 * i), size = (int) b.offsets.getLong(8 * (i + 1)) - start;
            if (replay.read.length < size) {
                replay.read = new byte[Math.max(size, 2 * replay.read.length)];
                if (withProbs) {
                    replay.readProbs = new byte[replay.read.length];
                }
            }
            for (int p = 0; p < size; p++) {
                replay.read[p] = b.bases.get(start + p);
            }
            replay.readSize = size;
            replay.readNo = b.readNo[i];
            replay.readDescriptorSize = b.descSize[i];
            System.arraycopy(b.text, b.descStart[i], replay.readDescriptor, 0, b.descSize[i]);
            replay.readProbsSize = b.probsSize[i];
            if (withProbs && b.probsSize[i] > 0) {
                if (replay.readProbs.length < b.probsSize[i]) {
                    replay.readProbs = new byte[b.probsSize[i]];
                }
                System.arraycopy(b.text, b.probsStart[i], replay.readProbs, 0, b.probsSize[i]);
            }
            replay.classNode = classVidx >= 0 ? nodeByIndex[classVidx] : null;
            replay.bufferPos = 0;
            if (out != null) {   // the kraken-style segments of the read, printed by the reference's own routine (:597-611)
                int r0 = (int) runOff.getLong(8 * i), r1 = (int) runOff.getLong(8 * (i + 1));
                for (int q = r0; q < r1; q++) {
                    int label = runs.getInt(8 * q), len = runs.getInt(8 * q + 4);
                    SmallTaxIdNode node = label == RUN_INVALID ? INVALID_NODE : label == RUN_MISS ? null : nodeByIndex[label];
                    printKrakenStyleOut(replay, node, len, q - r0);
                }
            }
            boolean found = (flags & READ_FOUND) != 0;
            if ((flags & READ_ACCEPTED) != 0) {   // the classified-read statistics (:509-526); the integer sums come from the device
                int max = size - k + 1;
                double err = ((double) taxErr) / max;
                double classErr = ((double) (max - readKmers)) / max;
                CountsPerTaxid stats = getCountsPerTaxid(replay.classNode, classVidx);
                stats.errorSum += err;
                stats.errorSquaredSum += err * err;
                stats.classErrorSum += classErr;
                stats.classErrorSquaredSum += classErr * classErr;
            }
            afterMatch(replay, found);
            if (callback != null) {
                callback.afterMatch(replay, found);
            }
        }
        inflight--;
    }

    @Override
    public MatchingResult runMatcher(StreamingResourceStream fastqs, File filteredFile, File krakenOutStyleFile,
            KMerUniqueCounterBits uniqueCounter) throws IOException {
        final int nValues = statsIndex.length;
        final boolean withCounts = uniqueCounter != null && uniqueCounter.isWithCounts() && maxKmerResCounts > 0;
        ByteBuffer counts = ByteBuffer.allocateDirect(Math.max(1, nValues) * COUNTS_BYTES).order(ByteOrder.LITTLE_ENDIAN);
        ByteBuffer top = withCounts ? ByteBuffer.allocateDirect((nValues + 1) * maxKmerResCounts * 2).order(ByteOrder.LITTLE_ENDIAN) : null;
        try (OutputStream lindexed = filteredFile != null ? StreamProvider.getOutputStreamForFile(filteredFile) : null;
                PrintStream lout = krakenOutStyleFile != null
                        ? new PrintStream(StreamProvider.getOutputStreamForFile(krakenOutStyleFile), false, StandardCharsets.UTF_8)
                        : null) {
            indexed = lindexed;
            out = lout;
            initStats();
            sess = GsNative.matchOpen(db, taxTree != null, uniqueCounter != null, withCounts ? maxKmerResCounts : 0, useBloom, maxPaths, threshold,
                    maxReadTaxErrorCount, maxReadClassErrorCount, lout != null);
            ordinal = 0;
            cur = 0;
            inflight = 0;
            batches[0].n = 0;
            try {
                processFastqStreams(fastqs);   // readFastq above drains the batches at the end of every file
                GsNative.matchFinish(sess, counts, top);
            } finally {
                GsNative.matchClose(sess);
                sess = 0;
            }
        }
        out = null;
        indexed = null;

        // the integer fields of CountsPerTaxid (core/.../match/CountsPerTaxid.java:127-151), as one matcher would hold them
        Map<String, CountsPerTaxid> taxid2Stats = new HashMap<>();
        Map<String, short[]> countMap = withCounts ? new HashMap<String, short[]>() : null;
        for (int v = 0; v < nValues; v++) {
            int o = v * COUNTS_BYTES;
            if (counts.getInt(o + 68) == 0 && statsIndex[v] == null) {
                continue;   // statsIndex[v] stays null: no row
            }
            CountsPerTaxid stats = getCountsPerTaxid(nodeByIndex[v], v);
            stats.kmers = counts.getLong(o);
            stats.contigs = (int) counts.getLong(o + 8);
            stats.contigLenSquaredSum = counts.getLong(o + 16);
            stats.reads1KMer = counts.getLong(o + 24);
            stats.reads = counts.getLong(o + 32);
            stats.readsKmers = counts.getLong(o + 40);
            stats.readsBPs = counts.getLong(o + 48);
            stats.uniqueKmers = uniqueCounter != null ? counts.getLong(o + 56) : -1;
            stats.maxContigLen = counts.getInt(o + 64);
            if (withCounts) {
                short[] row = new short[maxKmerResCounts];
                for (int i = 0; i < maxKmerResCounts; i++) {
                    row[i] = top.getShort(2 * (v * maxKmerResCounts + i));
                }
                stats.maxKMerCounts = row;
                countMap.put(stats.getTaxid(), row);
            }
            taxid2Stats.put(stats.getTaxid(), stats);
        }
        short[] totalMaxCounts = null;
        if (withCounts) {
            totalMaxCounts = new short[maxKmerResCounts];
            for (int i = 0; i < maxKmerResCounts; i++) {
                totalMaxCounts[i] = top.getShort(2 * (nValues * maxKmerResCounts + i));
            }
        }
        return new MatchingResult(kmerStore.getK(), taxid2Stats, dbMD5, totalReads, totalKMers, totalBPs, totalMaxCounts);
    }

    @Override
    public void dump() {
        for (Batch b : batches) {
            if (b != null && b.bases != null) {
                GsNative.freePinned(b.bases);
                GsNative.freePinned(b.offsets);
                b.bases = null;
            }
        }
        super.dump();
    }
}
