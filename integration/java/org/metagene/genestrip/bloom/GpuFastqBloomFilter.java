// GpuFastqBloomFilter.java -- the `filter` goal's per-read decision on the GPU (integration/build.sh compiles it against the
// reference's jar when a JDK is present).  FastqBloomFilter.nextEntry / isAcceptRead (core/.../bloom/FastqBloomFilter.java:92-161)
// are not final, so this subclass overrides nextEntry to RECORD the read in a pinned batch; full batches go through
// gs_filter_submit, and when a batch comes back its reads are rewritten in read order to the accepted / rejected stream exactly
// as the reference's nextEntry does.  The output streams of FastqBloomFilter are private, so runFilter is restated here around
// this class's own.  Lives in package org.metagene.genestrip.bloom for the protected fields of AbstractKMerBloomFilter
// (bits, hashes, hashFactors, bitVector: core/.../bloom/AbstractKMerBloomFilter.java:47-64), which are the index to upload.
package org.metagene.genestrip.bloom;

import java.io.File;
import java.io.IOException;
import java.io.InputStream;
import java.io.OutputStream;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.Arrays;

import org.metagene.genestrip.ExecutionContext;
import org.metagene.genestrip.gpu.GsNative;
import org.metagene.genestrip.io.StreamProvider;
import org.metagene.genestrip.io.StreamingResourceStream;

public class GpuFastqBloomFilter extends FastqBloomFilter {
    private static final int SLOTS = 3, KIND_XOR = 1, KIND_MURMUR = 2;

    private static final class Batch {
        ByteBuffer bases, offsets, accept;
        byte[] text = new byte[1 << 20];
        int textUsed, n;
        int[] descStart, descSize, probsStart, probsSize;
        long ticket;
    }

    private final long filter;
    private final int kSize, minPosCount, batchReads;
    private final double positiveRatio;
    private final boolean withProbs;
    private final Batch[] batches = new Batch[SLOTS];
    private final MyReadEntry replay;
    private long fsess;
    private int cur, inflight;
    private OutputStream accepted, rejected;

    /**
     * The index as LoadIndexGoal deserializes it (core/.../goals/LoadIndexGoal.java:92-104) -> device.  The serialized bit count,
     * hash factors and bit vector are authoritative (the factors are only regenerated when the vector grows, :104-110).
     */
    public static long upload(long ctx, KMerProbFilter index) {
        if (!(index instanceof AbstractKMerBloomFilter)) {
            throw new IllegalArgumentException("the filter goal's index is an XOR or Murmur Bloom filter, got " + index.getClass());
        }
        AbstractKMerBloomFilter f = (AbstractKMerBloomFilter) index;
        if (f.bitVector.isLarge()) {
            throw new IllegalArgumentException("bit vectors beyond 2^31 words: upload the segments of largeBits one after the other");
        }
        int words = (int) f.bitVector.size;
        return GsNative.filterCreate(ctx, f instanceof XORKMerBloomFilter ? KIND_XOR : KIND_MURMUR, f.bits, f.hashes, f.hashFactors,
                words == f.bitVector.bits.length ? f.bitVector.bits : Arrays.copyOf(f.bitVector.bits, words));
    }

    public GpuFastqBloomFilter(long filter, int batchReads, long batchBytes, int k, KMerProbFilter index, int minPosCount, double positiveRatio,
            int initialReadSize, int maxQueueSize, ExecutionContext bundle, boolean withProbs) {
        super(k, index, minPosCount, positiveRatio, initialReadSize, maxQueueSize, bundle, withProbs);
        if (bundle.getThreads() > 0) {
            throw new IllegalArgumentException("the GPU filter writes reads in input order: run it with threads = 0");
        }
        this.filter = filter;
        this.kSize = k;
        this.minPosCount = minPosCount;
        this.positiveRatio = positiveRatio;
        this.withProbs = withProbs;
        this.batchReads = batchReads;
        for (int i = 0; i < SLOTS; i++) {
            Batch b = batches[i] = new Batch();
            b.bases = GsNative.allocPinned(batchBytes).order(ByteOrder.LITTLE_ENDIAN);
            b.offsets = GsNative.allocPinned(8L * (batchReads + 1)).order(ByteOrder.LITTLE_ENDIAN);
            b.accept = ByteBuffer.allocateDirect(batchReads);
            b.descStart = new int[batchReads];
            b.descSize = new int[batchReads];
            b.probsStart = new int[batchReads];
            b.probsSize = new int[batchReads];
        }
        replay = new MyReadEntry(initialReadSize, withProbs);
    }

    @Override
    public void runFilter(StreamingResourceStream fastqs, File filteredFile, File restFile) throws IOException {
        try (OutputStream lindexed = filteredFile != null ? StreamProvider.getOutputStreamForFile(filteredFile) : null;
                OutputStream lnotIndexed = restFile != null ? StreamProvider.getOutputStreamForFile(restFile) : null) {
            accepted = lindexed;
            rejected = lnotIndexed;
            fsess = GsNative.filterOpen(filter, kSize, minPosCount, positiveRatio);
            cur = 0;
            inflight = 0;
            batches[0].n = 0;
            try {
                processFastqStreams(fastqs);
            } finally {
                GsNative.filterClose(fsess);
                fsess = 0;
            }
        }
        accepted = null;
        rejected = null;
    }

    @Override
    protected void readFastq(InputStream inputStream, boolean fasta) throws IOException {
        super.readFastq(inputStream, fasta);
        flush();
        while (inflight > 0) {
            collectOldest();
        }
    }

    @Override
    protected void nextEntry(ReadEntry entry, int index) throws IOException {
        Batch b = batches[cur];
        if (b.n == batchReads || b.bases.remaining() < entry.readSize) {
            flush();
            b = batches[cur];
        }
        if (b.n == 0) {
            b.bases.clear();
            b.offsets.clear();
            b.offsets.putLong(0);
            b.textUsed = 0;
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
        b.n++;
    }

    private void flush() throws IOException {
        Batch b = batches[cur];
        if (b.n == 0) {
            return;
        }
        b.ticket = GsNative.filterSubmit(fsess, b.bases, b.offsets, b.n);
        inflight++;
        if (inflight == SLOTS) {
            collectOldest();
        }
        cur = (cur + 1) % SLOTS;
        batches[cur].n = 0;
    }

    private void collectOldest() throws IOException {
        Batch b = batches[(cur + SLOTS - (inflight - 1)) % SLOTS];
        GsNative.filterCollect(fsess, b.ticket, b.accept);
        for (int i = 0; i < b.n; i++) {
            OutputStream target = b.accept.get(i) != 0 ? accepted : rejected;   // FastqBloomFilter.nextEntry :92-105
            if (target == null) {
                continue;
            }
            int start = (int) b.offsets.getLong(8 * i), size = (int) b.offsets.getLong(8 * (i + 1)) - start;
            if (replay.read.length < size) {
                replay.read = new byte[Math.max(size, 2 * replay.read.length)];
            }
            for (int p = 0; p < size; p++) {
                replay.read[p] = b.bases.get(start + p);
            }
            replay.readSize = size;
            replay.readDescriptorSize = b.descSize[i];
            System.arraycopy(b.text, b.descStart[i], replay.readDescriptor, 0, b.descSize[i]);
            replay.readProbsSize = b.probsSize[i];
            if (withProbs && b.probsSize[i] > 0) {
                if (replay.readProbs == null || replay.readProbs.length < b.probsSize[i]) {
                    replay.readProbs = new byte[b.probsSize[i]];
                }
                System.arraycopy(b.text, b.probsStart[i], replay.readProbs, 0, b.probsSize[i]);
            }
            rewriteInput(replay, target);
        }
        inflight--;
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
