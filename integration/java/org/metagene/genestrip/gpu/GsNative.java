// GsNative.java -- the JNI face of include/genestrip_b200.h (bound by integration/jni/gs_jni.cpp, one native per C entry
// point).  Handles travel as long; every failure surfaces as a RuntimeException carrying gs_last_error(), the way consumer
// thread failures surface in the reference (core/.../fastq/AbstractFastqReader.java:124-143).
package org.metagene.genestrip.gpu;

import java.nio.ByteBuffer;

public final class GsNative {
    static { System.loadLibrary("gsjni"); }
    private GsNative() {}

    public static native long ctxCreate(int[] devices);
    public static native void ctxDestroy(long ctx);
    // database (KMerSortedArray segments / RadixKMerStore buckets -> device), core/.../store/KMerSortedArray.java:63-71
    public static native long dbCreate(long ctx, int k, long nKmers, int nValues);
    public static native void dbPutKeys(long db, long offset, long[] keys, int n);
    public static native void dbPutValues(long db, long offset, short[] vals, int n);
    public static native void dbPutRadixBucket(long db, int radixBits, int radix, long[] entries, int n);
    public static native void dbSetTree(long db, int[] parentByValueIndex, int[] hasNode);
    public static native void dbSetBloomBlocked(long db, long seed, long buckets, long[] words);
    public static native void dbBuildBloom(long db);   // the store's optimized filter, built on the device (bit-identical)
    public static native void dbFinalize(long db);
    public static native void dbDestroy(long db);
    public static native long dbUpdate(long db, ByteBuffer seq, long nBytes, long[] regionOffsets, int[] regionValueIndex, boolean upperCase);
    public static native void dbGetValues(long db, long offset, short[] vals, int n);
    // match
    public static native long matchOpen(long db, boolean classify, boolean countUnique, int maxKmerResCounts, boolean useBloom, int maxPaths,
                                        int minKmersForClass, double maxTaxErr, double maxClassErr, boolean wantRuns);
    public static native ByteBuffer allocPinned(long bytes);
    public static native void freePinned(ByteBuffer b);
    public static native long matchSubmit(long sess, ByteBuffer bases, ByteBuffer offsets, int nReads, long firstReadNo);
    /** {results (16 bytes per read), max-contig events (16 bytes each)}: views of the session's pinned staging. */
    public static native ByteBuffer[] matchCollect(long sess, long ticket);
    /** Sessions opened with wantRuns: {results, events, run offsets (nReads + 1 longs), runs (label, len: 8 bytes each)}; copies. */
    public static native ByteBuffer[] matchCollectRuns(long sess, long ticket, int nReads, long runsCapacity);
    public static native void matchFinish(long sess, ByteBuffer counts, ByteBuffer topCounts);
    public static native void matchClose(long sess);
    public static native long[] matchSubmitFastq(long sess, ByteBuffer text, long nBytes, long firstReadNo);
    public static native ByteBuffer[] matchCollectFastq(long sess, long ticket);
    // filter
    public static native long filterCreate(long ctx, int kind, long p0, long p1, long[] factors, long[] words);
    public static native long filterLoadFile(long ctx, String path);   // flat GSF1 index file (GsfExporter), no Java object stream
    public static native void filterSaveFile(long filter, String path);
    public static native long filterOpen(long filter, int k, int minPosCount, double posRatio);
    public static native long filterSubmit(long fsess, ByteBuffer bases, ByteBuffer offsets, int nReads);
    public static native void filterCollect(long fsess, long ticket, ByteBuffer accept);
    public static native long[] filterSubmitFastq(long fsess, ByteBuffer text, long nBytes);
    public static native ByteBuffer[] filterCollectFastq(long fsess, long ticket);
    public static native void filterClose(long fsess);
    public static native void filterDestroy(long filter);
    // block-gzip input
    public static native void inflateBlocks(long ctx, ByteBuffer comp, long compBytes, ByteBuffer blocks, int nBlocks, ByteBuffer out, long outBytes);
}
