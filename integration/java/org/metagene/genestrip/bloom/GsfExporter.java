/*
 * GsfExporter -- writes a loaded Genestrip filter index (`*_index.ser.gz`, the object KMerProbFilter.load returns in
 * LoadIndexGoal, C/goals/LoadIndexGoal.java:92-104) as the flat little-endian "GSF1" file that libgenestrip_b200 loads without
 * a JVM (gs_filter_load_file; layout in genestrip_b200/csrc/gs_capi.cu and, byte for byte, in
 * tests/test_gpu_goals.py::test_filter_index_file_round_trip).  Source only: no JDK exists in the build image of
 * genestrip_b200, so this file was not compiled there (integration/build.sh compiles it where javac is present).
 * The hashed filters' fields are protected and this class sits in their package; BlockedKMerBloomFilter keeps seed / buckets /
 * data private (C/bloom/BlockedKMerBloomFilter.java:67-75), so those are read reflectively, like GsbExporter does.
 *
 *   KMerProbFilter index = KMerProbFilter.load(new File("human_virus_index.ser.gz"));
 *   GsfExporter.export(index, new File("human_virus_index.gsf"));
 *
 * The same arrays can be handed over in memory instead: GsNative.filterCreate(ctx, kind, p0, p1, factors, words) -- that is what
 * GpuFastqBloomFilter does with the object LoadIndexGoal delivers.
 */
package org.metagene.genestrip.bloom;

import java.io.BufferedOutputStream;
import java.io.DataOutputStream;
import java.io.File;
import java.io.FileOutputStream;
import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;

public class GsfExporter {
    public static final int BLOCKED = 0, XOR = 1, MURMUR = 2;   // GS_BLOOM_* of include/genestrip_b200.h

    public static void export(KMerProbFilter filter, File out) throws IOException {
        final int kind;
        final long p0, p1;
        long[] factors = new long[0];
        final long nWords;
        final long[] data;
        final long[][] large;
        if (filter instanceof BlockedKMerBloomFilter) {
            kind = BLOCKED;
            p0 = (Long) field(filter, BlockedKMerBloomFilter.class, "seed");
            p1 = (Long) field(filter, BlockedKMerBloomFilter.class, "buckets");
            nWords = p1 + 17;                                                     // BlockedKMerBloomFilter.java:91-125
            data = (long[]) field(filter, BlockedKMerBloomFilter.class, "data");
            large = (long[][]) field(filter, BlockedKMerBloomFilter.class, "largeData");
        } else if (filter instanceof AbstractKMerBloomFilter) {
            // same package: the protected fields are visible (AbstractKMerBloomFilter.java:52-62), LargeBitVector's arrays are public
            final AbstractKMerBloomFilter f = (AbstractKMerBloomFilter) filter;
            kind = filter instanceof XORKMerBloomFilter ? XOR : MURMUR;
            p0 = f.bits;
            p1 = f.hashes;
            factors = f.hashFactors;
            nWords = (p0 + 63) / 64;
            data = f.bitVector.bits;                 // small vectors: one long[] (C/util/LargeBitVector.java:56)
            large = f.bitVector.largeBits;           // large ones: fastutil BigArrays segments (:58)
        } else {
            throw new IOException("unknown filter class " + filter.getClass());
        }
        try (DataOutputStream o = new DataOutputStream(new BufferedOutputStream(new FileOutputStream(out), 1 << 20))) {
            ByteBuffer h = ByteBuffer.allocate(48).order(ByteOrder.LITTLE_ENDIAN);
            h.put(new byte[] { 'G', 'S', 'F', '1', 0, 0, 0, 0 });
            h.putInt(1).putInt(kind).putLong(p0).putLong(p1).putLong(kind == BLOCKED ? 0 : p1).putLong(nWords);
            o.write(h.array());
            ByteBuffer b = ByteBuffer.allocate(8 << 10).order(ByteOrder.LITTLE_ENDIAN);
            if (kind != BLOCKED) for (int i = 0; i < (int) p1; i++) b.putLong(factors[i]);   // at most a few dozen
            for (long i = 0; i < nWords; i++) {
                if (!b.hasRemaining()) { o.write(b.array(), 0, b.position()); b.clear(); }
                b.putLong(data != null ? data[(int) i] : large[(int) (i >>> 27)][(int) (i & ((1 << 27) - 1))]);   // fastutil BigArrays segments
            }
            o.write(b.array(), 0, b.position());
        }
    }

    private static Object field(Object o, Class<?> c, String name) throws IOException {
        try {
            java.lang.reflect.Field f = c.getDeclaredField(name);
            f.setAccessible(true);
            return f.get(o);
        } catch (ReflectiveOperationException e) {
            throw new IOException(e);
        }
    }
}
