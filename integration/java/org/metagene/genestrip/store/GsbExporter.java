/*
 * GsbExporter -- writes a loaded Genestrip database as the flat little-endian "GSB1" file that libgenestrip_b200 loads
 * without a JVM (gs_db_load_file; layout in genestrip_b200/csrc/gs_capi.cu).  Source only: no JDK exists in the build image
 * of genestrip_b200, so this file was not compiled there.  K-mers and value indices come from the store's own visitor
 * (KMerSortedArray.visit, C/store/KMerSortedArray.java:426-438: storage order, index = short - Short.MIN_VALUE); the blocked
 * Bloom filter keeps seed / buckets / data private (C/bloom/BlockedKMerBloomFilter.java:67-75), so they are read reflectively.
 *
 *   Database db = Database.load(new File("human_virus_db.zip"), true);     // C/store/Database.java:265-314
 *   GsbExporter.export(db, new File("human_virus_db.gsb"));
 */
package org.metagene.genestrip.store;

import java.io.BufferedOutputStream;
import java.io.DataOutputStream;
import java.io.File;
import java.io.FileOutputStream;
import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;

import org.metagene.genestrip.bloom.BlockedKMerBloomFilter;
import org.metagene.genestrip.tax.SmallTaxTree;
import org.metagene.genestrip.tax.SmallTaxTree.SmallTaxIdNode;

public class GsbExporter {
    public static void export(Database db, File out) throws IOException {
        @SuppressWarnings("unchecked")
        KMerSortedArray<SmallTaxIdNode> store = (KMerSortedArray<SmallTaxIdNode>) db.convertKMerStore();   // values = tree nodes (:136-143)
        SmallTaxTree tree = db.getTaxTree();
        final long n = store.getEntries();
        final int nValues = store.getNValues();
        BlockedKMerBloomFilter bloom = store.getFilter() instanceof BlockedKMerBloomFilter ? (BlockedKMerBloomFilter) store.getFilter() : null;
        try (DataOutputStream o = new DataOutputStream(new BufferedOutputStream(new FileOutputStream(out), 1 << 20))) {
            ByteBuffer h = ByteBuffer.allocate(56).order(ByteOrder.LITTLE_ENDIAN);
            h.put(new byte[] { 'G', 'S', 'B', '1', 0, 0, 0, 0 });
            h.putInt(1).putInt(store.getK()).putLong(n).putInt(nValues).putInt(bloom != null ? 1 : 0);
            final long seed = bloom != null ? (Long) field(bloom, "seed") : 0, buckets = bloom != null ? (Long) field(bloom, "buckets") : 0;
            h.putLong(seed).putLong(buckets).putLong(bloom != null ? buckets + 17 : 0);
            o.write(h.array());
            ByteBuffer b = ByteBuffer.allocate(8 << 10).order(ByteOrder.LITTLE_ENDIAN);
            for (long i = 0; i < n; i++) {                      // keys in storage order (position == index, :290-349)
                if (!b.hasRemaining()) { o.write(b.array(), 0, b.position()); b.clear(); }
                b.putLong(store.getKMerAt(i));
            }
            o.write(b.array(), 0, b.position()); b.clear();
            final ByteBuffer vb = b;
            final IOException[] err = new IOException[1];
            store.visit((s, kmer, index, pos) -> {              // the raw Java shorts: value index + Short.MIN_VALUE
                try {
                    if (!vb.hasRemaining()) { o.write(vb.array(), 0, vb.position()); vb.clear(); }
                    vb.putShort((short) (index + Short.MIN_VALUE));
                } catch (IOException e) { err[0] = e; }
            });
            if (err[0] != null) throw err[0];
            o.write(b.array(), 0, b.position()); b.clear();
            for (long pad = (8 - (n * 2) % 8) % 8; pad > 0; pad--) o.write(0);
            int[] parent = new int[nValues], hasNode = new int[nValues];
            for (int v = 0; v < nValues; v++) {                 // tree by value index (storeIndex, Database.java:107-128)
                SmallTaxIdNode node = store.getValueForIndex(v);
                hasNode[v] = node != null ? 1 : 0;
                parent[v] = node != null && node.getParent() != null ? node.getParent().getStoreIndex() : -1;
            }
            for (int v = 0; v < nValues; v++) { b.putInt(parent[v]); if (b.remaining() < 4) { o.write(b.array(), 0, b.position()); b.clear(); } }
            for (int v = 0; v < nValues; v++) { b.putInt(hasNode[v]); if (b.remaining() < 4) { o.write(b.array(), 0, b.position()); b.clear(); } }
            o.write(b.array(), 0, b.position()); b.clear();
            for (long pad = (8 - ((long) nValues * 8) % 8) % 8; pad > 0; pad--) o.write(0);
            if (bloom != null) {
                final long[] data = (long[]) field(bloom, "data");            // data[buckets + 17] (BlockedKMerBloomFilter.java:91-125)
                final long[][] large = (long[][]) field(bloom, "largeData");  // or 2^27-element segments (fastutil BigArrays)
                for (long i = 0; i < buckets + 17; i++) {
                    if (!b.hasRemaining()) { o.write(b.array(), 0, b.position()); b.clear(); }
                    b.putLong(data != null ? data[(int) i] : large[(int) (i >>> 27)][(int) (i & ((1 << 27) - 1))]);
                }
                o.write(b.array(), 0, b.position());
            }
        }
    }

    private static Object field(Object o, String name) throws IOException {
        try {
            java.lang.reflect.Field f = o.getClass().getDeclaredField(name);
            f.setAccessible(true);
            return f.get(o);
        } catch (ReflectiveOperationException e) {
            throw new IOException(e);
        }
    }
}
