// gs_jni.cpp -- the thin JNI shim a Genestrip maintainer adds to bind include/genestrip_b200.h.
//
// This image has no JDK (no jni.h): integration/build.sh builds libgsjni.so and the Java classes where one exists, and
// tests/test_capi_symbols.py compiles this file against tests/jni_stub/jni.h (the subset of the JNI API used here, signatures
// from the JNI specification) so that it cannot rot unnoticed.  With a JDK:
//   g++ -O2 -fPIC -shared -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include gs_jni.cpp -L../../genestrip_b200/_lib -lgenestrip_b200 -o libgsjni.so
// Java side: integration/java/org/metagene/genestrip/gpu/GsNative.java (one native method per function below).
// Conventions: handles travel as jlong; every failure becomes a RuntimeException carrying gs_last_error(), the same
// way consumer-thread failures surface in the reference (C/fastq/AbstractFastqReader.java:124-143).
#include <jni.h>

#include "genestrip_b200.h"

static void throwLast(JNIEnv* env) { env->ThrowNew(env->FindClass("java/lang/RuntimeException"), gs_last_error()); }
#define CHECK(rc) do { if ((rc) != GS_OK) { throwLast(env); return; } } while (0)
#define J(name) Java_org_metagene_genestrip_gpu_GsNative_##name

extern "C" {

JNIEXPORT jlong JNICALL J(ctxCreate)(JNIEnv* env, jclass, jintArray devices) {
    jint* d = devices ? env->GetIntArrayElements(devices, nullptr) : nullptr;
    gs_ctx* c = gs_ctx_create((const int*)d, devices ? env->GetArrayLength(devices) : 0);
    if (d) env->ReleaseIntArrayElements(devices, d, JNI_ABORT);
    if (!c) throwLast(env);
    return (jlong)c;
}
JNIEXPORT void JNICALL J(ctxDestroy)(JNIEnv*, jclass, jlong c) { gs_ctx_destroy((gs_ctx*)c); }

JNIEXPORT jlong JNICALL J(dbCreate)(JNIEnv* env, jclass, jlong ctx, jint k, jlong nKmers, jint nValues) {
    gs_db* db = gs_db_create((gs_ctx*)ctx, k, (uint64_t)nKmers, nValues);
    if (!db) throwLast(env);
    return (jlong)db;
}
// One segment of KMerSortedArray.kmers / valueIndexes (or of the BigArrays segments largeKmers[i] / largeValueIndexes[i],
// C/store/KMerSortedArray.java:63-71): pinned only for the duration of the copy.
JNIEXPORT void JNICALL J(dbPutKeys)(JNIEnv* env, jclass, jlong db, jlong offset, jlongArray keys, jint n) {
    void* p = env->GetPrimitiveArrayCritical(keys, nullptr);
    int rc = gs_db_put_keys((gs_db*)db, (uint64_t)offset, (const int64_t*)p, (uint64_t)n);
    env->ReleasePrimitiveArrayCritical(keys, p, JNI_ABORT);
    CHECK(rc);
}
JNIEXPORT void JNICALL J(dbPutValues)(JNIEnv* env, jclass, jlong db, jlong offset, jshortArray vals, jint n) {
    void* p = env->GetPrimitiveArrayCritical(vals, nullptr);
    int rc = gs_db_put_values((gs_db*)db, (uint64_t)offset, (const int16_t*)p, (uint64_t)n);
    env->ReleasePrimitiveArrayCritical(vals, p, JNI_ABORT);
    CHECK(rc);
}
JNIEXPORT void JNICALL J(dbPutRadixBucket)(JNIEnv* env, jclass, jlong db, jint radixBits, jint radix, jlongArray entries, jint n) {
    void* p = env->GetPrimitiveArrayCritical(entries, nullptr);
    int rc = gs_db_put_radix_bucket((gs_db*)db, radixBits, (uint32_t)radix, (const int64_t*)p, (uint32_t)n);
    env->ReleasePrimitiveArrayCritical(entries, p, JNI_ABORT);
    CHECK(rc);
}
JNIEXPORT void JNICALL J(dbSetTree)(JNIEnv* env, jclass, jlong db, jintArray parent, jintArray hasNode) {
    jint* p = env->GetIntArrayElements(parent, nullptr);
    jint* h = hasNode ? env->GetIntArrayElements(hasNode, nullptr) : nullptr;
    int rc = gs_db_set_tree((gs_db*)db, (const int32_t*)p, (const int32_t*)h, env->GetArrayLength(parent));
    env->ReleaseIntArrayElements(parent, p, JNI_ABORT);
    if (h) env->ReleaseIntArrayElements(hasNode, h, JNI_ABORT);
    CHECK(rc);
}
JNIEXPORT void JNICALL J(dbSetBloomBlocked)(JNIEnv* env, jclass, jlong db, jlong seed, jlong buckets, jlongArray words) {
    void* p = env->GetPrimitiveArrayCritical(words, nullptr);
    int rc = gs_db_set_bloom_blocked((gs_db*)db, seed, (uint64_t)buckets, (const int64_t*)p, (uint64_t)env->GetArrayLength(words));
    env->ReleasePrimitiveArrayCritical(words, p, JNI_ABORT);
    CHECK(rc);
}
// the store's optimized filter (KMerSortedArray.optimize, C/store/KMerSortedArray.java:409-422) built on the device: BlockedKMerBloomFilter
// keeps its words private, and the device-built filter is bit-identical (tests/test_gpu_match.py)
JNIEXPORT void JNICALL J(dbBuildBloom)(JNIEnv* env, jclass, jlong db) { CHECK(gs_db_build_bloom_blocked((gs_db*)db, nullptr, 0)); }
JNIEXPORT void JNICALL J(dbFinalize)(JNIEnv* env, jclass, jlong db) { CHECK(gs_db_finalize((gs_db*)db)); }
JNIEXPORT void JNICALL J(dbDestroy)(JNIEnv*, jclass, jlong db) { gs_db_destroy((gs_db*)db); }

JNIEXPORT jlong JNICALL J(matchOpen)(JNIEnv* env, jclass, jlong db, jboolean classify, jboolean countUnique, jint maxKmerResCounts,
                                      jboolean useBloom, jint maxPaths, jint minKmersForClass, jdouble maxTaxErr, jdouble maxClassErr,
                                      jboolean wantRuns) {
    gs_match_cfg c;
    gs_match_cfg_default(&c);
    c.classify_reads = classify; c.count_unique_kmers = countUnique; c.max_kmer_res_counts = maxKmerResCounts;
    c.use_bloom_filter = useBloom; c.max_classification_paths = maxPaths; c.min_kmers_for_class = minKmersForClass;
    c.max_read_tax_error_count = maxTaxErr; c.max_read_class_error_count = maxClassErr; c.want_runs = wantRuns;
    gs_sess* s = gs_match_open((gs_db*)db, &c);
    if (!s) throwLast(env);
    return (jlong)s;
}
// Batches live in pinned memory obtained from gs_alloc_pinned and wrapped as direct ByteBuffers: the parser thread
// writes reads straight into them (no copy on submit).
JNIEXPORT jobject JNICALL J(allocPinned)(JNIEnv* env, jclass, jlong bytes) {
    void* p = gs_alloc_pinned((size_t)bytes);
    if (!p) { throwLast(env); return nullptr; }
    return env->NewDirectByteBuffer(p, bytes);
}
JNIEXPORT void JNICALL J(freePinned)(JNIEnv* env, jclass, jobject buf) { gs_free_pinned(env->GetDirectBufferAddress(buf)); }
JNIEXPORT jlong JNICALL J(matchSubmit)(JNIEnv* env, jclass, jlong s, jobject bases, jobject offsets, jint nReads, jlong firstReadNo) {
    gs_ticket t = 0;
    if (gs_match_submit((gs_sess*)s, (const uint8_t*)env->GetDirectBufferAddress(bases),
                        (const uint64_t*)env->GetDirectBufferAddress(offsets), (uint32_t)nReads, (uint64_t)firstReadNo, &t) != GS_OK) throwLast(env);
    return (jlong)t;
}
// Zero-copy results: [0] = direct buffer over nReads x gs_read_result (16 bytes each), [1] = over the max-contig events.
JNIEXPORT jobjectArray JNICALL J(matchCollect)(JNIEnv* env, jclass, jlong s, jlong ticket) {
    const gs_read_result* out; const gs_maxcontig_event* ev; uint32_t n = 0, nev = 0;
    if (gs_match_collect_view((gs_sess*)s, (gs_ticket)ticket, &out, &n, &ev, &nev) != GS_OK) { throwLast(env); return nullptr; }
    jobjectArray arr = env->NewObjectArray(2, env->FindClass("java/nio/ByteBuffer"), nullptr);
    env->SetObjectArrayElement(arr, 0, env->NewDirectByteBuffer((void*)out, (jlong)n * (jlong)sizeof(gs_read_result)));
    env->SetObjectArrayElement(arr, 1, env->NewDirectByteBuffer((void*)ev, (jlong)nev * (jlong)sizeof(gs_maxcontig_event)));
    return arr;
}
// Sessions opened with wantRuns (kraken-style output): the per-read contig runs come back too.  Copies into fresh direct
// buffers: [0] results, [1] events, [2] run offsets (nReads + 1 x u64), [3] runs (label, len).  runsCapacity = the batch's k-mer
// count (an upper bound of its runs).
JNIEXPORT jobjectArray JNICALL J(matchCollectRuns)(JNIEnv* env, jclass, jlong s, jlong ticket, jint nReads, jlong runsCapacity) {
    jclass bb = env->FindClass("java/nio/ByteBuffer");
    jmethodID alloc = env->GetStaticMethodID(bb, "allocateDirect", "(I)Ljava/nio/ByteBuffer;");
    const jlong sizes[4] = {(jlong)nReads * (jlong)sizeof(gs_read_result), (jlong)nReads * (jlong)sizeof(gs_maxcontig_event),
                            ((jlong)nReads + 1) * 8, (runsCapacity > 0 ? runsCapacity : 1) * (jlong)sizeof(gs_run)};
    jobject bufs[4];
    void* p[4];
    for (int i = 0; i < 4; i++) {
        bufs[i] = env->CallStaticObjectMethod(bb, alloc, (jint)sizes[i]);
        if (!bufs[i]) return nullptr;
        p[i] = env->GetDirectBufferAddress(bufs[i]);
    }
    uint32_t nev = 0;
    if (gs_match_collect((gs_sess*)s, (gs_ticket)ticket, (gs_read_result*)p[0], (gs_maxcontig_event*)p[1], (uint32_t)nReads, &nev,
                         (uint64_t*)p[2], (gs_run*)p[3], (uint64_t)(runsCapacity > 0 ? runsCapacity : 1)) != GS_OK) { throwLast(env); return nullptr; }
    jobjectArray arr = env->NewObjectArray(4, bb, nullptr);
    env->SetObjectArrayElement(arr, 0, bufs[0]);
    env->SetObjectArrayElement(arr, 1, env->NewDirectByteBuffer(p[1], (jlong)nev * (jlong)sizeof(gs_maxcontig_event)));
    env->SetObjectArrayElement(arr, 2, bufs[2]);
    env->SetObjectArrayElement(arr, 3, bufs[3]);
    return arr;
}
// End of run: nValues x gs_taxon_counts (80 bytes each) into a caller-provided direct buffer.
JNIEXPORT void JNICALL J(matchFinish)(JNIEnv* env, jclass, jlong s, jobject counts, jobject topCounts) {
    CHECK(gs_match_finish((gs_sess*)s, (gs_taxon_counts*)env->GetDirectBufferAddress(counts),
                          topCounts ? (int16_t*)env->GetDirectBufferAddress(topCounts) : nullptr));
}
JNIEXPORT void JNICALL J(matchClose)(JNIEnv*, jclass, jlong s) { gs_match_close((gs_sess*)s); }

JNIEXPORT jlong JNICALL J(filterCreate)(JNIEnv* env, jclass, jlong ctx, jint kind, jlong p0, jlong p1, jlongArray factors, jlongArray words) {
    jlong* f = factors ? env->GetLongArrayElements(factors, nullptr) : nullptr;
    void* w = env->GetPrimitiveArrayCritical(words, nullptr);
    gs_filter* flt = gs_filter_create((gs_ctx*)ctx, kind, p0, p1, (const int64_t*)f, (const int64_t*)w, (uint64_t)env->GetArrayLength(words));
    env->ReleasePrimitiveArrayCritical(words, w, JNI_ABORT);
    if (f) env->ReleaseLongArrayElements(factors, f, JNI_ABORT);
    if (!flt) throwLast(env);
    return (jlong)flt;
}
JNIEXPORT jlong JNICALL J(filterLoadFile)(JNIEnv* env, jclass, jlong ctx, jstring path) {
    const char* p = env->GetStringUTFChars(path, nullptr);
    gs_filter* flt = gs_filter_load_file((gs_ctx*)ctx, p);
    env->ReleaseStringUTFChars(path, p);
    if (!flt) throwLast(env);
    return (jlong)flt;
}
JNIEXPORT void JNICALL J(filterSaveFile)(JNIEnv* env, jclass, jlong flt, jstring path) {
    const char* p = env->GetStringUTFChars(path, nullptr);
    const int rc = gs_filter_save_file((gs_filter*)flt, p);
    env->ReleaseStringUTFChars(path, p);
    if (rc) throwLast(env);
}
JNIEXPORT jlong JNICALL J(filterOpen)(JNIEnv* env, jclass, jlong flt, jint k, jint minPosCount, jdouble posRatio) {
    gs_fsess* s = gs_filter_open((gs_filter*)flt, k, minPosCount, posRatio);
    if (!s) throwLast(env);
    return (jlong)s;
}
JNIEXPORT jlong JNICALL J(filterSubmit)(JNIEnv* env, jclass, jlong s, jobject bases, jobject offsets, jint nReads) {
    gs_ticket t = 0;
    if (gs_filter_submit((gs_fsess*)s, (const uint8_t*)env->GetDirectBufferAddress(bases),
                         (const uint64_t*)env->GetDirectBufferAddress(offsets), (uint32_t)nReads, &t) != GS_OK) throwLast(env);
    return (jlong)t;
}
JNIEXPORT void JNICALL J(filterCollect)(JNIEnv* env, jclass, jlong s, jlong ticket, jobject accept) {
    CHECK(gs_filter_collect((gs_fsess*)s, (gs_ticket)ticket, (uint8_t*)env->GetDirectBufferAddress(accept)));
}
// ---- raw FASTQ text (GPU feeder): text = direct buffer over pinned memory holding whole records.
// Returns {ticket, nReads, status, totalKmers, totalBps}; ticket == 0 && status != 0: parse this chunk with the Java reader.
JNIEXPORT jlongArray JNICALL J(matchSubmitFastq)(JNIEnv* env, jclass, jlong s, jobject text, jlong nBytes, jlong firstReadNo) {
    gs_fastq_info info; gs_ticket t = 0;
    if (gs_match_submit_fastq((gs_sess*)s, (const uint8_t*)env->GetDirectBufferAddress(text), (uint64_t)nBytes, (uint64_t)firstReadNo, &info, &t) != GS_OK) { throwLast(env); return nullptr; }
    const jlong v[5] = {(jlong)t, (jlong)info.n_reads, (jlong)info.status, (jlong)info.total_kmers, (jlong)info.total_bps};
    jlongArray arr = env->NewLongArray(5);
    env->SetLongArrayRegion(arr, 0, 5, v);
    return arr;
}
// [0] results, [1] max-contig events, [2] header offset per event (u32), [3] record table (nReads + 1) x gs_fastq_rec
JNIEXPORT jobjectArray JNICALL J(matchCollectFastq)(JNIEnv* env, jclass, jlong s, jlong ticket) {
    const gs_read_result* out; const gs_maxcontig_event* ev; const uint32_t* eh; const gs_fastq_rec* recs; uint32_t n = 0, nev = 0;
    if (gs_match_collect_fastq((gs_sess*)s, (gs_ticket)ticket, &out, &n, &ev, &eh, &nev, &recs, nullptr, nullptr, 0) != GS_OK) { throwLast(env); return nullptr; }
    jobjectArray arr = env->NewObjectArray(4, env->FindClass("java/nio/ByteBuffer"), nullptr);
    env->SetObjectArrayElement(arr, 0, env->NewDirectByteBuffer((void*)out, (jlong)n * (jlong)sizeof(gs_read_result)));
    env->SetObjectArrayElement(arr, 1, env->NewDirectByteBuffer((void*)ev, (jlong)nev * (jlong)sizeof(gs_maxcontig_event)));
    env->SetObjectArrayElement(arr, 2, env->NewDirectByteBuffer((void*)eh, (jlong)nev * 4));
    env->SetObjectArrayElement(arr, 3, env->NewDirectByteBuffer((void*)recs, ((jlong)n + 1) * (jlong)sizeof(gs_fastq_rec)));
    return arr;
}
JNIEXPORT jlongArray JNICALL J(filterSubmitFastq)(JNIEnv* env, jclass, jlong s, jobject text, jlong nBytes) {
    gs_fastq_info info; gs_ticket t = 0;
    if (gs_filter_submit_fastq((gs_fsess*)s, (const uint8_t*)env->GetDirectBufferAddress(text), (uint64_t)nBytes, &info, &t) != GS_OK) { throwLast(env); return nullptr; }
    const jlong v[5] = {(jlong)t, (jlong)info.n_reads, (jlong)info.status, (jlong)info.total_kmers, (jlong)info.total_bps};
    jlongArray arr = env->NewLongArray(5);
    env->SetLongArrayRegion(arr, 0, 5, v);
    return arr;
}
JNIEXPORT jobjectArray JNICALL J(filterCollectFastq)(JNIEnv* env, jclass, jlong s, jlong ticket) {
    const uint8_t* acc; const gs_fastq_rec* recs; uint32_t n = 0;
    if (gs_filter_collect_fastq((gs_fsess*)s, (gs_ticket)ticket, &acc, &n, &recs) != GS_OK) { throwLast(env); return nullptr; }
    jobjectArray arr = env->NewObjectArray(2, env->FindClass("java/nio/ByteBuffer"), nullptr);
    env->SetObjectArrayElement(arr, 0, env->NewDirectByteBuffer((void*)acc, (jlong)n));
    env->SetObjectArrayElement(arr, 1, env->NewDirectByteBuffer((void*)recs, ((jlong)n + 1) * (jlong)sizeof(gs_fastq_rec)));
    return arr;
}
// ---- block-gzip input: comp / out = direct buffers over pinned memory, blocks = direct buffer of gs_deflate_block[nBlocks]
// (filled by the Java reader from the members' headers and trailers); throws on a corrupt member like GZIPInputStream does
JNIEXPORT void JNICALL J(inflateBlocks)(JNIEnv* env, jclass, jlong ctx, jobject comp, jlong compBytes, jobject blocks, jint nBlocks, jobject out, jlong outBytes) {
    CHECK(gs_inflate_blocks((gs_ctx*)ctx, (const uint8_t*)env->GetDirectBufferAddress(comp), (uint64_t)compBytes,
                            (gs_deflate_block*)env->GetDirectBufferAddress(blocks), (uint32_t)nBlocks,
                            (uint8_t*)env->GetDirectBufferAddress(out), (uint64_t)outBytes));
}
// ---- db goal, update phase (DBGoal.MyFastaReader.handleStore): the regions of one FASTA batch, line ends stripped
JNIEXPORT jlong JNICALL J(dbUpdate)(JNIEnv* env, jclass, jlong db, jobject seq, jlong nBytes, jlongArray regionOffsets, jintArray regionValueIndex, jboolean upperCase) {
    const jsize nr = env->GetArrayLength(regionValueIndex);
    jlong* off = env->GetLongArrayElements(regionOffsets, nullptr);
    jint* vid = env->GetIntArrayElements(regionValueIndex, nullptr);
    uint64_t changed = 0;
    const int rc = gs_db_update((gs_db*)db, (const uint8_t*)env->GetDirectBufferAddress(seq), (uint64_t)nBytes, (const uint64_t*)off, (const int32_t*)vid,
                                (uint32_t)nr, upperCase ? 1 : 0, &changed);
    env->ReleaseLongArrayElements(regionOffsets, off, JNI_ABORT);
    env->ReleaseIntArrayElements(regionValueIndex, vid, JNI_ABORT);
    if (rc != GS_OK) throwLast(env);
    return (jlong)changed;
}
JNIEXPORT void JNICALL J(dbGetValues)(JNIEnv* env, jclass, jlong db, jlong offset, jshortArray vals, jint n) {
    void* p = env->GetPrimitiveArrayCritical(vals, nullptr);
    const int rc = gs_db_get_values((gs_db*)db, (uint64_t)offset, (int16_t*)p, (uint64_t)n);
    env->ReleasePrimitiveArrayCritical(vals, p, 0);
    if (rc != GS_OK) throwLast(env);
}
JNIEXPORT void JNICALL J(filterClose)(JNIEnv*, jclass, jlong s) { gs_filter_close((gs_fsess*)s); }
JNIEXPORT void JNICALL J(filterDestroy)(JNIEnv*, jclass, jlong f) { gs_filter_destroy((gs_filter*)f); }

}  // extern "C"
