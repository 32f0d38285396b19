/* call_order.c -- a plain C99 host of include/genestrip_b200.h: the documented call order of the boundary, written the way
 * the JNI shim (integration/jni/gs_jni.cpp) or any other FFI would drive it.  It includes nothing but the public header and
 * links against libgenestrip_b200.so only.
 *
 *   cc -std=c99 -Wall -Werror -I include integration/c/call_order.c -L genestrip_b200/_lib -lgenestrip_b200 -o call_order
 *
 * Without a CUDA device it checks what can be checked on the host (ABI version, defaults of gs_match_cfg, the packer, the
 * error convention: no device -> NULL / GS_ERR_CUDA and a message, never a fallback) and exits 0.  With a device it runs the
 * whole order on a toy store: context -> database (keys, values, tree, Bloom filter built on the device) -> finalize ->
 * lookup -> match session (submit / collect / finish) -> filter index + session -> teardown, and checks the answers it can
 * derive by hand (tests/test_capi_symbols.py runs it both ways). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "genestrip_b200.h"

#define CHECK(cond)                                                                                   \
    do {                                                                                              \
        if (!(cond)) {                                                                                \
            fprintf(stderr, "%s:%d: check failed: %s (%s)\n", __FILE__, __LINE__, #cond, gs_last_error()); \
            return 1;                                                                                 \
        }                                                                                             \
    } while (0)

/* C=0 G=1 A=2 T=3, first base in the top bits (C/util/CGAT.java:66-69, 159-180); canonical = max(fwd, reverse complement) */
static int code_of(char c) { return c == 'C' ? 0 : c == 'G' ? 1 : c == 'A' ? 2 : c == 'T' ? 3 : -1; }
static int64_t canonical(const char* s, int k) {
    int64_t fwd = 0, rc = 0;
    for (int i = 0; i < k; i++) {
        int c = code_of(s[i]);
        fwd = (fwd << 2) | c;
        rc |= (int64_t)(c ^ 1) << (2 * i);
    }
    return fwd > rc ? fwd : rc;
}
static int cmp64(const void* a, const void* b) {
    int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return x < y ? -1 : x > y;
}

int main(void) {
    CHECK(gs_abi_version() == GS_ABI_VERSION);
    gs_match_cfg cfg;
    gs_match_cfg_default(&cfg);
    CHECK(cfg.classify_reads == 1 && cfg.count_unique_kmers == 1 && cfg.max_classification_paths == 10 && cfg.min_kmers_for_class == 1);
    CHECK(cfg.max_read_tax_error_count == -1.0 && cfg.max_read_class_error_count == -1.0 && cfg.layout == GS_LAYOUT_TABLE);

    /* the packer needs no device */
    const char* seq = "CGATCGATNCGATcgatCGATCGATCGATCGATCGAT";
    uint64_t codes[2];
    uint32_t valid[2];
    CHECK(gs_pack_bases((const uint8_t*)seq, strlen(seq), codes, valid, 1) == GS_OK);
    CHECK((codes[0] >> 56) == 0x1B /* C G A T = 00 01 10 11 */ && (valid[0] & 0x1FF) == 0x0FF /* the 9th base is N */);
    CHECK(((valid[0] >> 13) & 0xF) == 0 /* lower case is not a base (CGAT.java:60-69) */);

    gs_ctx* ctx = gs_ctx_create(NULL, 0);
    if (!ctx) {
        /* no CUDA device: the library says so and does nothing else */
        CHECK(strlen(gs_last_error()) > 0);
        CHECK(gs_alloc_pinned(1 << 20) == NULL);
        printf("call_order: host-only checks ok (no CUDA device: %s)\n", gs_last_error());
        return 0;
    }

    /* ---- a toy store: the 11 distinct canonical 5-mers of one sequence; value 0 is the root, 1 and 2 its children */
    enum { K = 5, V = 3 };
    const char* genome = "ACGTTGCAAGGCTTAC";
    int64_t keys[16];
    int nk = 0;
    for (int i = 0; i + K <= (int)strlen(genome); i++) keys[nk++] = canonical(genome + i, K);
    qsort(keys, (size_t)nk, sizeof(int64_t), cmp64);
    int nu = 0;
    for (int i = 0; i < nk; i++) if (i == 0 || keys[i] != keys[i - 1]) keys[nu++] = keys[i];
    int16_t vals[16];
    for (int i = 0; i < nu; i++) vals[i] = (int16_t)((i % 2 ? 1 : 2) - 32768); /* value index + Short.MIN_VALUE */
    const int32_t parent[V] = {-1, 0, 0};

    gs_db* db = gs_db_create(ctx, K, (uint64_t)nu, V);
    CHECK(db != NULL);
    CHECK(gs_db_put_keys(db, 0, keys, (uint64_t)nu) == GS_OK);
    CHECK(gs_db_put_values(db, 0, vals, (uint64_t)nu) == GS_OK);
    CHECK(gs_db_set_tree(db, parent, NULL, V) == GS_OK);
    CHECK(gs_db_build_bloom_blocked(db, NULL, 0) == GS_OK);
    CHECK(gs_db_put_keys(db, (uint64_t)nu, keys, 1) != GS_OK); /* out of range: an error code, not a crash */
    CHECK(gs_db_finalize(db) == GS_OK);
    CHECK(gs_db_finalize(db) == GS_ERR_STATE);

    int32_t vidx[16];
    int64_t pos[16];
    CHECK(gs_db_lookup(db, keys, (uint64_t)nu, 1, vidx, pos) == GS_OK);
    for (int i = 0; i < nu; i++) CHECK(vidx[i] == (i % 2 ? 1 : 2) && pos[i] == i);

    /* ---- match: three reads in pinned memory; the first is the genome itself */
    const char* reads[3] = {"ACGTTGCAAGGCTTAC", "GGGGGGGGGGGG", "ACG"};
    uint8_t* bases = (uint8_t*)gs_alloc_pinned(256);
    uint64_t* offsets = (uint64_t*)gs_alloc_pinned(4 * sizeof(uint64_t));
    CHECK(bases && offsets);
    offsets[0] = 0;
    for (int r = 0; r < 3; r++) {
        memcpy(bases + offsets[r], reads[r], strlen(reads[r]));
        offsets[r + 1] = offsets[r] + strlen(reads[r]);
    }
    gs_sess* sess = gs_match_open(db, &cfg);
    CHECK(sess != NULL);
    gs_ticket t = 0;
    CHECK(gs_match_submit(sess, bases, offsets, 3, 0, &t) == GS_OK && t != 0);
    gs_read_result res[3];
    gs_maxcontig_event ev[V];
    uint32_t nev = 0;
    CHECK(gs_match_collect(sess, t, res, ev, V, &nev, NULL, NULL, 0) == GS_OK);
    CHECK((res[0].flags & GS_READ_FOUND) && !(res[1].flags & GS_READ_FOUND) && !(res[2].flags & GS_READ_FOUND));
    CHECK(res[0].class_vidx == 2 /* 7 k-mers vote for value 2, 5 for value 1 */ && res[0].read_kmers == 7 && res[1].class_vidx == -1 && res[2].class_vidx == -1);
    CHECK(gs_match_collect(sess, t, res, ev, V, &nev, NULL, NULL, 0) == GS_ERR_STATE); /* a ticket is collected once */
    gs_taxon_counts counts[V];
    CHECK(gs_match_finish(sess, counts, NULL) == GS_OK);
    CHECK(counts[1].kmers == 5 && counts[2].kmers == 7 && counts[1].unique_kmers + counts[2].unique_kmers == nu && counts[2].reads == 1 && counts[0].reads == 0);
    gs_match_close(sess);

    /* ---- filter: the store's own blocked Bloom filter layout as a `filter` index (kind 0); an empty one accepts nothing */
    int64_t words[18];
    memset(words, 0, sizeof(words));
    gs_filter* flt = gs_filter_create(ctx, GS_BLOOM_BLOCKED, 42, 1, NULL, words, 18);
    CHECK(flt != NULL);
    gs_fsess* fs = gs_filter_open(flt, K, 1, 0.2);
    CHECK(fs != NULL);
    CHECK(gs_filter_submit(fs, bases, offsets, 3, &t) == GS_OK);
    uint8_t accept[3] = {9, 9, 9};
    CHECK(gs_filter_collect(fs, t, accept) == GS_OK);
    CHECK(accept[0] == 0 && accept[1] == 0 && accept[2] == 0);
    gs_filter_close(fs);
    gs_filter_destroy(flt);

    gs_free_pinned(bases);
    gs_free_pinned(offsets);
    gs_db_destroy(db);
    gs_ctx_destroy(ctx);
    printf("call_order: full call order ok (%d stored k-mers)\n", nu);
    return 0;
}
