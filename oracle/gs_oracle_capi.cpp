// gs_oracle_capi.cpp -- C API over the CPU ORACLE (test infrastructure; see gs_oracle.hpp header).
// Loaded through ctypes by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg only.
//
// Also holds the synthetic DB builder that restates the reference's `db` goal semantics
// (C/goals/refseq/FillDBGoal.java:100-166, 281-295; C/goals/refseq/DBGoal.java:104-128, 234-253;
// C/refseq/AbstractStoreFastaReader.java:88-120; C/goals/refseq/BloomIndexGoal.java:66-111) with one
// documented deviation: the fill-time duplicate check is exact instead of a 1e-11 Bloom filter.
#include "gs_oracle.hpp"

#include <atomic>
#include <thread>

using namespace gso;

struct Genome { std::string taxid; std::string seq; bool fill; };

struct gso_db {
    int k = 31;
    bool useRadix = false; int radixBits = 17; double optFpp = 0.01;
    TaxTree full;
    std::vector<Genome> genomes;
    std::vector<std::string> requested;
    Database db;
    std::unique_ptr<KMerProbFilter> index;  // `filter` goal index (BloomIndexGoal)
    std::string error;
    bool skipUpdate = false;  // tests of the GPU update phase: stop after the fill phase (FillDBGoal) so that the update can be replayed
};

struct gso_run {
    std::string csv, filtered, kraken, rest;
    jlong totalReads = 0, totalKMers = 0, totalBPs = 0;
    int nValues = 0;
    std::vector<jlong> kmers, contigs, sqsum, maxlen, reads1, reads, readsKmers, readsBPs, unique;
    std::vector<int> hasStats;
    std::vector<double> errSum, errSq, cerrSum, cerrSq;
    std::vector<std::string> desc;
    std::vector<jshort> maxCounts; int maxCountsN = 0;  // (nValues+1) x n ; row nValues = total
    std::vector<int> maxCountsHas;
    // per read
    std::vector<int> rFound, rClass, rReadKmers, rTaxErr, rAccepted, rSize;
    std::vector<jint> labels; std::vector<jlong> labelPos;
    std::vector<uint8_t> accept;
};

typedef struct {
    int k, classify, max_paths;
    double max_read_tax_err, max_read_class_err;
    int threshold, max_kmer_res_counts, count_unique, write_all, write_kraken, write_filtered, with_probs,
        initial_read_size, use_filter, dump_labels;
} gso_match_cfg;

static void scanGenome(int k, const std::string& seq, const std::function<void(jlong)>& f) {
    // AbstractStoreFastaReader.java:88-120 with stepSize=1, maxDust=-1, lowerCaseBases=true;
    // CGATLongBuffer.put (C/util/CGATLongBuffer.java): a non-CGAT byte resets the window; '>' lines reset it too.
    jlong kmer = 0, rev = 0; int filled = 0;
    size_t i = 0, n = seq.size();
    while (i < n) {
        if (seq[i] == '>') { while (i < n && seq[i] != '\n') i++; filled = 0; kmer = rev = 0; continue; }
        char ch = seq[i++];
        if (ch == '\n' || ch == '\r') continue;
        uint8_t c = cgat::cgatToUpperCase((uint8_t)ch);
        int bp = cgat::jump(c);
        if (bp == -1) { filled = 0; kmer = rev = 0; continue; }
        kmer = (jshl(kmer, 2) & cgat::shiftFilterStraight(k)) | (jlong)bp;
        rev = jushr(rev, 2) | jshl((jlong)cgat::revjump(c), cgat::shiftFilterReverse(k));
        if (filled < k) filled++;
        if (filled == k) f(cgat::standardKMer(kmer, rev));
    }
}

extern "C" {

// ------------------------------------------------------------------ primitives for the KATs
void gso_random_longs(int64_t seed, int n, int64_t* out) { JavaRandom r(seed); for (int i = 0; i < n; i++) out[i] = r.nextLong(); }
void gso_random_ints(int64_t seed, int bound, int n, int32_t* out) { JavaRandom r(seed); for (int i = 0; i < n; i++) out[i] = r.nextInt(bound); }
int64_t gso_kmer_straight(const uint8_t* s, int start, int k, int* bad) { return cgat::kMerToLongStraight(s, start, k, bad); }
int64_t gso_kmer_reverse(const uint8_t* s, int start, int k, int* bad) { return cgat::kMerToLongReverse(s, start, k, bad); }
int64_t gso_kmer_canonical(const uint8_t* s, int start, int k) { return cgat::kMerToLong(s, start, k, nullptr); }
int64_t gso_next_straight(int64_t kmer, uint8_t bp, int k) { return cgat::nextKMerStraight(kmer, bp, k); }
int64_t gso_next_reverse(int64_t kmer, uint8_t bp, int k) { return cgat::nextKMerReverse(kmer, bp, k); }
int64_t gso_murmur64(int64_t data, int64_t seed) { return murmurHash64(data, seed); }
int gso_java_double_to_string(double v, char* buf, int cap) {
    std::string s = javaDoubleToString(v);
    if ((int)s.size() + 1 > cap) return -1;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}
// canonical k-mers of all windows of a sequence (invalid windows -> -1)
void gso_all_kmers(const uint8_t* s, int n, int k, int64_t* out) {
    for (int i = 0; i + k <= n; i++) out[i] = cgat::kMerToLong(s, i, k, nullptr) , out[i] = (cgat::kMerToLongStraight(s, i, k, nullptr) == -1) ? -1 : out[i];
}

// ------------------------------------------------------------------ standalone Bloom filters
void* gso_bloom_new(int kind, double fpp) {
    if (kind == 0) return new BlockedKMerBloomFilter();
    return new HashedKMerBloomFilter(fpp, kind == 1);
}
// A hashed filter as LoadIndexGoal would deserialize it (C/goals/LoadIndexGoal.java:92-104): the serialized bit count, hash
// factors and bit vector words are authoritative (AbstractKMerBloomFilter.java:104-110 regenerates factors only on growth).
void* gso_bloom_from_words(int kind, int64_t bits, int hashes, const int64_t* factors, const int64_t* words, int64_t nWords) {
    auto* f = new HashedKMerBloomFilter(1e-8, kind == 1);
    f->bits = bits;
    f->hashes = hashes;
    f->hashFactors.assign(factors, factors + hashes);
    f->bitVector.size = nWords;
    f->bitVector.bits.assign(words, words + nWords);
    return f;
}
void gso_bloom_free(void* f) { delete (KMerProbFilter*)f; }
int64_t gso_bloom_ensure(void* f, int64_t n) { return ((KMerProbFilter*)f)->ensureExpectedSize(n, false); }
void gso_bloom_put(void* f, const int64_t* keys, int64_t n) { for (int64_t i = 0; i < n; i++) ((KMerProbFilter*)f)->putLong(keys[i]); }
void gso_bloom_contains(void* f, const int64_t* keys, int64_t n, uint8_t* out) { for (int64_t i = 0; i < n; i++) out[i] = ((KMerProbFilter*)f)->containsLong(keys[i]) ? 1 : 0; }
int gso_bloom_kind(void* f) { return ((KMerProbFilter*)f)->kind(); }
// blocked: p0=seed p1=buckets; hashed: p0=bits p1=hashes
void gso_bloom_params(void* f, int64_t* p0, int64_t* p1, int64_t* nwords) {
    KMerProbFilter* kf = (KMerProbFilter*)f;
    if (kf->kind() == 0) { auto* b = (BlockedKMerBloomFilter*)kf; *p0 = b->seed; *p1 = b->buckets; *nwords = (int64_t)b->data.size(); }
    else { auto* h = (HashedKMerBloomFilter*)kf; *p0 = h->bits; *p1 = h->hashes; *nwords = (int64_t)h->bitVector.bits.size(); }
}
const int64_t* gso_bloom_words(void* f) {
    KMerProbFilter* kf = (KMerProbFilter*)f;
    if (kf->kind() == 0) return ((BlockedKMerBloomFilter*)kf)->data.data();
    return ((HashedKMerBloomFilter*)kf)->bitVector.bits.data();
}
const int64_t* gso_bloom_factors(void* f) {
    KMerProbFilter* kf = (KMerProbFilter*)f;
    return kf->kind() == 0 ? nullptr : ((HashedKMerBloomFilter*)kf)->hashFactors.data();
}

// ------------------------------------------------------------------ DB builder
gso_db* gso_db_new(int k, const char* nodes, size_t nodesLen, const char* names, size_t namesLen,
                   int useRadix, int radixBits, double optFpp) {
    gso_db* d = new gso_db();
    d->k = k; d->useRadix = useRadix != 0; d->radixBits = radixBits; d->optFpp = optFpp;
    d->full.readNodes(std::string(nodes, nodesLen));
    d->full.readNames(std::string(names, namesLen));
    return d;
}
void gso_db_free(gso_db* d) { delete d; }
void gso_db_set_skip_update(gso_db* d, int skip) { d->skipUpdate = skip != 0; }
const char* gso_db_error(gso_db* d) { return d->error.c_str(); }
int gso_db_request(gso_db* d, const char* taxid) {  // taxids.txt entry: requested node; required up to the root
    TaxNode* n = d->full.getNodeByTaxId(taxid);
    if (!n) return -1;
    n->markRequired();
    d->requested.push_back(taxid);
    return 0;
}
int gso_db_add_genome(gso_db* d, const char* taxid, const uint8_t* seq, size_t len, int fill) {
    TaxNode* n = d->full.getNodeByTaxId(taxid);
    if (!n) return -1;
    n->markRequired();
    d->genomes.push_back(Genome{taxid, std::string((const char*)seq, len), fill != 0});
    return 0;
}
static void attachStore(gso_db* d, std::unique_ptr<KMerStoreBase> store) {
    d->db.store = std::move(store);
    d->db.taxTree = d->full.toSmallTaxTree();
    for (const std::string& r : d->requested) { TaxNode* n = d->db.taxTree->getNodeByTaxId(r); if (n) n->requested = true; }
    d->db.initStoreIndices();
}
int gso_db_finalize(gso_db* d, double indexFpp, int indexXor) {
    try {
        int k = d->k;
        // ---- fill: first writer wins (FillDBGoal.java:281-295 -> KMerSortedArray.putLong :168-202)
        std::unordered_map<jlong, int> first;
        std::vector<std::pair<jlong, std::string>> order;
        for (const Genome& g : d->genomes) {
            if (!g.fill) continue;
            scanGenome(k, g.seq, [&](jlong km) {
                if (first.emplace(km, 0).second) order.emplace_back(km, g.taxid);
            });
        }
        std::unique_ptr<KMerStoreBase> store;
        if (d->useRadix) {
            std::vector<int> sizes((size_t)1 << d->radixBits, 0);
            for (auto& kv : order) sizes[(size_t)RadixKMerStore::radixOf(kv.first, d->radixBits)]++;
            auto* rs = new RadixKMerStore(k, d->radixBits, sizes, 1e-11, d->optFpp, true);
            rs->useFilter = false;  // exact dedup already done (documented deviation)
            store.reset(rs);
            for (auto& kv : order) rs->putLong(kv.first, kv.second);
            rs->useFilter = true;
        } else {
            auto* sa = new KMerSortedArray(k, 1e-11, d->optFpp, true);
            sa->size = (jlong)order.size();
            sa->kmers.assign(order.size(), 0); sa->valueIndexes.assign(order.size(), 0);
            for (auto& kv : order) {  // putLong minus the probabilistic duplicate check
                jlong pos = sa->entries++;
                sa->kmers[(size_t)pos] = kv.first;
                sa->setIndexAtPosition(pos, sa->getAddValueIndex(kv.second));
            }
            store.reset(sa);
        }
        store->optimize();
        attachStore(d, std::move(store));
        // ---- filter index over the k-mers of requested nodes, on the FILLED store (BloomIndexGoal.java:66-111)
        if (indexFpp > 0) {
            jlong counter = 0;
            KMerStoreBase* st = d->db.store.get();
            auto isReq = [&](int vidx) { TaxNode* n = d->db.taxTree->getNodeByTaxId(st->indexMap[(size_t)vidx]); return n && n->requested; };
            st->visit([&](jlong, int vidx, jlong) { if (isReq(vidx)) counter++; });
            if (indexFpp == 0.01) d->index.reset(new BlockedKMerBloomFilter());
            else d->index.reset(new HashedKMerBloomFilter(indexFpp, indexXor != 0));
            d->index->ensureExpectedSize(counter, false);
            st->visit([&](jlong km, int vidx, jlong) { if (isReq(vidx)) d->index->putLong(km); });
        }
        // ---- update: value = LCA(value, genomeNode) for every genome (DBGoal.java:234-253, 300-311)
        KMerStoreBase* st = d->db.store.get();
        for (const Genome& g : d->genomes) {
            if (d->skipUpdate) break;
            TaxNode* node = d->full.getNodeByTaxId(g.taxid);
            scanGenome(k, g.seq, [&](jlong km) {
                jlong pos;
                int vi = st->getLong(km, &pos);
                if (vi < 0) return;
                TaxNode* oldNode = d->full.getNodeByTaxId(st->indexMap[(size_t)vi]);
                TaxNode* lca = TaxTree::getLowestCommonAncestor(oldNode, node);
                const std::string& nv = lca ? lca->taxId : st->indexMap[(size_t)vi];
                int ni = st->getAddValueIndex(nv);
                if (ni != vi) st->setIndexAtPosition(pos, ni);
            });
        }
        d->db.fix();
        d->db.convert();
        return 0;
    } catch (const std::exception& e) { d->error = e.what(); return -1; }
}
// DB straight from arrays (bench-scale spot checks): sorted distinct keys, raw Java shorts, tree by value index.
// Node taxid = decimal string of the value index + 1 (root must be value index 0 -> "1").
gso_db* gso_db_from_arrays(int k, const int64_t* keys, const int16_t* vals, int64_t n, int nValues,
                           const int32_t* parentByVidx, int buildBloom) {
    gso_db* d = new gso_db();
    d->k = k;
    std::string nodes;
    // children appear in value index order => pre-order positions need a proper tree; emit parent lines in index order
    for (int v = 0; v < nValues; v++) {
        int p = parentByVidx[v] < 0 ? v : parentByVidx[v];
        nodes += std::to_string(v + 1) + "\t|\t" + std::to_string(p + 1) + "\t|\tno rank\t|\t\t|\n";
    }
    d->full.readNodes(nodes);
    for (auto& kv : d->full.byId) kv.second->required = true;
    auto* sa = new KMerSortedArray(k, 1e-11, buildBloom ? 0.01 : 1.0, true);
    sa->size = sa->entries = n;
    sa->kmers.assign(keys, keys + n);
    sa->valueIndexes.assign(vals, vals + n);
    for (int v = 0; v < nValues; v++) sa->getAddValueIndex(std::to_string(v + 1));
    sa->sorted = true;
    sa->filter = sa->createOptimizedFilter();
    if (sa->filter) {
        // the store's filter over all keys (KMerSortedArray.java:409-422); several threads for the 1e8..2e9-key bench databases
        auto* bf = dynamic_cast<BlockedKMerBloomFilter*>(sa->filter.get());
        const int T = bf ? (int)std::max(1u, std::min(64u, std::thread::hardware_concurrency())) : 1;
        if (T > 1 && n > (1 << 20)) {
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++)
                th.emplace_back([=] { for (int64_t i = n * t / T, e = n * (t + 1) / T; i < e; i++) bf->putLongShared(keys[i]); });
            for (auto& x : th) x.join();
            bf->entries = n;
        } else {
            for (int64_t i = 0; i < n; i++) sa->filter->putLong(keys[i]);
        }
    }
    attachStore(d, std::unique_ptr<KMerStoreBase>(sa));
    d->db.fix();
    d->db.convert();
    return d;
}

int gso_db_k(gso_db* d) { return d->k; }
int64_t gso_db_n_kmers(gso_db* d) { return d->db.store->entries; }
int gso_db_n_values(gso_db* d) { return d->db.store->getNValues(); }
const char* gso_db_value_taxid(gso_db* d, int vidx) { return d->db.store->indexMap[(size_t)vidx].c_str(); }
// flattened (kmer, raw short value, pos) in storage-position order
void gso_db_export(gso_db* d, int64_t* keys, int16_t* vals) {
    d->db.store->visit([&](jlong km, int vidx, jlong pos) { keys[pos] = km; vals[pos] = (int16_t)(vidx + INT16_MIN); });
}
// by value index: parent value index (-1 root / no node), depth, pre-order position, has-node flag
void gso_db_tree(gso_db* d, int32_t* parent, int32_t* depth, int32_t* position, int32_t* hasNode) {
    int nv = d->db.store->getNValues();
    for (int v = 0; v < nv; v++) {
        TaxNode* n = d->db.nodeByValueIndex[(size_t)v];
        hasNode[v] = n ? 1 : 0;
        parent[v] = (n && n->parent) ? n->parent->storeIndex : -1;
        depth[v] = n ? n->depth : 0;
        position[v] = n ? n->position : -1;
    }
}
void gso_db_dbkmers(gso_db* d, int64_t* out) { for (size_t i = 0; i < d->db.dbKmersPerValueIndex.size(); i++) out[i] = d->db.dbKmersPerValueIndex[i]; }
void* gso_db_store_filter(gso_db* d) { return d->db.store->filter.get(); }
void* gso_db_index_filter(gso_db* d) { return d->index.get(); }
void gso_db_set_use_filter(gso_db* d, int use) { d->db.store->useFilter = use != 0; }
// single lookups (cross-store test): returns value index or -1
int gso_db_get(gso_db* d, int64_t kmer, int64_t* pos) { return d->db.store->getLong(kmer, pos); }
// node name / rank / taxid metadata for the product's CSV writer (by value index)
const char* gso_db_node_name(gso_db* d, int vidx) { TaxNode* n = d->db.nodeByValueIndex[(size_t)vidx]; return n ? n->name.c_str() : ""; }
int gso_db_node_rank(gso_db* d, int vidx) { TaxNode* n = d->db.nodeByValueIndex[(size_t)vidx]; return n ? n->rank : -1; }
int gso_db_node_requested(gso_db* d, int vidx) { TaxNode* n = d->db.nodeByValueIndex[(size_t)vidx]; return n && n->requested ? 1 : 0; }

// ------------------------------------------------------------------ match / filter runs
static MatchConfig toCfg(const gso_match_cfg* c) {
    MatchConfig m;
    m.k = c->k; m.classify = c->classify != 0; m.maxPaths = c->max_paths;
    m.maxReadTaxErrorCount = c->max_read_tax_err; m.maxReadClassErrorCount = c->max_read_class_err;
    m.threshold = c->threshold; m.maxKmerResCounts = c->max_kmer_res_counts; m.countUnique = c->count_unique != 0;
    m.writeAll = c->write_all != 0; m.writeKraken = c->write_kraken != 0; m.writeFiltered = c->write_filtered != 0;
    m.withProbs = c->with_probs != 0; m.initialReadSize = c->initial_read_size;
    return m;
}

// FastqKMerMatcher.runMatcher (C/match/FastqKMerMatcher.java:181-235) + MatchResultGoal (C/goals/MatchResultGoal.java:91-164)
// over n in-memory files (each FASTQ or FASTA), threads=0.
gso_run* gso_match_files(gso_db* d, const gso_match_cfg* c, const uint8_t* const* files, const size_t* lens,
                         const int* isFasta, int nFiles) {
    MatchConfig cfg = toCfg(c);
    d->db.store->useFilter = c->use_filter != 0;
    gso_run* r = new gso_run();
    FastqKMerMatcher m(&d->db, cfg);
    std::unique_ptr<KMerUniqueCounterBits> uc;
    if (cfg.countUnique) { uc.reset(new KMerUniqueCounterBits(d->db.store.get(), cfg.maxKmerResCounts > 0)); uc->clear(); }
    m.initStats();
    m.uniqueCounter = uc.get();
    if (cfg.writeKraken) m.krakenOut = &r->kraken;
    if (cfg.writeFiltered) m.filteredOut = &r->filtered;
    if (c->dump_labels) { m.dumpLabels = &r->labels; m.dumpPos = &r->labelPos; }
    ReadEntry e1, e2;
    e1.init(cfg.initialReadSize, cfg.withProbs, cfg.maxPaths);
    e2.init(cfg.initialReadSize, cfg.withProbs, cfg.maxPaths);
    FastqReader fr; fr.k = cfg.k;
    fr.nextEntry = [&](ReadEntry& e) {
        m.nextEntry(e);
        bool found = false;  // recompute `found` as returned by matchRead: filtered write happened iff found
        (void)found;
        r->rClass.push_back(e.classNode ? e.classNode->storeIndex : -1);
        r->rReadKmers.push_back(e.outReadKmers);
        r->rTaxErr.push_back(e.outReadTaxErrorCount);
        r->rAccepted.push_back(e.outAccepted ? 1 : 0);
        r->rSize.push_back(e.readSize);
    };
    for (int f = 0; f < nFiles; f++) {
        m.beginFile();
        if (isFasta[f]) fr.readFasta(files[f], lens[f], e1, e2); else fr.readFastq(files[f], lens[f], e1);
        // AbstractLoggingFastqStreamer.processFastqStreams (C/fastq/AbstractLoggingFastqStreamer.java:95-131): totals
        r->totalReads += fr.reads; r->totalKMers += fr.kMers; r->totalBPs += fr.readBPs;
    }
    std::vector<jlong> uniq;
    std::map<int, std::vector<jshort>> mc;
    if (uc) { uniq = uc->getUniqueKmerCounts(); if (uc->withCounts) mc = uc->getMaxCountsCounts(cfg.maxKmerResCounts); }
    MatchingResult res = completeResults(m, r->totalReads, r->totalKMers, r->totalBPs, uc ? &uniq : nullptr,
                                         (uc && uc->withCounts) ? &mc : nullptr);
    r->csv = printMatchResult(res);
    int nv = d->db.store->getNValues();
    r->nValues = nv;
    auto z = [&](std::vector<jlong>& v) { v.assign((size_t)nv, 0); };
    z(r->kmers); z(r->contigs); z(r->sqsum); z(r->maxlen); z(r->reads1); z(r->reads); z(r->readsKmers); z(r->readsBPs); z(r->unique);
    r->hasStats.assign((size_t)nv, 0);
    r->errSum.assign((size_t)nv, 0); r->errSq.assign((size_t)nv, 0); r->cerrSum.assign((size_t)nv, 0); r->cerrSq.assign((size_t)nv, 0);
    r->desc.assign((size_t)nv, "");
    for (int v = 0; v < nv; v++) {
        if (uc) r->unique[(size_t)v] = uniq[(size_t)v];
        CountsPerTaxid* s = m.statsIndex[(size_t)v].get();
        if (!s) continue;
        r->hasStats[(size_t)v] = 1;
        r->kmers[(size_t)v] = s->kmers; r->contigs[(size_t)v] = s->contigs; r->sqsum[(size_t)v] = s->contigLenSquaredSum;
        r->maxlen[(size_t)v] = s->maxContigLen; r->reads1[(size_t)v] = s->reads1KMer; r->reads[(size_t)v] = s->reads;
        r->readsKmers[(size_t)v] = s->readsKmers; r->readsBPs[(size_t)v] = s->readsBPs;
        r->errSum[(size_t)v] = s->errorSum; r->errSq[(size_t)v] = s->errorSquaredSum;
        r->cerrSum[(size_t)v] = s->classErrorSum; r->cerrSq[(size_t)v] = s->classErrorSquaredSum;
        r->desc[(size_t)v] = s->maxContigDescriptor;
    }
    if (uc && uc->withCounts) {
        int n = cfg.maxKmerResCounts;
        r->maxCountsN = n;
        r->maxCounts.assign((size_t)(nv + 1) * (size_t)n, 0);
        r->maxCountsHas.assign((size_t)nv + 1, 0);
        for (auto& kv : mc) {
            size_t row = kv.first < 0 ? (size_t)nv : (size_t)kv.first;
            r->maxCountsHas[row] = 1;
            for (int i = 0; i < n; i++) r->maxCounts[row * (size_t)n + (size_t)i] = kv.second[(size_t)i];
        }
    }
    return r;
}

// FastqBloomFilter.runFilter (C/bloom/FastqBloomFilter.java:80-105) over in-memory FASTQ files, threads=0.
gso_run* gso_filter_files(void* filter, int k, int minPosCount, double posRatio, int withProbs, int initialReadSize,
                          const uint8_t* const* files, const size_t* lens, const int* isFasta, int nFiles) {
    gso_run* r = new gso_run();
    KMerProbFilter* f = (KMerProbFilter*)filter;
    ReadEntry e1, e2;
    e1.init(initialReadSize, withProbs != 0, 1);
    e2.init(initialReadSize, withProbs != 0, 1);
    FastqReader fr; fr.k = k;
    fr.nextEntry = [&](ReadEntry& e) {
        bool acc = isAcceptRead(*f, k, minPosCount, posRatio, e.read.data(), e.readSize);
        r->accept.push_back(acc ? 1 : 0);
        r->rSize.push_back(e.readSize);
        if (acc) e.write(r->filtered); else e.write(r->rest);
    };
    for (int i = 0; i < nFiles; i++) {
        if (isFasta[i]) fr.readFasta(files[i], lens[i], e1, e2); else fr.readFastq(files[i], lens[i], e1);
        r->totalReads += fr.reads; r->totalKMers += fr.kMers; r->totalBPs += fr.readBPs;
    }
    return r;
}

void gso_run_free(gso_run* r) { delete r; }
const char* gso_run_text(gso_run* r, int which, size_t* len) {
    std::string* s = which == 0 ? &r->csv : which == 1 ? &r->filtered : which == 2 ? &r->kraken : &r->rest;
    *len = s->size();
    return s->data();
}
void gso_run_totals(gso_run* r, int64_t* out) { out[0] = r->totalReads; out[1] = r->totalKMers; out[2] = r->totalBPs; }
int64_t gso_run_n_reads(gso_run* r) { return (int64_t)std::max(r->rSize.size(), r->accept.size()); }
// which: 0 kmers 1 contigs 2 sqsum 3 maxlen 4 reads1 5 reads 6 readsKmers 7 readsBPs 8 unique
const int64_t* gso_run_stat(gso_run* r, int which) {
    std::vector<jlong>* v[] = {&r->kmers, &r->contigs, &r->sqsum, &r->maxlen, &r->reads1, &r->reads, &r->readsKmers, &r->readsBPs, &r->unique};
    return v[which]->data();
}
const int* gso_run_has_stats(gso_run* r) { return r->hasStats.data(); }
const double* gso_run_dstat(gso_run* r, int which) {
    std::vector<double>* v[] = {&r->errSum, &r->errSq, &r->cerrSum, &r->cerrSq};
    return v[which]->data();
}
const char* gso_run_desc(gso_run* r, int vidx) { return r->desc[(size_t)vidx].c_str(); }
// which: 0 class vidx 1 readKmers 2 taxErr 3 accepted 4 readSize
const int* gso_run_read(gso_run* r, int which) {
    std::vector<int>* v[] = {&r->rClass, &r->rReadKmers, &r->rTaxErr, &r->rAccepted, &r->rSize};
    return v[which]->data();
}
const uint8_t* gso_run_accept(gso_run* r) { return r->accept.data(); }
int64_t gso_run_n_labels(gso_run* r) { return (int64_t)r->labels.size(); }
const int32_t* gso_run_labels(gso_run* r) { return r->labels.data(); }
const int64_t* gso_run_label_pos(gso_run* r) { return r->labelPos.data(); }
int gso_run_max_counts(gso_run* r, const int16_t** counts, const int** has) { *counts = r->maxCounts.data(); *has = r->maxCountsHas.data(); return r->maxCountsN; }

// ------------------------------------------------------------------ multi-threaded CPU baseline
// The reference's threading model (C/fastq/AbstractFastqReader.java:85-185): reads are independent work
// units handed to `threads` consumers; per-consumer vote slots (SmallTaxTree.java:643-658) and
// readNoPerCPerStat rows (FastqKMerMatcher.java:78); shared stats.  This port gives every consumer its
// own stats object (no locks: generous to the CPU side) and a shared bitset set with atomic OR.
// Reads are pre-parsed (bases + offsets) so the timing covers the matchRead hot loop only, as the GPU
// kernel-only number does.  Returns the number of k-mers processed.
int64_t gso_match_reads_mt(gso_db* d, const gso_match_cfg* c, const uint8_t* bases, const uint64_t* offsets,
                           int64_t nReads, int threads, int64_t* kmersPerVidxOut) {
    MatchConfig cfg = toCfg(c);
    d->db.store->useFilter = c->use_filter != 0;
    if (threads < 1) threads = 1;
    TaxTree* tree = d->db.taxTree.get();
    tree->initCountSize(threads);
    tree->preorder([&](TaxNode* n) { n->counts.assign((size_t)threads, 0); n->countsInitKeys.assign((size_t)threads, -1); });
    std::vector<uint64_t> bitset(cfg.countUnique ? (size_t)((d->db.store->entries + 63) / 64) : 0, 0);
    std::vector<std::unique_ptr<FastqKMerMatcher>> ms;
    for (int t = 0; t < threads; t++) { ms.emplace_back(new FastqKMerMatcher(&d->db, cfg)); ms.back()->taxTree = cfg.classify ? tree : nullptr; }
    std::atomic<int64_t> next(0);
    std::atomic<int64_t> total(0);
    auto worker = [&](int t) {
        FastqKMerMatcher& m = *ms[(size_t)t];
        m.sharedBits = bitset.empty() ? nullptr : bitset.data();
        ReadEntry e; e.init(256, false, cfg.maxPaths);
        int64_t mine = 0;
        const int64_t CH = 1024;
        for (;;) {
            int64_t b = next.fetch_add(CH);
            if (b >= nReads) break;
            int64_t en = std::min(nReads, b + CH);
            for (int64_t i = b; i < en; i++) {
                size_t len = (size_t)(offsets[i + 1] - offsets[i]);
                if (e.read.size() < len + 1) e.read.resize(len + 1);
                std::memcpy(e.read.data(), bases + offsets[i], len);
                e.readSize = (int)len; e.readNo = i; e.readDescriptorSize = 0;
                e.usedPaths = 0; e.classNode = nullptr;
                for (int p = 0; p < cfg.maxPaths; p++) { e.readTaxIdNode[(size_t)p] = nullptr; e.counts[(size_t)p] = 0; }
                m.matchRead(e, t);
                if ((int)len >= cfg.k) mine += (int64_t)len - cfg.k + 1;
            }
        }
        total += mine;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(worker, t);
    worker(0);
    for (auto& x : th) x.join();
    if (kmersPerVidxOut) {
        int nv = d->db.store->getNValues();
        for (int v = 0; v < nv; v++) {
            int64_t s = 0;
            for (auto& m : ms) if (m->statsIndex[(size_t)v]) s += m->statsIndex[(size_t)v]->kmers;
            kmersPerVidxOut[v] = s;
        }
    }
    return total.load();
}

// FastqBloomFilter.isAcceptRead (C/bloom/FastqBloomFilter.java:120-161) over pre-parsed reads, `threads` consumer threads
// (the reference's consumers each run isAcceptRead on whole reads).  accept[i] = 0/1.  Returns the number of k-mers of the reads.
int64_t gso_filter_reads_mt(void* filter, int k, int minPosCount, double posRatio, const uint8_t* bases, const uint64_t* offsets,
                            int64_t nReads, int threads, uint8_t* accept) {
    KMerProbFilter* f = (KMerProbFilter*)filter;
    if (threads < 1) threads = 1;
    std::atomic<int64_t> next(0), total(0);
    auto worker = [&]() {
        int64_t mine = 0;
        const int64_t CH = 1024;
        for (;;) {
            int64_t b = next.fetch_add(CH);
            if (b >= nReads) break;
            int64_t en = std::min(nReads, b + CH);
            for (int64_t i = b; i < en; i++) {
                const int len = (int)(offsets[i + 1] - offsets[i]);
                accept[i] = isAcceptRead(*f, k, minPosCount, posRatio, bases + offsets[i], len) ? 1 : 0;
                if (len >= k) mine += len - k + 1;
            }
        }
        total += mine;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(worker);
    worker();
    for (auto& x : th) x.join();
    return total.load();
}

}  // extern "C"
