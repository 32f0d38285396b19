// gs_oracle_kat.cpp -- the reference's own known-answer / property tests restated against the CPU
// ORACLE (test infrastructure).  Each function returns the number of failed assertions and appends a
// human-readable reason to `msg`; tests/test_oracle_golden.py drives them.  Citations: T/ =
// core/src/test/java/org/metagene/genestrip/, R/ = core/src/test/resources/.
#include "gs_oracle.hpp"

#include <deque>
#include <set>
#include <unordered_set>

using namespace gso;

static thread_local std::string g_msg;
#define CHECK(cond, what) do { if (!(cond)) { fails++; if (g_msg.size() < 4000) { g_msg += what; g_msg += "; "; } } } while (0)

static jlong fullStraight(const std::deque<uint8_t>& b) { jlong r = 0; for (uint8_t c : b) { r = jrotl(r, 2); r += cgat::jump(c); } return r; }
static jlong fullReverse(const std::deque<uint8_t>& b) { jlong r = 0; for (size_t i = b.size(); i-- > 0;) { r = jrotl(r, 2); r += cgat::revjump(b[i]); } return r; }

// SmallTaxIdNode stand-ins for stores whose values are plain strings
static Database makeDbNoTree(std::unique_ptr<KMerStoreBase> store, std::vector<std::unique_ptr<TaxNode>>& nodes) {
    Database db;
    db.store = std::move(store);
    db.taxTree.reset(new TaxTree());
    db.nodeByValueIndex.assign((size_t)db.store->getNValues(), nullptr);
    for (int i = 0; i < db.store->getNValues(); i++) {
        nodes.emplace_back(new TaxNode());
        nodes.back()->taxId = db.store->indexMap[(size_t)i];
        nodes.back()->storeIndex = i;
        db.nodeByValueIndex[(size_t)i] = nodes.back().get();
    }
    return db;
}

static std::unique_ptr<KMerStoreBase> newStore(int type, int k, const std::vector<jlong>& distinct, const std::vector<std::string>& initialValues) {
    // T/match/FastqKMerMatcherTest.java:65-85 (SORTED_SMALL / SORTED_LARGE share the arithmetic; RADIX)
    std::unique_ptr<KMerStoreBase> s;
    if (type == 2) {
        std::vector<int> sizes((size_t)1 << 17, 0);
        for (jlong km : distinct) sizes[(size_t)RadixKMerStore::radixOf(km, 17)]++;
        s.reset(new RadixKMerStore(k, 17, sizes, 0.0001, 0.0001, true));
    } else {
        auto* a = new KMerSortedArray(k, 0.0001, 0.0001, true);
        a->initSize((jlong)distinct.size());
        s.reset(a);
    }
    for (const std::string& v : initialValues) s->getAddValueIndex(v);
    return s;
}

extern "C" {

const char* gso_kat_message() { return g_msg.c_str(); }

// java.util.Random check values (C/bloom/BlockedKMerBloomFilter.java:91-93)
int gso_kat_random() {
    int fails = 0; g_msg.clear();
    JavaRandom r(42);
    CHECK(r.nextLong() == -5025562857975149833LL, "Random(42).nextLong #1");
    CHECK(r.nextLong() == -5843495416241995736LL, "Random(42).nextLong #2");
    CHECK(r.nextLong() == 5694868678511409995LL, "Random(42).nextLong #3");
    JavaRandom q(42);
    // the widely published output of `new Random(42)` + ten `nextInt(10)` calls on any JDK
    int exp[10] = {0, 3, 8, 4, 0, 5, 5, 8, 9, 3};
    for (int i = 0; i < 10; i++) CHECK(q.nextInt(10) == exp[i], "Random(42).nextInt(10)");
    return fails;
}

// T/util/NextKMerTest.java:37-87
int gso_kat_next_kmer(int nBases) {
    int fails = 0; g_msg.clear();
    uint8_t buf[32];
    for (int k = 1; k < 31; k++) {
        for (int pass = 0; pass < 2; pass++) {
            std::deque<uint8_t> ring;
            JavaRandom random(10);
            jlong old = -1;
            for (int j = 0; j < nBases; j++) {
                uint8_t c = (uint8_t)cgat::DECODE_TABLE[random.nextInt(4)];
                ring.push_back(c);
                if ((int)ring.size() > k) ring.pop_front();
                if ((int)ring.size() == k) {
                    jlong kmer = pass == 0 ? fullStraight(ring) : fullReverse(ring);
                    if (old != -1) CHECK(kmer == (pass == 0 ? cgat::nextKMerStraight(old, c, k) : cgat::nextKMerReverse(old, c, k)), "rolling != full");
                    old = kmer;
                    cgat::longToKMerStraight(kmer, buf, 0, k);
                    if (pass == 1) cgat::reverse(buf, 0, k);
                    bool same = true;
                    for (int i = 0; i < k; i++) same = same && buf[i] == ring[(size_t)i];
                    CHECK(same, "decode round trip");
                    // and the array encoders agree with the ring encoders
                    std::vector<uint8_t> arr(ring.begin(), ring.end());
                    CHECK(cgat::kMerToLongStraight(arr.data(), 0, k, nullptr) == fullStraight(ring), "kMerToLongStraight");
                    CHECK(cgat::kMerToLongReverse(arr.data(), 0, k, nullptr) == fullReverse(ring), "kMerToLongReverse");
                }
            }
        }
    }
    return fails;
}

// T/bloom/KMerBloomFilterTest.java:47-130 (kind 0 blocked [fpp 0.01 default sizing], 1 xor, 2 murmur; fpp 0.0001)
int gso_kat_bloom(int kind, int size) {
    int fails = 0; g_msg.clear();
    const int k = 31;
    JavaRandom random(42);
    std::unique_ptr<KMerProbFilter> filter;
    double fpp = kind == 0 ? 0.01 : 0.0001;
    if (kind == 0) filter.reset(new BlockedKMerBloomFilter()); else filter.reset(new HashedKMerBloomFilter(fpp, kind == 1));
    filter->ensureExpectedSize(size, false);
    std::vector<std::vector<uint8_t>> reads;
    std::vector<uint8_t> read((size_t)k), rev((size_t)k);
    for (int i = 1; i < size; i++) {
        for (int j = 0; j < k; j++) { read[(size_t)j] = (uint8_t)cgat::DECODE_TABLE[random.nextInt(4)]; rev[(size_t)(k - j - 1)] = (uint8_t)cgat::complement(read[(size_t)j]); }
        reads.push_back(read);
        filter->putLong(cgat::kMerToLong(read.data(), 0, k, nullptr));
        CHECK(filter->containsLong(cgat::kMerToLong(read.data(), 0, k, nullptr)), "false negative");
        CHECK(filter->containsLong(cgat::kMerToLong(rev.data(), 0, k, nullptr)), "false negative (reverse complement)");
    }
    for (auto& rd : reads) {
        for (int j = 0; j < k; j++) rev[(size_t)(k - j - 1)] = (uint8_t)cgat::complement(rd[(size_t)j]);
        CHECK(filter->containsLong(cgat::kMerToLong(rd.data(), 0, k, nullptr)), "false negative (2nd pass)");
        CHECK(filter->containsLong(cgat::kMerToLong(rev.data(), 0, k, nullptr)), "false negative rc (2nd pass)");
    }
    int err = 0;
    for (int i = 1; i < size; i++) {
        for (int j = 0; j < k; j++) read[(size_t)j] = (uint8_t)cgat::DECODE_TABLE[random.nextInt(4)];
        if (filter->containsLong(cgat::kMerToLong(read.data(), 0, k, nullptr))) err++;
    }
    double testedFp = ((double)err) / (2.0 * size);
    CHECK(testedFp <= fpp * 1.1, "fpp above 1.1 x target: " + std::to_string(testedFp));
    return fails;
}

// T/bloom/XORKMerBloomFilterTest.java:50-58
int gso_kat_xor_min_value() {
    int fails = 0; g_msg.clear();
    HashedKMerBloomFilter f(0.0001, true);
    f.ensureExpectedSize(1000, false);
    CHECK(f.hashFactors[0] == -5025562857975149833LL, "hashFactors[0] = Random(42).nextLong()");
    jlong x = f.hashFactors[0] ^ INT64_MIN;
    jlong idx = f.reduce(f.hash(x, 0));
    CHECK(idx >= 0 && idx < f.bits, "bit index in range for hash == Long.MIN_VALUE");
    f.putLong(x);
    CHECK(f.containsLong(x), "contains after put");
    // sizing (C/bloom/AbstractKMerBloomFilter.java:167-180): p=1e-8 -> 38.34 bits/key, 27 hashes
    HashedKMerBloomFilter g(1e-8, true);
    g.ensureExpectedSize(100000000LL, false);
    CHECK(g.bits == 3834023350LL, "bits for n=1e8,p=1e-8: " + std::to_string(g.bits));
    CHECK(g.hashes == 27, "hashes for p=1e-8: " + std::to_string(g.hashes));
    return fails;
}

// T/store/AbstractKMerStoreTest.java:118-261 (put/get/visit; type 0 sorted, 2 radix) with Random(42)
int gso_kat_store(int type, int testSize, int negativeTestSize) {
    int fails = 0; g_msg.clear();
    const int k = 31;
    JavaRandom random(42);
    std::map<std::vector<uint8_t>, int> control;
    std::vector<std::pair<jlong, int>> kmerMap;
    std::unordered_set<jlong> seen;
    std::vector<uint8_t> read((size_t)k);
    while ((int)kmerMap.size() < testSize) {
        for (int j = 0; j < k; j++) read[(size_t)j] = (uint8_t)cgat::DECODE_TABLE[random.nextInt(4)];
        jlong km = cgat::kMerToLong(read.data(), 0, k, nullptr);
        if (!seen.insert(km).second) continue;
        int v = random.nextInt(3);
        control[read] = v;
        kmerMap.emplace_back(km, v);
    }
    std::vector<jlong> distinct;
    for (auto& kv : kmerMap) distinct.push_back(kv.first);
    std::unique_ptr<KMerStoreBase> store = newStore(type, k, distinct, {});
    for (auto& kv : kmerMap) store->putLong(kv.first, "v" + std::to_string(kv.second));
    store->optimize();
    jlong entries = store->entries;
    // the fill-time filter (fpp 1e-4) may reject a few puts as probable duplicates; the reference tolerates that silently
    CHECK(entries > (jlong)(testSize * 0.99), "too many rejected puts");
    int missing = 0;
    for (auto& kv : control) {
        std::vector<uint8_t> r = kv.first;
        jlong pos = -1;
        int vi = store->getLong(cgat::kMerToLong(r.data(), 0, k, nullptr), &pos);
        if (vi < 0) { missing++; continue; }
        CHECK(store->indexMap[(size_t)vi] == "v" + std::to_string(kv.second), "wrong value");
        CHECK(pos >= 0 && pos < entries, "position range");
        cgat::reverse(r.data(), 0, k);
        int vi2 = store->getLong(cgat::kMerToLong(r.data(), 0, k, nullptr), nullptr);
        CHECK(vi2 == vi, "reverse complement lookup");
    }
    CHECK(missing == (int)(testSize - entries), "missing == rejected puts");
    for (int i = 1; i <= negativeTestSize; i++) {
        for (int j = 0; j < k; j++) read[(size_t)j] = (uint8_t)cgat::DECODE_TABLE[random.nextInt(4)];
        if (!control.count(read)) {
            std::vector<uint8_t> rc = read; cgat::reverse(rc.data(), 0, k);
            if (!control.count(rc)) CHECK(store->getLong(cgat::kMerToLong(read.data(), 0, k, nullptr), nullptr) == -1, "negative lookup hit");
        }
    }
    std::set<jlong> positions;
    jlong visited = 0;
    std::unordered_map<jlong, int> remaining(kmerMap.begin(), kmerMap.end());
    store->visit([&](jlong km, int vidx, jlong pos) {
        visited++;
        auto it = remaining.find(km);
        if (it == remaining.end() || store->indexMap[(size_t)vidx] != "v" + std::to_string(it->second)) fails++;
        else remaining.erase(it);
        if (!(pos >= 0 && pos < entries && positions.insert(pos).second)) fails++;
    });
    CHECK(visited == entries && (jlong)positions.size() == entries, "visit covers [0,entries)");
    CHECK((jlong)remaining.size() == testSize - entries, "visit saw every stored k-mer");
    return fails;
}

// T/match/RadixKMerStoreBenchmarkTest.java:186-229: sorted and radix agree on hits and misses, k in {16,21,31}
int gso_kat_cross_store(int n) {
    int fails = 0; g_msg.clear();
    int ks[3] = {16, 21, 31};
    for (int ki = 0; ki < 3; ki++) {
        int k = ks[ki];
        JavaRandom rnd(54321);
        std::unordered_set<jlong> seen;
        std::vector<std::pair<jlong, std::string>> items;
        jlong mask = k == 32 ? -1LL : ((1LL << (2 * k)) - 1);
        while ((int)items.size() < n) {
            jlong km = rnd.nextLong() & mask;
            if (!seen.insert(km).second) continue;
            items.emplace_back(km, "t" + std::to_string(rnd.nextInt(5)));
        }
        std::vector<jlong> distinct;
        for (auto& kv : items) distinct.push_back(kv.first);
        auto a = newStore(0, k, distinct, {});
        auto b = newStore(2, k, distinct, {});
        // bypass the probabilistic duplicate check so both stores hold exactly the same set
        a->useFilter = false; b->useFilter = false;
        auto* sa = (KMerSortedArray*)a.get();
        for (auto& kv : items) { jlong pos = sa->entries++; sa->kmers[(size_t)pos] = kv.first; sa->setIndexAtPosition(pos, sa->getAddValueIndex(kv.second)); b->putLong(kv.first, kv.second); }
        a->optimize(); b->optimize();
        a->useFilter = true; b->useFilter = true;
        JavaRandom q(98765);
        std::set<jlong> posA, posB;
        for (int i = 0; i < n; i++) {
            jlong km = (i % 2 == 0) ? items[(size_t)q.nextInt(n)].first : (q.nextLong() & mask);
            jlong pa = -1, pb = -1;
            int va = a->getLong(km, &pa), vb = b->getLong(km, &pb);
            CHECK((va < 0) == (vb < 0), "hit/miss disagreement");
            if (va >= 0 && vb >= 0) {
                CHECK(a->indexMap[(size_t)va] == b->indexMap[(size_t)vb], "value disagreement");
                CHECK(pa >= 0 && pa < a->entries && pb >= 0 && pb < b->entries, "position range");
            }
        }
    }
    return fails;
}

// T/match/FastqKMerMatcherTest.java:96-210 (testMatchRead) -- store type 0 sorted / 2 radix, optimized
int gso_kat_match_read(int type, int optimize) {
    int fails = 0; g_msg.clear();
    const int readLength = 500, entries = 2000;
    const std::vector<std::string> TAXIDS = {"1", "2", "3"};
    const uint8_t CC[2] = {'C', 'C'}, GG[2] = {'G', 'G'}, TT[2] = {'T', 'T'}, AG[2] = {'A', 'G'};
    jlong cc = cgat::kMerToLong(CC, 0, 2, nullptr), gg = cgat::kMerToLong(GG, 0, 2, nullptr);
    jlong tt = cgat::kMerToLong(TT, 0, 2, nullptr), ag = cgat::kMerToLong(AG, 0, 2, nullptr);
    CHECK(cc == gg, "GG canonicalises to CC");
    std::unique_ptr<KMerStoreBase> store = newStore(type, 2, {cc, tt, ag}, TAXIDS);
    store->putLong(cc, TAXIDS[0]);
    CHECK(!store->putLong(gg, TAXIDS[1]), "duplicate put rejected");
    CHECK(store->putLong(tt, TAXIDS[1]), "TT put");
    store->putLong(ag, TAXIDS[2]);
    if (optimize) store->optimize();
    std::vector<std::unique_ptr<TaxNode>> nodes;
    Database db = makeDbNoTree(std::move(store), nodes);
    MatchConfig cfg; cfg.k = 2; cfg.classify = false; cfg.maxPaths = 4; cfg.maxReadTaxErrorCount = 0; cfg.maxReadClassErrorCount = 0;
    cfg.maxKmerResCounts = 10; cfg.initialReadSize = readLength * 10;
    FastqKMerMatcher matcher(&db, cfg);
    std::string out; matcher.krakenOut = &out;
    KMerUniqueCounterBits uc(db.store.get(), true);
    ReadEntry entry; entry.init(2000, true, 4);
    entry.readSize = readLength;
    JavaRandom random(42);  // the test class's field `random` (T/match/FastqKMerMatcherTest.java:56)
    for (int i = 1; i <= entries; i++) {
        jlong counters[3] = {0, 0, 0}; int contigs[3] = {0, 0, 0}, maxContigLen[3] = {0, 0, 0}; bool used[3] = {false, false, false};
        matcher.initStats();
        uc.clear(); matcher.uniqueCounter = &uc;
        entry.buffer.clear();
        int contigLen = 0, t = -1, lastT = -1;
        uint8_t* read = entry.read.data();
        for (int j = 0; j < readLength; j++) {
            read[j] = (uint8_t)cgat::DECODE_TABLE[random.nextInt(4)];
            if (j > 0) {
                lastT = t;
                if ((read[j - 1] == 'C' && read[j] == 'C') || (read[j - 1] == 'G' && read[j] == 'G')) { counters[0]++; used[0] = true; t = 0; }
                else if ((read[j - 1] == 'A' && read[j] == 'A') || (read[j - 1] == 'T' && read[j] == 'T')) { counters[1]++; used[1] = true; t = 1; }
                else if ((read[j - 1] == 'A' && read[j] == 'G') || (read[j - 1] == 'C' && read[j] == 'T')) { counters[2]++; used[2] = true; t = 2; }
                else t = -1;
                if (lastT != t && lastT != -1) { contigs[lastT]++; if (contigLen > maxContigLen[lastT]) maxContigLen[lastT] = contigLen; contigLen = 0; }
            }
            if (t != -1) contigLen++;
        }
        if (t != -1) { contigs[t]++; if (contigLen > maxContigLen[t]) maxContigLen[t] = contigLen; }
        matcher.matchRead(entry, 0);
        std::vector<jlong> uniq = uc.getUniqueKmerCounts();
        for (int j = 0; j < 3; j++) {
            int vi = db.store->getIndexForValue(TAXIDS[(size_t)j]);
            CountsPerTaxid* stats = matcher.statsIndex[(size_t)vi].get();
            if (!used[j]) CHECK(stats == nullptr, "stats for unused taxon");
            else {
                CHECK(stats != nullptr, "missing stats");
                if (!stats) continue;
                CHECK(stats->kmers == counters[j], "kmers");
                CHECK(uniq[(size_t)vi] == 1, "unique == 1");
                CHECK(stats->contigs == contigs[j], "contigs");
                CHECK(stats->maxContigLen == maxContigLen[j], "maxContigLen");
            }
        }
    }
    return fails;
}

static int classifyOne(FastqKMerMatcher& m, ReadEntry& e, const char* cgatStr, std::string& cls) {
    // fillInRead / initEntry (T/match/FastqKMerMatcherTest.java:418-438)
    e.buffer.clear(); e.usedPaths = 0; e.classNode = nullptr;
    for (size_t i = 0; i < e.counts.size(); i++) { e.readTaxIdNode[i] = nullptr; e.counts[i] = 0; }
    size_t n = std::strlen(cgatStr);
    std::memcpy(e.read.data(), cgatStr, n);
    e.read[n] = 0;
    e.readNo++;
    e.readSize = (int)n;
    m.matchRead(e, 0);
    cls = e.classNode ? e.classNode->taxId : std::string("null");
    return 0;
}

// T/match/FastqKMerMatcherTest.java:315-412 (testReadClassification) with R/taxtree/{nodes,names}.dmp
int gso_kat_classification(int type) {
    int fails = 0; g_msg.clear();
    const std::string nodesDmp =
        "1\t|\t1\t|\tno rank\t|\t\t|\t8\t|\t0\t|\t1\t|\t0\t|\t0\t|\t0\t|\t0\t|\t0\t|\t\t|\n"
        "2\t|\t1\t|\tno rank\t|\t\t|\t8\t|\t0\t|\t1\t|\t0\t|\t0\t|\t0\t|\t0\t|\t0\t|\t\t|\n"
        "3\t|\t1\t|\tno rank\t|\t\t|\t8\t|\t0\t|\t1\t|\t0\t|\t0\t|\t0\t|\t0\t|\t0\t|\t\t|\n";
    const std::string namesDmp =
        "1\t|\t1\t|\t\t|\tscientific name\t|\n2\t|\t2\t|\t\t|\tscientific name\t|\n3\t|\t3\t|\t\t|\tscientific name\t|";
    const std::vector<std::string> TAXIDS = {"1", "2", "3"};
    TaxTree tree;
    tree.readNodes(nodesDmp); tree.readNames(namesDmp);
    for (auto& t : TAXIDS) tree.getNodeByTaxId(t)->markRequired();
    const uint8_t CC[2] = {'C', 'C'}, CT[2] = {'C', 'T'}, CG[2] = {'C', 'G'};
    jlong cc = cgat::kMerToLong(CC, 0, 2, nullptr), ct = cgat::kMerToLong(CT, 0, 2, nullptr), cg = cgat::kMerToLong(CG, 0, 2, nullptr);
    std::unique_ptr<KMerStoreBase> store = newStore(type, 2, {cc, ct, cg}, TAXIDS);
    store->putLong(cc, TAXIDS[0]);
    CHECK(store->putLong(ct, TAXIDS[1]), "CT put");
    store->putLong(cg, TAXIDS[2]);
    // (the reference test does not optimize; lookups fall back to linear scans)
    Database db;
    db.store = std::move(store);
    db.taxTree = tree.toSmallTaxTree();
    db.initStoreIndices();
    db.convert();
    struct Case { double err; const char* read; const char* cls; };
    const Case cases[] = {
        {0, "CCCC", "1"}, {0, "GAGAGA", "null"}, {0, "CCCG", "3"}, {0, "AGGGG", "2"}, {0, "CCCCCCT", "2"},
        {1, "CTCCT", "2"}, {1, "CTCTCCT", "null"}, {1, "TAGGGG", "2"}, {1, "TAGGGGT", "null"},
        {0.5, "CCA", "1"}, {0.5, "CCAA", "null"},
        {0.1, "CC", "1"}, {0.1, "CCA", "null"}, {0.1, "CCAA", "null"},
        {0.99, "TTTT", "null"}, {0.99, "CTTT", "2"}};
    ReadEntry entry; entry.init(10, true, 4);
    double lastErr = -123;
    std::unique_ptr<FastqKMerMatcher> matcher;
    for (const Case& c : cases) {
        if (c.err != lastErr) {
            MatchConfig cfg; cfg.k = 2; cfg.classify = true; cfg.maxPaths = 4; cfg.maxReadTaxErrorCount = c.err;
            cfg.maxReadClassErrorCount = -1; cfg.threshold = 1; cfg.maxKmerResCounts = 10; cfg.initialReadSize = 1024;
            matcher.reset(new FastqKMerMatcher(&db, cfg));
            matcher->initStats();
            lastErr = c.err;
        }
        std::string cls;
        classifyOne(*matcher, entry, c.read, cls);
        CHECK(cls == c.cls, std::string("read ") + c.read + " err " + std::to_string(c.err) + ": got " + cls + " want " + c.cls);
    }
    return fails;
}

// T/tax/SmallTaxTreeLCATest.java:56-125
int gso_kat_lca() {
    int fails = 0; g_msg.clear();
    int edges[7][2] = {{1, 1}, {2, 1}, {3, 2}, {4, 2}, {5, 3}, {6, 5}, {7, 1}};
    std::string nodes, names;
    for (auto& e : edges) {
        nodes += std::to_string(e[0]) + "\t|\t" + std::to_string(e[1]) + "\t|\tno rank\t|\t\t|\n";
        names += std::to_string(e[0]) + "\t|\t" + std::to_string(e[0]) + "\t|\t\t|\tscientific name\t|\n";
    }
    TaxTree full; full.readNodes(nodes); full.readNames(names);
    const char* IDS[7] = {"1", "2", "3", "4", "5", "6", "7"};
    for (auto id : IDS) full.getNodeByTaxId(id)->markRequired();
    auto t = full.toSmallTaxTree();
    auto node = [&](const char* id) { return t->getNodeByTaxId(id); };
    auto lca = [&](const char* a, const char* b) { return TaxTree::getLowestCommonAncestor(node(a), node(b)); };
    CHECK(lca("5", "6") == node("5"), "LCA(5,6)"); CHECK(lca("6", "5") == node("5"), "LCA(6,5)");
    CHECK(lca("1", "6") == node("1"), "LCA(1,6)"); CHECK(lca("6", "4") == node("2"), "LCA(6,4)");
    CHECK(lca("3", "4") == node("2"), "LCA(3,4)"); CHECK(lca("6", "7") == node("1"), "LCA(6,7)");
    CHECK(lca("6", "6") == node("6"), "LCA(6,6)");
    CHECK(TaxTree::getLowestCommonAncestor(nullptr, node("6")) == nullptr, "LCA(null,6)");
    CHECK(TaxTree::getLowestCommonAncestor(node("6"), nullptr) == nullptr, "LCA(6,null)");
    for (auto x : IDS) for (auto y : IDS) {
        TaxNode* bf = nullptr;
        for (TaxNode* a = node(x); a && !bf; a = a->parent) for (TaxNode* b = node(y); b; b = b->parent) if (a == b) { bf = a; break; }
        CHECK(lca(x, y) == bf, "brute force LCA");
    }
    int expectedLevel[7] = {0, 1, 2, 2, 3, 4, 1};
    for (int i = 0; i < 7; i++) { CHECK(node(IDS[i])->getLevel() == expectedLevel[i], "level"); CHECK(node(IDS[i])->depth == expectedLevel[i], "depth"); }
    return fails;
}

// T/fastq/FastqReaderTest.java:38-75 with fixture R/fastq/SimpleTest.fastq (initial buffer size 3 forces growth)
int gso_kat_fastq_reader(int withProbs) {
    int fails = 0; g_msg.clear();
    const std::string fq =
        "@S\nGATTTG\nGGGTTCAAAGCAGTATCGATCA\nA\nA\nTAGTAAATCCATTTGTTCAACTCACA\nGTT\nT\n+\n"
        "!''*((((**\n*+))%%%++)(%%%%).1**\n*-+*''))**55CCF>>>\n>>>C\nCCCCCC65\n"
        "@T\nC\nG\nA\nT\n+\n!\n*\n*\n>\n";
    int calls = 0;
    FastqReader fr; fr.k = 2;
    ReadEntry e; e.init(3, withProbs != 0, 1);
    fr.nextEntry = [&](ReadEntry& r) {
        calls++;
        std::string desc((const char*)r.readDescriptor.data(), (size_t)r.readDescriptorSize);
        std::string rd((const char*)r.read.data(), (size_t)r.readSize);
        if (calls == 1) {
            CHECK(desc == "@S", "descriptor 1: " + desc);
            CHECK(rd == "GATTTGGGGTTCAAAGCAGTATCGATCAAATAGTAAATCCATTTGTTCAACTCACAGTTT", "read 1: " + rd);
            if (withProbs) CHECK(std::string((const char*)r.readProbs.data(), (size_t)r.readProbsSize) ==
                                     "!''*((((***+))%%%++)(%%%%).1***-+*''))**55CCF>>>>>>CCCCCCC65", "probs 1");
            CHECK(r.readProbsSize == (int)rd.size(), "probs size 1");
        } else if (calls == 2) {
            CHECK(desc == "@T", "descriptor 2"); CHECK(rd == "CGAT", "read 2: " + rd);
            if (withProbs) CHECK(std::string((const char*)r.readProbs.data(), (size_t)r.readProbsSize) == "!**>", "probs 2");
            CHECK(r.readProbsSize == 4, "probs size 2");
        }
    };
    fr.readFastq((const uint8_t*)fq.data(), fq.size(), e);
    CHECK(calls == 2, "two reads");
    return fails;
}

}  // extern "C"
