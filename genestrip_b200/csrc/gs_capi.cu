// gs_capi.cu -- the C ABI of include/genestrip_b200.h: device-resident database, match / filter sessions with
// double-buffered batches on side streams, end-of-run merge.  No CPU fallback: every compute entry point needs a
// CUDA device and fails with GS_ERR_CUDA otherwise.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types only: the functions are resolved with dlopen (gs_nccl), the library does not link against NCCL

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "gs_kernels.cuh"
#include "gs_pack.hpp"

// ---------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int gs_fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return gs_fail(GS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define CUP(call)                                                                                             \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) {                                                                              \
            gs_fail(GS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return nullptr;                                                                                   \
        }                                                                                                     \
    } while (0)

template <typename T>
static cudaError_t dmalloc(T** p, size_t count) {
    *p = nullptr;
    return cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T));
}

// grow a device buffer (contents are not preserved); the caller guarantees no kernel still uses it
template <typename T>
static cudaError_t dgrow(T** p, size_t* cap, size_t need) {
    if (need <= *cap && *p) return cudaSuccess;
    if (*p) { cudaError_t e = cudaFree(*p); if (e != cudaSuccess) return e; *p = nullptr; }
    size_t ncap = need + need / 4 + 64;
    cudaError_t e = cudaMalloc((void**)p, ncap * sizeof(T));
    if (e == cudaSuccess) *cap = ncap;
    return e;
}
template <typename T>
static cudaError_t hgrow(T** p, size_t* cap, size_t need) {
    if (need <= *cap && *p) return cudaSuccess;
    if (*p) { cudaError_t e = cudaFreeHost(*p); if (e != cudaSuccess) return e; *p = nullptr; }
    size_t ncap = need + need / 4 + 64;
    cudaError_t e = cudaMallocHost((void**)p, ncap * sizeof(T));
    if (e == cudaSuccess) *cap = ncap;
    return e;
}

static u64 magic_for(u64 d) { return d ? (~0ULL) / d : 0; }

// ---------------------------------------------------------------------------------------------------------
// NCCL, resolved at run time: "libnccl.so.2" is the copy the host process already loaded (torch's, the JVM's) or the system's.
// Only the end-of-run merge across GPUs needs it; everything else works without.
// ---------------------------------------------------------------------------------------------------------
struct GsNccl {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};
static GsNccl* gs_nccl() {
    static GsNccl N;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("GS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            N.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (N.handle) break;
        }
        if (!N.handle) { N.error = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found") + "; set GS_NCCL_LIB"; return; }
        bool ok = true;
        auto sym = [&](const char* name) { void* p = dlsym(N.handle, name); if (!p) { ok = false; N.error = std::string("NCCL symbol missing: ") + name; } return p; };
        *(void**)&N.GetUniqueId = sym("ncclGetUniqueId"); *(void**)&N.CommInitRank = sym("ncclCommInitRank"); *(void**)&N.CommInitAll = sym("ncclCommInitAll");
        *(void**)&N.CommDestroy = sym("ncclCommDestroy"); *(void**)&N.AllReduce = sym("ncclAllReduce"); *(void**)&N.AllGather = sym("ncclAllGather");
        *(void**)&N.Send = sym("ncclSend"); *(void**)&N.Recv = sym("ncclRecv"); *(void**)&N.GroupStart = sym("ncclGroupStart"); *(void**)&N.GroupEnd = sym("ncclGroupEnd");
        *(void**)&N.GetErrorString = sym("ncclGetErrorString");
        if (!ok) { dlclose(N.handle); N.handle = nullptr; }
    });
    return N.handle ? &N : nullptr;
}
#define NC(call)                                                                                              \
    do {                                                                                                      \
        ncclResult_t r_ = (call);                                                                             \
        if (r_ != ncclSuccess)                                                                                \
            return gs_fail(GS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, gs_nccl()->GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------
struct gs_comm;
struct gs_ctx {
    std::vector<int> devs;
    std::vector<int> sms;
    bool peerAll = true;            // every device of the context can map every other one's memory
    gs_comm* allComm = nullptr;     // communicator over the context's own devices (gs_match_finish with n_devices > 1), made on first use
    // gs_inflate_blocks: scratch on device 0, one call at a time
    std::mutex infMutex;
    cudaStream_t infStream = nullptr;
    uint8_t* infComp = nullptr; size_t infCompCap = 0;
    uint8_t* infText = nullptr; size_t infTextCap = 0;
    gs_deflate_block* infBlocks = nullptr; size_t infBlocksCap = 0;
    ~gs_ctx() {
        if (infStream || infComp || infText || infBlocks) {
            cudaSetDevice(devs.empty() ? 0 : devs[0]);
            cudaFree(infComp); cudaFree(infText); cudaFree(infBlocks);
            if (infStream) cudaStreamDestroy(infStream);
        }
    }
};

extern "C" int gs_abi_version(void) { return GS_ABI_VERSION; }
extern "C" const char* gs_last_error(void) { return g_err.c_str(); }

extern "C" gs_ctx* gs_ctx_create(const int* device_ordinals, int n_devices) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        gs_fail(GS_ERR_CUDA, "no CUDA device (%s); this library has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return nullptr;
    }
    gs_ctx* c = new gs_ctx();
    if (!device_ordinals || n_devices <= 0) c->devs.push_back(0);
    else c->devs.assign(device_ordinals, device_ordinals + n_devices);
    for (int d : c->devs) {
        if (d < 0 || d >= count) { gs_fail(GS_ERR_ARG, "device ordinal %d out of range [0,%d)", d, count); delete c; return nullptr; }
        int sm = 0;
        cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, d);
        c->sms.push_back(sm > 0 ? sm : 148);
    }
    // peer access between the devices of the context (database replication, end-of-run merge)
    for (size_t i = 0; i < c->devs.size(); i++)
        for (size_t j = 0; j < c->devs.size(); j++) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, c->devs[i], c->devs[j]);
            if (can) { cudaSetDevice(c->devs[i]); cudaDeviceEnablePeerAccess(c->devs[j], 0); cudaGetLastError(); }
            else c->peerAll = false;
        }
    // The path is random 8-byte probes into multi-GB arrays: keep the L2 from promoting every 32-byte sector miss to a
    // 64/128-byte DRAM fetch (measured with ncu: 2.5 DRAM sectors per missed sector at the default granularity).
    size_t gran = 32;
    if (const char* e = getenv("GS_L2_FETCH_GRANULARITY")) gran = (size_t)atoi(e);
    for (int d : c->devs) {
        cudaSetDevice(d);
        if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        cudaGetLastError();
    }
    cudaSetDevice(c->devs[0]);
    return c;
}
extern "C" void gs_comm_destroy(gs_comm*);
extern "C" void gs_ctx_destroy(gs_ctx* c) {
    if (c && c->allComm) { gs_comm_destroy(c->allComm); c->allComm = nullptr; }
    delete c;
}
extern "C" int gs_ctx_n_devices(const gs_ctx* c) { return c ? (int)c->devs.size() : 0; }

static_assert(sizeof(gs_deflate_block) == 32, "gs_deflate_block is part of the ABI: 2 x u64 + 4 x u32");
// Block-gzip members -> text on the device (gs_inflate.cu), see genestrip_b200.h
extern "C" int gs_inflate_blocks(gs_ctx* c, const uint8_t* comp, uint64_t comp_bytes, gs_deflate_block* blocks, uint32_t n_blocks,
                                 uint8_t* out, uint64_t out_bytes) {
    if (!c) return gs_fail(GS_ERR_STATE, "null context");
    if (n_blocks == 0) return GS_OK;
    if (!comp || !blocks || (!out && out_bytes)) return gs_fail(GS_ERR_ARG, "null argument");
    for (uint32_t i = 0; i < n_blocks; i++) {
        const gs_deflate_block& b = blocks[i];
        if (b.in_off > comp_bytes || b.in_len > comp_bytes - b.in_off || b.out_off > out_bytes || b.out_len > out_bytes - b.out_off)
            return gs_fail(GS_ERR_ARG, "deflate block %u lies outside the buffers", i);
    }
    // the output ranges must not overlap (the blocks are inflated concurrently: an overlap would make text and CRC depend on
    // the order of the warps); they normally tile [0, out_bytes) -- if they leave gaps, the gaps come back as zero bytes, not
    // as whatever an earlier call left in the device buffer
    bool gaps = false;
    {
        uint64_t end = 0;
        for (uint32_t i = 0; i < n_blocks; i++) {
            if (blocks[i].out_off < end) return gs_fail(GS_ERR_ARG, "deflate block %u: output ranges must be ascending and must not overlap", i);
            gaps = gaps || blocks[i].out_off > end;
            end = blocks[i].out_off + blocks[i].out_len;
        }
        gaps = gaps || end < out_bytes;
    }
    std::lock_guard<std::mutex> lock(c->infMutex);
    CU(cudaSetDevice(c->devs[0]));
    if (!c->infStream) CU(cudaStreamCreateWithFlags(&c->infStream, cudaStreamNonBlocking));
    CU(dgrow(&c->infComp, &c->infCompCap, (size_t)comp_bytes + 16));
    CU(dgrow(&c->infText, &c->infTextCap, (size_t)out_bytes + 16));
    CU(dgrow(&c->infBlocks, &c->infBlocksCap, (size_t)n_blocks));
    if (gaps) CU(cudaMemsetAsync(c->infText, 0, (size_t)out_bytes, c->infStream));
    CU(cudaMemcpyAsync(c->infComp, comp, comp_bytes, cudaMemcpyHostToDevice, c->infStream));
    CU(cudaMemcpyAsync(c->infBlocks, blocks, (size_t)n_blocks * sizeof(gs_deflate_block), cudaMemcpyHostToDevice, c->infStream));
    gs_launch_inflate_blocks(c->infComp, c->infText, c->infBlocks, n_blocks, c->infStream);
    CU(cudaGetLastError());
    if (out_bytes) CU(cudaMemcpyAsync(out, c->infText, out_bytes, cudaMemcpyDeviceToHost, c->infStream));
    CU(cudaMemcpyAsync(blocks, c->infBlocks, (size_t)n_blocks * sizeof(gs_deflate_block), cudaMemcpyDeviceToHost, c->infStream));
    CU(cudaStreamSynchronize(c->infStream));
    for (uint32_t i = 0; i < n_blocks; i++)
        if (blocks[i].status) return gs_fail(GS_ERR_DATA, "block-gzip member %u is corrupt (%s)", i,
                                             blocks[i].status == 1 ? "malformed deflate stream" : blocks[i].status == 2 ? "size mismatch" : "CRC-32 mismatch");
    return GS_OK;
}

extern "C" void* gs_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, std::max<size_t>(bytes, 1));
    if (e != cudaSuccess) { gs_fail(GS_ERR_CUDA, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e)); return nullptr; }
    return p;
}
extern "C" void gs_free_pinned(void* p) { if (p) cudaFreeHost(p); }

// Frees the device temporaries of an entry point when its scope is left, early CU() returns included (the watched variables
// are read at that moment, so buffers grown or allocated later in the function are covered as well).
struct DevTemps {
    std::vector<void**> p;
    void watch(void** q) { p.push_back(q); }
    template <typename T> void watch(T** q) { p.push_back((void**)q); }
    void release() { for (void** q : p) if (*q) { cudaFree(*q); *q = nullptr; } }
    ~DevTemps() { release(); }
};

// ---------------------------------------------------------------------------------------------------------
// database
// ---------------------------------------------------------------------------------------------------------
struct DevDb {
    int dev = 0;
    u64* keys = nullptr;
    uint16_t* vals = nullptr;
    u32* bstart = nullptr;
    u64* bloom = nullptr;
    u64* tab = nullptr;
    u64* mzFilter = nullptr;
    int *parent = nullptr, *depth = nullptr, *pre = nullptr, *last = nullptr;
    GsDbView view;
};

struct gs_db {
    gs_ctx* ctx = nullptr;
    int k = 0;
    u64 n = 0;
    int V = 0;
    bool finalized = false;
    std::vector<DevDb> d;
    int16_t* rawVals = nullptr;  // device 0 staging of the Java shorts
    std::vector<int> hParent, hHasNode, hDepth, hPre, hLast;
    bool hasTree = false;
    // bloom
    bool hasBloom = false;
    long long bloomSeed = 0;
    u64 bloomBuckets = 0, bloomWords = 0;
    // bucket index
    int bbits = 0, bshift = 0;
    u64 nBuckets = 0;
    // probe table
    int tbits = 0, rbits = 0;
    u64 tabBuckets = 0;       // 2^tbits + GS_TAB_PAD_BUCKETS
    int mzBits = 0;  // log2 of the minimizer prefilter's size in bits; 0 = none
    bool mzWide = false;  // minimizers ordered by a 64-bit hash (stores whose filter would use most of the 32-bit hash space)
    bool seenLeased = false;  // the table's in-line seen bits belong to at most one unique-counting session at a time
    int openSessions = 0;     // match sessions that read this database (gs_db_update needs none)
    // radix source staging
    std::vector<std::pair<u64, int16_t>> radixItems;
    u64 bytes = 0;
};

extern "C" gs_db* gs_db_create(gs_ctx* ctx, int k, uint64_t n_kmers, int n_values) {
    if (!ctx) { gs_fail(GS_ERR_ARG, "null context"); return nullptr; }
    if (k < 1 || k > 31) { gs_fail(GS_ERR_ARG, "k=%d out of range [1,31]", k); return nullptr; }
    if (n_values < 0 || n_values > 65535) { gs_fail(GS_ERR_LIMIT, "n_values=%d exceeds 65535 (KMerSortedArray.MAX_VALUES)", n_values); return nullptr; }
    if (n_kmers >= 0xFFFFFFF0ULL) { gs_fail(GS_ERR_LIMIT, "n_kmers=%llu exceeds the 32-bit position limit of this build", (unsigned long long)n_kmers); return nullptr; }
    gs_db* db = new gs_db();
    db->ctx = ctx; db->k = k; db->n = n_kmers; db->V = n_values;
    db->d.resize(ctx->devs.size());
    for (size_t i = 0; i < ctx->devs.size(); i++) db->d[i].dev = ctx->devs[i];
    CUP(cudaSetDevice(ctx->devs[0]));
    CUP(dmalloc(&db->d[0].keys, n_kmers + 1));
    CUP(dmalloc(&db->rawVals, n_kmers));
    CUP(cudaMemset(db->d[0].keys, 0, (n_kmers + 1) * sizeof(u64)));
    CUP(cudaMemset(db->rawVals, 0x80, std::max<u64>(n_kmers, 1) * sizeof(int16_t)));  // 0x8080 -> index 128; overwritten by put_values
    return db;
}

extern "C" int gs_db_put_keys(gs_db* db, uint64_t offset, const int64_t* keys, uint64_t n) {
    if (!db || db->finalized) return gs_fail(GS_ERR_STATE, "database missing or already finalized");
    if (offset > db->n || n > db->n - offset) return gs_fail(GS_ERR_ARG, "key segment [%llu,%llu) exceeds n_kmers=%llu", (unsigned long long)offset, (unsigned long long)(offset + n), (unsigned long long)db->n);
    CU(cudaSetDevice(db->d[0].dev));
    CU(cudaMemcpy(db->d[0].keys + offset, keys, n * sizeof(u64), cudaMemcpyDefault));  // host or device source
    return GS_OK;
}

extern "C" int gs_db_put_values(gs_db* db, uint64_t offset, const int16_t* vidx_raw, uint64_t n) {
    if (!db || db->finalized) return gs_fail(GS_ERR_STATE, "database missing or already finalized");
    if (offset > db->n || n > db->n - offset) return gs_fail(GS_ERR_ARG, "value segment exceeds n_kmers");
    CU(cudaSetDevice(db->d[0].dev));
    CU(cudaMemcpy(db->rawVals + offset, vidx_raw, n * sizeof(int16_t), cudaMemcpyDefault));
    return GS_OK;
}

extern "C" int gs_db_put_radix_bucket(gs_db* db, int radix_bits, uint32_t radix, const int64_t* entries, uint32_t n) {
    if (!db || db->finalized) return gs_fail(GS_ERR_STATE, "database missing or already finalized");
    if (radix_bits < 16 || radix_bits > 30) return gs_fail(GS_ERR_ARG, "radix_bits=%d out of range [16,30]", radix_bits);
    if (radix >= (1u << radix_bits)) return gs_fail(GS_ERR_ARG, "radix %u out of range", radix);
    const int remainingBits = 62 - radix_bits;  // RadixKMerStore.java:165-168
    const u64 remainingMask = (1ULL << remainingBits) - 1;
    for (uint32_t i = 0; i < n; i++) {
        const u64 e = (u64)entries[i];
        const u64 kmer = ((e & remainingMask) << radix_bits) | (u64)radix;  // RadixKMerStore.visit :714-730
        const int vi = (int)(e >> remainingBits);
        if (vi >= db->V) return gs_fail(GS_ERR_ARG, "radix entry value index %d >= n_values", vi);
        db->radixItems.emplace_back(kmer, (int16_t)(vi - 32768));
    }
    return GS_OK;
}

extern "C" int gs_db_set_tree(gs_db* db, const int32_t* parent_by_vidx, const int32_t* has_node, int n_values) {
    if (!db || db->finalized) return gs_fail(GS_ERR_STATE, "database missing or already finalized");
    if (n_values != db->V) return gs_fail(GS_ERR_ARG, "n_values mismatch (%d vs %d)", n_values, db->V);
    const int V = db->V;
    db->hParent.assign(parent_by_vidx, parent_by_vidx + V);
    if (has_node) db->hHasNode.assign(has_node, has_node + V); else db->hHasNode.assign(V, 1);
    for (int v = 0; v < V; v++) {
        if (!db->hHasNode[v]) { db->hParent[v] = -1; continue; }
        int p = db->hParent[v];
        if (p < -1 || p >= V || p == v) return gs_fail(GS_ERR_ARG, "parent of value index %d is %d", v, p);
        if (p >= 0 && !db->hHasNode[p]) return gs_fail(GS_ERR_ARG, "parent %d of value index %d has no node", p, v);
    }
    // depth + DFS interval labels (a is ancestor-or-self of b <=> pre[a] <= pre[b] <= last[a])
    std::vector<std::vector<int>> kids(V);
    std::vector<int> roots;
    for (int v = 0; v < V; v++) {
        if (!db->hHasNode[v]) continue;
        if (db->hParent[v] < 0) roots.push_back(v); else kids[db->hParent[v]].push_back(v);
    }
    db->hDepth.assign(V, 0); db->hPre.assign(V, -1); db->hLast.assign(V, -1);
    int counter = 0, visited = 0, nodes = 0;
    for (int v = 0; v < V; v++) nodes += db->hHasNode[v] ? 1 : 0;
    std::vector<std::pair<int, size_t>> stack;
    for (int r : roots) {
        stack.emplace_back(r, 0);
        db->hPre[r] = counter++; db->hDepth[r] = 0; visited++;
        while (!stack.empty()) {
            auto& top = stack.back();
            if (top.second < kids[top.first].size()) {
                int c = kids[top.first][top.second++];
                db->hPre[c] = counter++; db->hDepth[c] = db->hDepth[stack.back().first] + 1; visited++;
                stack.emplace_back(c, 0);
            } else {
                db->hLast[top.first] = counter - 1;
                stack.pop_back();
            }
        }
    }
    if (visited != nodes) return gs_fail(GS_ERR_ARG, "tax tree has a cycle (%d of %d nodes reachable from roots)", visited, nodes);
    // nodes without a tree node never appear as labels; give them an empty interval
    for (int v = 0; v < V; v++) if (!db->hHasNode[v]) { db->hPre[v] = counter; db->hLast[v] = counter - 1; }
    db->hasTree = true;
    return GS_OK;
}

extern "C" int gs_db_set_bloom_blocked(gs_db* db, int64_t seed, uint64_t buckets, const int64_t* words, uint64_t n_words) {
    if (!db || db->finalized) return gs_fail(GS_ERR_STATE, "database missing or already finalized");
    if (buckets == 0 || n_words < buckets + 17) return gs_fail(GS_ERR_ARG, "blocked Bloom filter needs buckets+17 words (buckets=%llu, n_words=%llu)", (unsigned long long)buckets, (unsigned long long)n_words);
    CU(cudaSetDevice(db->d[0].dev));
    if (db->d[0].bloom) { CU(cudaFree(db->d[0].bloom)); db->d[0].bloom = nullptr; }
    CU(dmalloc(&db->d[0].bloom, buckets + 17));
    CU(cudaMemcpy(db->d[0].bloom, words, (buckets + 17) * sizeof(u64), cudaMemcpyHostToDevice));
    db->hasBloom = true; db->bloomSeed = seed; db->bloomBuckets = buckets; db->bloomWords = buckets + 17;
    return GS_OK;
}

static int upload_radix_items(gs_db* db) {
    if (db->radixItems.empty()) return GS_OK;
    if (db->radixItems.size() != db->n) return gs_fail(GS_ERR_ARG, "radix buckets hold %zu entries, n_kmers=%llu", db->radixItems.size(), (unsigned long long)db->n);
    std::sort(db->radixItems.begin(), db->radixItems.end());
    std::vector<u64> keys(db->n);
    std::vector<int16_t> vals(db->n);
    for (size_t i = 0; i < db->radixItems.size(); i++) { keys[i] = db->radixItems[i].first; vals[i] = db->radixItems[i].second; }
    db->radixItems.clear(); db->radixItems.shrink_to_fit();
    CU(cudaMemcpy(db->d[0].keys, keys.data(), db->n * sizeof(u64), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(db->rawVals, vals.data(), db->n * sizeof(int16_t), cudaMemcpyHostToDevice));
    return GS_OK;
}

extern "C" int gs_db_build_bloom_blocked(gs_db* db, int64_t* words_out, uint64_t n_words_out) {
    if (!db || db->finalized) return gs_fail(GS_ERR_STATE, "database missing or already finalized");
    CU(cudaSetDevice(db->d[0].dev));
    int rc = upload_radix_items(db);
    if (rc) return rc;
    // BlockedKMerBloomFilter.ensureExpectedSize (C/bloom/BlockedKMerBloomFilter.java:201-219) with bitsPerKey = 10
    const u64 entries = std::max<u64>(1, db->n);
    const u64 buckets = (entries * 10 + 63) / 64;
    const long long seed = -5025562857975149833LL;  // new Random(42).nextLong() (:91-93)
    if (db->d[0].bloom) { CU(cudaFree(db->d[0].bloom)); db->d[0].bloom = nullptr; }
    CU(dmalloc(&db->d[0].bloom, buckets + 17));
    CU(cudaMemset(db->d[0].bloom, 0, (buckets + 17) * sizeof(u64)));
    gs_launch_bloom_build(db->d[0].keys, db->n, db->d[0].bloom, buckets, magic_for(buckets), seed, 0);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    db->hasBloom = true; db->bloomSeed = seed; db->bloomBuckets = buckets; db->bloomWords = buckets + 17;
    if (words_out) {
        if (n_words_out < buckets + 17) return gs_fail(GS_ERR_ARG, "words_out too small: %llu < %llu", (unsigned long long)n_words_out, (unsigned long long)(buckets + 17));
        CU(cudaMemcpy(words_out, db->d[0].bloom, (buckets + 17) * sizeof(u64), cudaMemcpyDeviceToHost));
    }
    return GS_OK;
}

// Probe table of device 0 from its sorted keys / converted values: deterministic placement (gs_kernels.cu "probe table build").
static int build_table(gs_db* db) {
    DevDb& d0 = db->d[0];
    const u64 nB = db->tabBuckets;
    const u64 nBlocks = gs_table_scan_blocks(nB);
    u32 *counts = nullptr, *delta = nullptr, *blockIn = nullptr, *bad = nullptr;
    long long* agg = nullptr;
    struct Scratch { u32 **a, **b, **c, **d; long long** e; ~Scratch() { cudaFree(*a); cudaFree(*b); cudaFree(*c); cudaFree(*d); cudaFree(*e); } } guard{&counts, &delta, &blockIn, &bad, &agg};
    if (!d0.tab) CU(dmalloc(&d0.tab, nB * 4));
    CU(dmalloc(&counts, nB)); CU(dmalloc(&delta, nB + 1)); CU(dmalloc(&blockIn, nBlocks)); CU(dmalloc(&agg, nBlocks * 2)); CU(dmalloc(&bad, 1));
    CU(cudaMemset(d0.tab, 0, nB * 32));
    CU(cudaMemset(counts, 0, nB * sizeof(u32)));
    CU(cudaMemset(bad, 0, sizeof(u32)));
    gs_launch_table_build(d0.keys, d0.vals, db->n, d0.tab, counts, delta, agg, blockIn, bad, nB, db->rbits, 0);
    CU(cudaGetLastError());
    u32 tail = 0, hbad = 0;
    CU(cudaMemcpy(&tail, delta + nB, sizeof(u32), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&hbad, bad, sizeof(u32), cudaMemcpyDeviceToHost));
    if (tail != 0 || hbad != 0)
        return gs_fail(GS_ERR_LIMIT, "probe table: a collision chain does not fit (%s)", (hbad & 1u) ? "a bucket with 2^16 or more keys" : (hbad & 2u) ? "a key further from its home bucket than an entry can say" : "ran past the pad buckets");
    return GS_OK;
}

extern "C" int gs_db_finalize(gs_db* db) {
    if (!db || db->finalized) return gs_fail(GS_ERR_STATE, "database missing or already finalized");
    DevDb& d0 = db->d[0];
    CU(cudaSetDevice(d0.dev));
    int rc = upload_radix_items(db);
    if (rc) return rc;
    if (!db->hasTree) {  // no tree given: every value index is its own root (classification off / tests without tree)
        std::vector<int> parent(db->V, -1);
        rc = gs_db_set_tree(db, parent.data(), nullptr, db->V);
        if (rc) return rc;
    }
    const int V = db->V;
    u32* dBad = nullptr;
    CU(dmalloc(&dBad, 1));
    CU(cudaMemset(dBad, 0, sizeof(u32)));
    gs_launch_check_sorted(d0.keys, db->n, db->k, dBad, 0);
    CU(cudaGetLastError());
    u32 bad = 0;
    CU(cudaMemcpy(&bad, dBad, sizeof(u32), cudaMemcpyDeviceToHost));
    if (bad) { cudaFree(dBad); return gs_fail(GS_ERR_ARG, "keys are not strictly ascending 2k-bit values (%u violations)", bad); }
    // values: Java short -> value index, GS_VAL_NONODE for values without a tree node
    int* dHas = nullptr;
    CU(dmalloc(&dHas, (size_t)V));
    CU(cudaMemcpy(dHas, db->hHasNode.data(), (size_t)V * sizeof(int), cudaMemcpyHostToDevice));
    CU(dmalloc(&d0.vals, db->n));
    gs_launch_convert_values(db->rawVals, dHas, db->n, V, d0.vals, dBad, 0);
    CU(cudaGetLastError());
    CU(cudaMemcpy(&bad, dBad, sizeof(u32), cudaMemcpyDeviceToHost));
    CU(cudaFree(dHas));
    CU(cudaFree(dBad));
    if (bad) return gs_fail(GS_ERR_ARG, "%u stored values have an index >= n_values", bad);
    // Values whose tax id has no tree node were collapsed to GS_VAL_NONODE above; their original indices are only needed again
    // by gs_db_get_values / gs_db_save_file, so the raw shorts are kept exactly when such values exist.
    bool allNodes = true;
    for (int v = 0; v < V; v++) allNodes = allNodes && db->hHasNode[(size_t)v] != 0;
    if (allNodes) { CU(cudaFree(db->rawVals)); db->rawVals = nullptr; }
    // bucket index over the top key bits: ~4 keys per bucket on average
    int lg = 0;
    while ((1ULL << (lg + 1)) <= std::max<u64>(db->n, 1)) lg++;
    db->bbits = std::max(0, std::min(std::min(2 * db->k, 30), lg - 1));
    db->bshift = 2 * db->k - db->bbits;
    db->nBuckets = 1ULL << db->bbits;
    CU(dmalloc(&d0.bstart, db->nBuckets + 1));
    gs_launch_bucket_index(d0.keys, db->n, db->bshift, db->nBuckets, d0.bstart, 0);
    CU(cudaGetLastError());
    // minimizer prefilter: ~0.22 distinct minimizers per stored k-mer (window of 9), one bit each, fill <= ~12 %
    if (db->k >= GS_MZ_MIN_K && db->n > 0) {
        int fb = 16;
        while (fb < 32 && (double)(1ULL << fb) < 1.8 * (double)db->n) fb++;
        db->mzBits = fb;
        db->mzWide = fb > 30;
        if (const char* e = getenv("GS_DEBUG_MZ_WIDE")) db->mzWide = atoi(e) != 0;   // tests: force the 64-bit minimizer order on small stores
        CU(dmalloc(&d0.mzFilter, (size_t)(1ULL << (fb - 6))));
        CU(cudaMemset(d0.mzFilter, 0, (size_t)(1ULL << (fb - 3))));
        gs_launch_mz_build(d0.keys, db->n, db->k, d0.mzFilter, (u32)((1ULL << fb) - 1), db->mzWide ? 1 : 0, 0);
        CU(cudaGetLastError());
    }
    // probe table: 1..2 keys per 4-slot bucket
    {
        int tb = GS_TAB_MIN_BITS;
        while ((2ULL << tb) < db->n) tb++;
        if (const char* e = getenv("GS_DEBUG_TAB_BITS")) { const int v = atoi(e); if (v >= GS_TAB_MIN_BITS && v <= 40) tb = v; }   // tests: an overfull table must be refused / doubled
        for (int attempt = 0;; attempt++) {
            db->tbits = tb; db->rbits = 62 - tb;
            db->tabBuckets = (1ULL << tb) + GS_TAB_PAD_BUCKETS;
            rc = build_table(db);
            if (rc == GS_OK) break;
            // a chain that does not fit (possible, if astronomically unlikely, at the default load): same keys, twice the buckets
            if (rc != GS_ERR_LIMIT || attempt == 2) return rc;
            CU(cudaFree(d0.tab)); d0.tab = nullptr;
            tb++;
        }
    }
    // tree
    CU(dmalloc(&d0.parent, (size_t)V)); CU(dmalloc(&d0.depth, (size_t)V)); CU(dmalloc(&d0.pre, (size_t)V)); CU(dmalloc(&d0.last, (size_t)V));
    CU(cudaMemcpy(d0.parent, db->hParent.data(), (size_t)V * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d0.depth, db->hDepth.data(), (size_t)V * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d0.pre, db->hPre.data(), (size_t)V * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d0.last, db->hLast.data(), (size_t)V * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaDeviceSynchronize());
    // replicate to the other devices of the context
    for (size_t i = 1; i < db->d.size(); i++) {
        DevDb& di = db->d[i];
        CU(cudaSetDevice(di.dev));
        CU(dmalloc(&di.keys, db->n + 1)); CU(dmalloc(&di.vals, db->n)); CU(dmalloc(&di.bstart, db->nBuckets + 1));
        CU(dmalloc(&di.parent, (size_t)V)); CU(dmalloc(&di.depth, (size_t)V)); CU(dmalloc(&di.pre, (size_t)V)); CU(dmalloc(&di.last, (size_t)V));
        CU(cudaMemcpyPeer(di.keys, di.dev, d0.keys, d0.dev, (db->n + 1) * sizeof(u64)));
        CU(cudaMemcpyPeer(di.vals, di.dev, d0.vals, d0.dev, db->n * sizeof(uint16_t)));
        CU(cudaMemcpyPeer(di.bstart, di.dev, d0.bstart, d0.dev, (db->nBuckets + 1) * sizeof(u32)));
        CU(dmalloc(&di.tab, db->tabBuckets * 4));
        CU(cudaMemcpyPeer(di.tab, di.dev, d0.tab, d0.dev, db->tabBuckets * 32));
        CU(cudaMemcpyPeer(di.parent, di.dev, d0.parent, d0.dev, (size_t)V * sizeof(int)));
        CU(cudaMemcpyPeer(di.depth, di.dev, d0.depth, d0.dev, (size_t)V * sizeof(int)));
        CU(cudaMemcpyPeer(di.pre, di.dev, d0.pre, d0.dev, (size_t)V * sizeof(int)));
        CU(cudaMemcpyPeer(di.last, di.dev, d0.last, d0.dev, (size_t)V * sizeof(int)));
        if (db->mzBits) {
            CU(dmalloc(&di.mzFilter, (size_t)(1ULL << (db->mzBits - 6))));
            CU(cudaMemcpyPeer(di.mzFilter, di.dev, d0.mzFilter, d0.dev, (size_t)(1ULL << (db->mzBits - 3))));
        }
        if (db->hasBloom) {
            CU(dmalloc(&di.bloom, db->bloomWords));
            CU(cudaMemcpyPeer(di.bloom, di.dev, d0.bloom, d0.dev, db->bloomWords * sizeof(u64)));
        }
        CU(cudaDeviceSynchronize());
    }
    for (DevDb& d : db->d) {
        GsDbView& v = d.view;
        v.keys = d.keys; v.vals = d.vals; v.n = db->n; v.bstart = d.bstart; v.bshift = db->bshift; v.nBuckets = db->nBuckets;
        v.k = db->k; v.bloom = d.bloom; v.bloomBuckets = db->bloomBuckets; v.bloomMagic = magic_for(db->bloomBuckets);
        v.bloomSeed = db->bloomSeed; v.hasBloom = db->hasBloom ? 1 : 0;
        v.tab = d.tab; v.tbits = db->tbits; v.rbits = db->rbits; v.tabSlots = db->tabBuckets * GS_TAB_SLOT_STRIDE;
        v.mzFilter = d.mzFilter; v.mzMask = db->mzBits ? (u32)((1ULL << db->mzBits) - 1) : 0u; v.mzWide = db->mzWide ? 1 : 0;
        v.parent = d.parent; v.depth = d.depth; v.pre = d.pre; v.last = d.last; v.nValues = V;
    }
    db->bytes = db->tabBuckets * 32 + (db->n + 1) * 8 + db->n * 2 + (db->nBuckets + 1) * 4 + (db->hasBloom ? db->bloomWords * 8 : 0) + (u64)V * 16 + (db->mzBits ? (1ULL << (db->mzBits - 3)) : 0);
    CU(cudaSetDevice(d0.dev));
    db->finalized = true;
    return GS_OK;
}

// ---- flat database file ("GSB1") -----------------------------------------------------------------------------
// What Database.save puts into a zip of Java-serialized objects (C/store/Database.java:201-260), as flat little-endian
// arrays a C program (and a 40-line Java exporter, integration/java/.../GsbExporter.java) can stream:
//   header: char magic[8] = "GSB1", u32 version, u32 k, u64 n_kmers, u32 n_values, u32 flags (bit 0: blocked Bloom filter),
//           i64 bloom_seed, u64 bloom_buckets, u64 bloom_words                       (56 bytes)
//   i64 keys[n_kmers]; i16 values[n_kmers] (Java shorts) + pad to 8; i32 parent[n_values]; i32 has_node[n_values] + pad to 8;
//   i64 bloom[bloom_words]
struct GsbHeader {
    char magic[8];
    uint32_t version, k;
    uint64_t n_kmers;
    uint32_t n_values, flags;
    int64_t bloom_seed;
    uint64_t bloom_buckets, bloom_words;
};
static_assert(sizeof(GsbHeader) == 56, "GSB1 header layout");

extern "C" int gs_db_info(const gs_db* db, int* k, uint64_t* n_kmers, int* n_values) {
    if (!db) return gs_fail(GS_ERR_ARG, "null database");
    if (k) *k = db->k;
    if (n_kmers) *n_kmers = db->n;
    if (n_values) *n_values = db->V;
    return GS_OK;
}

extern "C" int gs_db_save_file(gs_db* db, const char* path) {
    if (!db || !db->finalized) return gs_fail(GS_ERR_STATE, "database not finalized");
    if (!path) return gs_fail(GS_ERR_ARG, "null path");
    FILE* f = fopen(path, "wb");
    if (!f) return gs_fail(GS_ERR_ARG, "cannot create %s", path);
    struct Closer { FILE* f; ~Closer() { if (f) fclose(f); } } closer{f};
    GsbHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, "GSB1", 4);
    h.version = 1; h.k = (uint32_t)db->k; h.n_kmers = db->n; h.n_values = (uint32_t)db->V; h.flags = db->hasBloom ? 1u : 0u;
    h.bloom_seed = db->bloomSeed; h.bloom_buckets = db->bloomBuckets; h.bloom_words = db->hasBloom ? db->bloomWords : 0;
    if (fwrite(&h, sizeof(h), 1, f) != 1) return gs_fail(GS_ERR_ARG, "write error on %s", path);
    DevDb& d0 = db->d[0];
    CU(cudaSetDevice(d0.dev));
    const u64 seg = 1ULL << 24;
    std::vector<u64> buf((size_t)std::min<u64>(seg, std::max<u64>(db->n, 1)));
    for (u64 off = 0; off < db->n; off += seg) {
        const u64 n = std::min(seg, db->n - off);
        CU(cudaMemcpy(buf.data(), d0.keys + off, n * sizeof(u64), cudaMemcpyDeviceToHost));
        if (fwrite(buf.data(), sizeof(u64), n, f) != n) return gs_fail(GS_ERR_ARG, "write error on %s", path);
    }
    std::vector<int16_t> vb((size_t)std::min<u64>(seg, std::max<u64>(db->n, 1)));
    for (u64 off = 0; off < db->n; off += seg) {
        const u64 n = std::min(seg, db->n - off);
        int rc = gs_db_get_values(db, off, vb.data(), n);
        if (rc) return rc;
        if (fwrite(vb.data(), sizeof(int16_t), n, f) != n) return gs_fail(GS_ERR_ARG, "write error on %s", path);
    }
    const char zeros[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const size_t padV = (size_t)((8 - (db->n * 2) % 8) % 8);
    if (padV && fwrite(zeros, 1, padV, f) != padV) return gs_fail(GS_ERR_ARG, "write error on %s", path);
    const size_t V = (size_t)db->V;
    if (V && (fwrite(db->hParent.data(), sizeof(int), V, f) != V || fwrite(db->hHasNode.data(), sizeof(int), V, f) != V))
        return gs_fail(GS_ERR_ARG, "write error on %s", path);
    const size_t padT = (size_t)((8 - (V * 8) % 8) % 8);
    if (padT && fwrite(zeros, 1, padT, f) != padT) return gs_fail(GS_ERR_ARG, "write error on %s", path);
    if (db->hasBloom) {
        std::vector<u64> bw((size_t)db->bloomWords);
        CU(cudaMemcpy(bw.data(), d0.bloom, db->bloomWords * sizeof(u64), cudaMemcpyDeviceToHost));
        if (fwrite(bw.data(), sizeof(u64), bw.size(), f) != bw.size()) return gs_fail(GS_ERR_ARG, "write error on %s", path);
    }
    if (fflush(f) != 0) return gs_fail(GS_ERR_ARG, "write error on %s", path);
    return GS_OK;
}

static gs_db* db_load_file(gs_ctx* ctx, const char* path);
extern "C" gs_db* gs_db_load_file(gs_ctx* ctx, const char* path) {
    // nothing may unwind through the C boundary (a JVM sits on the other side): a file whose header asks for absurd sizes
    // ends as an error code, not as std::bad_alloc
    try {
        return db_load_file(ctx, path);
    } catch (const std::exception& e) {
        gs_fail(GS_ERR_ARG, "%s: %s", path ? path : "(null)", e.what());
        return nullptr;
    }
}
static gs_db* db_load_file(gs_ctx* ctx, const char* path) {
    if (!ctx || !path) { gs_fail(GS_ERR_ARG, "null argument"); return nullptr; }
    FILE* f = fopen(path, "rb");
    if (!f) { gs_fail(GS_ERR_ARG, "cannot open %s", path); return nullptr; }
    struct Closer { FILE* f; ~Closer() { if (f) fclose(f); } } closer{f};
    GsbHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "GSB1\0\0\0\0", 8) != 0 || h.version != 1) { gs_fail(GS_ERR_ARG, "%s is not a GSB1 database file", path); return nullptr; }
    {   // the header's sizes must add up to the file's size before anything is allocated from them
        const long at = ftell(f);
        if (fseek(f, 0, SEEK_END) != 0) { gs_fail(GS_ERR_ARG, "%s: cannot seek", path); return nullptr; }
        const u64 fileBytes = (u64)ftell(f);
        fseek(f, at, SEEK_SET);
        const bool bloom = (h.flags & 1u) != 0;
        const bool sane = h.k >= 1 && h.k <= 31 && h.n_values <= 65536 && h.n_kmers <= fileBytes / 10 &&
                          (!bloom || (h.bloom_words == h.bloom_buckets + 17 && h.bloom_words <= fileBytes / 8));
        const u64 want = sizeof(GsbHeader) + h.n_kmers * 8 + ((h.n_kmers * 2 + 7) & ~7ULL) + (((u64)h.n_values * 8 + 7) & ~7ULL) + (bloom ? h.bloom_words * 8 : 0);
        if (!sane || want != fileBytes) { gs_fail(GS_ERR_ARG, "%s: header does not match the file (%llu bytes, header implies %llu)", path, (unsigned long long)fileBytes, (unsigned long long)want); return nullptr; }
    }
    gs_db* db = gs_db_create(ctx, (int)h.k, h.n_kmers, (int)h.n_values);
    if (!db) return nullptr;
    auto bail = [&](const char* what) -> gs_db* { std::string msg = gs_last_error(); gs_db_destroy(db); gs_fail(GS_ERR_ARG, "%s: %s %s", path, what, msg.c_str()); return nullptr; };
    const u64 seg = 1ULL << 24;
    {
        std::vector<int64_t> buf((size_t)std::min<u64>(seg, std::max<u64>(h.n_kmers, 1)));
        for (u64 off = 0; off < h.n_kmers; off += seg) {
            const u64 n = std::min(seg, h.n_kmers - off);
            if (fread(buf.data(), sizeof(int64_t), n, f) != n) return bail("truncated key section");
            if (gs_db_put_keys(db, off, buf.data(), n) != GS_OK) return bail("keys:");
        }
        std::vector<int16_t> vb((size_t)std::min<u64>(seg, std::max<u64>(h.n_kmers, 1)));
        for (u64 off = 0; off < h.n_kmers; off += seg) {
            const u64 n = std::min(seg, h.n_kmers - off);
            if (fread(vb.data(), sizeof(int16_t), n, f) != n) return bail("truncated value section");
            if (gs_db_put_values(db, off, vb.data(), n) != GS_OK) return bail("values:");
        }
        char pad[8];
        const size_t padV = (size_t)((8 - (h.n_kmers * 2) % 8) % 8);
        if (padV && fread(pad, 1, padV, f) != padV) return bail("truncated");
    }
    {
        const size_t V = h.n_values;
        std::vector<int> parent(std::max<size_t>(V, 1)), has(std::max<size_t>(V, 1));
        if (V && (fread(parent.data(), sizeof(int), V, f) != V || fread(has.data(), sizeof(int), V, f) != V)) return bail("truncated tree section");
        char pad[8];
        const size_t padT = (size_t)((8 - (V * 8) % 8) % 8);
        if (padT && fread(pad, 1, padT, f) != padT) return bail("truncated");
        if (gs_db_set_tree(db, parent.data(), has.data(), (int)V) != GS_OK) return bail("tree:");
    }
    if (h.flags & 1u) {
        std::vector<int64_t> bw((size_t)h.bloom_words);
        if (fread(bw.data(), sizeof(int64_t), bw.size(), f) != bw.size()) return bail("truncated Bloom filter section");
        if (gs_db_set_bloom_blocked(db, h.bloom_seed, h.bloom_buckets, bw.data(), h.bloom_words) != GS_OK) return bail("bloom:");
    }
    if (gs_db_finalize(db) != GS_OK) return bail("finalize:");
    return db;
}

// ---- database update phase --------------------------------------------------------------------------------
extern "C" int gs_db_update(gs_db* db, const uint8_t* seq, uint64_t n_bytes, const uint64_t* region_offsets, const int32_t* region_vidx,
                            uint32_t n_regions, int upper_case, uint64_t* n_changed) {
    if (!db || !db->finalized) return gs_fail(GS_ERR_STATE, "database not finalized");
    if (db->seenLeased || db->openSessions > 0) return gs_fail(GS_ERR_STATE, "%d match session(s) are open on this database: close them before updating", db->openSessions);
    if (n_changed) *n_changed = 0;
    if (n_regions == 0 || n_bytes == 0) return GS_OK;
    if (!seq || !region_offsets || !region_vidx) return gs_fail(GS_ERR_ARG, "null argument");
    if (region_offsets[0] != 0 || region_offsets[n_regions] != n_bytes) return gs_fail(GS_ERR_ARG, "region offsets must span [0, n_bytes]");
    for (u32 i = 0; i < n_regions; i++) {
        if (region_offsets[i + 1] < region_offsets[i]) return gs_fail(GS_ERR_ARG, "region offsets not ascending at region %u", i);
        if (region_offsets[i + 1] - region_offsets[i] > 0x7FFFFFF0ULL) return gs_fail(GS_ERR_LIMIT, "region %u longer than 2^31 bases", i);
    }
    DevDb& d0 = db->d[0];
    CU(cudaSetDevice(d0.dev));
    const u64 kGroup = 64ULL << 20;  // bases per pass (labels + positions take 12 bytes per base)
    unsigned long long* dChanged = nullptr;
    u32* dCtr = nullptr;
    uint8_t* dSeq = nullptr; u64* dOff = nullptr; int* dNode = nullptr; u32* dLab = nullptr; long long* dPos = nullptr; u32* dValid = nullptr; u32* dStart = nullptr;
    DevTemps temps;
    temps.watch(&dChanged); temps.watch(&dCtr); temps.watch(&dSeq); temps.watch(&dOff); temps.watch(&dNode); temps.watch(&dLab); temps.watch(&dPos);
    temps.watch(&dValid); temps.watch(&dStart);
    CU(dmalloc(&dChanged, 1)); CU(cudaMemset(dChanged, 0, sizeof(unsigned long long)));
    CU(dmalloc(&dCtr, 8));
    size_t seqCap = 0, offCap = 0, nodeCap = 0, labCap = 0, posCap = 0, validCap = 0, startCap = 0;
    int rc = GS_OK;
    std::vector<u64> rel;
    for (u32 r0 = 0; r0 < n_regions && rc == GS_OK;) {
        u32 r1 = r0 + 1;
        while (r1 < n_regions && region_offsets[r1 + 1] - region_offsets[r0] <= kGroup) r1++;
        const u64 b0 = region_offsets[r0], nb = region_offsets[r1] - b0;
        const u32 nr = r1 - r0;
        rel.resize((size_t)nr + 1);
        for (u32 i = 0; i <= nr; i++) rel[i] = region_offsets[r0 + i] - b0;
        const u64 nSeg = (nb + GS_SEG_POS - 1) / GS_SEG_POS;
        const size_t words = (size_t)nSeg * GS_SEG_CHUNKS + 64;
        CU(dgrow(&dSeq, &seqCap, (size_t)nb + 64)); CU(dgrow(&dOff, &offCap, (size_t)nr + 1)); CU(dgrow(&dNode, &nodeCap, (size_t)nr));
        CU(dgrow(&dLab, &labCap, (size_t)nb + 32)); CU(dgrow(&dPos, &posCap, (size_t)nb + 32));
        CU(dgrow(&dValid, &validCap, words)); CU(dgrow(&dStart, &startCap, words));
        if (nb) CU(cudaMemcpy(dSeq, seq + b0, nb, cudaMemcpyHostToDevice));
        CU(cudaMemset(dSeq + nb, 0, 64));
        CU(cudaMemcpy(dOff, rel.data(), ((size_t)nr + 1) * sizeof(u64), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(dNode, region_vidx + r0, (size_t)nr * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemset(dCtr, 0, 8 * sizeof(u32)));
        CU(cudaMemset(dStart, 0, words * sizeof(u32)));
        if (upper_case) gs_launch_cgat_upper(dSeq, nb, 0);
        GsMatchParams P;
        memset(&P, 0, sizeof(P));
        P.db = d0.view;
        P.bases = dSeq; P.offsets = dOff; P.nReads = nr; P.layout = GS_LAYOUT_CLASSIC; P.useBloom = db->hasBloom ? 1 : 0;
        P.off0 = 0; P.lead = 0; P.flatLen = nb; P.labels = dLab; P.validBits = dValid; P.startBits = dStart; P.segCounter = dCtr + 2;
        P.flatPos = dPos;
        gs_launch_mark_starts(P, 0);
        gs_launch_label(P, true, (int)std::max<u64>(1, std::min<u64>((u64)db->ctx->sms[0] * 4, (nSeg + GS_WARPS_PER_BLOCK - 1) / GS_WARPS_PER_BLOCK)), 0);
        gs_launch_db_update(d0.view, dLab, dPos, nb, dOff, dNode, nr, d0.vals, dChanged, 0);
        CU(cudaGetLastError());
        CU(cudaDeviceSynchronize());
        r0 = r1;
    }
    unsigned long long changed = 0;
    CU(cudaMemcpy(&changed, dChanged, sizeof(changed), cudaMemcpyDeviceToHost));
    if (n_changed) *n_changed = changed;
    temps.release();   // before the table is rebuilt (early returns above are covered by the guard's destructor)
    if (changed) {  // the probe table carries the values in its entries: rebuild it, then refresh the replicas
        int rcb = build_table(db);
        if (rcb) return rcb;
        CU(cudaDeviceSynchronize());
        for (size_t i = 1; i < db->d.size(); i++) {
            DevDb& di = db->d[i];
            CU(cudaMemcpyPeer(di.vals, di.dev, d0.vals, d0.dev, db->n * sizeof(uint16_t)));
            CU(cudaMemcpyPeer(di.tab, di.dev, d0.tab, d0.dev, db->tabBuckets * 32));
        }
        CU(cudaDeviceSynchronize());
    }
    return rc;
}

extern "C" int gs_db_get_values(gs_db* db, uint64_t offset, int16_t* vidx_raw, uint64_t n) {
    if (!db || !db->finalized) return gs_fail(GS_ERR_STATE, "database not finalized");
    if (offset > db->n || n > db->n - offset) return gs_fail(GS_ERR_ARG, "range beyond the %llu stored k-mers", (unsigned long long)db->n);
    if (n == 0) return GS_OK;
    if (!vidx_raw) return gs_fail(GS_ERR_ARG, "null argument");
    CU(cudaSetDevice(db->d[0].dev));
    int16_t* dRaw = nullptr;
    CU(dmalloc(&dRaw, (size_t)n));
    struct Free { void* p; ~Free() { cudaFree(p); } } fr{dRaw};
    gs_launch_values_to_raw(db->d[0].vals + offset, db->rawVals ? db->rawVals + offset : nullptr, n, dRaw, 0);
    CU(cudaGetLastError());
    CU(cudaMemcpy(vidx_raw, dRaw, n * sizeof(int16_t), cudaMemcpyDeviceToHost));
    return GS_OK;
}

extern "C" void gs_db_destroy(gs_db* db) {
    if (!db) return;
    for (DevDb& d : db->d) {
        cudaSetDevice(d.dev);
        cudaFree(d.keys); cudaFree(d.vals); cudaFree(d.bstart); cudaFree(d.bloom); cudaFree(d.tab); cudaFree(d.mzFilter);
        cudaFree(d.parent); cudaFree(d.depth); cudaFree(d.pre); cudaFree(d.last);
    }
    if (db->rawVals) { cudaSetDevice(db->d[0].dev); cudaFree(db->rawVals); }
    delete db;
}

extern "C" gs_ctx* gs_db_context(gs_db* db) { return db ? db->ctx : nullptr; }
extern "C" uint64_t gs_db_device_bytes(const gs_db* db) { return db ? db->bytes : 0; }
extern "C" int gs_db_n_devices(const gs_db* db) { return db ? (int)db->d.size() : 0; }

extern "C" int gs_db_lookup(gs_db* db, const int64_t* kmers, uint64_t n, int use_bloom, int32_t* vidx_out, int64_t* pos_out) {
    if (!db || !db->finalized) return gs_fail(GS_ERR_STATE, "database not finalized");
    CU(cudaSetDevice(db->d[0].dev));
    u64* dK = nullptr; int* dV = nullptr; long long* dP = nullptr;
    DevTemps temps; temps.watch(&dK); temps.watch(&dV); temps.watch(&dP);
    CU(dmalloc(&dK, n)); CU(dmalloc(&dV, n)); CU(dmalloc(&dP, n));
    CU(cudaMemcpy(dK, kmers, n * sizeof(u64), cudaMemcpyHostToDevice));
    if (n) gs_launch_lookup(db->d[0].view, dK, n, use_bloom, dV, dP, 0);
    CU(cudaGetLastError());
    CU(cudaMemcpy(vidx_out, dV, n * sizeof(int), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(pos_out, dP, n * sizeof(long long), cudaMemcpyDeviceToHost));
    return GS_OK;
}

// ---------------------------------------------------------------------------------------------------------
// match sessions
// ---------------------------------------------------------------------------------------------------------
template <typename Slot> static void text_stage_free(Slot& sl);
struct MatchSlot {
    bool pending = false;
    gs_ticket ticket = 0;
    u32 nReads = 0;
    u64 firstReadNo = 0;
    u64 totalKmers = 0;
    uint8_t* dBases = nullptr; size_t basesCap = 0;
    u64* dOffsets = nullptr; size_t offCap = 0;
    gs_read_result* dOut = nullptr; size_t outCap = 0;
    gs_read_result* hOut = nullptr; size_t hOutCap = 0;
    gs_maxcontig_event* dEv = nullptr; gs_maxcontig_event* hEv = nullptr;
    u32* dNEv = nullptr; u32* hNEv = nullptr;  // [0] = number of events, [1] = malformed-offsets flag
    // want_runs
    u64* dKmerOff = nullptr; size_t kmerOffCap = 0;
    u64* hKmerOff = nullptr; size_t hKmerOffCap = 0;
    gs_run* dRuns = nullptr; size_t runsCap = 0;
    u32* dRunCounts = nullptr; size_t runCountsCap = 0;
    u32* hRunCounts = nullptr; size_t hRunCountsCap = 0;
    // raw FASTQ text batches (gs_match_submit_fastq)
    uint8_t* dText = nullptr; size_t textCap = 0;
    u32* dLineEnd = nullptr; size_t lineCap = 0;
    u32* dBlockCounts = nullptr; size_t blockCountsCap = 0;
    u32* dTextMeta = nullptr;              // [0] lines, [1] records, [2] status bits, [3] pad, then u64 totals[2] (k-mers, bases)
    u32* hTextMeta = nullptr;              // pinned copy
    gs_fastq_rec* dRecs = nullptr; size_t recsCap = 0;
    gs_fastq_rec* hRecs = nullptr; size_t hRecsCap = 0;
    u32* dLens = nullptr; size_t lensCap = 0;
    u64* dTileSums = nullptr; size_t tileSumsCap = 0;
    u32* dEvHdr = nullptr; u32* hEvHdr = nullptr;
    bool isText = false;
    // bases packed on the host (gs_match_cfg.host_pack_threads != 0): pinned staging + device copies of codes / validity words
    u64* hCodes = nullptr; size_t hCodesCap = 0;
    u32* hValid = nullptr; size_t hValidCap = 0;
    u64* dCodes = nullptr; size_t dCodesCap = 0;
    u32* dValid = nullptr; size_t dValidCap = 0;
    // split of the last submit (adaptive link split, see gs_match_submit): ASCII bytes copied and the events around that copy
    u64 asciiBytes = 0, packedBases = 0;
    double packSec = 0;
    cudaEvent_t evAscii0 = nullptr, evAscii1 = nullptr;
    cudaEvent_t evH2D = nullptr, evCompute = nullptr, evDone = nullptr;
};

struct DevSess {
    int dev = 0, devIndex = 0, sms = 148;
    cudaStream_t sCopyIn = nullptr, sCompute = nullptr, sCopyOut = nullptr;
    size_t l2WindowBytes = 0, l2CarveBytes = 0;   // access policy window on sCompute (minimizer prefilter), 0 = none
    float l2HitRatio = 0.f;
    long long* counters = nullptr;  // [7][V]
    u64* maxcontig = nullptr;       // [V]
    u64* bitset = nullptr; u64 bitsetWords = 0;
    uint16_t* hitCounts = nullptr;
    long long* unique = nullptr;    // [V]
    u32* popPartial = nullptr; size_t popPartialCap = 0;   // per-CTA count tables of the popcount kernel
    u32* slowTable = nullptr;
    int fastBlocks = 0, slowBlocks = 0, labelBlocks = 0, labelBlocksOverlap = 0;
    // label kernel -> reduce kernels hand-over.  Two sets per device, used by alternate batches: the reduce kernels of batch i
    // run on their own stream (sReduce) while the label kernel of batch i + 1 already fills the other set -- the thread-per-read
    // reduce is a latency-bound kernel of 0.3-0.4 ms that hides behind the label kernel when that leaves it one CTA slot per SM.
    // GS_OVERLAP_REDUCE=0: sReduce is the compute stream itself and everything runs in stream order as before (A/B).
    struct HandOver {
        u32* labels = nullptr; size_t labelsCap = 0;
        u32* validBits = nullptr; size_t validCap = 0;
        u32* startBits = nullptr; size_t startCap = 0;
        u32* bmask = nullptr; size_t bmaskCap = 0;
        u32* redoList = nullptr; size_t redoCap = 0;
        u32* overflowList = nullptr; size_t ovCap = 0;
        u32* overflowCount = nullptr;   // [0] overflow list length, [1] read claim counter, [2] segment claim counter, [3] redo list length, [4] read-group claim counter
        cudaEvent_t evLabel = nullptr, evReduce = nullptr;
        bool reduceRecorded = false;
    } ho[2];
    cudaStream_t sReduce = nullptr;     // == sCompute when the overlap is off
    bool overlap = false;
    u64 batchNo = 0;
    std::vector<cudaEvent_t> timingEv;  // gs_match_set_timing: 4 events per batch (before / after the label kernel, before / after the reduce kernels)
    MatchSlot slots[GS_MAX_INFLIGHT];
};

// The compute stream waits for the reduce kernels still in flight on sReduce (no host wait): whatever is launched on sCompute
// next -- merge, popcount, a dump, the caller's own event -- is ordered behind all the work of the batches submitted so far.
static int join_reduce(DevSess& D) {
    if (!D.overlap) return GS_OK;
    for (DevSess::HandOver& H : D.ho) if (H.reduceRecorded) CU(cudaStreamWaitEvent(D.sCompute, H.evReduce, 0));
    return GS_OK;
}
static int sync_compute(DevSess& D) {
    CU(cudaStreamSynchronize(D.sCompute));
    if (D.overlap) CU(cudaStreamSynchronize(D.sReduce));
    return GS_OK;
}

struct gs_sess {
    gs_db* db = nullptr;
    gs_match_cfg cfg;
    std::vector<DevSess> devs;
    u64 nextTicket = 1;
    u64 launches = 0;
    bool finished = false;
    bool timing = false;
    int layout = GS_LAYOUT_TABLE;
    bool inlineSeen = false;  // unique k-mer bits are kept in the probe-table lines (leased from the database)
    bool dualBits = false;    // ... and mirrored into the compact bitset as they are set (gs_match_prepare_merge before the first batch):
                              // the merge then starts from a current bitset instead of streaming the whole table for its seen bits
    u64 batchesLaunched = 0;
    u64 nPos = 0;  // "storage positions" addressed by the unique-k-mer bitset: table slot ids or sorted-array indices
    // end-of-run merge across GPUs (merge_state)
    bool merged = false;
    double mergeMs = 0, mergeBitsetMs = 0;   // CUDA events on device 0's compute stream: whole merge / bitset part
    u64 mergeBytes = 0;                      // bitset bytes this rank fetched from the other ranks
    int mergePath = 0;                       // 1 = peer mappings read in place, 2 = ncclSend/ncclRecv slice exchange
    // gs_match_prepare_merge: peer mappings of the other ranks' bitsets / hit counters made ahead of the merge
    gs_comm* prepComm = nullptr;
    bool prepPeer = false;
    std::vector<void*> prepBits, prepHits;   // [world]
    std::vector<void*> prepOpened;           // mappings to close (cudaIpcCloseMemHandle)
    gsp::Packer* packer = nullptr;           // host_pack_threads != 0: made on the first gs_match_submit
    double packFrac = 0.7;                   // share of a batch's bases that goes over the link packed (the rest as ASCII)
    bool packSeeded = false;                 // the smoothed route costs hold a first measurement
    double packP = 0, packA = 0;             // smoothed seconds per byte: packing on the host's cores, ASCII on the link
    // search on the time per batch around the model's split (GS_PACK_SEARCH=0 switches it off): phases of four batches below / above it
    double packBias = 0, probeSum[2] = {0, 0}, probePrevBytes = 0;
    int probeN[2] = {0, 0}, probePhase = 0, probeIdx = -1;
    std::chrono::steady_clock::time_point probePrev;
    double packSeconds = 0;                  // host time spent packing (gs_match_pack_stats)
    u64 packBytes = 0, h2dBytes = 0;         // bases packed / bytes of base data put on the link
};

extern "C" void gs_match_cfg_default(gs_match_cfg* c) {
    if (!c) return;
    memset(c, 0, sizeof(*c));
    c->classify_reads = 1;              // C/GSConfigKey.java:302
    c->count_unique_kmers = 1;          // :305
    c->max_kmer_res_counts = 0;         // :344
    c->use_bloom_filter = 1;            // :320
    c->max_classification_paths = 10;   // :350
    c->min_kmers_for_class = 1;         // :341
    c->max_read_tax_error_count = -1;   // :328
    c->max_read_class_error_count = -1; // :337
    c->want_runs = 0;
    c->layout = GS_LAYOUT_TABLE;
    c->prefilter = 1;
    c->host_pack_threads = -1;
    c->host_pack_percent = -1;
}

static int sess_alloc_dev(gs_sess* s, DevSess& D) {
    const int V = s->db->V;
    CU(cudaSetDevice(D.dev));
    CU(cudaStreamCreateWithFlags(&D.sCopyIn, cudaStreamNonBlocking));
    {   // the label kernel's stream gets the greater priority: its CTAs are placed first and the reduce kernels of the batch
        // before take the slot per SM it leaves free (the other way round the reduce kernel fills the SMs for its 0.4 ms and the
        // label kernel waits: measured, no overlap at all)
        const char* e = getenv("GS_OVERLAP_REDUCE");
        D.overlap = !(e && atoi(e) == 0);
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // hi = numerically lowest = greatest priority
        CU(cudaStreamCreateWithPriority(&D.sCompute, cudaStreamNonBlocking, D.overlap ? hi : lo));
        if (D.overlap) CU(cudaStreamCreateWithPriority(&D.sReduce, cudaStreamNonBlocking, lo));
        else D.sReduce = D.sCompute;
    }
    CU(cudaStreamCreateWithFlags(&D.sCopyOut, cudaStreamNonBlocking));
    CU(dmalloc(&D.counters, (size_t)7 * V));
    CU(cudaMemset(D.counters, 0, std::max<size_t>((size_t)7 * V, 1) * sizeof(long long)));
    CU(dmalloc(&D.maxcontig, (size_t)V));
    CU(cudaMemset(D.maxcontig, 0, std::max<size_t>(V, 1) * sizeof(u64)));
    CU(dmalloc(&D.unique, (size_t)V));
    if (s->cfg.count_unique_kmers) {
        D.bitsetWords = (s->nPos + 63) / 64;
        if (s->inlineSeen) {
            gs_launch_table_clear_seen(s->db->d[D.devIndex].tab, s->nPos, 0);
            CU(cudaGetLastError());
            CU(cudaDeviceSynchronize());
        } else {
            CU(dmalloc(&D.bitset, D.bitsetWords));
            CU(cudaMemset(D.bitset, 0, std::max<u64>(D.bitsetWords, 1) * sizeof(u64)));
        }
        if (s->cfg.max_kmer_res_counts > 0) {
            CU(dmalloc(&D.hitCounts, s->nPos + 2));
            CU(cudaMemset(D.hitCounts, 0, (s->nPos + 2) * sizeof(uint16_t)));
        }
    }
    {
        for (DevSess::HandOver& H : D.ho) {
            CU(dmalloc(&H.overflowCount, 8));
            CU(cudaEventCreateWithFlags(&H.evLabel, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&H.evReduce, cudaEventDisableTiming));
        }
    }
    const int occ0 = std::max(1, gs_match_kernel_occupancy(0));
    D.fastBlocks = D.sms * occ0;
    const int labelOcc = std::max(1, gs_match_kernel_occupancy(s->layout == GS_LAYOUT_TABLE ? 3 : 4));
    D.labelBlocks = D.sms * labelOcc;
    if (const char* e = getenv("GS_DEBUG_LABEL_BLOCKS_PER_SM")) { const int v = atoi(e); if (v > 0) D.labelBlocks = D.sms * v; }  // tuning experiments
    D.labelBlocksOverlap = D.labelBlocks;   // GS_OVERLAP_SHORT=1 experiments: CTAs per SM the label kernel of a short-read batch gets
    if (const char* e = getenv("GS_OVERLAP_LABEL_BLOCKS_PER_SM")) { const int v = atoi(e); if (v > 0) D.labelBlocksOverlap = D.sms * v; }
    D.slowBlocks = std::max(1, D.sms / 4);
    // L2-persisting access window over the minimizer prefilter: the one structure every k-mer of every batch reads, next to
    // 600 MB of bases and 2 GB of labels per batch that stream through the same L2.  Set when the whole filter fits the
    // device's persisting carve-out (79 MB on B200).  GS_L2_PERSIST=0 switches the window off (A/B).
    {
        const gs_db* db = s->db;
        const char* e = getenv("GS_L2_PERSIST");
        const bool want = !(e && atoi(e) == 0);
        cudaDeviceProp prop;
        if (want && s->layout == GS_LAYOUT_TABLE && s->cfg.prefilter && db->mzBits && db->d[D.devIndex].mzFilter &&
            cudaGetDeviceProperties(&prop, D.dev) == cudaSuccess && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
            const size_t fBytes = (size_t)1 << (db->mzBits - 3);
            const size_t carve = fBytes;
            // measured (profiles/r02/bench_l_*.json): a filter that fits (32 MB at 1e8 k-mers) gains 3-4 %; of the 512 MB filter of
            // the 2e9-k-mer store only 79 MB can be pinned and the kernel gets 2.6 % slower -- no window then
            if (fBytes <= (size_t)prop.persistingL2CacheMaxSize && fBytes <= (size_t)prop.accessPolicyMaxWindowSize && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
                cudaStreamAttrValue av;
                memset(&av, 0, sizeof(av));
                av.accessPolicyWindow.base_ptr = (void*)db->d[D.devIndex].mzFilter;
                av.accessPolicyWindow.num_bytes = std::min<size_t>(fBytes, (size_t)prop.accessPolicyMaxWindowSize);
                av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)av.accessPolicyWindow.num_bytes);
                av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                if (cudaStreamSetAttribute(D.sCompute, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess) {
                    D.l2WindowBytes = av.accessPolicyWindow.num_bytes; D.l2CarveBytes = carve; D.l2HitRatio = av.accessPolicyWindow.hitRatio;
                }
            }
            cudaGetLastError();  // a refused window is not an error of the session
        }
    }
    CU(dmalloc(&D.slowTable, (size_t)D.slowBlocks * GS_WARPS_PER_BLOCK * 2 * std::max(V, 1)));
    for (MatchSlot& sl : D.slots) {
        CU(cudaEventCreateWithFlags(&sl.evH2D, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&sl.evCompute, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&sl.evDone, cudaEventDisableTiming));
        CU(cudaEventCreate(&sl.evAscii0)); CU(cudaEventCreate(&sl.evAscii1));
        CU(dmalloc(&sl.dEv, (size_t)std::max(V, 1)));
        CU(dmalloc(&sl.dNEv, 2));
        CU(cudaMallocHost((void**)&sl.hEv, std::max<size_t>(V, 1) * sizeof(gs_maxcontig_event)));
        CU(cudaMallocHost((void**)&sl.hNEv, 2 * sizeof(u32)));
    }
    return GS_OK;
}

extern "C" void gs_match_close(gs_sess* s) {
    if (!s) return;
    for (DevSess& D : s->devs) {
        cudaSetDevice(D.dev);
        cudaDeviceSynchronize();
        if (D.l2WindowBytes) cudaCtxResetPersistingL2Cache();   // hand the carve-out back: nothing of this session stays pinned in the L2
        for (MatchSlot& sl : D.slots) {
            cudaFree(sl.dBases); cudaFree(sl.dOffsets); cudaFree(sl.dOut); cudaFree(sl.dEv); cudaFree(sl.dNEv);
            cudaFree(sl.dKmerOff); cudaFree(sl.dRuns); cudaFree(sl.dRunCounts);
            cudaFree(sl.dCodes); cudaFree(sl.dValid);
            if (sl.hCodes) cudaFreeHost(sl.hCodes);
            if (sl.hValid) cudaFreeHost(sl.hValid);
            if (sl.hOut) cudaFreeHost(sl.hOut);
            if (sl.hEv) cudaFreeHost(sl.hEv);
            if (sl.hNEv) cudaFreeHost(sl.hNEv);
            if (sl.hKmerOff) cudaFreeHost(sl.hKmerOff);
            if (sl.hRunCounts) cudaFreeHost(sl.hRunCounts);
            text_stage_free(sl);
            if (sl.evH2D) cudaEventDestroy(sl.evH2D);
            if (sl.evCompute) cudaEventDestroy(sl.evCompute);
            if (sl.evDone) cudaEventDestroy(sl.evDone);
            if (sl.evAscii0) cudaEventDestroy(sl.evAscii0);
            if (sl.evAscii1) cudaEventDestroy(sl.evAscii1);
        }
        cudaFree(D.counters); cudaFree(D.maxcontig); cudaFree(D.bitset); cudaFree(D.hitCounts); cudaFree(D.unique); cudaFree(D.popPartial);
        cudaFree(D.slowTable);
        for (DevSess::HandOver& H : D.ho) {
            cudaFree(H.overflowList); cudaFree(H.overflowCount);
            cudaFree(H.labels); cudaFree(H.validBits); cudaFree(H.startBits); cudaFree(H.redoList); cudaFree(H.bmask);
            if (H.evLabel) cudaEventDestroy(H.evLabel);
            if (H.evReduce) cudaEventDestroy(H.evReduce);
        }
        if (D.overlap && D.sReduce) cudaStreamDestroy(D.sReduce);
        for (cudaEvent_t e : D.timingEv) cudaEventDestroy(e);
        D.timingEv.clear();
        if (D.sCopyIn) cudaStreamDestroy(D.sCopyIn);
        if (D.sCompute) cudaStreamDestroy(D.sCompute);
        if (D.sCopyOut) cudaStreamDestroy(D.sCopyOut);
    }
    if (s->inlineSeen) s->db->seenLeased = false;
    if (s->db->openSessions > 0) s->db->openSessions--;
    for (void* p : s->prepOpened) if (p) cudaIpcCloseMemHandle(p);
    delete s->packer;
    delete s;
}

extern "C" gs_sess* gs_match_open(gs_db* db, const gs_match_cfg* cfg) {
    if (!db || !db->finalized) { gs_fail(GS_ERR_STATE, "database not finalized"); return nullptr; }
    gs_match_cfg c;
    if (cfg) c = *cfg; else gs_match_cfg_default(&c);
    if (c.max_classification_paths < 1 || c.max_classification_paths > GS_MAX_PATHS) {
        gs_fail(GS_ERR_ARG, "max_classification_paths=%d out of range [1,%d]", c.max_classification_paths, GS_MAX_PATHS);
        return nullptr;
    }
    if (c.use_bloom_filter && !db->hasBloom) c.use_bloom_filter = 0;  // store without filter: KMerSortedArray.getLong :299 skips it
    if (c.layout != GS_LAYOUT_TABLE && c.layout != GS_LAYOUT_CLASSIC) { gs_fail(GS_ERR_ARG, "unknown layout %d", c.layout); return nullptr; }
    gs_sess* s = new gs_sess();
    s->db = db; s->cfg = c;
    db->openSessions++;
    s->layout = c.layout;
    s->nPos = c.layout == GS_LAYOUT_TABLE ? db->tabBuckets * GS_TAB_SLOT_STRIDE : db->n;
    if (c.layout == GS_LAYOUT_TABLE && c.count_unique_kmers && !db->seenLeased) { s->inlineSeen = true; db->seenLeased = true; }
    s->devs.resize(db->d.size());
    for (size_t i = 0; i < db->d.size(); i++) {
        s->devs[i].dev = db->d[i].dev; s->devs[i].devIndex = (int)i; s->devs[i].sms = db->ctx->sms[i];
        if (sess_alloc_dev(s, s->devs[i]) != GS_OK) { gs_match_close(s); return nullptr; }
    }
    cudaSetDevice(db->d[0].dev);
    return s;
}

static void fill_params(gs_sess* s, DevSess& D, GsMatchParams& P) {
    memset(&P, 0, sizeof(P));
    P.db = s->db->d[D.devIndex].view;
    if (!s->cfg.prefilter) P.db.mzFilter = nullptr;
    P.counters = D.counters; P.maxcontig = D.maxcontig; P.hitCounts = D.hitCounts;
    if (s->inlineSeen) { P.bitset = s->dualBits ? D.bitset : nullptr; P.seenTab = (u32*)s->db->d[D.devIndex].tab; }
    else { P.bitset = D.bitset; P.seenTab = nullptr; }
    P.classify = s->cfg.classify_reads ? 1 : 0;
    P.useBloom = s->cfg.use_bloom_filter ? 1 : 0;
    P.maxPaths = s->cfg.max_classification_paths;
    P.threshold = s->cfg.min_kmers_for_class;
    P.layout = s->layout;
    P.maxTaxErr = s->cfg.max_read_tax_error_count;
    P.maxClassErr = s->cfg.max_read_class_error_count;
    P.slowTable = D.slowTable;   // (the per-batch hand-over pointers are set by prepare_flat / launch_batch from the batch's set)
}

// flat geometry of a batch whose bases cover byte offsets [off0, off0 + nBytes) of P.bases; grows the hand-over buffers of set H
static int prepare_flat(DevSess& D, DevSess::HandOver& H, GsMatchParams& P, u64 off0, u64 nBytes) {
    P.off0 = off0;
    P.lead = P.packCodes ? 0u : (u32)((uintptr_t)(P.bases + off0) & 15);   // packed words start at the batch's first base
    P.flatLen = (u64)P.lead + nBytes;
    const u64 nSeg = (P.flatLen + GS_SEG_POS - 1) / GS_SEG_POS;
    const size_t words = (size_t)nSeg * GS_SEG_CHUNKS + 64;
    if (P.flatLen + 32 > H.labelsCap || words > H.validCap || words > H.startCap) {
        int rc = sync_compute(D);
        if (rc) return rc;
        CU(dgrow(&H.labels, &H.labelsCap, (size_t)P.flatLen + 32));
        CU(dgrow(&H.validBits, &H.validCap, words));
        CU(dgrow(&H.startBits, &H.startCap, words));
    }
    P.labels = H.labels; P.validBits = H.validBits; P.startBits = H.startBits;
    P.overflowCount = H.overflowCount; P.workCounter = H.overflowCount + 1; P.segCounter = H.overflowCount + 2;
    return GS_OK;
}

// Kernels of one batch.  Compute stream: read-start bitmap, label kernel.  Reduce stream (the compute stream itself when the
// overlap is off): reduce kernels (thread per read / warp per read, then the slow path over the overflow list), max-contig
// events, and whatever the caller appends through `tail` (it runs on the reduce stream behind them).  The batch uses one of the
// device's two hand-over sets; the set's previous user is two batches back, and the compute stream waits for that batch's
// reduce kernels before the label kernel overwrites the labels.
template <typename Tail>
static int launch_batch(gs_sess* s, DevSess& D, GsMatchParams& P, gs_maxcontig_event* dEv, u32* dNEv, u64 off0, u64 nBytes, Tail&& tail) {
    DevSess::HandOver& H = D.ho[D.overlap ? (D.batchNo & 1) : 0];
    D.batchNo++;
    s->batchesLaunched++;
    if (D.overlap && H.reduceRecorded) CU(cudaStreamWaitEvent(D.sCompute, H.evReduce, 0));
    if (P.nReads > H.ovCap || !H.overflowList) {
        int rcs = sync_compute(D);
        if (rcs) return rcs;
        CU(dgrow(&H.overflowList, &H.ovCap, (size_t)P.nReads));
    }
    P.overflowList = H.overflowList;
    int rc = prepare_flat(D, H, P, off0, nBytes);
    if (rc) return rc;
    CU(cudaMemsetAsync(H.overflowCount, 0, 8 * sizeof(u32), D.sCompute));
    if (dNEv) CU(cudaMemsetAsync(dNEv, 0, 2 * sizeof(u32), D.sCompute));
    P.errFlag = dNEv ? dNEv + 1 : nullptr;
    if (P.nReads == 0) {
        rc = tail(D.sCompute);
        if (rc) return rc;
        if (D.overlap) { CU(cudaEventRecord(H.evReduce, D.sCompute)); H.reduceRecorded = true; }
        return GS_OK;
    }
    const u64 nSeg = (P.flatLen + GS_SEG_POS - 1) / GS_SEG_POS;
    CU(cudaMemsetAsync(H.startBits, 0, ((size_t)nSeg * GS_SEG_CHUNKS + 64) * sizeof(u32), D.sCompute));
    gs_launch_mark_starts(P, D.sCompute);
    CU(cudaGetLastError());
    // short reads: one thread per read; reads it cannot take (too many taxa / too long) go to the warp-per-read kernel
    const bool threadPath = !s->cfg.want_runs && nBytes / P.nReads <= 512;
    static const bool noMask = getenv("GS_DEBUG_NO_BMASK") != nullptr;   // A/B: reduce kernels that walk every label
    if (!noMask) {  // the label kernel also marks the run boundaries: the reduce kernels visit boundaries, not labels
        const size_t words = (size_t)nSeg * GS_SEG_CHUNKS + 64;
        if (words > H.bmaskCap) {
            int rcs = sync_compute(D);
            if (rcs) return rcs;
            CU(dgrow(&H.bmask, &H.bmaskCap, words));
        }
        P.bmask = H.bmask;
    }
    if (threadPath && P.nReads > H.redoCap) {
        int rcs = sync_compute(D);
        if (rcs) return rcs;
        CU(dgrow(&H.redoList, &H.redoCap, (size_t)P.nReads));
    }
    // Which batches overlap (measured, profiles/r02/ab/bench_overlap_*.json): the warp-per-read reduce of long reads fills the
    // tail of the next label kernel -- the SMs whose persistent CTAs have run out of segments -- and the step gets 3.8 % shorter.
    // The thread-per-read reduce of short reads gains nothing there (its 0.38 ms just move), and leaving it a CTA slot per SM
    // costs the label kernel more (6.08 -> 6.50 ms) than it hides: short-read batches keep the stream order.
    // GS_OVERLAP_SHORT=1 overlaps them too (label kernel at GS_OVERLAP_LABEL_BLOCKS_PER_SM CTAs per SM, default all).
    static const bool overlapShort = [] { const char* e = getenv("GS_OVERLAP_SHORT"); return e && atoi(e) != 0; }();
    const bool ov = D.overlap && (!threadPath || overlapShort);
    const cudaStream_t sR = ov ? D.sReduce : D.sCompute;
    const int gridCap = (threadPath && ov) ? D.labelBlocksOverlap : D.labelBlocks;
    const int labelBlocks = (int)std::max<u64>(1, std::min<u64>((u64)gridCap, (nSeg + GS_WARPS_PER_BLOCK - 1) / GS_WARPS_PER_BLOCK));
    cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (s->timing) {
        for (int i = 0; i < 4; i++) { CU(cudaEventCreate(&tev[i])); D.timingEv.push_back(tev[i]); }
        CU(cudaEventRecord(tev[0], D.sCompute));
    }
    gs_launch_label(P, false, labelBlocks, D.sCompute);
    CU(cudaGetLastError());
    if (s->timing) CU(cudaEventRecord(tev[1], D.sCompute));
    if (ov) {
        CU(cudaEventRecord(H.evLabel, D.sCompute));
        CU(cudaStreamWaitEvent(sR, H.evLabel, 0));
    }
    if (s->timing) CU(cudaEventRecord(tev[2], sR));
    if (threadPath) {
        P.redoList = H.redoList; P.redoCount = H.overflowCount + 3; P.groupCounter = H.overflowCount + 4;
        gs_launch_reduce_thread(P, (int)std::min<u64>((u64)D.sms * 8, ((u64)P.nReads + 127) / 128), sR);
        CU(cudaGetLastError());
        s->launches += 1;
    }
    const int fastBlocks = threadPath ? D.slowBlocks : (int)std::min<u64>((u64)D.fastBlocks, ((u64)P.nReads + GS_WARPS_PER_BLOCK - 1) / GS_WARPS_PER_BLOCK);
    gs_launch_reduce(P, 0, false, fastBlocks, sR);
    CU(cudaGetLastError());
    gs_launch_reduce(P, 1, false, D.slowBlocks, sR);
    CU(cudaGetLastError());
    if (s->timing) CU(cudaEventRecord(tev[3], sR));
    s->launches += 4;
    if (dEv) {
        gs_launch_maxcontig_events(D.maxcontig, s->db->V, P.firstReadNo, P.nReads, dEv, dNEv, sR);
        CU(cudaGetLastError());
        s->launches += 1;
    }
    rc = tail(sR);
    if (rc) return rc;
    if (D.overlap) { CU(cudaEventRecord(H.evReduce, sR)); H.reduceRecorded = true; }
    return GS_OK;
}
static int launch_batch(gs_sess* s, DevSess& D, GsMatchParams& P, gs_maxcontig_event* dEv, u32* dNEv, u64 off0, u64 nBytes) {
    return launch_batch(s, D, P, dEv, dNEv, off0, nBytes, [](cudaStream_t) { return GS_OK; });
}

extern "C" int gs_match_submit(gs_sess* s, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads,
                               uint64_t first_read_no, gs_ticket* ticket) {
    if (!s || s->finished) return gs_fail(GS_ERR_STATE, "session missing or finished");
    if (!ticket || !offsets || (!bases && n_reads && offsets[n_reads] > offsets[0])) return gs_fail(GS_ERR_ARG, "null argument");
    if (first_read_no + n_reads > GS_ORDINAL_MASK) return gs_fail(GS_ERR_LIMIT, "read ordinal exceeds 2^40");
    const gs_ticket t = s->nextTicket;
    const size_t nDev = s->devs.size();
    DevSess& D = s->devs[(t - 1) % nDev];
    MatchSlot& sl = D.slots[((t - 1) / nDev) % GS_MAX_INFLIGHT];
    if (sl.pending) return gs_fail(GS_ERR_STATE, "more than %d batches in flight on a device: collect ticket %llu first", GS_MAX_INFLIGHT, (unsigned long long)sl.ticket);
    const u64 base0 = offsets[0];
    const u64 nBytes = offsets[n_reads] - base0;
    const int k = s->db->k;
    if (offsets[n_reads] < base0) return gs_fail(GS_ERR_ARG, "offsets not ascending");
    u64 totalKmers = 0;  // per-read offsets are validated on the device (errFlag); only want_runs needs a host pass
    if (s->cfg.want_runs)
        for (u32 i = 0; i < n_reads; i++) {
            if (offsets[i + 1] < offsets[i]) return gs_fail(GS_ERR_ARG, "offsets not ascending at read %u", i);
            const u64 L = offsets[i + 1] - offsets[i];
            if (L >= (u64)k) totalKmers += L - k + 1;
        }
    CU(cudaSetDevice(D.dev));
    CU(dgrow(&sl.dOffsets, &sl.offCap, (size_t)n_reads + 1));
    CU(dgrow(&sl.dOut, &sl.outCap, (size_t)n_reads));
    CU(hgrow(&sl.hOut, &sl.hOutCap, (size_t)n_reads));
    GsMatchParams P;
    fill_params(s, D, P);
    // inputs: host -> device on the copy-in stream (cudaMemcpyAsync; truly asynchronous for pinned buffers)
    // Two resources move a batch: the link (1 byte per base as ASCII, no CPU work) and the host's cores (which can pack the
    // bases to 0.375 bytes each, gs_pack.hpp).  Either alone is the bottleneck somewhere -- the link with one GPU, the cores
    // (host memory bandwidth, really) with eight -- so a batch is split at a segment boundary: the tail goes out as ASCII
    // first, its copy runs while the pool packs the head, then the packed words follow.  The label kernel stages a segment
    // from whichever form holds it.  The split follows the measured cost per byte of the two routes (gs_match_collect*: adaptive
    // unless host_pack_percent fixes it); results do not depend on it.
    const u64 nSegAll = (nBytes + GS_SEG_POS - 1) / GS_SEG_POS;
    u64 packSegs = 0;
    double probeOffset = 0;
    // The model balances this rank's cores against this rank's link.  When several ranks share one host's memory system that is
    // not the whole picture (the packed route costs 2.1 bytes of host memory traffic per base, the ASCII route 1.0), so the split
    // is also searched on what the caller sees, the time per byte from one submit to the next: phases of four batches a step
    // below and a step above the model's split, the last two of each phase counted, and a bias that moves towards the cheaper
    // side when the two differ by more than 1.5 %.  A caller that is slower than both routes sees no difference: the bias stays.
    // Measured: one GPU 64.7-64.8 G k-mers/s with the search, 64.5 G without (the bias ends at the upper clamp); four ranks on one
    // 32-core host 100.6 G with it, 74.5 G without (share 0.47 and still moving after 66 batches).  GS_PACK_SEARCH=0: model only.
    static const bool packSearch = [] { const char* e = getenv("GS_PACK_SEARCH"); return !(e && atoi(e) == 0); }();
    if (packSearch && s->cfg.host_pack_threads != 0 && s->cfg.host_pack_percent < 0 && nBytes >= (1u << 22)) {
        const double step = 0.08;
        const auto now = std::chrono::steady_clock::now();
        if (s->probeIdx >= 0 && s->probePrevBytes > 0 && s->probeIdx >= 2) {
            s->probeSum[s->probePhase] += std::chrono::duration<double>(now - s->probePrev).count() / s->probePrevBytes;
            s->probeN[s->probePhase]++;
        }
        if (++s->probeIdx == 4) {
            s->probeIdx = 0;
            s->probePhase ^= 1;
            if (s->probePhase == 0 && s->probeN[0] > 0 && s->probeN[1] > 0) {   // a low and a high phase are complete
                const double lo = s->probeSum[0] / s->probeN[0], hi = s->probeSum[1] / s->probeN[1];
                if (hi < lo * 0.985) s->packBias += step;
                else if (lo < hi * 0.985) s->packBias -= step;
                s->packBias = std::min(0.15, std::max(-0.9, s->packBias));
                s->probeSum[0] = s->probeSum[1] = 0; s->probeN[0] = s->probeN[1] = 0;
            }
        }
        s->probePrev = now; s->probePrevBytes = (double)nBytes;
        probeOffset = s->packBias + (s->probePhase ? step : -step);
    }
    if (s->cfg.host_pack_threads != 0 && nBytes > 0) {
        const double frac = s->cfg.host_pack_percent >= 0 ? std::min(100, s->cfg.host_pack_percent) / 100.0
                                                           : std::min(0.95, std::max(0.05, s->packFrac + probeOffset));
        packSegs = std::min<u64>(nSegAll, (u64)llround(frac * (double)nSegAll));
    }
    const u64 asciiFrom = std::min<u64>(nBytes, packSegs * GS_SEG_POS);   // first byte that travels as ASCII (a multiple of 16)
    sl.asciiBytes = nBytes - asciiFrom; sl.packedBases = 0; sl.packSec = 0;
    if (asciiFrom < nBytes) {
        CU(dgrow(&sl.dBases, &sl.basesCap, (size_t)nBytes + 64));   // sized for the whole batch: the split moves, the buffers stay
        CU(cudaEventRecord(sl.evAscii0, D.sCopyIn));
        CU(cudaMemcpyAsync(sl.dBases, bases + base0 + asciiFrom, nBytes - asciiFrom, cudaMemcpyHostToDevice, D.sCopyIn));
        CU(cudaEventRecord(sl.evAscii1, D.sCopyIn));
        s->h2dBytes += nBytes - asciiFrom;
        P.bases = sl.dBases - base0 - asciiFrom;   // byte (base0 + f) of the host buffer <-> P.bases + base0 + f, for f >= asciiFrom
    }
    if (packSegs) {
        // the kernel reads 32 words per segment of 31: one word past the last packed segment, zero words behind the batch
        if (!s->packer) s->packer = new gsp::Packer(s->cfg.host_pack_threads);
        const u64 wordsAll = (nBytes + 31) / 32;
        const size_t words = (size_t)std::min<u64>(wordsAll, packSegs * GS_SEG_CHUNKS + 2);
        const u64 packBases = std::min<u64>(nBytes, (u64)words * 32);
        CU(hgrow(&sl.hCodes, &sl.hCodesCap, (size_t)wordsAll));
        CU(hgrow(&sl.hValid, &sl.hValidCap, (size_t)wordsAll));
        CU(dgrow(&sl.dCodes, &sl.dCodesCap, (size_t)wordsAll + 64));
        CU(dgrow(&sl.dValid, &sl.dValidCap, (size_t)wordsAll + 64));
        // packed in a few pieces, each handed to the link as soon as it is ready: the copy of the packed words runs under the
        // packing of the next piece instead of behind all of it (2-3 ms less latency per batch, which counts with three
        // batches in flight)
        const auto t0 = std::chrono::steady_clock::now();
        const size_t nPieces = words >= ((size_t)1 << 20) ? 4 : 1;
        for (size_t pc = 0; pc < nPieces; pc++) {
            const size_t w0 = (words * pc / nPieces) & ~(size_t)1023, w1 = pc + 1 == nPieces ? words : ((words * (pc + 1) / nPieces) & ~(size_t)1023);
            if (w1 <= w0) continue;
            const u64 b0 = (u64)w0 * 32, b1 = std::min<u64>(packBases, (u64)w1 * 32);
            s->packer->pack(bases + base0 + b0, b1 > b0 ? b1 - b0 : 0, (uint64_t*)sl.hCodes + w0, sl.hValid + w0);
            CU(cudaMemcpyAsync(sl.dCodes + w0, sl.hCodes + w0, (w1 - w0) * sizeof(u64), cudaMemcpyHostToDevice, D.sCopyIn));
            CU(cudaMemcpyAsync(sl.dValid + w0, sl.hValid + w0, (w1 - w0) * sizeof(u32), cudaMemcpyHostToDevice, D.sCopyIn));
        }
        sl.packSec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        sl.packedBases = packBases;
        s->packSeconds += sl.packSec;
        s->packBytes += packBases; s->h2dBytes += words * 12;
        CU(cudaMemsetAsync(sl.dCodes + words, 0, 64 * sizeof(u64), D.sCopyIn));
        CU(cudaMemsetAsync(sl.dValid + words, 0, 64 * sizeof(u32), D.sCopyIn));
        P.packCodes = sl.dCodes; P.packValid = sl.dValid; P.packSegs = (u32)std::min<u64>(packSegs, 0xFFFFFFFFu);
    }
    CU(cudaMemcpyAsync(sl.dOffsets, offsets, ((size_t)n_reads + 1) * sizeof(u64), cudaMemcpyHostToDevice, D.sCopyIn));
    if (s->cfg.want_runs) {
        CU(hgrow(&sl.hKmerOff, &sl.hKmerOffCap, (size_t)n_reads + 1));
        u64 acc = 0;
        for (u32 i = 0; i < n_reads; i++) {
            sl.hKmerOff[i] = acc;
            const u64 L = offsets[i + 1] - offsets[i];
            if (L >= (u64)k) acc += L - k + 1;
        }
        sl.hKmerOff[n_reads] = acc;
        CU(dgrow(&sl.dKmerOff, &sl.kmerOffCap, (size_t)n_reads + 1));
        CU(dgrow(&sl.dRuns, &sl.runsCap, (size_t)totalKmers));
        CU(dgrow(&sl.dRunCounts, &sl.runCountsCap, (size_t)n_reads));
        CU(hgrow(&sl.hRunCounts, &sl.hRunCountsCap, (size_t)n_reads));
        CU(cudaMemcpyAsync(sl.dKmerOff, sl.hKmerOff, ((size_t)n_reads + 1) * sizeof(u64), cudaMemcpyHostToDevice, D.sCopyIn));
        CU(cudaMemsetAsync(sl.dRunCounts, 0, std::max<size_t>(n_reads, 1) * sizeof(u32), D.sCopyIn));
        P.runs = sl.dRuns; P.runOffsets = sl.dKmerOff; P.runsCap = totalKmers; P.runCounts = sl.dRunCounts;
    }
    CU(cudaEventRecord(sl.evH2D, D.sCopyIn));
    CU(cudaStreamWaitEvent(D.sCompute, sl.evH2D, 0));
    P.offsets = sl.dOffsets; P.nReads = n_reads; P.firstReadNo = first_read_no; P.out = sl.dOut;
    // offsets are relative to bases + offsets[0] on the device: the kernel subtracts nothing, so rebase here
    // (the device copy of the base stream starts at host offset base0)
    if (!P.bases) P.bases = (const uint8_t*)nullptr - base0;   // nothing travels as ASCII: only the alignment of P.bases + base0 is looked at
    int rc = launch_batch(s, D, P, sl.dEv, sl.dNEv, base0, nBytes, [&](cudaStream_t st) -> int { CU(cudaEventRecord(sl.evCompute, st)); return GS_OK; });
    if (rc) return rc;
    // results: device -> pinned host on the copy-out stream
    CU(cudaStreamWaitEvent(D.sCopyOut, sl.evCompute, 0));
    if (n_reads) CU(cudaMemcpyAsync(sl.hOut, sl.dOut, (size_t)n_reads * sizeof(gs_read_result), cudaMemcpyDeviceToHost, D.sCopyOut));
    CU(cudaMemcpyAsync(sl.hNEv, sl.dNEv, 2 * sizeof(u32), cudaMemcpyDeviceToHost, D.sCopyOut));
    CU(cudaMemcpyAsync(sl.hEv, sl.dEv, std::max<size_t>(s->db->V, 1) * sizeof(gs_maxcontig_event), cudaMemcpyDeviceToHost, D.sCopyOut));
    if (s->cfg.want_runs && n_reads) CU(cudaMemcpyAsync(sl.hRunCounts, sl.dRunCounts, (size_t)n_reads * sizeof(u32), cudaMemcpyDeviceToHost, D.sCopyOut));
    CU(cudaEventRecord(sl.evDone, D.sCopyOut));
    sl.pending = true; sl.ticket = t; sl.nReads = n_reads; sl.firstReadNo = first_read_no; sl.totalKmers = totalKmers; sl.isText = false;
    s->nextTicket++;
    *ticket = t;
    return GS_OK;
}

// Shared by the match and filter sessions: text chunk -> device, record splitter, verdict to the host (the only wait), then
// read offsets + compacted bases on the same stream.  On return info->status != 0 means "not strict 4-line FASTQ" (nothing
// else was done); otherwise sl.dBases / sl.dOffsets / sl.dRecs describe info->n_reads reads.
template <typename Slot>
static int text_stage(Slot& sl, cudaStream_t st, const uint8_t* text, u64 n_bytes, int k, gs_fastq_info* info, u64* launches) {
    const size_t lineCap = (size_t)(n_bytes / 16 + 64);
    const size_t nBlocks = (size_t)((n_bytes + GS_TEXT_SEG - 1) / GS_TEXT_SEG);
    CU(dgrow(&sl.dText, &sl.textCap, (size_t)n_bytes + 64));
    CU(dgrow(&sl.dLineEnd, &sl.lineCap, lineCap));
    CU(dgrow(&sl.dBlockCounts, &sl.blockCountsCap, nBlocks + 1));
    CU(dgrow(&sl.dRecs, &sl.recsCap, lineCap / 4 + 2));
    CU(dgrow(&sl.dLens, &sl.lensCap, lineCap / 4 + 2));
    if (!sl.dTextMeta) { CU(dmalloc(&sl.dTextMeta, 8)); CU(cudaMallocHost((void**)&sl.hTextMeta, 8 * sizeof(u32))); }
    CU(cudaMemsetAsync(sl.dTextMeta, 0, 8 * sizeof(u32), st));
    if (n_bytes) CU(cudaMemcpyAsync(sl.dText, text, n_bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(sl.dText + n_bytes, 0, 64, st));
    gs_launch_text_split(sl.dText, n_bytes, sl.dBlockCounts, sl.dLineEnd, (u32)std::min<size_t>(lineCap, 0xFFFFFFFFu), sl.dTextMeta, sl.dRecs, sl.dLens,
                         k, (unsigned long long*)(sl.dTextMeta + 4), st);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(sl.hTextMeta, sl.dTextMeta, 8 * sizeof(u32), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *launches += n_bytes ? 4 : 0;
    info->status = sl.hTextMeta[2];
    if (info->status) return GS_OK;
    const u32 n_reads = sl.hTextMeta[1];
    info->n_reads = n_reads;
    memcpy(&info->total_kmers, sl.hTextMeta + 4, sizeof(u64));
    memcpy(&info->total_bps, sl.hTextMeta + 6, sizeof(u64));
    CU(dgrow(&sl.dBases, &sl.basesCap, (size_t)info->total_bps + 64));
    CU(dgrow(&sl.dOffsets, &sl.offCap, (size_t)n_reads + 1));
    CU(hgrow(&sl.hRecs, &sl.hRecsCap, (size_t)n_reads + 1));
    CU(dgrow(&sl.dTileSums, &sl.tileSumsCap, (size_t)n_reads / 1024 + 2));
    if (n_reads == 0) CU(cudaMemsetAsync(sl.dOffsets, 0, sizeof(u64), st));
    gs_launch_text_compact(sl.dText, sl.dRecs, sl.dLens, n_reads, sl.dTileSums, sl.dOffsets, sl.dBases, st);
    CU(cudaGetLastError());
    *launches += n_reads ? 4 : 0;
    return GS_OK;
}
template <typename Slot>
static void text_stage_free(Slot& sl) {
    cudaFree(sl.dText); cudaFree(sl.dLineEnd); cudaFree(sl.dBlockCounts); cudaFree(sl.dTextMeta); cudaFree(sl.dRecs);
    cudaFree(sl.dLens); cudaFree(sl.dTileSums); cudaFree(sl.dEvHdr);
    if (sl.hTextMeta) cudaFreeHost(sl.hTextMeta);
    if (sl.hRecs) cudaFreeHost(sl.hRecs);
    if (sl.hEvHdr) cudaFreeHost(sl.hEvHdr);
}

// Raw FASTQ text: the chunk goes to the device as it is; the record splitter (gs_text.cu) runs on the copy-in stream right
// behind the copy, so it overlaps the match kernels of the previous batch; the host waits only for the splitter's verdict
// (number of reads, strict-format flags), then the bases are compacted and the usual kernels follow on the compute stream.
extern "C" int gs_match_submit_fastq(gs_sess* s, const uint8_t* text, uint64_t n_bytes, uint64_t first_read_no, gs_fastq_info* info,
                                     gs_ticket* ticket) {
    if (!s || s->finished) return gs_fail(GS_ERR_STATE, "session missing or finished");
    if (!ticket || !info || (!text && n_bytes)) return gs_fail(GS_ERR_ARG, "null argument");
    if (n_bytes >= 0xFFFFFF00ULL) return gs_fail(GS_ERR_LIMIT, "text chunk of %llu bytes (limit 2^32 - 256)", (unsigned long long)n_bytes);
    memset(info, 0, sizeof(*info));
    *ticket = 0;
    const gs_ticket t = s->nextTicket;
    const size_t nDev = s->devs.size();
    DevSess& D = s->devs[(t - 1) % nDev];
    MatchSlot& sl = D.slots[((t - 1) / nDev) % GS_MAX_INFLIGHT];
    if (sl.pending) return gs_fail(GS_ERR_STATE, "more than %d batches in flight on a device: collect ticket %llu first", GS_MAX_INFLIGHT, (unsigned long long)sl.ticket);
    const int V = s->db->V;
    CU(cudaSetDevice(D.dev));
    if (!sl.dEvHdr) { CU(dmalloc(&sl.dEvHdr, (size_t)std::max(V, 1))); CU(cudaMallocHost((void**)&sl.hEvHdr, (size_t)std::max(V, 1) * sizeof(u32))); }
    int rcs = text_stage(sl, D.sCopyIn, text, n_bytes, s->db->k, info, &s->launches);
    if (rcs) return rcs;
    if (info->status) return GS_OK;   // not strict 4-line FASTQ: the caller parses this chunk on the CPU (nothing is pending)
    const u32 n_reads = info->n_reads;
    if (first_read_no + n_reads > GS_ORDINAL_MASK) return gs_fail(GS_ERR_LIMIT, "read ordinal exceeds 2^40");
    const u64 nBytes = info->total_bps;
    CU(dgrow(&sl.dOut, &sl.outCap, (size_t)n_reads));
    CU(hgrow(&sl.hOut, &sl.hOutCap, (size_t)n_reads));
    GsMatchParams P;
    fill_params(s, D, P);
    if (s->cfg.want_runs) {  // contig runs of read i go to dRuns[kmerOff[i] ...): offsets by a scan of the k-mer counts on the device
        CU(dgrow(&sl.dKmerOff, &sl.kmerOffCap, (size_t)n_reads + 1));
        CU(hgrow(&sl.hKmerOff, &sl.hKmerOffCap, (size_t)n_reads + 1));
        CU(dgrow(&sl.dRuns, &sl.runsCap, (size_t)info->total_kmers));
        CU(dgrow(&sl.dRunCounts, &sl.runCountsCap, (size_t)n_reads));
        CU(hgrow(&sl.hRunCounts, &sl.hRunCountsCap, (size_t)n_reads));
        gs_launch_text_kmer_offsets(sl.dLens, n_reads, s->db->k, sl.dRunCounts, sl.dTileSums, sl.dKmerOff, D.sCopyIn);
        CU(cudaGetLastError());
        CU(cudaMemsetAsync(sl.dRunCounts, 0, std::max<size_t>(n_reads, 1) * sizeof(u32), D.sCopyIn));
        CU(cudaMemcpyAsync(sl.hKmerOff, sl.dKmerOff, ((size_t)n_reads + 1) * sizeof(u64), cudaMemcpyDeviceToHost, D.sCopyIn));
        P.runs = sl.dRuns; P.runOffsets = sl.dKmerOff; P.runsCap = info->total_kmers; P.runCounts = sl.dRunCounts;
        s->launches += n_reads ? 4 : 0;
    }
    CU(cudaEventRecord(sl.evH2D, D.sCopyIn));
    CU(cudaStreamWaitEvent(D.sCompute, sl.evH2D, 0));
    // the record table goes back while the match kernels run
    CU(cudaStreamWaitEvent(D.sCopyOut, sl.evH2D, 0));
    CU(cudaMemcpyAsync(sl.hRecs, sl.dRecs, ((size_t)n_reads + 1) * sizeof(gs_fastq_rec), cudaMemcpyDeviceToHost, D.sCopyOut));
    P.bases = sl.dBases; P.offsets = sl.dOffsets; P.nReads = n_reads; P.firstReadNo = first_read_no; P.out = sl.dOut;
    int rc = launch_batch(s, D, P, sl.dEv, sl.dNEv, 0, nBytes, [&](cudaStream_t st) -> int {
        gs_launch_text_event_headers(sl.dEv, sl.dNEv, (u32)std::max(V, 1), sl.dRecs, first_read_no, n_reads, sl.dEvHdr, st);
        CU(cudaGetLastError());
        s->launches += 1;
        CU(cudaEventRecord(sl.evCompute, st));
        return GS_OK;
    });
    if (rc) return rc;
    CU(cudaStreamWaitEvent(D.sCopyOut, sl.evCompute, 0));
    if (n_reads) CU(cudaMemcpyAsync(sl.hOut, sl.dOut, (size_t)n_reads * sizeof(gs_read_result), cudaMemcpyDeviceToHost, D.sCopyOut));
    CU(cudaMemcpyAsync(sl.hNEv, sl.dNEv, 2 * sizeof(u32), cudaMemcpyDeviceToHost, D.sCopyOut));
    CU(cudaMemcpyAsync(sl.hEv, sl.dEv, std::max<size_t>(V, 1) * sizeof(gs_maxcontig_event), cudaMemcpyDeviceToHost, D.sCopyOut));
    CU(cudaMemcpyAsync(sl.hEvHdr, sl.dEvHdr, std::max<size_t>(V, 1) * sizeof(u32), cudaMemcpyDeviceToHost, D.sCopyOut));
    if (s->cfg.want_runs && n_reads) CU(cudaMemcpyAsync(sl.hRunCounts, sl.dRunCounts, (size_t)n_reads * sizeof(u32), cudaMemcpyDeviceToHost, D.sCopyOut));
    CU(cudaEventRecord(sl.evDone, D.sCopyOut));
    sl.pending = true; sl.ticket = t; sl.nReads = n_reads; sl.firstReadNo = first_read_no; sl.totalKmers = info->total_kmers; sl.isText = true;
    s->nextTicket++;
    *ticket = t;
    return GS_OK;
}

static int wait_ticket(gs_sess* s, gs_ticket t, DevSess** Dout, MatchSlot** slOut) {
    if (!s) return gs_fail(GS_ERR_STATE, "null session");
    if (t == 0 || t >= s->nextTicket) return gs_fail(GS_ERR_STATE, "unknown ticket %llu", (unsigned long long)t);
    const size_t nDev = s->devs.size();
    DevSess& D = s->devs[(t - 1) % nDev];
    MatchSlot& sl = D.slots[((t - 1) / nDev) % GS_MAX_INFLIGHT];
    if (!sl.pending || sl.ticket != t) return gs_fail(GS_ERR_STATE, "ticket %llu is not pending", (unsigned long long)t);
    CU(cudaSetDevice(D.dev));
    CU(cudaEventSynchronize(sl.evDone));
    sl.pending = false;
    *Dout = &D; *slOut = &sl;
    if (!sl.isText && s->cfg.host_pack_threads != 0 && s->cfg.host_pack_percent < 0 && sl.packedBases >= (1u << 22) && sl.asciiBytes >= (1u << 22)) {
        // The split from the two routes' own costs, measured on every batch inside the pipeline: packing time p per byte on the
        // cores (next to the DMA traffic it competes with), copy time a per byte on the link.  Cores and link are busy equally
        // long when  p x = a (1 - x) + 0.375 a x,  i.e.  x = a / (p + 0.625 a);  p itself falls as x rises (less ASCII traffic
        // through the same DRAM), so the split is iterated on smoothed measurements until it stops moving.  (Round 2, first
        // version: seeded once from the first -- cold -- batch and then a blind hill-climb on the time per batch, which stayed
        // at 0.36 as often as it found 0.6-0.7: e2e 57 vs 61 G k-mers/s on the same box.)
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sl.evAscii0, sl.evAscii1) == cudaSuccess && ms > 0 && sl.packSec > 0) {
            const double pB = sl.packSec / (double)sl.packedBases, aB = ms * 1e-3 / (double)sl.asciiBytes;
            if (!s->packSeeded) { s->packP = pB; s->packA = aB; s->packSeeded = true; }
            else { s->packP += 0.3 * (pB - s->packP); s->packA += 0.3 * (aB - s->packA); }
            const double target = s->packA / (s->packP + 0.625 * s->packA);
            s->packFrac = std::min(0.92, std::max(0.08, s->packFrac + 0.5 * (target - s->packFrac)));
        } else {
            cudaGetLastError();
        }
    }
    if (sl.hNEv[1]) return gs_fail(GS_ERR_ARG, "batch of ticket %llu holds malformed read offsets (descending, or a read longer than 2^31 bases)", (unsigned long long)t);
    return GS_OK;
}

extern "C" int gs_match_collect_view(gs_sess* s, gs_ticket t, const gs_read_result** out, uint32_t* n_reads,
                                     const gs_maxcontig_event** events, uint32_t* n_events) {
    DevSess* D; MatchSlot* sl;
    int rc = wait_ticket(s, t, &D, &sl);
    if (rc) return rc;
    if (out) *out = sl->hOut;
    if (n_reads) *n_reads = sl->nReads;
    if (events) *events = sl->hEv;
    if (n_events) *n_events = sl->hNEv[0];
    return GS_OK;
}

// contig runs of a collected batch: run_offsets[n + 1] = prefix sums of the per-read run counts, runs = the lists back to back
static int collect_runs(MatchSlot& sl, uint64_t* run_offsets, gs_run* runs, uint64_t runs_cap) {
    u64 acc = 0;
    for (u32 i = 0; i < sl.nReads; i++) { run_offsets[i] = acc; acc += sl.hRunCounts[i]; }
    run_offsets[sl.nReads] = acc;
    if (runs) {
        if (acc > runs_cap) return gs_fail(GS_ERR_LIMIT, "%llu contig runs, capacity %llu", (unsigned long long)acc, (unsigned long long)runs_cap);
        // runs of read i sit at dRuns[kmerOff[i] .. + count): copy the dense prefix region back, then compact
        std::vector<gs_run> tmp((size_t)sl.totalKmers);
        if (sl.totalKmers) CU(cudaMemcpy(tmp.data(), sl.dRuns, (size_t)sl.totalKmers * sizeof(gs_run), cudaMemcpyDeviceToHost));
        for (u32 i = 0; i < sl.nReads; i++)
            if (sl.hRunCounts[i]) memcpy(runs + run_offsets[i], tmp.data() + sl.hKmerOff[i], (size_t)sl.hRunCounts[i] * sizeof(gs_run));
    }
    return GS_OK;
}

extern "C" int gs_match_collect_fastq(gs_sess* s, gs_ticket t, const gs_read_result** out, uint32_t* n_reads, const gs_maxcontig_event** events,
                                      const uint32_t** event_hdr_start, uint32_t* n_events, const gs_fastq_rec** recs,
                                      uint64_t* run_offsets, gs_run* runs, uint64_t runs_cap) {
    DevSess* D; MatchSlot* sl;
    int rc = wait_ticket(s, t, &D, &sl);
    if (rc) return rc;
    if (!sl->isText) return gs_fail(GS_ERR_STATE, "ticket %llu was not submitted as FASTQ text", (unsigned long long)t);
    if (s->cfg.want_runs && run_offsets) { rc = collect_runs(*sl, run_offsets, runs, runs_cap); if (rc) return rc; }
    if (out) *out = sl->hOut;
    if (n_reads) *n_reads = sl->nReads;
    if (events) *events = sl->hEv;
    if (event_hdr_start) *event_hdr_start = sl->hEvHdr;
    if (n_events) *n_events = sl->hNEv[0];
    if (recs) *recs = sl->hRecs;
    return GS_OK;
}

extern "C" int gs_match_collect(gs_sess* s, gs_ticket t, gs_read_result* out, gs_maxcontig_event* events, uint32_t ev_cap,
                                uint32_t* n_events, uint64_t* run_offsets, gs_run* runs, uint64_t runs_cap) {
    DevSess* Dp; MatchSlot* slp;
    int rc = wait_ticket(s, t, &Dp, &slp);
    if (rc) return rc;
    MatchSlot& sl = *slp;
    if (out && sl.nReads) memcpy(out, sl.hOut, (size_t)sl.nReads * sizeof(gs_read_result));
    if (n_events) {
        const u32 ne = sl.hNEv[0];
        if (events) {
            if (ne > ev_cap) return gs_fail(GS_ERR_LIMIT, "%u max-contig events, capacity %u", ne, ev_cap);
            memcpy(events, sl.hEv, (size_t)ne * sizeof(gs_maxcontig_event));
        }
        *n_events = ne;
    }
    if (s->cfg.want_runs && run_offsets) return collect_runs(sl, run_offsets, runs, runs_cap);
    return GS_OK;
}

extern "C" int gs_match_run_device(gs_sess* s, const uint8_t* d_bases, const uint64_t* d_offsets, uint32_t n_reads,
                                   uint64_t n_bases, uint64_t first_read_no, gs_read_result* d_out) {
    if (!s || s->finished) return gs_fail(GS_ERR_STATE, "session missing or finished");
    if (((uintptr_t)d_bases & 15) != 0) return gs_fail(GS_ERR_ARG, "d_bases must be 16-byte aligned");
    DevSess& D = s->devs[0];
    CU(cudaSetDevice(D.dev));
    GsMatchParams P;
    fill_params(s, D, P);
    P.bases = d_bases; P.offsets = (const u64*)d_offsets; P.nReads = n_reads; P.firstReadNo = first_read_no; P.out = d_out;
    return launch_batch(s, D, P, nullptr, nullptr, 0, n_bases);
}

extern "C" int gs_match_sync(gs_sess* s) {
    if (!s) return gs_fail(GS_ERR_STATE, "null session");
    for (DevSess& D : s->devs) {
        CU(cudaSetDevice(D.dev));
        CU(cudaStreamSynchronize(D.sCopyIn));
        { int rc = sync_compute(D); if (rc) return rc; }
        CU(cudaStreamSynchronize(D.sCopyOut));
    }
    return GS_OK;
}

// Orders the compute stream (gs_match_stream) behind everything submitted so far, without a host wait: a caller that times or
// chains device work on that stream (bench.py's CUDA events, a host that decodes on the GPU) calls this before its own event.
extern "C" int gs_match_join(gs_sess* s) {
    if (!s) return gs_fail(GS_ERR_STATE, "null session");
    for (DevSess& D : s->devs) { CU(cudaSetDevice(D.dev)); int rc = join_reduce(D); if (rc) return rc; }
    return GS_OK;
}

// In-line seen bits -> the session's compact bitset (same addressing as the external one: slot id = bucket * 16 + j)
static int materialize_bitset(gs_sess* s, DevSess& D) {
    if (!s->inlineSeen || s->dualBits) return GS_OK;   // dual mode: the label kernel has kept the compact bitset current
    CU(cudaSetDevice(D.dev));
    { int rc = join_reduce(D); if (rc) return rc; }
    if (!D.bitset) CU(dmalloc(&D.bitset, D.bitsetWords));
    gs_launch_table_extract_seen(s->db->d[D.devIndex].tab, s->nPos, D.bitset, D.sCompute);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(D.sCompute));
    s->launches += 1;
    return GS_OK;
}

extern "C" int gs_match_device_state(gs_sess* s, int64_t** counters, uint64_t** maxcontig, uint64_t** bitset, uint64_t* bitset_words) {
    if (!s) return gs_fail(GS_ERR_STATE, "null session");
    DevSess& D = s->devs[0];
    if (bitset && s->cfg.count_unique_kmers) {
        int rc = gs_match_sync(s);
        if (rc) return rc;
        rc = materialize_bitset(s, D);
        if (rc) return rc;
    }
    if (counters) *counters = (int64_t*)D.counters;
    if (maxcontig) *maxcontig = (uint64_t*)D.maxcontig;
    if (bitset) *bitset = (uint64_t*)D.bitset;
    if (bitset_words) *bitset_words = D.bitsetWords;
    return GS_OK;
}

extern "C" int gs_match_unique_popcount(gs_sess* s, const uint64_t* d_bitset, uint64_t word_begin, uint64_t word_end, int64_t* d_unique) {
    if (!s) return gs_fail(GS_ERR_STATE, "null session");
    DevSess& D = s->devs[0];
    CU(cudaSetDevice(D.dev));
    { int rc = join_reduce(D); if (rc) return rc; }
    CU(dgrow(&D.popPartial, &D.popPartialCap, (size_t)gs_popcount_scratch_words(D.sms, s->db->V)));
    gs_launch_unique_popcount((const u64*)d_bitset, word_begin, word_end, s->db->d[0].view, s->layout, (long long*)d_unique, D.popPartial, D.sms, D.sCompute);
    CU(cudaGetLastError());
    s->launches += 1;
    return GS_OK;
}

extern "C" void* gs_match_stream(gs_sess* s) { return s ? (void*)s->devs[0].sCompute : nullptr; }
extern "C" uint64_t gs_match_kernel_launches(const gs_sess* s) { return s ? s->launches : 0; }

static void timing_clear(DevSess& D) {
    for (cudaEvent_t e : D.timingEv) cudaEventDestroy(e);
    D.timingEv.clear();
}
extern "C" int gs_match_set_timing(gs_sess* s, int on) {
    if (!s) return gs_fail(GS_ERR_ARG, "null session");
    for (DevSess& D : s->devs) { CU(cudaSetDevice(D.dev)); int rc = sync_compute(D); if (rc) return rc; timing_clear(D); }
    s->timing = on != 0;
    return GS_OK;
}
extern "C" int gs_match_kernel_times(gs_sess* s, double* label_ms, double* reduce_ms, uint64_t* n_batches) {
    if (!s) return gs_fail(GS_ERR_ARG, "null session");
    double lab = 0, red = 0; u64 n = 0;
    for (DevSess& D : s->devs) {
        CU(cudaSetDevice(D.dev));
        { int rc = sync_compute(D); if (rc) return rc; }
        for (size_t i = 0; i + 3 < D.timingEv.size(); i += 4) {   // reduce: on its own stream, next to the following batch's label kernel when the overlap is on
            float a = 0, b = 0;
            CU(cudaEventElapsedTime(&a, D.timingEv[i], D.timingEv[i + 1]));
            CU(cudaEventElapsedTime(&b, D.timingEv[i + 2], D.timingEv[i + 3]));
            lab += a; red += b; n++;
        }
    }
    if (label_ms) *label_ms = n ? lab / (double)n : 0.0;
    if (reduce_ms) *reduce_ms = n ? red / (double)n : 0.0;
    if (n_batches) *n_batches = n;
    return GS_OK;
}

extern "C" int gs_match_dump_labels(gs_sess* s, const uint8_t* d_bases, const uint64_t* d_offsets, uint32_t n_reads,
                                    const uint64_t* d_kmer_offsets, int32_t* d_labels, int64_t* d_pos) {
    if (!s) return gs_fail(GS_ERR_STATE, "null session");
    DevSess& D = s->devs[0];
    CU(cudaSetDevice(D.dev));
    // side-effect free apart from the dump: private scratch accumulators
    const int V = s->db->V;
    long long* counters = nullptr; u64* maxcontig = nullptr; gs_read_result* out = nullptr; u32* ovList = nullptr; u32* ovCount = nullptr;
    long long* flatPos = nullptr;
    DevTemps temps;
    temps.watch(&counters); temps.watch(&maxcontig); temps.watch(&out); temps.watch(&ovList); temps.watch(&ovCount); temps.watch(&flatPos);
    CU(dmalloc(&counters, (size_t)7 * V)); CU(dmalloc(&maxcontig, (size_t)V)); CU(dmalloc(&out, (size_t)n_reads));
    CU(dmalloc(&ovList, (size_t)n_reads)); CU(dmalloc(&ovCount, 4));
    CU(cudaMemset(counters, 0, std::max<size_t>((size_t)7 * V, 1) * sizeof(long long)));
    CU(cudaMemset(maxcontig, 0, std::max<size_t>(V, 1) * sizeof(u64)));
    CU(cudaMemset(ovCount, 0, 4 * sizeof(u32)));
    GsMatchParams P;
    fill_params(s, D, P);
    P.counters = counters; P.maxcontig = maxcontig; P.bitset = nullptr; P.seenTab = nullptr; P.hitCounts = nullptr;
    P.overflowList = ovList; P.overflowCount = ovCount; P.workCounter = ovCount + 1;
    P.bases = d_bases; P.offsets = (const u64*)d_offsets; P.nReads = n_reads; P.firstReadNo = 0; P.out = out;
    P.kmerOffsets = (const u64*)d_kmer_offsets; P.dumpLabels = d_labels; P.dumpPos = (long long*)d_pos;
    { int rc = sync_compute(D); if (rc) return rc; }
    if (n_reads) {
        u64 ends[2] = {0, 0};
        CU(cudaMemcpy(&ends[0], d_offsets, sizeof(u64), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(&ends[1], d_offsets + n_reads, sizeof(u64), cudaMemcpyDeviceToHost));
        if (ends[1] < ends[0]) return gs_fail(GS_ERR_ARG, "offsets not ascending");
        int rc = prepare_flat(D, D.ho[0], P, ends[0], ends[1] - ends[0]);
        if (rc) return rc;
        P.overflowCount = ovCount; P.workCounter = ovCount + 1; P.segCounter = ovCount + 2;
        const u64 nSeg = (P.flatLen + GS_SEG_POS - 1) / GS_SEG_POS;
        CU(dmalloc(&flatPos, (size_t)P.flatLen + 32));
        P.flatPos = flatPos;
        CU(cudaMemsetAsync(D.ho[0].startBits, 0, ((size_t)nSeg * GS_SEG_CHUNKS + 64) * sizeof(u32), D.sCompute));
        gs_launch_mark_starts(P, D.sCompute);
        gs_launch_label(P, true, D.labelBlocks, D.sCompute);
        gs_launch_reduce(P, 0, true, D.fastBlocks, D.sCompute);
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(D.sCompute));
    return GS_OK;
}

// getMaxCountsCounts (C/store/KMerUniqueCounterBits.java:173-211): per value index the n largest per-position hit
// counters (Java shorts, signed compare, rows start as zeros) among the positions whose bit is set.
static void update_max_counts(int16_t count, int16_t* target, int n) {
    for (int j = 0; j < n; j++) {
        if (count > target[j]) {
            for (int q = n - 1; q > j; q--) target[q] = target[q - 1];
            target[j] = count;
            return;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// end-of-run merge across GPUs (the only exchange step of the path).  Stands in for the tail of
// FastqKMerMatcher.runMatcher (C/match/FastqKMerMatcher.java:199-234), where ONE statsIndex / KMerUniqueCounterBits has seen
// all reads: here every GPU has its share, so
//   counters  int64[7][V]  ncclAllReduce(sum)        maxcontig  uint64[V]  ncclAllReduce(max)  (ties -> lowest read ordinal)
//   bitset    rank r owns words [r * per, (r + 1) * per): one kernel ORs that slice of every rank's bitset -- read in place over
//             NVLink through peer mappings (same process: cudaDeviceEnablePeerAccess; other processes: cudaIpc handles), or,
//             where no mapping is possible, after an ncclSend/ncclRecv slice exchange -- and counts the merged words per value
//             index (KMerUniqueCounterBits.getUniqueKmerCounts :146-163); then ncclAllReduce(sum) of unique[V].
//   hit counters (maxKMerResCounts > 0): summed per slice the same way, then the merged slices are passed round so that every
//             rank holds the whole merged bitset and counters for getMaxCountsCounts (:173-199).
// Every rank ends up with the complete result.
// ---------------------------------------------------------------------------------------------------------
struct gs_comm {
    gs_ctx* ctx = nullptr;
    int world = 0;
    std::vector<ncclComm_t> comms;  // one per device of the context
    std::vector<int> ranks;
};

extern "C" int gs_comm_unique_id(uint8_t* id) {
    if (!id) return gs_fail(GS_ERR_ARG, "null argument");
    GsNccl* N = gs_nccl();
    if (!N) return gs_fail(GS_ERR_STATE, "NCCL is not available (libnccl.so.2 could not be loaded; set GS_NCCL_LIB)");
    static_assert(GS_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "gs_comm_unique_id hands out an ncclUniqueId");
    ncclUniqueId u;
    NC(N->GetUniqueId(&u));
    memcpy(id, u.internal, GS_COMM_ID_BYTES);
    return GS_OK;
}

extern "C" void gs_comm_destroy(gs_comm* cm) {
    if (!cm) return;
    GsNccl* N = gs_nccl();
    for (size_t i = 0; i < cm->comms.size(); i++)
        if (N && cm->comms[i]) { cudaSetDevice(cm->ctx->devs[i]); N->CommDestroy(cm->comms[i]); }
    delete cm;
}

static void comm_warm_up_peers(gs_comm* cm);
extern "C" gs_comm* gs_comm_create(gs_ctx* ctx, const uint8_t* id, int world, int rank) {
    if (!ctx || !id) { gs_fail(GS_ERR_ARG, "null argument"); return nullptr; }
    if (ctx->devs.size() != 1) { gs_fail(GS_ERR_ARG, "gs_comm_create joins ONE GPU per process to a job (context has %zu devices; their merge needs no communicator)", ctx->devs.size()); return nullptr; }
    if (world < 1 || world > GS_MAX_RANKS || rank < 0 || rank >= world) { gs_fail(GS_ERR_ARG, "rank %d of %d (at most %d ranks)", rank, world, GS_MAX_RANKS); return nullptr; }
    GsNccl* N = gs_nccl();
    if (!N) { gs_fail(GS_ERR_STATE, "NCCL is not available (libnccl.so.2 could not be loaded; set GS_NCCL_LIB)"); return nullptr; }
    CUP(cudaSetDevice(ctx->devs[0]));
    ncclUniqueId u;
    memcpy(u.internal, id, GS_COMM_ID_BYTES);
    ncclComm_t c = nullptr;
    ncclResult_t r = N->CommInitRank(&c, world, u, rank);
    if (r != ncclSuccess) { gs_fail(GS_ERR_CUDA, "ncclCommInitRank failed: %s", N->GetErrorString(r)); return nullptr; }
    gs_comm* cm = new gs_comm();
    cm->ctx = ctx; cm->world = world; cm->comms.push_back(c); cm->ranks.push_back(rank);
    // NCCL connects lazily: the first collective of every kind pays for the channel set-up (270 ms measured).  That belongs to
    // creating the communicator, not to the first merge.
    {
        long long* w = nullptr;
        if (dmalloc(&w, (size_t)world * 2) == cudaSuccess) {
            cudaMemset(w, 0, (size_t)world * 2 * sizeof(long long));
            N->AllReduce(w, w, (size_t)world, ncclInt64, ncclSum, c, 0);
            N->AllReduce(w, w, (size_t)world, ncclUint64, ncclMax, c, 0);
            N->AllGather(w + rank, w, 1, ncclInt64, c, 0);
            N->GroupStart();
            for (int q = 0; q < world; q++) { N->Send(w + q, 1, ncclInt64, q, c, 0); N->Recv(w + world + q, 1, ncclInt64, q, c, 0); }
            N->GroupEnd();
            cudaStreamSynchronize(0);
            cudaFree(w);
        }
    }
    comm_warm_up_peers(cm);
    return cm;
}
extern "C" int gs_comm_world(const gs_comm* cm) { return cm ? cm->world : 0; }
extern "C" int gs_comm_rank(const gs_comm* cm) { return cm && cm->ranks.size() == 1 ? cm->ranks[0] : -1; }

// communicator over the devices of a multi-GPU context (one process drives them all)
static gs_comm* ctx_all_comm(gs_ctx* ctx) {
    if (ctx->allComm) return ctx->allComm;
    GsNccl* N = gs_nccl();
    if (!N) { gs_fail(GS_ERR_STATE, "NCCL is not available (libnccl.so.2 could not be loaded; set GS_NCCL_LIB): a context with %zu devices needs it for the end-of-run merge", ctx->devs.size()); return nullptr; }
    if (ctx->devs.size() > GS_MAX_RANKS) { gs_fail(GS_ERR_LIMIT, "more than %d devices", GS_MAX_RANKS); return nullptr; }
    gs_comm* cm = new gs_comm();
    cm->ctx = ctx; cm->world = (int)ctx->devs.size();
    cm->comms.assign(ctx->devs.size(), nullptr);
    ncclResult_t r = N->CommInitAll(cm->comms.data(), (int)ctx->devs.size(), ctx->devs.data());
    if (r != ncclSuccess) { gs_fail(GS_ERR_CUDA, "ncclCommInitAll failed: %s", N->GetErrorString(r)); delete cm; return nullptr; }
    for (size_t i = 0; i < ctx->devs.size(); i++) cm->ranks.push_back((int)i);
    ctx->allComm = cm;
    return cm;
}

static void merge_slice(u64 n, int world, int rank, u64* per, u64* lo, u64* hi) {
    u64 p = (n + (u64)world - 1) / (u64)world;
    p = (p + 63) & ~63ULL;   // slices start at multiples of 64 words: 128-bit loads and 512-byte aligned exchanges
    *per = p;
    *lo = std::min(n, (u64)rank * p);
    *hi = std::min(n, *lo + p);
}

// peer mappings of one array of every rank (IPC handles exchanged over the communicator); closes them on destruction
struct IpcPeers {
    std::vector<void*> opened;
    ~IpcPeers() { for (void* p : opened) if (p) cudaIpcCloseMemHandle(p); }
};
// ptrs[q] = rank q's `mine` as seen from this process; *ok = 0 if a mapping failed somewhere (every rank gets the same verdict)
static int ipc_exchange(GsNccl* N, gs_comm* cm, cudaStream_t st, void* mine, void** ptrs, IpcPeers& keep, int* ok) {
    const int W = cm->world, r = cm->ranks[0];
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof(h));
    int good = cudaIpcGetMemHandle(&h, mine) == cudaSuccess ? 1 : 0;
    cudaGetLastError();
    uint8_t* dH = nullptr;
    int* dOk = nullptr;
    CU(dmalloc(&dH, sizeof(h) * (size_t)W)); CU(dmalloc(&dOk, 1));
    struct Free { void *a, *b; ~Free() { cudaFree(a); cudaFree(b); } } fr{dH, dOk};
    CU(cudaMemcpyAsync(dH + sizeof(h) * (size_t)r, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    NC(N->AllGather(dH + sizeof(h) * (size_t)r, dH, sizeof(h), ncclUint8, cm->comms[0], st));
    std::vector<cudaIpcMemHandle_t> all((size_t)W);
    CU(cudaMemcpyAsync(all.data(), dH, sizeof(h) * (size_t)W, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int q = 0; q < W && good; q++) {
        if (q == r) { ptrs[q] = mine; continue; }
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[(size_t)q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); good = 0; break; }
        keep.opened.push_back(p);
        ptrs[q] = p;
    }
    CU(cudaMemcpyAsync(dOk, &good, sizeof(int), cudaMemcpyHostToDevice, st));
    NC(N->AllReduce(dOk, dOk, 1, ncclInt32, ncclMin, cm->comms[0], st));
    CU(cudaMemcpyAsync(&good, dOk, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *ok = good;
    return GS_OK;
}

// The first peer mapping between two processes enables peer access between their contexts (74 ms measured); that is
// communicator set-up too.  A throw-away exchange of a small buffer pays for it here.
static void comm_warm_up_peers(gs_comm* cm) {
    GsNccl* N = gs_nccl();
    if (!N || cm->world < 2 || cm->comms.size() != 1) return;
    void* buf = nullptr;
    if (cudaMalloc(&buf, 256) != cudaSuccess) { cudaGetLastError(); return; }
    std::vector<void*> ptrs((size_t)cm->world, nullptr);
    int ok = 0;
    {
        IpcPeers keep;
        ipc_exchange(N, cm, 0, buf, ptrs.data(), keep, &ok);
    }
    // every rank has closed its mappings before anybody frees: one more collective as the barrier
    int* d = nullptr;
    if (cudaMalloc((void**)&d, sizeof(int)) == cudaSuccess) {
        cudaMemset(d, 0, sizeof(int));
        N->AllReduce(d, d, 1, ncclInt32, ncclSum, cm->comms[0], 0);
        cudaStreamSynchronize(0);
        cudaFree(d);
    }
    cudaFree(buf);
}

static int materialize_bitset(gs_sess* s, DevSess& D);
extern "C" int gs_match_prepare_merge(gs_sess* s, gs_comm* cm) {
    if (!s || !cm) return gs_fail(GS_ERR_ARG, "null argument");
    if (s->finished || s->merged) return gs_fail(GS_ERR_STATE, "session already finished");
    if (s->devs.size() != 1 || cm->comms.size() != 1) return GS_OK;   // one process, several GPUs: the devices address each other directly
    if (cm->ctx != s->db->ctx) return gs_fail(GS_ERR_ARG, "communicator and session belong to different contexts");
    if (s->prepComm == cm) return GS_OK;
    GsNccl* N = gs_nccl();
    if (!N) return gs_fail(GS_ERR_STATE, "NCCL is not available");
    DevSess& D = s->devs[0];
    CU(cudaSetDevice(D.dev));
    const int W = cm->world;
    s->prepBits.assign((size_t)W, nullptr); s->prepHits.assign((size_t)W, nullptr);
    s->prepPeer = false;
    if (s->cfg.count_unique_kmers) {
        if (!D.bitset) {   // in-line seen bits: the compact bitset only exists from the end of the run on, but its address can be handed out now
            CU(dmalloc(&D.bitset, D.bitsetWords));
            CU(cudaMemsetAsync(D.bitset, 0, std::max<u64>(D.bitsetWords, 1) * sizeof(u64), D.sCompute));
            // no batch has run yet: from here on every first-time seen bit also goes into the compact bitset (one more RED per
            // distinct k-mer of the run), and the merge needs no pass over the table (5.5 ms for the 34 GB table of the 2e9-k-mer
            // store).  GS_MERGE_DUAL_BITS=0: extract at the end as before (A/B).
            const char* e = getenv("GS_MERGE_DUAL_BITS");
            if (s->inlineSeen && s->batchesLaunched == 0 && !(e && atoi(e) == 0)) s->dualBits = true;
        }
        const char* force = getenv("GS_MERGE_PATH");
        if (!(force && strcmp(force, "nccl") == 0)) {
            IpcPeers keep;
            int ok = 0;
            int rc = ipc_exchange(N, cm, D.sCompute, D.bitset, s->prepBits.data(), keep, &ok);
            if (rc) return rc;
            if (ok && D.hitCounts) { rc = ipc_exchange(N, cm, D.sCompute, D.hitCounts, s->prepHits.data(), keep, &ok); if (rc) return rc; }
            s->prepPeer = ok != 0;
            s->prepOpened.swap(keep.opened);   // the session keeps the mappings until the merge is over
            if (!s->prepPeer) { for (void* p : s->prepOpened) if (p) cudaIpcCloseMemHandle(p); s->prepOpened.clear(); }
        }
        // the merge kernel: scratch allocated and the function loaded (an empty slice) before the run, not at its end
        CU(dgrow(&D.popPartial, &D.popPartialCap, (size_t)gs_popcount_scratch_words(D.sms, s->db->V)));
        GsPeerPtrs none;
        memset(&none, 0, sizeof(none));
        gs_launch_merge_or_popcount(none, 0, D.bitset, 0, 0, s->db->d[D.devIndex].view, s->layout, D.unique, D.popPartial, 1, D.sCompute);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(D.sCompute));
    }
    {   // the merge's collectives once at their own sizes, types and operators on scratch: NCCL sets up the protocol and the
        // channels a message size needs the first time it sees it (measured on 8 GPUs: 20 ms inside the first merge of the
        // 2e9-k-mer store's 54 k value indices after two workloads with 11 k), and that belongs in front of the run
        const int V = s->db->V;
        long long* scratch = nullptr;
        CU(dmalloc(&scratch, (size_t)7 * std::max(V, 1)));
        CU(cudaMemsetAsync(scratch, 0, (size_t)7 * std::max(V, 1) * sizeof(long long), D.sCompute));
        int rcn = GS_OK;
        auto nc = [&](ncclResult_t r) { if (r != ncclSuccess && rcn == GS_OK) rcn = gs_fail(GS_ERR_CUDA, "NCCL warm-up: %s", N->GetErrorString(r)); };
        nc(N->AllReduce(scratch, scratch, (size_t)7 * V, ncclInt64, ncclSum, cm->comms[0], D.sCompute));
        nc(N->AllReduce(scratch, scratch, (size_t)V, ncclUint64, ncclMax, cm->comms[0], D.sCompute));
        nc(N->AllReduce(scratch, scratch, (size_t)V, ncclInt64, ncclSum, cm->comms[0], D.sCompute));
        cudaStreamSynchronize(D.sCompute);
        cudaFree(scratch);
        if (rcn) return rcn;
    }
    s->prepComm = cm;
    return GS_OK;
}

static int merge_state(gs_sess* s, gs_comm* cm) {
    GsNccl* N = gs_nccl();
    if (!N) return gs_fail(GS_ERR_STATE, "NCCL is not available");
    if (cm->ctx != s->db->ctx || cm->comms.size() != s->devs.size()) return gs_fail(GS_ERR_ARG, "communicator and session belong to different contexts");
    const int V = s->db->V, W = cm->world, L = (int)s->devs.size();
    const bool uniq = s->cfg.count_unique_kmers != 0;
    int rc;
    DevSess& D0 = s->devs[0];
    for (DevSess& D : s->devs) { CU(cudaSetDevice(D.dev)); int rcj = join_reduce(D); if (rcj) return rcj; }   // the merge follows every batch's reduce kernels
    CU(cudaSetDevice(D0.dev));
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1)); CU(cudaEventCreate(&e2));
    struct Ev { cudaEvent_t a, b, c; ~Ev() { cudaEventDestroy(a); cudaEventDestroy(b); cudaEventDestroy(c); } } evs{e0, e1, e2};
    CU(cudaEventRecord(e0, D0.sCompute));
    static const bool dbgMerge = getenv("GS_DEBUG_MERGE") != nullptr;   // phase times on stderr (each phase drained: not for timing runs)
    auto tPrev = std::chrono::steady_clock::now();
    auto stamp = [&](const char* what) {
        if (!dbgMerge) return;
        cudaStreamSynchronize(D0.sCompute);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[gs merge rank %d] %-28s %8.3f ms\n", cm->ranks[0], what, std::chrono::duration<double, std::milli>(now - tPrev).count());
        tPrev = now;
    };
    stamp("drain before merge");
    // in-line seen bits (probe table entries) -> compact bitset: part of the end-of-run work, so inside the measured span
    if (uniq) for (DevSess& D : s->devs) { rc = materialize_bitset(s, D); if (rc) return rc; }
    CU(cudaSetDevice(D0.dev));
    stamp("seen bits -> bitset");
    // ---- counters and max-contigs
    NC(N->GroupStart());
    for (int l = 0; l < L; l++) {
        DevSess& D = s->devs[l];
        NC(N->AllReduce(D.counters, D.counters, (size_t)7 * V, ncclInt64, ncclSum, cm->comms[l], D.sCompute));
        NC(N->AllReduce(D.maxcontig, D.maxcontig, (size_t)V, ncclUint64, ncclMax, cm->comms[l], D.sCompute));
    }
    NC(N->GroupEnd());
    stamp("all-reduce counters/maxcontig");
    s->mergeBytes = 0; s->mergePath = 0;
    if (uniq && D0.bitset) {
        CU(cudaSetDevice(D0.dev));
        CU(cudaEventRecord(e1, D0.sCompute));
        const u64 words = D0.bitsetWords;
        // ---- where do the other ranks' bitsets come from?
        const char* force = getenv("GS_MERGE_PATH");   // "nccl": slice exchange even where peer mappings exist (A/B measurements)
        bool peer = !(force && strcmp(force, "nccl") == 0);
        std::vector<std::vector<void*>> bitPtr((size_t)L, std::vector<void*>((size_t)W, nullptr)), hitPtr = bitPtr;
        IpcPeers keep;
        if (peer) {
            if (L == W) {   // one process: every device's arrays are directly addressable
                peer = cm->ctx->peerAll;
                for (int l = 0; l < L && peer; l++) for (int q = 0; q < W; q++) { bitPtr[l][q] = s->devs[q].bitset; hitPtr[l][q] = s->devs[q].hitCounts; }
            } else if (s->prepComm == cm) {   // mapped ahead of the run (gs_match_prepare_merge)
                peer = s->prepPeer;
                for (int q = 0; q < W; q++) { bitPtr[0][q] = s->prepBits[(size_t)q]; hitPtr[0][q] = s->prepHits[(size_t)q]; }
            } else {        // one process per GPU: map the other ranks' arrays through IPC handles
                int ok = 0;
                rc = ipc_exchange(N, cm, D0.sCompute, D0.bitset, bitPtr[0].data(), keep, &ok);
                if (rc) return rc;
                if (ok && D0.hitCounts) { rc = ipc_exchange(N, cm, D0.sCompute, D0.hitCounts, hitPtr[0].data(), keep, &ok); if (rc) return rc; }
                peer = ok != 0;
            }
        }
        s->mergePath = peer ? 1 : 2;
        stamp("peer mappings");
        // ---- bitset: OR of the own slice over all ranks + per-taxon popcount, then the sum of the slices' counts
        std::vector<u64*> recv((size_t)L, nullptr);
        struct FreeAll { std::vector<u64*>& v; std::vector<DevSess>& d; ~FreeAll() { for (size_t i = 0; i < v.size(); i++) if (v[i]) { cudaSetDevice(d[i].dev); cudaFree(v[i]); } } } fr{recv, s->devs};
        u64 per = 0, lo = 0, hi = 0;
        if (!peer) {
            for (int l = 0; l < L; l++) {
                merge_slice(words, W, cm->ranks[l], &per, &lo, &hi);
                CU(cudaSetDevice(s->devs[l].dev));
                CU(dmalloc(&recv[l], per * (u64)W));
            }
            NC(N->GroupStart());
            for (int l = 0; l < L; l++) {
                DevSess& D = s->devs[l];
                merge_slice(words, W, cm->ranks[l], &per, &lo, &hi);
                for (int q = 0; q < W; q++) {
                    u64 pq, loq, hiq;
                    merge_slice(words, W, q, &pq, &loq, &hiq);
                    if (hiq > loq) NC(N->Send(D.bitset + loq, (size_t)(hiq - loq), ncclUint64, q, cm->comms[l], D.sCompute));
                    if (hi > lo) NC(N->Recv(recv[l] + (u64)q * per, (size_t)(hi - lo), ncclUint64, q, cm->comms[l], D.sCompute));
                }
            }
            NC(N->GroupEnd());
        }
        for (int l = 0; l < L; l++) {
            DevSess& D = s->devs[l];
            merge_slice(words, W, cm->ranks[l], &per, &lo, &hi);
            CU(cudaSetDevice(D.dev));
            GsPeerPtrs src;
            memset(&src, 0, sizeof(src));
            for (int q = 0; q < W; q++) src.p[q] = peer ? (const u64*)bitPtr[l][q] : recv[l] + (u64)q * per - lo;
            CU(cudaMemsetAsync(D.unique, 0, std::max<size_t>(V, 1) * sizeof(long long), D.sCompute));
            CU(dgrow(&D.popPartial, &D.popPartialCap, (size_t)gs_popcount_scratch_words(D.sms, V)));
            gs_launch_merge_or_popcount(src, W, D.bitset, lo, hi, s->db->d[D.devIndex].view, s->layout, D.unique, D.popPartial, D.sms, D.sCompute);
            CU(cudaGetLastError());
            s->launches += 1;
            if (l == 0) s->mergeBytes += (hi - lo) * 8 * (u64)(W - 1);
        }
        stamp("OR + popcount kernel");
        NC(N->GroupStart());
        for (int l = 0; l < L; l++) NC(N->AllReduce(s->devs[l].unique, s->devs[l].unique, (size_t)V, ncclInt64, ncclSum, cm->comms[l], s->devs[l].sCompute));
        NC(N->GroupEnd());
        stamp("all-reduce unique");
        // ---- per-position hit counters: slice sums, then every rank collects the merged slices of bitset and counters
        if (D0.hitCounts) {
            std::vector<uint16_t*> hrecv((size_t)L, nullptr);
            struct FreeH { std::vector<uint16_t*>& v; std::vector<DevSess>& d; ~FreeH() { for (size_t i = 0; i < v.size(); i++) if (v[i]) { cudaSetDevice(d[i].dev); cudaFree(v[i]); } } } frh{hrecv, s->devs};
            const u64 nPos = s->nPos;
            auto posSlice = [&](int rank, u64* a, u64* b) { u64 p, l0, h0; merge_slice(words, W, rank, &p, &l0, &h0); *a = std::min(nPos, l0 * 64); *b = std::min(nPos, h0 * 64); };
            u64 a = 0, b = 0;
            if (!peer) {
                for (int l = 0; l < L; l++) { CU(cudaSetDevice(s->devs[l].dev)); CU(dmalloc(&hrecv[l], per * 64 * (u64)W)); }
                NC(N->GroupStart());
                for (int l = 0; l < L; l++) {
                    DevSess& D = s->devs[l];
                    posSlice(cm->ranks[l], &a, &b);
                    for (int q = 0; q < W; q++) {
                        u64 aq, bq;
                        posSlice(q, &aq, &bq);
                        if (bq > aq) NC(N->Send(D.hitCounts + aq, (size_t)(bq - aq) * 2, ncclUint8, q, cm->comms[l], D.sCompute));
                        if (b > a) NC(N->Recv(hrecv[l] + (u64)q * per * 64, (size_t)(b - a) * 2, ncclUint8, q, cm->comms[l], D.sCompute));
                    }
                }
                NC(N->GroupEnd());
            }
            for (int l = 0; l < L; l++) {
                DevSess& D = s->devs[l];
                posSlice(cm->ranks[l], &a, &b);
                CU(cudaSetDevice(D.dev));
                GsPeerPtrs src;
                memset(&src, 0, sizeof(src));
                for (int q = 0; q < W; q++) src.p[q] = peer ? (const u64*)hitPtr[l][q] : (const u64*)(hrecv[l] + (u64)q * per * 64 - a);
                gs_launch_merge_add_u16(src, W, D.hitCounts, a, b, D.sCompute);
                CU(cudaGetLastError());
                s->launches += 1;
            }
            // (rank q's kernels above have read every other rank's slice q before q sends its merged slice: stream order)
            NC(N->GroupStart());
            for (int l = 0; l < L; l++) {
                DevSess& D = s->devs[l];
                merge_slice(words, W, cm->ranks[l], &per, &lo, &hi);
                posSlice(cm->ranks[l], &a, &b);
                for (int q = 0; q < W; q++) {
                    if (q == cm->ranks[l]) continue;
                    u64 pq, loq, hiq, aq, bq;
                    merge_slice(words, W, q, &pq, &loq, &hiq);
                    posSlice(q, &aq, &bq);
                    if (hi > lo) NC(N->Send(D.bitset + lo, (size_t)(hi - lo), ncclUint64, q, cm->comms[l], D.sCompute));
                    if (hiq > loq) NC(N->Recv(D.bitset + loq, (size_t)(hiq - loq), ncclUint64, q, cm->comms[l], D.sCompute));
                    if (b > a) NC(N->Send(D.hitCounts + a, (size_t)(b - a) * 2, ncclUint8, q, cm->comms[l], D.sCompute));
                    if (bq > aq) NC(N->Recv(D.hitCounts + aq, (size_t)(bq - aq) * 2, ncclUint8, q, cm->comms[l], D.sCompute));
                }
            }
            NC(N->GroupEnd());
        }
    }
    CU(cudaSetDevice(D0.dev));
    if (!(uniq && D0.bitset)) CU(cudaEventRecord(e1, D0.sCompute));
    CU(cudaEventRecord(e2, D0.sCompute));
    // Once these streams have drained nobody reads this rank's arrays any more: the all-reduce of `unique` (or the last slice
    // exchange) on rank q sits behind q's merge kernels in stream order, and it cannot complete here before q has reached it.
    for (DevSess& D : s->devs) { CU(cudaSetDevice(D.dev)); CU(cudaStreamSynchronize(D.sCompute)); }
    float msAll = 0, msBits = 0;
    CU(cudaSetDevice(D0.dev));
    CU(cudaEventElapsedTime(&msAll, e0, e2));
    CU(cudaEventElapsedTime(&msBits, e1, e2));
    s->mergeMs = msAll; s->mergeBitsetMs = msBits;
    s->merged = true;
    for (void* p : s->prepOpened) if (p) cudaIpcCloseMemHandle(p);
    s->prepOpened.clear(); s->prepComm = nullptr;
    return GS_OK;
}

// results of the finished run from device 0 (after the merge every device holds the same totals)
static int read_out(gs_sess* s, gs_taxon_counts* counts, int16_t* top_counts, bool popcountLocal) {
    const int V = s->db->V;
    DevSess& D0 = s->devs[0];
    CU(cudaSetDevice(D0.dev));
    std::vector<long long> acc((size_t)7 * V, 0), uniq((size_t)V, 0);
    std::vector<u64> mc((size_t)V, 0);
    CU(cudaMemcpy(acc.data(), D0.counters, (size_t)7 * V * sizeof(long long), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(mc.data(), D0.maxcontig, (size_t)V * sizeof(u64), cudaMemcpyDeviceToHost));
    if (D0.bitset) {
        if (popcountLocal) {
            CU(cudaMemsetAsync(D0.unique, 0, std::max<size_t>(V, 1) * sizeof(long long), D0.sCompute));
            CU(dgrow(&D0.popPartial, &D0.popPartialCap, (size_t)gs_popcount_scratch_words(D0.sms, V)));
            gs_launch_unique_popcount(D0.bitset, 0, D0.bitsetWords, s->db->d[0].view, s->layout, D0.unique, D0.popPartial, D0.sms, D0.sCompute);
            CU(cudaGetLastError());
            s->launches += 1;
            CU(cudaStreamSynchronize(D0.sCompute));
        }
        CU(cudaMemcpy(uniq.data(), D0.unique, (size_t)V * sizeof(long long), cudaMemcpyDeviceToHost));
    }
    if (counts) {
        for (int v = 0; v < V; v++) {
            gs_taxon_counts& c = counts[v];
            c.kmers = acc[0 * (size_t)V + v]; c.contigs = acc[1 * (size_t)V + v]; c.contig_len_squared_sum = acc[2 * (size_t)V + v];
            c.reads_1kmer = acc[3 * (size_t)V + v]; c.reads = acc[4 * (size_t)V + v]; c.reads_kmers = acc[5 * (size_t)V + v];
            c.reads_bps = acc[6 * (size_t)V + v];
            c.unique_kmers = D0.bitset ? uniq[v] : -1;  // FastqKMerMatcher.java:224-229
            c.max_contig_len = (int32_t)(mc[v] >> GS_MAXCONTIG_SHIFT);
            c.max_contig_read_no = mc[v] ? GS_ORDINAL_MASK - (mc[v] & GS_ORDINAL_MASK) : 0;
            c.touched = (c.reads_1kmer > 0 || c.reads > 0) ? 1 : 0;  // getCountsPerTaxid was called (:545-556)
        }
    }
    if (top_counts && D0.hitCounts && s->cfg.max_kmer_res_counts > 0) {
        const int nTop = s->cfg.max_kmer_res_counts;
        memset(top_counts, 0, (size_t)(V + 1) * nTop * sizeof(int16_t));
        u64 total = 0;
        for (int v = 0; v < V; v++) total += (u64)uniq[v];
        u32* dHits = nullptr; unsigned long long* dN = nullptr;
        CU(dmalloc(&dHits, (size_t)total)); CU(dmalloc(&dN, 1));
        struct Free { void *a, *b; ~Free() { cudaFree(a); cudaFree(b); } } fr{dHits, dN};
        CU(cudaMemsetAsync(dN, 0, sizeof(unsigned long long), D0.sCompute));
        gs_launch_collect_hits(D0.bitset, D0.bitsetWords, D0.hitCounts, s->db->d[0].view, s->layout, dHits, dN, total, D0.sCompute);
        CU(cudaGetLastError());
        s->launches += 1;
        CU(cudaStreamSynchronize(D0.sCompute));
        unsigned long long got = 0;
        CU(cudaMemcpy(&got, dN, sizeof(got), cudaMemcpyDeviceToHost));
        if (got != total) return gs_fail(GS_ERR_STATE, "hit list holds %llu entries, expected %llu", got, (unsigned long long)total);
        std::vector<u32> hits((size_t)total);
        if (total) CU(cudaMemcpy(hits.data(), dHits, (size_t)total * sizeof(u32), cudaMemcpyDeviceToHost));
        for (u32 hcv : hits) {  // the n largest counters per row: independent of the visiting order
            update_max_counts((int16_t)(hcv & 0xFFFF), top_counts + (size_t)(hcv >> 16) * nTop, nTop);
            update_max_counts((int16_t)(hcv & 0xFFFF), top_counts + (size_t)V * nTop, nTop);
        }
    }
    return GS_OK;
}

static int finish_checks(gs_sess* s) {
    if (!s) return gs_fail(GS_ERR_STATE, "null session");
    for (DevSess& D : s->devs)
        for (MatchSlot& sl : D.slots)
            if (sl.pending) return gs_fail(GS_ERR_STATE, "ticket %llu still pending", (unsigned long long)sl.ticket);
    return gs_match_sync(s);
}

extern "C" int gs_match_finish(gs_sess* s, gs_taxon_counts* counts, int16_t* top_counts) {
    int rc = finish_checks(s);
    if (rc) return rc;
    if (s->devs.size() > 1 && !s->merged) {   // one process, several GPUs: merge over the context's own communicator
        gs_comm* cm = ctx_all_comm(s->db->ctx);
        if (!cm) return GS_ERR_STATE;
        rc = merge_state(s, cm);
        if (rc) return rc;
    } else if (s->cfg.count_unique_kmers && !s->merged) {
        rc = materialize_bitset(s, s->devs[0]);
        if (rc) return rc;
    }
    s->finished = true;
    return read_out(s, counts, top_counts, !s->merged);
}

extern "C" int gs_match_finish_comm(gs_sess* s, gs_comm* cm, gs_taxon_counts* counts, int16_t* top_counts) {
    if (!cm) return gs_match_finish(s, counts, top_counts);
    int rc = finish_checks(s);
    if (rc) return rc;
    if (s->devs.size() != 1) return gs_fail(GS_ERR_ARG, "gs_match_finish_comm: one GPU per process (the session spans %zu devices)", s->devs.size());
    if (!s->merged) { rc = merge_state(s, cm); if (rc) return rc; }
    s->finished = true;
    return read_out(s, counts, top_counts, false);
}

extern "C" int gs_pack_bases(const uint8_t* bases, uint64_t n, uint64_t* codes, uint32_t* valid, int threads) {
    if ((!bases && n) || !codes || !valid) return gs_fail(GS_ERR_ARG, "null argument");
    if (threads == 1) { gsp::pack_range(bases, n, codes, valid); return GS_OK; }
    gsp::Packer pk(threads);
    pk.pack(bases, n, codes, valid);
    return GS_OK;
}
extern "C" const char* gs_pack_isa(void) { return gsp::pack_isa(); }

extern "C" int gs_match_l2_window(const gs_sess* s, uint64_t* window_bytes, uint64_t* persisting_bytes, double* hit_ratio) {
    if (!s || s->devs.empty()) return gs_fail(GS_ERR_ARG, "gs_match_l2_window: no session");
    const DevSess& D = s->devs[0];
    if (window_bytes) *window_bytes = D.l2WindowBytes;
    if (persisting_bytes) *persisting_bytes = D.l2CarveBytes;
    if (hit_ratio) *hit_ratio = D.l2HitRatio;
    return GS_OK;
}
extern "C" double gs_match_pack_fraction(const gs_sess* s) { return s ? (s->cfg.host_pack_threads == 0 ? 0.0 : s->cfg.host_pack_percent >= 0 ? s->cfg.host_pack_percent / 100.0 : std::min(0.95, std::max(0.05, s->packFrac + s->packBias))) : 0.0; }
extern "C" int gs_match_pack_stats(const gs_sess* s, int* threads, double* pack_seconds, uint64_t* bases_packed, uint64_t* h2d_base_bytes) {
    if (!s) return gs_fail(GS_ERR_ARG, "null session");
    if (threads) *threads = s->packer ? s->packer->threads() : 0;
    if (pack_seconds) *pack_seconds = s->packSeconds;
    if (bases_packed) *bases_packed = s->packBytes;
    if (h2d_base_bytes) *h2d_base_bytes = s->h2dBytes;
    return GS_OK;
}

extern "C" int gs_match_merge_stats(const gs_sess* s, double* total_ms, double* bitset_ms, uint64_t* bytes_from_peers, int* path) {
    if (!s) return gs_fail(GS_ERR_ARG, "null session");
    if (total_ms) *total_ms = s->mergeMs;
    if (bitset_ms) *bitset_ms = s->mergeBitsetMs;
    if (bytes_from_peers) *bytes_from_peers = s->mergeBytes;
    if (path) *path = s->mergePath;
    return GS_OK;
}

// ---------------------------------------------------------------------------------------------------------
// filter
// ---------------------------------------------------------------------------------------------------------
struct DevFilter {
    int dev = 0;
    u64* words = nullptr;
    long long* factors = nullptr;
    GsFilterView view;
};
struct gs_filter {
    gs_ctx* ctx = nullptr;
    std::vector<DevFilter> d;
    int kind = 0;
    long long p0 = 0, p1 = 0;
    u64 nWords = 0;
    std::vector<long long> factors;   // host copy (hashed kinds): what gs_filter_save_file writes
};

extern "C" void gs_filter_destroy(gs_filter* f) {
    if (!f) return;
    for (DevFilter& d : f->d) { cudaSetDevice(d.dev); cudaFree(d.words); cudaFree(d.factors); }
    delete f;
}

extern "C" int gs_filter_n_devices(const gs_filter* f) { return f ? (int)f->d.size() : 0; }
extern "C" gs_ctx* gs_filter_context(gs_filter* f) { return f ? f->ctx : nullptr; }

extern "C" gs_filter* gs_filter_create(gs_ctx* ctx, int kind, int64_t p0, int64_t p1, const int64_t* factors,
                                       const int64_t* words, uint64_t n_words) {
    if (!ctx) { gs_fail(GS_ERR_ARG, "null context"); return nullptr; }
    if (!words) { gs_fail(GS_ERR_ARG, "null words"); return nullptr; }
    u64 modulus = 0;
    if (kind == GS_BLOOM_BLOCKED) {
        if (p1 <= 0 || n_words < (u64)p1 + 17) { gs_fail(GS_ERR_ARG, "blocked filter: buckets=%lld needs buckets+17 words, got %llu", (long long)p1, (unsigned long long)n_words); return nullptr; }
        modulus = (u64)p1;
    } else if (kind == GS_BLOOM_XOR || kind == GS_BLOOM_MURMUR) {
        if (p0 <= 0 || p1 <= 0 || !factors || n_words * 64 < (u64)p0) { gs_fail(GS_ERR_ARG, "hashed filter: bits=%lld hashes=%lld words=%llu", (long long)p0, (long long)p1, (unsigned long long)n_words); return nullptr; }
        modulus = (u64)p0;
    } else { gs_fail(GS_ERR_ARG, "unknown filter kind %d", kind); return nullptr; }
    gs_filter* f = new gs_filter();
    f->ctx = ctx;
    f->kind = kind; f->p0 = p0; f->p1 = p1; f->nWords = n_words;
    if (kind != GS_BLOOM_BLOCKED) f->factors.assign(factors, factors + p1);
    f->d.resize(ctx->devs.size());
    for (size_t i = 0; i < ctx->devs.size(); i++) {
        DevFilter& d = f->d[i];
        d.dev = ctx->devs[i];
        if (cudaSetDevice(d.dev) != cudaSuccess || dmalloc(&d.words, n_words) != cudaSuccess ||
            cudaMemcpy(d.words, words, n_words * sizeof(u64), cudaMemcpyHostToDevice) != cudaSuccess) {
            gs_fail(GS_ERR_CUDA, "filter upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            gs_filter_destroy(f); return nullptr;
        }
        if (kind != GS_BLOOM_BLOCKED) {
            if (dmalloc(&d.factors, (size_t)p1) != cudaSuccess ||
                cudaMemcpy(d.factors, factors, (size_t)p1 * sizeof(long long), cudaMemcpyHostToDevice) != cudaSuccess) {
                gs_fail(GS_ERR_CUDA, "filter upload failed: %s", cudaGetErrorString(cudaGetLastError()));
                gs_filter_destroy(f); return nullptr;
            }
        }
        d.view.kind = kind; d.view.p0 = p0; d.view.p1 = p1; d.view.magic = magic_for(modulus);
        d.view.factors = d.factors; d.view.words = d.words;
    }
    cudaSetDevice(ctx->devs[0]);
    return f;
}

// Flat little-endian filter index file "GSF1": what KMerProbFilter.save / load move as a Java object stream
// (C/bloom/KMerProbFilter.java, C/goals/LoadIndexGoal.java:92-104), as plain arrays.
//   header (48 bytes): magic "GSF1\0\0\0\0", i32 version = 1, i32 kind (GS_BLOOM_*), i64 p0, i64 p1, u64 n_factors, u64 n_words
//   i64 factors[n_factors]   (hashed kinds: hashFactors, AbstractKMerBloomFilter.java:104-110; blocked: none)
//   i64 words[n_words]       (the bit array: long[] or the segments of the large array, in order)
struct GsfHeader { char magic[8]; int32_t version; int32_t kind; int64_t p0, p1; uint64_t n_factors, n_words; };
static_assert(sizeof(GsfHeader) == 48, "GSF1 header layout");

extern "C" int gs_filter_save_file(gs_filter* f, const char* path) {
    if (!f || !path) return gs_fail(GS_ERR_ARG, "null argument");
    try {
        std::vector<int64_t> words((size_t)f->nWords);
        CU(cudaSetDevice(f->d[0].dev));
        CU(cudaMemcpy(words.data(), f->d[0].words, (size_t)f->nWords * sizeof(int64_t), cudaMemcpyDeviceToHost));
        FILE* o = fopen(path, "wb");
        if (!o) return gs_fail(GS_ERR_ARG, "cannot create %s", path);
        GsfHeader h;
        memset(&h, 0, sizeof(h));
        memcpy(h.magic, "GSF1\0\0\0\0", 8);
        h.version = 1; h.kind = f->kind; h.p0 = f->p0; h.p1 = f->p1; h.n_factors = f->factors.size(); h.n_words = f->nWords;
        bool ok = fwrite(&h, sizeof(h), 1, o) == 1;
        ok = ok && (f->factors.empty() || fwrite(f->factors.data(), sizeof(long long), f->factors.size(), o) == f->factors.size());
        ok = ok && (words.empty() || fwrite(words.data(), sizeof(int64_t), words.size(), o) == words.size());
        ok = (fclose(o) == 0) && ok;
        if (!ok) return gs_fail(GS_ERR_ARG, "short write to %s", path);
        return GS_OK;
    } catch (const std::exception& e) {
        return gs_fail(GS_ERR_ARG, "%s: %s", path, e.what());
    }
}

extern "C" gs_filter* gs_filter_load_file(gs_ctx* ctx, const char* path) {
    if (!ctx || !path) { gs_fail(GS_ERR_ARG, "null argument"); return nullptr; }
    try {
        FILE* in = fopen(path, "rb");
        if (!in) { gs_fail(GS_ERR_ARG, "cannot open %s", path); return nullptr; }
        struct Closer { FILE* f; ~Closer() { if (f) fclose(f); } } closer{in};
        GsfHeader h;
        if (fread(&h, sizeof(h), 1, in) != 1 || memcmp(h.magic, "GSF1\0\0\0\0", 8) != 0 || h.version != 1) { gs_fail(GS_ERR_ARG, "%s is not a GSF1 filter file", path); return nullptr; }
        if (fseek(in, 0, SEEK_END) != 0) { gs_fail(GS_ERR_ARG, "%s: cannot seek", path); return nullptr; }
        const u64 fileBytes = (u64)ftell(in);
        fseek(in, (long)sizeof(h), SEEK_SET);
        // the header's sizes must add up to the file's size before anything is allocated from them
        const bool hashed = h.kind == GS_BLOOM_XOR || h.kind == GS_BLOOM_MURMUR;
        const bool sane = (hashed || h.kind == GS_BLOOM_BLOCKED) && h.n_words <= fileBytes / 8 && h.n_factors <= 4096 &&
                          (hashed ? (h.p1 > 0 && (u64)h.p1 == h.n_factors) : h.n_factors == 0);
        if (!sane || sizeof(h) + h.n_factors * 8 + h.n_words * 8 != fileBytes) { gs_fail(GS_ERR_ARG, "%s: header does not match the file (%llu bytes)", path, (unsigned long long)fileBytes); return nullptr; }
        std::vector<int64_t> factors((size_t)h.n_factors), words((size_t)h.n_words);
        if ((h.n_factors && fread(factors.data(), 8, factors.size(), in) != factors.size()) || (h.n_words && fread(words.data(), 8, words.size(), in) != words.size())) {
            gs_fail(GS_ERR_ARG, "%s: truncated", path); return nullptr;
        }
        return gs_filter_create(ctx, h.kind, h.p0, h.p1, hashed ? factors.data() : nullptr, words.data(), h.n_words);   // validates bits / buckets against n_words
    } catch (const std::exception& e) {
        gs_fail(GS_ERR_ARG, "%s: %s", path, e.what());
        return nullptr;
    }
}

extern "C" int gs_filter_contains(gs_filter* f, const int64_t* kmers, uint64_t n, uint8_t* out) {
    if (!f) return gs_fail(GS_ERR_STATE, "null filter");
    CU(cudaSetDevice(f->d[0].dev));
    u64* dK = nullptr; uint8_t* dO = nullptr;
    CU(dmalloc(&dK, n)); CU(dmalloc(&dO, n));
    CU(cudaMemcpy(dK, kmers, n * sizeof(u64), cudaMemcpyHostToDevice));
    if (n) gs_launch_filter_contains(f->d[0].view, dK, n, dO, 0);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, dO, n, cudaMemcpyDeviceToHost));
    CU(cudaFree(dK)); CU(cudaFree(dO));
    return GS_OK;
}

struct FilterSlot {
    bool pending = false;
    gs_ticket ticket = 0;
    u32 nReads = 0;
    uint8_t* dBases = nullptr; size_t basesCap = 0;
    u64* dOffsets = nullptr; size_t offCap = 0;
    uint8_t* dAccept = nullptr; size_t accCap = 0;
    uint8_t* hAccept = nullptr; size_t hAccCap = 0;
    u32* dErr = nullptr; u32* hErr = nullptr;
    // raw FASTQ text batches (gs_match_submit_fastq)
    uint8_t* dText = nullptr; size_t textCap = 0;
    u32* dLineEnd = nullptr; size_t lineCap = 0;
    u32* dBlockCounts = nullptr; size_t blockCountsCap = 0;
    u32* dTextMeta = nullptr;              // [0] lines, [1] records, [2] status bits, [3] pad, then u64 totals[2] (k-mers, bases)
    u32* hTextMeta = nullptr;              // pinned copy
    gs_fastq_rec* dRecs = nullptr; size_t recsCap = 0;
    gs_fastq_rec* hRecs = nullptr; size_t hRecsCap = 0;
    u32* dLens = nullptr; size_t lensCap = 0;
    u64* dTileSums = nullptr; size_t tileSumsCap = 0;
    u32* dEvHdr = nullptr; u32* hEvHdr = nullptr;
    bool isText = false;
    cudaEvent_t evH2D = nullptr, evCompute = nullptr, evDone = nullptr;
};
struct DevFsess {
    int dev = 0, devIndex = 0, blocks = 0;
    cudaStream_t sCopyIn = nullptr, sCompute = nullptr, sCopyOut = nullptr;
    FilterSlot slots[GS_MAX_INFLIGHT];
    // scratch of the flat filter kernels (kernels of one device run back to back on sCompute: one set is enough)
    u32* startBits = nullptr; u32* hitBits = nullptr; size_t bitsCap = 0;
    u32* segCounter = nullptr;
    u64 launches = 0;
};
struct gs_fsess {
    gs_filter* f = nullptr;
    int k = 31, minPosCount = 1;
    double posRatio = 0.2;
    std::vector<DevFsess> devs;
    u64 nextTicket = 1;
};

extern "C" void gs_filter_close(gs_fsess* s) {
    if (!s) return;
    for (DevFsess& D : s->devs) {
        cudaSetDevice(D.dev);
        cudaDeviceSynchronize();
        for (FilterSlot& sl : D.slots) {
            cudaFree(sl.dBases); cudaFree(sl.dOffsets); cudaFree(sl.dAccept); cudaFree(sl.dErr);
            text_stage_free(sl);
            if (sl.hAccept) cudaFreeHost(sl.hAccept);
            if (sl.hErr) cudaFreeHost(sl.hErr);
            if (sl.evH2D) cudaEventDestroy(sl.evH2D);
            if (sl.evCompute) cudaEventDestroy(sl.evCompute);
            if (sl.evDone) cudaEventDestroy(sl.evDone);
        }
        cudaFree(D.startBits); cudaFree(D.hitBits); cudaFree(D.segCounter);
        if (D.sCopyIn) cudaStreamDestroy(D.sCopyIn);
        if (D.sCompute) cudaStreamDestroy(D.sCompute);
        if (D.sCopyOut) cudaStreamDestroy(D.sCopyOut);
    }
    delete s;
}

extern "C" gs_fsess* gs_filter_open(gs_filter* f, int k, int min_pos_count, double pos_ratio) {
    if (!f) { gs_fail(GS_ERR_ARG, "null filter"); return nullptr; }
    if (k < 1 || k > 31) { gs_fail(GS_ERR_ARG, "k=%d out of range [1,31]", k); return nullptr; }
    gs_fsess* s = new gs_fsess();
    s->f = f; s->k = k; s->minPosCount = min_pos_count; s->posRatio = pos_ratio;
    s->devs.resize(f->d.size());
    for (size_t i = 0; i < f->d.size(); i++) {
        DevFsess& D = s->devs[i];
        D.dev = f->d[i].dev; D.devIndex = (int)i;
        bool ok = cudaSetDevice(D.dev) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&D.sCopyIn, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&D.sCompute, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&D.sCopyOut, cudaStreamNonBlocking) == cudaSuccess;
        for (FilterSlot& sl : D.slots)
            ok = ok && cudaEventCreateWithFlags(&sl.evH2D, cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&sl.evCompute, cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&sl.evDone, cudaEventDisableTiming) == cudaSuccess &&
                 dmalloc(&sl.dErr, 1) == cudaSuccess && cudaMallocHost((void**)&sl.hErr, sizeof(u32)) == cudaSuccess;
        if (!ok) { gs_fail(GS_ERR_CUDA, "filter session setup failed: %s", cudaGetErrorString(cudaGetLastError())); gs_filter_close(s); return nullptr; }
        D.blocks = f->ctx->sms[i] * std::max(1, gs_match_kernel_occupancy(2));
        ok = dmalloc(&D.segCounter, 2) == cudaSuccess;
        if (!ok) { gs_fail(GS_ERR_CUDA, "filter session setup failed: %s", cudaGetErrorString(cudaGetLastError())); gs_filter_close(s); return nullptr; }
    }
    cudaSetDevice(f->d[0].dev);
    return s;
}

static void fill_fparams(gs_fsess* s, DevFsess& D, GsFilterParams& P) {
    memset(&P, 0, sizeof(P));
    P.f = s->f->d[D.devIndex].view;
    P.k = s->k; P.minPosCount = s->minPosCount; P.posRatio = s->posRatio;
}

// The three kernels of a filter batch on the device's compute stream: read-start bitmap, hit bit per k-mer position, accept
// byte per read.  `bases + off0` = first base of the batch, nBytes = its length (reads back to back).
static int filter_launch(DevFsess& D, GsFilterParams& P, u64 off0, u64 nBytes) {
    if (!P.nReads) return GS_OK;
    P.off0 = off0;
    P.lead = (u32)((uintptr_t)(P.bases + off0) & 15);
    P.flatLen = nBytes + P.lead;
    const u64 nSeg = (P.flatLen + GS_SEG_POS - 1) / GS_SEG_POS;
    const size_t words = (size_t)nSeg * GS_SEG_CHUNKS + 64;
    if (words > D.bitsCap) {
        CU(cudaStreamSynchronize(D.sCompute));
        cudaFree(D.startBits); cudaFree(D.hitBits); D.startBits = D.hitBits = nullptr; D.bitsCap = 0;
        CU(dmalloc(&D.startBits, words + words / 4));
        CU(dmalloc(&D.hitBits, words + words / 4));
        D.bitsCap = words + words / 4;
    }
    P.startBits = D.startBits; P.hitBits = D.hitBits; P.segCounter = D.segCounter;
    CU(cudaMemsetAsync(D.startBits, 0, words * sizeof(u32), D.sCompute));
    CU(cudaMemsetAsync(D.segCounter, 0, 2 * sizeof(u32), D.sCompute));
    const int blocks = (int)std::max<u64>(1, std::min<u64>((u64)D.blocks, (nSeg + GS_WARPS_PER_BLOCK - 1) / GS_WARPS_PER_BLOCK));
    gs_launch_filter(P, blocks, D.sCompute);
    CU(cudaGetLastError());
    D.launches += 3;
    return GS_OK;
}

extern "C" int gs_filter_submit(gs_fsess* s, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, gs_ticket* ticket) {
    if (!s) return gs_fail(GS_ERR_STATE, "null session");
    if (!ticket || !offsets) return gs_fail(GS_ERR_ARG, "null argument");
    const gs_ticket t = s->nextTicket;
    const size_t nDev = s->devs.size();
    DevFsess& D = s->devs[(t - 1) % nDev];
    FilterSlot& sl = D.slots[((t - 1) / nDev) % GS_MAX_INFLIGHT];
    if (sl.pending) return gs_fail(GS_ERR_STATE, "more than %d batches in flight on a device: collect ticket %llu first", GS_MAX_INFLIGHT, (unsigned long long)sl.ticket);
    if (offsets[n_reads] < offsets[0]) return gs_fail(GS_ERR_ARG, "offsets not ascending");
    const u64 base0 = offsets[0];
    const u64 nBytes = offsets[n_reads] - base0;
    CU(cudaSetDevice(D.dev));
    CU(dgrow(&sl.dBases, &sl.basesCap, (size_t)nBytes + 64));
    CU(dgrow(&sl.dOffsets, &sl.offCap, (size_t)n_reads + 1));
    CU(dgrow(&sl.dAccept, &sl.accCap, (size_t)n_reads));
    CU(hgrow(&sl.hAccept, &sl.hAccCap, (size_t)n_reads));
    if (nBytes) CU(cudaMemcpyAsync(sl.dBases, bases + base0, nBytes, cudaMemcpyHostToDevice, D.sCopyIn));
    CU(cudaMemcpyAsync(sl.dOffsets, offsets, ((size_t)n_reads + 1) * sizeof(u64), cudaMemcpyHostToDevice, D.sCopyIn));
    CU(cudaEventRecord(sl.evH2D, D.sCopyIn));
    CU(cudaStreamWaitEvent(D.sCompute, sl.evH2D, 0));
    GsFilterParams P;
    fill_fparams(s, D, P);
    P.bases = sl.dBases - base0; P.offsets = sl.dOffsets; P.nReads = n_reads; P.accept = sl.dAccept; P.errFlag = sl.dErr;
    CU(cudaMemsetAsync(sl.dErr, 0, sizeof(u32), D.sCompute));
    { const int rcl = filter_launch(D, P, base0, nBytes); if (rcl) return rcl; }
    CU(cudaEventRecord(sl.evCompute, D.sCompute));
    CU(cudaStreamWaitEvent(D.sCopyOut, sl.evCompute, 0));
    if (n_reads) CU(cudaMemcpyAsync(sl.hAccept, sl.dAccept, n_reads, cudaMemcpyDeviceToHost, D.sCopyOut));
    CU(cudaMemcpyAsync(sl.hErr, sl.dErr, sizeof(u32), cudaMemcpyDeviceToHost, D.sCopyOut));
    CU(cudaEventRecord(sl.evDone, D.sCopyOut));
    sl.pending = true; sl.ticket = t; sl.nReads = n_reads; sl.isText = false;
    s->nextTicket++;
    *ticket = t;
    return GS_OK;
}

extern "C" int gs_filter_submit_fastq(gs_fsess* s, const uint8_t* text, uint64_t n_bytes, gs_fastq_info* info, gs_ticket* ticket) {
    if (!s) return gs_fail(GS_ERR_STATE, "null session");
    if (!ticket || !info || (!text && n_bytes)) return gs_fail(GS_ERR_ARG, "null argument");
    if (n_bytes >= 0xFFFFFF00ULL) return gs_fail(GS_ERR_LIMIT, "text chunk of %llu bytes (limit 2^32 - 256)", (unsigned long long)n_bytes);
    memset(info, 0, sizeof(*info));
    *ticket = 0;
    const gs_ticket t = s->nextTicket;
    const size_t nDev = s->devs.size();
    DevFsess& D = s->devs[(t - 1) % nDev];
    FilterSlot& sl = D.slots[((t - 1) / nDev) % GS_MAX_INFLIGHT];
    if (sl.pending) return gs_fail(GS_ERR_STATE, "more than %d batches in flight on a device: collect ticket %llu first", GS_MAX_INFLIGHT, (unsigned long long)sl.ticket);
    CU(cudaSetDevice(D.dev));
    u64 launches = 0;
    int rcs = text_stage(sl, D.sCopyIn, text, n_bytes, s->k, info, &launches);
    if (rcs) return rcs;
    if (info->status) return GS_OK;
    const u32 n_reads = info->n_reads;
    CU(dgrow(&sl.dAccept, &sl.accCap, (size_t)n_reads));
    CU(hgrow(&sl.hAccept, &sl.hAccCap, (size_t)n_reads));
    CU(cudaEventRecord(sl.evH2D, D.sCopyIn));
    CU(cudaStreamWaitEvent(D.sCompute, sl.evH2D, 0));
    CU(cudaStreamWaitEvent(D.sCopyOut, sl.evH2D, 0));
    CU(cudaMemcpyAsync(sl.hRecs, sl.dRecs, ((size_t)n_reads + 1) * sizeof(gs_fastq_rec), cudaMemcpyDeviceToHost, D.sCopyOut));
    GsFilterParams P;
    fill_fparams(s, D, P);
    P.bases = sl.dBases; P.offsets = sl.dOffsets; P.nReads = n_reads; P.accept = sl.dAccept; P.errFlag = sl.dErr;
    CU(cudaMemsetAsync(sl.dErr, 0, sizeof(u32), D.sCompute));
    { const int rcl = filter_launch(D, P, 0, info->total_bps); if (rcl) return rcl; }
    CU(cudaEventRecord(sl.evCompute, D.sCompute));
    CU(cudaStreamWaitEvent(D.sCopyOut, sl.evCompute, 0));
    if (n_reads) CU(cudaMemcpyAsync(sl.hAccept, sl.dAccept, n_reads, cudaMemcpyDeviceToHost, D.sCopyOut));
    CU(cudaMemcpyAsync(sl.hErr, sl.dErr, sizeof(u32), cudaMemcpyDeviceToHost, D.sCopyOut));
    CU(cudaEventRecord(sl.evDone, D.sCopyOut));
    sl.pending = true; sl.ticket = t; sl.nReads = n_reads; sl.isText = true;
    s->nextTicket++;
    *ticket = t;
    return GS_OK;
}

extern "C" int gs_filter_collect_fastq(gs_fsess* s, gs_ticket t, const uint8_t** accept, uint32_t* n_reads, const gs_fastq_rec** recs) {
    if (!s) return gs_fail(GS_ERR_STATE, "null session");
    if (t == 0 || t >= s->nextTicket) return gs_fail(GS_ERR_STATE, "unknown ticket %llu", (unsigned long long)t);
    const size_t nDev = s->devs.size();
    DevFsess& D = s->devs[(t - 1) % nDev];
    FilterSlot& sl = D.slots[((t - 1) / nDev) % GS_MAX_INFLIGHT];
    if (!sl.pending || sl.ticket != t || !sl.isText) return gs_fail(GS_ERR_STATE, "ticket %llu is not a pending FASTQ text batch", (unsigned long long)t);
    CU(cudaSetDevice(D.dev));
    CU(cudaEventSynchronize(sl.evDone));
    sl.pending = false;
    if (sl.hErr[0]) return gs_fail(GS_ERR_ARG, "batch of ticket %llu holds malformed read offsets", (unsigned long long)t);
    if (accept) *accept = sl.hAccept;
    if (n_reads) *n_reads = sl.nReads;
    if (recs) *recs = sl.hRecs;
    return GS_OK;
}

extern "C" int gs_filter_collect(gs_fsess* s, gs_ticket t, uint8_t* accept) {
    if (!s) return gs_fail(GS_ERR_STATE, "null session");
    if (t == 0 || t >= s->nextTicket) return gs_fail(GS_ERR_STATE, "unknown ticket %llu", (unsigned long long)t);
    const size_t nDev = s->devs.size();
    DevFsess& D = s->devs[(t - 1) % nDev];
    FilterSlot& sl = D.slots[((t - 1) / nDev) % GS_MAX_INFLIGHT];
    if (!sl.pending || sl.ticket != t) return gs_fail(GS_ERR_STATE, "ticket %llu is not pending", (unsigned long long)t);
    CU(cudaSetDevice(D.dev));
    CU(cudaEventSynchronize(sl.evDone));
    sl.pending = false;
    if (*sl.hErr) return gs_fail(GS_ERR_ARG, "batch of ticket %llu holds malformed read offsets", (unsigned long long)t);
    if (accept && sl.nReads) memcpy(accept, sl.hAccept, sl.nReads);
    return GS_OK;
}

extern "C" int gs_filter_run_device(gs_fsess* s, const uint8_t* d_bases, const uint64_t* d_offsets, uint32_t n_reads, uint64_t n_bases, uint8_t* d_accept) {
    if (!s) return gs_fail(GS_ERR_STATE, "null session");
    if (((uintptr_t)d_bases & 15) != 0) return gs_fail(GS_ERR_ARG, "d_bases must be 16-byte aligned");
    DevFsess& D = s->devs[0];
    CU(cudaSetDevice(D.dev));
    GsFilterParams P;
    fill_fparams(s, D, P);
    P.bases = d_bases; P.offsets = (const u64*)d_offsets; P.nReads = n_reads; P.accept = d_accept; P.errFlag = nullptr;
    return filter_launch(D, P, 0, n_bases);
}

extern "C" uint64_t gs_filter_kernel_launches(const gs_fsess* s) {
    u64 n = 0;
    if (s) for (const DevFsess& D : s->devs) n += D.launches;
    return n;
}


extern "C" int gs_filter_sync(gs_fsess* s) {
    if (!s) return gs_fail(GS_ERR_STATE, "null session");
    for (DevFsess& D : s->devs) {
        CU(cudaSetDevice(D.dev));
        CU(cudaStreamSynchronize(D.sCopyIn));
        CU(cudaStreamSynchronize(D.sCompute));
        CU(cudaStreamSynchronize(D.sCopyOut));
    }
    return GS_OK;
}

extern "C" void* gs_filter_stream(gs_fsess* s) { return s ? (void*)s->devs[0].sCompute : nullptr; }
