// Host side of the C ABI (include/fugu_gpu.h): context, index snapshot upload (flat CSR ->
// block-compressed HBM layout), plan lowering (fg_query_batch -> device plan + work items) and
// batch execution. No CPU evaluation path exists here: without a CUDA device every compute entry
// point returns FG_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <time.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fugu_gpu.h"
#include "fg_error.h"
#include "fg_internal.h"
#include "fg_pool.h"

using namespace fg;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int32_t fail(int32_t code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int32_t fg::host_fail(int32_t code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(e_ == cudaErrorMemoryAllocation ? FG_ERR_OOM : FG_ERR_CUDA,           \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" const char* fg_last_error(void) { return g_err; }
extern "C" const char* fg_version(void) { return "fugu_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------
// tantivy fieldnorm code (Lucene SmallFloat byte4ToInt) and BM25 weight (A.4)
// ------------------------------------------------------------------------------------------
static uint32_t g_fn_table[256];
static std::once_flag g_fn_once;
static void fn_init() {
    for (uint32_t b = 0; b < 256; b++) {
        if (b < 24) { g_fn_table[b] = b; continue; }
        uint32_t i = b - 24, bits = i & 7, shift = i >> 3;
        uint32_t dec = shift == 0 ? bits : ((bits | 8u) << (shift - 1));
        g_fn_table[b] = 24 + dec;
    }
}
extern "C" uint32_t fg_id_to_fieldnorm(uint8_t id) {
    std::call_once(g_fn_once, fn_init);
    return g_fn_table[id];
}
extern "C" uint8_t fg_fieldnorm_to_id(uint32_t n) {
    std::call_once(g_fn_once, fn_init);
    // largest id with table[id] <= n
    const uint32_t* p = std::upper_bound(g_fn_table, g_fn_table + 256, n);
    return (uint8_t)((p - g_fn_table) - 1);
}
extern "C" float fg_bm25_idf(uint64_t df, uint64_t n) {
    float x = ((float)(n - df) + 0.5f) / ((float)df + 0.5f);
    return logf(1.0f + x);
}
static const float K1 = 1.2f, B = 0.75f;

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
struct LeadScratch;
struct fg_ctx {
    int device = 0;
    std::vector<std::unique_ptr<LeadScratch>> lead_scratch;  // reusable host vectors of the plan lowering (under pool_mu)
    cudaStream_t own = nullptr;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    int n_sms = 0;
    // side streams + events: the per-class search kernels of one batch run concurrently
    cudaStream_t aux[NCLS - 1] = {};
    cudaEvent_t fork_ev = nullptr, join_ev[NCLS - 1] = {};
    // small cache of device blocks for the per-call buffers (plans, partial lists, results):
    // cudaMalloc/cudaFree per request would dominate the host side of a 5000-query batch
    std::mutex pool_mu;
    std::vector<std::pair<void*, size_t>> pool;
    size_t pool_bytes = 0;
    // plan uploads run on their own stream: a pageable-memory H2D copy first waits for everything queued
    // on its stream, which on the compute stream would serialise the upload of batch i+1 behind the
    // kernels of batch i (fg_batch_submit / fgh_search_batch pipeline host work under device work)
    cudaStream_t up = nullptr;
    std::vector<std::pair<void*, size_t>> hpool;  // page-locked staging blocks for asynchronous result copies
    // development switches, read ONCE when the context is created (never per request)
    bool env_legacy = false;      // FG_LEGACY=1: lower every plan for the round-1 window kernels (fg_kernels.cu)
    bool env_no_columns = false;  // FG_NO_COLUMNS=1
    bool env_no_prune = false;    // FG_NO_PRUNE=1: exhaustive evaluation (A/B runs; results are identical)
    bool env_timing = false;      // FG_TIMING=1
    bool env_prof = false;        // FG_PROF=1
    // fg_batch_submit alternates between these streams: the tail of one pipeline chunk's persistent kernel (warps running out of
    // work) overlaps the start of the next chunk's instead of serialising behind it
    cudaStream_t sub[2] = {nullptr, nullptr};
    uint32_t sub_seq = 0, sub_streams = 2;
    uint32_t lead_par_blocks = 16, lead_max_par = 64, lead_chunk = 16, lead_tma = 0, lead_chunk_req = 48, lead_max_par_req = 1024, lead_union_work = 32;
};
static uint64_t env_u64_early(const char* name, uint64_t dflt);
static double now_ms();

static cudaError_t pinned_alloc(fg_ctx* c, void** out, size_t bytes) {
    bytes = std::max<size_t>((bytes + 4095) & ~(size_t)4095, 4096);
    {
        std::lock_guard<std::mutex> g(c->pool_mu);
        int best = -1;
        for (int i = 0; i < (int)c->hpool.size(); i++)
            if (c->hpool[i].second >= bytes && c->hpool[i].second <= 4 * bytes + (1 << 16) &&
                (best < 0 || c->hpool[i].second < c->hpool[best].second)) best = i;
        if (best >= 0) {
            *out = c->hpool[best].first;
            c->hpool.erase(c->hpool.begin() + best);
            return cudaSuccess;
        }
    }
    return cudaHostAlloc(out, bytes, cudaHostAllocDefault);
}
static void pinned_free(fg_ctx* c, void* p, size_t bytes) {
    if (!p) return;
    bytes = std::max<size_t>((bytes + 4095) & ~(size_t)4095, 4096);
    std::lock_guard<std::mutex> g(c->pool_mu);
    if (c->hpool.size() < 16) c->hpool.emplace_back(p, bytes);
    else cudaFreeHost(p);
}

static cudaError_t pool_alloc(fg_ctx* c, void** out, size_t bytes) {
    bytes = std::max<size_t>((bytes + 255) & ~(size_t)255, 256);
    {
        std::lock_guard<std::mutex> g(c->pool_mu);
        int best = -1;
        for (int i = 0; i < (int)c->pool.size(); i++)
            if (c->pool[i].second >= bytes && c->pool[i].second <= 4 * bytes + (1 << 16) &&
                (best < 0 || c->pool[i].second < c->pool[best].second)) best = i;
        if (best >= 0) {
            *out = c->pool[best].first;
            c->pool_bytes -= c->pool[best].second;
            c->pool.erase(c->pool.begin() + best);
            return cudaSuccess;
        }
    }
    return cudaMalloc(out, bytes);
}
// NOTE: callers must have synchronised the stream that last used `p`
static void pool_free(fg_ctx* c, void* p, size_t bytes) {
    if (!p) return;
    bytes = std::max<size_t>((bytes + 255) & ~(size_t)255, 256);
    std::lock_guard<std::mutex> g(c->pool_mu);
    if (c->pool.size() < 64 && c->pool_bytes + bytes <= ((size_t)1 << 30)) {
        c->pool.emplace_back(p, bytes);
        c->pool_bytes += bytes;
    } else {
        cudaFree(p);
    }
}

extern "C" void fg_ctx_destroy(fg_ctx* c);
extern "C" int32_t fg_ctx_create(int32_t device, fg_ctx** out) {
    if (!out) return fail(FG_ERR_INVALID, "fg_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(FG_ERR_NO_DEVICE, "no CUDA device (%s); the query path has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(FG_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    CU(cudaSetDevice(device));
    cudaDeviceProp pr;
    CU(cudaGetDeviceProperties(&pr, device));
    if (pr.major < 10)
        return fail(FG_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                    device, pr.major, pr.minor);
    std::unique_ptr<fg_ctx, void (*)(fg_ctx*)> guard(new fg_ctx(), fg_ctx_destroy);  // a failing call below must not leak streams / events
    fg_ctx* c = guard.get();
    c->device = device;
    c->n_sms = pr.multiProcessorCount;
    c->env_legacy = env_u64_early("FG_LEGACY", 0) != 0;
    c->env_no_columns = getenv("FG_NO_COLUMNS") != nullptr;
    c->env_no_prune = getenv("FG_NO_PRUNE") != nullptr;
    c->env_timing = getenv("FG_TIMING") != nullptr;
    c->env_prof = getenv("FG_PROF") != nullptr;
    c->lead_par_blocks = (uint32_t)std::max<uint64_t>(1, env_u64_early("FG_LEAD_PAR_BLOCKS", 16));
    c->lead_max_par = (uint32_t)std::max<uint64_t>(1, env_u64_early("FG_LEAD_MAX_PAR", 64));
    c->lead_chunk = (uint32_t)std::max<uint64_t>(1, env_u64_early("FG_LEAD_CHUNK", 16));
    c->lead_chunk_req = (uint32_t)std::max<uint64_t>(1, env_u64_early("FG_LEAD_REQ_WORK", 48));
    c->lead_max_par_req = (uint32_t)std::max<uint64_t>(1, env_u64_early("FG_LEAD_MAX_PAR_REQ", 1024));
    c->lead_union_work = (uint32_t)std::max<uint64_t>(1, env_u64_early("FG_LEAD_UNION_WORK", 32));
    c->lead_tma = (uint32_t)env_u64_early("FG_LEAD_TMA", 0);
    CU(cudaStreamCreateWithFlags(&c->own, cudaStreamNonBlocking));
    c->stream = c->own;
    for (int i = 0; i < NCLS - 1; i++) {
        CU(cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->join_ev[i], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
    CU(cudaStreamCreateWithFlags(&c->up, cudaStreamNonBlocking));
    c->sub_streams = (uint32_t)std::min<uint64_t>(2, env_u64_early("FG_SUBMIT_STREAMS", 2));
    for (uint32_t i = 0; i < c->sub_streams; i++) CU(cudaStreamCreateWithFlags(&c->sub[i], cudaStreamNonBlocking));
    *out = guard.release();
    return FG_OK;
}
extern "C" void fg_ctx_destroy(fg_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->own) cudaStreamDestroy(c->own);
    for (int i = 0; i < NCLS - 1; i++) { if (c->aux[i]) cudaStreamDestroy(c->aux[i]); if (c->join_ev[i]) cudaEventDestroy(c->join_ev[i]); }
    if (c->fork_ev) cudaEventDestroy(c->fork_ev);
    for (auto& b : c->pool) cudaFree(b.first);
    for (auto& b : c->hpool) cudaFreeHost(b.first);
    if (c->up) cudaStreamDestroy(c->up);
    for (auto& st : c->sub) if (st) cudaStreamDestroy(st);
    delete c;
}
extern "C" int32_t fg_ctx_set_stream(fg_ctx* c, void* s) {
    if (!c) return fail(FG_ERR_INVALID, "ctx is NULL");
    std::lock_guard<std::mutex> g(c->mu);
    c->stream = s ? (cudaStream_t)s : c->own;
    return FG_OK;
}
extern "C" int32_t fg_ctx_synchronize(fg_ctx* c) {
    if (!c) return fail(FG_ERR_INVALID, "ctx is NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return FG_OK;
}

// ------------------------------------------------------------------------------------------
// index snapshot
// ------------------------------------------------------------------------------------------
struct TermInfo {
    uint32_t blk_begin, n_blocks, df_local, df_global;
    uint64_t bytes;  // packed payload + 16 B skip per block
    int32_t col;     // dense tf column of the term (index into fg_index::d_cols) or -1
    float idf_w;     // idf(df_global, N) * (1 + K1): the leaf weight before the boost (one logf per term at upload,
                     // not one per leaf per query in the lowering)
    float max_factor;  // max over the term's postings of tf / (tf + norm): weight * max_factor bounds the leaf's score
    uint32_t top_off, top_n;  // the term's top_n largest block maxima, descending, at fg_index::topbm[top_off ..]
    int32_t bm;        // membership bitmap of the term (slot in fg_index::d_bits) or -1
};
struct HostField {
    uint32_t flags = 0;
    uint32_t n_terms = 0;
    uint64_t total_tokens = 0;
    float cnorm = 0.f;
    std::vector<TermInfo> terms;
};
struct DeviceArena {
    int device = 0;
    std::vector<void*> allocs;
    ~DeviceArena() {
        cudaSetDevice(device);
        for (void* p : allocs) cudaFree(p);
    }
};
struct fg_index {
    fg_ctx* ctx = nullptr;
    uint32_t n_docs = 0, doc_base = 0;
    uint64_t global_n_docs = 0;
    std::vector<HostField> fields;
    fg_index_info info{};
    DevIndex dev{};
    // device arrays of the snapshot. Shared by the snapshots fg_index_with_alive derives from this
    // one (same postings, another alive bitset); freed when the last of them is released.
    std::shared_ptr<DeviceArena> arena;
    void* own_alive = nullptr;  // alive bitset of a derived snapshot (not in the arena)
    // dense tf columns of the most frequent terms (see build_columns)
    const uint8_t* d_cols = nullptr;
    uint64_t col_stride = 0;
    uint32_t n_cols = 0;
    // per term with at least TOP_MIN_BLOCKS blocks: its largest block maxima (descending, at most TOP_MAX). The
    // k-th of them is a lower bound of the k-th best score of any pure union containing the term.
    std::shared_ptr<std::vector<float>> topbm;
    // membership bitmaps + rank directories of the mid-frequency terms (see build at upload)
    const uint32_t* d_bits = nullptr;
    const uint32_t* d_rank = nullptr;
    uint64_t bm_stride_words = 0;
    uint32_t n_bitmaps = 0;
};
static constexpr uint32_t TOP_MIN_BLOCKS = 4, TOP_MAX = 128;

static uint64_t env_u64_early(const char* name, uint64_t dflt) {
    const char* e = getenv(name);
    return e ? strtoull(e, nullptr, 10) : dflt;
}
static int host_threads() {
    unsigned h = std::thread::hardware_concurrency();
    const char* e = getenv("FG_HOST_THREADS");
    if (e) h = (unsigned)atoi(e);
    return (int)std::max(1u, std::min(h ? h : 4u, 64u));
}
static inline uint32_t bits_of(uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; }

template <class F>
static void parallel_for(uint64_t n, int T, F f) {
    if (n == 0) return;
    T = (int)std::min<uint64_t>((uint64_t)T, n);
    if (T <= 1) { f(0, n, 0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back([=]() { f(n * t / T, n * (t + 1) / T, t); });
    for (auto& x : th) x.join();
}

static void pack_stream(uint32_t* words, const uint32_t* vals, uint32_t n, uint32_t b) {
    // little-endian horizontal stream, value i at bit i*b; caller zeroed `words` (4*b words)
    if (b == 0) return;
    for (uint32_t i = 0; i < n; i++) {
        uint64_t bp = (uint64_t)i * b;
        uint32_t wi = (uint32_t)(bp >> 5), sh = (uint32_t)(bp & 31);
        uint64_t v = (uint64_t)vals[i] << sh;
        words[wi] |= (uint32_t)v;
        if (sh + b > 32) words[wi + 1] |= (uint32_t)(v >> 32);
    }
}

static uint32_t count_alive(const uint32_t* bits, uint32_t n_docs) {
    uint32_t na = 0;
    for (uint32_t w = 0; w < (n_docs + 31) / 32; w++) {
        uint32_t x = bits[w];
        if (w == n_docs / 32 && (n_docs & 31)) x &= (1u << (n_docs & 31)) - 1u;
        na += (uint32_t)__builtin_popcount(x);
    }
    return na;
}

extern "C" void fg_index_release(fg_index* ix) {
    if (!ix) return;
    if (ix->ctx) cudaSetDevice(ix->ctx->device);
    if (ix->own_alive) cudaFree(ix->own_alive);
    delete ix;  // drops this snapshot's share of the arena
}

// One f32 per block: the largest tf / (tf + norm(doc)) of its postings, computed on the device with the
// scoring path's own arithmetic (so weight * bmax >= every score of the block, exactly). The host keeps
// each term's maximum and its largest block maxima for the MaxScore bounds of the lowering.
// Needs ix->dev.{skip, blk, fnorm, cache} and the term table; allocates ix->dev.bmax in the snapshot's arena.
static int32_t compute_block_max(fg_index* ix, uint64_t n_blocks, int T) {
    fg_ctx* ctx = ix->ctx;
    const uint32_t n_fields = (uint32_t)ix->fields.size();
    float* d_bmax = nullptr;
    CU(cudaMalloc((void**)&d_bmax, std::max<size_t>(n_blocks * 4, 16)));
    ix->arena->allocs.push_back(d_bmax);
    ix->info.device_bytes += n_blocks * 4;
    ix->dev.bmax = d_bmax;
    for (uint32_t f = 0; f < n_fields; f++) {
        const HostField& hf = ix->fields[f];
        if (hf.terms.empty()) continue;
        const uint32_t b0 = hf.terms.front().blk_begin, b1 = hf.terms.back().blk_begin + hf.terms.back().n_blocks;
        launch_blockmax(ix->dev, b0, b1, (hf.flags & FG_FIELD_HAS_FIELDNORMS) ? (int)f : -1, hf.cnorm, d_bmax, ctx->stream);
    }
    CU(cudaGetLastError());
    std::vector<float> bm(n_blocks);
    CU(cudaMemcpyAsync(bm.data(), d_bmax, n_blocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    auto top = std::make_shared<std::vector<float>>();
    // offsets first (serial), then the per-term selection in parallel
    uint64_t top_total = 0;
    for (uint32_t f = 0; f < n_fields; f++)
        for (auto& ti : ix->fields[f].terms) {
            ti.top_off = ti.top_n = 0;
            if (ti.n_blocks >= TOP_MIN_BLOCKS) {
                ti.top_off = (uint32_t)top_total;
                ti.top_n = std::min<uint32_t>(ti.n_blocks, TOP_MAX);
                top_total += ti.top_n;
            }
        }
    if (top_total > 0xFFFFFFF0ull) return fail(FG_ERR_UNSUPPORTED, "block-max table too large");
    top->resize(top_total);
    for (uint32_t f = 0; f < n_fields; f++) {
        HostField& hf = ix->fields[f];
        parallel_for(hf.n_terms, T, [&](uint64_t a, uint64_t b, int) {
            std::vector<float> tmp;
            for (uint64_t t = a; t < b; t++) {
                TermInfo& ti = hf.terms[t];
                const float* v = bm.data() + ti.blk_begin;
                float mx = 0.f;
                for (uint32_t i = 0; i < ti.n_blocks; i++) mx = std::max(mx, v[i]);
                ti.max_factor = mx;
                if (!ti.top_n) continue;
                tmp.assign(v, v + ti.n_blocks);
                std::partial_sort(tmp.begin(), tmp.begin() + ti.top_n, tmp.end(), std::greater<float>());
                std::copy(tmp.begin(), tmp.begin() + ti.top_n, top->begin() + ti.top_off);
            }
        });
    }
    ix->topbm = top;
    return FG_OK;
}

extern "C" int32_t fg_index_upload(fg_ctx* ctx, const fg_index_desc* d, fg_index** out) {
    if (!ctx || !d || !out) return fail(FG_ERR_INVALID, "fg_index_upload: NULL argument");
    *out = nullptr;
    if (d->n_fields == 0 || d->n_fields > (uint32_t)MAX_FIELDS || !d->fields)
        return fail(FG_ERR_INVALID, "n_fields must be in [1,%d]", MAX_FIELDS);
    std::call_once(g_fn_once, fn_init);
    CU(cudaSetDevice(ctx->device));
    std::unique_ptr<fg_index, void (*)(fg_index*)> ix(new fg_index(), fg_index_release);
    ix->ctx = ctx;
    ix->arena = std::make_shared<DeviceArena>();
    ix->arena->device = ctx->device;
    ix->n_docs = d->n_docs;
    ix->doc_base = d->doc_id_base;
    ix->global_n_docs = d->global_n_docs ? d->global_n_docs : d->n_docs;
    ix->fields.resize(d->n_fields);
    const int T = host_threads();

    // ---- block counts per term ----
    uint64_t n_blocks = 0, n_postings = 0;
    for (uint32_t f = 0; f < d->n_fields; f++) {
        const fg_field_desc& fd = d->fields[f];
        HostField& hf = ix->fields[f];
        hf.flags = fd.flags;
        hf.n_terms = fd.n_terms;
        hf.total_tokens = fd.total_num_tokens;
        hf.terms.resize(fd.n_terms);
        if (fd.n_terms && (!fd.term_offsets || !fd.doc_ids))
            return fail(FG_ERR_INVALID, "field %u: term_offsets/doc_ids NULL", f);
        if ((fd.flags & FG_FIELD_HAS_FIELDNORMS) && !fd.fieldnorm_ids && d->n_docs)
            return fail(FG_ERR_INVALID, "field %u: HAS_FIELDNORMS but fieldnorm_ids NULL", f);
        for (uint32_t t = 0; t < fd.n_terms; t++) {
            if (fd.term_offsets[t + 1] < fd.term_offsets[t])
                return fail(FG_ERR_INVALID, "field %u: term_offsets not monotone at %u", f, t);
            uint64_t n = fd.term_offsets[t + 1] - fd.term_offsets[t];
            if (n > d->n_docs) return fail(FG_ERR_INVALID, "field %u term %u: df > n_docs", f, t);
            TermInfo& ti = hf.terms[t];
            ti.df_local = (uint32_t)n;
            ti.df_global = fd.global_doc_freq ? fd.global_doc_freq[t] : (uint32_t)n;
            ti.idf_w = fg_bm25_idf(ti.df_global, ix->global_n_docs) * (1.0f + K1);
            if (ti.df_global < ti.df_local)
                return fail(FG_ERR_INVALID, "field %u term %u: global df < local df", f, t);
            ti.n_blocks = (uint32_t)((n + BLOCK - 1) / BLOCK);
            if (n_blocks + ti.n_blocks > 0xFFFFFFF0ull) return fail(FG_ERR_UNSUPPORTED, "too many blocks");
            ti.blk_begin = (uint32_t)n_blocks;
            ti.bytes = 0;
            ti.col = -1;
            ti.max_factor = 0.f;
            ti.top_off = ti.top_n = 0;
            ti.bm = -1;
            n_blocks += ti.n_blocks;
            n_postings += n;
        }
    }

    // ---- pass A: per-block bit widths (parallel over terms of each field) ----
    std::vector<SkipEntry> skip(n_blocks);
    std::vector<uint8_t> wsum(n_blocks);  // bd + bt (16-byte units of payload)
    std::vector<int> bad(T, 0);
    for (uint32_t f = 0; f < d->n_fields; f++) {
        const fg_field_desc& fd = d->fields[f];
        HostField& hf = ix->fields[f];
        const bool freqs = (fd.flags & FG_FIELD_HAS_FREQS) && fd.term_freqs;
        parallel_for(fd.n_terms, T, [&](uint64_t a, uint64_t b, int tix) {
            for (uint64_t t = a; t < b; t++) {
                const TermInfo& ti = hf.terms[t];
                const uint32_t* docs = fd.doc_ids + fd.term_offsets[t];
                const uint32_t* tfs = freqs ? fd.term_freqs + fd.term_offsets[t] : nullptr;
                uint32_t prev_plus1 = 0;  // previous doc + 1
                for (uint32_t bi = 0; bi < ti.n_blocks; bi++) {
                    const uint32_t i0 = bi * BLOCK, n = std::min<uint32_t>(BLOCK, ti.df_local - i0);
                    uint32_t gor = 0, tor = 0, p = prev_plus1;
                    for (uint32_t i = 0; i < n; i++) {
                        const uint32_t dd = docs[i0 + i];
                        if (dd < p || dd >= d->n_docs) { bad[tix] = 1; break; }
                        gor |= dd - p;
                        p = dd + 1;
                        if (tfs) {
                            if (tfs[i0 + i] == 0) { bad[tix] = 1; break; }
                            tor |= tfs[i0 + i] - 1;
                        }
                    }
                    const uint32_t bd = bits_of(gor), bt = bits_of(tor);
                    SkipEntry& e = skip[ti.blk_begin + bi];
                    e.last_doc = docs[i0 + n - 1];
                    e.first_base = prev_plus1;
                    e.off16 = 0;
                    e.packed = pack_meta(bd, bt, n);
                    wsum[ti.blk_begin + bi] = (uint8_t)(bd + bt);
                    prev_plus1 = p;
                }
            }
        });
    }
    for (int t = 0; t < T; t++)
        if (bad[t])
            return fail(FG_ERR_INVALID, "postings must be strictly ascending, < n_docs, with tf >= 1");

    // ---- offsets ----
    uint64_t off16 = 0;
    for (uint64_t b = 0; b < n_blocks; b++) {
        if (off16 > 0xFFFFFFFFull) return fail(FG_ERR_UNSUPPORTED, "packed payload exceeds 64 GiB");
        skip[b].off16 = (uint32_t)off16;
        off16 += wsum[b];
    }
    const uint64_t payload = off16 * 16;
    for (uint32_t f = 0; f < d->n_fields; f++)
        for (auto& ti : ix->fields[f].terms) {
            uint64_t by = 0;
            for (uint32_t bi = 0; bi < ti.n_blocks; bi++) by += (uint64_t)wsum[ti.blk_begin + bi] * 16 + 16;
            ti.bytes = by;
        }

    // ---- pass B: pack (parallel) ----
    std::vector<uint32_t> blk((payload + 64) / 4, 0);
    for (uint32_t f = 0; f < d->n_fields; f++) {
        const fg_field_desc& fd = d->fields[f];
        HostField& hf = ix->fields[f];
        const bool freqs = (fd.flags & FG_FIELD_HAS_FREQS) && fd.term_freqs;
        parallel_for(fd.n_terms, T, [&](uint64_t a, uint64_t b, int) {
            uint32_t gaps[BLOCK], tfm[BLOCK];
            for (uint64_t t = a; t < b; t++) {
                const TermInfo& ti = hf.terms[t];
                const uint32_t* docs = fd.doc_ids + fd.term_offsets[t];
                const uint32_t* tfs = freqs ? fd.term_freqs + fd.term_offsets[t] : nullptr;
                for (uint32_t bi = 0; bi < ti.n_blocks; bi++) {
                    const SkipEntry& e = skip[ti.blk_begin + bi];
                    const uint32_t bd = e.packed & 63, bt = (e.packed >> 6) & 63;
                    const uint32_t i0 = bi * BLOCK, n = ((e.packed >> 12) & 127) + 1;
                    uint32_t p = e.first_base;
                    for (uint32_t i = 0; i < n; i++) {
                        gaps[i] = docs[i0 + i] - p;
                        p = docs[i0 + i] + 1;
                        tfm[i] = tfs ? tfs[i0 + i] - 1 : 0;
                    }
                    uint32_t* w = blk.data() + (size_t)e.off16 * 4;
                    pack_stream(w, gaps, n, bd);
                    pack_stream(w + 4 * bd, tfm, n, bt);
                }
            }
        });
    }

    // ---- BM25 norm caches from GLOBAL statistics ----
    std::vector<float> cache((size_t)MAX_FIELDS * 256, 0.f);
    for (uint32_t f = 0; f < d->n_fields; f++) {
        const float avg = (float)ix->fields[f].total_tokens / (float)ix->global_n_docs;
        for (int i = 0; i < 256; i++)
            cache[f * 256 + i] = K1 * (1.0f - B + B * (float)g_fn_table[i] / avg);
        ix->fields[f].cnorm = cache[f * 256 + fg_fieldnorm_to_id(1)];
    }

    // ---- upload ----
    auto dev_copy = [&](const void* src, size_t bytes, const void** dst) -> int32_t {
        void* p = nullptr;
        CU(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
        ix->arena->allocs.push_back(p);
        if (bytes) CU(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
        *dst = p;
        ix->info.device_bytes += bytes;
        return FG_OK;
    };
    int32_t rc;
    if ((rc = dev_copy(skip.data(), n_blocks * sizeof(SkipEntry), (const void**)&ix->dev.skip))) return rc;
    if ((rc = dev_copy(blk.data(), blk.size() * 4, (const void**)&ix->dev.blk))) return rc;
    if ((rc = dev_copy(cache.data(), cache.size() * 4, (const void**)&ix->dev.cache))) return rc;
    for (uint32_t f = 0; f < d->n_fields; f++) {
        ix->dev.fnorm[f] = nullptr;
        if (d->fields[f].flags & FG_FIELD_HAS_FIELDNORMS) {
            // padded: the slot scan reads the fieldnorm ids of 4 docs with one 32-bit load
            std::vector<uint8_t> fnp(((size_t)d->n_docs + 15) / 16 * 16 + 16, 0);
            if (d->n_docs) memcpy(fnp.data(), d->fields[f].fieldnorm_ids, d->n_docs);
            if ((rc = dev_copy(fnp.data(), fnp.size(), (const void**)&ix->dev.fnorm[f]))) return rc;
        }
    }
    ix->dev.alive = nullptr;
    if (d->alive_bitset)
        if ((rc = dev_copy(d->alive_bitset, ((size_t)d->n_docs + 31) / 32 * 4, (const void**)&ix->dev.alive))) return rc;
    ix->dev.n_docs = d->n_docs;
    ix->dev.doc_base = d->doc_id_base;
    ix->dev.n_alive = d->alive_bitset ? count_alive(d->alive_bitset, d->n_docs) : d->n_docs;

    // ---- block-max metadata (tantivy stores the block-max fieldnorm/tf pair in its skip entries, A.7) ----
    if ((rc = compute_block_max(ix.get(), n_blocks, T))) return rc;

    // ---- dense tf columns -----------------------------------------------------------------
    // A term that occurs in at least 1/FG_COL_DIV of the shard's docs additionally gets a dense
    // column: one byte per doc holding its term frequency (0 = absent). Leaves on such terms skip
    // block decode entirely: the slot scan reads 4 docs per 32-bit load and computes
    // w * tf / (tf + cache[fieldnorm]) from the column byte and the fieldnorm byte, and an
    // intersection looks a candidate up with one byte load instead of a skip search + block decode.
    // 1 B/doc/term of HBM buys ~20x fewer instructions per posting on the lists that carry most of
    // the postings of a Zipfian query mix. The posting blocks stay (exact accounting, tf > 255).
    {
        const uint64_t col_div = env_u64_early("FG_COL_DIV", 16);
        const uint64_t col_min_df = env_u64_early("FG_COL_MIN_DF", 128);
        const uint64_t budget = env_u64_early("FG_COL_MAX_MB", 16384) << 20;
        const uint64_t stride = (((uint64_t)d->n_docs + 15) & ~(uint64_t)15) + 16;
        struct Cand { uint32_t f, t, df; };
        std::vector<Cand> cand;
        if (col_div && d->n_docs)
            for (uint32_t f = 0; f < d->n_fields; f++) {
                const fg_field_desc& fd = d->fields[f];
                const bool freqs = (fd.flags & FG_FIELD_HAS_FREQS) && fd.term_freqs;
                for (uint32_t t = 0; t < fd.n_terms; t++) {
                    const TermInfo& ti = ix->fields[f].terms[t];
                    if ((uint64_t)ti.df_local * col_div < d->n_docs || ti.df_local < col_min_df) continue;
                    bool ok = true;  // a byte must hold every tf of the list
                    if (freqs) {
                        const uint32_t* tfs = fd.term_freqs + fd.term_offsets[t];
                        for (uint32_t i = 0; i < ti.df_local && ok; i++) ok = tfs[i] <= 255;
                    }
                    if (ok) cand.push_back({f, t, ti.df_local});
                }
            }
        std::stable_sort(cand.begin(), cand.end(), [](const Cand& a, const Cand& b) { return a.df > b.df; });
        while (!cand.empty() && cand.size() * stride > budget) cand.pop_back();
        if (!cand.empty()) {
            std::vector<uint8_t> cols(cand.size() * stride, 0);
            parallel_for(cand.size(), T, [&](uint64_t a, uint64_t b, int) {
                for (uint64_t c = a; c < b; c++) {
                    const fg_field_desc& fd = d->fields[cand[c].f];
                    const bool freqs = (fd.flags & FG_FIELD_HAS_FREQS) && fd.term_freqs;
                    const uint64_t o = fd.term_offsets[cand[c].t];
                    uint8_t* col = cols.data() + c * stride;
                    for (uint32_t i = 0; i < cand[c].df; i++)
                        col[fd.doc_ids[o + i]] = freqs ? (uint8_t)fd.term_freqs[o + i] : (uint8_t)1;
                }
            });
            if ((rc = dev_copy(cols.data(), cols.size(), (const void**)&ix->d_cols))) return rc;
            for (size_t c = 0; c < cand.size(); c++) ix->fields[cand[c].f].terms[cand[c].t].col = (int32_t)c;
            ix->col_stride = stride;
            ix->n_cols = (uint32_t)cand.size();
        }
        ix->info.n_columns = ix->n_cols;
        ix->info.column_bytes = ix->n_cols * stride;
    }

    // ---- membership bitmaps of the mid-frequency terms ---------------------------------------
    // A term below the column threshold whose list still spans many blocks is the expensive lookup target
    // (every candidate lands in a different block: skip search + block decode per candidate). Such a term gets
    // a bitmap (1 bit per doc) + a rank directory (postings before every 256-doc chunk): a lookup is one 4-byte
    // gather, and only a hit goes on to its posting (rank -> block, position -> tf stream). Built on the device
    // from the posting blocks. 180 GB of HBM3e pays for it: n_docs/8 * 1.125 bytes per term.
    {
        const uint64_t bm_min_df = env_u64_early("FG_BITMAP_MIN_DF", 512);
        const uint64_t budget = env_u64_early("FG_BITMAP_MAX_MB", 16384) << 20;
        const uint64_t stride_words = (((uint64_t)d->n_docs + 255) / 256) * 8 + 8;
        const uint64_t per_term = stride_words * 4 + stride_words / 2;
        struct Cand { uint32_t f, t, df; };
        std::vector<Cand> cand;
        if (bm_min_df && d->n_docs)
            for (uint32_t f = 0; f < d->n_fields; f++)
                for (uint32_t t = 0; t < ix->fields[f].n_terms; t++) {
                    const TermInfo& ti = ix->fields[f].terms[t];
                    if (ti.col < 0 && ti.df_local >= bm_min_df) cand.push_back({f, t, ti.df_local});
                }
        std::stable_sort(cand.begin(), cand.end(), [](const Cand& a, const Cand& b) { return a.df > b.df; });
        while (!cand.empty() && cand.size() * per_term > budget) cand.pop_back();
        if (!cand.empty()) {
            std::vector<uint2> sel;
            for (size_t c = 0; c < cand.size(); c++) {
                TermInfo& ti = ix->fields[cand[c].f].terms[cand[c].t];
                ti.bm = (int32_t)c;
                for (uint32_t b = 0; b < ti.n_blocks; b++) sel.push_back(make_uint2(ti.blk_begin + b, (uint32_t)c));
            }
            uint32_t *d_bits = nullptr, *d_rank = nullptr;
            uint2* d_sel = nullptr;
            const size_t bits_bytes = cand.size() * stride_words * 4, rank_bytes = cand.size() * (stride_words / 8) * 4;
            CU(cudaMalloc((void**)&d_bits, bits_bytes));
            ix->arena->allocs.push_back(d_bits);
            CU(cudaMalloc((void**)&d_rank, rank_bytes));
            ix->arena->allocs.push_back(d_rank);
            CU(cudaMalloc((void**)&d_sel, sel.size() * sizeof(uint2)));
            CU(cudaMemsetAsync(d_bits, 0, bits_bytes, ctx->stream));
            CU(cudaMemcpyAsync(d_sel, sel.data(), sel.size() * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
            launch_bitmap_build(ix->dev, d_sel, (uint32_t)sel.size(), d_bits, stride_words, ctx->stream);
            launch_bitmap_rank(d_bits, d_rank, (uint32_t)cand.size(), stride_words, ctx->stream);
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(ctx->stream));
            CU(cudaFree(d_sel));
            ix->d_bits = d_bits;
            ix->d_rank = d_rank;
            ix->bm_stride_words = stride_words;
            ix->n_bitmaps = (uint32_t)cand.size();
            ix->info.device_bytes += bits_bytes + rank_bytes;
        }
        ix->info.n_bitmaps = ix->n_bitmaps;
        ix->info.bitmap_bytes = (uint64_t)ix->n_bitmaps * per_term;
    }

    ix->info.n_postings = n_postings;
    ix->info.n_blocks = n_blocks;
    ix->info.packed_bytes = payload;
    ix->info.skip_bytes = n_blocks * 16;
    ix->info.n_docs = d->n_docs;
    ix->info.n_fields = d->n_fields;
    *out = ix.release();
    return FG_OK;
}

extern "C" int32_t fg_index_get_info(const fg_index* ix, fg_index_info* out) {
    if (!ix || !out) return fail(FG_ERR_INVALID, "NULL argument");
    *out = ix->info;
    return FG_OK;
}
extern "C" int32_t fg_index_term_info(const fg_index* ix, uint32_t field, uint32_t term,
                                      uint32_t* ldf, uint32_t* gdf, uint32_t* nb, uint64_t* bytes) {
    if (!ix) return fail(FG_ERR_INVALID, "NULL index");
    if (field >= ix->fields.size() || term >= ix->fields[field].n_terms)
        return fail(FG_ERR_INVALID, "field/term out of range");
    const TermInfo& t = ix->fields[field].terms[term];
    if (ldf) *ldf = t.df_local;
    if (gdf) *gdf = t.df_global;
    if (nb) *nb = t.n_blocks;
    if (bytes) *bytes = t.bytes;
    return FG_OK;
}

// A snapshot that differs from `base` only in which docs are alive (deletes since the last commit,
// src/db/document.rs:38-41,65): shares every device array of `base`, uploads n_docs/8 bytes.
extern "C" int32_t fg_index_with_alive(fg_index* base, const uint32_t* alive_bitset, fg_index** out) {
    if (!base || !out) return fail(FG_ERR_INVALID, "fg_index_with_alive: NULL argument");
    *out = nullptr;
    CU(cudaSetDevice(base->ctx->device));
    std::unique_ptr<fg_index, void (*)(fg_index*)> ix(new fg_index(*base), fg_index_release);
    ix->own_alive = nullptr;  // the copy must not adopt base's private bitset
    ix->dev.alive = nullptr;
    ix->dev.n_alive = base->n_docs;
    if (alive_bitset) {
        // same padding rule as the other per-doc arrays: kernels read the bits of 8 docs as one byte
        const size_t bytes = ((size_t)base->n_docs + 31) / 32 * 4, padded = (bytes + 16 + 15) & ~(size_t)15;
        CU(cudaMalloc(&ix->own_alive, padded));
        CU(cudaMemset(ix->own_alive, 0, padded));
        if (bytes) CU(cudaMemcpy(ix->own_alive, alive_bitset, bytes, cudaMemcpyHostToDevice));
        ix->dev.alive = (const uint32_t*)ix->own_alive;
        ix->dev.n_alive = count_alive(alive_bitset, base->n_docs);
        ix->info.device_bytes += bytes;
    }
    *out = ix.release();
    return FG_OK;
}

// ------------------------------------------------------------------------------------------
// snapshot append: base snapshot + one new segment -> new snapshot
// ------------------------------------------------------------------------------------------
// The reference commits on every ingest call (src/db/document.rs:65,97); tantivy then adds one new segment and
// leaves the old ones alone. The equivalent here: the new documents take the doc ids after the base's (a new
// tantivy segment's doc-id base is the sum of the earlier max_docs too), so every posting list only grows at its
// tail. Only the segment's own postings cross PCIe; everything else is rebuilt from what is already in HBM:
//   * full blocks of the base keep their payload bytes and offsets (one device-to-device copy of the payload
//     region); a term's partial last block is decoded on the device, merged with the term's new postings on the
//     host and re-encoded together with them behind the old payload;
//   * skip entries are laid out anew (terms stay contiguous) by a range-copy kernel + the new blocks' entries;
//   * N and the average field length change, so norm caches, idf weights and the block-max metadata of EVERY
//     block are recomputed (blockmax_kernel over the whole payload: HBM speed);
//   * fieldnorm ids, dense tf columns and membership bitmaps are extended in place of being rebuilt: old part
//     copied on the device, the new documents' part uploaded / built from the new blocks.
// Which terms own a column or a bitmap is decided at full uploads only (a term crossing a threshold later is
// served from its blocks until the next one; a column term whose new tf exceeds a byte loses its column).
namespace {
struct DevTmp {  // scratch device buffer of a build step: freed on every path out of the scope
    void* p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)); }
    template <class T> T* as() const { return static_cast<T*>(p); }
    ~DevTmp() { if (p) cudaFree(p); }
};
struct NewBlocks {  // new blocks of one touched term, encoded on the host
    std::vector<SkipEntry> skips;
    std::vector<uint32_t> words;  // payload, 4 * (bd + bt) words per block
};
void encode_postings(const uint32_t* docs, const uint32_t* tfs, uint32_t n, uint32_t first_base, NewBlocks& out) {
    uint32_t prev_plus1 = first_base;
    uint32_t gaps[BLOCK], tfm[BLOCK];
    for (uint32_t i0 = 0; i0 < n; i0 += BLOCK) {
        const uint32_t m = std::min<uint32_t>(BLOCK, n - i0);
        uint32_t gor = 0, tor = 0, p = prev_plus1;
        for (uint32_t i = 0; i < m; i++) {
            gaps[i] = docs[i0 + i] - p;
            p = docs[i0 + i] + 1;
            tfm[i] = tfs ? tfs[i0 + i] - 1 : 0;
            gor |= gaps[i];
            tor |= tfm[i];
        }
        const uint32_t bd = bits_of(gor), bt = bits_of(tor);
        SkipEntry e;
        e.last_doc = docs[i0 + m - 1];
        e.first_base = prev_plus1;
        e.off16 = (uint32_t)(out.words.size() / 4);  // relative to the term's first new block; rebased by the caller
        e.packed = pack_meta(bd, bt, m);
        out.skips.push_back(e);
        const size_t w0 = out.words.size();
        out.words.resize(w0 + 4 * (size_t)(bd + bt), 0);
        pack_stream(out.words.data() + w0, gaps, m, bd);
        pack_stream(out.words.data() + w0 + 4 * bd, tfm, m, bt);
        prev_plus1 = p;
    }
}
}  // namespace

extern "C" int32_t fg_index_append(fg_index* base, const fg_index_desc* seg, const uint32_t* alive_bitset, fg_index** out) {
    if (!base || !seg || !out) return fail(FG_ERR_INVALID, "fg_index_append: NULL argument");
    *out = nullptr;
    if (seg->n_fields != base->fields.size() || !seg->fields) return fail(FG_ERR_INVALID, "fg_index_append: the segment must have the base's %zu fields", base->fields.size());
    if (base->global_n_docs != base->n_docs || base->doc_base != 0 || seg->global_n_docs || seg->doc_id_base)
        return fail(FG_ERR_UNSUPPORTED, "fg_index_append: sharded snapshots (global statistics) are rebuilt with fg_index_upload");
    if ((uint64_t)base->n_docs + seg->n_docs > 0xFFFFFFF0ull) return fail(FG_ERR_UNSUPPORTED, "too many docs");
    fg_ctx* ctx = base->ctx;
    CU(cudaSetDevice(ctx->device));
    std::call_once(g_fn_once, fn_init);
    const int T = host_threads();
    const uint32_t n_old = base->n_docs, n_seg = seg->n_docs, n_new = n_old + n_seg, n_fields = seg->n_fields;
    for (uint32_t f = 0; f < n_fields; f++) {
        const fg_field_desc& fd = seg->fields[f];
        const HostField& bf = base->fields[f];
        if (fd.flags != bf.flags) return fail(FG_ERR_INVALID, "field %u: flags differ from the base snapshot", f);
        if (fd.n_terms < bf.n_terms) return fail(FG_ERR_INVALID, "field %u: the segment's dictionary must extend the base's (%u < %u terms)", f, fd.n_terms, bf.n_terms);
        if (fd.n_terms && (!fd.term_offsets || (fd.term_offsets[fd.n_terms] && !fd.doc_ids))) return fail(FG_ERR_INVALID, "field %u: term_offsets/doc_ids NULL", f);
        if ((fd.flags & FG_FIELD_HAS_FIELDNORMS) && !fd.fieldnorm_ids && n_seg) return fail(FG_ERR_INVALID, "field %u: HAS_FIELDNORMS but fieldnorm_ids NULL", f);
    }
    double t_mark = now_ms();
    std::string t_log;
#define APPEND_MARK(label)                                                        \
    do {                                                                          \
        if (ctx->env_timing) {                                                    \
            char b_[96];                                                          \
            const double t_ = now_ms();                                           \
            snprintf(b_, sizeof(b_), " %s %.1f", label, t_ - t_mark);             \
            t_log += b_;                                                          \
            t_mark = t_;                                                          \
        }                                                                         \
    } while (0)
    std::unique_ptr<fg_index, void (*)(fg_index*)> ix(new fg_index(), fg_index_release);
    ix->ctx = ctx;
    ix->arena = std::make_shared<DeviceArena>();
    ix->arena->device = ctx->device;
    ix->n_docs = n_new;
    ix->doc_base = 0;
    ix->global_n_docs = n_new;
    ix->fields.resize(n_fields);
    auto dev_alloc = [&](size_t bytes, void** dst) -> int32_t {
        void* p = nullptr;
        CU(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
        ix->arena->allocs.push_back(p);
        ix->info.device_bytes += bytes;
        *dst = p;
        return FG_OK;
    };
    int32_t rc;

    // ---- term table of the new snapshot; which blocks of the base are kept, which are decoded for the host ----
    struct Touched { uint32_t f, t, old_last_block /* global, EMPTY32 = the term is new */, tail, list_idx; };
    constexpr uint32_t EMPTY32 = 0xFFFFFFFFu;
    std::vector<Touched> touched;
    std::vector<uint32_t> decode_list;  // global block indices of the base handed to tail_decode_kernel: postings wanted ...
    std::vector<uint32_t> meta_list;    // ... and those of which only the skip entry's (last_doc, first_base) is needed
    uint64_t n_blocks = 0, n_postings = 0;
    for (uint32_t f = 0; f < n_fields; f++) {
        const fg_field_desc& fd = seg->fields[f];
        const HostField& bf = base->fields[f];
        HostField& hf = ix->fields[f];
        hf.flags = fd.flags;
        hf.n_terms = fd.n_terms;
        hf.total_tokens = bf.total_tokens + fd.total_num_tokens;
        hf.terms.resize(fd.n_terms);
        for (uint32_t t = 0; t < fd.n_terms; t++) {
            if (fd.term_offsets[t + 1] < fd.term_offsets[t]) return fail(FG_ERR_INVALID, "field %u: term_offsets not monotone at %u", f, t);
            const uint64_t add = fd.term_offsets[t + 1] - fd.term_offsets[t];
            if (add > n_seg) return fail(FG_ERR_INVALID, "field %u term %u: df > n_docs of the segment", f, t);
            TermInfo ti{};
            const TermInfo* old = t < bf.n_terms ? &bf.terms[t] : nullptr;
            const uint32_t old_df = old ? old->df_local : 0;
            ti.df_local = ti.df_global = old_df + (uint32_t)add;
            // A bitmap term finds a hit's posting through its rank (block = rank / 128): every block of it but the last
            // stays full, so its partial last block is re-encoded together with the new postings. Any other term keeps
            // its blocks as they are and gets new blocks behind them (a partial block in the middle of a list is fine
            // for the block walk, the gallop and the block maxima; the next full upload packs the list again).
            const uint32_t tail = (old && add && old->bm >= 0) ? old_df % BLOCK : 0u;
            ti.n_blocks = (old ? old->n_blocks : 0u) - (tail ? 1u : 0u) + (uint32_t)((tail + add + BLOCK - 1) / BLOCK);
            ti.blk_begin = (uint32_t)n_blocks;
            ti.col = old ? old->col : -1;
            ti.bm = old ? old->bm : -1;
            ti.idf_w = fg_bm25_idf(ti.df_global, n_new) * (1.0f + K1);
            if (n_blocks + ti.n_blocks > 0xFFFFFFF0ull) return fail(FG_ERR_UNSUPPORTED, "too many blocks");
            n_blocks += ti.n_blocks;
            n_postings += ti.df_local;
            if (add) {
                Touched x{f, t, EMPTY32, tail, EMPTY32};
                if (old && old->n_blocks) {
                    x.old_last_block = old->blk_begin + old->n_blocks - 1;
                    std::vector<uint32_t>& list = tail ? decode_list : meta_list;  // postings wanted, or just (last_doc, first_base)
                    x.list_idx = (uint32_t)list.size();
                    list.push_back(x.old_last_block);
                }
                touched.push_back(x);
            }
            hf.terms[t] = ti;
        }
    }

    APPEND_MARK("term table");
    // ---- last blocks of the touched terms: (last_doc, first_base) and, for partial ones, their postings ----
    std::vector<uint32_t> h_meta(2 * decode_list.size()), h_docs((size_t)BLOCK * decode_list.size()), h_tfs((size_t)BLOCK * decode_list.size());
    if (!decode_list.empty()) {
        DevTmp t_list, t_meta, t_docs, t_tfs;
        const size_t nl = decode_list.size();
        CU(t_list.alloc(nl * 4));
        CU(t_meta.alloc(nl * 8));
        CU(t_docs.alloc(nl * BLOCK * 4));
        CU(t_tfs.alloc(nl * BLOCK * 4));
        CU(cudaMemcpyAsync(t_list.p, decode_list.data(), nl * 4, cudaMemcpyHostToDevice, ctx->stream));
        launch_tail_decode(base->dev, t_list.as<uint32_t>(), (uint32_t)nl, t_meta.as<uint32_t>(), t_docs.as<uint32_t>(), t_tfs.as<uint32_t>(), ctx->stream);
        CU(cudaMemcpyAsync(h_meta.data(), t_meta.p, nl * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(h_docs.data(), t_docs.p, nl * BLOCK * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(h_tfs.data(), t_tfs.p, nl * BLOCK * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaGetLastError());
    }
    std::vector<uint32_t> h_meta2(2 * meta_list.size());
    if (!meta_list.empty()) {
        DevTmp t_list, t_meta;
        const size_t nl = meta_list.size();
        CU(t_list.alloc(nl * 4));
        CU(t_meta.alloc(nl * 8));
        CU(cudaMemcpyAsync(t_list.p, meta_list.data(), nl * 4, cudaMemcpyHostToDevice, ctx->stream));
        launch_tail_decode(base->dev, t_list.as<uint32_t>(), (uint32_t)nl, t_meta.as<uint32_t>(), nullptr, nullptr, ctx->stream);
        CU(cudaMemcpyAsync(h_meta2.data(), t_meta.p, nl * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaGetLastError());
    }

    APPEND_MARK("tail decode");
    // ---- encode the new blocks of every touched term (parallel): old partial tail + the segment's postings ----
    std::vector<NewBlocks> enc(touched.size());
    std::vector<int> bad(T, 0);
    parallel_for(touched.size(), T, [&](uint64_t a, uint64_t b, int tix) {
        std::vector<uint32_t> docs, tfs;
        for (uint64_t i = a; i < b; i++) {
            const Touched& x = touched[i];
            const fg_field_desc& fd = seg->fields[x.f];
            const bool freqs = (fd.flags & FG_FIELD_HAS_FREQS) && fd.term_freqs;
            const uint64_t o = fd.term_offsets[x.t];
            const uint32_t add = (uint32_t)(fd.term_offsets[x.t + 1] - o);
            docs.clear(); tfs.clear();
            uint32_t first_base = 0;
            if (x.list_idx != EMPTY32) {
                if (!x.tail) {
                    first_base = h_meta2[2 * x.list_idx] + 1;  // behind the base's last posting
                } else {                                       // the partial last block is re-encoded with the new postings
                    first_base = h_meta[2 * x.list_idx + 1];
                    docs.assign(h_docs.begin() + (size_t)x.list_idx * BLOCK, h_docs.begin() + (size_t)x.list_idx * BLOCK + x.tail);
                    tfs.assign(h_tfs.begin() + (size_t)x.list_idx * BLOCK, h_tfs.begin() + (size_t)x.list_idx * BLOCK + x.tail);
                }
            }
            uint32_t prev = EMPTY32;
            for (uint32_t j = 0; j < add; j++) {
                const uint32_t dl = fd.doc_ids[o + j], tf = freqs ? fd.term_freqs[o + j] : 1u;
                if (dl >= n_seg || (prev != EMPTY32 && dl <= prev) || tf == 0) { bad[tix] = 1; break; }
                prev = dl;
                docs.push_back(n_old + dl);
                tfs.push_back(tf);
            }
            if (bad[tix]) break;
            encode_postings(docs.data(), freqs ? tfs.data() : nullptr, (uint32_t)docs.size(), first_base, enc[i]);
        }
    });
    for (int t = 0; t < T; t++)
        if (bad[t]) return fail(FG_ERR_INVALID, "segment postings must be strictly ascending, < n_docs of the segment, with tf >= 1");

    APPEND_MARK("encode");
    // ---- layout: payload of the new blocks behind the base's, skip entries term by term ----
    const uint64_t old_payload16 = base->info.packed_bytes / 16;
    uint64_t off16 = old_payload16;
    std::vector<uint32_t> new_words;
    std::vector<SkipEntry> new_skips;          // compact: the new blocks of the touched terms, in term order
    {
        size_t nw = 0, nsk = 0;
        for (const NewBlocks& nb : enc) { nw += nb.words.size(); nsk += nb.skips.size(); }
        new_words.reserve(nw);
        new_skips.reserve(nsk);
    }
    std::vector<uint3> old_ranges, new_ranges;  // {src_begin, dst_begin, count} for copy_ranges_kernel
    {
        size_t ti_touched = 0;
        for (uint32_t f = 0; f < n_fields; f++) {
            const HostField& bf = base->fields[f];
            HostField& hf = ix->fields[f];
            for (uint32_t t = 0; t < hf.n_terms; t++) {
                TermInfo& ti = hf.terms[t];
                const TermInfo* old = t < bf.n_terms ? &bf.terms[t] : nullptr;
                const bool is_touched = ti_touched < touched.size() && touched[ti_touched].f == f && touched[ti_touched].t == t;
                uint32_t kept = old ? old->n_blocks : 0;
                if (is_touched && touched[ti_touched].tail) kept--;  // the partial last block is re-encoded
                if (kept) old_ranges.push_back(make_uint3(old->blk_begin, ti.blk_begin, kept));
                uint64_t bytes = old ? old->bytes : 0;
                if (is_touched) {
                    NewBlocks& nb = enc[ti_touched];
                    if (kept + nb.skips.size() != ti.n_blocks) return fail(FG_ERR_CUDA, "fg_index_append: internal block count mismatch (field %u term %u)", f, t);
                    new_ranges.push_back(make_uint3((uint32_t)new_skips.size(), ti.blk_begin + kept, (uint32_t)nb.skips.size()));
                    for (SkipEntry e : nb.skips) {
                        e.off16 += (uint32_t)off16;
                        new_skips.push_back(e);
                    }
                    if (off16 + nb.words.size() / 4 > 0xFFFFFFFFull) return fail(FG_ERR_UNSUPPORTED, "packed payload exceeds 64 GiB");
                    off16 += nb.words.size() / 4;
                    new_words.insert(new_words.end(), nb.words.begin(), nb.words.end());
                    bytes += nb.words.size() * 4 + nb.skips.size() * 16;  // (the replaced partial block's bytes stay counted: dead payload)
                    ti_touched++;
                }
                ti.bytes = bytes;
            }
        }
    }
    const uint64_t payload = off16 * 16;

    APPEND_MARK("layout");
    // ---- device arrays ----
    uint8_t* d_blk = nullptr;
    uint4* d_skip = nullptr;
    if ((rc = dev_alloc(payload + 64, (void**)&d_blk))) return rc;
    if ((rc = dev_alloc(n_blocks * sizeof(SkipEntry), (void**)&d_skip))) return rc;
    CU(cudaMemcpyAsync(d_blk, base->dev.blk, old_payload16 * 16, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaMemsetAsync(d_blk + payload, 0, 64, ctx->stream));
    if (!new_words.empty()) CU(cudaMemcpyAsync(d_blk + old_payload16 * 16, new_words.data(), new_words.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    {
        DevTmp t_r0, t_r1, t_ns;
        CU(t_r0.alloc(old_ranges.size() * sizeof(uint3)));
        CU(t_r1.alloc(new_ranges.size() * sizeof(uint3)));
        CU(t_ns.alloc(new_skips.size() * sizeof(SkipEntry)));
        if (!old_ranges.empty()) CU(cudaMemcpyAsync(t_r0.p, old_ranges.data(), old_ranges.size() * sizeof(uint3), cudaMemcpyHostToDevice, ctx->stream));
        if (!new_ranges.empty()) CU(cudaMemcpyAsync(t_r1.p, new_ranges.data(), new_ranges.size() * sizeof(uint3), cudaMemcpyHostToDevice, ctx->stream));
        if (!new_skips.empty()) CU(cudaMemcpyAsync(t_ns.p, new_skips.data(), new_skips.size() * sizeof(SkipEntry), cudaMemcpyHostToDevice, ctx->stream));
        launch_copy_ranges(base->dev.skip, d_skip, t_r0.p, (uint32_t)old_ranges.size(), ctx->stream);
        launch_copy_ranges(t_ns.p, d_skip, t_r1.p, (uint32_t)new_ranges.size(), ctx->stream);
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaGetLastError());
    }
    ix->dev.skip = d_skip;
    ix->dev.blk = d_blk;

    // norm caches from the new statistics
    std::vector<float> cache((size_t)MAX_FIELDS * 256, 0.f);
    for (uint32_t f = 0; f < n_fields; f++) {
        const float avg = (float)ix->fields[f].total_tokens / (float)ix->global_n_docs;
        for (int i = 0; i < 256; i++) cache[f * 256 + i] = K1 * (1.0f - B + B * (float)g_fn_table[i] / avg);
        ix->fields[f].cnorm = cache[f * 256 + fg_fieldnorm_to_id(1)];
    }
    float* d_cache = nullptr;
    if ((rc = dev_alloc(cache.size() * 4, (void**)&d_cache))) return rc;
    CU(cudaMemcpyAsync(d_cache, cache.data(), cache.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    ix->dev.cache = d_cache;
    // fieldnorm ids: the base's on the device, the segment's from the host
    for (uint32_t f = 0; f < n_fields; f++) {
        ix->dev.fnorm[f] = nullptr;
        if (!(seg->fields[f].flags & FG_FIELD_HAS_FIELDNORMS)) continue;
        const size_t padded = ((size_t)n_new + 15) / 16 * 16 + 16;
        uint8_t* p = nullptr;
        if ((rc = dev_alloc(padded, (void**)&p))) return rc;
        CU(cudaMemsetAsync(p, 0, padded, ctx->stream));
        if (n_old) CU(cudaMemcpyAsync(p, base->dev.fnorm[f], n_old, cudaMemcpyDeviceToDevice, ctx->stream));
        if (n_seg) CU(cudaMemcpyAsync(p + n_old, seg->fields[f].fieldnorm_ids, n_seg, cudaMemcpyHostToDevice, ctx->stream));
        ix->dev.fnorm[f] = p;
    }
    ix->dev.alive = nullptr;
    ix->dev.n_docs = n_new;
    ix->dev.doc_base = 0;
    ix->dev.n_alive = n_new;
    if (alive_bitset) {
        const size_t bytes = ((size_t)n_new + 31) / 32 * 4, padded = (bytes + 16 + 15) & ~(size_t)15;
        uint32_t* p = nullptr;
        if ((rc = dev_alloc(padded, (void**)&p))) return rc;
        CU(cudaMemsetAsync(p, 0, padded, ctx->stream));
        CU(cudaMemcpyAsync(p, alive_bitset, bytes, cudaMemcpyHostToDevice, ctx->stream));
        ix->dev.alive = p;
        ix->dev.n_alive = count_alive(alive_bitset, n_new);
    }
    CU(cudaStreamSynchronize(ctx->stream));  // (host staging vectors above may go out of use)

    APPEND_MARK("payload + skip + norms + alive");
    // block maxima of every block under the new norms, per-term maxima and top tables
    if ((rc = compute_block_max(ix.get(), n_blocks, T))) return rc;

    APPEND_MARK("block maxima");
    // ---- dense tf columns: the base's columns, extended by the segment's docs ----
    if (base->n_cols) {
        const uint64_t stride = (((uint64_t)n_new + 15) & ~(uint64_t)15) + 16;
        uint8_t* d_cols = nullptr;
        if ((rc = dev_alloc((size_t)base->n_cols * stride, (void**)&d_cols))) return rc;
        CU(cudaMemsetAsync(d_cols, 0, (size_t)base->n_cols * stride, ctx->stream));
        if (n_old) CU(cudaMemcpy2DAsync(d_cols, stride, base->d_cols, base->col_stride, n_old, base->n_cols, cudaMemcpyDeviceToDevice, ctx->stream));
        std::vector<uint8_t> part((size_t)base->n_cols * std::max<uint32_t>(n_seg, 1), 0);
        for (uint32_t f = 0; f < n_fields; f++) {
            const fg_field_desc& fd = seg->fields[f];
            const bool freqs = (fd.flags & FG_FIELD_HAS_FREQS) && fd.term_freqs;
            for (uint32_t t = 0; t < base->fields[f].n_terms; t++) {
                TermInfo& ti = ix->fields[f].terms[t];
                if (ti.col < 0) continue;
                const uint64_t o = fd.term_offsets[t], add = fd.term_offsets[t + 1] - o;
                uint8_t* col = part.data() + (size_t)ti.col * n_seg;
                for (uint64_t j = 0; j < add; j++) {
                    const uint32_t tf = freqs ? fd.term_freqs[o + j] : 1u;
                    if (tf > 255) { ti.col = -1; break; }  // a byte no longer holds every tf: the term is served from its blocks
                    col[fd.doc_ids[o + j]] = (uint8_t)tf;
                }
            }
        }
        if (n_seg) CU(cudaMemcpy2DAsync(d_cols + n_old, stride, part.data(), n_seg, n_seg, base->n_cols, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ix->d_cols = d_cols;
        ix->col_stride = stride;
        ix->n_cols = base->n_cols;
    }
    ix->info.n_columns = ix->n_cols;
    ix->info.column_bytes = ix->n_cols * ix->col_stride;

    APPEND_MARK("columns");
    // ---- membership bitmaps: old bits copied, the new blocks' docs set, rank directories rebuilt ----
    if (base->n_bitmaps) {
        const uint64_t stride_words = (((uint64_t)n_new + 255) / 256) * 8 + 8;
        const size_t bits_bytes = (size_t)base->n_bitmaps * stride_words * 4, rank_bytes = (size_t)base->n_bitmaps * (stride_words / 8) * 4;
        uint32_t *d_bits = nullptr, *d_rank = nullptr;
        if ((rc = dev_alloc(bits_bytes, (void**)&d_bits))) return rc;
        if ((rc = dev_alloc(rank_bytes, (void**)&d_rank))) return rc;
        CU(cudaMemsetAsync(d_bits, 0, bits_bytes, ctx->stream));
        const uint64_t old_used_words = std::min<uint64_t>(base->bm_stride_words, stride_words);
        CU(cudaMemcpy2DAsync(d_bits, stride_words * 4, base->d_bits, base->bm_stride_words * 4, old_used_words * 4, base->n_bitmaps, cudaMemcpyDeviceToDevice, ctx->stream));
        std::vector<uint2> sel;
        for (size_t i = 0; i < touched.size(); i++) {
            const TermInfo& ti = ix->fields[touched[i].f].terms[touched[i].t];
            if (ti.bm < 0) continue;
            const uint32_t nnew = (uint32_t)enc[i].skips.size();
            for (uint32_t b = ti.n_blocks - nnew; b < ti.n_blocks; b++) sel.push_back(make_uint2(ti.blk_begin + b, (uint32_t)ti.bm));
        }
        DevTmp t_sel;
        CU(t_sel.alloc(sel.size() * sizeof(uint2)));
        if (!sel.empty()) CU(cudaMemcpyAsync(t_sel.p, sel.data(), sel.size() * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
        launch_bitmap_build(ix->dev, t_sel.as<uint2>(), (uint32_t)sel.size(), d_bits, stride_words, ctx->stream);
        launch_bitmap_rank(d_bits, d_rank, base->n_bitmaps, stride_words, ctx->stream);
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaGetLastError());
        ix->d_bits = d_bits;
        ix->d_rank = d_rank;
        ix->bm_stride_words = stride_words;
        ix->n_bitmaps = base->n_bitmaps;
        ix->info.bitmap_bytes = (uint64_t)ix->n_bitmaps * (stride_words * 4 + stride_words / 2);
    }
    APPEND_MARK("bitmaps");
    ix->info.n_bitmaps = ix->n_bitmaps;
    ix->info.n_postings = n_postings;
    ix->info.n_blocks = n_blocks;
    ix->info.packed_bytes = payload;
    ix->info.skip_bytes = n_blocks * 16;
    ix->info.n_docs = n_new;
    ix->info.n_fields = n_fields;
    if (ctx->env_timing) fprintf(stderr, "[fg_index_append] ms:%s\n", t_log.c_str());
#undef APPEND_MARK
    ix->info.appended_bytes_h2d = new_words.size() * 4 + new_skips.size() * sizeof(SkipEntry) + (old_ranges.size() + new_ranges.size()) * sizeof(uint3) +
                                  (uint64_t)n_seg * (ix->n_cols + 2) + (decode_list.size() + meta_list.size()) * 4;
    *out = ix.release();
    return FG_OK;
}

// ------------------------------------------------------------------------------------------
// plan lowering
// ------------------------------------------------------------------------------------------
struct fg_batch {
    fg_index* ix = nullptr;
    uint32_t n_queries = 0, n_items = 0, kcap = 0;
    int ks = 1;
    uint64_t sum_k = 0;
    DevQuery* d_queries = nullptr;
    DevLeaf* d_leaves = nullptr;
    DevItem* d_items = nullptr;
    uint64_t* d_partial = nullptr;
    uint32_t* d_partial_count = nullptr;
    unsigned long long* d_stats = nullptr;
    uint32_t* d_qtheta = nullptr;
    size_t sz[7] = {0, 0, 0, 0, 0, 0, 0};
    uint32_t class_count[NCLS] = {};
    uint64_t n_launches = 0;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};  // before search, after search, after merge
    cudaEvent_t ev_up = nullptr, ev_done = nullptr;   // plan uploaded / results staged on the host
    // fg_batch_submit: device result block, its page-locked host copy, layout
    void* d_out = nullptr;
    void* h_out = nullptr;
    size_t out_sz = 0;
    uint32_t sub_k_stride = 0;
    bool sub_counts = false;
    // lead-driven evaluation (fg_lead.cu): the default lowering
    bool lead = false;
    LQuery* l_queries = nullptr;
    LLeaf* l_leaves = nullptr;
    LItem* l_items = nullptr;
    uint32_t* l_state = nullptr;   // [1 work | n_cursors | qtheta n_queries | qcount n_queries | qmatch n_queries]
    uint32_t n_cursors = 0;
    uint64_t partial_entries = 0;
    size_t lsz[4] = {0, 0, 0, 0};
    void* d_plan = nullptr;        // one device block [LQuery | LLeaf | LItemRec] (l_queries / l_leaves point into it; l_items is expanded from the records on the device)
    void* h_plan = nullptr;        // its page-locked source (kept until release: the upload is asynchronous)
    size_t plan_sz = 0;
    uint64_t* d_sel = nullptr;     // deep-page batches (ks == 0): scratch of lead_select_kernel
    cudaStream_t exec_stream = nullptr;  // fg_batch_submit: the stream this batch runs on (null = the context's stream)
    uint32_t combine_k = 0;        // fg_search_union_of: the queries are the disjuncts of ONE query with this page limit
    uint32_t combine_filters = 0;  // ... the last combine_filters of them are filter children (fg_search_union_of_filtered)
    uint64_t* d_comb = nullptr;    // ... and its two scratch arrays of comb_cap2 keys
    uint32_t comb_cap2 = 0;
    size_t comb_sz = 0;
    size_t sel_sz = 0;
    std::vector<int32_t> qstatus;  // FG_PREP_PER_QUERY_STATUS: per-query lowering status
    std::string first_bad;
};

static double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static uint64_t env_u64(const char* name, uint64_t dflt) {
    const char* e = getenv(name);
    return e ? strtoull(e, nullptr, 10) : dflt;
}

extern "C" void fg_batch_release(fg_batch* b) {
    if (!b) return;
    if (b->ix && b->ix->ctx) cudaSetDevice(b->ix->ctx->device);
    if (b->ix && b->ix->ctx) {
        fg_ctx* c = b->ix->ctx;
        cudaStreamSynchronize(c->stream);  // nothing in flight may still use the blocks we recycle
        if (b->exec_stream) cudaStreamSynchronize(b->exec_stream);
        pool_free(c, b->d_queries, b->sz[0]); pool_free(c, b->d_leaves, b->sz[1]); pool_free(c, b->d_items, b->sz[2]);
        pool_free(c, b->d_partial, b->sz[3]); pool_free(c, b->d_partial_count, b->sz[4]);
        pool_free(c, b->d_stats, b->sz[5]); pool_free(c, b->d_qtheta, b->sz[6]);
        pool_free(c, b->d_out, b->out_sz);
        pinned_free(c, b->h_out, b->out_sz);
        if (b->ev_up) cudaEventSynchronize(b->ev_up);  // the plan upload reads h_plan
        pool_free(c, b->d_plan, b->plan_sz);
        pinned_free(c, b->h_plan, b->plan_sz);
        pool_free(c, b->l_state, b->lsz[3]);
        pool_free(c, b->l_items, b->lsz[2]);
        pool_free(c, b->d_sel, b->sel_sz);
        pool_free(c, b->d_comb, b->comb_sz);
    }
    for (auto& e : b->ev) if (e) cudaEventDestroy(e);
    if (b->ev_up) cudaEventDestroy(b->ev_up);
    if (b->ev_done) cudaEventDestroy(b->ev_done);
    delete b;
}



// ------------------------------------------------------------------------------------------
// plan lowering for the lead-driven kernels (fg_lead.cu; the scheme is described in fg_internal.h)
// ------------------------------------------------------------------------------------------
constexpr int LKEYS = 16 * 32;  // item sort keys: lead index (clamped to 15), then list length (log2)
struct LeadPart {
    std::vector<LLeaf> leaves;
    std::vector<LItem> items;       // one per (query, lead group); cursor = index of the (query, lead) pair inside this part
    std::vector<uint32_t> item_key; // sort key: lead index, then list length
    std::vector<uint32_t> item_copies;  // copies of the item in the queue (a long lead is walked by several warps)
    uint64_t n_item_copies = 0;
    std::vector<uint32_t> q_items;  // items per query of this part
    uint32_t key_count[LKEYS];
    uint32_t n_cursors = 0, kmax = 1;
    uint64_t sum_k = 0;
    int32_t rc = FG_OK;
    std::string err;
    void reset() {
        leaves.clear(); items.clear(); item_key.clear(); item_copies.clear(); q_items.clear();
        memset(key_count, 0, sizeof(key_count));
        n_cursors = 0; kmax = 1; sum_k = 0; rc = FG_OK; err.clear(); n_item_copies = 0;
    }
};
struct LeadScratch {
    std::vector<LeadPart> parts;
    std::vector<LQuery> lq;
};
static std::unique_ptr<LeadScratch> take_scratch(fg_ctx* c) {
    std::lock_guard<std::mutex> g(c->pool_mu);
    if (!c->lead_scratch.empty()) {
        std::unique_ptr<LeadScratch> s = std::move(c->lead_scratch.back());
        c->lead_scratch.pop_back();
        return s;
    }
    return std::unique_ptr<LeadScratch>(new LeadScratch());
}
static void give_scratch(fg_ctx* c, std::unique_ptr<LeadScratch> s) {
    if (!s) return;
    std::lock_guard<std::mutex> g(c->pool_mu);
    if (c->lead_scratch.size() < 8) c->lead_scratch.push_back(std::move(s));
}

static inline uint32_t host_sortable(float f) {
    uint32_t b;
    memcpy(&b, &f, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

static int32_t prepare_lead(fg_index* ix, const fg_query_batch* qb, uint32_t prep_flags, fg_batch** out) {
    fg_ctx* ctx = ix->ctx;
    const bool use_cols = !(prep_flags & FG_PREP_NO_COLUMNS) && !ctx->env_no_columns;  // tf columns and membership bitmaps
    const uint32_t PAR_BLOCKS_UNION = ctx->lead_par_blocks, MAX_PAR = ctx->lead_max_par;
    constexpr int MAXC = 40;
    struct CRec { uint32_t occur, begin, count; uint64_t df; };
    using Part = LeadPart;
    // scratch vectors are recycled through the context: fresh multi-megabyte vectors cost more in page faults than the
    // lowering itself
    std::unique_ptr<LeadScratch> scratch = take_scratch(ctx);
    struct Giveback {
        fg_ctx* c; std::unique_ptr<LeadScratch>& s;
        ~Giveback() { give_scratch(c, std::move(s)); }
    } giveback{ctx, scratch};
    std::vector<LQuery>& lq = scratch->lq;
    lq.resize(qb->n_queries);
    const bool per_query = (prep_flags & FG_PREP_PER_QUERY_STATUS) != 0;
    std::vector<int32_t> qstatus(per_query ? qb->n_queries : 0, FG_OK);
    std::string first_bad;  // (written by the thread that owns the query's part; read after the join)
    auto lfail = [](Part& o, int32_t code, const char* fmt, ...) -> int32_t {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        o.rc = code;
        o.err = buf;
        return code;
    };
    const std::vector<float>& topbm = *ix->topbm;
    const bool has_deletes = ix->dev.alive != nullptr;
    auto lower_range = [&](uint32_t q_begin, uint32_t q_end, Part& o) -> int32_t {
        LLeaf tmp[LMAX_LEAVES * 2];
        uint32_t ttop[LMAX_LEAVES * 2], ttopn[LMAX_LEAVES * 2];
        CRec crec[MAXC];
        o.reset();
        o.q_items.assign(q_end - q_begin, 0);
        o.leaves.reserve((size_t)((uint64_t)qb->n_leaves * (q_end - q_begin) / std::max<uint32_t>(qb->n_queries, 1)) + 64);
        auto lower_query = [&](uint32_t qi) -> int32_t {
            const fg_query& q = qb->queries[qi];
            LQuery& D = lq[qi];
            memset(&D, 0, sizeof(D));
            D.k = 1;
            if (q.k == 0) return lfail(o, FG_ERR_INVALID, "query %u: k == 0 (TopDocs::with_limit requires limit >= 1)", qi);
            if ((uint64_t)q.clause_begin + q.n_clauses > qb->n_clauses)
                return lfail(o, FG_ERR_INVALID, "query %u: clause range out of bounds", qi);
            o.kmax = std::max(o.kmax, q.k);
            o.sum_k += q.k;
            D.k = q.k;
            D.leaf_begin = (uint32_t)o.leaves.size();
            int nc = 0, nt = 0;
            int must_idx[MAXC], n_must = 0, n_should = 0, n_not = 0;
            float const_score = 0.f;
            bool empty = false, has_all_only = false, positive = true;
            for (uint32_t ci = 0; ci < q.n_clauses; ci++) {
                const fg_clause& c = qb->clauses[q.clause_begin + ci];
                if ((uint64_t)c.leaf_begin + c.n_leaves > qb->n_leaves)
                    return lfail(o, FG_ERR_INVALID, "query %u: leaf range out of bounds", qi);
                if (c.occur > FG_OCCUR_MUST_NOT) return lfail(o, FG_ERR_INVALID, "query %u: bad occur", qi);
                CRec cr{c.occur, (uint32_t)nt, 0, 0};
                bool all = false;
                float all_boost = 0.f;
                for (uint32_t li = 0; li < c.n_leaves; li++) {
                    const fg_leaf& lf = qb->leaves[c.leaf_begin + li];
                    if (lf.term_ord == FG_TERM_ALL) { all = true; all_boost += lf.boost; continue; }
                    if (lf.term_ord == FG_TERM_MISSING) continue;
                    if (lf.field >= ix->fields.size()) return lfail(o, FG_ERR_INVALID, "query %u: field %u out of range", qi, lf.field);
                    const HostField& hf = ix->fields[lf.field];
                    // a term the planner's dictionary learned after this snapshot was taken (a commit raced the request):
                    // the snapshot does not contain it -- an empty scorer, like an uncommitted document in the reference
                    if (lf.term_ord >= hf.n_terms) continue;
                    const TermInfo& ti = hf.terms[lf.term_ord];
                    if (ti.df_global == 0 || ti.n_blocks == 0) continue;  // empty scorer (globally, or in this shard)
                    if (nt >= LMAX_LEAVES * 2) return lfail(o, FG_ERR_UNSUPPORTED, "query %u: too many leaves", qi);
                    LLeaf& L = tmp[nt];
                    memset(&L, 0, sizeof(L));
                    L.blk_begin = ti.blk_begin;
                    L.n_blocks = ti.n_blocks;
                    L.weight = lf.boost * ti.idf_w;
                    L.cnorm = hf.cnorm;
                    L.fn_field = (hf.flags & FG_FIELD_HAS_FIELDNORMS) ? (int32_t)lf.field : -1;
                    L.df = ti.df_local;
                    if (c.occur != FG_OCCUR_MUST_NOT) {
                        if (!(L.weight >= 0.f)) positive = false;
                        L.ub = L.weight * ti.max_factor;
                    }
                    if (use_cols && ti.col >= 0) L.col = ix->d_cols + (uint64_t)ti.col * ix->col_stride;
                    else if (use_cols && ti.bm >= 0) {
                        L.bits = ix->d_bits + (uint64_t)ti.bm * ix->bm_stride_words;
                        L.rank = ix->d_rank + (uint64_t)ti.bm * (ix->bm_stride_words / 8);
                    }
                    ttop[nt] = ti.top_off;
                    ttopn[nt] = ti.top_n;
                    nt++;
                    cr.df += ti.df_local;
                    cr.count++;
                }
                if (all) {
                    if (c.occur == FG_OCCUR_MUST && cr.count == 0) { const_score += all_boost; has_all_only = true; continue; }
                    return lfail(o, FG_ERR_UNSUPPORTED, "query %u: AllQuery leaf outside a Must clause is not supported", qi);
                }
                if (cr.count == 0) {
                    if (c.occur == FG_OCCUR_MUST) empty = true;
                    continue;
                }
                if (nc >= MAXC) return lfail(o, FG_ERR_UNSUPPORTED, "query %u: too many clauses", qi);
                if (c.occur == FG_OCCUR_MUST) must_idx[n_must++] = nc;
                else if (c.occur == FG_OCCUR_SHOULD) n_should++;
                else n_not++;
                crec[nc++] = cr;
            }
            D.const_score = const_score;
            if (has_all_only && n_must == 0 && !empty) {
                if (n_should == 0 && n_not == 0) {  // pure AllQuery: every alive doc, score = boost
                    D.flags |= LQ_ALL;
                    return FG_OK;
                }
                return lfail(o, FG_ERR_UNSUPPORTED, "query %u: AllQuery Must with only Should/MustNot siblings not supported", qi);
            }
            if (empty || (n_must == 0 && n_should == 0)) return FG_OK;
            if (nt > LMAX_LEAVES) return lfail(o, FG_ERR_UNSUPPORTED, "query %u: %d live leaves > %d", qi, nt, LMAX_LEAVES);
            // Must clauses by ascending Sum(df) = tantivy's Intersection order (stable insertion sort)
            for (int i = 1; i < n_must; i++) {
                const int x = must_idx[i];
                int j = i - 1;
                while (j >= 0 && crec[must_idx[j]].df > crec[x].df) { must_idx[j + 1] = must_idx[j]; j--; }
                must_idx[j + 1] = x;
            }
            // ---- leads, shortest list first ----
            // A lead's bound carries the upper bounds of the leads walked AFTER it (a doc of an earlier lead is
            // scored there). Short lists are cheap to walk completely and usually carry the large idf weights:
            // walked first, they drop out of the bounds of the long lists, whose own weights are small -- the
            // long lists are then pruned by their own block maxima alone (on a 1 M-doc Zipfian mix 3x fewer
            // blocks and 7x fewer candidates than descending-upper-bound order).
            int lead_src[LMAX_LEAVES], nl = 0;
            if (n_must) {
                const CRec& cr = crec[must_idx[0]];
                for (uint32_t i = 0; i < cr.count; i++) lead_src[nl++] = (int)(cr.begin + i);
            } else {
                for (int ci = 0; ci < nc; ci++)
                    if (crec[ci].occur == FG_OCCUR_SHOULD)
                        for (uint32_t i = 0; i < crec[ci].count; i++) lead_src[nl++] = (int)(crec[ci].begin + i);
            }
            for (int i = 1; i < nl; i++) {
                const int x = lead_src[i];
                int j = i - 1;
                while (j >= 0 && tmp[lead_src[j]].df > tmp[x].df) { lead_src[j + 1] = lead_src[j]; j--; }
                lead_src[j + 1] = x;
            }
            const size_t l0 = o.leaves.size();
            for (int i = 0; i < nl; i++) {
                LLeaf L = tmp[lead_src[i]];
                L.role = LR_LEAD;
                o.leaves.push_back(L);
            }
            uint32_t n_req = 0, n_opt = 0;
            for (int ci = 1; ci < n_must; ci++) {
                const CRec& cr = crec[must_idx[ci]];
                for (uint32_t i = 0; i < cr.count; i++) {
                    LLeaf L = tmp[cr.begin + i];
                    L.role = LR_REQ | ((uint32_t)ci << 8);
                    o.leaves.push_back(L);
                    n_req++;
                }
            }
            if (n_must)
                for (int ci = 0; ci < nc; ci++)
                    if (crec[ci].occur == FG_OCCUR_SHOULD)
                        for (uint32_t i = 0; i < crec[ci].count; i++) {
                            LLeaf L = tmp[crec[ci].begin + i];
                            L.role = LR_OPT;
                            o.leaves.push_back(L);
                            n_opt++;
                        }
            for (int ci = 0; ci < nc; ci++)
                if (crec[ci].occur == FG_OCCUR_MUST_NOT)
                    for (uint32_t i = 0; i < crec[ci].count; i++) {
                        LLeaf L = tmp[crec[ci].begin + i];
                        L.role = LR_NOT;
                        L.ub = 0.f;
                        o.leaves.push_back(L);
                    }
            D.n_leaves = (uint32_t)(o.leaves.size() - l0);
            {
                uint32_t slot = 0;  // decoded-block cache slots of the leaves that are looked up by gallop + decode
                for (size_t i = l0; i < o.leaves.size(); i++)
                    if (!o.leaves[i].col && !o.leaves[i].bits) o.leaves[i].slot = slot++;
            }
            D.n_lead = (uint32_t)nl;
            D.n_req = n_req;
            D.n_opt = n_opt;
            {
                uint64_t ld = 0;  // a document is offered at most once, by one of the leads
                for (int i = 0; i < nl; i++) ld += o.leaves[l0 + i].df;
                D.lead_docs = (uint32_t)std::min<uint64_t>(ld, ix->n_docs);
            }
            // rest of a lead = what a candidate of it can still collect: later leads + required + optional leaves
            float tail = 0.f, total = 0.f;
            for (uint32_t i = (uint32_t)nl; i < (uint32_t)nl + n_req + n_opt; i++) tail += o.leaves[l0 + i].ub;
            total = tail;
            for (int i = nl - 1; i >= 0; i--) {
                o.leaves[l0 + i].rest = tail;
                tail += o.leaves[l0 + i].ub;
            }
            total = tail;
            D.slack = total * 1e-5f + 1e-30f;
            {
                const float ubt = total * 1.0001f + const_score;  // no score of this query lies above it
                uint32_t bits;
                memcpy(&bits, &ubt, 4);
                D.hist_base = ubt > 0.f ? std::max<uint32_t>(bits >> LHIST_SHIFT, (uint32_t)LHIST_B) : (uint32_t)LHIST_B;
            }
            if (positive) D.flags |= LQ_PRUNE;
            // a lower bound of the k-th best score known before anything runs: the k-th largest block maximum of
            // a lead of a pure union (k distinct docs reach it with that one leaf alone, the others only add)
            if (positive && n_must == 0 && n_not == 0 && !has_deletes) {
                float t0 = 0.f;
                for (int i = 0; i < nl; i++) {
                    const int sidx = lead_src[i];
                    if (ttopn[sidx] >= q.k) t0 = std::max(t0, tmp[sidx].weight * topbm[ttop[sidx] + q.k - 1]);
                }
                if (t0 > 0.f) D.theta0 = host_sortable(t0);
            }
            // ---- work items ----
            // Short leads that follow one another share one item (one warp walks them in order); a long lead gets an item
            // of its own in `par` copies (one warp each) that share its block cursor. (A lead that cannot contribute a hit
            // given the threshold known at lowering time still gets its place: whether the execution may prune is decided
            // per execution; in the pruned form such a lead ends at its first threshold test.)
            uint32_t qitems = 0;
            const float bound_slack = D.slack + const_score;
            // (required clauses: every posting of a lead block is looked up in them -- a block is a whole round of dependent
            // gathers, so such leads are shared out among warps in much smaller pieces than the mostly skipped blocks of a pruned union)
            const uint32_t PAR_BLOCKS = n_req ? std::min<uint32_t>(PAR_BLOCKS_UNION, std::max<uint32_t>(2u, ctx->lead_chunk_req / n_req))
                                              : std::min<uint32_t>(PAR_BLOCKS_UNION, std::max<uint32_t>(2u, ctx->lead_union_work / std::max<uint32_t>(1u, D.n_leaves - 1)));
            int i = 0;
            while (i < nl) {
                const LLeaf& L0 = o.leaves[l0 + i];
                const uint32_t nb = L0.n_blocks;
                uint32_t par = 1, lg = 31u - (uint32_t)__builtin_clz(nb | 1u);
                int j = i + 1;
                if (nb > PAR_BLOCKS) {
                    // (every copy appends up to k hits to the query's partial region: fewer copies for deep pages)
                    const uint32_t max_par = n_req ? std::max<uint32_t>(4, std::min<uint32_t>(ctx->lead_max_par_req, 16384u / q.k))
                                                   : std::max<uint32_t>(4, std::min<uint32_t>(MAX_PAR, 4096u / q.k));
                    par = std::max<uint32_t>(1, std::min<uint32_t>(max_par, (nb + PAR_BLOCKS - 1) / PAR_BLOCKS));
                } else {
                    uint32_t sum = nb;
                    while (j < nl && sum + o.leaves[l0 + j].n_blocks <= PAR_BLOCKS) {
                        sum += o.leaves[l0 + j].n_blocks;
                        j++;
                    }
                    lg = 31u - (uint32_t)__builtin_clz(sum | 1u);
                }
                const uint32_t key = std::min<uint32_t>((uint32_t)i, 15u) * 32u + (31u - lg);
                float item_bound = -INFINITY;  // (only meaningful when the bounds hold: LQ_PRUNE)
                for (int l = i; l < j; l++) item_bound = std::max(item_bound, o.leaves[l0 + l].ub + o.leaves[l0 + l].rest + bound_slack);
                if (!positive) item_bound = INFINITY;
                o.items.push_back(LItem{qi, (uint32_t)i | ((uint32_t)j << 16), o.n_cursors, item_bound});
                o.item_key.push_back(key);
                o.item_copies.push_back(par);
                o.key_count[key] += par;
                o.n_item_copies += par;
                o.n_cursors += (uint32_t)(j - i);
                qitems += par;
                i = j;
            }
            o.q_items[qi - q_begin] = qitems;
            return FG_OK;
        };
        for (uint32_t qi = q_begin; qi < q_end; qi++) {
            const int32_t rc = lower_query(qi);
            if (rc == FG_OK) continue;
            if (!per_query) return rc;
            // FG_PREP_PER_QUERY_STATUS: the offender becomes an empty query, its siblings are answered
            qstatus[qi] = rc;
            if (first_bad.empty()) first_bad = o.err;
            o.rc = FG_OK;
            LQuery& D = lq[qi];
            memset(&D, 0, sizeof(D));
            D.k = 1;
            D.leaf_begin = (uint32_t)o.leaves.size();
        }
        return FG_OK;
    };
    const double t_begin = now_ms();
    const int LT = (int)std::max<uint64_t>(1, std::min<uint64_t>({(uint64_t)HostPool::get().size(), 16, (uint64_t)qb->n_queries / 96 + 1}));
    std::vector<Part>& parts = scratch->parts;
    if ((int)parts.size() < LT) parts.resize((size_t)LT);
    auto q_lo = [&](int t) { return (uint32_t)((uint64_t)qb->n_queries * t / LT); };
    HostPool::get().run(LT, [&](int t) { lower_range(q_lo(t), q_lo(t + 1), parts[t]); });
    for (int t = 0; t < LT; t++)
        if (parts[t].rc != FG_OK) return fail(parts[t].rc, "%s", parts[t].err.c_str());
    // ---- layout of the concatenated plan: per-part bases, item positions by sort key, partial regions ----
    // Items are ordered by (lead index, list length): every query's first lead is queued ahead of all second leads,
    // and so on: a later lead mostly finds the query's threshold already set.
    std::vector<uint32_t> leaf_base((size_t)LT + 1, 0), cur_base((size_t)LT + 1, 0);
    std::vector<uint32_t> key_start((size_t)LT * LKEYS);
    uint32_t kmax = 1;
    uint64_t sum_k = 0, part_entries = 0, ni_tot = 0, nrec_tot = 0, sel_entries = 0;
    for (int t = 0; t < LT; t++) {
        leaf_base[t + 1] = leaf_base[t] + (uint32_t)parts[t].leaves.size();
        cur_base[t + 1] = cur_base[t] + parts[t].n_cursors;
        kmax = std::max(kmax, parts[t].kmax);
        sum_k += parts[t].sum_k;
        ni_tot += parts[t].n_item_copies;
        nrec_tot += parts[t].items.size();
    }
    {
        uint32_t run = 0;
        for (int key = 0; key < LKEYS; key++)
            for (int t = 0; t < LT; t++) {
                key_start[(size_t)t * LKEYS + key] = run;
                run += parts[t].key_count[key];
            }
    }
    for (int t = 0; t < LT; t++) {
        const uint32_t q0 = q_lo(t), q1 = q_lo(t + 1);
        for (uint32_t qi = q0; qi < q1; qi++) {
            LQuery& D = lq[qi];
            D.leaf_begin += leaf_base[t];
            D.part_begin = (uint32_t)std::min<uint64_t>(part_entries, 0xFFFFFFFFull);
            if (kmax > 1024) {
                // deep pages: the warps append every accepted hit (no per-warp queue); at most lead_docs of them exist.
                // A scratch region for the selected page follows (lead_select_kernel sorts it: power of two).
                D.part_cap = D.lead_docs;
                uint64_t cap2 = 1;
                while (cap2 < std::min<uint64_t>(D.k, D.part_cap)) cap2 <<= 1;
                D.sel_begin = (uint32_t)std::min<uint64_t>(sel_entries, 0xFFFFFFFFull);
                sel_entries += cap2;
            } else {
                D.part_cap = parts[t].q_items[qi - q0] * D.k;
            }
            part_entries += D.part_cap;
        }
    }
    // (deep pages share a batch with other queries at the price of list space for all of them: callers batch them apart)
    if (part_entries > 0x7FFFFFF0ull || sel_entries > 0x7FFFFFF0ull || ni_tot > 0xFFFFFFF0ull)
        return fail(FG_ERR_UNSUPPORTED, "partial result lists exceed 2^31 entries%s", kmax > 1024 ? " (a batch with a page limit above 1024 keeps every match of every query: send deep pages in batches of their own)" : "");
    const uint32_t n_cursors = cur_base[LT];
    const size_t nl_tot = leaf_base[LT];
    const size_t off_leaves = ((size_t)qb->n_queries * sizeof(LQuery) + 255) & ~(size_t)255;
    // uploaded: [LQuery | LLeaf | LItemRec]; the item array itself (one entry per copy) is written on the device
    const size_t off_recs = (off_leaves + nl_tot * sizeof(LLeaf) + 255) & ~(size_t)255;
    const size_t plan_bytes = off_recs + nrec_tot * sizeof(LItemRec) + 256;
    std::vector<uint32_t> rec_base((size_t)LT + 1, 0);
    for (int t = 0; t < LT; t++) rec_base[t + 1] = rec_base[t] + (uint32_t)parts[t].items.size();

    CU(cudaSetDevice(ctx->device));
    std::unique_ptr<fg_batch, void (*)(fg_batch*)> b(new fg_batch(), fg_batch_release);
    b->ix = ix;
    b->lead = true;
    b->n_queries = qb->n_queries;
    b->n_items = (uint32_t)ni_tot;
    b->kcap = kmax;
    b->ks = kmax <= 32 ? 1 : kmax <= 128 ? 4 : kmax <= 1024 ? 32 : 0;  // 0: deep pages (append + select)
    b->sum_k = sum_k;
    b->n_cursors = n_cursors;
    b->partial_entries = part_entries;
    b->qstatus.swap(qstatus);
    b->first_bad.swap(first_bad);
    // the plan is assembled in page-locked memory (one block, one asynchronous copy) by the same worker threads
    b->plan_sz = plan_bytes;
    CU(pinned_alloc(ctx, &b->h_plan, plan_bytes));
    char* hp = (char*)b->h_plan;
    if (qb->n_queries) memcpy(hp, lq.data(), (size_t)qb->n_queries * sizeof(LQuery));
    HostPool::get().run(LT, [&](int t) {
        Part& o = parts[t];
        if (!o.leaves.empty()) memcpy(hp + off_leaves + (size_t)leaf_base[t] * sizeof(LLeaf), o.leaves.data(), o.leaves.size() * sizeof(LLeaf));
        LItemRec* recs = reinterpret_cast<LItemRec*>(hp + off_recs) + rec_base[t];  // (sequential writes: the queue order is in `dst`)
        uint32_t* pos = key_start.data() + (size_t)t * LKEYS;
        for (size_t i = 0; i < o.items.size(); i++) {
            LItemRec r{o.items[i], pos[o.item_key[i]], o.item_copies[i]};
            r.item.cursor += cur_base[t];
            pos[o.item_key[i]] += o.item_copies[i];
            recs[i] = r;
        }
    });
    const double t_lowered = now_ms();
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(pool_alloc(ctx, &b->d_plan, plan_bytes));
    CU(cudaMemcpyAsync(b->d_plan, b->h_plan, plan_bytes, cudaMemcpyHostToDevice, ctx->up));
    b->l_queries = reinterpret_cast<LQuery*>((char*)b->d_plan);
    b->l_leaves = reinterpret_cast<LLeaf*>((char*)b->d_plan + off_leaves);
    b->lsz[2] = std::max<size_t>(ni_tot * sizeof(LItem), 16);
    CU(pool_alloc(ctx, (void**)&b->l_items, b->lsz[2]));
    launch_expand_items(reinterpret_cast<const LItemRec*>((char*)b->d_plan + off_recs), (uint32_t)nrec_tot, b->l_items, ctx->up);
    b->lsz[3] = ((size_t)4 + n_cursors + (3 + (size_t)LHIST_B) * (size_t)b->n_queries) * 4 + 32;
    CU(pool_alloc(ctx, (void**)&b->l_state, b->lsz[3]));
    b->sz[3] = std::max<size_t>((size_t)part_entries * 8, 16);
    CU(pool_alloc(ctx, (void**)&b->d_partial, b->sz[3]));
    b->sz[5] = 32 * sizeof(unsigned long long);
    CU(pool_alloc(ctx, (void**)&b->d_stats, b->sz[5]));
    if (sel_entries) {
        b->sel_sz = (size_t)sel_entries * 8;
        CU(pool_alloc(ctx, (void**)&b->d_sel, b->sel_sz));
    }
    for (auto& e : b->ev) CU(cudaEventCreate(&e));
    CU(cudaEventCreateWithFlags(&b->ev_up, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&b->ev_done, cudaEventDisableTiming));
    CU(cudaEventRecord(b->ev_up, ctx->up));  // (the page-locked block lives as long as the batch: no wait here)
    if (ctx->env_timing)
        fprintf(stderr, "[prepare_lead] lowering + assembly %.2f ms (%d threads), alloc + upload %.2f ms (%llu items, %zu leaves)\n", t_lowered - t_begin, LT,
                now_ms() - t_lowered, (unsigned long long)ni_tot, nl_tot);
    *out = b.release();
    return FG_OK;
}

extern "C" int32_t fg_batch_prepare(fg_index* ix, const fg_query_batch* qb, fg_batch** out) {
    return fg_batch_prepare_ex(ix, qb, 0, out);
}

extern "C" int32_t fg_batch_prepare_ex(fg_index* ix, const fg_query_batch* qb, uint32_t prep_flags, fg_batch** out) {
    if (!ix || !qb || !out) return fail(FG_ERR_INVALID, "fg_batch_prepare: NULL argument");
    *out = nullptr;
    if (qb->n_queries && (!qb->queries || (qb->n_clauses && !qb->clauses) || (qb->n_leaves && !qb->leaves)))
        return fail(FG_ERR_INVALID, "fg_batch_prepare: NULL arrays");
    if (!(prep_flags & FG_PREP_LEGACY) && !ix->ctx->env_legacy) return prepare_lead(ix, qb, prep_flags, out);
    const bool use_cols = !(prep_flags & FG_PREP_NO_COLUMNS) && ix->n_cols && !ix->ctx->env_no_columns;
    const uint64_t COL_COST_DIV = std::max<uint64_t>(1, env_u64("FG_COL_COST_DIV", 16));
    const uint64_t COL_COST_DIV_PHASES = std::max<uint64_t>(1, env_u64("FG_COL_COST_DIV_PHASES", 2));
    const uint64_t WINDOW_COST = env_u64("FG_WINDOW_COST", 4096);
    const bool USE_COLSCAN = env_u64("FG_COLSCAN", 1) != 0;
    const uint64_t STREAM_MAX_BPW = env_u64("FG_STREAM_MAX_BPW", 6);  // blocks per dense window of a streamed leaf
    *out = nullptr;
    if (qb->n_queries && (!qb->queries || (qb->n_clauses && !qb->clauses) || (qb->n_leaves && !qb->leaves)))
        return fail(FG_ERR_INVALID, "fg_batch_prepare: NULL arrays");
    const double t_begin = now_ms();
    const uint64_t ITEM_BYTES = env_u64("FG_ITEM_BYTES", 131072);
    const uint64_t ITEM_BYTES_HASH = env_u64("FG_ITEM_BYTES_HASH", 24576);      // pure unions in hash mode
    const uint64_t ITEM_BYTES_MASKED = env_u64("FG_ITEM_BYTES_MASKED", 98304);  // plans with Must / MustNot clauses
    const uint64_t DENSE_MIN = env_u64("FG_DENSE_MIN", 1024);  // insert postings per dense window
    const uint64_t DENSE_MIN_MUST = env_u64("FG_DENSE_MIN_MUST", 256);  // same, plans with Must clauses
    const uint32_t HASH_MIN_SPAN = 4096;

    std::vector<DevQuery> dq(qb->n_queries);
    const uint64_t SOLO_MIN_BLOCKS = env_u64("FG_SOLO_MIN_BLOCKS", 0xFFFFFFFFull);
    constexpr int MAXC = 32, MAXT = 64;
    struct CRec { uint32_t occur, begin, count, ncol; uint64_t cost, df; };
    // The lowering of one query is independent of the others: large batches are lowered by several host
    // threads into per-thread leaf / item arrays that are concatenated afterwards (leaf_begin, item_begin
    // and item slots are rebased).
    struct LowerOut {
        std::vector<DevLeaf> dl;
        std::vector<DevItem> items;
        std::vector<uint64_t> item_cost;
        uint32_t kmax = 1;
        uint64_t sum_k = 0, n_colscan_items = 0;
        int32_t rc = FG_OK;
        std::string err;
    };
    auto lfail = [](LowerOut& o, int32_t code, const char* fmt, ...) -> int32_t {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        o.rc = code;
        o.err = buf;
        return code;
    };
    auto lower_range = [&](uint32_t q_begin, uint32_t q_end, LowerOut& o) -> int32_t {
        std::vector<DevLeaf>& dl = o.dl;
        std::vector<DevItem>& items = o.items;
        std::vector<uint64_t>& item_cost = o.item_cost;
        uint32_t& kmax = o.kmax;
        uint64_t& sum_k = o.sum_k;
        uint64_t& n_colscan_items = o.n_colscan_items;
        // per-query scratch (fixed arrays: the lowering of a 5000-query batch must not allocate per query)
        DevLeaf ctmp[MAXT];  // column leaves of the query (appended after its block leaves)
        CRec crec[MAXC];
        DevLeaf tmp[MAXT];
        const size_t nq_part = q_end - q_begin;
        dl.reserve((size_t)((uint64_t)qb->n_leaves * nq_part / std::max<uint32_t>(qb->n_queries, 1)) + 64);
        items.reserve(nq_part * 4);
        item_cost.reserve(nq_part * 4);
        for (uint32_t qi = q_begin; qi < q_end; qi++) {
            const fg_query& q = qb->queries[qi];
            if (q.k == 0) return lfail(o, FG_ERR_INVALID, "query %u: k == 0 (TopDocs::with_limit requires limit >= 1)", qi);
            if (q.k > 1024) return lfail(o, FG_ERR_UNSUPPORTED, "query %u: k = %u > 1024: deep pages run on the lead-driven kernels (default lowering)", qi, q.k);
            if ((uint64_t)q.clause_begin + q.n_clauses > qb->n_clauses)
                return lfail(o, FG_ERR_INVALID, "query %u: clause range out of bounds", qi);
            kmax = std::max(kmax, q.k);
            sum_k += q.k;
            int nc = 0, nt = 0;          // live clauses / leaves of this query
            int must_idx[MAXC], n_must = 0, n_should = 0, n_not = 0;
            float const_score = 0.f;
            bool empty = false;          // a Must clause that can never match
            bool has_all_only = false;
            for (uint32_t ci = 0; ci < q.n_clauses; ci++) {
                const fg_clause& c = qb->clauses[q.clause_begin + ci];
                if ((uint64_t)c.leaf_begin + c.n_leaves > qb->n_leaves)
                    return lfail(o, FG_ERR_INVALID, "query %u: leaf range out of bounds", qi);
                if (c.occur > FG_OCCUR_MUST_NOT) return lfail(o, FG_ERR_INVALID, "query %u: bad occur", qi);
                CRec cr{c.occur, (uint32_t)nt, 0, 0, 0, 0};
                bool all = false;
                float all_boost = 0.f;
                for (uint32_t li = 0; li < c.n_leaves; li++) {
                    const fg_leaf& lf = qb->leaves[c.leaf_begin + li];
                    if (lf.term_ord == FG_TERM_ALL) { all = true; all_boost += lf.boost; continue; }
                    if (lf.term_ord == FG_TERM_MISSING) continue;
                    if (lf.field >= ix->fields.size()) return lfail(o, FG_ERR_INVALID, "query %u: field %u out of range", qi, lf.field);
                    const HostField& hf = ix->fields[lf.field];
                    if (lf.term_ord >= hf.n_terms) continue;  // not in this snapshot's dictionary (see prepare_lead)
                    const TermInfo& ti = hf.terms[lf.term_ord];
                    if (ti.df_global == 0 || ti.n_blocks == 0) continue;  // empty scorer (globally, or in this shard)
                    if (nt >= MAXT) return lfail(o, FG_ERR_UNSUPPORTED, "query %u: too many leaves", qi);
                    DevLeaf& L = tmp[nt++];
                    memset(&L, 0, sizeof(L));
                    L.blk_begin = ti.blk_begin;
                    L.n_blocks = ti.n_blocks;
                    L.weight = lf.boost * ti.idf_w;
                    L.cnorm = hf.cnorm;
                    L.fn_field = (hf.flags & FG_FIELD_HAS_FIELDNORMS) ? (int32_t)lf.field : -1;
                    if (use_cols && ti.col >= 0) {
                        L.col = ix->d_cols + (uint64_t)ti.col * ix->col_stride;
                        cr.ncol++;
                    } else {
                        cr.cost += ti.bytes;
                    }
                    cr.df += ti.df_local;
                    cr.count++;
                }
                if (all) {
                    if (c.occur == FG_OCCUR_MUST && cr.count == 0) { const_score += all_boost; has_all_only = true; continue; }
                    return lfail(o, FG_ERR_UNSUPPORTED, "query %u: AllQuery leaf outside a Must clause is not supported", qi);
                }
                if (cr.count == 0) {
                    if (c.occur == FG_OCCUR_MUST) empty = true;
                    continue;
                }
                if (nc >= MAXC) return lfail(o, FG_ERR_UNSUPPORTED, "query %u: too many clauses", qi);
                if (c.occur == FG_OCCUR_MUST) must_idx[n_must++] = nc;
                else if (c.occur == FG_OCCUR_SHOULD) n_should++;
                else n_not++;
                crec[nc++] = cr;
            }
            DevQuery& D = dq[qi];
            memset(&D, 0, sizeof(D));
            D.k = q.k;
            D.const_score = const_score;
            D.leaf_begin = (uint32_t)dl.size();
            D.item_begin = (uint32_t)items.size();
            if (has_all_only && n_must == 0 && !empty) {
                if (n_should == 0 && n_not == 0) {  // pure AllQuery: every alive doc, score = boost
                    D.flags |= QF_ALL;
                    continue;
                }
                return lfail(o, FG_ERR_UNSUPPORTED, "query %u: AllQuery Must with only Should/MustNot siblings not supported", qi);
            }
            if (empty || (n_must == 0 && n_should == 0)) continue;
            if (n_must > MAX_MUST) return lfail(o, FG_ERR_UNSUPPORTED, "query %u: more than %d Must clauses", qi, MAX_MUST);
            // Must clauses by ascending Sum(df) = tantivy's Intersection order (stable insertion sort)
            for (int i = 1; i < n_must; i++) {
                const int x = must_idx[i];
                int j = i - 1;
                while (j >= 0 && crec[must_idx[j]].df > crec[x].df) { must_idx[j + 1] = must_idx[j]; j--; }
                must_idx[j + 1] = x;
            }
            const size_t ql0 = dl.size();
            uint64_t insert_postings = 0, total_bytes = 0;
            int n_ctmp = 0;
            bool col_insert = false;
            // Block leaves go to `dl` in evaluation order; column leaves are collected in ctmp and
            // appended behind them (they are applied in the slot scan, after every block phase).
            // `req` = mask bits a slot must already carry when a filter leaf touches it; only bits that
            // block phases decide completely may be required: a clause with a column leaf sets its bit
            // as late as the slot scan. A filter leaf left without any usable precondition is LF_NOFILT.
            auto emit = [&](const CRec& cr, uint32_t bit, uint32_t role, uint32_t req, bool ins, bool nofilt) {
                for (uint32_t i = 0; i < cr.count; i++) {
                    DevLeaf L = tmp[cr.begin + i];
                    L.bit = bit; L.role = role; L.req = req;
                    if (L.col) {
                        if (ins) col_insert = true;
                        ctmp[n_ctmp++] = L;
                        continue;
                    }
                    if (nofilt) L.lflags |= LF_NOFILT;
                    if (ins) insert_postings += (uint64_t)L.n_blocks * BLOCK;
                    dl.push_back(L);
                }
            };
            uint32_t complete = 0, need_not = 0;
            if (n_must) {
                D.all_must = (1u << n_must) - 1u;
                for (int ci = 0; ci < n_must; ci++) {
                    const CRec& cr = crec[must_idx[ci]];
                    const uint32_t req = ((1u << ci) - 1u) & complete;
                    emit(cr, 1u << ci, ci == 0 ? ROLE_INSERT : ROLE_MUST, req, ci == 0, ci != 0 && req == 0);
                    if (cr.ncol == 0) complete |= 1u << ci;
                    total_bytes += cr.cost;
                }
                D.n_insert = 0;
                for (size_t i = ql0; i < dl.size(); i++) D.n_insert += dl[i].role == ROLE_INSERT;
                for (int i = 0; i < nc; i++)
                    if (crec[i].occur == FG_OCCUR_SHOULD) { emit(crec[i], 0, ROLE_SHOULD, complete, false, complete == 0); total_bytes += crec[i].cost / 4; }
                need_not = complete;
            } else {
                D.all_must = BIT_SHOULD;
                D.flags |= QF_NO_MUST;
                bool any_col = false, positive = true;
                for (int i = 0; i < nc; i++)
                    if (crec[i].occur == FG_OCCUR_SHOULD) {
                        emit(crec[i], BIT_SHOULD, ROLE_INSERT, 0, true, false);
                        total_bytes += crec[i].cost;
                        any_col = any_col || crec[i].ncol;
                    }
                D.n_insert = (uint32_t)(dl.size() - ql0);
                for (size_t i = ql0; i < dl.size(); i++) positive = positive && dl[i].weight > 0.f;
                for (int i = 0; i < n_ctmp; i++) positive = positive && ctmp[i].weight > 0.f;
                if (n_not == 0 && positive) D.flags |= QF_PURE_UNION;
                complete = any_col ? 0u : BIT_SHOULD;
                need_not = complete;
            }
            for (int i = 0; i < nc; i++)
                if (crec[i].occur == FG_OCCUR_MUST_NOT) { emit(crec[i], BIT_NOT, ROLE_NOT, 0, false, need_not == 0); total_bytes += crec[i].cost / 4; }
            const size_t nbl = dl.size() - ql0;  // block leaves
            if (nbl + (size_t)n_ctmp > (size_t)MAX_LEAVES)
                return lfail(o, FG_ERR_UNSUPPORTED, "query %u: %zu live leaves > %d", qi, nbl + (size_t)n_ctmp, MAX_LEAVES);
            // candidate-bitmap rebuild points: after the last leaf of a clause when the next leaf filters
            for (size_t i = ql0; i + 1 < dl.size(); i++) {
                const DevLeaf& nx = dl[i + 1];
                if (nx.role == ROLE_INSERT || (nx.lflags & LF_NOFILT)) continue;
                const bool boundary = dl[i].role != nx.role || dl[i].bit != nx.bit;
                if (!boundary) continue;
                if (dl[i].role == ROLE_SHOULD) continue;  // bitmap of all-Must candidates is still valid
                // mask the NEXT leaf's docs must carry (MustNot filters on the matching candidates)
                dl[i].build_cb = nx.role == ROLE_NOT ? need_not : nx.req;
            }
            D.n_leaves = (uint32_t)nbl;
            D.n_stream = 0;
            D.n_col = (uint32_t)n_ctmp;
            D.col_req = n_must ? complete : 0u;
            if (col_insert) D.flags |= QF_COL_INSERT;
            for (int i = 0; i < n_ctmp; i++) dl.push_back(ctmp[i]);

            // ---- mode + work items ----
            const uint32_t nd = ix->n_docs;
            const uint64_t dmin = n_must ? DENSE_MIN_MUST : DENSE_MIN;
            const uint32_t mode = (col_insert || insert_postings * (uint64_t)DW >= dmin * (uint64_t)std::max<uint32_t>(nd, 1)) ? MODE_DENSE : MODE_HASH;
            // column-scan class: a pure union with a column insert leaf whose block leaves are all sparse enough
            // to be streamed (one warp per leaf, at most NW of them) and whose column leaves share one
            // fieldnorm field. Other dense plans keep their clause phases (all warps share the decode).
            bool colscan = USE_COLSCAN && mode == MODE_DENSE && (D.flags & QF_PURE_UNION) && col_insert && nbl <= (size_t)NW &&
                           q.k <= 128 && STREAM_MAX_BPW;
            for (int i = 0; colscan && i < n_ctmp; i++) colscan = ctmp[i].fn_field >= 0 && ctmp[i].fn_field == ctmp[0].fn_field;
            for (size_t i = 0; colscan && i < nbl; i++)
                colscan = (uint64_t)dl[ql0 + i].n_blocks * DW <= STREAM_MAX_BPW * (uint64_t)std::max<uint32_t>(nd, 1);
            if (colscan) {
                for (size_t i = 0; i < nbl; i++) dl[ql0 + i].lflags |= LF_STREAM;
                D.n_stream = (uint32_t)nbl;
            }
            // a dense window reads 1 B per doc per column leaf (+ the fieldnorm byte), at a fraction of the
            // per-byte cost of packed blocks; hash rounds only gather the columns at their candidates
            if (mode == MODE_DENSE && n_ctmp) total_bytes += (uint64_t)nd * (uint64_t)(n_ctmp + 1) / (colscan ? COL_COST_DIV : COL_COST_DIV_PHASES);
            // a dense window of the phase kernels has a fixed cost (barriers, skip scans, slot scan) whatever it decodes
            if (mode == MODE_DENSE && !colscan) total_bytes += (uint64_t)(nd / DW + 1) * WINDOW_COST;
            const uint64_t ib = !(D.flags & QF_PURE_UNION) ? ITEM_BYTES_MASKED : (mode == MODE_DENSE ? ITEM_BYTES : ITEM_BYTES_HASH);
            uint64_t want = std::max<uint64_t>(1, (total_bytes + ib / 2) / ib);
            const uint32_t min_span = mode == MODE_DENSE ? (uint32_t)DW : HASH_MIN_SPAN;
            const uint64_t max_items = std::max<uint64_t>(1, nd / min_span);
            const uint32_t ni = (uint32_t)std::min(want, max_items);
            if (mode == MODE_DENSE && SOLO_MIN_BLOCKS != 0xFFFFFFFFull) {
                // long insert lists get a phase of their own: a slot is then touched by one thread only
                for (size_t i = ql0; i < dl.size(); i++)
                    if (dl[i].role == ROLE_INSERT && dl[i].n_blocks >= SOLO_MIN_BLOCKS) dl[i].solo = 1;
            }
            D.n_items = ni;
            n_colscan_items += colscan ? ni : 0;
            if (getenv("FG_DEBUG_PLAN") && !colscan && mode == MODE_DENSE)
                fprintf(stderr, "[plan] q%u dense %s: %zu block leaves (%u streamed), %d column leaves, col_insert %d, k %u, items %u, bytes %llu\n", qi,
                        (D.flags & QF_PURE_UNION) ? "pure" : "masked", nbl, D.n_stream, n_ctmp, (int)col_insert, q.k, ni, (unsigned long long)total_bytes);
            for (uint32_t j = 0; j < ni; j++) {
                DevItem it{};
                it.query = qi;
                // 16-aligned cuts: dense windows read columns / fieldnorms of 8 docs per 64-bit load
                it.doc_lo = (uint32_t)((uint64_t)nd * j / ni) & ~15u;
                it.doc_hi = j + 1 == ni ? nd : ((uint32_t)((uint64_t)nd * (j + 1) / ni) & ~15u);
                it.mode = mode;
                it.slot = D.item_begin + j;
                it.cls = colscan ? 4u : (mode == MODE_DENSE ? 0u : 2u) + ((D.flags & QF_PURE_UNION) ? 0u : 1u);
                items.push_back(it);
                item_cost.push_back(total_bytes / ni);
            }
        }
        return FG_OK;
    };
    const int LT = (int)std::max<uint64_t>(1, std::min<uint64_t>({(uint64_t)HostPool::get().size(), 8, (uint64_t)qb->n_queries / 256 + 1}));
    std::vector<LowerOut> parts((size_t)LT);
    HostPool::get().run(LT, [&](int t) {
        lower_range((uint32_t)((uint64_t)qb->n_queries * t / LT), (uint32_t)((uint64_t)qb->n_queries * (t + 1) / LT), parts[t]);
    });
    for (auto& o : parts)
        if (o.rc != FG_OK) return fail(o.rc, "%s", o.err.c_str());
    std::vector<DevLeaf> dl;
    std::vector<DevItem> items;
    std::vector<uint64_t> item_cost;
    uint32_t kmax = 1;
    uint64_t sum_k = 0, n_colscan_items = 0;
    if (LT == 1) {
        dl.swap(parts[0].dl); items.swap(parts[0].items); item_cost.swap(parts[0].item_cost);
        kmax = parts[0].kmax; sum_k = parts[0].sum_k; n_colscan_items = parts[0].n_colscan_items;
    } else {
        size_t nl_tot = 0, ni_tot = 0;
        for (auto& o : parts) { nl_tot += o.dl.size(); ni_tot += o.items.size(); }
        dl.reserve(nl_tot); items.reserve(ni_tot); item_cost.reserve(ni_tot);
        for (int t = 0; t < LT; t++) {
            LowerOut& o = parts[t];
            const uint32_t lb = (uint32_t)dl.size(), ib = (uint32_t)items.size();
            const uint32_t q0 = (uint32_t)((uint64_t)qb->n_queries * t / LT), q1 = (uint32_t)((uint64_t)qb->n_queries * (t + 1) / LT);
            for (uint32_t qi = q0; qi < q1; qi++) { dq[qi].leaf_begin += lb; dq[qi].item_begin += ib; }
            for (auto& it : o.items) it.slot += ib;
            dl.insert(dl.end(), o.dl.begin(), o.dl.end());
            items.insert(items.end(), o.items.begin(), o.items.end());
            item_cost.insert(item_cost.end(), o.item_cost.begin(), o.item_cost.end());
            kmax = std::max(kmax, o.kmax); sum_k += o.sum_k; n_colscan_items += o.n_colscan_items;
        }
    }

    if (kmax > 128 && n_colscan_items)  // the column-scan kernel has no 32-row queue variant
        for (auto& it : items) if (it.cls == 4u) it.cls = 0u;
    const double t_lower = now_ms();
    // one launch per kernel class, heavy items first inside a class (the hardware CTA scheduler is
    // the work queue): counting sort on (class, 64 half-octave cost buckets), O(n)
    std::vector<DevItem> sorted(items.size());
    {
        auto bucket = [&](size_t i) -> uint32_t {
            const uint64_t c = item_cost[i] >> 8;  // 256-byte granules
            const uint32_t lg = c ? 63u - (uint32_t)__builtin_clzll(c) : 0u;
            const uint32_t sub = (lg && ((c >> (lg - 1)) & 1u)) ? 1u : 0u;
            const uint32_t bk = std::min<uint32_t>(63u, lg * 2u + sub);
            // Column-scan class: the FIRST work item of every query is queued ahead of all sibling items. An
            // item's first window runs cold (the query's threshold is not known yet: fully scanned, no
            // sparse-hit gating); siblings that start after a first item has published its threshold through
            // qtheta are gated from their first window on. Items of one query have equal cost, so plain
            // heavy-first order would start them all at once, all cold.
            uint32_t group = items[i].cls;
            if (group == 4u && items[i].slot != dq[items[i].query].item_begin) group = (uint32_t)NCLS;
            return group * 64u + (63u - bk);
        };
        uint32_t hist[(NCLS + 1) * 64 + 1] = {0};
        for (size_t i = 0; i < items.size(); i++) hist[bucket(i) + 1]++;
        for (int i = 0; i < (NCLS + 1) * 64; i++) hist[i + 1] += hist[i];
        for (size_t i = 0; i < items.size(); i++) sorted[hist[bucket(i)]++] = items[i];
    }
    uint32_t class_count[NCLS] = {};
    for (auto& it : items) class_count[it.cls]++;

    fg_ctx* ctx = ix->ctx;
    CU(cudaSetDevice(ctx->device));
    std::unique_ptr<fg_batch, void (*)(fg_batch*)> b(new fg_batch(), fg_batch_release);
    b->ix = ix;
    b->n_queries = qb->n_queries;
    b->n_items = (uint32_t)items.size();
    b->kcap = kmax;
    b->ks = kmax <= 32 ? 1 : kmax <= 128 ? 4 : 32;
    b->sum_k = sum_k;
    for (int i = 0; i < NCLS; i++) b->class_count[i] = class_count[i];
    auto up = [&](const void* src, size_t bytes, void** dst, size_t* sz) -> int32_t {
        *sz = std::max<size_t>(bytes, 16);
        CU(pool_alloc(ctx, dst, *sz));
        if (bytes) CU(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->up));
        return FG_OK;
    };
    int32_t rc;
    std::lock_guard<std::mutex> g(ctx->mu);
    if ((rc = up(dq.data(), dq.size() * sizeof(DevQuery), (void**)&b->d_queries, &b->sz[0]))) return rc;
    if ((rc = up(dl.data(), dl.size() * sizeof(DevLeaf), (void**)&b->d_leaves, &b->sz[1]))) return rc;
    if ((rc = up(sorted.data(), sorted.size() * sizeof(DevItem), (void**)&b->d_items, &b->sz[2]))) return rc;
    b->sz[3] = std::max<size_t>((size_t)b->n_items * b->kcap * 8, 16);
    b->sz[4] = std::max<size_t>((size_t)b->n_items * 4, 16);
    b->sz[5] = 16 * sizeof(unsigned long long);
    b->sz[6] = std::max<size_t>((size_t)b->n_queries * 4, 16);
    CU(pool_alloc(ctx, (void**)&b->d_partial, b->sz[3]));
    CU(pool_alloc(ctx, (void**)&b->d_partial_count, b->sz[4]));
    CU(pool_alloc(ctx, (void**)&b->d_stats, b->sz[5]));
    CU(pool_alloc(ctx, (void**)&b->d_qtheta, b->sz[6]));
    for (auto& e : b->ev) CU(cudaEventCreate(&e));
    CU(cudaEventCreateWithFlags(&b->ev_up, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&b->ev_done, cudaEventDisableTiming));
    CU(cudaEventRecord(b->ev_up, ctx->up));
    CU(cudaStreamSynchronize(ctx->up));  // host vectors go out of scope (the compute stream is not touched)
    if (ctx->env_timing)
        fprintf(stderr, "[fg_batch_prepare] lowering %.2f ms, sort+upload %.2f ms (%zu items)\n", t_lower - t_begin, now_ms() - t_lower, items.size());
    *out = b.release();
    return FG_OK;
}

extern "C" int32_t fg_batch_execute(fg_batch* b, uint32_t flags, uint32_t k_stride, void* d_hits,
                                    void* d_n_hits, void* d_match_count, void* d_match_bitmap) {
    if (!b || !d_hits || !d_n_hits) return fail(FG_ERR_INVALID, "fg_batch_execute: NULL argument");
    if (!b->combine_k && k_stride < b->kcap) return fail(FG_ERR_INVALID, "k_stride %u < max k %u", k_stride, b->kcap);
    fg_index* ix = b->ix;
    fg_ctx* ctx = ix->ctx;
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaStream_t st = b->exec_stream ? b->exec_stream : ctx->stream;
    CU(cudaStreamWaitEvent(st, b->ev_up, 0));
    CU(cudaMemsetAsync(b->d_stats, 0, (b->lead ? 32 : 16) * sizeof(unsigned long long), st));
    if (b->lead) {
        if (flags & FG_EXEC_EXACT_ACCOUNTING)
            return fail(FG_ERR_INVALID, "FG_EXEC_EXACT_ACCOUNTING needs a batch prepared with FG_PREP_LEGACY | FG_PREP_NO_COLUMNS");
        LeadParams p{};
        p.ix = ix->dev;
        p.queries = b->l_queries;
        p.leaves = b->l_leaves;
        p.items = b->l_items;
        p.n_items = b->n_items;
        p.n_queries = b->n_queries;
        p.work = b->l_state;
        p.cursors = b->l_state + 1;
        p.n_cursors = b->n_cursors;
        p.qtheta = p.cursors + b->n_cursors;
        p.qcount = p.qtheta + b->n_queries;
        p.qmatch = p.qcount + b->n_queries;
        p.qhist = b->l_state + (((size_t)1 + b->n_cursors + 3 * (size_t)b->n_queries + 3) & ~(size_t)3);  // 16-byte aligned
        p.exhaustive = (d_match_count || d_match_bitmap || (flags & FG_EXEC_NO_PRUNE) || ctx->env_no_prune) ? 1 : 0;
        if (!p.exhaustive) CU(cudaMemsetAsync(p.qhist, 0, (size_t)b->n_queries * LHIST_B * 4, st));
        p.partial = b->d_partial;
        p.stats = b->d_stats;
        p.match_bitmap = (uint32_t*)d_match_bitmap;
        p.bitmap_words = (ix->n_docs + 31) / 32;
        p.want_counts = d_match_count ? 1 : 0;
        p.exhaustive = (d_match_count || d_match_bitmap || (flags & FG_EXEC_NO_PRUNE) || ctx->env_no_prune) ? 1 : 0;
        p.acct = (flags & FG_EXEC_COUNTERS) ? 1 : 0;
        p.tma = ctx->lead_tma;
        p.chunk = ctx->lead_chunk;
        p.chunk_req = ctx->lead_chunk_req;
        p.union_work = ctx->lead_union_work;
        p.prof = (ctx->env_prof && p.acct) ? 1u : 0u;
        CU(cudaEventRecord(b->ev[0], st));
        launch_lead(p, b->ks, ctx->n_sms, st);
        CU(cudaEventRecord(b->ev[1], st));
        LeadMergeParams m{};
        m.queries = b->l_queries;
        m.n_queries = b->n_queries;
        m.partial = b->d_partial;
        m.qcount = p.qcount;
        m.qmatch = p.qmatch;
        m.k_stride = k_stride;
        m.doc_base = ix->doc_base;
        m.alive = ix->dev.alive;
        m.n_docs = ix->n_docs;
        m.n_alive = ix->dev.n_alive;
        m.out_hits = d_hits;
        m.out_n = (uint32_t*)d_n_hits;
        m.out_count = (uint32_t*)d_match_count;
        m.sel = b->d_sel;
        if (b->combine_k) launch_lead_combine(m, b->combine_k, b->d_comb, b->d_comb + b->comb_cap2, b->comb_cap2, b->combine_filters, st);
        else launch_lead_merge(m, b->ks, st);
        CU(cudaEventRecord(b->ev[2], st));
        b->n_launches = b->n_queries ? (b->n_items ? 3 : 2) : 0;
        CU(cudaGetLastError());
        return FG_OK;
    }
    CU(cudaMemsetAsync(b->d_qtheta, 0, std::max<size_t>((size_t)b->n_queries * 4, 16), st));
    SearchParams p{};
    p.ix = ix->dev;
    p.queries = b->d_queries;
    p.leaves = b->d_leaves;
    p.items = b->d_items;
    p.n_items = b->n_items;
    p.kcap = b->kcap;
    p.partial = b->d_partial;
    p.partial_count = b->d_partial_count;
    p.stats = b->d_stats;
    p.match_bitmap = (uint32_t*)d_match_bitmap;
    p.bitmap_words = (ix->n_docs + 31) / 32;
    p.exact_filter = (flags & FG_EXEC_EXACT_ACCOUNTING) ? 1 : 0;
    p.deterministic = (flags & FG_EXEC_DETERMINISTIC) ? 1 : 0;
    p.want_counts = d_match_count ? 1 : 0;
    p.acct = (flags & (FG_EXEC_EXACT_ACCOUNTING | FG_EXEC_COUNTERS)) ? 1 : 0;
    p.qtheta = (flags & FG_EXEC_DETERMINISTIC) ? nullptr : b->d_qtheta;
    p.no_prune = ((flags & FG_EXEC_NO_PRUNE) || ctx->env_no_prune) ? 1 : 0;
    p.prof = ctx->env_prof ? b->d_stats + 8 : nullptr;
    CU(cudaEventRecord(b->ev[0], st));
    {
        // fork: class 0 on the main stream, classes 1..3 on side streams, join before the merge
        void* streams[NCLS] = {st, ctx->aux[0], ctx->aux[1], ctx->aux[2], ctx->aux[3]};
        CU(cudaEventRecord(ctx->fork_ev, st));
        for (int i = 0; i < NCLS - 1; i++)
            if (b->class_count[i + 1]) CU(cudaStreamWaitEvent(ctx->aux[i], ctx->fork_ev, 0));
        launch_search(p, b->ks, b->class_count, streams);
        for (int i = 0; i < NCLS - 1; i++)
            if (b->class_count[i + 1]) {
                CU(cudaEventRecord(ctx->join_ev[i], ctx->aux[i]));
                CU(cudaStreamWaitEvent(st, ctx->join_ev[i], 0));
            }
    }
    CU(cudaEventRecord(b->ev[1], st));
    MergeParams m{};
    m.queries = b->d_queries;
    m.n_queries = b->n_queries;
    m.kcap = b->kcap;
    m.partial = b->d_partial;
    m.partial_count = b->d_partial_count;
    m.k_stride = k_stride;
    m.doc_base = ix->doc_base;
    m.alive = ix->dev.alive;
    m.n_docs = ix->n_docs;
    m.n_alive = ix->dev.n_alive;
    m.out_hits = d_hits;
    m.out_n = (uint32_t*)d_n_hits;
    m.out_count = (uint32_t*)d_match_count;
    launch_merge(m, b->ks, st);
    CU(cudaEventRecord(b->ev[2], st));
    b->n_launches = (b->n_queries ? 1 : 0);
    for (int i = 0; i < NCLS; i++) b->n_launches += b->class_count[i] ? 1 : 0;
    CU(cudaGetLastError());
    return FG_OK;
}

extern "C" int32_t fg_batch_query_status(const fg_batch* b, int32_t* out_status) {
    if (!b || !out_status) return fail(FG_ERR_INVALID, "fg_batch_query_status: NULL argument");
    for (uint32_t i = 0; i < b->n_queries; i++) out_status[i] = i < b->qstatus.size() ? b->qstatus[i] : FG_OK;
    for (int32_t st : b->qstatus)
        if (st != FG_OK) { fail(st, "%s", b->first_bad.c_str()); break; }  // fg_last_error() describes the first offender
    return FG_OK;
}

extern "C" int32_t fg_batch_get_stats(fg_batch* b, fg_batch_stats* out) {
    if (!b || !out) return fail(FG_ERR_INVALID, "NULL argument");
    fg_ctx* ctx = b->ix->ctx;
    CU(cudaSetDevice(ctx->device));
    unsigned long long h[8];
    {
        std::lock_guard<std::mutex> g(ctx->mu);
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaMemcpy(h, b->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
    }
    if (ctx->env_prof && b->lead) {
        unsigned long long pr[8], ph[10];
        cudaMemcpy(pr, b->d_stats + 8, sizeof(pr), cudaMemcpyDeviceToHost);
        cudaMemcpy(ph, b->d_stats + 16, sizeof(ph), cudaMemcpyDeviceToHost);
        unsigned long long ph_tot = 0;
        for (int i = 0; i < 10; i++) ph_tot += ph[i];
        if (ph_tot)  // (only a library built with make PROFILE=1 fills these)
            fprintf(stderr, "[prof lead phases] %% of warp busy time: plan load %.1f, block walk %.1f, decode %.1f, lookups %.1f, scoring %.1f, "
                            "top-k + histogram %.1f, append %.1f, other %.1f; lookups by kind: tf column %.1f, bitmap %.1f, gallop + block %.1f\n", 100.0 * ph[0] / ph_tot, 100.0 * ph[1] / ph_tot, 100.0 * ph[2] / ph_tot,
                    100.0 * (ph[3] + ph[8] + ph[9]) / ph_tot, 100.0 * ph[4] / ph_tot, 100.0 * ph[5] / ph_tot, 100.0 * ph[6] / ph_tot, 100.0 * ph[7] / ph_tot,
                    100.0 * ph[3] / ph_tot, 100.0 * ph[8] / ph_tot, 100.0 * ph[9] / ph_tot);
        if (pr[4]) {
            const double start = (double)~pr[5], first_idle = (double)~pr[2] - start, done = (double)pr[3] - start;
            fprintf(stderr, "[prof lead] warps %llu: kernel %.1f us, first warp out of work at %.1f us, busy share %.1f %% of warp-time, longest item %.1f us, "
                            "mean busy per warp %.1f us\n", pr[4], done * 1e-3, first_idle * 1e-3, 100.0 * (double)pr[0] / ((double)pr[4] * done),
                    (double)pr[1] * 1e-3, (double)pr[0] / (double)pr[4] * 1e-3);
        }
    }
    if (ctx->env_prof && !b->lead) {
        unsigned long long pr[8];
        cudaMemcpy(pr, b->d_stats + 8, sizeof(pr), cudaMemcpyDeviceToHost);
        const double ni = b->n_items ? (double)b->n_items : 1.0;
        fprintf(stderr, "[prof] per item (cycles): init %.0f setup %.0f scan %.0f decode %.0f slotscan %.0f epilogue %.0f rounds %.1f\n",
                pr[0] / ni, pr[1] / ni, pr[2] / ni, pr[3] / ni, pr[4] / ni, pr[5] / ni, pr[7] / ni);
    }
    out->bytes_blocks = h[0];
    out->bytes_redecode = h[1];
    out->scored_postings = h[2];
    out->colscan_chunks = h[3];
    out->colscan_chunks_skipped = h[4];
    out->bytes_meta = b->lead ? h[5] : 0;
    out->lead_blocks = b->lead ? h[6] : 0;
    out->lead_blocks_seen = b->lead ? h[7] : 0;
    out->n_work_items = b->n_items;
    out->plan_bytes = b->plan_sz;
    out->n_launches = b->n_launches;
    out->n_queries = b->n_queries;
    out->sum_k = b->sum_k;
    out->search_kernel_ms = out->merge_kernel_ms = 0.f;
    if (b->n_launches) {
        CU(cudaEventElapsedTime(&out->search_kernel_ms, b->ev[0], b->ev[1]));
        CU(cudaEventElapsedTime(&out->merge_kernel_ms, b->ev[1], b->ev[2]));
    }
    return FG_OK;
}

extern "C" int32_t fg_search_batch(fg_index* ix, const fg_query_batch* qb, uint32_t k_stride,
                                   fg_hit* out_hits, uint32_t* out_n_hits, uint32_t* out_match_count) {
    if (!ix || !qb || !out_hits || !out_n_hits) return fail(FG_ERR_INVALID, "fg_search_batch: NULL argument");
    const bool timing = ix->ctx->env_timing;
    const double t0 = now_ms();
    fg_batch* b = nullptr;
    // match counts need every matching document visited: the windowed accumulator kernels are the exhaustive engine;
    // the TopDocs form (no counts) runs on the lead-driven kernels, which prune
    int32_t rc = out_match_count ? fg_batch_prepare_ex(ix, qb, FG_PREP_LEGACY, &b) : FG_ERR_UNSUPPORTED;
    if (rc) rc = fg_batch_prepare(ix, qb, &b);
    const double t1 = now_ms();
    if (rc) return rc;
    std::unique_ptr<fg_batch, void (*)(fg_batch*)> guard(b, fg_batch_release);
    if (k_stride < b->kcap) return fail(FG_ERR_INVALID, "k_stride %u < max k %u", k_stride, b->kcap);
    fg_ctx* ctx = ix->ctx;
    const size_t nq = qb->n_queries;
    if (nq == 0) return FG_OK;
    void *d_hits = nullptr, *d_n = nullptr, *d_c = nullptr;
    struct Guard {
        fg_ctx* c; void* p = nullptr; size_t sz = 0;
        ~Guard() { if (p) { cudaStreamSynchronize(c->stream); pool_free(c, p, sz); } }
    } g1{ctx}, g2{ctx}, g3{ctx};
    CU(pool_alloc(ctx, &d_hits, nq * k_stride * sizeof(fg_hit)));
    g1.p = d_hits; g1.sz = nq * k_stride * sizeof(fg_hit);
    CU(pool_alloc(ctx, &d_n, nq * 4));
    g2.p = d_n; g2.sz = nq * 4;
    if (out_match_count) {  // optional: the reference's TopDocs collector does not count matches
        CU(pool_alloc(ctx, &d_c, nq * 4));
        g3.p = d_c; g3.sz = nq * 4;
    }
    rc = fg_batch_execute(b, 0, k_stride, d_hits, d_n, d_c, nullptr);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaMemcpyAsync(out_hits, d_hits, nq * k_stride * sizeof(fg_hit), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out_n_hits, d_n, nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_match_count) CU(cudaMemcpyAsync(out_match_count, d_c, nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (timing)
        fprintf(stderr, "[fg_search_batch] prepare %.2f ms, alloc+execute+d2h %.2f ms (n=%zu)\n", t1 - t0, now_ms() - t1, nq);
    return FG_OK;
}

// One query whose Should children are boolean queries themselves (tantivy: BooleanQuery of BooleanQuerys, e.g.
// `(a AND b) OR (c AND d)`): the batch's queries are its disjuncts. Every disjunct is evaluated exhaustively by the
// lead-driven kernels in the deep-page form (all matches kept), lead_combine_kernel sums the scores per document and
// selects the page. The disjuncts' own k is ignored.
extern "C" int32_t fg_search_union_of(fg_index* ix, const fg_query_batch* disjuncts, uint32_t k, fg_hit* out_hits, uint32_t* out_n_hits,
                                      uint32_t* out_match_count) {
    return fg_search_union_of_filtered(ix, disjuncts, 0, k, out_hits, out_n_hits, out_match_count);
}

// The same with FILTER children: the last n_filters queries of the batch. A document matches when at least one ordinary
// child and every filter child match it; its score is the sum over the ordinary children plus the filter scores (tantivy:
// Bool[Must(union of the children), Must(filter)..], src/db/search.rs:140-144 with a nested text query). Scores must not
// be negative (the combine step tags filter entries in the sign position of the sortable score).
extern "C" int32_t fg_search_union_of_filtered(fg_index* ix, const fg_query_batch* disjuncts, uint32_t n_filters, uint32_t k, fg_hit* out_hits,
                                               uint32_t* out_n_hits, uint32_t* out_match_count) {
    if (!ix || !disjuncts || !out_hits || !out_n_hits) return fail(FG_ERR_INVALID, "fg_search_union_of: NULL argument");
    if (k == 0) return fail(FG_ERR_INVALID, "k == 0 (TopDocs::with_limit requires limit >= 1)");
    if (disjuncts->n_queries == 0 || disjuncts->n_queries > 64) return fail(FG_ERR_INVALID, "fg_search_union_of: 1 to 64 disjuncts");
    if (n_filters >= disjuncts->n_queries) return fail(FG_ERR_INVALID, "fg_search_union_of_filtered: at least one child must not be a filter");
    if (n_filters)
        for (uint32_t i = 0; i < disjuncts->n_leaves; i++)
            if (disjuncts->leaves[i].boost < 0.f) return fail(FG_ERR_UNSUPPORTED, "fg_search_union_of_filtered: negative boosts");
    std::vector<fg_query> qs(disjuncts->queries, disjuncts->queries + disjuncts->n_queries);
    for (fg_query& q : qs) q.k = 0x7FFFFFFFu;  // every match of every disjunct: no per-warp queue, no threshold
    fg_query_batch qb = *disjuncts;
    qb.queries = qs.data();
    fg_batch* b = nullptr;
    int32_t rc = prepare_lead(ix, &qb, 0, &b);
    if (rc) return rc;
    std::unique_ptr<fg_batch, void (*)(fg_batch*)> guard(b, fg_batch_release);
    fg_ctx* ctx = ix->ctx;
    CU(cudaSetDevice(ctx->device));
    uint64_t cap2 = 1;
    while (cap2 < std::max<uint64_t>(b->partial_entries, 1)) cap2 <<= 1;
    if (cap2 > 0x40000000ull) return fail(FG_ERR_UNSUPPORTED, "fg_search_union_of: the disjuncts match too many documents");
    b->combine_k = k;
    b->combine_filters = n_filters;
    b->comb_cap2 = (uint32_t)cap2;
    b->comb_sz = (size_t)cap2 * 16;
    CU(pool_alloc(ctx, (void**)&b->d_comb, b->comb_sz));
    const uint32_t k_stride = (uint32_t)std::min<uint64_t>(k, std::max<uint64_t>(b->partial_entries, 1));
    void *d_hits = nullptr, *d_n = nullptr;
    struct Guard {
        fg_ctx* c; void* p = nullptr; size_t sz = 0;
        ~Guard() { if (p) { cudaStreamSynchronize(c->stream); pool_free(c, p, sz); } }
    } g1{ctx}, g2{ctx};
    CU(pool_alloc(ctx, &d_hits, (size_t)k_stride * sizeof(fg_hit)));
    g1.p = d_hits; g1.sz = (size_t)k_stride * sizeof(fg_hit);
    CU(pool_alloc(ctx, &d_n, 2 * sizeof(uint32_t)));
    g2.p = d_n; g2.sz = 2 * sizeof(uint32_t);
    // (match counts come from the combine step itself: distinct documents over all disjuncts)
    rc = fg_batch_execute(b, 0, k_stride, d_hits, d_n, out_match_count ? (char*)d_n + 4 : nullptr, nullptr);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(ctx->mu);
    uint32_t h_n[2] = {0, 0};
    CU(cudaMemcpyAsync(h_n, d_n, sizeof(h_n), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out_hits, d_hits, (size_t)k_stride * sizeof(fg_hit), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *out_n_hits = h_n[0];
    if (out_match_count) *out_match_count = h_n[1];
    return FG_OK;
}

extern "C" int32_t fg_batch_submit(fg_batch* b, uint32_t flags, uint32_t k_stride, int32_t want_counts) {
    if (!b) return fail(FG_ERR_INVALID, "fg_batch_submit: NULL batch");
    if (k_stride < b->kcap) return fail(FG_ERR_INVALID, "k_stride %u < max k %u", k_stride, b->kcap);
    if (b->d_out) return fail(FG_ERR_INVALID, "fg_batch_submit: batch already submitted");
    fg_ctx* ctx = b->ix->ctx;
    CU(cudaSetDevice(ctx->device));
    const size_t nq = b->n_queries;
    b->sub_k_stride = k_stride;
    b->sub_counts = want_counts != 0;
    if (nq == 0) return FG_OK;
    const size_t hits_b = nq * k_stride * sizeof(fg_hit);
    b->out_sz = hits_b + 2 * nq * 4;
    CU(pool_alloc(ctx, &b->d_out, b->out_sz));
    CU(pinned_alloc(ctx, &b->h_out, b->out_sz));
    char* d = (char*)b->d_out;
    if (b->lead && ctx->sub_streams) {  // (the window kernels share per-context side streams: they stay on the context's stream)
        std::lock_guard<std::mutex> g(ctx->mu);
        b->exec_stream = ctx->sub[ctx->sub_seq++ % ctx->sub_streams];
    }
    int32_t rc = fg_batch_execute(b, flags, k_stride, d, d + hits_b, want_counts ? d + hits_b + nq * 4 : nullptr, nullptr);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaStream_t st = b->exec_stream ? b->exec_stream : ctx->stream;
    CU(cudaMemcpyAsync(b->h_out, b->d_out, b->out_sz, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(b->ev_done, st));
    return FG_OK;
}

// fg_batch_submit for a sharded index: local top-k -> all-gather -> merge inside the library, then the device->host copy
// of the GLOBAL result into the batch's staging; fg_batch_collect hands it out. Collective.
extern "C" int32_t fg_batch_submit_sharded(fg_batch* b, fg_comm* comm, uint32_t flags, uint32_t k_stride) {
    if (!b || !comm) return fail(FG_ERR_INVALID, "fg_batch_submit_sharded: NULL argument");
    if (k_stride < b->kcap) return fail(FG_ERR_INVALID, "k_stride %u < max k %u", k_stride, b->kcap);
    if (b->d_out) return fail(FG_ERR_INVALID, "fg_batch_submit_sharded: batch already submitted");
    fg_ctx* ctx = b->ix->ctx;
    CU(cudaSetDevice(ctx->device));
    const size_t nq = b->n_queries;
    b->sub_k_stride = k_stride;
    b->sub_counts = false;
    if (nq == 0) return FG_OK;
    const size_t hits_b = nq * k_stride * sizeof(fg_hit);
    b->out_sz = hits_b + 2 * nq * 4;
    CU(pool_alloc(ctx, &b->d_out, b->out_sz));
    CU(pinned_alloc(ctx, &b->h_out, b->out_sz));
    char* d = (char*)b->d_out;
    if (b->lead && ctx->sub_streams) {  // as fg_batch_submit: consecutive chunks overlap (every rank submits in the same order)
        std::lock_guard<std::mutex> g(ctx->mu);
        b->exec_stream = ctx->sub[ctx->sub_seq++ % ctx->sub_streams];
    }
    int32_t rc = fg_batch_execute_sharded(b, comm, flags, k_stride, d, d + hits_b);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaStream_t st = b->exec_stream ? b->exec_stream : ctx->stream;
    CU(cudaMemcpyAsync(b->h_out, b->d_out, hits_b + nq * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(b->ev_done, st));
    return FG_OK;
}

extern "C" int32_t fg_batch_collect(fg_batch* b, fg_hit* out_hits, uint32_t* out_n_hits, uint32_t* out_match_count) {
    if (!b || !out_hits || !out_n_hits) return fail(FG_ERR_INVALID, "fg_batch_collect: NULL argument");
    const size_t nq = b->n_queries;
    if (nq == 0) return FG_OK;
    if (!b->h_out) return fail(FG_ERR_INVALID, "fg_batch_collect: batch was not submitted");
    if (out_match_count && !b->sub_counts) return fail(FG_ERR_INVALID, "fg_batch_collect: the batch was submitted without match counts");
    CU(cudaSetDevice(b->ix->ctx->device));
    CU(cudaEventSynchronize(b->ev_done));
    const size_t hits_b = nq * b->sub_k_stride * sizeof(fg_hit);
    const char* h = (const char*)b->h_out;
    memcpy(out_hits, h, hits_b);
    memcpy(out_n_hits, h + hits_b, nq * 4);
    if (out_match_count) memcpy(out_match_count, h + hits_b + nq * 4, nq * 4);
    return FG_OK;
}

extern "C" int32_t fg_merge_topk_device(fg_ctx* ctx, const void* d_hits, const void* d_n,
                                        uint32_t n_ranks, uint32_t n_queries, uint32_t k,
                                        uint32_t k_stride, void* d_out_hits, void* d_out_n) {
    if (!ctx || !d_hits || !d_n || !d_out_hits || !d_out_n) return fail(FG_ERR_INVALID, "NULL argument");
    if (k == 0 || k > 1024 || k > k_stride) return fail(FG_ERR_INVALID, "k must be in [1, min(1024, k_stride)]");
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(ctx->mu);
    launch_merge_gathered(d_hits, (const uint32_t*)d_n, n_ranks, n_queries, k, k_stride, d_out_hits,
                          (uint32_t*)d_out_n, k <= 32 ? 1 : k <= 128 ? 4 : 32, ctx->stream);
    CU(cudaGetLastError());
    return FG_OK;
}

// ------------------------------------------------------------------------------------------
// multi-GPU, one process per GPU: NCCL communicator + sharded execution
// ------------------------------------------------------------------------------------------
namespace {
// NCCL is loaded on first use (dlopen): a single-GPU host never needs the library, and a host that already
// carries an NCCL (torch's bundled copy has the same soname) shares it instead of getting a second one.
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};
NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (api.h) break;
        }
        if (!api.h) { api.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
        auto sym = [&](const char* n) -> void* {
            void* p = dlsym(api.h, n);
            if (!p && api.err.empty()) api.err = std::string("libnccl lacks ") + n;
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    });
    return api;
}
}  // namespace
#define NC(call)                                                                                          \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess)                                                                            \
            return fail(FG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, nccl_api().GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

struct fg_comm {
    fg_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    // exchange buffers, one set per stream a sharded batch may run on (0 = the context's stream, 1 / 2 = the submit streams):
    void* d_local[3] = {nullptr, nullptr, nullptr};    // this shard's [hits n*k_stride | n_hits n]
    void* d_gather[3] = {nullptr, nullptr, nullptr};   // [world][hits] then [world][n_hits]
    size_t local_sz[3] = {0, 0, 0};
    void* d_tmp = nullptr;      // all-reduce staging
    size_t tmp_sz = 0;
    void* h_tmp = nullptr;      // page-locked staging of fg_comm_allgather_bytes ([send | recv x world])
    size_t h_tmp_sz = 0;
};

extern "C" int32_t fg_comm_unique_id(void* out_id) {
    if (!out_id) return fail(FG_ERR_INVALID, "fg_comm_unique_id: NULL argument");
    NcclApi& n = nccl_api();
    if (!n.err.empty()) return fail(FG_ERR_UNSUPPORTED, "%s", n.err.c_str());
    static_assert(sizeof(ncclUniqueId) == FG_COMM_ID_BYTES, "FG_COMM_ID_BYTES");
    ncclUniqueId id;
    NC(n.GetUniqueId(&id));
    memcpy(out_id, &id, sizeof(id));
    return FG_OK;
}
extern "C" int32_t fg_comm_create(fg_ctx* ctx, int32_t rank, int32_t n_ranks, const void* id, fg_comm** out) {
    if (!ctx || !id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(FG_ERR_INVALID, "fg_comm_create: bad argument");
    *out = nullptr;
    NcclApi& n = nccl_api();
    if (!n.err.empty()) return fail(FG_ERR_UNSUPPORTED, "%s", n.err.c_str());
    CU(cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    std::unique_ptr<fg_comm> c(new fg_comm());
    c->ctx = ctx;
    c->rank = rank;
    c->world = n_ranks;
    NC(n.CommInitRank(&c->comm, n_ranks, uid, rank));
    *out = c.release();
    return FG_OK;
}
extern "C" void fg_comm_destroy(fg_comm* c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->comm) nccl_api().CommDestroy(c->comm);
    for (int i = 0; i < 3; i++) {
        if (i && c->ctx->sub[i - 1]) cudaStreamSynchronize(c->ctx->sub[i - 1]);
        cudaFree(c->d_local[i]);
        cudaFree(c->d_gather[i]);
    }
    cudaFree(c->d_tmp);
    if (c->h_tmp) cudaFreeHost(c->h_tmp);
    delete c;
}
extern "C" int32_t fg_comm_info(const fg_comm* c, int32_t* rank, int32_t* n_ranks) {
    if (!c) return fail(FG_ERR_INVALID, "fg_comm_info: NULL communicator");
    if (rank) *rank = c->rank;
    if (n_ranks) *n_ranks = c->world;
    return FG_OK;
}
// all-gather of `bytes` bytes per rank, HOST buffers (staged through the device like the all-reduce); collective
extern "C" int32_t fg_comm_allgather_bytes(fg_comm* c, const void* send, size_t bytes, void* recv) {
    if (!c || (bytes && (!send || !recv))) return fail(FG_ERR_INVALID, "fg_comm_allgather_bytes: NULL argument");
    if (bytes == 0) return FG_OK;
    fg_ctx* ctx = c->ctx;
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(ctx->mu);
    const size_t need = bytes * ((size_t)c->world + 1);
    if (c->tmp_sz < need) {
        CU(cudaStreamSynchronize(ctx->stream));
        cudaFree(c->d_tmp);
        c->d_tmp = nullptr;
        c->tmp_sz = 0;
        CU(cudaMalloc(&c->d_tmp, need + (need >> 1)));
        c->tmp_sz = need + (need >> 1);
    }
    if (c->h_tmp_sz < need) {  // page-locked staging: the copies run at bus speed and truly asynchronously
        if (c->h_tmp) cudaFreeHost(c->h_tmp);
        c->h_tmp = nullptr;
        c->h_tmp_sz = 0;
        CU(cudaHostAlloc(&c->h_tmp, need + (need >> 1), cudaHostAllocDefault));
        c->h_tmp_sz = need + (need >> 1);
    }
    char* d_send = (char*)c->d_tmp;
    char* d_recv = d_send + bytes;
    char* h_send = (char*)c->h_tmp;
    char* h_recv = h_send + bytes;
    memcpy(h_send, send, bytes);
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(d_send, h_send, bytes, cudaMemcpyHostToDevice, st));
    NC(nccl_api().AllGather(d_send, d_recv, bytes, ncclUint8, c->comm, st));
    CU(cudaMemcpyAsync(h_recv, d_recv, bytes * c->world, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    memcpy(recv, h_recv, bytes * c->world);
    return FG_OK;
}
static int32_t comm_allreduce(fg_comm* c, void* values, size_t n, size_t elem, ncclDataType_t dt) {
    if (!c || (n && !values)) return fail(FG_ERR_INVALID, "fg_comm_allreduce: NULL argument");
    if (n == 0) return FG_OK;
    fg_ctx* ctx = c->ctx;
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(ctx->mu);
    if (c->tmp_sz < n * elem) {
        cudaFree(c->d_tmp);
        c->d_tmp = nullptr;
        c->tmp_sz = 0;
        CU(cudaMalloc(&c->d_tmp, n * elem));
        c->tmp_sz = n * elem;
    }
    CU(cudaMemcpyAsync(c->d_tmp, values, n * elem, cudaMemcpyHostToDevice, ctx->stream));
    NC(nccl_api().AllReduce(c->d_tmp, c->d_tmp, n, dt, ncclSum, c->comm, ctx->stream));
    CU(cudaMemcpyAsync(values, c->d_tmp, n * elem, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return FG_OK;
}
extern "C" int32_t fg_comm_allreduce_sum_u64(fg_comm* c, uint64_t* v, size_t n) { return comm_allreduce(c, v, n, 8, ncclUint64); }
extern "C" int32_t fg_comm_allreduce_sum_u32(fg_comm* c, uint32_t* v, size_t n) { return comm_allreduce(c, v, n, 4, ncclUint32); }

extern "C" int32_t fg_batch_execute_sharded(fg_batch* b, fg_comm* c, uint32_t flags, uint32_t k_stride, void* d_hits, void* d_n_hits) {
    if (!b || !c || !d_hits || !d_n_hits) return fail(FG_ERR_INVALID, "fg_batch_execute_sharded: NULL argument");
    fg_ctx* ctx = b->ix->ctx;
    if (ctx != c->ctx) return fail(FG_ERR_INVALID, "fg_batch_execute_sharded: the batch and the communicator belong to different contexts");
    if (b->lead && b->ks == 0) return fail(FG_ERR_UNSUPPORTED, "fg_batch_execute_sharded: page limits above 1024 are not merged across shards");
    if (k_stride < b->kcap) return fail(FG_ERR_INVALID, "k_stride %u < max k %u", k_stride, b->kcap);
    CU(cudaSetDevice(ctx->device));
    const size_t nq = b->n_queries;
    if (nq == 0) return FG_OK;
    const size_t hits_b = nq * k_stride * sizeof(fg_hit), n_b = nq * 4, local = hits_b + n_b;
    cudaStream_t st = b->exec_stream ? b->exec_stream : ctx->stream;
    const int slot = b->exec_stream == nullptr ? 0 : (b->exec_stream == ctx->sub[0] ? 1 : 2);
    {
        std::lock_guard<std::mutex> g(ctx->mu);
        if (c->local_sz[slot] < local) {
            CU(cudaStreamSynchronize(st));
            cudaFree(c->d_local[slot]);
            cudaFree(c->d_gather[slot]);
            c->d_local[slot] = c->d_gather[slot] = nullptr;
            c->local_sz[slot] = 0;
            CU(cudaMalloc(&c->d_local[slot], local));
            CU(cudaMalloc(&c->d_gather[slot], local * c->world));
            c->local_sz[slot] = local;
        }
    }
    char* dl = (char*)c->d_local[slot];
    char* dg = (char*)c->d_gather[slot];
    int32_t rc = fg_batch_execute(b, flags, k_stride, dl, dl + hits_b, nullptr, nullptr);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(ctx->mu);
    NcclApi& n = nccl_api();
    // one fused NCCL launch: the hit lists and their lengths of every shard
    NC(n.GroupStart());
    NC(n.AllGather(dl, dg, hits_b, ncclUint8, c->comm, st));
    NC(n.AllGather(dl + hits_b, dg + hits_b * c->world, n_b, ncclUint8, c->comm, st));
    NC(n.GroupEnd());
    const void* qrec = b->lead ? (const void*)b->l_queries : (const void*)b->d_queries;
    const uint32_t q_words = b->lead ? sizeof(LQuery) / 4 : sizeof(DevQuery) / 4;
    const uint32_t k_word = b->lead ? offsetof(LQuery, k) / 4 : offsetof(DevQuery, k) / 4;
    launch_merge_ranks(dg, (const uint32_t*)(dg + hits_b * c->world), (uint32_t)c->world, (uint32_t)nq, qrec, q_words, k_word, k_stride,
                       d_hits, (uint32_t*)d_n_hits, b->ks, st);
    CU(cudaEventRecord(b->ev[2], st));
    b->n_launches += 2;
    CU(cudaGetLastError());
    return FG_OK;
}
