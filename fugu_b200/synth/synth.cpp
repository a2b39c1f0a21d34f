// Synthetic Zipfian corpus generator — bench/test input
// infrastructure, NOT part of the query hot path and NOT part of the oracle.
//
// Everything is counter-based: token i of doc d depends only on (seed, d, i), so any
// doc range can be generated independently (doc-ID shards on different ranks see the
// same corpus without materialising it on one host).
//
//   rng(seed, d, i) = splitmix64(splitmix64(seed ^ d*0xD6E8FEB86659FD93) + i)
//   len(d)   = clamp(round(exp(ln 60 + 0.5 z)), 8, 400), z ~ N(0,1) by Box-Muller on counters 0,1
//   tok(d,i) = inverse-CDF Zipf(s, V) sample of counter 2+i          (term string "w<rank>")
//   name(d)  : with probability name_pct/100 a 3..6-token `metadata.name` (counters 2^32+..)
//   facets(d): ns = d mod n_ns -> the ObjectRecord namespace facets of
//              /root/reference/src/object.rs:81-111 with all ancestor paths as terms.
//
// Output is the flat per-field CSR (term -> ascending local doc ids + term freqs) that the
// reference-side index loader would hand to fg_index_upload (include/fugu_gpu.h).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

inline uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline uint64_t rng(uint64_t seed, uint64_t d, uint64_t i) {
    return splitmix64(splitmix64(seed ^ (d * 0xD6E8FEB86659FD93ull)) + i);
}
inline double u01(uint64_t x) { return (double)(x >> 11) * (1.0 / 9007199254740992.0); }

struct Corpus {
    uint64_t seed;
    uint32_t vocab;
    double zipf_s;
    int name_pct;
    uint32_t n_ns;  // 0 = no facets
    std::vector<double> cdf;
    std::vector<uint32_t> guide;  // guide[j] = first rank index whose cdf >= j/G
    static constexpr uint32_t G = 1 << 16;

    void init() {
        cdf.resize(vocab);
        double h = 0;
        for (uint32_t r = 1; r <= vocab; r++) h += std::pow((double)r, -zipf_s);
        double acc = 0;
        for (uint32_t r = 1; r <= vocab; r++) {
            acc += std::pow((double)r, -zipf_s) / h;
            cdf[r - 1] = acc;
        }
        cdf[vocab - 1] = 1.0;
        guide.resize(G + 1);
        uint32_t k = 0;
        for (uint32_t j = 0; j <= G; j++) {
            double t = (double)j / G;
            while (k < vocab - 1 && cdf[k] < t) k++;
            guide[j] = k;
        }
    }
    // smallest index k with cdf[k] > u  (term ordinal = rank-1)
    inline uint32_t sample(double u) const {
        uint32_t j = (uint32_t)(u * G);
        uint32_t lo = guide[j], hi = guide[j + 1];
        while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (cdf[mid] > u) hi = mid; else lo = mid + 1;
        }
        return lo;
    }
    inline uint32_t doc_len(uint64_t d) const {
        double u1 = u01(rng(seed, d, 0)), u2 = u01(rng(seed, d, 1));
        u1 = (u1 * 9007199254740992.0 + 1.0) / 9007199254740993.0;  // (0,1]
        double z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
        double l = std::nearbyint(std::exp(std::log(60.0) + 0.5 * z));
        if (l < 8) l = 8;
        if (l > 400) l = 400;
        return (uint32_t)l;
    }
    inline uint32_t name_len(uint64_t d) const {
        if (name_pct <= 0) return 0;
        const uint64_t B = 1ull << 32;
        if ((int)(rng(seed, d, B) % 100) >= name_pct) return 0;
        return 3 + (uint32_t)(rng(seed, d, B + 1) % 4);
    }
    // tokens of (doc, field): field 0 = text, field 1 = name
    inline uint32_t tokens(uint64_t d, int field, uint32_t* out) const {
        if (field == 0) {
            uint32_t n = doc_len(d);
            for (uint32_t i = 0; i < n; i++) out[i] = sample(u01(rng(seed, d, 2 + i)));
            return n;
        }
        uint32_t n = name_len(d);
        const uint64_t B = 1ull << 32;
        for (uint32_t i = 0; i < n; i++) out[i] = sample(u01(rng(seed, d, B + 2 + i)));
        return n;
    }
};

// ---- facet vocabulary (field 2) ------------------------------------------------------
// ns = d mod n_ns, org = (d / n_ns) mod 16, type = (d / (n_ns*16)) mod 4.
// Term ordinals (ancestors included, each its own term as tantivy indexes facets):
//   0                       /namespace
//   1 + ns*23 + 0           /namespace/nsXX
//   1 + ns*23 + 1           /namespace/nsXX/organization
//   1 + ns*23 + 2 + org     /namespace/nsXX/organization/orgY      (16)
//   1 + ns*23 + 18          /namespace/nsXX/data
//   1 + ns*23 + 19 + type   /namespace/nsXX/data/typeZ             (4)
constexpr uint32_t FACET_PER_NS = 23;

struct Csr {
    uint32_t n_terms = 0;
    uint64_t n_postings = 0;
    uint64_t n_docs = 0;
    uint64_t total_tokens = 0;
    uint64_t* offsets = nullptr;
    uint32_t* docs = nullptr;
    uint32_t* tfs = nullptr;
    uint32_t* doc_len = nullptr;
};

int n_threads() {
    unsigned h = std::thread::hardware_concurrency();
    if (h == 0) h = 4;
    const char* e = getenv("FGS_THREADS");
    if (e) h = (unsigned)atoi(e);
    return (int)std::max(1u, std::min(h, 32u));
}

}  // namespace

extern "C" {

void* fgs_corpus_create(uint64_t seed, uint32_t vocab, double zipf_s, int name_pct, uint32_t n_ns) {
    Corpus* c = new Corpus();
    c->seed = seed; c->vocab = vocab; c->zipf_s = zipf_s; c->name_pct = name_pct; c->n_ns = n_ns;
    c->init();
    return c;
}
void fgs_corpus_destroy(void* h) { delete (Corpus*)h; }

uint32_t fgs_doc_len(void* h, uint64_t d) { return ((Corpus*)h)->doc_len(d); }

// Tokens (term ordinals) of one doc/field into out (capacity >= 400). Returns count.
uint32_t fgs_doc_tokens(void* h, uint64_t d, int field, uint32_t* out) {
    return ((Corpus*)h)->tokens(d, field, out);
}

uint32_t fgs_facet_vocab(void* h) {
    Corpus* c = (Corpus*)h;
    return c->n_ns ? 1 + c->n_ns * FACET_PER_NS : 0;
}
// facet term ordinals of doc d (6 of them); returns count
uint32_t fgs_doc_facets(void* h, uint64_t d, uint32_t* out) {
    Corpus* c = (Corpus*)h;
    if (!c->n_ns) return 0;
    uint32_t ns = (uint32_t)(d % c->n_ns), org = (uint32_t)((d / c->n_ns) % 16),
             ty = (uint32_t)((d / ((uint64_t)c->n_ns * 16)) % 4);
    uint32_t b = 1 + ns * FACET_PER_NS;
    out[0] = 0; out[1] = b; out[2] = b + 1; out[3] = b + 2 + org; out[4] = b + 18; out[5] = b + 19 + ty;
    return 6;
}
// facet path string of a facet term ordinal
int fgs_facet_path(void* h, uint32_t ord, char* buf, int cap) {
    Corpus* c = (Corpus*)h;
    if (!c->n_ns) return -1;
    if (ord == 0) return snprintf(buf, cap, "/namespace");
    uint32_t ns = (ord - 1) / FACET_PER_NS, k = (ord - 1) % FACET_PER_NS;
    if (k == 0) return snprintf(buf, cap, "/namespace/ns%02u", ns);
    if (k == 1) return snprintf(buf, cap, "/namespace/ns%02u/organization", ns);
    if (k < 18) return snprintf(buf, cap, "/namespace/ns%02u/organization/org%u", ns, k - 2);
    if (k == 18) return snprintf(buf, cap, "/namespace/ns%02u/data", ns);
    return snprintf(buf, cap, "/namespace/ns%02u/data/type%u", ns, k - 19);
}

// Build the flat CSR of one field over docs [d0, d1) with LOCAL doc ids (d - d0).
// field: 0 text, 1 name, 2 facet. Caller frees with fgs_csr_free.
void* fgs_build_csr(void* h, uint64_t d0, uint64_t d1, int field) {
    Corpus* c = (Corpus*)h;
    Csr* out = new Csr();
    const uint64_t nd = d1 - d0;
    const uint32_t V = field == 2 ? fgs_facet_vocab(h) : c->vocab;
    out->n_terms = V;
    out->n_docs = nd;
    out->offsets = (uint64_t*)calloc((size_t)V + 1, sizeof(uint64_t));
    out->doc_len = (uint32_t*)calloc((size_t)std::max<uint64_t>(nd, 1), sizeof(uint32_t));
    const int T = (int)std::min<uint64_t>((uint64_t)n_threads(), std::max<uint64_t>(1, nd / 1024 + 1));
    std::vector<std::vector<uint32_t>> df((size_t)T, std::vector<uint32_t>(V, 0));
    std::vector<uint64_t> tot((size_t)T, 0);
    auto range = [&](int t, uint64_t& a, uint64_t& b) {
        a = d0 + nd * (uint64_t)t / T;
        b = d0 + nd * (uint64_t)(t + 1) / T;
    };
    auto doc_terms = [&](uint64_t d, uint32_t* buf) -> uint32_t {
        if (field == 2) return fgs_doc_facets(h, d, buf);
        uint32_t n = c->tokens(d, field, buf);
        std::sort(buf, buf + n);
        return n;
    };
    // pass 1: per-thread document frequencies
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t]() {
                uint64_t a, b; range(t, a, b);
                uint32_t buf[512];
                for (uint64_t d = a; d < b; d++) {
                    uint32_t n = doc_terms(d, buf);
                    out->doc_len[d - d0] = n;
                    tot[t] += n;
                    for (uint32_t i = 0; i < n; i++)
                        if (i == 0 || buf[i] != buf[i - 1]) df[t][buf[i]]++;
                }
            });
        for (auto& x : th) x.join();
    }
    // offsets + per-thread write cursors (thread t's slice of term v comes after threads < t)
    std::vector<std::vector<uint64_t>> cur((size_t)T);
    for (int t = 0; t < T; t++) cur[t].resize(V);
    uint64_t np = 0;
    for (uint32_t v = 0; v < V; v++) {
        out->offsets[v] = np;
        for (int t = 0; t < T; t++) { cur[t][v] = np; np += df[t][v]; }
    }
    out->offsets[V] = np;
    out->n_postings = np;
    for (int t = 0; t < T; t++) out->total_tokens += tot[t];
    out->docs = (uint32_t*)malloc(std::max<uint64_t>(np, 1) * sizeof(uint32_t));
    out->tfs = (uint32_t*)malloc(std::max<uint64_t>(np, 1) * sizeof(uint32_t));
    // pass 2: fill
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t]() {
                uint64_t a, b; range(t, a, b);
                uint32_t buf[512];
                for (uint64_t d = a; d < b; d++) {
                    uint32_t n = doc_terms(d, buf);
                    uint32_t i = 0;
                    while (i < n) {
                        uint32_t j = i + 1;
                        while (j < n && buf[j] == buf[i]) j++;
                        uint64_t p = cur[t][buf[i]]++;
                        out->docs[p] = (uint32_t)(d - d0);
                        out->tfs[p] = j - i;
                        i = j;
                    }
                }
            });
        for (auto& x : th) x.join();
    }
    return out;
}

uint32_t fgs_csr_n_terms(void* p) { return ((Csr*)p)->n_terms; }
uint64_t fgs_csr_n_postings(void* p) { return ((Csr*)p)->n_postings; }
uint64_t fgs_csr_n_docs(void* p) { return ((Csr*)p)->n_docs; }
uint64_t fgs_csr_total_tokens(void* p) { return ((Csr*)p)->total_tokens; }
uint64_t* fgs_csr_offsets(void* p) { return ((Csr*)p)->offsets; }
uint32_t* fgs_csr_docs(void* p) { return ((Csr*)p)->docs; }
uint32_t* fgs_csr_tfs(void* p) { return ((Csr*)p)->tfs; }
uint32_t* fgs_csr_doc_len(void* p) { return ((Csr*)p)->doc_len; }
void fgs_csr_free(void* p) {
    Csr* c = (Csr*)p;
    if (!c) return;
    free(c->offsets); free(c->docs); free(c->tfs); free(c->doc_len);
    delete c;
}

// Text of one doc ("w12 w7 w12 ...") for the ObjectRecord replay kit / tokenizer-in-the-loop
// tests. field 0 = text, 1 = metadata.name. Returns length written (excluding NUL).
int fgs_doc_text(void* h, uint64_t d, int field, char* buf, int cap) {
    Corpus* c = (Corpus*)h;
    uint32_t toks[512];
    uint32_t n = c->tokens(d, field, toks);
    int pos = 0;
    for (uint32_t i = 0; i < n; i++) {
        int w = snprintf(buf + pos, cap - pos, i ? " w%u" : "w%u", toks[i] + 1);
        if (w < 0 || w >= cap - pos) return -1;
        pos += w;
    }
    if (n == 0 && cap > 0) buf[0] = 0;
    return pos;
}

}  // extern "C"
