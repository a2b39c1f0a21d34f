// oracle.cpp — TEST INFRASTRUCTURE ONLY. Never linked into, imported by, or called from the
// product path (fugu_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it.
//
// PARITY UNPINNED: the arithmetic of fugu's query path lives in the un-vendored crate
// tantivy 0.24.1 (/root/reference/Cargo.toml:48, Cargo.lock:4609-4612; call site
// /root/reference/src/db/search.rs:162). Its source is not in /root/reference, there is no Rust
// toolchain here, and the reference ships no test that pins a score, a ranking or a doc set
// (SURVEY.md 4, 8(c)). This file restates tantivy's published algorithm:
//   A.3 fieldnorm code  (Lucene SmallFloat byte4ToInt table, id = largest entry <= n_tokens)
//   A.4 Bm25Weight      (K1 = 1.2, B = 0.75, f32; idf = ln(1 + (N - df + .5)/(df + .5));
//                        cache[i] = K1*(1 - B + B*table[i]/avg); score = w*tf/(tf + cache[id]))
//   A.5 boolean scorers (TermScorer, BufferedUnionScorer with a 4096-doc horizon, Intersection
//                        ordered by cost with leap-frog seeks, RequiredOptionalScorer, Exclude)
//   A.6 TopDocs         (TopNComputer: 2k buffer + median truncation; score desc, doc asc)
//   A.5 pruning         (orc_search_batch_pruned: block maxima for unions of plain term scorers, the shape tantivy hands
//                        to block-max WAND; results identical to the exhaustive scorers, tests/test_oracle_golden.py)
// and is anchored on hand-computed known-answer vectors under tests/golden/ plus a second,
// independently written pure-Python twin (oracle/oracle_py.py).
//
// It works on the same flat CSR the reference-side loader would pass to fg_index_upload and on
// the same fg_query_batch plans (types from include/fugu_gpu.h; no product code is used).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "../include/fugu_gpu.h"

namespace {

constexpr uint32_t TERMINATED = 0x7FFFFFFFu;  // tantivy: DocId = u32, TERMINATED = i32::MAX
constexpr float K1 = 1.2f, B = 0.75f;

// ---- A.3 fieldnorm code ----------------------------------------------------------------------
uint32_t int4_decode(uint32_t i) {  // Lucene SmallFloat.byte4ToInt for the part above 24
    uint32_t bits = i & 0x07, shift = i >> 3;
    return shift == 0 ? bits : ((bits | 0x08) << (shift - 1));
}
struct FnTable {
    uint32_t t[256];
    FnTable() {
        for (uint32_t b = 0; b < 256; b++) t[b] = b < 24 ? b : 24 + int4_decode(b - 24);
    }
};
const FnTable FN;
uint8_t fieldnorm_to_id(uint32_t n) {
    int lo = 0, hi = 255;  // largest id with t[id] <= n
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (FN.t[mid] <= n) lo = mid; else hi = mid - 1;
    }
    return (uint8_t)lo;
}

// ---- A.4 BM25 ------------------------------------------------------------------------------
float idf(uint64_t df, uint64_t n) {
    float x = ((float)(n - df) + 0.5f) / ((float)df + 0.5f);
    return std::log(1.0f + x);
}
struct Bm25 {
    float weight;
    float cache[256];
    Bm25(float boost, uint64_t df, uint64_t n_docs, uint64_t total_tokens) {
        const float avg = (float)total_tokens / (float)n_docs;
        weight = boost * (idf(df, n_docs) * (1.0f + K1));
        for (int i = 0; i < 256; i++) cache[i] = K1 * (1.0f - B + B * (float)FN.t[i] / avg);
    }
    float score(uint8_t fn_id, uint32_t tf) const {
        const float t = (float)tf;
        return weight * (t / (t + cache[fn_id]));
    }
};

struct Index {
    uint32_t n_docs, doc_base;
    uint64_t global_n_docs;
    std::vector<fg_field_desc> fields;
    const uint32_t* alive;
    bool is_alive(uint32_t d) const { return !alive || ((alive[d >> 5] >> (d & 31)) & 1u); }
};

// ---- A.5 scorers -----------------------------------------------------------------------------
struct Scorer {
    virtual ~Scorer() {}
    virtual uint32_t doc() const = 0;
    virtual uint32_t advance() = 0;
    virtual uint32_t seek(uint32_t target) {
        uint32_t d = doc();
        while (d < target) d = advance();
        return d;
    }
    virtual float score() = 0;
    virtual uint64_t cost() const = 0;
};

struct EmptyScorer : Scorer {
    uint32_t doc() const override { return TERMINATED; }
    uint32_t advance() override { return TERMINATED; }
    float score() override { return 0.f; }
    uint64_t cost() const override { return 0; }
};

struct TermScorer : Scorer {
    const uint32_t* docs;
    const uint32_t* tfs;  // may be null (tf == 1)
    const uint8_t* fn;    // may be null (constant fieldnorm 1)
    size_t n, i = 0;
    Bm25 bm;
    uint8_t const_id;
    TermScorer(const uint32_t* d, const uint32_t* t, const uint8_t* f, size_t n_, const Bm25& b)
        : docs(d), tfs(t), fn(f), n(n_), bm(b), const_id(fieldnorm_to_id(1)) {}
    uint32_t doc() const override { return i < n ? docs[i] : TERMINATED; }
    uint32_t advance() override { i++; return doc(); }
    uint32_t seek(uint32_t target) override {
        if (doc() >= target) return doc();
        // gallop then binary search (tantivy: skip list + in-block search)
        size_t step = 1, lo = i, hi = i + 1;
        while (hi < n && docs[hi] < target) { lo = hi; step <<= 1; hi = std::min(n, hi + step); }
        i = std::lower_bound(docs + lo, docs + std::min(hi + 1, n), target) - docs;
        return doc();
    }
    float score() override { return bm.score(fn ? fn[docs[i]] : const_id, tfs ? tfs[i] : 1u); }
    uint64_t cost() const override { return n; }
};

// BufferedUnionScorer: 64 x 64-bit bitset = 4096-doc horizon, per-slot score sum
struct UnionScorer : Scorer {
    static constexpr uint32_t HORIZON = 4096;
    std::vector<std::unique_ptr<Scorer>> subs;
    uint64_t bits[64];
    float sums[HORIZON];
    uint32_t cursor = 64, offset = 0, cur = 0;
    float cur_score = 0.f;
    uint64_t cost_ = 0;
    explicit UnionScorer(std::vector<std::unique_ptr<Scorer>> s) : subs(std::move(s)) {
        std::memset(bits, 0, sizeof(bits));
        for (uint32_t i = 0; i < HORIZON; i++) sums[i] = 0.f;
        for (auto& x : subs) cost_ += x->cost();
        drop_terminated();
        if (refill()) advance(); else cur = TERMINATED;
    }
    void drop_terminated() {
        size_t w = 0;
        for (size_t r = 0; r < subs.size(); r++)
            if (subs[r]->doc() != TERMINATED) { if (w != r) subs[w] = std::move(subs[r]); w++; }
        subs.resize(w);
    }
    bool refill() {
        if (subs.empty()) return false;
        uint32_t mn = TERMINATED;
        for (auto& s : subs) mn = std::min(mn, s->doc());
        offset = mn;
        cursor = 0;
        for (auto& s : subs) {
            if (TermScorer* t = dynamic_cast<TermScorer*>(s.get())) {
                // plain term scorer (the common child): the same loop on its arrays, without a virtual call per posting
                // (tantivy's union is generic over its scorer type and gets the same effect from monomorphisation)
                const uint64_t lim = (uint64_t)mn + HORIZON;
                size_t i = t->i;
                while (i < t->n && t->docs[i] < lim) {
                    const uint32_t delta = t->docs[i] - mn;
                    bits[delta >> 6] |= 1ull << (delta & 63);
                    sums[delta] += t->bm.score(t->fn ? t->fn[t->docs[i]] : t->const_id, t->tfs ? t->tfs[i] : 1u);
                    i++;
                }
                t->i = i;
                continue;
            }
            while (true) {
                uint32_t d = s->doc();
                if (d >= mn + HORIZON || d == TERMINATED) break;
                uint32_t delta = d - mn;
                bits[delta >> 6] |= 1ull << (delta & 63);
                sums[delta] += s->score();
                s->advance();
            }
        }
        drop_terminated();
        return true;
    }
    bool advance_buffered() {
        while (cursor < 64) {
            if (bits[cursor]) {
                uint32_t b = (uint32_t)__builtin_ctzll(bits[cursor]);
                bits[cursor] &= bits[cursor] - 1;
                uint32_t delta = (cursor << 6) | b;
                cur = offset + delta;
                cur_score = sums[delta];
                sums[delta] = 0.f;
                return true;
            }
            cursor++;
        }
        return false;
    }
    uint32_t doc() const override { return cur; }
    uint32_t advance() override {
        if (advance_buffered()) return cur;
        if (!refill()) { cur = TERMINATED; return cur; }
        if (!advance_buffered()) cur = TERMINATED;
        return cur;
    }
    float score() override { return cur_score; }
    uint64_t cost() const override { return cost_; }
};

struct IntersectionScorer : Scorer {
    std::vector<std::unique_ptr<Scorer>> subs;  // ascending cost
    explicit IntersectionScorer(std::vector<std::unique_ptr<Scorer>> s) : subs(std::move(s)) {
        std::stable_sort(subs.begin(), subs.end(),
                         [](const std::unique_ptr<Scorer>& a, const std::unique_ptr<Scorer>& b) {
                             return a->cost() < b->cost();
                         });
        align(subs[0]->doc());
    }
    uint32_t align(uint32_t cand) {
        while (cand != TERMINATED) {
            bool ok = true;
            for (size_t i = 1; i < subs.size(); i++) {
                uint32_t d = subs[i]->seek(cand);
                if (d > cand) { cand = subs[0]->seek(d); ok = false; break; }
            }
            if (ok) break;
        }
        return cand;
    }
    uint32_t doc() const override { return subs[0]->doc(); }
    uint32_t advance() override { return align(subs[0]->advance()); }
    float score() override {
        float s = 0.f;
        for (auto& x : subs) s += x->score();
        return s;
    }
    uint64_t cost() const override { return subs[0]->cost(); }
};

struct RequiredOptionalScorer : Scorer {
    std::unique_ptr<Scorer> req, opt;
    RequiredOptionalScorer(std::unique_ptr<Scorer> r, std::unique_ptr<Scorer> o)
        : req(std::move(r)), opt(std::move(o)) {}
    uint32_t doc() const override { return req->doc(); }
    uint32_t advance() override { return req->advance(); }
    float score() override {
        float s = req->score();
        const uint32_t d = req->doc();
        if (opt->doc() <= d && opt->seek(d) == d) s += opt->score();
        return s;
    }
    uint64_t cost() const override { return req->cost(); }
};

struct ExcludeScorer : Scorer {
    std::unique_ptr<Scorer> und, exc;
    ExcludeScorer(std::unique_ptr<Scorer> u, std::unique_ptr<Scorer> e)
        : und(std::move(u)), exc(std::move(e)) {
        while (und->doc() != TERMINATED && excluded()) und->advance();
    }
    bool excluded() {
        const uint32_t d = und->doc();
        return exc->doc() <= d && exc->seek(d) == d;
    }
    uint32_t doc() const override { return und->doc(); }
    uint32_t advance() override {
        do { if (und->advance() == TERMINATED) break; } while (excluded());
        return und->doc();
    }
    float score() override { return und->score(); }
    uint64_t cost() const override { return und->cost(); }
};

struct ConstAddScorer : Scorer {  // Must(AllQuery) sibling: +boost on every hit, no restriction
    std::unique_ptr<Scorer> und;
    float add;
    ConstAddScorer(std::unique_ptr<Scorer> u, float a) : und(std::move(u)), add(a) {}
    uint32_t doc() const override { return und->doc(); }
    uint32_t advance() override { return und->advance(); }
    uint32_t seek(uint32_t t) override { return und->seek(t); }
    float score() override { return und->score() + add; }
    uint64_t cost() const override { return und->cost(); }
};

std::unique_ptr<Scorer> make_union(std::vector<std::unique_ptr<Scorer>> v) {
    if (v.empty()) return std::unique_ptr<Scorer>(new EmptyScorer());
    if (v.size() == 1) return std::move(v[0]);
    return std::unique_ptr<Scorer>(new UnionScorer(std::move(v)));
}

struct BuiltQuery {
    std::unique_ptr<Scorer> scorer;
    bool unsupported = false;
};

std::unique_ptr<Scorer> leaf_scorer(const Index& ix, const fg_leaf& lf) {
    if (lf.term_ord == FG_TERM_MISSING) return std::unique_ptr<Scorer>(new EmptyScorer());
    const fg_field_desc& f = ix.fields[lf.field];
    const uint64_t a = f.term_offsets[lf.term_ord], b = f.term_offsets[lf.term_ord + 1];
    const uint64_t gdf = f.global_doc_freq ? f.global_doc_freq[lf.term_ord] : (b - a);
    if (gdf == 0 || a == b) return std::unique_ptr<Scorer>(new EmptyScorer());
    Bm25 bm(lf.boost, gdf, ix.global_n_docs, f.total_num_tokens);
    const bool freqs = (f.flags & FG_FIELD_HAS_FREQS) && f.term_freqs;
    const bool norms = (f.flags & FG_FIELD_HAS_FIELDNORMS) && f.fieldnorm_ids;
    return std::unique_ptr<Scorer>(new TermScorer(f.doc_ids + a, freqs ? f.term_freqs + a : nullptr,
                                                  norms ? f.fieldnorm_ids : nullptr, b - a, bm));
}

// BooleanWeight::complex_scorer restated for one level of grouping
BuiltQuery build(const Index& ix, const fg_query_batch& qb, const fg_query& q) {
    std::vector<std::unique_ptr<Scorer>> must, should, mnot;
    float const_add = 0.f;
    bool all_must = false;
    BuiltQuery out;
    for (uint32_t ci = 0; ci < q.n_clauses; ci++) {
        const fg_clause& c = qb.clauses[q.clause_begin + ci];
        std::vector<std::unique_ptr<Scorer>> leaves;
        bool all = false;
        float all_boost = 0.f;
        for (uint32_t li = 0; li < c.n_leaves; li++) {
            const fg_leaf& lf = qb.leaves[c.leaf_begin + li];
            if (lf.term_ord == FG_TERM_ALL) { all = true; all_boost += lf.boost; continue; }
            leaves.push_back(leaf_scorer(ix, lf));
        }
        if (all) {
            if (c.occur == FG_OCCUR_MUST && leaves.empty()) { const_add += all_boost; all_must = true; continue; }
            out.unsupported = true;
            return out;
        }
        auto u = make_union(std::move(leaves));
        (c.occur == FG_OCCUR_MUST ? must : c.occur == FG_OCCUR_SHOULD ? should : mnot).push_back(std::move(u));
    }
    std::unique_ptr<Scorer> pos;
    if (!must.empty()) {
        std::unique_ptr<Scorer> m = must.size() == 1 ? std::move(must[0])
                                                     : std::unique_ptr<Scorer>(new IntersectionScorer(std::move(must)));
        if (!should.empty())
            pos.reset(new RequiredOptionalScorer(std::move(m), make_union(std::move(should))));
        else
            pos = std::move(m);
    } else if (all_must) {
        out.unsupported = true;  // pure AllQuery (+ optional siblings): answered by the host layer
        return out;
    } else if (!should.empty()) {
        pos = make_union(std::move(should));
    } else {
        pos.reset(new EmptyScorer());
    }
    if (!mnot.empty()) pos.reset(new ExcludeScorer(std::move(pos), make_union(std::move(mnot))));
    if (const_add != 0.f) pos.reset(new ConstAddScorer(std::move(pos), const_add));
    out.scorer = std::move(pos);
    return out;
}

// ---- A.6 TopNComputer: buffer of 2k, truncate at the median, (score desc, doc asc) ------------
struct Hit { float score; uint32_t doc; };
inline bool better(const Hit& a, const Hit& b) { return a.score > b.score || (a.score == b.score && a.doc < b.doc); }
struct TopN {
    size_t k;
    std::vector<Hit> buf;
    bool has_thr = false;
    Hit thr{};
    explicit TopN(size_t k_) : k(k_) { buf.reserve(2 * k_); }
    void push(float s, uint32_t d) {
        Hit h{s, d};
        if (has_thr && !better(h, thr)) return;
        if (buf.size() == 2 * k) truncate();
        buf.push_back(h);
    }
    void truncate() {
        std::nth_element(buf.begin(), buf.begin() + (k - 1), buf.end(), better);
        thr = buf[k - 1];
        has_thr = true;
        buf.resize(k);
    }
    std::vector<Hit> finish() {
        std::sort(buf.begin(), buf.end(), better);
        if (buf.size() > k) buf.resize(k);
        return buf;
    }
};

int run_query(const Index& ix, const fg_query_batch& qb, uint32_t qi, uint32_t k_stride, fg_hit* hits,
              uint32_t* n_hits, uint32_t* counts, uint32_t* match_bitmap, uint32_t bitmap_words) {
    const fg_query& q = qb.queries[qi];
    if (q.k == 0) return FG_ERR_INVALID;
    BuiltQuery bq = build(ix, qb, q);
    if (bq.unsupported) return FG_ERR_UNSUPPORTED;
    TopN top(q.k);
    uint32_t cnt = 0;
    Scorer& s = *bq.scorer;
    for (uint32_t d = s.doc(); d != TERMINATED; d = s.advance()) {
        if (!ix.is_alive(d)) continue;
        cnt++;
        if (match_bitmap) match_bitmap[(size_t)qi * bitmap_words + (d >> 5)] |= 1u << (d & 31);
        top.push(s.score(), d);
    }
    std::vector<Hit> r = top.finish();
    const uint32_t n = (uint32_t)std::min<size_t>(r.size(), k_stride);
    for (uint32_t i = 0; i < n; i++) {
        hits[(size_t)qi * k_stride + i].score = r[i].score;
        hits[(size_t)qi * k_stride + i].doc = r[i].doc + ix.doc_base;
    }
    for (uint32_t i = n; i < k_stride; i++) {
        hits[(size_t)qi * k_stride + i].score = 0.f;
        hits[(size_t)qi * k_stride + i].doc = 0xFFFFFFFFu;
    }
    n_hits[qi] = n;
    if (counts) counts[qi] = cnt;
    return FG_OK;
}

// ---- block-max pruning for unions of plain term scorers (A.5: BooleanWeight hands a Should-only query whose
// children are all TermScorers to block-max WAND instead of the buffered union; fugu's two default fields make a
// single-word query such a union, (text:w OR name:w), while a multi-word query is a union of unions and keeps the
// buffered union). tantivy's index stores the block-max (fieldnorm id, tf) pair of every 128-doc block in its skip
// entries; the equivalent here is built once per index (orc_blockmax_build) and only the TopDocs form uses it
// (counting visits every match). Restated as MaxScore with block maxima: same result set and scores as the exhaustive
// evaluation (scores are summed in leaf order either way).
struct BlockMax {
    // per field: offset of each term's first block in `factor`, then one float per 128 postings = the largest
    // tf / (tf + cache[fieldnorm id]) of the block under the index's statistics
    std::vector<std::vector<uint64_t>> term_block0;
    std::vector<std::vector<float>> factor;
};
constexpr size_t PBLOCK = 128;

struct PrunedLeaf {
    const uint32_t* docs;
    const uint32_t* tfs;
    const uint8_t* fn;
    size_t n, i = 0;
    const float* bmax;  // per block of 128 postings
    float w, gmax;      // weight, largest possible score of the list
    Bm25 bm;
    uint8_t const_id;
    float score_at(size_t j) const { return bm.score(fn ? fn[docs[j]] : const_id, tfs ? tfs[j] : 1u); }
};

bool prunable(const fg_query_batch& qb, const fg_query& q) {
    if (q.n_clauses == 0) return false;
    uint32_t multi = 0;
    for (uint32_t ci = 0; ci < q.n_clauses; ci++) {
        const fg_clause& c = qb.clauses[q.clause_begin + ci];
        if (c.occur != FG_OCCUR_SHOULD) return false;
        for (uint32_t li = 0; li < c.n_leaves; li++)
            if (qb.leaves[c.leaf_begin + li].term_ord == FG_TERM_ALL || !(qb.leaves[c.leaf_begin + li].boost > 0.f)) return false;
        if (c.n_leaves > 1) multi++;
    }
    // one clause of several leaves (a bare word over the default fields) or several single-leaf clauses: every child of
    // the top-level union is a term scorer. A clause of several leaves next to other clauses is a union inside a union.
    return multi == 0 || q.n_clauses == 1;
}

int run_query_pruned(const Index& ix, const BlockMax& bmx, const fg_query_batch& qb, uint32_t qi, uint32_t k_stride, fg_hit* hits, uint32_t* n_hits) {
    const fg_query& q = qb.queries[qi];
    if (q.k == 0) return FG_ERR_INVALID;
    std::vector<PrunedLeaf> L;
    for (uint32_t ci = 0; ci < q.n_clauses; ci++) {
        const fg_clause& c = qb.clauses[q.clause_begin + ci];
        for (uint32_t li = 0; li < c.n_leaves; li++) {
            const fg_leaf& lf = qb.leaves[c.leaf_begin + li];
            if (lf.term_ord == FG_TERM_MISSING) continue;
            const fg_field_desc& f = ix.fields[lf.field];
            const uint64_t a = f.term_offsets[lf.term_ord], b = f.term_offsets[lf.term_ord + 1];
            const uint64_t gdf = f.global_doc_freq ? f.global_doc_freq[lf.term_ord] : (b - a);
            if (gdf == 0 || a == b) continue;
            const bool freqs = (f.flags & FG_FIELD_HAS_FREQS) && f.term_freqs;
            const bool norms = (f.flags & FG_FIELD_HAS_FIELDNORMS) && f.fieldnorm_ids;
            PrunedLeaf p{f.doc_ids + a, freqs ? f.term_freqs + a : nullptr, norms ? f.fieldnorm_ids : nullptr, (size_t)(b - a), 0,
                         bmx.factor[lf.field].data() + bmx.term_block0[lf.field][lf.term_ord], 0.f, 0.f,
                         Bm25(lf.boost, gdf, ix.global_n_docs, f.total_num_tokens), fieldnorm_to_id(1)};
            p.w = p.bm.weight;
            float mx = 0.f;
            for (size_t blk = 0; blk < (p.n + PBLOCK - 1) / PBLOCK; blk++) mx = std::max(mx, p.bmax[blk]);
            p.gmax = p.w * mx * 1.00001f;  // (bounds carry a relative slack: the block maxima were computed in another expression order)
            L.push_back(p);
        }
    }
    const size_t m = L.size();
    // MaxScore over the lists, shortest first: list i offers the docs none of the shorter lists holds (those are scored
    // when their list is walked), looked up in the longer ones; a block of list i is skipped when
    // w_i * bmax_i(block) + the maxima of the longer lists cannot reach the threshold. The hit collector does not depend
    // on the order docs arrive in ((score, doc) decides), and a doc's score is summed in leaf order as the union does.
    std::vector<size_t> order(m);
    for (size_t i = 0; i < m; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return L[x].n < L[y].n; });
    TopN top(q.k);
    std::vector<size_t> cur(m);
    std::vector<float> part(m);
    for (size_t oi = 0; oi < m; oi++) {
        PrunedLeaf& P = L[order[oi]];
        float rest = 0.f;
        for (size_t oj = oi + 1; oj < m; oj++) rest += L[order[oj]].gmax;
        for (size_t j = 0; j < m; j++) cur[j] = 0;
        const size_t nblk = (P.n + PBLOCK - 1) / PBLOCK;
        for (size_t blk = 0; blk < nblk; blk++) {
            if (top.has_thr && P.w * P.bmax[blk] * 1.00001f + rest < top.thr.score) continue;
            const size_t e = std::min(P.n, (blk + 1) * PBLOCK);
            for (size_t x = blk * PBLOCK; x < e; x++) {
                const uint32_t d = P.docs[x];
                bool owned = false;  // a shorter list holds d
                for (size_t oj = 0; oj < m; oj++) part[order[oj]] = 0.f;
                for (size_t oj = 0; oj < m && !owned; oj++) {
                    if (oj == oi) { part[order[oi]] = P.score_at(x); continue; }
                    PrunedLeaf& Q = L[order[oj]];
                    size_t& c = cur[order[oj]];
                    // gallop from the cursor (docs of P ascend)
                    size_t step = 1, lo = c, hi = c;
                    while (hi < Q.n && Q.docs[hi] < d) { lo = hi; hi = std::min(Q.n, hi + step); step <<= 1; }
                    c = (size_t)(std::lower_bound(Q.docs + lo, Q.docs + hi, d) - Q.docs);
                    if (c < Q.n && Q.docs[c] == d) {
                        if (oj < oi) owned = true;
                        else part[order[oj]] = Q.score_at(c);
                    }
                }
                if (owned || !ix.is_alive(d)) continue;
                float sc = 0.f;
                for (size_t j = 0; j < m; j++) sc += part[j];  // (absent leaves add +0: the sum equals the union's)
                top.push(sc, d);
            }
        }
    }
    std::vector<Hit> r = top.finish();
    const uint32_t n = (uint32_t)std::min<size_t>(r.size(), k_stride);
    for (uint32_t i = 0; i < n; i++) {
        hits[(size_t)qi * k_stride + i].score = r[i].score;
        hits[(size_t)qi * k_stride + i].doc = r[i].doc + ix.doc_base;
    }
    for (uint32_t i = n; i < k_stride; i++) {
        hits[(size_t)qi * k_stride + i].score = 0.f;
        hits[(size_t)qi * k_stride + i].doc = 0xFFFFFFFFu;
    }
    n_hits[qi] = n;
    return FG_OK;
}

Index make_index(const fg_index_desc* d) {
    Index ix;
    ix.n_docs = d->n_docs;
    ix.doc_base = d->doc_id_base;
    ix.global_n_docs = d->global_n_docs ? d->global_n_docs : d->n_docs;
    ix.fields.assign(d->fields, d->fields + d->n_fields);
    ix.alive = d->alive_bitset;
    return ix;
}

// ---- algorithmic bytes on the library's block layout rule, recomputed here ----
inline uint32_t bits_of(uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; }
struct Blk { uint32_t first_base, last_doc, bytes; size_t i0, n; };
std::vector<Blk> blocks_of(const uint32_t* docs, const uint32_t* tfs, size_t n) {
    std::vector<Blk> out;
    uint32_t prev1 = 0;
    for (size_t i0 = 0; i0 < n; i0 += 128) {
        size_t m = std::min<size_t>(128, n - i0);
        uint32_t gor = 0, tor = 0, p = prev1;
        for (size_t i = 0; i < m; i++) {
            gor |= docs[i0 + i] - p;
            p = docs[i0 + i] + 1;
            if (tfs) tor |= tfs[i0 + i] - 1;
        }
        Blk b;
        b.first_base = prev1;
        b.last_doc = docs[i0 + m - 1];
        b.bytes = (uint32_t)((m * bits_of(gor) + 7) / 8 + (m * bits_of(tor) + 7) / 8 + 16);
        b.i0 = i0;
        b.n = m;
        out.push_back(b);
        prev1 = p;
    }
    return out;
}
struct LeafData { const uint32_t* docs; const uint32_t* tfs; size_t n; };
bool leaf_data(const Index& ix, const fg_leaf& lf, LeafData& o) {
    if (lf.term_ord == FG_TERM_MISSING || lf.term_ord == FG_TERM_ALL) return false;
    const fg_field_desc& f = ix.fields[lf.field];
    const uint64_t a = f.term_offsets[lf.term_ord], b = f.term_offsets[lf.term_ord + 1];
    if (a == b) return false;
    const bool freqs = (f.flags & FG_FIELD_HAS_FREQS) && f.term_freqs;
    o.docs = f.doc_ids + a;
    o.tfs = freqs ? f.term_freqs + a : nullptr;
    o.n = b - a;
    return true;
}
std::vector<uint32_t> union_docs(const std::vector<LeafData>& ls) {
    std::vector<uint32_t> u;
    for (auto& l : ls) u.insert(u.end(), l.docs, l.docs + l.n);
    std::sort(u.begin(), u.end());
    u.erase(std::unique(u.begin(), u.end()), u.end());
    return u;
}
// bytes of the blocks of leaf `l` whose range (first_base .. last_doc) holds >= 1 candidate, and
// the number of its postings that are candidates
void filtered_cost(const LeafData& l, const std::vector<uint32_t>& cand, uint64_t& bytes, uint64_t& scored) {
    for (const Blk& b : blocks_of(l.docs, l.tfs, l.n)) {
        auto it = std::lower_bound(cand.begin(), cand.end(), b.first_base);
        if (it != cand.end() && *it <= b.last_doc) bytes += b.bytes;
    }
    size_t i = 0, j = 0;
    while (i < l.n && j < cand.size()) {
        if (l.docs[i] < cand[j]) i++;
        else if (l.docs[i] > cand[j]) j++;
        else { scored++; i++; j++; }
    }
}

void algo_query(const Index& ix, const fg_query_batch& qb, const fg_query& q, uint64_t& bytes, uint64_t& scored) {
    struct Cl { uint64_t cost; std::vector<LeafData> leaves; };
    std::vector<Cl> must, should, mnot;
    bool empty = false;
    for (uint32_t ci = 0; ci < q.n_clauses; ci++) {
        const fg_clause& c = qb.clauses[q.clause_begin + ci];
        Cl cl{0, {}};
        bool all = false;
        for (uint32_t li = 0; li < c.n_leaves; li++) {
            const fg_leaf& lf = qb.leaves[c.leaf_begin + li];
            if (lf.term_ord == FG_TERM_ALL) { all = true; continue; }
            LeafData d;
            if (leaf_data(ix, lf, d)) { cl.leaves.push_back(d); cl.cost += d.n; }
        }
        if (all && cl.leaves.empty()) continue;
        if (cl.leaves.empty()) { if (c.occur == FG_OCCUR_MUST) empty = true; continue; }
        (c.occur == FG_OCCUR_MUST ? must : c.occur == FG_OCCUR_SHOULD ? should : mnot).push_back(cl);
    }
    if (empty || (must.empty() && should.empty())) return;
    auto full_cost = [&](const LeafData& l) {
        for (const Blk& b : blocks_of(l.docs, l.tfs, l.n)) bytes += b.bytes;
        scored += l.n;
    };
    std::vector<uint32_t> cand;
    if (!must.empty()) {
        std::stable_sort(must.begin(), must.end(), [](const Cl& a, const Cl& b) { return a.cost < b.cost; });
        for (auto& l : must[0].leaves) full_cost(l);
        cand = union_docs(must[0].leaves);
        for (size_t ci = 1; ci < must.size(); ci++) {
            for (auto& l : must[ci].leaves) filtered_cost(l, cand, bytes, scored);
            std::vector<uint32_t> u = union_docs(must[ci].leaves), nx;
            std::set_intersection(cand.begin(), cand.end(), u.begin(), u.end(), std::back_inserter(nx));
            cand.swap(nx);
        }
        for (auto& c : should)
            for (auto& l : c.leaves) filtered_cost(l, cand, bytes, scored);
    } else {
        std::vector<LeafData> all;
        for (auto& c : should)
            for (auto& l : c.leaves) { full_cost(l); all.push_back(l); }
        if (!mnot.empty()) cand = union_docs(all);
    }
    for (auto& c : mnot)
        for (auto& l : c.leaves) {
            uint64_t dummy = 0;
            filtered_cost(l, cand, bytes, dummy);  // MustNot postings are never scored
        }
}

}  // namespace

extern "C" {

uint8_t orc_fieldnorm_to_id(uint32_t n) { return fieldnorm_to_id(n); }
uint32_t orc_id_to_fieldnorm(uint8_t id) { return FN.t[id]; }
float orc_idf(uint64_t df, uint64_t n) { return idf(df, n); }
float orc_bm25_score(float boost, uint64_t df, uint64_t n_docs, uint64_t total_tokens, uint8_t fn_id, uint32_t tf) {
    return Bm25(boost, df, n_docs, total_tokens).score(fn_id, tf);
}

// Same signature shape as fg_search_batch, but over the raw CSR description (nothing is copied;
// the descriptor's arrays must stay alive for the call). match_bitmap may be NULL.
int32_t orc_search_batch(const fg_index_desc* d, const fg_query_batch* qb, uint32_t k_stride,
                         fg_hit* hits, uint32_t* n_hits, uint32_t* counts, uint32_t* match_bitmap,
                         int32_t n_threads) {
    Index ix = make_index(d);
    const uint32_t words = (d->n_docs + 31) / 32;
    std::atomic<uint32_t> next(0);
    std::atomic<int32_t> rc(FG_OK);
    auto work = [&]() {
        while (true) {
            uint32_t qi = next.fetch_add(1);
            if (qi >= qb->n_queries) break;
            int r = run_query(ix, *qb, qi, k_stride, hits, n_hits, counts, match_bitmap, words);
            if (r != FG_OK) rc.store(r);
        }
    };
    if (n_threads <= 1) work();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(work);
        for (auto& t : th) t.join();
    }
    return rc.load();
}

// One query whose Should children are the batch's queries (each a one-level plan): tantivy's BooleanQuery of boolean
// queries, e.g. `(a AND b) OR (c AND d)`. complex_scorer builds each child's scorer and puts them under the buffered
// union (A.5): a document matches when any child matches and scores the sum of the matching children.
// With n_filters > 0 the LAST n_filters queries are Must siblings of that union -- Bool[Must(union), Must(f)..], what
// Dataset::search builds from a nested text query and a facet query (/root/reference/src/db/search.rs:140-144): an
// Intersection of the union and the filters (counterpart of fg_search_union_of_filtered).
int32_t orc_search_union_of_filtered(const fg_index_desc* d, const fg_query_batch* disjuncts, uint32_t n_filters, uint32_t k, fg_hit* hits,
                                     uint32_t* n_hits, uint32_t* match_count) {
    if (k == 0 || n_filters >= disjuncts->n_queries) return FG_ERR_INVALID;
    Index ix = make_index(d);
    std::vector<std::unique_ptr<Scorer>> kids, filters;
    for (uint32_t qi = 0; qi < disjuncts->n_queries; qi++) {
        BuiltQuery bq = build(ix, *disjuncts, disjuncts->queries[qi]);
        if (bq.unsupported) return FG_ERR_UNSUPPORTED;
        (qi + n_filters >= disjuncts->n_queries ? filters : kids).push_back(std::move(bq.scorer));
    }
    std::unique_ptr<Scorer> u = make_union(std::move(kids));
    if (!filters.empty()) {
        std::vector<std::unique_ptr<Scorer>> musts;
        musts.push_back(std::move(u));
        for (auto& f : filters) musts.push_back(std::move(f));
        u.reset(new IntersectionScorer(std::move(musts)));
    }
    TopN top(k);
    uint32_t cnt = 0;
    for (uint32_t doc = u->doc(); doc != TERMINATED; doc = u->advance()) {
        if (!ix.is_alive(doc)) continue;
        cnt++;
        top.push(u->score(), doc);
    }
    std::vector<Hit> r = top.finish();
    for (size_t i = 0; i < r.size(); i++) {
        hits[i].score = r[i].score;
        hits[i].doc = r[i].doc + ix.doc_base;
    }
    *n_hits = (uint32_t)r.size();
    if (match_count) *match_count = cnt;
    return FG_OK;
}
int32_t orc_search_union_of(const fg_index_desc* d, const fg_query_batch* disjuncts, uint32_t k, fg_hit* hits, uint32_t* n_hits,
                            uint32_t* match_count) {
    return orc_search_union_of_filtered(d, disjuncts, 0, k, hits, n_hits, match_count);
}

// Block-max metadata of an index (what tantivy keeps in its skip entries): built once, released with orc_blockmax_free.
void* orc_blockmax_build(const fg_index_desc* d, int32_t n_threads) {
    Index ix = make_index(d);
    BlockMax* bm = new BlockMax();
    bm->term_block0.resize(ix.fields.size());
    bm->factor.resize(ix.fields.size());
    for (size_t f = 0; f < ix.fields.size(); f++) {
        const fg_field_desc& fd = ix.fields[f];
        std::vector<uint64_t>& b0 = bm->term_block0[f];
        b0.resize((size_t)fd.n_terms + 1);
        uint64_t nb = 0;
        for (uint32_t t = 0; t < fd.n_terms; t++) {
            b0[t] = nb;
            nb += (fd.term_offsets[t + 1] - fd.term_offsets[t] + PBLOCK - 1) / PBLOCK;
        }
        b0[fd.n_terms] = nb;
        bm->factor[f].assign(nb, 0.f);
        if (!fd.n_terms) continue;
        const bool freqs = (fd.flags & FG_FIELD_HAS_FREQS) && fd.term_freqs;
        const bool norms = (fd.flags & FG_FIELD_HAS_FIELDNORMS) && fd.fieldnorm_ids;
        const Bm25 unit(1.0f, 1, ix.global_n_docs, fd.total_num_tokens);  // (only its norm cache is used)
        const uint8_t cid = fieldnorm_to_id(1);
        std::atomic<uint32_t> next(0);
        auto work = [&]() {
            while (true) {
                const uint32_t t0 = next.fetch_add(256);
                if (t0 >= fd.n_terms) break;
                for (uint32_t t = t0; t < std::min(fd.n_terms, t0 + 256); t++) {
                    const uint64_t a = fd.term_offsets[t], b = fd.term_offsets[t + 1];
                    for (uint64_t i = a; i < b; i++) {
                        const float tf = (float)(freqs ? fd.term_freqs[i] : 1u);
                        const float fac = tf / (tf + unit.cache[norms ? fd.fieldnorm_ids[fd.doc_ids[i]] : cid]);
                        float& slot = bm->factor[f][b0[t] + (i - a) / PBLOCK];
                        slot = std::max(slot, fac);
                    }
                }
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < std::max(1, n_threads); t++) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
    }
    return bm;
}
void orc_blockmax_free(void* h) { delete (BlockMax*)h; }

// The TopDocs form with block-max pruning where tantivy prunes (unions whose children are all term scorers); every other
// query runs the exhaustive scorers of orc_search_batch. No match counts (pruning does not visit every match).
int32_t orc_search_batch_pruned(const fg_index_desc* d, const void* blockmax, const fg_query_batch* qb, uint32_t k_stride,
                                fg_hit* hits, uint32_t* n_hits, int32_t n_threads) {
    Index ix = make_index(d);
    const BlockMax& bm = *(const BlockMax*)blockmax;
    std::atomic<uint32_t> next(0);
    std::atomic<int32_t> rc(FG_OK);
    auto work = [&]() {
        while (true) {
            uint32_t qi = next.fetch_add(1);
            if (qi >= qb->n_queries) break;
            int r = prunable(*qb, qb->queries[qi]) ? run_query_pruned(ix, bm, *qb, qi, k_stride, hits, n_hits)
                                                   : run_query(ix, *qb, qi, k_stride, hits, n_hits, nullptr, nullptr, 0);
            if (r != FG_OK) rc.store(r);
        }
    };
    if (n_threads <= 1) work();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(work);
        for (auto& t : th) t.join();
    }
    return rc.load();
}

// Algorithmic posting bytes / scored postings per query; out arrays [n_queries].
int32_t orc_algorithmic_bytes(const fg_index_desc* d, const fg_query_batch* qb, uint64_t* out_block_bytes,
                              uint64_t* out_scored, int32_t n_threads) {
    Index ix = make_index(d);
    std::atomic<uint32_t> next(0);
    auto work = [&]() {
        while (true) {
            uint32_t qi = next.fetch_add(1);
            if (qi >= qb->n_queries) break;
            uint64_t b = 0, s = 0;
            algo_query(ix, *qb, qb->queries[qi], b, s);
            out_block_bytes[qi] = b;
            out_scored[qi] = s;
        }
    };
    if (n_threads <= 1) work();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(work);
        for (auto& t : th) t.join();
    }
    return FG_OK;
}

}  // extern "C"
