// Host side of the query path above the device ABI: tokenizer, tantivy-style query parser,
// the planning half of Dataset::search and a small dataset builder (include/fugu_host.h).
// Mirrors /root/reference/src/db/search.rs:74-324,594-610 (fugu side) and SURVEY.md Appendix
// A.1/A.2 (tantivy 0.24.1 side, upstream-recalled). No search arithmetic happens here: every
// search goes to the device through fg_search_batch.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <deque>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/fugu_host.h"
#include <dlfcn.h>
#include <time.h>
#include "fg_pool.h"
#include "fg_unicode_tables.h"

// ------------------------------------------------------------------------------------------
// errors, and the device library behind this one
// ------------------------------------------------------------------------------------------
// libfugu_host.so (this file) contains no CUDA code and does not link against libfugu_gpu.so: planning,
// tokenising and the dataset builder work in a process that never touches a GPU (the reference arm of the
// bench, a GPU-less build of the Rust host). The device entry points are bound on first use: from the
// library this code is linked into when that one exports them (a monolithic build), else from
// libfugu_gpu.so in the directory of libfugu_host.so. Without them every call that needs the device
// fails with FG_ERR_NO_DEVICE.
namespace {

thread_local char g_herr[512] = "";
int32_t host_fail(int32_t code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_herr, sizeof(g_herr), fmt, ap);
    va_end(ap);
    return code;
}

struct DevApi {
#define FGH_DEV_FN(name) decltype(&::name) name = nullptr
    FGH_DEV_FN(fg_last_error);
    FGH_DEV_FN(fg_index_upload);
    FGH_DEV_FN(fg_index_release);
    FGH_DEV_FN(fg_index_with_alive);
    FGH_DEV_FN(fg_index_append);
    FGH_DEV_FN(fg_index_term_info);
    FGH_DEV_FN(fg_search_batch);
    FGH_DEV_FN(fg_search_union_of);
    FGH_DEV_FN(fg_search_union_of_filtered);
    FGH_DEV_FN(fg_batch_prepare_ex);
    FGH_DEV_FN(fg_batch_query_status);
    FGH_DEV_FN(fg_batch_submit);
    FGH_DEV_FN(fg_batch_submit_sharded);
    FGH_DEV_FN(fg_comm_info);
    FGH_DEV_FN(fg_comm_allgather_bytes);
    FGH_DEV_FN(fg_batch_collect);
    FGH_DEV_FN(fg_batch_release);
#undef FGH_DEV_FN
    bool ok = false;
    std::string why;
};

const DevApi& dev_api() {
    static const DevApi api = [] {
        DevApi a;
        Dl_info self;
        memset(&self, 0, sizeof(self));
        std::string own = dladdr((const void*)&host_fail, &self) && self.dli_fname ? self.dli_fname : "";
        void* h = own.empty() ? nullptr : dlopen(own.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!h || !dlsym(h, "fg_batch_prepare_ex")) {
            const size_t slash = own.rfind('/');
            const std::string sibling = (slash == std::string::npos ? std::string("") : own.substr(0, slash + 1)) + "libfugu_gpu.so";
            h = dlopen(sibling.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (!h) {
                const char* e = dlerror();
                a.why = "device library " + sibling + " cannot be loaded (" + (e ? e : "?") + "); there is no CPU fallback";
                return a;
            }
        }
        bool all = true;
#define FGH_DEV_BIND(name) all = ((a.name = reinterpret_cast<decltype(a.name)>(dlsym(h, #name))) != nullptr) && all
        FGH_DEV_BIND(fg_last_error);
        FGH_DEV_BIND(fg_index_upload);
        FGH_DEV_BIND(fg_index_release);
        FGH_DEV_BIND(fg_index_with_alive);
        FGH_DEV_BIND(fg_index_append);
        FGH_DEV_BIND(fg_index_term_info);
        FGH_DEV_BIND(fg_search_batch);
        FGH_DEV_BIND(fg_search_union_of);
        FGH_DEV_BIND(fg_search_union_of_filtered);
        FGH_DEV_BIND(fg_batch_prepare_ex);
        FGH_DEV_BIND(fg_batch_query_status);
        FGH_DEV_BIND(fg_batch_submit);
        FGH_DEV_BIND(fg_batch_submit_sharded);
        FGH_DEV_BIND(fg_comm_info);
        FGH_DEV_BIND(fg_comm_allgather_bytes);
        FGH_DEV_BIND(fg_batch_collect);
        FGH_DEV_BIND(fg_batch_release);
#undef FGH_DEV_BIND
        a.ok = all;
        if (!all) a.why = "the device library does not export the fugu_gpu.h entry points this host library needs (version mismatch)";
        return a;
    }();
    return api;
}
// status of a device-library call: on failure its message becomes this library's (fgh_last_error)
int32_t D(int32_t rc) {
    if (rc != FG_OK) {
        const char* m = dev_api().fg_last_error ? dev_api().fg_last_error() : "";
        snprintf(g_herr, sizeof(g_herr), "%s", m ? m : "");
    }
    return rc;
}
#define DEV_OR_FAIL()                                                                  \
    do {                                                                               \
        if (!dev_api().ok) return host_fail(FG_ERR_NO_DEVICE, "%s", dev_api().why.c_str()); \
    } while (0)

// tantivy's fieldnorm code (Lucene SmallFloat, 24 exact values, then 3 mantissa bits per
// power of two): the largest id whose decoded value is <= n. fg_fieldnorm_to_id of the device library
// computes the same table; it is repeated here so that building documents needs no device library.
uint8_t fieldnorm_id(uint32_t n) {
    static const std::vector<uint64_t> table = [] {
        std::vector<uint64_t> t(256);
        for (uint32_t id = 0; id < 256; id++) {
            if (id < 24) { t[id] = id; continue; }
            const uint32_t m = (id - 24) & 7u, e = (id - 24) >> 3;
            t[id] = 24u + (e == 0 ? (uint64_t)m : ((uint64_t)(m | 8u) << (e - 1)));
        }
        return t;
    }();
    return (uint8_t)((std::upper_bound(table.begin(), table.end(), (uint64_t)n) - table.begin()) - 1);
}

// ------------------------------------------------------------------------------------------
// A.1 default analyzer: SimpleTokenizer -> RemoveLongFilter(40) -> LowerCaser
// ------------------------------------------------------------------------------------------
inline bool decode_utf8(const unsigned char* s, size_t n, size_t& i, uint32_t& cp) {
    const unsigned char c = s[i];
    if (c < 0x80) { cp = c; i += 1; return true; }
    int len = (c >> 5) == 0x6 ? 2 : (c >> 4) == 0xE ? 3 : (c >> 3) == 0x1E ? 4 : 0;
    if (!len || i + len > n) { cp = 0xFFFD; i += 1; return false; }
    cp = c & (0xFF >> (len + 1));
    for (int k = 1; k < len; k++) {
        if ((s[i + k] & 0xC0) != 0x80) { cp = 0xFFFD; i += 1; return false; }
        cp = (cp << 6) | (s[i + k] & 0x3F);
    }
    i += len;
    return true;
}
inline bool is_alnum_cp(uint32_t cp) {
    if (cp < 0x80) return (cp >= '0' && cp <= '9') || (cp >= 'a' && cp <= 'z') || (cp >= 'A' && cp <= 'Z');
    int lo = 0, hi = fg::kNumAlnumRanges - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        if (cp < fg::kAlnumRanges[mid].lo) hi = mid - 1;
        else if (cp > fg::kAlnumRanges[mid].hi) lo = mid + 1;
        else return true;
    }
    return false;
}
inline void append_utf8(std::string& o, uint32_t cp) {
    if (cp < 0x80) o.push_back((char)cp);
    else if (cp < 0x800) { o.push_back((char)(0xC0 | (cp >> 6))); o.push_back((char)(0x80 | (cp & 0x3F))); }
    else if (cp < 0x10000) { o.push_back((char)(0xE0 | (cp >> 12))); o.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); o.push_back((char)(0x80 | (cp & 0x3F))); }
    else { o.push_back((char)(0xF0 | (cp >> 18))); o.push_back((char)(0x80 | ((cp >> 12) & 0x3F))); o.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); o.push_back((char)(0x80 | (cp & 0x3F))); }
}
inline void append_lower(std::string& o, uint32_t cp) {
    if (cp < 0x80) { o.push_back((char)((cp >= 'A' && cp <= 'Z') ? cp + 32 : cp)); return; }
    int lo = 0, hi = fg::kNumLowerMap - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        if (cp < fg::kLowerMap[mid].cp) hi = mid - 1;
        else if (cp > fg::kLowerMap[mid].cp) lo = mid + 1;
        else { o += fg::kLowerMap[mid].utf8; return; }
    }
    append_utf8(o, cp);
}
void tokenize(const char* text, size_t n, std::vector<std::string>& out) {
    const unsigned char* s = (const unsigned char*)text;
    size_t i = 0;
    while (i < n) {
        // skip separators
        size_t j = i;
        uint32_t cp;
        decode_utf8(s, n, j, cp);
        if (!is_alnum_cp(cp)) { i = j; continue; }
        // maximal alphanumeric run [i, e)
        size_t e = j;
        while (e < n) {
            size_t k = e;
            decode_utf8(s, n, k, cp);
            if (!is_alnum_cp(cp)) break;
            e = k;
        }
        if (e - i < 40) {  // RemoveLongFilter::limit(40): keep tokens with len_bytes < 40 (before lowercasing)
            std::string tok;
            tok.reserve(e - i);
            size_t k = i;
            while (k < e) { decode_utf8(s, e, k, cp); append_lower(tok, cp); }
            out.push_back(std::move(tok));
        }
        i = e;
    }
}

// ------------------------------------------------------------------------------------------
// facet paths: Facet::from_text("/a/b/c") -> one term per ancestor ("/a", "/a/b", "/a/b/c")
// ------------------------------------------------------------------------------------------
bool facet_segments(const std::string& path, std::vector<std::string>& segs) {
    if (path.empty() || path[0] != '/') return false;
    std::string cur;
    bool esc = false;
    for (size_t i = 1; i < path.size(); i++) {
        char c = path[i];
        if (esc) { cur.push_back(c); esc = false; continue; }
        if (c == '\\') { esc = true; continue; }
        if (c == '/') { segs.push_back(cur); cur.clear(); continue; }
        cur.push_back(c);
    }
    if (path.size() > 1) segs.push_back(cur);
    return true;
}
std::string facet_key(const std::vector<std::string>& segs, size_t n) {
    std::string k;
    for (size_t i = 0; i < n; i++) { k.push_back('/'); k += segs[i]; }
    return k;
}
// normalize_facet_path, src/db/search.rs:594-600
std::string normalize_facet_path(const std::string& p) { return (!p.empty() && p[0] == '/') ? p : "/" + p; }

// ------------------------------------------------------------------------------------------
// A.2 query grammar (tantivy-query-grammar 0.24 subset) -> user AST
// ------------------------------------------------------------------------------------------
enum Occ : int { O_NONE = -1, O_SHOULD = 0, O_MUST = 1, O_NOT = 2 };
struct Ast {
    enum Kind { LEAF, ALL, CLAUSE } kind = LEAF;
    std::string field;   // LEAF: explicit field or ""
    std::string text;    // LEAF: literal
    bool quoted = false;
    float boost = 1.f;
    std::vector<std::pair<int, Ast>> kids;  // CLAUSE: (occur or O_NONE, child)
};
struct ParseError { std::string msg; bool unsupported = false; };

struct Parser {
    const std::string& s;
    size_t i = 0;
    explicit Parser(const std::string& q) : s(q) {}
    void ws() { while (i < s.size() && isspace((unsigned char)s[i])) i++; }
    bool eof() const { return i >= s.size(); }
    static bool special(char c) { return strchr("()[]{}\":^~", c) != nullptr; }
    bool keyword(const char* kw) {
        size_t n = strlen(kw);
        if (s.compare(i, n, kw) != 0) return false;
        if (i + n < s.size() && !isspace((unsigned char)s[i + n]) && s[i + n] != '(') return false;
        return true;
    }
    std::string word() {
        std::string w;
        while (i < s.size() && !isspace((unsigned char)s[i]) && !special(s[i])) {
            if (s[i] == '\\' && i + 1 < s.size()) { w.push_back(s[i + 1]); i += 2; continue; }
            w.push_back(s[i++]);
        }
        return w;
    }
    void boost(Ast& a) {
        if (i < s.size() && s[i] == '^') {
            size_t j = i + 1;
            char* end = nullptr;
            float b = strtof(s.c_str() + j, &end);
            if (end == s.c_str() + j) throw ParseError{"expected a number after '^'"};
            a.boost *= b;
            i = end - s.c_str();
        }
    }
    Ast leaf() {
        ws();
        if (eof()) throw ParseError{"unexpected end of query"};
        Ast a;
        char c = s[i];
        if (c == '(') {
            i++;
            a = seq(true);
            ws();
            if (eof() || s[i] != ')') throw ParseError{"missing ')'"};
            i++;
            boost(a);
            return a;
        }
        if (c == '[' || c == '{') throw ParseError{"range queries are not evaluated on the device", true};
        if (c == ')' || c == ']' || c == '}' || c == '^' || c == '~' || c == ':') throw ParseError{std::string("unexpected '") + c + "'"};
        if (c == '*' && (i + 1 == s.size() || isspace((unsigned char)s[i + 1]) || s[i + 1] == ')' || s[i + 1] == '^')) {
            i++;
            a.kind = Ast::ALL;
            boost(a);  // `*^2`: a boosted AllQuery (every document scores the boost)
            return a;
        }
        if (c == '"') {
            size_t j = s.find('"', i + 1);
            if (j == std::string::npos) throw ParseError{"unbalanced '\"'"};
            a.text = s.substr(i + 1, j - i - 1);
            a.quoted = true;
            i = j + 1;
            if (i < s.size() && s[i] == '~') throw ParseError{"phrase slop is not evaluated on the device", true};
            boost(a);
            return a;
        }
        std::string w = word();
        if (w.empty()) throw ParseError{"empty term"};
        if (i < s.size() && s[i] == ':') {  // field:literal
            i++;
            if (eof() || isspace((unsigned char)s[i])) throw ParseError{"missing value after ':'"};
            Ast v = leaf();
            if (v.kind != Ast::LEAF) {
                if (v.kind == Ast::CLAUSE) throw ParseError{"field:(...) groups are not supported", true};
                throw ParseError{"field:* is not supported", true};
            }
            v.field = w;
            return v;
        }
        if (i < s.size() && s[i] == '~') throw ParseError{"fuzzy terms are not evaluated on the device", true};
        a.text = w;
        boost(a);
        return a;
    }
    // (occur, leaf): '+', '-' prefixes and the NOT keyword
    std::pair<int, Ast> occur_leaf() {
        ws();
        int occ = O_NONE;
        if (!eof() && (s[i] == '+' || s[i] == '-') && i + 1 < s.size() && !isspace((unsigned char)s[i + 1])) {
            occ = s[i] == '+' ? O_MUST : O_NOT;
            i++;
        } else if (keyword("NOT")) {
            i += 3;
            occ = O_NOT;
        }
        Ast a = leaf();
        return {occ, std::move(a)};
    }
    // one item of the top-level sequence: a single occur_leaf or a chain `x AND y OR z`
    std::pair<int, Ast> item() {
        auto first = occur_leaf();
        std::vector<std::vector<std::pair<int, Ast>>> dnf;  // aggregate_binary_expressions
        dnf.emplace_back();
        dnf.back().push_back(std::move(first));
        bool chain = false;
        while (true) {
            size_t save = i;
            ws();
            bool is_and = keyword("AND"), is_or = !is_and && keyword("OR");
            if (!is_and && !is_or) { i = save; break; }
            i += is_and ? 3 : 2;
            ws();
            if (eof()) throw ParseError{"dangling boolean operator"};
            auto nx = occur_leaf();
            chain = true;
            if (is_and) dnf.back().push_back(std::move(nx));
            else { dnf.emplace_back(); dnf.back().push_back(std::move(nx)); }
        }
        if (!chain) return std::move(dnf[0][0]);
        auto make_and = [](std::vector<std::pair<int, Ast>>& g) -> std::pair<int, Ast> {
            if (g.size() == 1) return std::move(g[0]);
            Ast c;
            c.kind = Ast::CLAUSE;
            for (auto& x : g) c.kids.emplace_back(x.first == O_NONE ? O_MUST : x.first, std::move(x.second));
            return {O_NONE, std::move(c)};
        };
        if (dnf.size() == 1) return make_and(dnf[0]);
        Ast c;
        c.kind = Ast::CLAUSE;
        for (auto& g : dnf) {
            auto x = make_and(g);
            c.kids.emplace_back(x.first == O_NONE ? O_SHOULD : x.first, std::move(x.second));
        }
        return {O_NONE, std::move(c)};
    }
    Ast seq(bool in_paren) {
        Ast c;
        c.kind = Ast::CLAUSE;
        while (true) {
            ws();
            if (eof()) break;
            if (s[i] == ')') { if (in_paren) break; throw ParseError{"unbalanced ')'"}; }
            c.kids.push_back(item());
        }
        if (c.kids.empty()) throw ParseError{"empty query"};
        if (c.kids.size() == 1 && c.kids[0].first == O_NONE) return std::move(c.kids[0].second);
        return c;
    }
};

// escape_query_string, src/db/search.rs:603-610
std::string escape_query_string(const std::string& q) {
    std::string o;
    for (char c : q)
        if (!strchr("()[]{}\":+-!~*?\\^", c)) o.push_back(c);
    return o;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// dataset
// ------------------------------------------------------------------------------------------
// string -> term ordinal, open addressing over a byte pool (lookups take (ptr, len): no allocation)
struct TermDict {
    struct Ent { uint64_t off; uint32_t len, ord; };
    std::vector<char> pool;
    std::vector<Ent> ents;
    std::vector<uint32_t> table;  // index into ents + 1, 0 = empty
    static uint64_t hash(const char* s, size_t n) {
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < n; i++) { h ^= (unsigned char)s[i]; h *= 1099511628211ull; }
        return h ^ (h >> 29);
    }
    void reserve(size_t n) { if (table.size() < 2 * n + 16) rehash(2 * n + 16); }
    void rehash(size_t want) {
        size_t cap = 16;
        while (cap < want) cap <<= 1;
        table.assign(cap, 0);
        for (uint32_t i = 0; i < ents.size(); i++) place(i);
    }
    void place(uint32_t i) {
        size_t m = table.size() - 1, h = hash(pool.data() + ents[i].off, ents[i].len) & m;
        while (table[h]) h = (h + 1) & m;
        table[h] = i + 1;
    }
    uint32_t find(const char* s, size_t n) const {
        if (table.empty()) return FG_TERM_MISSING;
        size_t m = table.size() - 1, h = hash(s, n) & m;
        while (table[h]) {
            const Ent& e = ents[table[h] - 1];
            if (e.len == n && memcmp(pool.data() + e.off, s, n) == 0) return e.ord;
            h = (h + 1) & m;
        }
        return FG_TERM_MISSING;
    }
    // find() with the hash already computed (one word is looked up in several dictionaries), and a prefetch of the
    // slot the probe will start at: the planner issues the prefetches of all words of a query before it resolves any
    // (a lookup is three dependent cache misses: table slot, entry, pooled string)
    void prefetch_slot(uint64_t hv) const {
        if (!table.empty()) __builtin_prefetch(&table[hv & (table.size() - 1)]);
    }
    void prefetch_entry(uint64_t hv) const {
        if (table.empty()) return;
        const uint32_t t = table[hv & (table.size() - 1)];
        if (t) __builtin_prefetch(&ents[t - 1]);
    }
    uint32_t find_hashed(uint64_t hv, const char* s, size_t n) const {
        if (table.empty()) return FG_TERM_MISSING;
        size_t m = table.size() - 1, h = hv & m;
        while (table[h]) {
            const Ent& e = ents[table[h] - 1];
            if (e.len == n && memcmp(pool.data() + e.off, s, n) == 0) return e.ord;
            h = (h + 1) & m;
        }
        return FG_TERM_MISSING;
    }
    void insert(const char* s, size_t n, uint32_t ord) {
        if ((ents.size() + 1) * 2 > table.size()) rehash((ents.size() + 1) * 4);
        ents.push_back({(uint64_t)pool.size(), (uint32_t)n, ord});
        pool.insert(pool.end(), s, s + n);
        place((uint32_t)ents.size() - 1);
    }
    size_t size() const { return ents.size(); }
};

struct FieldBuild {
    TermDict dict;
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> postings;  // term -> (doc, tf)
    uint64_t total_tokens = 0;
    std::vector<uint32_t> doc_len;
    uint32_t term(const std::string& t) {
        uint32_t o = dict.find(t.data(), t.size());
        if (o != FG_TERM_MISSING) return o;
        o = (uint32_t)postings.size();
        dict.insert(t.data(), t.size(), o);
        postings.emplace_back();
        return o;
    }
};

struct fgh_dataset {
    fg_ctx* ctx = nullptr;
    // current device snapshot. Searches take a reference for the duration of the call, so a concurrent
    // commit can swap the snapshot without pulling it from under them (a tantivy Searcher behaves the
    // same way: it keeps the segments it was opened on).
    std::shared_ptr<fg_index> index;
    uint32_t committed_docs = 0;  // n_docs of the snapshot in `index`
    uint32_t full_docs = 0;       // n_docs of the last snapshot built by a full upload (appends since: committed_docs - full_docs)
    uint64_t version = 0;         // bumped by every upsert / delete (a commit detects changes made while it was building)
    std::mutex commit_mu;         // commits serialize among themselves (they hold `mu` only to capture and to publish)
    uint64_t n_appends = 0, n_full_uploads = 0;
    // dictionary entries the current snapshot knows, per field: ordinals are assigned in insertion order, so
    // a term first seen after the last commit has an ordinal >= this and is planned as MISSING (an
    // uncommitted document is invisible to the reference's searcher too). A dataset without a device
    // context never commits: its dictionary is its state and everything in it is visible.
    uint32_t committed_terms[3] = {0, 0, 0};
    // writers (upsert, delete, commit, adopt) exclusive; planners, dictionary walks and snapshot grabs shared
    mutable std::shared_mutex mu;
    uint32_t lookup(uint32_t field, const char* s, size_t n) const {
        const uint32_t o = f[field].dict.find(s, n);
        return (o != FG_TERM_MISSING && ctx && o >= committed_terms[field]) ? FG_TERM_MISSING : o;
    }
    uint32_t lookup_hashed(uint32_t field, uint64_t hv, const char* s, size_t n) const {
        const uint32_t o = f[field].dict.find_hashed(hv, s, n);
        return (o != FG_TERM_MISSING && ctx && o >= committed_terms[field]) ? FG_TERM_MISSING : o;
    }
    FieldBuild f[3];
    std::vector<std::string> ids;
    std::unordered_map<std::string, uint32_t> id2doc;
    std::vector<uint8_t> alive;  // per doc
    bool adopted = false;
    bool dirty = false;
    uint32_t n_docs = 0;
    uint32_t doc_base = 0;
};

extern "C" const char* fgh_last_error(void) { return g_herr; }

extern "C" int32_t fgh_dataset_create(fg_ctx* ctx, fgh_dataset** out) {
    if (!out) return host_fail(FG_ERR_INVALID, "fgh_dataset_create: out is NULL");
    *out = new fgh_dataset();
    (*out)->ctx = ctx;
    return FG_OK;
}
extern "C" void fgh_dataset_destroy(fgh_dataset* ds) {
    if (!ds) return;
    delete ds;  // releases the snapshot unless a search still holds it
}
extern "C" uint32_t fgh_dataset_num_docs(const fgh_dataset* ds) { return ds ? ds->n_docs : 0; }
extern "C" fg_index* fgh_dataset_index(fgh_dataset* ds) {
    if (!ds) return nullptr;
    std::shared_lock<std::shared_mutex> g(ds->mu);
    return ds->index.get();
}
static std::shared_ptr<fg_index> current_snapshot(fgh_dataset* ds) {
    std::shared_lock<std::shared_mutex> g(ds->mu);
    return ds->index;
}

extern "C" int32_t fgh_tokenize(const char* text, char* buf, uint32_t cap) {
    if (!text || !buf) return -1;
    std::vector<std::string> t;
    tokenize(text, strlen(text), t);
    size_t pos = 0;
    for (auto& x : t) {
        if (pos + x.size() + 1 > cap) return -1;
        memcpy(buf + pos, x.c_str(), x.size() + 1);
        pos += x.size() + 1;
    }
    return (int32_t)t.size();
}

static void delete_doc_locked(fgh_dataset* ds, const std::string& id) {
    auto it = ds->id2doc.find(id);
    if (it != ds->id2doc.end()) {
        ds->alive[it->second] = 0;
        ds->id2doc.erase(it);
        ds->dirty = true;
        ds->version++;
    }
}

extern "C" int32_t fgh_dataset_delete(fgh_dataset* ds, const char* id) {
    if (!ds || !id) return host_fail(FG_ERR_INVALID, "NULL argument");
    std::unique_lock<std::shared_mutex> g(ds->mu);
    if (ds->adopted) return host_fail(FG_ERR_UNSUPPORTED, "adopted datasets are immutable");
    delete_doc_locked(ds, id);
    return FG_OK;
}

extern "C" int32_t fgh_dataset_upsert(fgh_dataset* ds, const char* id, const char* text, const char* name,
                                      const char* const* facets, uint32_t n_facets) {
    if (!ds || !id || !text) return host_fail(FG_ERR_INVALID, "fgh_dataset_upsert: NULL argument");
    // ObjectRecord::validate, src/object.rs:31-78
    const size_t idl = strlen(id), tl = strlen(text);
    if (idl == 0) return host_fail(FG_ERR_INVALID, "Object ID cannot be empty");
    if (idl > 256) return host_fail(FG_ERR_INVALID, "Object ID too long (max 256 characters)");
    if (tl == 0) return host_fail(FG_ERR_INVALID, "Object text cannot be empty");
    if (tl > 10000) return host_fail(FG_ERR_INVALID, "Text too long (max 10000 characters)");
    if (n_facets > 100) return host_fail(FG_ERR_INVALID, "Too many facets (max 100 per object)");
    for (uint32_t i = 0; i < n_facets; i++) {
        if (!facets[i] || !facets[i][0]) return host_fail(FG_ERR_INVALID, "Facet at index %u cannot be empty", i);
        if (strlen(facets[i]) > 512) return host_fail(FG_ERR_INVALID, "Facet at index %u too long (max 512 characters)", i);
    }
    std::unique_lock<std::shared_mutex> g(ds->mu);
    if (ds->adopted) return host_fail(FG_ERR_UNSUPPORTED, "adopted datasets are immutable");
    delete_doc_locked(ds, id);  // delete_term(id) then add_document, src/db/document.rs:38-48
    const uint32_t doc = ds->n_docs++;
    ds->ids.emplace_back(id);
    ds->id2doc[id] = doc;
    ds->alive.push_back(1);
    auto index_text = [&](FieldBuild& fb, const char* s) {
        std::vector<std::string> toks;
        if (s) tokenize(s, strlen(s), toks);
        fb.doc_len.push_back((uint32_t)toks.size());
        fb.total_tokens += toks.size();
        std::sort(toks.begin(), toks.end());
        for (size_t i = 0; i < toks.size();) {
            size_t j = i + 1;
            while (j < toks.size() && toks[j] == toks[i]) j++;
            fb.postings[fb.term(toks[i])].emplace_back(doc, (uint32_t)(j - i));
            i = j;
        }
    };
    index_text(ds->f[FGH_FIELD_TEXT], text);
    index_text(ds->f[FGH_FIELD_NAME], name);
    {
        FieldBuild& fb = ds->f[FGH_FIELD_FACET];
        std::vector<std::string> keys;
        for (uint32_t i = 0; i < n_facets; i++) {
            std::vector<std::string> segs;
            if (!facet_segments(normalize_facet_path(facets[i]), segs)) continue;  // Facet::from_text Err -> skipped (document.rs:322-329)
            for (size_t n = 1; n <= segs.size(); n++) keys.push_back(facet_key(segs, n));
        }
        fb.doc_len.push_back((uint32_t)keys.size());
        fb.total_tokens += keys.size();
        std::sort(keys.begin(), keys.end());
        keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
        for (auto& k : keys) fb.postings[fb.term(k)].emplace_back(doc, 1u);
    }
    ds->dirty = true;
    ds->version++;
    return FG_OK;
}

extern "C" int32_t fgh_dataset_commit(fgh_dataset* ds) {
    if (!ds) return host_fail(FG_ERR_INVALID, "NULL dataset");
    // Three phases, so that searches (and upserts) are not held up by the device work of a commit -- tantivy's commit does
    // not stall its searchers either: (1) under the dataset's lock, capture what the new snapshot will contain (the CSR of
    // the new segment / of everything, the alive bits); (2) without the lock, build the snapshot on the device; (3) under
    // the lock again, publish it. Commits serialize among themselves. Documents upserted during (2) stay invisible
    // (their doc ids and term ordinals lie beyond the captured counts) and keep the dataset dirty.
    std::lock_guard<std::mutex> one_commit(ds->commit_mu);
    std::unique_lock<std::shared_mutex> g(ds->mu);
    if (ds->adopted) return FG_OK;
    if (!ds->ctx) return host_fail(FG_ERR_NO_DEVICE, "dataset has no device context (planning only)");
    if (ds->index && !ds->dirty) return FG_OK;  // nothing changed since the last commit
    const uint32_t n_docs = ds->n_docs;
    const uint64_t version = ds->version;
    const std::shared_ptr<fg_index> base = ds->index;
    std::vector<uint32_t> alive((n_docs + 31) / 32, 0);
    bool any_dead = false;
    for (uint32_t d = 0; d < n_docs; d++) {
        if (ds->alive[d]) alive[d >> 5] |= 1u << (d & 31);
        else any_dead = true;
    }
    const bool incremental = base && !getenv("FG_NO_INCREMENTAL");
    enum { ALIVE_ONLY, APPEND, FULL } mode = FULL;
    if (incremental && ds->committed_docs == n_docs) mode = ALIVE_ONLY;  // only deletes since the last commit
    // New documents since the last commit: hand over only them, as one new segment (fg_index_append: the postings
    // already in HBM stay there). Like tantivy's merge policy, a full rebuild follows once the appended part has
    // outgrown the part that was built whole (it re-decides which terms own tf columns / membership bitmaps).
    else if (incremental && ds->committed_docs && n_docs > ds->committed_docs && (uint64_t)n_docs <= 2ull * ds->full_docs) mode = APPEND;
    struct Csr { std::vector<uint64_t> off; std::vector<uint32_t> docs, tfs; std::vector<uint8_t> fn; };
    Csr c[3];
    fg_field_desc fd[3];
    memset(fd, 0, sizeof(fd));
    const uint32_t d0 = mode == APPEND ? ds->committed_docs : 0u, ns = n_docs - d0;  // docs [d0, n_docs) are handed over
    if (mode != ALIVE_ONLY) {
        for (int f = 0; f < 3; f++) {
            FieldBuild& fb = ds->f[f];
            c[f].off.resize(fb.postings.size() + 1);
            uint64_t n = 0;
            for (size_t t = 0; t < fb.postings.size(); t++) {
                c[f].off[t] = n;
                const auto& pl = fb.postings[t];
                auto it = d0 ? std::lower_bound(pl.begin(), pl.end(), std::make_pair(d0, 0u)) : pl.begin();
                for (; it != pl.end(); ++it) {
                    c[f].docs.push_back(it->first - d0);
                    c[f].tfs.push_back(it->second);
                    n++;
                }
            }
            c[f].off[fb.postings.size()] = n;
            uint64_t tokens = 0;
            for (uint32_t d = d0; d < n_docs; d++) tokens += fb.doc_len[d];
            fd[f].n_terms = (uint32_t)fb.postings.size();
            fd[f].total_num_tokens = tokens;  // (of the docs handed over: the whole corpus for a full upload)
            fd[f].term_offsets = c[f].off.data();
            fd[f].doc_ids = c[f].docs.data();
            if (f != (int)FGH_FIELD_FACET) {
                c[f].fn.resize(ns);
                for (uint32_t d = 0; d < ns; d++) c[f].fn[d] = fieldnorm_id(fb.doc_len[d0 + d]);
                fd[f].flags = FG_FIELD_HAS_FIELDNORMS | FG_FIELD_HAS_FREQS;
                fd[f].fieldnorm_ids = c[f].fn.data();
                fd[f].term_freqs = c[f].tfs.data();
            }
        }
    }
    g.unlock();

    // ---- device work, no lock held ----
    DEV_OR_FAIL();
    fg_index* nx = nullptr;
    int32_t rc;
    if (mode == ALIVE_ONLY) {
        rc = D(dev_api().fg_index_with_alive(base.get(), any_dead ? alive.data() : nullptr, &nx));
    } else {
        fg_index_desc desc;
        memset(&desc, 0, sizeof(desc));
        desc.n_docs = ns;
        desc.n_fields = 3;
        desc.fields = fd;
        if (mode == APPEND) {
            rc = D(dev_api().fg_index_append(base.get(), &desc, any_dead ? alive.data() : nullptr, &nx));
        } else {
            desc.alive_bitset = any_dead ? alive.data() : nullptr;
            rc = D(dev_api().fg_index_upload(ds->ctx, &desc, &nx));
        }
    }
    if (rc) return rc;

    // ---- publish ----
    g.lock();
    ds->index.reset(nx, dev_api().fg_index_release);
    if (mode != ALIVE_ONLY) {
        ds->committed_docs = n_docs;
        for (int f = 0; f < 3; f++) ds->committed_terms[f] = fd[f].n_terms;
        if (mode == APPEND) ds->n_appends++;
        else { ds->full_docs = n_docs; ds->n_full_uploads++; }
    }
    ds->dirty = ds->version != version;  // something was upserted / deleted while the snapshot was being built
    return FG_OK;
}

extern "C" int32_t fgh_dataset_commit_counts(const fgh_dataset* ds, uint64_t* n_full_uploads, uint64_t* n_appends) {
    if (!ds) return host_fail(FG_ERR_INVALID, "NULL dataset");
    std::shared_lock<std::shared_mutex> g(ds->mu);
    if (n_full_uploads) *n_full_uploads = ds->n_full_uploads;
    if (n_appends) *n_appends = ds->n_appends;
    return FG_OK;
}

extern "C" int32_t fgh_dataset_adopt(fgh_dataset* ds, const fg_index_desc* desc, const char* const* terms,
                                     const uint64_t* terms_bytes) {
    if (!ds || !desc) return host_fail(FG_ERR_INVALID, "NULL argument");
    if (desc->n_fields > 3) return host_fail(FG_ERR_INVALID, "adopt expects fields [text, name, facet]");
    std::unique_lock<std::shared_mutex> g(ds->mu);
    if (ds->ctx) {
        fg_index* nx = nullptr;
        DEV_OR_FAIL();
        int32_t rc = D(dev_api().fg_index_upload(ds->ctx, desc, &nx));
        if (rc) return rc;
        ds->index.reset(nx, dev_api().fg_index_release);
    }
    for (uint32_t f = 0; f < desc->n_fields; f++) {
        ds->f[f] = FieldBuild();
        if (!terms || !terms[f]) continue;
        const char* p = terms[f];
        const char* end = p + terms_bytes[f];
        ds->f[f].dict.reserve(desc->fields[f].n_terms);
        for (uint32_t t = 0; t < desc->fields[f].n_terms && p < end; t++) {
            size_t l = strlen(p);
            ds->f[f].dict.insert(p, l, t);
            p += l + 1;
        }
    }
    for (uint32_t f = 0; f < 3; f++) ds->committed_terms[f] = f < desc->n_fields ? desc->fields[f].n_terms : 0;
    ds->adopted = true;
    ds->n_docs = desc->n_docs;
    ds->doc_base = desc->doc_id_base;
    return FG_OK;
}

extern "C" int32_t fgh_dataset_doc_id(const fgh_dataset* ds, uint32_t doc, char* buf, uint32_t cap) {
    if (!ds || !buf) return -1;
    if (ds->adopted) {  // synthetic corpora: "d%09u" 
        int n = snprintf(buf, cap, "d%09u", doc);
        return n < (int)cap ? n : -1;
    }
    if (doc >= ds->ids.size() || ds->ids[doc].size() + 1 > cap) return -1;
    memcpy(buf, ds->ids[doc].c_str(), ds->ids[doc].size() + 1);
    return (int32_t)ds->ids[doc].size();
}

extern "C" uint32_t fgh_dataset_term_ord(const fgh_dataset* ds, uint32_t field, const char* token) {
    if (!ds || field > 2 || !token) return FG_TERM_MISSING;
    return ds->lookup(field, token, strlen(token));
}

// ------------------------------------------------------------------------------------------
// planning: user AST -> one level of clauses (what the device evaluates)
// ------------------------------------------------------------------------------------------
namespace {

struct Group {                 // OR-ed leaves whose scores add up (word -> fields, facet group)
    std::vector<fg_leaf> leaves;
    bool all = false;          // AllQuery
    float all_boost = 0.f;
};
struct LNode {                 // logical AST after field expansion / tokenisation
    bool is_group = true;
    Group g;
    std::vector<std::pair<int, LNode>> kids;
};

struct Planner {
    const fgh_dataset* ds;
    explicit Planner(const fgh_dataset* d) : ds(d) {}

    int field_of(const std::string& name) const {
        if (name == "text") return FGH_FIELD_TEXT;
        if (name == "name") return FGH_FIELD_NAME;
        return -1;
    }
    void term_leaf(Group& g, uint32_t field, const std::string& literal, float boost) const {
        std::vector<std::string> toks;
        tokenize(literal.c_str(), literal.size(), toks);
        if (toks.empty()) return;  // the analyzer dropped everything: no query for this field
        if (toks.size() > 1) throw ParseError{"a literal that tokenises to several tokens is a PhraseQuery (needs positions)", true};
        fg_leaf l;
        l.field = field;
        l.term_ord = fgh_dataset_term_ord(ds, field, toks[0].c_str());
        l.boost = boost;
        g.leaves.push_back(l);
    }
    LNode lower(const Ast& a, float boost) const {
        LNode n;
        if (a.kind == Ast::ALL) {
            n.g.all = true;
            n.g.all_boost = boost * a.boost;
            return n;
        }
        if (a.kind == Ast::LEAF) {
            const float b = boost * a.boost;
            if (!a.field.empty()) {
                int f = field_of(a.field);
                if (f < 0) {
                    if (a.field == "id" || a.field == "namespace" || a.field == "organization" || a.field == "conversation_id" ||
                        a.field == "data_type" || a.field == "facet" || a.field == "metadata" || a.field.rfind("date_", 0) == 0)
                        throw ParseError{"field '" + a.field + "' is not loaded on the device", true};
                    throw ParseError{"Field does not exist: '" + a.field + "'"};
                }
                term_leaf(n.g, (uint32_t)f, a.text, b);
            } else {  // default fields [text, name], src/db/search.rs:108-112
                term_leaf(n.g, FGH_FIELD_TEXT, a.text, b);
                term_leaf(n.g, FGH_FIELD_NAME, a.text, b);
            }
            return n;
        }
        n.is_group = false;
        for (auto& k : a.kids) n.kids.emplace_back(k.first == O_NONE ? O_SHOULD : k.first, lower(k.second, boost * a.boost));
        return n;
    }
};

struct FlatClause { int occ; Group g; };

bool all_should_groups(const LNode& n) {
    if (n.is_group) return true;
    for (auto& k : n.kids)
        if (k.first != O_SHOULD || !all_should_groups(k.second)) return false;
    return true;
}
void collect_groups(const LNode& n, Group& into) {
    if (n.is_group) {
        into.leaves.insert(into.leaves.end(), n.g.leaves.begin(), n.g.leaves.end());
        if (n.g.all) { into.all = true; into.all_boost += n.g.all_boost; }
        return;
    }
    for (auto& k : n.kids) collect_groups(k.second, into);
}
// Lift `node` (a boolean node seen under occur `occ`) into the flat clause list.
void flatten(const LNode& n, int occ, std::vector<FlatClause>& out, int depth) {
    if (n.is_group) { out.push_back({occ, n.g}); return; }
    if (all_should_groups(n)) {  // Should-of-Should is a union whose scores add: one group
        if (occ == O_SHOULD && depth == 0) {  // keep top-level Should words as separate clauses
            for (auto& k : n.kids) flatten(k.second, O_SHOULD, out, depth + 1);
            return;
        }
        FlatClause c{occ, {}};
        collect_groups(n, c.g);
        out.push_back(std::move(c));
        return;
    }
    if (occ == O_MUST || depth == 0) {
        bool has_must = false;
        for (auto& k : n.kids) has_must = has_must || k.first == O_MUST;
        if (depth > 0 && !has_must) {
            // Must(Bool[Should b.., MustNot c..]): without a Must child at least one Should child has to match, so lifted
            // to this level the Should children are ONE Must group (a union whose scores add), not optional clauses
            FlatClause m{O_MUST, {}};
            bool any = false;
            for (auto& k : n.kids) {
                if (k.first != O_SHOULD) continue;
                if (!(k.second.is_group || all_should_groups(k.second)))
                    throw ParseError{"nested boolean (OR of AND groups / negated groups) is not evaluated on the device", true};
                collect_groups(k.second, m.g);
                any = true;
            }
            if (!any) throw ParseError{"a required group of only negated terms (matches nothing) is not evaluated on the device", true};
            out.push_back(std::move(m));
            for (auto& k : n.kids) {
                if (k.first != O_NOT) continue;
                if (!(k.second.is_group || all_should_groups(k.second)))
                    throw ParseError{"nested boolean (OR of AND groups / negated groups) is not evaluated on the device", true};
                flatten(k.second, O_NOT, out, depth + 1);
            }
            return;
        }
        // Must(Bool[Must a, Should b, MustNot c]) == Must a, Should b, MustNot c at this level
        for (auto& k : n.kids) {
            if (k.second.is_group || all_should_groups(k.second)) flatten(k.second, k.first, out, depth + 1);
            else if (k.first == O_MUST) flatten(k.second, O_MUST, out, depth + 1);
            else throw ParseError{"nested boolean (OR of AND groups / negated groups) is not evaluated on the device", true};
        }
        return;
    }
    throw ParseError{"nested boolean (OR of AND groups / negated groups) is not evaluated on the device", true};
}

// A union (every child Should) some of whose children are boolean queries of their own: each such child must flatten to
// one level by itself; the plain children (words, unions of words) together form one more child. tantivy builds exactly
// this tree -- BooleanQuery[(Should, BooleanQuery[..]), ..] -- and sums the scores of the children that match.
bool all_kids_should(const LNode& n) {
    if (n.is_group) return false;
    for (auto& k : n.kids)
        if (k.first != O_SHOULD) return false;
    return true;
}
// (a union inside the union -- `x OR (a AND b) y` parses as ((x OR (a AND b)) y) -- is the same union: its children join)
void collect_union_children(const LNode& n, std::vector<FlatClause>& plain, std::vector<std::vector<FlatClause>>& nested) {
    for (auto& k : n.kids) {
        if (k.second.is_group || all_should_groups(k.second)) flatten(k.second, O_SHOULD, plain, 1);
        else if (all_kids_should(k.second)) collect_union_children(k.second, plain, nested);
        else {
            std::vector<FlatClause> sub;
            flatten(k.second, O_SHOULD, sub, 0);  // (throws when the child itself is nested)
            nested.push_back(std::move(sub));
        }
    }
}
bool split_disjuncts(const LNode& root, std::vector<std::vector<FlatClause>>& out) {
    if (!all_kids_should(root)) return false;  // a Must / MustNot sibling of a nested group: deeper than this form
    std::vector<FlatClause> plain;
    collect_union_children(root, plain, out);
    if (out.empty()) return false;
    if (!plain.empty()) out.insert(out.begin(), std::move(plain));
    return out.size() <= 64;
}

// parse_filters + build_facet_query, src/db/search.rs:221-324
void facet_group(const fgh_dataset* ds, const char* const* filters, uint32_t n, Group& g, bool& any_term) {
    for (uint32_t i = 0; i < n; i++) {
        std::string f = filters[i] ? filters[i] : "";
        // Dataset::search drops filters of the form *x* before build_facet_query (:98-105); the
        // wildcard post-filter branch is dead because parse_filters never emits Wildcard (F8)
        if (!f.empty() && f.front() == '*' && f.back() == '*') continue;
        std::string norm = normalize_facet_path(f);
        std::string path;
        if (norm.size() >= 2 && norm.compare(norm.size() - 2, 2, "/*") == 0) path = norm.substr(0, norm.size() - 2);  // Prefix -> exact term (:273-281)
        else if (norm.find('=') != std::string::npos) path = norm.substr(0, norm.find('='));                         // a=b -> path a (:306-313)
        else path = norm;
        std::vector<std::string> segs;
        if (!facet_segments(path, segs)) continue;  // Facet::from_text Err: filter silently dropped (:234,277)
        any_term = true;
        fg_leaf l;
        l.field = FGH_FIELD_FACET;
        l.term_ord = fgh_dataset_term_ord(ds, FGH_FIELD_FACET, facet_key(segs, segs.size()).c_str());
        l.boost = 1.f;
        g.leaves.push_back(l);
    }
}

int32_t plan_impl(const fgh_dataset* ds, const char* query, const char* const* filters, uint32_t n_filters,
                  uint32_t page, uint32_t per_page, fgh_plan_t* out) {
    memset(out, 0, sizeof(*out));
    const uint64_t limit = (uint64_t)page * per_page + per_page;  // src/db/search.rs:154-160
    if (limit == 0) return host_fail(FG_ERR_INVALID, "per_page == 0: TopDocs::with_limit panics on a zero limit");
    if (limit > 0xFFFFFFFFull) return host_fail(FG_ERR_INVALID, "page*per_page overflows");
    out->k = (uint32_t)limit;
    out->offset = page * per_page;
    std::string q = query ? query : "";
    bool blank = true;
    for (char c : q) if (!isspace((unsigned char)c)) blank = false;

    std::vector<FlatClause> flat;
    std::vector<std::vector<FlatClause>> disjuncts;  // nested query: the children of the top-level union (fgh_plan_t::n_disjuncts)
    Planner pl(ds);
    bool text_all = false;
    float text_all_boost = 1.f;
    try {
        if (blank) {
            text_all = true;  // AllQuery, :115-116
        } else {
            LNode root;
            try {
                Parser p(q);
                Ast a = p.seq(false);
                root = pl.lower(a, 1.f);
            } catch (ParseError& e) {
                if (e.unsupported) throw;
                // fallback: strip special characters and retry (:120-125); a second error is Err
                out->used_fallback = 1;
                std::string esc = escape_query_string(q);
                Parser p(esc);
                Ast a = p.seq(false);
                root = pl.lower(a, 1.f);
            }
            if (root.is_group && root.g.all && root.g.leaves.empty()) { text_all = true; text_all_boost = root.g.all_boost; }
            else {
                try {
                    flatten(root, O_SHOULD, flat, 0);
                } catch (ParseError& e) {
                    // not one level: a union whose children are one-level boolean queries? (`(a AND b) OR (c AND d)`)
                    flat.clear();
                    if (!e.unsupported || !split_disjuncts(root, disjuncts)) throw;
                }
            }
        }
    } catch (ParseError& e) {
        return host_fail(e.unsupported ? FG_ERR_UNSUPPORTED : FG_ERR_INVALID, "query '%s': %s", q.c_str(), e.msg.c_str());
    }

    // non-wildcard filters present -> Must-join with the facet query (:131-144)
    uint32_t n_nonwild = 0;
    for (uint32_t i = 0; i < n_filters; i++) {
        std::string f = filters[i] ? filters[i] : "";
        if (!(!f.empty() && f.front() == '*' && f.back() == '*')) n_nonwild++;
    }
    bool filter_child = false;
    if (n_nonwild && !disjuncts.empty()) {
        // Bool[Must(text_query = the union of the children), Must(facet_query)] (:140-144): the facet group has to hold for
        // every hit and to score ONCE, not once per child. It becomes a FILTER child of the union (fg_search_union_of_filtered):
        // Must(facet group) AND Must(any positive leaf of any child, boost 0) -- the second clause only bounds the child's
        // matches by the text query's (a superset of them) instead of by the size of the namespace.
        Group fg_;
        bool any = false;
        facet_group(ds, filters, n_filters, fg_, any);
        if (!any)
            return host_fail(FG_ERR_UNSUPPORTED, "query '%s': a nested boolean query combined with filters that hold no facet term (AllQuery) is not evaluated on the device", q.c_str());
        FlatClause any_text{O_MUST, {}};
        for (auto& d : disjuncts)
            for (auto& c : d) {
                if (c.g.all) return host_fail(FG_ERR_UNSUPPORTED, "query '%s': '*' inside a nested boolean query", q.c_str());
                for (fg_leaf l : c.g.leaves) {
                    if (l.boost < 0.f) return host_fail(FG_ERR_UNSUPPORTED, "query '%s': negative boosts in a nested boolean query combined with facet filters", q.c_str());
                    if (c.occ == O_NOT || l.term_ord == FG_TERM_MISSING) continue;
                    bool seen = false;
                    for (const fg_leaf& x : any_text.g.leaves) seen = seen || (x.field == l.field && x.term_ord == l.term_ord);
                    l.boost = 0.f;
                    if (!seen) any_text.g.leaves.push_back(l);
                }
            }
        for (const fg_leaf& l : fg_.leaves)
            if (l.boost < 0.f) return host_fail(FG_ERR_UNSUPPORTED, "query '%s': negative boost", q.c_str());
        std::vector<FlatClause> fchild;
        fchild.push_back({O_MUST, fg_});
        fchild.push_back(std::move(any_text));
        disjuncts.push_back(std::move(fchild));
        filter_child = true;
    } else if (n_nonwild) {
        Group fg_;
        bool any = false;
        facet_group(ds, filters, n_filters, fg_, any);
        if (text_all && !blank) {
            // `*` is a query string like any other (only query.trim().is_empty() drops the text query, :136): it parses to
            // AllQuery and is Must-joined with the facet query, so every hit scores 1.0 (times its boost) + the facet score
            flat.clear();
            FlatClause ac{O_MUST, {}};
            ac.g.all = true;
            ac.g.all_boost = text_all_boost;
            flat.push_back(std::move(ac));
        }
        if (text_all && blank) {
            // empty text query: the facet query alone (:136-138); no valid term -> AllQuery (:258-261)
            flat.clear();
            if (any) flat.push_back({O_SHOULD, fg_});
            else out->is_all = 1;
        } else {
            // Bool[(Must, text_query), (Must, facet_query)]
            std::vector<FlatClause> joined;
            bool has_must = false;
            for (auto& c : flat) if (c.occ == O_MUST) has_must = true;
            if (!has_must) {
                // Must(Bool[Should.., MustNot..]): without a Must child at least one Should has to
                // match, so the Should clauses become ONE Must group (a union whose scores add)
                FlatClause m{O_MUST, {}};
                bool any_should = false;
                for (auto& c : flat) {
                    if (c.occ != O_SHOULD) continue;
                    any_should = true;
                    m.g.leaves.insert(m.g.leaves.end(), c.g.leaves.begin(), c.g.leaves.end());
                    if (c.g.all) { m.g.all = true; m.g.all_boost += c.g.all_boost; }
                }
                if (any_should) joined.push_back(std::move(m));
                for (auto& c : flat) if (c.occ == O_NOT) joined.push_back(c);
                if (!any_should) {  // only MustNot: matches nothing, and so does the Must-join
                    joined.clear();
                    flat.clear();
                    out->n_clauses = 0;
                    out->n_leaves = 0;
                    return FG_OK;
                }
            } else {
                joined = flat;
            }
            FlatClause fc{O_MUST, {}};
            if (any) fc.g = fg_;
            else { fc.g.all = true; fc.g.all_boost = 1.f; }  // AllQuery Must: +1.0 on every hit
            joined.push_back(std::move(fc));
            flat.swap(joined);
        }
    } else if (text_all) {
        out->is_all = 1;
    }
    if (!disjuncts.empty()) {
        out->n_disjuncts = (uint32_t)disjuncts.size();
        for (size_t d = 0; d < disjuncts.size(); d++)
            for (auto& c : disjuncts[d]) {
                if (out->n_clauses >= FGH_MAX_PLAN_CLAUSES) return host_fail(FG_ERR_UNSUPPORTED, "too many clauses");
                if (c.g.all) return host_fail(FG_ERR_UNSUPPORTED, "'*' inside a nested boolean query");
                fg_clause& oc = out->clauses[out->n_clauses++];
                oc.occur = (c.occ == O_MUST ? FG_OCCUR_MUST : c.occ == O_NOT ? FG_OCCUR_MUST_NOT : FG_OCCUR_SHOULD) | ((uint32_t)(d + 1) << FGH_DISJUNCT_SHIFT);
                if (filter_child && d + 1 == disjuncts.size()) oc.occur |= FGH_FILTER_CHILD;
                oc.leaf_begin = out->n_leaves;
                for (auto& l : c.g.leaves) {
                    if (out->n_leaves >= FGH_MAX_PLAN_LEAVES) return host_fail(FG_ERR_UNSUPPORTED, "too many leaves");
                    out->leaves[out->n_leaves++] = l;
                }
                oc.n_leaves = out->n_leaves - oc.leaf_begin;
            }
        return FG_OK;
    }
    if (out->is_all) {
        out->n_clauses = 1;
        out->clauses[0] = {FG_OCCUR_MUST, 0, 1};
        out->n_leaves = 1;
        out->leaves[0] = {0, FG_TERM_ALL, text_all ? text_all_boost : 1.f};
        return FG_OK;
    }
    for (auto& c : flat) {
        if (out->n_clauses >= FGH_MAX_PLAN_CLAUSES) return host_fail(FG_ERR_UNSUPPORTED, "too many clauses");
        fg_clause& oc = out->clauses[out->n_clauses++];
        oc.occur = c.occ == O_MUST ? FG_OCCUR_MUST : c.occ == O_NOT ? FG_OCCUR_MUST_NOT : FG_OCCUR_SHOULD;
        oc.leaf_begin = out->n_leaves;
        auto push = [&](const fg_leaf& l) -> bool {
            if (out->n_leaves >= FGH_MAX_PLAN_LEAVES) return false;
            out->leaves[out->n_leaves++] = l;
            return true;
        };
        if (c.g.all) {
            if (!c.g.leaves.empty() || c.occ != O_MUST)
                return host_fail(FG_ERR_UNSUPPORTED, "'*' mixed with other terms outside a Must clause");
            if (!push({0, FG_TERM_ALL, c.g.all_boost})) return host_fail(FG_ERR_UNSUPPORTED, "too many leaves");
        }
        for (auto& l : c.g.leaves)
            if (!push(l)) return host_fail(FG_ERR_UNSUPPORTED, "too many leaves");
        oc.n_leaves = out->n_leaves - oc.leaf_begin;
    }
    return FG_OK;
}

}  // namespace

extern "C" int32_t fgh_plan(const fgh_dataset* ds, const char* query, const char* const* filters,
                            uint32_t n_filters, uint32_t page, uint32_t per_page, fgh_plan_t* out) {
    if (!ds || !out) return host_fail(FG_ERR_INVALID, "fgh_plan: NULL argument");
    std::shared_lock<std::shared_mutex> g(ds->mu);  // the dictionaries must not grow under the planner
    return plan_impl(ds, query, filters, n_filters, page, per_page, out);
}

namespace {
// Internal per-query marker of the shared planner (never leaves the library): a nested boolean query, which a sharded
// request answers apart (every shard evaluates it locally, the pages are merged on the host).
constexpr int32_t RC_NESTED = 0x4E45;
// plan n requests (multi-threaded) into one flat batch; failed plans become empty queries
struct PlannedBatch {
    std::vector<fg_query> q;
    std::vector<fg_clause> c;
    std::vector<fg_leaf> l;
    std::vector<uint32_t> offset;
    std::vector<int32_t> rc;
    uint32_t kmax = 1;
    int32_t first_err = FG_OK;
    std::string first_msg;
    // nested queries (fgh_plan_t::n_disjuncts): (index in the batch, plan); the flat batch holds an empty query in their
    // place, the search path answers them with fg_search_union_of
    std::vector<std::pair<uint32_t, fgh_plan_t>> composites;
};

// Fast path for the overwhelmingly common request shape: no filters, only ASCII alphanumeric
// words separated by blanks, optionally all joined by AND. Produces exactly what the general
// parser produces for such strings (bare words -> Should groups over [text, name]; `a AND b` ->
// Must groups) without building an AST. Returns false when the string needs the general parser.
bool plan_fast(const fgh_dataset* ds, const char* q, uint32_t page, uint32_t per_page,
               std::vector<fg_clause>& c, std::vector<fg_leaf>& l, fg_query& out, uint32_t& offset) {
    const uint64_t limit = (uint64_t)page * per_page + per_page;
    if (limit == 0 || limit > 0xFFFFFFFFull || !q) return false;
    struct W { const char* p; uint32_t n; };
    W words[16];
    int nw = 0;
    const char* p = q;
    while (*p) {
        while (*p == ' ' || *p == '\t') p++;
        if (!*p) break;
        const char* s0 = p;
        while (*p && *p != ' ' && *p != '\t') {
            const unsigned char ch = (unsigned char)*p;
            if (!((ch >= '0' && ch <= '9') || (ch >= 'a' && ch <= 'z') || (ch >= 'A' && ch <= 'Z'))) return false;
            p++;
        }
        if (nw == 16 || p - s0 >= 40) return false;
        words[nw++] = {s0, (uint32_t)(p - s0)};
    }
    if (nw == 0) return false;  // blank query = AllQuery: general path
    auto is = [](const W& w, const char* kw) { return w.n == strlen(kw) && memcmp(w.p, kw, w.n) == 0; };
    bool conj = false;
    for (int i = 0; i < nw; i++) {
        if (is(words[i], "OR") || is(words[i], "NOT")) return false;
        if (is(words[i], "AND")) conj = true;
    }
    if (conj) {  // strictly  w AND w AND w ...
        if (nw % 2 == 0) return false;
        for (int i = 0; i < nw; i++)
            if ((i % 2 == 1) != is(words[i], "AND")) return false;
    }
    const size_t c0 = c.size();
    out.k = (uint32_t)limit;
    out.clause_begin = (uint32_t)c0;
    offset = page * per_page;
    // lower-case all words and hash them first, start the table-slot loads of both dictionaries for every word, then
    // the entry loads, and only then resolve: the cache misses of a query's lookups overlap instead of queueing up
    char low[16][40];
    uint64_t hv[16];
    const int step = conj ? 2 : 1;
    const TermDict &dt = ds->f[FGH_FIELD_TEXT].dict, &dn = ds->f[FGH_FIELD_NAME].dict;
    for (int i = 0; i < nw; i += step) {
        for (uint32_t j = 0; j < words[i].n; j++) {
            const char ch = words[i].p[j];
            low[i][j] = (ch >= 'A' && ch <= 'Z') ? (char)(ch + 32) : ch;
        }
        hv[i] = TermDict::hash(low[i], words[i].n);
        dt.prefetch_slot(hv[i]);
        dn.prefetch_slot(hv[i]);
    }
    for (int i = 0; i < nw; i += step) {
        dt.prefetch_entry(hv[i]);
        dn.prefetch_entry(hv[i]);
    }
    for (int i = 0; i < nw; i += step) {
        fg_clause cl;
        cl.occur = conj ? FG_OCCUR_MUST : FG_OCCUR_SHOULD;
        cl.leaf_begin = (uint32_t)l.size();
        cl.n_leaves = 2;
        l.push_back({FGH_FIELD_TEXT, ds->lookup_hashed(FGH_FIELD_TEXT, hv[i], low[i], words[i].n), 1.f});
        l.push_back({FGH_FIELD_NAME, ds->lookup_hashed(FGH_FIELD_NAME, hv[i], low[i], words[i].n), 1.f});
        c.push_back(cl);
    }
    out.n_clauses = (uint32_t)(c.size() - c0);
    return true;
}

void plan_batch(const fgh_dataset* ds, uint32_t n, const char* const* queries, const char* const* filters,
                const uint32_t* filter_offsets, const uint32_t* pages, const uint32_t* per_pages, PlannedBatch& pb,
                bool keep_composites = false, bool mark_composites = false) {
    std::shared_lock<std::shared_mutex> dict_lock(ds->mu);  // held for the worker threads too: the dictionaries must not grow under the planner
    pb.rc.assign(n, FG_OK);
    pb.q.resize(n);
    pb.offset.assign(n, 0);
    unsigned hw = std::thread::hardware_concurrency();
    const char* e = getenv("FG_HOST_THREADS");
    if (e) hw = (unsigned)atoi(e);
    // the fast path plans ~3M requests/s per thread; a thread costs ~30 us to start
    const int T = (int)std::max(1u, std::min<unsigned>({hw ? hw : 4u, (unsigned)fg::HostPool::get().size(), 16u, n / 96 + 1}));
    struct Part { std::vector<fg_clause> c; std::vector<fg_leaf> l; std::vector<std::string> errs; uint32_t a, b;
                  std::vector<std::pair<uint32_t, fgh_plan_t>> comp; };
    std::vector<Part> parts((size_t)T);
    auto work = [&](int t) {
        Part& P = parts[t];
        P.a = (uint32_t)((uint64_t)n * t / T);
        P.b = (uint32_t)((uint64_t)n * (t + 1) / T);
        P.errs.resize(P.b - P.a);
        fgh_plan_t plan;
        for (uint32_t i = P.a; i < P.b; i++) {
            const uint32_t f0 = filter_offsets ? filter_offsets[i] : 0, f1 = filter_offsets ? filter_offsets[i + 1] : 0;
            const uint32_t page = pages ? pages[i] : 0, pp = per_pages ? per_pages[i] : 20;
            if (f1 == f0 && plan_fast(ds, queries[i], page, pp, P.c, P.l, pb.q[i], pb.offset[i])) continue;
            pb.rc[i] = plan_impl(ds, queries[i], filters ? filters + f0 : nullptr, f1 - f0, page, pp, &plan);
            if (pb.rc[i] == FG_OK && plan.n_disjuncts) {
                if (keep_composites || mark_composites) {  // answered apart (fg_search_union_of): an empty query holds its place in the flat batch
                    if (keep_composites) P.comp.emplace_back(i, plan);
                    else pb.rc[i] = RC_NESTED;  // (sharded requests: every rank re-plans it for its own shard)
                    pb.q[i].k = 1;
                    pb.q[i].clause_begin = (uint32_t)P.c.size();
                    pb.q[i].n_clauses = 0;
                    pb.offset[i] = plan.offset;
                    continue;
                }
                pb.rc[i] = host_fail(FG_ERR_UNSUPPORTED, "query '%s': a nested boolean query does not fit a flat fg_query_batch (use fgh_plan / fgh_search)", queries[i]);
            }
            const bool bad = pb.rc[i] != FG_OK;
            if (bad) P.errs[i - P.a] = g_herr;
            pb.q[i].k = bad ? 1 : plan.k;
            pb.q[i].clause_begin = (uint32_t)P.c.size();
            pb.q[i].n_clauses = bad ? 0 : plan.n_clauses;
            if (!bad) {
                pb.offset[i] = plan.offset;
                for (uint32_t c = 0; c < plan.n_clauses; c++) {
                    fg_clause cl = plan.clauses[c];
                    cl.leaf_begin += (uint32_t)P.l.size();
                    P.c.push_back(cl);
                }
                P.l.insert(P.l.end(), plan.leaves, plan.leaves + plan.n_leaves);
            }
        }
    };
    static const bool timing = getenv("FG_TIMING") != nullptr;
    timespec ts0, ts1, ts2;
    if (timing) clock_gettime(CLOCK_MONOTONIC, &ts0);
    fg::HostPool::get().run(T, work);
    if (timing) clock_gettime(CLOCK_MONOTONIC, &ts1);
    // concatenate the per-thread parts (clause / leaf indices become global): offsets first, then every thread copies its own part
    std::vector<uint32_t> cb((size_t)T + 1, 0), lb((size_t)T + 1, 0);
    for (int t = 0; t < T; t++) {
        cb[t + 1] = cb[t] + (uint32_t)parts[t].c.size();
        lb[t + 1] = lb[t] + (uint32_t)parts[t].l.size();
    }
    pb.c.resize(cb[T]);
    pb.l.resize(lb[T]);
    std::vector<uint32_t> kmax_t((size_t)T, 1);
    fg::HostPool::get().run(T, [&](int t) {
        Part& P = parts[t];
        fg_clause* dc = pb.c.data() + cb[t];
        for (size_t i = 0; i < P.c.size(); i++) {
            fg_clause cl = P.c[i];
            cl.leaf_begin += lb[t];
            dc[i] = cl;
        }
        if (!P.l.empty()) memcpy(pb.l.data() + lb[t], P.l.data(), P.l.size() * sizeof(fg_leaf));
        uint32_t km = 1;
        for (uint32_t i = P.a; i < P.b; i++) {
            pb.q[i].clause_begin += cb[t];
            km = std::max(km, pb.q[i].k);
        }
        kmax_t[t] = km;
    });
    pb.composites.clear();
    for (int t = 0; t < T; t++) {
        pb.composites.insert(pb.composites.end(), parts[t].comp.begin(), parts[t].comp.end());
        pb.kmax = std::max(pb.kmax, kmax_t[t]);
        if (pb.first_err) continue;
        Part& P = parts[t];
        for (uint32_t i = P.a; i < P.b; i++)
            if (pb.rc[i] != FG_OK && pb.rc[i] != RC_NESTED) { pb.first_err = pb.rc[i]; pb.first_msg = P.errs[i - P.a]; break; }
    }
    if (timing) {
        clock_gettime(CLOCK_MONOTONIC, &ts2);
        fprintf(stderr, "[plan_batch] n=%u threads=%d: parallel %.3f ms, concatenation %.3f ms\n", n, T,
                (ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6, (ts2.tv_sec - ts1.tv_sec) * 1e3 + (ts2.tv_nsec - ts1.tv_nsec) * 1e-6);
    }
}
}  // namespace

extern "C" int32_t fgh_plan_batch(const fgh_dataset* ds, uint32_t n, const char* const* queries,
                                  const char* const* filters, const uint32_t* filter_offsets,
                                  const uint32_t* pages, const uint32_t* per_pages, fg_query* out_queries,
                                  fg_clause* out_clauses, uint32_t cap_clauses, fg_leaf* out_leaves,
                                  uint32_t cap_leaves, uint32_t* n_clauses, uint32_t* n_leaves, int32_t* status) {
    if (!ds || (n && (!queries || !out_queries || !out_clauses || !out_leaves)) || !n_clauses || !n_leaves)
        return host_fail(FG_ERR_INVALID, "fgh_plan_batch: NULL argument");
    PlannedBatch pb;
    plan_batch(ds, n, queries, filters, filter_offsets, pages, per_pages, pb);
    if (pb.c.size() > cap_clauses || pb.l.size() > cap_leaves) return host_fail(FG_ERR_INVALID, "fgh_plan_batch: output capacity too small");
    if (n) memcpy(out_queries, pb.q.data(), n * sizeof(fg_query));
    if (!pb.c.empty()) memcpy(out_clauses, pb.c.data(), pb.c.size() * sizeof(fg_clause));
    if (!pb.l.empty()) memcpy(out_leaves, pb.l.data(), pb.l.size() * sizeof(fg_leaf));
    *n_clauses = (uint32_t)pb.c.size();
    *n_leaves = (uint32_t)pb.l.size();
    if (status) memcpy(status, pb.rc.data(), n * sizeof(int32_t));
    if (pb.first_err && !status) return host_fail(pb.first_err, "%s", pb.first_msg.c_str());
    return FG_OK;
}

namespace {
// Sharded datasets (one process per GPU, fg_comm): every rank is handed the same request. Planning does not depend on
// the shard (term ordinals and statistics are global), so each rank plans 1/n_ranks of the queries and the plans are
// all-gathered (two small collectives: sizes, then the padded arrays); lowering is per shard.
int32_t plan_batch_shared(const fgh_dataset* ds, fg_comm* comm, uint32_t n, const char* const* queries, const char* const* filters,
                          const uint32_t* filter_offsets, const uint32_t* pages, const uint32_t* per_pages, PlannedBatch& pb) {
    int32_t rank = 0, world = 1;
    if (comm) D(dev_api().fg_comm_info(comm, &rank, &world));
    if (world <= 1 || n < (uint32_t)world * 32u) {
        plan_batch(ds, n, queries, filters, filter_offsets, pages, per_pages, pb, false, /*mark_composites=*/true);
        return FG_OK;
    }
    const uint32_t a = (uint32_t)((uint64_t)n * rank / world), b = (uint32_t)((uint64_t)n * (rank + 1) / world);
    PlannedBatch mine;
    plan_batch(ds, b - a, queries + a, filters, filter_offsets ? filter_offsets + a : nullptr, pages ? pages + a : nullptr,
               per_pages ? per_pages + a : nullptr, mine, false, /*mark_composites=*/true);
    // One collective in the common case: every rank sends a fixed-size record sized for plain word queries (<= 4 clauses
    // and 8 leaves per query) with its array lengths in a header; a rank whose plans do not fit says so in the header
    // and all ranks repeat the exchange with records sized for the largest rank.
    struct Hdr { uint32_t nq, nc, nl; int32_t first_err; uint32_t overflow, pad[3]; };
    const uint32_t mq = (n + (uint32_t)world - 1) / (uint32_t)world;
    std::vector<Hdr> hs((size_t)world);
    std::vector<char> send, recv;
    size_t o_off = 0, o_rc = 0, o_c = 0, o_l = 0, rec = 0;
    auto exchange = [&](uint32_t mc, uint32_t ml) -> int32_t {
        // record: [header | queries | offsets | status | clauses | leaves], padded to the capacities
        const size_t o_q = sizeof(Hdr);
        o_off = o_q + (size_t)mq * sizeof(fg_query);
        o_rc = o_off + (size_t)mq * 4;
        o_c = o_rc + (size_t)mq * 4;
        o_l = o_c + (size_t)mc * sizeof(fg_clause);
        rec = ((o_l + (size_t)ml * sizeof(fg_leaf)) + 15) & ~(size_t)15;
        send.assign(rec, 0);
        recv.resize(rec * (size_t)world);
        Hdr h{b - a, (uint32_t)mine.c.size(), (uint32_t)mine.l.size(), mine.first_err, 0u, {0u, 0u, 0u}};
        h.overflow = (h.nc > mc || h.nl > ml) ? 1u : 0u;
        memcpy(send.data(), &h, sizeof(h));
        if (!h.overflow) {
            if (h.nq) {
                memcpy(send.data() + o_q, mine.q.data(), (size_t)h.nq * sizeof(fg_query));
                memcpy(send.data() + o_off, mine.offset.data(), (size_t)h.nq * 4);
                memcpy(send.data() + o_rc, mine.rc.data(), (size_t)h.nq * 4);
            }
            if (h.nc) memcpy(send.data() + o_c, mine.c.data(), (size_t)h.nc * sizeof(fg_clause));
            if (h.nl) memcpy(send.data() + o_l, mine.l.data(), (size_t)h.nl * sizeof(fg_leaf));
        }
        if (int32_t rc = D(dev_api().fg_comm_allgather_bytes(comm, send.data(), rec, recv.data()))) return rc;
        for (int r = 0; r < world; r++) memcpy(&hs[(size_t)r], recv.data() + rec * (size_t)r, sizeof(Hdr));
        return FG_OK;
    };
    static const bool force_small = getenv("FG_PLAN_EXCHANGE_SMALL") != nullptr;  // dev: exercise the repeat path
    if (int32_t rc = exchange(force_small ? 1u : 4u * mq, force_small ? 1u : 8u * mq)) return rc;
    {
        bool any_overflow = false;
        uint32_t mc = 0, ml = 0;
        for (const Hdr& x : hs) { any_overflow = any_overflow || x.overflow; mc = std::max(mc, x.nc); ml = std::max(ml, x.nl); }
        if (any_overflow)  // (every rank sees the same headers: all of them repeat)
            if (int32_t rc = exchange(mc, ml)) return rc;
    }
    const size_t o_q = sizeof(Hdr);
    // unpack: sizes first, then every rank's arrays with one copy each and a fix-up of the indices they carry
    size_t tq = 0, tc = 0, tl = 0;
    for (const Hdr& x : hs) { tq += x.nq; tc += x.nc; tl += x.nl; }
    pb.q.resize(tq); pb.offset.resize(tq); pb.rc.resize(tq); pb.c.resize(tc); pb.l.resize(tl);
    pb.kmax = 1;
    pb.first_err = FG_OK;
    size_t qb_ = 0, cb = 0, lb = 0;
    for (int r = 0; r < world; r++) {
        const char* p = recv.data() + rec * (size_t)r;
        const Hdr& x = hs[(size_t)r];
        if (x.nq) {
            memcpy(pb.q.data() + qb_, p + o_q, (size_t)x.nq * sizeof(fg_query));
            memcpy(pb.offset.data() + qb_, p + o_off, (size_t)x.nq * 4);
            memcpy(pb.rc.data() + qb_, p + o_rc, (size_t)x.nq * 4);
            for (size_t i = qb_; i < qb_ + x.nq; i++) {
                pb.q[i].clause_begin += (uint32_t)cb;
                pb.kmax = std::max(pb.kmax, pb.q[i].k);
            }
        }
        if (x.nc) {
            memcpy(pb.c.data() + cb, p + o_c, (size_t)x.nc * sizeof(fg_clause));
            for (size_t i = cb; i < cb + x.nc; i++) pb.c[i].leaf_begin += (uint32_t)lb;
        }
        if (x.nl) memcpy(pb.l.data() + lb, p + o_l, (size_t)x.nl * sizeof(fg_leaf));
        qb_ += x.nq; cb += x.nc; lb += x.nl;
        if (x.first_err && !pb.first_err) {
            pb.first_err = x.first_err;
            pb.first_msg = r == rank ? mine.first_msg : std::string("a query of the request failed to plan (planned by rank ") + std::to_string(r) + ")";
        }
    }
    return FG_OK;
}

int32_t search_batch_impl(fgh_dataset* ds, fg_comm* comm, uint32_t n, const char* const* queries,
                          const char* const* filters, const uint32_t* filter_offsets,
                          const uint32_t* pages, const uint32_t* per_pages, uint32_t stride,
                          fg_hit* out_hits, uint32_t* out_n, uint32_t* out_match_count,
                          int32_t* status);

// TopDocs order (score descending, doc id ascending inside ties): what the device's 64-bit keys sort by.
inline bool hit_before(const fg_hit& a, const fg_hit& b) { return a.score > b.score || (a.score == b.score && a.doc < b.doc); }

// Merge per-shard result lists (each in TopDocs order, doc ids global and disjoint between shards) and cut the page
// [offset, offset + per_page) out of the first `limit` hits of the merged order.
uint32_t merge_shard_pages(const fg_hit* const* lists, const uint32_t* lens, uint32_t n_lists, uint64_t limit, uint64_t offset,
                           uint32_t per_page, fg_hit* out) {
    std::vector<fg_hit> all;
    size_t total = 0;
    for (uint32_t r = 0; r < n_lists; r++) total += lens[r];
    all.reserve(total);
    for (uint32_t r = 0; r < n_lists; r++) all.insert(all.end(), lists[r], lists[r] + lens[r]);
    const size_t keep = (size_t)std::min<uint64_t>(limit, all.size());
    std::partial_sort(all.begin(), all.begin() + keep, all.end(), hit_before);
    if (offset >= keep) return 0;
    const uint32_t take = (uint32_t)std::min<uint64_t>(keep - offset, per_page);
    if (take) memcpy(out, all.data() + offset, (size_t)take * sizeof(fg_hit));
    return take;
}

// Sharded requests, the queries the fused exchange does not take (deep pages, nested boolean queries): every rank
// evaluates them on its own shard through the one-GPU path (statistics are global, so scores are final; a document
// lives in exactly one shard, so a nested query's sums are complete there), then two collectives for all of them
// together -- (status, length) per query, then the hit lists padded to the longest rank -- and every rank merges.
// A shard cannot contribute more than `limit` hits to the first `limit` of the merged order.
int32_t answer_apart_sharded(fgh_dataset* ds, fg_comm* comm, const std::vector<uint32_t>& apart, const char* const* queries,
                             const char* const* filters, const uint32_t* filter_offsets, const uint32_t* pages,
                             const uint32_t* per_pages, uint32_t stride, fg_hit* out_hits, uint32_t* out_n, int32_t* status) {
    int32_t rank = 0, world = 1;
    if (int32_t r = D(dev_api().fg_comm_info(comm, &rank, &world))) return r;
    const uint32_t m = (uint32_t)apart.size();
    struct Rec { int32_t st; uint32_t n; };
    std::vector<Rec> mine(m), all((size_t)m * world);
    std::vector<fg_hit> loc, buf;
    std::vector<std::string> msgs(m);
    const uint32_t local_docs = std::max(1u, ds->n_docs);
    for (uint32_t j = 0; j < m; j++) {
        const uint32_t i = apart[j];
        const uint32_t page = pages ? pages[i] : 0, pp = per_pages ? per_pages[i] : 20;
        const uint64_t limit = (uint64_t)page * pp + pp;
        const uint32_t want = (uint32_t)std::min<uint64_t>(limit, local_docs), zero = 0;
        const uint32_t f0 = filter_offsets ? filter_offsets[i] : 0, f1 = filter_offsets ? filter_offsets[i + 1] : 0;
        const uint32_t fo[2] = {0, f1 - f0};
        buf.resize(want);
        uint32_t nh = 0;
        int32_t st = FG_OK;
        // (a hard failure is kept as this query's status: the collectives below must still be entered by every rank)
        const int32_t rc = search_batch_impl(ds, nullptr, 1, queries + i, filters ? filters + f0 : nullptr, fo, &zero, &want, want,
                                             buf.data(), &nh, nullptr, &st);
        if (rc || st) { mine[j] = Rec{rc ? rc : st, 0u}; msgs[j] = g_herr; continue; }
        mine[j] = Rec{FG_OK, nh};
        loc.insert(loc.end(), buf.begin(), buf.begin() + nh);
    }
    if (int32_t r = D(dev_api().fg_comm_allgather_bytes(comm, mine.data(), (size_t)m * sizeof(Rec), all.data()))) return r;
    std::vector<size_t> tot((size_t)world, 0);
    size_t longest = 0;
    for (int r = 0; r < world; r++) {
        for (uint32_t j = 0; j < m; j++) tot[(size_t)r] += all[(size_t)r * m + j].n;
        longest = std::max(longest, tot[(size_t)r]);
    }
    std::vector<fg_hit> gathered(longest * (size_t)world);
    if (longest) {
        loc.resize(longest, fg_hit{0.f, 0u});
        if (int32_t r = D(dev_api().fg_comm_allgather_bytes(comm, loc.data(), longest * sizeof(fg_hit), gathered.data()))) return r;
    }
    std::vector<size_t> at((size_t)world, 0);  // start of query j's hits inside every rank's list
    std::vector<const fg_hit*> lists((size_t)world);
    std::vector<uint32_t> lens((size_t)world);
    for (uint32_t j = 0; j < m; j++) {
        const uint32_t i = apart[j];
        int32_t st = FG_OK, bad_rank = -1;
        for (int r = 0; r < world; r++) {
            const Rec& x = all[(size_t)r * m + j];
            if (x.st != FG_OK && st == FG_OK) { st = x.st; bad_rank = r; }
            lists[(size_t)r] = gathered.data() + (size_t)r * longest + at[(size_t)r];
            lens[(size_t)r] = x.n;
            at[(size_t)r] += x.n;
        }
        out_n[i] = 0;
        if (st != FG_OK) {
            if (bad_rank == rank) host_fail(st, "%s", msgs[j].c_str());
            else host_fail(st, "query '%s' failed on the shard of rank %d", queries[i], bad_rank);
            if (!status) return st;
            status[i] = st;
            continue;
        }
        const uint32_t page = pages ? pages[i] : 0, pp = per_pages ? per_pages[i] : 20;
        out_n[i] = merge_shard_pages(lists.data(), lens.data(), (uint32_t)world, (uint64_t)page * pp + pp, (uint64_t)page * pp,
                                     std::min(pp, stride), out_hits + (size_t)i * stride);
    }
    return FG_OK;
}

int32_t search_batch_impl(fgh_dataset* ds, fg_comm* comm, uint32_t n, const char* const* queries,
                          const char* const* filters, const uint32_t* filter_offsets,
                          const uint32_t* pages, const uint32_t* per_pages, uint32_t stride,
                          fg_hit* out_hits, uint32_t* out_n, uint32_t* out_match_count,
                          int32_t* status) {
    if (!ds || (n && (!queries || !out_hits || !out_n))) return host_fail(FG_ERR_INVALID, "fgh_search_batch: NULL argument");
    const std::shared_ptr<fg_index> snap = current_snapshot(ds);  // held until the results are collected
    if (!snap) return host_fail(ds->ctx ? FG_ERR_INVALID : FG_ERR_NO_DEVICE, "dataset has no device snapshot (commit first)");
    if (n == 0) return FG_OK;
    // Deep pages (limit = page*per_page + per_page above 1024; handlers/search.rs:370-374 clamps per_page, not page) are
    // evaluated without per-warp queues: every match above the running threshold is kept and the page is selected
    // afterwards, which costs list space for every query of the batch they are in. They are rare: answer each one in a
    // batch of its own and the rest together.
    if (n > 1 && !comm) {
        std::vector<uint32_t> deep, rest;
        for (uint32_t i = 0; i < n; i++) {
            const uint64_t limit = (uint64_t)(pages ? pages[i] : 0) * (per_pages ? per_pages[i] : 20) + (per_pages ? per_pages[i] : 20);
            (limit > 1024 ? deep : rest).push_back(i);
        }
        if (!deep.empty()) {
            auto sub = [&](const std::vector<uint32_t>& idx) -> int32_t {
                const uint32_t m = (uint32_t)idx.size();
                if (!m) return FG_OK;
                std::vector<const char*> q(m), f;
                std::vector<uint32_t> fo(m + 1, 0), pg(m), pp(m), nh(m), cnt(m);
                std::vector<int32_t> st(m, FG_OK);
                std::vector<fg_hit> hits((size_t)m * stride);
                for (uint32_t j = 0; j < m; j++) {
                    const uint32_t i = idx[j];
                    q[j] = queries[i];
                    pg[j] = pages ? pages[i] : 0;
                    pp[j] = per_pages ? per_pages[i] : 20;
                    if (filter_offsets)
                        for (uint32_t x = filter_offsets[i]; x < filter_offsets[i + 1]; x++) f.push_back(filters[x]);
                    fo[j + 1] = (uint32_t)f.size();
                }
                const int32_t rc = search_batch_impl(ds, nullptr, m, q.data(), f.empty() ? nullptr : f.data(), fo.data(), pg.data(), pp.data(), stride, hits.data(),
                                                     nh.data(), out_match_count ? cnt.data() : nullptr, status ? st.data() : nullptr);
                if (rc) return rc;
                for (uint32_t j = 0; j < m; j++) {
                    const uint32_t i = idx[j];
                    memcpy(out_hits + (size_t)i * stride, hits.data() + (size_t)j * stride, (size_t)nh[j] * sizeof(fg_hit));
                    out_n[i] = nh[j];
                    if (out_match_count) out_match_count[i] = cnt[j];
                    if (status) status[i] = st[j];
                }
                return FG_OK;
            };
            if (int32_t rc = sub(rest)) return rc;
            for (uint32_t i : deep)
                if (int32_t rc = sub(std::vector<uint32_t>{i})) return rc;
            return FG_OK;
        }
    }
    // Large requests are cut into a few chunks and pipelined: while the device evaluates chunk i
    // (fg_batch_submit returns at once) this thread parses, plans and lowers chunk i+1, so the host
    // side of the call hides under the kernels instead of adding to them.
    // (fg_batch_submit alternates between two streams, so the tail of the first chunk's persistent kernel overlaps the start of
    // the second's: 1.71 -> 1.55 ms per request; more chunks cost the host more than the overlap gains,
    // profiles/r02h_e2e_submit_streams.txt.)
    // Chunk count (measured on a B200, 5000-query C2 batch, round 2): planning + lowering of the whole request takes
    // ~0.9 ms on 16 host threads, the kernels ~1.5 ms. One chunk: 2.4 ms; two chunks (20 % / 80 %): 2.3 ms; three:
    // 2.5 ms; four: 2.9 ms -- every extra chunk costs the device more (smaller launches fill it worse) than the
    // overlap gains once the host side is this short.
    uint32_t nch = n >= 3072 ? 2u : 1u;
    if (const char* e = getenv("FG_PIPELINE_CHUNKS")) nch = (uint32_t)std::max(1, atoi(e));
    nch = std::max<uint32_t>(1, std::min(nch, n));
    double first_frac = 1.0 / nch;
    if (nch == 2) first_frac = 0.35;  // (measured, round 2 end: 0.2 -> 1.97 ms, 0.35 -> 1.79 ms per 5000-query C2 request)
    else if (nch > 2) first_frac = 0.4 / nch;
    if (const char* e = getenv("FG_PIPELINE_FIRST")) first_frac = std::min(0.9, std::max(0.05, atof(e)));
    struct Chunk { uint32_t a, b; PlannedBatch pb; fg_batch* batch = nullptr; };
    std::vector<Chunk> ch(nch);
    // Sharded: the whole request is planned first, shared among the ranks (the exchange of the plans is a collective on
    // the context's stream: between chunks it would queue up behind the previous chunk's kernels); the chunks then only
    // lower and submit.
    PlannedBatch whole;
    std::vector<uint32_t> apart;
    const bool planned_whole = comm != nullptr;
    if (planned_whole) {
        if (int32_t prc = plan_batch_shared(ds, comm, n, queries, filters, filter_offsets, pages, per_pages, whole)) return prc;
        // Deep pages (limit above 1024) and nested boolean queries do not go through the fused exchange (its lists are
        // k_stride wide; a nested query is several device queries): an empty query holds their place in the flat batch
        // and answer_apart_sharded evaluates them per shard and merges the pages on the host.
        for (uint32_t i = 0; i < n; i++) {
            if (whole.rc[i] == RC_NESTED) {
                whole.rc[i] = FG_OK;
                apart.push_back(i);
            } else if (whole.rc[i] == FG_OK && whole.q[i].k > 1024) {
                whole.q[i].k = 1;
                whole.q[i].n_clauses = 0;
                apart.push_back(i);
            }
        }
        if (status) memcpy(status, whole.rc.data(), n * sizeof(int32_t));
        if (whole.first_err && !status) return host_fail(whole.first_err, "%s", whole.first_msg.c_str());
    }
    struct Cleanup {
        std::vector<Chunk>& c;
        ~Cleanup() { for (auto& x : c) if (x.batch) dev_api().fg_batch_release(x.batch); }
    } cleanup{ch};
    const bool timing = getenv("FG_TIMING") != nullptr;
    auto now_ms = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_start = now_ms();
    double t_plan = 0, t_prep = 0, t_sub = 0;
    for (uint32_t i = 0; i < nch; i++) {
        Chunk& C = ch[i];
        const double t0 = now_ms();
        // uneven cut: the first chunk is small, so the device starts early, and sized so that it keeps the
        // device busy about as long as the host needs for the rest (host ~ 1/3 of the device time per query)
        auto cut = [&](uint32_t j) -> uint32_t {
            if (j == 0) return 0;
            if (j >= nch) return n;
            const double first = first_frac, rest = (1.0 - first_frac) / (nch - 1);
            return (uint32_t)std::min<double>(n, n * (first + rest * (j - 1)));
        };
        C.a = cut(i);
        C.b = cut(i + 1);
        const uint32_t m = C.b - C.a;
        fg_query_batch qb;
        memset(&qb, 0, sizeof(qb));
        qb.n_queries = m;
        if (planned_whole) {  // a slice of the request's plan (its queries index the request's clause / leaf arrays)
            C.pb.rc.assign(whole.rc.begin() + C.a, whole.rc.begin() + C.b);
            C.pb.offset.assign(whole.offset.begin() + C.a, whole.offset.begin() + C.b);
            C.pb.kmax = 1;
            for (uint32_t j = C.a; j < C.b; j++) C.pb.kmax = std::max(C.pb.kmax, whole.q[j].k);
            qb.n_clauses = (uint32_t)whole.c.size();
            qb.n_leaves = (uint32_t)whole.l.size();
            qb.queries = whole.q.data() + C.a;
            qb.clauses = whole.c.data();
            qb.leaves = whole.l.data();
        } else {
            plan_batch(ds, m, queries + C.a, filters, filter_offsets ? filter_offsets + C.a : nullptr,
                       pages ? pages + C.a : nullptr, per_pages ? per_pages + C.a : nullptr, C.pb, /*keep_composites=*/true);
            if (status) memcpy(status + C.a, C.pb.rc.data(), m * sizeof(int32_t));
            if (C.pb.first_err && !status) return host_fail(C.pb.first_err, "%s", C.pb.first_msg.c_str());  // single-status callers see the first failure
            qb.n_clauses = (uint32_t)C.pb.c.size();
            qb.n_leaves = (uint32_t)C.pb.l.size();
            qb.queries = C.pb.q.data();
            qb.clauses = C.pb.c.data();
            qb.leaves = C.pb.l.data();
        }
        const double t1 = now_ms();
        // The TopDocs form (no match counts: what the reference's collector does) runs on the lead-driven kernels,
        // which prune; match counts need every matching document visited, which is what the windowed accumulator
        // kernels do best (FG_PREP_LEGACY). A query the device path cannot take fails alone, not its siblings.
        int32_t r = FG_ERR_UNSUPPORTED;
        if (out_match_count) r = D(dev_api().fg_batch_prepare_ex(snap.get(), &qb, FG_PREP_LEGACY, &C.batch));
        if (r) r = D(dev_api().fg_batch_prepare_ex(snap.get(), &qb, FG_PREP_PER_QUERY_STATUS, &C.batch));
        if (r) return r;
        {
            std::vector<int32_t> qs(m);
            D(dev_api().fg_batch_query_status(C.batch, qs.data()));
            for (uint32_t j = 0; j < m; j++)
                if (qs[j] != FG_OK && C.pb.rc[j] == FG_OK) {
                    C.pb.rc[j] = qs[j];
                    if (status) status[C.a + j] = qs[j];
                    else return qs[j];  // single-status callers see the first failure (message set by fg_batch_query_status)
                }
        }
        const double t2 = now_ms();
        // match counts are optional: the reference's TopDocs collector does not count (src/db/search.rs:162)
        r = comm ? D(dev_api().fg_batch_submit_sharded(C.batch, comm, 0, C.pb.kmax))
                 : D(dev_api().fg_batch_submit(C.batch, 0, C.pb.kmax, out_match_count ? 1 : 0));
        if (r) return r;
        t_plan += t1 - t0; t_prep += t2 - t1; t_sub += now_ms() - t2;
    }
    const double t_submitted = now_ms();
    std::vector<fg_hit> hits;
    std::vector<uint32_t> nh, cnt;
    for (uint32_t i = 0; i < nch; i++) {
        Chunk& C = ch[i];
        const uint32_t m = C.b - C.a, kmax = C.pb.kmax;
        hits.resize((size_t)m * kmax);
        nh.resize(m);
        cnt.resize(m);
        int32_t r = D(dev_api().fg_batch_collect(C.batch, hits.data(), nh.data(), out_match_count ? cnt.data() : nullptr));
        if (r) return r;
        for (uint32_t j = 0; j < m; j++) {
            const uint32_t qi = C.a + j;
            const bool bad = C.pb.rc[j] != FG_OK;
            const uint32_t off = C.pb.offset[j], pp = per_pages ? per_pages[qi] : 20;
            uint32_t take = (!bad && nh[j] > off) ? std::min(nh[j] - off, std::min(pp, stride)) : 0;  // skip(offset).take(per_page)
            for (uint32_t x = 0; x < take; x++) out_hits[(size_t)qi * stride + x] = hits[(size_t)j * kmax + off + x];
            out_n[qi] = take;
            if (out_match_count) out_match_count[qi] = bad ? 0 : cnt[j];
        }
        // nested queries of the chunk: one fg_search_union_of each (their children are the disjuncts of the plan)
        for (const auto& cp : C.pb.composites) {
            const uint32_t qi = C.a + cp.first;
            const fgh_plan_t& P = cp.second;
            std::vector<fg_query> dq;
            std::vector<fg_clause> dc(P.clauses, P.clauses + P.n_clauses);
            uint32_t n_filter_children = 0;
            for (uint32_t ci = 0; ci < P.n_clauses; ci++) {
                const uint32_t d = dc[ci].occur >> FGH_DISJUNCT_SHIFT;
                if (dc[ci].occur & FGH_FILTER_CHILD) n_filter_children = 1;  // (the last child)
                dc[ci].occur &= FGH_FILTER_CHILD - 1u;
                if (dq.size() < d) dq.push_back(fg_query{1u, ci, 0u});
                dq.back().n_clauses++;
            }
            fg_query_batch qb;
            memset(&qb, 0, sizeof(qb));
            qb.n_queries = (uint32_t)dq.size();
            qb.n_clauses = P.n_clauses;
            qb.n_leaves = P.n_leaves;
            qb.queries = dq.data();
            qb.clauses = dc.data();
            qb.leaves = P.leaves;
            std::vector<fg_hit> page(P.k);
            uint32_t n_page = 0, n_match = 0;
            const int32_t rc = D(dev_api().fg_search_union_of_filtered(snap.get(), &qb, n_filter_children, P.k, page.data(), &n_page, out_match_count ? &n_match : nullptr));
            if (rc) {
                if (!status) return rc;
                status[qi] = rc;
                out_n[qi] = 0;
                continue;
            }
            const uint32_t pp = per_pages ? per_pages[qi] : 20;
            const uint32_t take = n_page > P.offset ? std::min(n_page - P.offset, std::min(pp, stride)) : 0;  // skip(offset).take(per_page)
            for (uint32_t x = 0; x < take; x++) out_hits[(size_t)qi * stride + x] = page[P.offset + x];
            out_n[qi] = take;
            if (out_match_count) out_match_count[qi] = n_match;
        }
    }
    if (!apart.empty())
        if (int32_t r = answer_apart_sharded(ds, comm, apart, queries, filters, filter_offsets, pages, per_pages, stride, out_hits, out_n, status)) return r;
    if (timing)
        fprintf(stderr, "[fgh_search_batch] n=%u chunks=%u: plan %.2f prepare %.2f submit %.2f | all submitted at %.2f, done at %.2f ms\n",
                n, nch, t_plan, t_prep, t_sub, t_submitted - t_start, now_ms() - t_start);
    return FG_OK;
}

}  // namespace

extern "C" int32_t fgh_search_batch(fgh_dataset* ds, uint32_t n, const char* const* queries,
                                    const char* const* filters, const uint32_t* filter_offsets,
                                    const uint32_t* pages, const uint32_t* per_pages, uint32_t stride,
                                    fg_hit* out_hits, uint32_t* out_n, uint32_t* out_match_count,
                                    int32_t* status) {
    return search_batch_impl(ds, nullptr, n, queries, filters, filter_offsets, pages, per_pages, stride, out_hits, out_n, out_match_count, status);
}

extern "C" int32_t fgh_search_batch_sharded(fgh_dataset* ds, fg_comm* comm, uint32_t n, const char* const* queries,
                                            const char* const* filters, const uint32_t* filter_offsets,
                                            const uint32_t* pages, const uint32_t* per_pages, uint32_t stride,
                                            fg_hit* out_hits, uint32_t* out_n, int32_t* status) {
    if (!comm) return host_fail(FG_ERR_INVALID, "fgh_search_batch_sharded: NULL communicator");
    return search_batch_impl(ds, comm, n, queries, filters, filter_offsets, pages, per_pages, stride, out_hits, out_n, nullptr, status);
}

extern "C" uint32_t fgh_merge_shard_pages(const fg_hit* const* lists, const uint32_t* lens, uint32_t n_lists, uint32_t page,
                                          uint32_t per_page, fg_hit* out_hits) {
    if (!n_lists || !lists || !lens || !out_hits || !per_page) return 0;
    return merge_shard_pages(lists, lens, n_lists, (uint64_t)page * per_page + per_page, (uint64_t)page * per_page, per_page, out_hits);
}

extern "C" int32_t fgh_search(fgh_dataset* ds, const char* query, const char* const* filters, uint32_t n_filters,
                              uint32_t page, uint32_t per_page, fg_hit* out_hits, uint32_t* out_n,
                              uint32_t* out_match_count) {
    const uint32_t offs[2] = {0, n_filters};
    return fgh_search_batch(ds, 1, &query, filters, offs, &page, &per_page, per_page, out_hits, out_n,
                            out_match_count, nullptr);
}

// ------------------------------------------------------------------------------------------
// micro-batcher
// ------------------------------------------------------------------------------------------
// The HTTP API answers one query per request (search_endpoint, /root/reference/src/server/handlers/search.rs:152;
// query_json_post :210), each on its own tokio worker thread inside Dataset::search; the device wants thousands
// of queries per launch. fgh_batcher_search is what such a thread calls instead of fgh_search: the request is
// queued, a dispatcher thread gathers what has arrived -- everything that queued up while the previous batch was
// on the device, or, when the device is idle, whatever arrives within `max_wait_us` of the first request, up to
// `max_batch` -- and answers it with ONE fgh_search_batch; every caller wakes up with its own page. A request the
// device path cannot take fails alone with its own status (the caller keeps tantivy for it).
struct fgh_batcher {
    struct Req {
        const char* query;
        const char* const* filters;
        uint32_t n_filters, page, per_page;
        fg_hit* out_hits;
        uint32_t* out_n;
        int32_t rc = FG_OK;
        std::string err;
        bool done = false;
        std::chrono::steady_clock::time_point t_in;
    };
    fgh_dataset* ds = nullptr;
    uint32_t max_batch = 4096, max_wait_us = 200;
    std::mutex mu;
    std::condition_variable cv_in, cv_out;
    std::deque<Req*> queue;
    bool stop = false;
    std::thread worker;
    fgh_batcher_stats st{};

    void run() {
        std::vector<Req*> reqs;
        std::vector<const char*> qs, fl;
        std::vector<uint32_t> foff, pages, pps, nh;
        std::vector<int32_t> status;
        std::vector<fg_hit> hits;
        std::unique_lock<std::mutex> g(mu);
        while (true) {
            cv_in.wait(g, [&] { return stop || !queue.empty(); });
            if (queue.empty()) return;  // stop requested and nothing pending
            // device idle: give concurrent callers a short window to join the first request's batch
            const auto deadline = queue.front()->t_in + std::chrono::microseconds(max_wait_us);
            while (!stop && queue.size() < max_batch && std::chrono::steady_clock::now() < deadline) cv_in.wait_until(g, deadline);
            reqs.clear();
            while (!queue.empty() && reqs.size() < max_batch) { reqs.push_back(queue.front()); queue.pop_front(); }
            g.unlock();
            const uint32_t n = (uint32_t)reqs.size();
            qs.resize(n); pages.resize(n); pps.resize(n); nh.assign(n, 0); status.assign(n, FG_OK);
            fl.clear(); foff.assign(n + 1, 0);
            uint32_t stride = 1;
            for (uint32_t i = 0; i < n; i++) {
                const Req& r = *reqs[i];
                qs[i] = r.query;
                pages[i] = r.page;
                pps[i] = r.per_page;
                stride = std::max(stride, r.per_page);
                for (uint32_t f = 0; f < r.n_filters; f++) fl.push_back(r.filters[f]);
                foff[i + 1] = (uint32_t)fl.size();
            }
            hits.resize((size_t)n * stride);
            const int32_t rc = fgh_search_batch(ds, n, qs.data(), fl.empty() ? nullptr : fl.data(), foff.data(), pages.data(), pps.data(), stride,
                                                hits.data(), nh.data(), nullptr, status.data());
            const std::string batch_err = rc ? g_herr : "";
            const auto t_done = std::chrono::steady_clock::now();
            g.lock();
            st.n_batches++;
            st.n_requests += n;
            st.max_batch_seen = std::max<uint64_t>(st.max_batch_seen, n);
            for (uint32_t i = 0; i < n; i++) {
                Req& r = *reqs[i];
                st.wait_us_total += (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(t_done - r.t_in).count();
                if (rc) { r.rc = rc; r.err = batch_err; }
                else if (status[i]) { r.rc = status[i]; r.err = "query not evaluated on the device path (status " + std::to_string(status[i]) + ")"; }
                else {
                    memcpy(r.out_hits, hits.data() + (size_t)i * stride, (size_t)nh[i] * sizeof(fg_hit));
                    *r.out_n = nh[i];
                }
                r.done = true;
            }
            cv_out.notify_all();
        }
    }
};

extern "C" int32_t fgh_batcher_create(fgh_dataset* ds, uint32_t max_batch, uint32_t max_wait_us, fgh_batcher** out) {
    if (!ds || !out) return host_fail(FG_ERR_INVALID, "fgh_batcher_create: NULL argument");
    fgh_batcher* b = new fgh_batcher();
    b->ds = ds;
    b->max_batch = max_batch ? max_batch : 4096;
    b->max_wait_us = max_wait_us;
    b->worker = std::thread([b] { b->run(); });
    *out = b;
    return FG_OK;
}
extern "C" void fgh_batcher_destroy(fgh_batcher* b) {
    if (!b) return;
    {
        std::lock_guard<std::mutex> g(b->mu);
        b->stop = true;  // requests already queued are still answered
    }
    b->cv_in.notify_all();
    if (b->worker.joinable()) b->worker.join();
    delete b;
}
extern "C" int32_t fgh_batcher_search(fgh_batcher* b, const char* query, const char* const* filters, uint32_t n_filters,
                                      uint32_t page, uint32_t per_page, fg_hit* out_hits, uint32_t* out_n) {
    if (!b || !out_hits || !out_n) return host_fail(FG_ERR_INVALID, "fgh_batcher_search: NULL argument");
    if (per_page == 0) return host_fail(FG_ERR_INVALID, "per_page == 0: TopDocs::with_limit panics on a zero limit");
    fgh_batcher::Req r;
    r.query = query ? query : "";
    r.filters = filters;
    r.n_filters = filters ? n_filters : 0;
    r.page = page;
    r.per_page = per_page;
    r.out_hits = out_hits;
    r.out_n = out_n;
    *out_n = 0;
    r.t_in = std::chrono::steady_clock::now();
    std::unique_lock<std::mutex> g(b->mu);
    if (b->stop) return host_fail(FG_ERR_INVALID, "fgh_batcher_search: the batcher is shutting down");
    b->queue.push_back(&r);
    b->cv_in.notify_one();
    b->cv_out.wait(g, [&] { return r.done; });
    g.unlock();
    if (r.rc) return host_fail(r.rc, "%s", r.err.c_str());
    return FG_OK;
}
extern "C" int32_t fgh_batcher_get_stats(fgh_batcher* b, fgh_batcher_stats* out) {
    if (!b || !out) return host_fail(FG_ERR_INVALID, "fgh_batcher_get_stats: NULL argument");
    std::lock_guard<std::mutex> g(b->mu);
    *out = b->st;
    return FG_OK;
}

// ------------------------------------------------------------------------------------------
// facet counting: FacetCollector::for_field("facet") + add_facet(root)
// over AllQuery, src/db/facet.rs:35-103,206-233
// ------------------------------------------------------------------------------------------
namespace {

struct FacetEnt { const char* p; uint32_t len, ord, depth; };

// tantivy orders facets by their encoded form, in which the segments are joined by '\0'
// (Facet's Ord is the byte order of that string): '/' compares below every other byte
inline bool facet_less(const FacetEnt& a, const FacetEnt& b) {
    const uint32_t n = std::min(a.len, b.len);
    for (uint32_t i = 0; i < n; i++) {
        const unsigned ca = a.p[i] == '/' ? 0u : (unsigned char)a.p[i], cb = b.p[i] == '/' ? 0u : (unsigned char)b.p[i];
        if (ca != cb) return ca < cb;
    }
    return a.len < b.len;
}

// dictionary entries strictly below `root`, at most max_depth segments below it (0 = no limit), in
// facet order, which is also the pre-order of collect_facets_recursive (src/db/facet.rs:206-233).
// Caller holds ds->mu (the pool the entries point into grows with upserts).
int32_t facet_enumerate(const fgh_dataset* ds, const char* root, uint32_t max_depth, std::vector<FacetEnt>& out) {
    std::vector<std::string> segs;
    if (!root || !facet_segments(root, segs))  // Facet::from(path) panics on a path without the leading '/'
        return host_fail(FG_ERR_INVALID, "facet root '%s' must start with '/'", root ? root : "(null)");
    const std::string rk = facet_key(segs, segs.size());  // "" for the root facet "/"
    const TermDict& dict = ds->f[FGH_FIELD_FACET].dict;
    for (const TermDict::Ent& e : dict.ents) {
        const char* s = dict.pool.data() + e.off;
        if (e.len <= rk.size() + 1 || memcmp(s, rk.data(), rk.size()) != 0 || s[rk.size()] != '/') continue;
        uint32_t depth = 0;
        for (uint32_t i = (uint32_t)rk.size(); i < e.len; i++) depth += s[i] == '/';
        if (max_depth && depth > max_depth) continue;
        out.push_back({s, e.len, e.ord, depth});
    }
    std::sort(out.begin(), out.end(), facet_less);
    return FG_OK;
}

int32_t facet_emit(const std::vector<FacetEnt>& ents, const std::vector<uint64_t>* counts, fgh_facet_entry* out,
                   uint32_t cap, char* path_buf, uint32_t path_cap, uint32_t* n_out, uint32_t* path_bytes_out) {
    uint32_t n = 0;
    uint64_t bytes = 0;
    for (size_t i = 0; i < ents.size(); i++) {
        if (counts && (*counts)[i] == 0) continue;  // the collector reports facets of matching docs only
        if (out) {
            if (n >= cap || bytes + ents[i].len + 1 > path_cap)
                return host_fail(FG_ERR_INVALID, "facet output too small (call with out == NULL for the required sizes)");
            memcpy(path_buf + bytes, ents[i].p, ents[i].len);
            path_buf[bytes + ents[i].len] = 0;
            out[n] = {ents[i].ord, ents[i].depth, counts ? (*counts)[i] : 0, (uint32_t)bytes, ents[i].len};
        }
        n++;
        bytes += ents[i].len + 1;
    }
    if (n_out) *n_out = n;
    if (path_bytes_out) *path_bytes_out = (uint32_t)std::min<uint64_t>(bytes, 0xFFFFFFFFull);
    return FG_OK;
}

}  // namespace

extern "C" int32_t fgh_facet_children(const fgh_dataset* ds, const char* root, uint32_t max_depth,
                                      fgh_facet_entry* out, uint32_t cap, char* path_buf, uint32_t path_cap,
                                      uint32_t* n_out, uint32_t* path_bytes_out) {
    if (!ds || (out && !path_buf)) return host_fail(FG_ERR_INVALID, "fgh_facet_children: NULL argument");
    std::shared_lock<std::shared_mutex> g(ds->mu);
    std::vector<FacetEnt> ents;
    if (int32_t rc = facet_enumerate(ds, root, max_depth, ents)) return rc;
    return facet_emit(ents, nullptr, out, cap, path_buf, path_cap, n_out, path_bytes_out);
}

extern "C" int32_t fgh_facet_counts(fgh_dataset* ds, const char* root, uint32_t max_depth,
                                    fgh_facet_entry* out, uint32_t cap, char* path_buf, uint32_t path_cap,
                                    uint32_t* n_out, uint32_t* path_bytes_out) {
    if (!ds || (out && !path_buf)) return host_fail(FG_ERR_INVALID, "fgh_facet_counts: NULL argument");
    std::vector<FacetEnt> ents;
    std::vector<std::string> keep;  // copies: the dictionary pool may grow once the lock is dropped
    std::shared_ptr<fg_index> snap;
    {
        std::shared_lock<std::shared_mutex> g(ds->mu);
        if (int32_t rc = facet_enumerate(ds, root, max_depth, ents)) return rc;
        snap = ds->index;
        keep.reserve(ents.size());
        for (auto& e : ents) { keep.emplace_back(e.p, e.len); }
        for (size_t i = 0; i < ents.size(); i++) ents[i].p = keep[i].data();
    }
    if (!snap) return host_fail(ds->ctx ? FG_ERR_INVALID : FG_ERR_NO_DEVICE, "dataset has no device snapshot (commit first)");
    if (!out) return facet_emit(ents, nullptr, nullptr, 0, nullptr, 0, n_out, path_bytes_out);  // size query: upper bounds
    {
        // facets first seen after the last commit are not in the snapshot's dictionary (ordinals are
        // assigned in insertion order): they count 0 docs, like an uncommitted document in the reference
        std::vector<FacetEnt> kept;
        for (auto& e : ents)
            if (dev_api().fg_index_term_info(snap.get(), FGH_FIELD_FACET, e.ord, nullptr, nullptr, nullptr, nullptr) == FG_OK) kept.push_back(e);
        ents.swap(kept);
    }
    // one single-leaf query per facet: its match count is the number of ALIVE docs that carry the
    // facet or a descendant (every ancestor path of a document's facets is its own term)
    const uint32_t n = (uint32_t)ents.size();
    std::vector<uint64_t> counts(n, 0);
    if (n) {
        std::vector<fg_query> q(n);
        std::vector<fg_clause> c(n);
        std::vector<fg_leaf> l(n);
        for (uint32_t i = 0; i < n; i++) {
            q[i] = {1u, i, 1u};
            c[i] = {FG_OCCUR_SHOULD, i, 1u};
            l[i] = {FGH_FIELD_FACET, ents[i].ord, 1.f};
        }
        fg_query_batch qb;
        memset(&qb, 0, sizeof(qb));
        qb.n_queries = qb.n_clauses = qb.n_leaves = n;
        qb.queries = q.data();
        qb.clauses = c.data();
        qb.leaves = l.data();
        std::vector<fg_hit> hits(n);
        std::vector<uint32_t> nh(n), cnt(n);
        if (int32_t rc = D(dev_api().fg_search_batch(snap.get(), &qb, 1, hits.data(), nh.data(), cnt.data()))) return rc;
        for (uint32_t i = 0; i < n; i++) counts[i] = cnt[i];
    }
    return facet_emit(ents, &counts, out, cap, path_buf, path_cap, n_out, path_bytes_out);
}
