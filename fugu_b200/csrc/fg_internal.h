// Internal structures shared by the host side (layout builder, plan lowering) and the kernels.
// Not part of the ABI (include/fugu_gpu.h is).
#pragma once
#include <stdint.h>

namespace fg {

// ---- HBM posting layout ----------------------------------------------------------------------
// A term's postings are cut into blocks of 128. Block b stores
//   gaps  g_i = doc_i - doc_{i-1} - 1  (doc_{-1} = first_base - 1), bit-packed with `bd` bits each,
//   tf_i - 1                             bit-packed with `bt` bits each,
// as two little-endian horizontal bit streams (value i at bit i*width), doc stream first, each
// padded to 128 values, so a block occupies 16*(bd+bt) bytes and starts 16-byte aligned.
// One 16-byte skip entry per block:
struct SkipEntry {
    uint32_t last_doc;    // doc id of the last posting of the block
    uint32_t first_base;  // previous block's last_doc + 1 (0 for a term's first block)
    uint32_t off16;       // payload offset in 16-byte units
    uint32_t packed;      // bd[0:6) | bt[6:12) | (n-1)[12:19)
};
constexpr int BLOCK = 128;
static inline uint32_t pack_meta(uint32_t bd, uint32_t bt, uint32_t n) {
    return bd | (bt << 6) | ((n - 1) << 12);
}

constexpr int MAX_FIELDS = 8;     // fields per index
constexpr int MAX_LEAVES = 16;    // live leaves per query on the device
constexpr int MAX_MUST = 6;       // Must clauses per query (mask bits 0..5)
constexpr uint32_t BIT_SHOULD = 0x40, BIT_NOT = 0x80;

// leaf roles, in evaluation order inside a query
enum : uint32_t { ROLE_INSERT = 0, ROLE_MUST = 1, ROLE_SHOULD = 2, ROLE_NOT = 3 };

struct DevLeaf {
    uint32_t blk_begin;  // first skip entry of the term
    uint32_t n_blocks;
    float weight;        // boost * idf * (1 + K1), from GLOBAL statistics
    float cnorm;         // constant BM25 norm when the field has no fieldnorms
    int32_t fn_field;    // field slot for fieldnorm ids / cache table, -1 = use cnorm
    uint32_t bit;        // mask bit this leaf's clause sets
    uint32_t role;
    uint32_t req;        // mask bits a slot must already carry (filter roles)
    uint32_t build_cb;   // rebuild the candidate bitmap after this leaf (mask value to test), 0 = no
    uint32_t solo;       // long list in a dense-mode plan: gets a phase of its own (plain adds, no atomics)
    uint32_t lflags;     // LF_*
    uint32_t pad0;
    const uint8_t* col;  // dense tf column of the term (1 byte per doc, 0 = absent) or nullptr: a column
                         // leaf has no block phases, it is applied to the round's slots in the slot scan
    uint32_t pad1[2];
};
static_assert(sizeof(DevLeaf) == 64, "DevLeaf is uploaded as a flat array");
// a filter-role leaf whose precondition bits are not known before the slot scan (an earlier clause
// has a column leaf): applied to every doc of the dense window, no candidate pre-test
constexpr uint32_t LF_NOFILT = 1;
constexpr uint32_t LF_STREAM = 2;  // streamed by one warp (see DevQuery::n_stream)

constexpr uint32_t MODE_DENSE = 0, MODE_HASH = 1;
constexpr uint32_t QF_NO_MUST = 1;
constexpr uint32_t QF_ALL = 4;         // pure AllQuery: first k alive docs, score = const_score
constexpr uint32_t QF_PURE_UNION = 2;  // only Should clauses, all weights > 0: no clause masks needed
constexpr uint32_t QF_COL_INSERT = 8;  // an insert-role leaf is a column leaf: every doc of the range is a candidate

struct DevQuery {
    uint32_t leaf_begin;
    uint32_t n_leaves;
    uint32_t n_insert;    // leading leaves with ROLE_INSERT
    uint32_t k;
    uint32_t all_must;    // mask a matching slot must equal (low 7 bits)
    uint32_t flags;
    float const_score;    // sum of Must AllQuery scores
    uint32_t item_begin;  // work items of this query are contiguous in the partial arrays
    uint32_t n_items;
    uint32_t n_col;       // column leaves, stored after the n_leaves block leaves
    uint32_t col_req;     // mask bits block phases have fully decided: a slot lacking one can never match
    uint32_t n_stream;    // trailing block leaves (of the n_leaves) that are streamed, one warp per leaf
};
static_assert(sizeof(DevQuery) == 48, "DevQuery is uploaded as a flat array");

struct DevItem {
    uint32_t query;
    uint32_t doc_lo, doc_hi;
    uint32_t mode;
    uint32_t slot;  // index into partial arrays (query.item_begin + j)
    uint32_t cls;   // kernel class: 0 dense pure, 1 dense masked, 2 hash pure, 3 hash masked, 4 column scan
    uint32_t pad[2];
};

struct DevIndex {
    const uint4* skip;
    const float* bmax;   // per block: max over its postings of tf / (tf + norm(doc)) (block-max metadata)
    const uint8_t* blk;
    const uint8_t* fnorm[MAX_FIELDS];
    const float* cache;  // [MAX_FIELDS][256]
    const uint32_t* alive;
    uint32_t n_docs;
    uint32_t doc_base;
    uint32_t n_alive;
};

struct SearchParams {
    DevIndex ix;
    const DevQuery* queries;
    const DevLeaf* leaves;
    const DevItem* items;
    uint32_t n_items;
    uint32_t item_begin;         // first item of this launch (items are grouped by kernel class)
    uint32_t kcap;               // entries per partial list
    uint64_t* partial;           // [n_items][kcap] sortable keys, 0 = empty
    uint32_t* partial_count;     // [n_items] matching docs
    unsigned long long* stats;   // [5]: bytes_blocks, bytes_redecode, scored, column-scan chunks seen, ... skipped
    uint32_t* match_bitmap;      // optional
    uint32_t bitmap_words;
    uint32_t exact_filter;
    uint32_t deterministic;     // one phase per leaf: bit-reproducible sums in leaf order
    uint32_t want_counts;       // the caller asked for match counts (the reference's TopDocs does not)
    uint32_t acct;              // maintain the byte / scored-posting counters
    unsigned long long* prof;   // optional [8] cycle counters (dev tool)
    uint32_t* qtheta;           // [n_queries] sortable f32: score every work item may prune below (0 = none)
    uint32_t no_prune;          // column scan: visit every chunk even when no doc of it can reach the top-k (A/B switch)
};

struct MergeParams {
    const DevQuery* queries;
    uint32_t n_queries;
    uint32_t kcap;
    const uint64_t* partial;
    const uint32_t* partial_count;
    uint32_t k_stride;
    uint32_t doc_base;
    const uint32_t* alive;
    uint32_t n_docs, n_alive;
    void* out_hits;        // fg_hit [n_queries][k_stride]
    uint32_t* out_n;
    uint32_t* out_count;   // may be null
};

// kernel geometry
constexpr int NT = 256;          // threads per CTA
constexpr int NW = NT / 32;      // warps per CTA
constexpr int DW = 8192;         // dense window: docs per round
constexpr int HS_LOG2 = 12;
constexpr int HS = 1 << HS_LOG2; // hash slots per round
constexpr int HBLK = 16;         // insert-leaf blocks per hash round (<= HS/2/128)
constexpr int SLOTS = DW > HS ? DW : HS;
constexpr int CBW = 1024;        // candidate bitmap words (32768 bits; dense mode uses the first DW bits)
constexpr int CB_LOG2 = 15;

constexpr int NCLS = 5;          // kernel classes (DevItem::cls)
#ifndef FG_CW
#define FG_CW 6144
#endif
constexpr int CW = FG_CW;       // column-scan window (two accumulator buffers of CW floats)
void launch_search(const SearchParams& p, int ks, const uint32_t class_count[NCLS], void* const streams[NCLS]);
void launch_merge(const MergeParams& p, int ks, void* stream);
void launch_merge_gathered(const void* hits, const uint32_t* n, uint32_t n_ranks, uint32_t n_queries,
                           uint32_t k, uint32_t k_stride, void* out_hits, uint32_t* out_n, int ks,
                           void* stream);
int search_smem_bytes(int ks, bool pure);

// ---- lead-driven evaluation (fg_lead.cu) ---------------------------------------------------------
// A query is answered by walking some of its leaves as LEADS, block after block, and looking every
// candidate doc of a lead block up in the query's other leaves (dense tf column byte, or skip-table
// gallop + block decode). Leaves of a query are stored [leads | required | optional | excluded]:
//   pure union      : every Should leaf is a lead, shortest list first (ub = weight * max tf factor bounds a
//                     leaf's contribution); a candidate of lead i that also occurs in a lead j < i is dropped
//                     (lead j scores it)
//   plan with Must  : the leaves of the Must clause with the smallest document frequency are the leads,
//                     the other Must clauses are required lookups (ascending frequency), Should leaves
//                     optional lookups, MustNot leaves excluding lookups.
// MaxScore / block-max pruning (TopDocs form only; results are identical to exhaustive evaluation): a
// lead is not walked at all once ub + rest < theta (theta = k-th best score known for the query), a
// block is decoded only if weight * bmax[block] + rest >= theta, a candidate is looked up only while
// its partial score + the upper bounds of the leaves not applied yet >= theta.
constexpr int LMAX_LEAVES = 32;   // live leaves per query (16 words x [text, name])
constexpr int LHIST_B = 128;
constexpr int LHIST_SHIFT = 17;
enum : uint32_t { LR_LEAD = 0, LR_REQ = 1, LR_OPT = 2, LR_NOT = 3 };
struct LLeaf {
    uint32_t blk_begin, n_blocks;
    float weight;        // boost * idf * (1 + K1), GLOBAL statistics
    float cnorm;         // constant BM25 norm of a field without fieldnorms
    int32_t fn_field;    // fieldnorm field or -1
    float ub;            // upper bound of this leaf's score contribution (0 for LR_NOT)
    float rest;          // LR_LEAD: sum of ub over the later leads, required and optional leaves
    uint32_t role;       // LR_* | clause index << 8 (required leaves of one Must clause share the index)
    const uint8_t* col;  // dense tf column or nullptr
    const uint32_t* bits; // membership bitmap of the term (1 bit per doc) or nullptr; 8 words per 256-doc chunk
    const uint32_t* rank; // with bits: number of postings before each 256-doc chunk
    uint32_t df;         // local document frequency
    uint32_t slot;       // lookup-cache slot of a leaf that is looked up by gallop + decode (running index over the query's such leaves)
};
static_assert(sizeof(LLeaf) == 64, "LLeaf is uploaded as a flat array");
constexpr uint32_t LQ_ALL = 1;    // pure AllQuery: first k alive docs, score = const_score
constexpr uint32_t LQ_PRUNE = 2;  // every scoring weight is > 0: the upper bounds hold
struct LQuery {
    uint32_t leaf_begin, n_leaves, n_lead, n_req, n_opt;
    uint32_t k;
    uint32_t flags;
    float const_score;   // Must AllQuery scores
    float slack;         // added to an upper bound before it is compared with theta (f32 summation order)
    uint32_t part_begin; // this query's region of the partial array
    uint32_t part_cap;
    uint32_t theta0;     // sortable f32 lower bound of the k-th best score known at lowering time, 0 = none
    uint32_t hist_base;  // (f32 bits of the query's total upper bound) >> LHIST_SHIFT: top bucket of its score histogram
    uint32_t sel_begin;  // deep-page batches (k > 1024): this query's region of the selection scratch (next power of two >= min(k, part_cap) keys)
    uint32_t lead_docs;  // min(sum of the leads' document frequencies, n_docs): no more documents are ever offered
    uint32_t pad[1];
};
static_assert(sizeof(LQuery) == 64, "LQuery is uploaded as a flat array");
// Shared threshold of a query: every warp adds the score of each hit it accepts to a per-query histogram
// (global atomics; a document is offered exactly once across all warps). Bucket = the score's f32 bits
// >> LHIST_SHIFT (sign, exponent, 6 mantissa bits: 64 buckets per octave, exact and monotone), relative to
// the bucket of the query's total upper bound; LHIST_B buckets cover the two octaves below it. The lower edge
// of the highest bucket at which the counts from the top reach k is a lower bound of the k-th best score.
struct LItem {
    uint32_t query;
    uint32_t lead;    // lead leaves this item walks, one after the other: [lead & 0xFFFF, lead >> 16) (indices inside the query)
    uint32_t cursor;  // block cursor of the first of them (the following leads use the following cursors); shared by
                      // the copies of the item (each copy is one warp)
    float bound;      // no document this item can offer scores above it (max over its leads of ub + rest + slack): an
                      // item whose bound is below the query's threshold is skipped before its plan is even loaded
};
// Work items as the host writes them: one record per (query, lead group) with the number of copies the queue holds of it
// (a long lead is walked by several warps sharing its block cursor) and the position of the first copy in the device's
// item array; expand_items_kernel writes the copies. (A 5000-query C2 batch has 183 k items but 28 k records: the host
// neither writes nor uploads the copies.)
struct LItemRec {
    LItem item;
    uint32_t dst;     // index of the first copy in the item array
    uint32_t copies;
};
static_assert(sizeof(LItemRec) == 24, "LItemRec is uploaded as a flat array");
void launch_expand_items(const LItemRec* recs, uint32_t n_recs, LItem* items, void* stream);

struct LeadParams {
    DevIndex ix;
    const LQuery* queries;
    const LLeaf* leaves;
    const LItem* items;
    uint32_t n_items;
    uint32_t n_queries;
    uint32_t* work;            // [1] next item (dynamic queue: items start in array order)
    uint32_t* cursors;         // [n_cursors] next block of a (query, lead) pair
    uint32_t n_cursors;
    uint32_t* qtheta;          // [n_queries] sortable f32 threshold shared by all warps of a query
    uint32_t* qcount;          // [n_queries] entries appended to the query's partial region
    uint32_t* qmatch;          // [n_queries] matching docs (exhaustive form)
    uint32_t* qhist;           // [n_queries][LHIST_B] score histogram of the accepted hits (pruned form)
    uint64_t* partial;
    unsigned long long* stats; // [8]
    uint32_t* match_bitmap;
    uint32_t bitmap_words;
    uint32_t exhaustive;       // visit every posting (match counts / bitmaps wanted, or pruning switched off)
    uint32_t want_counts;
    uint32_t acct;
    uint32_t chunk;            // blocks of a lead a warp claims per step
    uint32_t chunk_req;        // plans with required clauses: chunk = min(chunk, max(2, chunk_req / required leaves))
    uint32_t union_work;       // unions: chunk = min(chunk, max(2, union_work / (n_leaves - 1)))
    uint32_t tma;              // stage lead-block payloads in shared memory with 1-D bulk copies (cp.async.bulk)
    uint32_t prof;             // dev tool: per-warp busy time / longest item into stats[8..13]
};
struct LeadMergeParams {
    const LQuery* queries;
    uint32_t n_queries;
    const uint64_t* partial;
    const uint32_t* qcount;
    const uint32_t* qmatch;
    uint32_t k_stride;
    uint32_t doc_base;
    const uint32_t* alive;
    uint32_t n_docs, n_alive;
    void* out_hits;
    uint32_t* out_n;
    uint32_t* out_count;
    uint64_t* sel;  // deep-page batches: selection scratch (LQuery::sel_begin)
};
void launch_lead(const LeadParams& p, int ks, int n_sms, void* stream);
void launch_lead_merge(const LeadMergeParams& p, int ks, void* stream);
// union of the batch's queries as disjuncts of ONE query (fg_search_union_of): one output row; a, b = scratch of cap2 keys
void launch_lead_combine(const LeadMergeParams& p, uint32_t k, uint64_t* a, uint64_t* b, uint32_t cap2, uint32_t n_filter, void* stream);
// merge of all-gathered per-rank lists with per-query k (word k_word of the q_words-word query records at qrec)
void launch_merge_ranks(const void* hits, const uint32_t* n, uint32_t n_ranks, uint32_t n_queries, const void* qrec, uint32_t q_words,
                        uint32_t k_word, uint32_t k_stride, void* out_hits, uint32_t* out_n, int ks, void* stream);
// block-max metadata of the blocks [b0, b1) of one field (upload time)
void launch_blockmax(const DevIndex& ix, uint32_t b0, uint32_t b1, int fn_field, float cnorm, float* bmax, void* stream);
// membership bitmaps + rank directories of selected terms (upload time): sel[i] = {global block, bitmap slot};
// bits is zeroed by the caller, stride_words is a multiple of 8
void launch_bitmap_build(const DevIndex& ix, const uint2* sel, uint32_t n_sel, uint32_t* bits, uint64_t stride_words, void* stream);
void launch_bitmap_rank(const uint32_t* bits, uint32_t* rank, uint32_t n_slots, uint64_t stride_words, void* stream);

// snapshot append (fg_index_append): decode the listed blocks for the host (out_meta[2i] = last_doc, [2i+1] = first_base;
// out_docs / out_tfs 128 slots per block), and copy ranges {src_begin, dst_begin, count} of 16-byte skip entries
void launch_tail_decode(const DevIndex& ix, const uint32_t* blocks, uint32_t n_list, uint32_t* out_meta, uint32_t* out_docs, uint32_t* out_tfs, void* stream);
void launch_copy_ranges(const void* src, void* dst, const void* ranges, uint32_t n_ranges, void* stream);

}  // namespace fg
