/* fugu_host.h — host side of the query path, above the device ABI (include/fugu_gpu.h).
 *
 * The reference's host code is Rust; no Rust toolchain exists in the build image, so the host side
 * is C++ behind this C ABI, mirroring the reference interface for the path name by name:
 *
 *   fgh_dataset_*        <->  Dataset / NamedIndex "docs" index   (src/db/core.rs:39-188,205-497)
 *   fgh_dataset_upsert   <->  DocumentOperations::upsert          (src/db/document.rs:23-67:
 *                              delete_term(id) + add_document; the docs schema src/db/schemas.rs:7-31
 *                              decides what is tokenised: `text`, `name` TEXT, `facet` Facet)
 *   fgh_dataset_commit   <->  writer.commit() + reader reload     (src/db/document.rs:65; core.rs:86)
 *   fgh_search           <->  Dataset::search(query, filters, page, per_page)   (src/db/search.rs:74-218)
 *   fgh_plan             <->  the planning half of it: QueryParser::for_index(.., [text, name]) +
 *                              parse_query + escape fallback (:108-127,603-610), parse_filters (:292-324),
 *                              build_facet_query (:221-289), Must-join (:132-151), limit (:154-160)
 *   fgh_tokenize         <->  tantivy's "default" analyzer selected by TEXT
 *                              (SimpleTokenizer -> RemoveLongFilter(40) -> LowerCaser)
 *
 * Hit hydration (searcher.doc + convert_doc_to_search_result, src/db/search.rs:172-207,534-590) is
 * reduced to the doc -> id side table; stored fields stay with the caller.
 *
 * Same error convention as fugu_gpu.h: int32 status, and fgh_last_error() returns the thread-local message of
 * the last failing fgh_* call on this thread (a failure inside the device library is passed through).
 *
 * Two libraries: libfugu_host.so (this header; plain C++, no CUDA code, no link-time dependency on the device
 * library) and libfugu_gpu.so (fugu_gpu.h). The host library binds the device entry points on first use, from
 * libfugu_gpu.so in its own directory; a process that only plans, tokenises or builds documents (the reference
 * arm of the bench, a GPU-less build of the Rust host) never maps any CUDA code. Calls that need the device
 * fail with FG_ERR_NO_DEVICE when the device library or a CUDA device is missing: there is no CPU fallback.
 */
#ifndef FUGU_HOST_H
#define FUGU_HOST_H

#include "fugu_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* field ids of the docs index as this library lays them out */
#define FGH_FIELD_TEXT 0u
#define FGH_FIELD_NAME 1u
#define FGH_FIELD_FACET 2u

typedef struct fgh_dataset fgh_dataset;

const char* fgh_last_error(void);

/* ctx may be NULL: the dataset then only plans (fgh_plan / fgh_tokenize work, searches fail). */
int32_t fgh_dataset_create(fg_ctx* ctx, fgh_dataset** out);
void fgh_dataset_destroy(fgh_dataset* ds);

/* Upsert one record. `name` may be NULL (no metadata.name). `facets` are facet path strings
 * ("/namespace/ns01/organization/org3"); every ancestor path becomes its own term, as tantivy
 * indexes a Facet. If `id` already exists the old document is deleted (alive bit cleared; statistics
 * keep counting it until a merge, as tantivy's do) and a new document is appended.
 * Validation limits of ObjectRecord::validate (src/object.rs:31-78) are enforced. */
int32_t fgh_dataset_upsert(fgh_dataset* ds, const char* id, const char* text, const char* name,
                           const char* const* facets, uint32_t n_facets);
int32_t fgh_dataset_delete(fgh_dataset* ds, const char* id);
/* Visibility follows the reference's reader: searches, plans and facet counts see the state of the last
 * commit. A document upserted since is invisible, and a term or facet that only such documents contain
 * plans as FG_TERM_MISSING (an empty scorer), never as an error. Planning, searching and counting may run
 * concurrently with upserts and commits (readers share the dataset's lock, writers take it exclusively).
 *
 * Publish the pending state as a new device snapshot (swaps the fg_index; searches already running
 * keep the snapshot they started on). Only deletes since the last commit: the new snapshot shares the
 * posting arrays of the old one and uploads just the alive bitset (fg_index_with_alive, SURVEY.md 8(f)
 * row f3). New documents: only they are handed over, as one new segment (fg_index_append: the postings already
 * in HBM stay there, the derived structures are rebuilt on the device); once the appended part has outgrown the
 * part that was uploaded whole, the next commit rebuilds the snapshot from scratch (tantivy merges segments on a
 * similar schedule). Nothing pending: no-op. Searches, plans and upserts are not held up while the snapshot is
 * built on the device (the dataset's lock is taken only to capture the pending state and to publish the result);
 * commits serialize among themselves; what is upserted while a commit runs becomes visible with the next one. */
int32_t fgh_dataset_commit(fgh_dataset* ds);
/* how the snapshots of this dataset were built so far: full uploads / appended segments */
int32_t fgh_dataset_commit_counts(const fgh_dataset* ds, uint64_t* n_full_uploads, uint64_t* n_appends);

/* Alternative to upsert+commit for large pre-built corpora: adopt a flat CSR (uploaded as is) plus
 * the term dictionaries needed for planning. `terms[f]` = n_terms NUL-terminated strings packed
 * back to back, in term-ordinal order (NULL for a field nobody queries by string). */
int32_t fgh_dataset_adopt(fgh_dataset* ds, const fg_index_desc* desc, const char* const* terms,
                          const uint64_t* terms_bytes);

uint32_t fgh_dataset_num_docs(const fgh_dataset* ds);
fg_index* fgh_dataset_index(fgh_dataset* ds); /* current snapshot (NULL before the first commit); owned by
                                                 the dataset, valid until the next commit */
/* external id of a doc (global doc id as reported in fg_hit.doc); returns length or -1 */
int32_t fgh_dataset_doc_id(const fgh_dataset* ds, uint32_t doc, char* buf, uint32_t cap);
/* term ordinal of a token in a field, FG_TERM_MISSING when absent */
uint32_t fgh_dataset_term_ord(const fgh_dataset* ds, uint32_t field, const char* token);

/* tokens of `text` under the default analyzer, written NUL-separated into buf; returns the token
 * count, or -1 when buf is too small */
int32_t fgh_tokenize(const char* text, char* buf, uint32_t cap);

/* ---- planning ---- */
#define FGH_MAX_PLAN_CLAUSES 16
#define FGH_MAX_PLAN_LEAVES 64
typedef struct {
    uint32_t k;          /* page*per_page + per_page */
    uint32_t offset;     /* page*per_page */
    uint32_t n_clauses;
    uint32_t n_leaves;
    uint32_t is_all;     /* the whole query is AllQuery (empty query, no usable filter) */
    uint32_t used_fallback; /* parse_query failed and the escaped retry was used (src/db/search.rs:120-125) */
    fg_clause clauses[FGH_MAX_PLAN_CLAUSES];
    fg_leaf leaves[FGH_MAX_PLAN_LEAVES];
    /* Nested query (a union whose children are boolean queries themselves, e.g. `(a AND b) OR (c AND d)`, `a OR (b AND c)`):
     * the number of children; clause i then belongs to child `clauses[i].occur >> FGH_DISJUNCT_SHIFT` (1-based; the
     * clauses of a child are contiguous, the plain words of the top level form one child) and the low bits hold its
     * occur. Such a plan is answered with fg_search_union_of (fgh_search / fgh_search_batch do that by themselves);
     * fg_batch_prepare rejects its clauses. 0 = an ordinary one-level plan. */
    uint32_t n_disjuncts;
} fgh_plan_t;
#define FGH_DISJUNCT_SHIFT 8
/* A nested query combined with facet filters (Bool[Must(text_query), Must(facet_query)], src/db/search.rs:140-144): the
 * plan's LAST child is a filter child, its clauses carry this bit in `occur` -- a document matches when an ordinary child
 * and the filter child hold it, and scores the children's sum plus the filter's score (fg_search_union_of_filtered). */
#define FGH_FILTER_CHILD 0x80u
/* FG_ERR_INVALID = parse error even after the fallback (the reference returns Err -> HTTP 500);
 * FG_ERR_UNSUPPORTED = valid tantivy query the device path does not evaluate (phrase, range, fuzzy, boolean
 * trees nested deeper than a union of one-level boolean queries). */
int32_t fgh_plan(const fgh_dataset* ds, const char* query, const char* const* filters,
                 uint32_t n_filters, uint32_t page, uint32_t per_page, fgh_plan_t* out);

/* Plans n requests (multi-threaded) into one flat fg_query_batch in caller-allocated arrays
 * (worst case FGH_MAX_PLAN_CLAUSES / FGH_MAX_PLAN_LEAVES per query). Requests that fail to plan
 * become empty queries and report their code in status[q] (when status is NULL the first failure is
 * returned instead). */
int32_t fgh_plan_batch(const fgh_dataset* ds, uint32_t n, const char* const* queries,
                       const char* const* filters, const uint32_t* filter_offsets,
                       const uint32_t* pages, const uint32_t* per_pages, fg_query* out_queries,
                       fg_clause* out_clauses, uint32_t cap_clauses, fg_leaf* out_leaves,
                       uint32_t cap_leaves, uint32_t* n_clauses, uint32_t* n_leaves, int32_t* status);

/* ---- search: Dataset::search for one request / a batch of requests ----
 * Writes the requested page (after skip(offset).take(per_page), src/db/search.rs:210-211) to
 * out_hits[q*per_page_stride ..], the number of hits in the page to out_n[q] and, when not NULL,
 * the number of matching docs to out_match_count[q]. status[q] (may be NULL) receives a per-query
 * FG_* code so that one bad query does not fail a batch. */
int32_t fgh_search(fgh_dataset* ds, const char* query, const char* const* filters, uint32_t n_filters,
                   uint32_t page, uint32_t per_page, fg_hit* out_hits, uint32_t* out_n,
                   uint32_t* out_match_count);
int32_t fgh_search_batch(fgh_dataset* ds, uint32_t n, const char* const* queries,
                         const char* const* filters, const uint32_t* filter_offsets /* [n+1] or NULL */,
                         const uint32_t* pages /* or NULL = 0 */, const uint32_t* per_pages /* or NULL = 20 */,
                         uint32_t per_page_stride, fg_hit* out_hits, uint32_t* out_n,
                         uint32_t* out_match_count, int32_t* status);

/* The same request against a dataset whose documents are sharded over several GPUs, one process per GPU
 * (fg_comm in fugu_gpu.h): a COLLECTIVE call -- every rank passes the same queries and gets the
 * same, global, result. The ranks share the planning (each plans 1/n_ranks of the request, the plans are
 * all-gathered), every rank lowers the plan for its own shard, and each pipeline chunk ends in the library's fused
 * exchange (local top-k -> ncclAllGather -> merge). Match counts are not available. Requests the fused exchange does
 * not take -- page limits above 1024 (the reference bounds per_page, not page: src/server/handlers/search.rs:370-374)
 * and nested boolean queries -- are evaluated by every rank on its own shard through the one-GPU path and merged on
 * the host (fgh_merge_shard_pages) after two collectives for all of them together. The dataset must have been
 * adopted from this rank's shard with the GLOBAL statistics (fgh_dataset_adopt with fg_index_desc.global_*). */
int32_t fgh_search_batch_sharded(fgh_dataset* ds, fg_comm* comm, uint32_t n, const char* const* queries,
                                 const char* const* filters, const uint32_t* filter_offsets, const uint32_t* pages,
                                 const uint32_t* per_pages, uint32_t per_page_stride, fg_hit* out_hits,
                                 uint32_t* out_n, int32_t* status);

/* Merge of per-shard result lists on the host (a host that drives several shards from one process, or gathers the
 * shards' lists itself, uses the same merge fgh_search_batch_sharded does): lists[r][0..lens[r]) is shard r's result
 * in TopDocs order (score descending, doc id ascending inside ties; global doc ids, a document lives in one shard),
 * each at least min(page*per_page + per_page, matches of the shard) long. Writes the page
 * skip(page*per_page).take(per_page) of the merged order (src/db/search.rs:154-160, 210-211) and returns its length. */
uint32_t fgh_merge_shard_pages(const fg_hit* const* lists, const uint32_t* lens, uint32_t n_lists, uint32_t page,
                               uint32_t per_page, fg_hit* out_hits);

/* ---- micro-batcher ----------------------------------------------------
 * The reference's HTTP API is one query per request (search_endpoint, src/server/handlers/search.rs:152;
 * query_json_post :210), each request blocking its own worker thread in Dataset::search (src/db/search.rs:74).
 * fgh_batcher_search has fgh_search's contract (same page semantics, same errors, blocking, callable from any
 * number of threads) but answers concurrent callers together: a dispatcher thread takes everything that queued
 * up while the previous batch was on the device -- or, when idle, what arrives within max_wait_us of the first
 * request, at most max_batch (0 = 4096) -- through ONE fgh_search_batch. A request the device path does not
 * evaluate fails alone (its own FG_ERR_*), its siblings are answered. out_hits must hold per_page entries.
 * fgh_batcher_destroy answers what is queued, then stops the dispatcher; no call may be made after it. */
typedef struct fgh_batcher fgh_batcher;
typedef struct {
    uint64_t n_requests;     /* requests answered */
    uint64_t n_batches;      /* fgh_search_batch calls they were answered with */
    uint64_t max_batch_seen;
    uint64_t wait_us_total;  /* sum over requests of (answered - queued) */
} fgh_batcher_stats;
int32_t fgh_batcher_create(fgh_dataset* ds, uint32_t max_batch, uint32_t max_wait_us, fgh_batcher** out);
void fgh_batcher_destroy(fgh_batcher* b);
int32_t fgh_batcher_search(fgh_batcher* b, const char* query, const char* const* filters, uint32_t n_filters,
                           uint32_t page, uint32_t per_page, fg_hit* out_hits, uint32_t* out_n);
int32_t fgh_batcher_get_stats(fgh_batcher* b, fgh_batcher_stats* out);

/* ---- facet counting -------------------------------------------------
 * Mirrors `FacetCollector::for_field("facet")` + `add_facet(root)` run over `AllQuery`
 * (src/db/facet.rs:35-103: get_namespace_facets / get_available_namespaces / list_facet) and the
 * recursive walk of get_facet_tree (collect_facets_recursive, src/db/facet.rs:206-233).
 * fgh_facet_counts reports every facet strictly below `root` down to `max_depth` segments (1 = the
 * direct children, what one FacetCollector call returns; 0 = no limit) that at least one ALIVE document
 * carries, with the number of such documents, in facet order (segments compared bytewise, a parent
 * before its children = the pre-order of the reference's recursion). The counts are the match counts of
 * one single-term query per facet, evaluated on the device by the search kernels (every ancestor path
 * of a document's facets is its own term). fgh_facet_children lists the dictionary entries only
 * (count = 0, no device needed).
 * Paths are written NUL-terminated into path_buf; entry i's path starts at path_off. With out == NULL
 * the calls only report upper bounds of the entries / path bytes needed in n_out / path_bytes_out.
 * `root` must start with '/' (Facet::from panics otherwise): FG_ERR_INVALID. */
typedef struct {
    uint32_t term_ord; /* ordinal in the facet field's dictionary */
    uint32_t depth;    /* segments below root (1 = direct child) */
    uint64_t count;    /* alive documents under this facet */
    uint32_t path_off;
    uint32_t path_len;
} fgh_facet_entry;
int32_t fgh_facet_children(const fgh_dataset* ds, const char* root, uint32_t max_depth,
                           fgh_facet_entry* out, uint32_t cap, char* path_buf, uint32_t path_cap,
                           uint32_t* n_out, uint32_t* path_bytes_out);
int32_t fgh_facet_counts(fgh_dataset* ds, const char* root, uint32_t max_depth,
                         fgh_facet_entry* out, uint32_t cap, char* path_buf, uint32_t path_cap,
                         uint32_t* n_out, uint32_t* path_bytes_out);

#ifdef __cplusplus
}
#endif
#endif /* FUGU_HOST_H */
