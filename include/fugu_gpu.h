/* fugu_gpu.h — C ABI of the B200-native query hot path for fugu.
 *
 * What this replaces in the reference (paths relative to /root/reference):
 *   the single call  `searcher.search(&base_query, &TopDocs::with_limit(search_limit))`
 *   at src/db/search.rs:162, i.e. posting decode -> boolean AND/OR -> BM25 -> top-k, which today
 *   runs inside the un-vendored crate tantivy 0.24.1 (Cargo.toml:48, Cargo.lock:4609-4612).
 *   Everything before that line (query parsing / planning, src/db/search.rs:86-160) and after it
 *   (hit hydration / pagination, src/db/search.rs:169-217) stays on the host.
 *
 * Contract:
 *   - plain C, plain pointers and sizes; no torch / C++ types cross the boundary;
 *   - every entry point returns an int32 status (FG_OK = 0, negative = error) and never unwinds;
 *     fg_last_error() returns a thread-local message for the last failing call on this thread;
 *   - thread-safe and re-entrant: axum handlers call Dataset::search concurrently
 *     (src/server/server_main.rs:50, src/db/config.rs:188-190); calls on one fg_ctx serialise on its
 *     stream; an fg_index is an immutable snapshot (mirrors a tantivy `Searcher`);
 *   - inputs are borrowed for the duration of the call, fg_index_upload copies, results are written
 *     into caller-allocated buffers, handles are released by explicit *_destroy / *_release;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     FG_ERR_NO_DEVICE.
 *
 * Semantics implemented on the device are tantivy 0.24.1's as fugu configures them
 * (SURVEY.md Appendix A): BM25 with K1 = 1.2, B = 0.75 in f32, 256-entry fieldnorm table,
 * boolean Must / Should / MustNot with one level of grouping, TopDocs order
 * (score descending, doc id ascending).
 */
#ifndef FUGU_GPU_H
#define FUGU_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (src/db/search.rs:79 returns Result<_, Box<dyn Error>>; HTTP 500 on Err,
 *      src/server/handlers/search.rs:196-205) ---- */
#define FG_OK 0
#define FG_ERR_INVALID (-1)     /* bad argument (NULL, k == 0: TopDocs::with_limit asserts limit >= 1) */
#define FG_ERR_UNSUPPORTED (-2) /* plan node the device path does not evaluate (phrase, range, deep trees) */
#define FG_ERR_CUDA (-3)        /* CUDA runtime error; message in fg_last_error() */
#define FG_ERR_OOM (-4)
#define FG_ERR_NO_DEVICE (-5)   /* no CUDA device: the product never falls back to a CPU path */

const char* fg_last_error(void);
const char* fg_version(void);

/* ---- context: one per (process, device) ---- */
typedef struct fg_ctx fg_ctx;
int32_t fg_ctx_create(int32_t device, fg_ctx** out);
void fg_ctx_destroy(fg_ctx* ctx);
/* Run all work of this context on `cuda_stream` (a cudaStream_t / CUstream, e.g. torch's current
 * stream) instead of the context's own stream. NULL restores the own stream. */
int32_t fg_ctx_set_stream(fg_ctx* ctx, void* cuda_stream);
int32_t fg_ctx_synchronize(fg_ctx* ctx);

/* ---- index snapshot ----------------------------------------------------------------------
 * The reference-side loader walks tantivy's PUBLIC per-segment API
 * (SegmentReader::inverted_index(field), term stream, read_postings(.., WithFreqs),
 * get_fieldnorms_reader, alive_bitset) and hands over one flat CSR per field, so this library never
 * parses tantivy's private files (src/db/core.rs:53-55,238-245 own them).
 * Doc ids are shard-local, dense, 0..n_docs; `doc_id_base` is added to every reported hit. */
#define FG_FIELD_HAS_FIELDNORMS 1u /* TEXT fields (src/db/schemas.rs:9-17); facet fields have none */
#define FG_FIELD_HAS_FREQS 2u      /* term_freqs given; otherwise tf == 1 (facet / Basic postings) */

typedef struct {
    uint32_t flags;
    uint32_t n_terms;
    uint64_t total_num_tokens;       /* GLOBAL over all shards/segments (Bm25Weight average fieldnorm) */
    const uint8_t* fieldnorm_ids;    /* [n_docs] tantivy fieldnorm ids, or NULL (constant norm 1) */
    const uint64_t* term_offsets;    /* [n_terms + 1] into doc_ids / term_freqs */
    const uint32_t* doc_ids;         /* strictly ascending within a term */
    const uint32_t* term_freqs;      /* or NULL */
    const uint32_t* global_doc_freq; /* [n_terms] df summed over all shards, or NULL = local df */
} fg_field_desc;

typedef struct {
    uint32_t n_docs;              /* docs of this shard = sum of max_doc of its segments */
    uint32_t doc_id_base;         /* global doc id of local doc 0 */
    uint64_t global_n_docs;       /* sum of max_doc over ALL shards; 0 = n_docs */
    uint32_t n_fields;
    uint32_t reserved;
    const fg_field_desc* fields;  /* field id = position in this array */
    const uint32_t* alive_bitset; /* [ceil(n_docs/32)] bit set = alive, or NULL = all alive */
} fg_index_desc;

typedef struct fg_index fg_index;
int32_t fg_index_upload(fg_ctx* ctx, const fg_index_desc* desc, fg_index** out);
void fg_index_release(fg_index* index);
/* Snapshot refresh after deletes only (DocumentOperations::upsert's delete_term + commit,
 * src/db/document.rs:38-41,65): a new snapshot that shares every device array
 * of `base` (postings, skip entries, fieldnorms, tf columns) and differs only in its alive bitset
 * (NULL = all alive). Costs n_docs/8 bytes of upload. Statistics (N, df, total_num_tokens) keep
 * counting deleted docs, as tantivy's do until a merge. `base` and the derived snapshot are released
 * independently, in any order; the shared arrays are freed with the last of them. */
int32_t fg_index_with_alive(fg_index* base, const uint32_t* alive_bitset, fg_index** out);
/* Snapshot refresh after a commit that added documents (DocumentOperations::upsert / add + commit,
 * src/db/document.rs:23-67,97): `segment` describes ONLY the new documents, the way a new
 * tantivy segment would (doc ids 0..segment->n_docs local to the segment; they become base.n_docs + id, as a new
 * segment's doc-id base is the sum of the earlier max_docs). Its fields are the base's, in the same order with the
 * same flags; its term ordinals extend the base's dictionary (ordinals < the base's n_terms mean the same terms, new
 * terms follow); total_num_tokens counts the segment's tokens; global_doc_freq / global_n_docs / doc_id_base must be
 * NULL / 0 (single-shard snapshots only). `alive_bitset` covers base.n_docs + segment->n_docs docs (NULL = all alive).
 * Only the segment's postings, fieldnorm ids and column bytes cross PCIe: the blocks of the base keep their payload
 * (copied device to device), new blocks go behind them (a bitmap term's partial last block is re-encoded with its new
 * postings so that every block of it but the last stays full), and norm
 * caches, idf weights and the block-max metadata of every block are recomputed on the device for the new N and
 * average field length, so the result scores exactly like a full fg_index_upload of the whole corpus. Terms keep the
 * lookup structures (tf column / membership bitmap) they had; membership is re-decided by the next full upload.
 * `base` is not modified and is released independently. */
int32_t fg_index_append(fg_index* base, const fg_index_desc* segment, const uint32_t* alive_bitset, fg_index** out);

typedef struct {
    uint64_t n_postings;
    uint64_t n_blocks;
    uint64_t packed_bytes;  /* bit-packed doc/tf payload */
    uint64_t skip_bytes;    /* 16 B per block */
    uint64_t device_bytes;  /* everything resident in HBM for this snapshot */
    uint32_t n_docs;
    uint32_t n_fields;
    uint64_t column_bytes;  /* dense tf columns of the most frequent terms (1 B per doc per column) */
    uint32_t n_columns;
    uint32_t n_bitmaps;     /* membership bitmaps (+ rank directories) of the mid-frequency terms */
    uint64_t bitmap_bytes;
    uint64_t appended_bytes_h2d; /* fg_index_append: bytes that crossed PCIe to build this snapshot (0 after a full upload) */
} fg_index_info;
int32_t fg_index_get_info(const fg_index* index, fg_index_info* out);
/* document frequency / layout of one term (host copy of the term table) */
int32_t fg_index_term_info(const fg_index* index, uint32_t field, uint32_t term_ord,
                           uint32_t* local_df, uint32_t* global_df, uint32_t* n_blocks,
                           uint64_t* packed_bytes);

/* ---- query plan ---------------------------------------------------------------------------
 * One level of grouping, which is what Dataset::search builds (src/db/search.rs:108-151):
 *   clause = (occur, leaves[]); leaves inside a clause are OR-ed and their scores summed
 *            (word -> (text:w OR name:w) expansion; the facet Should-group of build_facet_query,
 *             src/db/search.rs:221-289);
 *   query  = clauses[]; Must clauses intersect, Should clauses add score (and select docs when there
 *            is no Must clause), MustNot clauses exclude. */
#define FG_OCCUR_SHOULD 0u
#define FG_OCCUR_MUST 1u
#define FG_OCCUR_MUST_NOT 2u
#define FG_TERM_MISSING 0xFFFFFFFFu /* term absent from the dictionary: empty scorer */
#define FG_TERM_ALL 0xFFFFFFFEu     /* AllQuery leaf (src/db/search.rs:115-116,258-261): score 1.0 */

typedef struct {
    uint32_t field;
    uint32_t term_ord;
    float boost; /* 1.0 unless `term^boost` */
} fg_leaf;

typedef struct {
    uint32_t occur;
    uint32_t leaf_begin; /* into fg_query_batch.leaves */
    uint32_t n_leaves;
} fg_clause;

typedef struct {
    uint32_t k;            /* TopDocs limit = page*per_page + per_page (src/db/search.rs:154-160) */
    uint32_t clause_begin; /* into fg_query_batch.clauses */
    uint32_t n_clauses;
} fg_query;

typedef struct {
    uint32_t n_queries;
    uint32_t n_clauses;
    uint32_t n_leaves;
    uint32_t reserved;
    const fg_query* queries;
    const fg_clause* clauses;
    const fg_leaf* leaves;
} fg_query_batch;

typedef struct {
    float score;
    uint32_t doc; /* global doc id = doc_id_base + local id */
} fg_hit;

/* The reference-facing call: host buffers in, host buffers out, blocking.
 * out_hits[q*k_stride + r] = r-th best hit of query q (score desc, doc asc); out_n_hits[q] <= k;
 * out_match_count[q] (may be NULL) = number of documents matching query q on this shard. */
int32_t fg_search_batch(fg_index* index, const fg_query_batch* batch, uint32_t k_stride,
                        fg_hit* out_hits, uint32_t* out_n_hits, uint32_t* out_match_count);

/* One query whose Should children are boolean queries themselves -- tantivy: a BooleanQuery of BooleanQuerys, what the
 * QueryParser builds for `(a AND b) OR (c AND d)` or `a OR (b AND c)` (src/db/search.rs:118): the queries of `disjuncts`
 * (1 to 64; each a one-level plan as above; their own k is ignored) are the children; a document matches when any of them
 * matches and scores the SUM of the scores of those that do. Every disjunct is evaluated on the device with all its matches
 * kept, a combine step sums per document and selects the page. out_hits holds min(k, matches) entries (caller-allocated for k);
 * out_match_count (may be NULL) = distinct matching documents. Blocking, host buffers. */
int32_t fg_search_union_of(fg_index* index, const fg_query_batch* disjuncts, uint32_t k, fg_hit* out_hits,
                           uint32_t* out_n_hits, uint32_t* out_match_count);
/* The same with FILTER children -- the last n_filters queries of `disjuncts`: a document matches when at least one
 * ordinary child and every filter child match it, and scores the sum over the ordinary children plus the filters' scores.
 * This is what Dataset::search builds from a nested text query and facet filters: Bool[Must(text_query), Must(facet_query)]
 * (src/db/search.rs:140-144). Boosts must not be negative. */
int32_t fg_search_union_of_filtered(fg_index* index, const fg_query_batch* disjuncts, uint32_t n_filters, uint32_t k,
                                    fg_hit* out_hits, uint32_t* out_n_hits, uint32_t* out_match_count);

/* ---- split-phase form (batches resident in HBM; CUDA-event timing; multi-GPU merge) --------- */
typedef struct fg_batch fg_batch;
/* lowers the plan (weights from GLOBAL statistics, work items) and uploads it to the device */
int32_t fg_batch_prepare(fg_index* index, const fg_query_batch* batch, fg_batch** out);
/* same with lowering options. FG_PREP_NO_COLUMNS evaluates every leaf from its posting blocks even
 * when the term has a dense tf column: the block path is what the algorithmic-byte definition
 *  and the exact-accounting counters are stated on. */
#define FG_PREP_NO_COLUMNS 1u
/* FG_PREP_LEGACY lowers the plan for the windowed accumulator kernels of round 1 (kept for A/B runs and for
 * the exact algorithmic-byte accounting pass, which is defined on exhaustive block evaluation); the default
 * lowering targets the lead-driven kernels (block-max / MaxScore pruning, skip-table gallop lookups). */
#define FG_PREP_LEGACY 2u
/* Page limits: k <= 1024 is served from per-warp register queues (the common case); a batch whose largest k exceeds 1024
 * (deep pages: src/db/search.rs:154-160 puts no bound on page * per_page) keeps every match above the running threshold
 * of every query in it and selects the pages afterwards (radix select + sort on the device): exact, any k, but list
 * space grows with the number of matches, so callers send deep pages in batches of their own (fgh_search_batch does).
 * FG_PREP_PER_QUERY_STATUS: a query the device path cannot take (k == 0, more than 32 live leaves, a plan
 * node outside the supported shapes) does not fail the batch: it becomes an empty query (0 hits) and its status is
 * reported by fg_batch_query_status; the other queries of the batch are answered. Without the flag the first such
 * query fails the call (default lowering only; FG_PREP_LEGACY batches always fail as a whole). */
#define FG_PREP_PER_QUERY_STATUS 4u
int32_t fg_batch_prepare_ex(fg_index* index, const fg_query_batch* batch, uint32_t prep_flags, fg_batch** out);
void fg_batch_release(fg_batch* b);
/* per-query lowering status of a prepared batch, [n_queries] (all FG_OK unless FG_PREP_PER_QUERY_STATUS was given) */
int32_t fg_batch_query_status(const fg_batch* b, int32_t* out_status);
#define FG_EXEC_EXACT_ACCOUNTING 1u /* exact block-need test for the algorithmic-byte counters (slow) */
#define FG_EXEC_COUNTERS 4u         /* maintain bytes_blocks / bytes_redecode / scored_postings (cheap) */
#define FG_EXEC_NO_PRUNE 8u         /* column scan: visit every 256-doc chunk even when the query's k-th best score
                                       already exceeds what a doc without a sparse-term posting can reach (A/B switch;
                                       results are identical either way) */
#define FG_EXEC_DETERMINISTIC 2u    /* apply leaves one at a time: bit-reproducible f32 sums (slower);
                                       default sums the leaves of a clause with float atomics, which can
                                       differ in the last bit for docs with >= 3 contributions */
/* Launches the search kernels for the prepared batch on the context's stream (asynchronous).
 * d_hits [n_queries*k_stride] fg_hit, d_n_hits [n_queries], d_match_count [n_queries] or NULL:
 * DEVICE pointers. d_match_bitmap: NULL, or DEVICE [n_queries * ceil(n_docs/32)] words, zeroed by
 * the caller, in which every matching local doc id gets its bit set (parity tests). */
int32_t fg_batch_execute(fg_batch* b, uint32_t flags, uint32_t k_stride, void* d_hits,
                         void* d_n_hits, void* d_match_count, void* d_match_bitmap);

/* Asynchronous host-buffer form. fg_batch_submit launches the prepared batch and queues the
 * device->host copy of its results into page-locked staging owned by the batch; it returns at once.
 * fg_batch_collect blocks until the results have arrived and copies them into the caller's buffers
 * (same layout as fg_search_batch; out_match_count may be NULL, and must be when want_counts was 0).
 * A host thread can plan and prepare batch i+1 while batch i runs: fgh_search_batch pipelines large
 * requests this way. One submit per prepared batch. Submitted batches run on two internal streams in turn
 * (FG_SUBMIT_STREAMS, read at fg_ctx_create; 0 = the context's stream): consecutive batches overlap on the device
 * where one's persistent kernel runs out of work; each batch's results are ordered by its own fg_batch_collect. */
int32_t fg_batch_submit(fg_batch* b, uint32_t flags, uint32_t k_stride, int32_t want_counts);
int32_t fg_batch_collect(fg_batch* b, fg_hit* out_hits, uint32_t* out_n_hits, uint32_t* out_match_count);

typedef struct {
    uint64_t bytes_blocks;    /* packed bytes + 16 B skip entry of every block decoded (counted once) */
    uint64_t bytes_redecode;  /* bytes of blocks decoded again (round/work-item boundaries) */
    uint64_t scored_postings; /* (doc, leaf) pairs scored = 1 B fieldnorm gathers */
    uint64_t n_work_items;
    uint64_t n_launches;      /* kernels launched by the last fg_batch_execute */
    uint64_t n_queries;
    uint64_t sum_k;           /* sum of k over queries (8 B result per hit) */
    float search_kernel_ms;   /* CUDA-event time of the search kernel of the last execute */
    float merge_kernel_ms;    /* ... and of the per-query merge kernel */
    uint64_t colscan_chunks;         /* with FG_EXEC_COUNTERS: 256-doc chunks of windowed column-scan plans reached */
    uint64_t colscan_chunks_skipped; /* ... of which skipped because no doc in them could enter the top-k */
    /* lead-driven kernels, with FG_EXEC_COUNTERS: bytes_blocks = payload + skip entry of every block decoded
     * (lead or lookup), scored_postings = 1-byte gathers (fieldnorm ids, tf-column bytes), and: */
    uint64_t bytes_meta;       /* block-max words and skip entries read while skipping / galloping */
    uint64_t lead_blocks;      /* lead blocks decoded ... */
    uint64_t lead_blocks_seen; /* ... of the lead blocks whose block maximum was tested */
    uint64_t plan_bytes;       /* bytes of the lowered plan fg_batch_prepare uploaded (queries, leaves, work-item records) */
} fg_batch_stats;
/* synchronises the stream and reads the device counters of the last fg_batch_execute */
int32_t fg_batch_get_stats(fg_batch* b, fg_batch_stats* out);

/* Merge of per-shard results after the all-gather: `d_gathered_hits` is
 * [n_ranks][n_queries][k_stride] fg_hit, `d_gathered_n` is [n_ranks][n_queries]; writes the k best
 * per query under (score desc, doc asc) to d_out_hits [n_queries][k_stride] / d_out_n. All DEVICE
 * pointers; asynchronous on the context's stream. */
int32_t fg_merge_topk_device(fg_ctx* ctx, const void* d_gathered_hits, const void* d_gathered_n,
                             uint32_t n_ranks, uint32_t n_queries, uint32_t k, uint32_t k_stride,
                             void* d_out_hits, void* d_out_n);

/* ---- multi-GPU, one process per GPU --------------------------------------------
 * Documents are sharded by contiguous doc-id range; rank r uploads its shard with doc_id_base = first doc of the
 * shard and the GLOBAL statistics (global_n_docs, global_doc_freq, total_num_tokens), so weights and norm caches
 * are identical on every rank. One exchange per batch, inside the library: local top-k -> ncclAllGather of the
 * per-shard lists over NVLink -> on-device select of the k best per query (score desc, global doc id asc). Every
 * rank ends up with the global result. The communicator is NCCL's (libnccl.so.2, loaded on first use); the host
 * distributes the 128-byte id of rank 0 to the other ranks by whatever transport it has (the reference is a
 * single-process server, src/main.rs:11: there is no transport to mirror). */
typedef struct fg_comm fg_comm;
#define FG_COMM_ID_BYTES 128
int32_t fg_comm_unique_id(void* out_id);
int32_t fg_comm_create(fg_ctx* ctx, int32_t rank, int32_t n_ranks, const void* id, fg_comm** out);
void fg_comm_destroy(fg_comm* comm);
/* sum over all ranks, in place, HOST buffers (global statistics at upload time); collective */
int32_t fg_comm_allreduce_sum_u64(fg_comm* comm, uint64_t* values, size_t n);
int32_t fg_comm_allreduce_sum_u32(fg_comm* comm, uint32_t* values, size_t n);
/* Collective: executes the prepared batch on this rank's shard, all-gathers the per-shard lists and merges them.
 * d_hits [n_queries*k_stride] fg_hit and d_n_hits [n_queries] are DEVICE pointers receiving the GLOBAL result on
 * every rank; asynchronous on the context's stream. Every rank must pass a batch prepared from the same queries. */
int32_t fg_batch_execute_sharded(fg_batch* b, fg_comm* comm, uint32_t flags, uint32_t k_stride, void* d_hits, void* d_n_hits);
/* fg_batch_submit for a sharded index (collective): local top-k -> all-gather -> merge, then the device->host copy of
 * the GLOBAL result into the batch's staging; returns at once, fg_batch_collect (out_match_count NULL) hands it out. */
int32_t fg_batch_submit_sharded(fg_batch* b, fg_comm* comm, uint32_t flags, uint32_t k_stride);
int32_t fg_comm_info(const fg_comm* comm, int32_t* rank, int32_t* n_ranks);
/* all-gather of `bytes` bytes per rank between HOST buffers (recv holds n_ranks * bytes, in rank order); collective.
 * The host layer shares the planning of a request among the ranks with it. */
int32_t fg_comm_allgather_bytes(fg_comm* comm, const void* send, size_t bytes, void* recv);

/* ---- scoring helpers shared with the host planner (tantivy fieldnorm / Bm25Weight) ---------- */
uint8_t fg_fieldnorm_to_id(uint32_t num_tokens);
uint32_t fg_id_to_fieldnorm(uint8_t id);
float fg_bm25_idf(uint64_t doc_freq, uint64_t doc_count);

#ifdef __cplusplus
}
#endif
#endif /* FUGU_GPU_H */
