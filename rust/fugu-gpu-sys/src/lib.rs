//! Raw declarations of include/fugu_gpu.h and include/fugu_host.h (NOT compiled in the build image: no toolchain).
//! Layouts are `#[repr(C)]` mirrors of the headers; sizes in comments are asserted by tests/test_abi.py.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

macro_rules! opaque { ($($n:ident),*) => { $(#[repr(C)] pub struct $n { _p: [u8; 0] })* } }
opaque!(fg_ctx, fg_index, fg_batch, fg_comm, fgh_dataset);

pub const FG_OK: i32 = 0;
pub const FG_ERR_INVALID: i32 = -1;
pub const FG_ERR_UNSUPPORTED: i32 = -2; // keep tantivy for this query
pub const FG_ERR_CUDA: i32 = -3;
pub const FG_ERR_OOM: i32 = -4;
pub const FG_ERR_NO_DEVICE: i32 = -5;

pub const FG_FIELD_HAS_FIELDNORMS: u32 = 1;
pub const FG_FIELD_HAS_FREQS: u32 = 2;
pub const FG_OCCUR_SHOULD: u32 = 0;
pub const FG_OCCUR_MUST: u32 = 1;
pub const FG_OCCUR_MUST_NOT: u32 = 2;
pub const FG_TERM_MISSING: u32 = 0xFFFF_FFFF;
pub const FG_TERM_ALL: u32 = 0xFFFF_FFFE;
pub const FG_PREP_NO_COLUMNS: u32 = 1;
pub const FG_PREP_LEGACY: u32 = 2;
pub const FG_PREP_PER_QUERY_STATUS: u32 = 4;
pub const FG_EXEC_NO_PRUNE: u32 = 8;
pub const FG_COMM_ID_BYTES: usize = 128;

#[repr(C)]
pub struct fg_field_desc {
    // 56 bytes
    pub flags: u32,
    pub n_terms: u32,
    pub total_num_tokens: u64,
    pub fieldnorm_ids: *const u8,
    pub term_offsets: *const u64,
    pub doc_ids: *const u32,
    pub term_freqs: *const u32,
    pub global_doc_freq: *const u32,
}
#[repr(C)]
pub struct fg_index_desc {
    // 40 bytes
    pub n_docs: u32,
    pub doc_id_base: u32,
    pub global_n_docs: u64,
    pub n_fields: u32,
    pub reserved: u32,
    pub fields: *const fg_field_desc,
    pub alive_bitset: *const u32,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct fg_leaf { pub field: u32, pub term_ord: u32, pub boost: f32 }
#[repr(C)]
#[derive(Clone, Copy)]
pub struct fg_clause { pub occur: u32, pub leaf_begin: u32, pub n_leaves: u32 }
#[repr(C)]
#[derive(Clone, Copy)]
pub struct fg_query { pub k: u32, pub clause_begin: u32, pub n_clauses: u32 }
#[repr(C)]
pub struct fg_query_batch {
    // 40 bytes
    pub n_queries: u32,
    pub n_clauses: u32,
    pub n_leaves: u32,
    pub reserved: u32,
    pub queries: *const fg_query,
    pub clauses: *const fg_clause,
    pub leaves: *const fg_leaf,
}
#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct fg_hit { pub score: f32, pub doc: u32 }

extern "C" {
    pub fn fg_last_error() -> *const c_char;
    pub fn fg_version() -> *const c_char;
    pub fn fg_ctx_create(device: i32, out: *mut *mut fg_ctx) -> i32;
    pub fn fg_ctx_destroy(ctx: *mut fg_ctx);
    pub fn fg_ctx_set_stream(ctx: *mut fg_ctx, cuda_stream: *mut c_void) -> i32;
    pub fn fg_ctx_synchronize(ctx: *mut fg_ctx) -> i32;
    pub fn fg_index_upload(ctx: *mut fg_ctx, desc: *const fg_index_desc, out: *mut *mut fg_index) -> i32;
    pub fn fg_index_with_alive(base: *mut fg_index, alive_bitset: *const u32, out: *mut *mut fg_index) -> i32;
    pub fn fg_index_release(index: *mut fg_index);
    /// The reference-facing call (replaces src/db/search.rs:162): host buffers in, host buffers out, blocking.
    pub fn fg_search_batch(index: *mut fg_index, batch: *const fg_query_batch, k_stride: u32,
                           out_hits: *mut fg_hit, out_n_hits: *mut u32, out_match_count: *mut u32) -> i32;
    // asynchronous host-buffer form (the micro-batcher keeps several batches in flight)
    pub fn fg_batch_prepare_ex(index: *mut fg_index, batch: *const fg_query_batch, prep_flags: u32, out: *mut *mut fg_batch) -> i32;
    pub fn fg_batch_query_status(b: *const fg_batch, out_status: *mut i32) -> i32;
    pub fn fg_batch_submit(b: *mut fg_batch, flags: u32, k_stride: u32, want_counts: i32) -> i32;
    pub fn fg_batch_collect(b: *mut fg_batch, out_hits: *mut fg_hit, out_n_hits: *mut u32, out_match_count: *mut u32) -> i32;
    pub fn fg_batch_release(b: *mut fg_batch);
    // multi-GPU, one process per GPU
    pub fn fg_comm_unique_id(out_id: *mut c_void) -> i32;
    pub fn fg_comm_create(ctx: *mut fg_ctx, rank: i32, n_ranks: i32, id: *const c_void, out: *mut *mut fg_comm) -> i32;
    pub fn fg_comm_destroy(comm: *mut fg_comm);
    pub fn fg_comm_allreduce_sum_u64(comm: *mut fg_comm, values: *mut u64, n: usize) -> i32;
    pub fn fg_comm_allreduce_sum_u32(comm: *mut fg_comm, values: *mut u32, n: usize) -> i32;
    pub fn fg_batch_execute_sharded(b: *mut fg_batch, comm: *mut fg_comm, flags: u32, k_stride: u32,
                                    d_hits: *mut c_void, d_n_hits: *mut c_void) -> i32;
    pub fn fg_fieldnorm_to_id(num_tokens: u32) -> u8;
    pub fn fg_id_to_fieldnorm(id: u8) -> u32;
    pub fn fg_bm25_idf(doc_freq: u64, doc_count: u64) -> f32;
    // one query whose Should children are boolean queries themselves: `(a AND b) OR (c AND d)`
    pub fn fg_search_union_of(index: *mut fg_index, disjuncts: *const fg_query_batch, k: u32, out_hits: *mut fg_hit,
                              out_n_hits: *mut u32, out_match_count: *mut u32) -> i32;
    pub fn fg_search_union_of_filtered(index: *mut fg_index, disjuncts: *const fg_query_batch, n_filters: u32, k: u32, out_hits: *mut fg_hit,
                                       out_n_hits: *mut u32, out_match_count: *mut u32) -> i32;
    // snapshot refresh after a commit that added documents: only the new segment crosses PCIe
    pub fn fg_index_append(base: *mut fg_index, segment: *const fg_index_desc, alive_bitset: *const u32, out: *mut *mut fg_index) -> i32;
    // sharded submit (collective), communicator info, host-buffer all-gather (shared planning)
    pub fn fg_batch_submit_sharded(b: *mut fg_batch, comm: *mut fg_comm, flags: u32, k_stride: u32) -> i32;
    pub fn fg_comm_info(comm: *const fg_comm, rank: *mut i32, n_ranks: *mut i32) -> i32;
    pub fn fg_comm_allgather_bytes(comm: *mut fg_comm, send: *const c_void, bytes: usize, recv: *mut c_void) -> i32;
}

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct fgh_batcher_stats { pub n_requests: u64, pub n_batches: u64, pub max_batch_seen: u64, pub wait_us_total: u64 }
opaque!(fgh_batcher);

// libfugu_host.so: no CUDA code; device entry points are bound on first use (FG_ERR_NO_DEVICE when they are missing)
#[link(name = "fugu_host")]
extern "C" {
    pub fn fgh_last_error() -> *const c_char;
    pub fn fgh_dataset_commit_counts(ds: *const fgh_dataset, n_full_uploads: *mut u64, n_appends: *mut u64) -> i32;
    pub fn fgh_search_batch_sharded(ds: *mut fgh_dataset, comm: *mut fg_comm, n: u32, queries: *const *const c_char,
                                    filters: *const *const c_char, filter_offsets: *const u32, pages: *const u32,
                                    per_pages: *const u32, stride: u32, out_hits: *mut fg_hit, out_n: *mut u32, status: *mut i32) -> i32;
    // host-side merge of per-shard result lists (what fgh_search_batch_sharded uses for deep pages and nested queries)
    pub fn fgh_merge_shard_pages(lists: *const *const fg_hit, lens: *const u32, n_lists: u32, page: u32, per_page: u32,
                                 out_hits: *mut fg_hit) -> u32;
    // micro-batcher for the one-query-per-request handlers (src/server/handlers/search.rs:152): blocking, any thread
    pub fn fgh_batcher_create(ds: *mut fgh_dataset, max_batch: u32, max_wait_us: u32, out: *mut *mut fgh_batcher) -> i32;
    pub fn fgh_batcher_destroy(b: *mut fgh_batcher);
    pub fn fgh_batcher_search(b: *mut fgh_batcher, query: *const c_char, filters: *const *const c_char, n_filters: u32,
                              page: u32, per_page: u32, out_hits: *mut fg_hit, out_n: *mut u32) -> i32;
    pub fn fgh_batcher_get_stats(b: *mut fgh_batcher, out: *mut fgh_batcher_stats) -> i32;

    // include/fugu_host.h (libfugu_host.so): the planning half of Dataset::search + a batched search entry
    pub fn fgh_dataset_create(ctx: *mut fg_ctx, out: *mut *mut fgh_dataset) -> i32;
    pub fn fgh_dataset_destroy(ds: *mut fgh_dataset);
    pub fn fgh_dataset_upsert(ds: *mut fgh_dataset, id: *const c_char, text: *const c_char, name: *const c_char,
                              facets: *const *const c_char, n_facets: u32) -> i32;
    pub fn fgh_dataset_delete(ds: *mut fgh_dataset, id: *const c_char) -> i32;
    pub fn fgh_dataset_commit(ds: *mut fgh_dataset) -> i32;
    pub fn fgh_dataset_doc_id(ds: *const fgh_dataset, doc: u32, buf: *mut c_char, cap: u32) -> i32;
    pub fn fgh_search_batch(ds: *mut fgh_dataset, n: u32, queries: *const *const c_char, filters: *const *const c_char,
                            filter_offsets: *const u32, pages: *const u32, per_pages: *const u32, stride: u32,
                            out_hits: *mut fg_hit, out_n: *mut u32, out_match_count: *mut u32, status: *mut i32) -> i32;
}
