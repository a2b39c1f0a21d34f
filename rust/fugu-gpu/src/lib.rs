//! Safe wrapper around libfugu_gpu.so for the fugu server (NOT compiled in the build image: no toolchain).
//!
//! * `Context` / `Snapshot`: one CUDA context per process and device, one immutable HBM snapshot per commit
//!   (mirrors a tantivy `Searcher`; swapped with `ArcSwap` after `commit()`, /root/reference/src/db/document.rs:65,97).
//! * `lower()`: `tantivy::query::Query` tree -> flat plan (one level of grouping), `None` for shapes the device does not
//!   evaluate (phrase, range, fuzzy, OR of AND groups): the caller keeps tantivy for those.
//! * `search_top_docs()`: the replacement of `searcher.search(&q, &TopDocs::with_limit(n))` at src/db/search.rs:162.
//! * `batcher`: micro-batcher for the one-query-per-request HTTP handlers.
pub mod batcher;

use fugu_gpu_sys as sys;
use std::ffi::CStr;
use std::ptr;
use std::sync::Arc;
use tantivy::{DocAddress, Score};

#[derive(Debug)]
pub struct GpuError { pub code: i32, pub message: String }
impl std::fmt::Display for GpuError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result { write!(f, "fugu_gpu error {}: {}", self.code, self.message) }
}
impl std::error::Error for GpuError {}

pub(crate) fn check(rc: i32) -> Result<(), GpuError> {
    if rc == sys::FG_OK { return Ok(()); }
    // (calls into libfugu_host.so report through fgh_last_error, which also carries device-library failures)
    let message = unsafe { CStr::from_ptr(sys::fg_last_error()) }.to_string_lossy().into_owned();
    Err(GpuError { code: rc, message })
}

pub struct Context { raw: *mut sys::fg_ctx }
unsafe impl Send for Context {}
unsafe impl Sync for Context {} // the ABI is thread-safe: calls on one context serialise on its stream
impl Context {
    pub fn new(device: i32) -> Result<Arc<Self>, GpuError> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::fg_ctx_create(device, &mut raw) })?; // FG_ERR_NO_DEVICE without a GPU: there is no CPU fallback
        Ok(Arc::new(Self { raw }))
    }
    pub(crate) fn raw(&self) -> *mut sys::fg_ctx { self.raw }
}
impl Drop for Context { fn drop(&mut self) { unsafe { sys::fg_ctx_destroy(self.raw) } } }

/// Flat CSR of one field, as the loader collects it from tantivy's public per-segment API
/// (`SegmentReader::inverted_index(field)`, term stream, `read_postings(.., WithFreqs)`, `get_fieldnorms_reader`).
pub struct FieldCsr {
    pub has_fieldnorms: bool,
    pub total_num_tokens: u64,
    pub fieldnorm_ids: Vec<u8>,
    pub term_offsets: Vec<u64>,
    pub doc_ids: Vec<u32>,
    pub term_freqs: Option<Vec<u32>>,
    /// sorted term bytes -> ordinal: the host-side term dictionary of this snapshot
    pub terms: Vec<Vec<u8>>,
}

pub struct Snapshot {
    raw: *mut sys::fg_index,
    _ctx: Arc<Context>,
    /// global doc id = segment base + local id; kept for hydration (`searcher.doc(addr)`, src/db/search.rs:173)
    pub segment_bases: Vec<u32>,
    pub fields: Vec<FieldCsr>,
}
unsafe impl Send for Snapshot {}
unsafe impl Sync for Snapshot {}
impl Snapshot {
    pub fn upload(ctx: Arc<Context>, n_docs: u32, segment_bases: Vec<u32>, fields: Vec<FieldCsr>, alive: Option<&[u32]>) -> Result<Arc<Self>, GpuError> {
        let descs: Vec<sys::fg_field_desc> = fields.iter().map(|f| sys::fg_field_desc {
            flags: if f.has_fieldnorms { sys::FG_FIELD_HAS_FIELDNORMS } else { 0 } | if f.term_freqs.is_some() { sys::FG_FIELD_HAS_FREQS } else { 0 },
            n_terms: (f.term_offsets.len() - 1) as u32,
            total_num_tokens: f.total_num_tokens,
            fieldnorm_ids: if f.has_fieldnorms { f.fieldnorm_ids.as_ptr() } else { ptr::null() },
            term_offsets: f.term_offsets.as_ptr(),
            doc_ids: f.doc_ids.as_ptr(),
            term_freqs: f.term_freqs.as_ref().map_or(ptr::null(), |v| v.as_ptr()),
            global_doc_freq: ptr::null(),
        }).collect();
        let desc = sys::fg_index_desc { n_docs, doc_id_base: 0, global_n_docs: 0, n_fields: descs.len() as u32, reserved: 0,
                                        fields: descs.as_ptr(), alive_bitset: alive.map_or(ptr::null(), |a| a.as_ptr()) };
        let mut raw = ptr::null_mut();
        check(unsafe { sys::fg_index_upload(ctx.raw(), &desc, &mut raw) })?;
        Ok(Arc::new(Self { raw, _ctx: ctx, segment_bases, fields }))
    }
    pub(crate) fn raw(&self) -> *mut sys::fg_index { self.raw }
    pub fn term_ord(&self, field: usize, term: &[u8]) -> u32 {
        self.fields[field].terms.binary_search_by(|t| t.as_slice().cmp(term)).map_or(sys::FG_TERM_MISSING, |i| i as u32)
    }
    pub fn doc_address(&self, doc: u32) -> DocAddress {
        let seg = self.segment_bases.partition_point(|&b| b <= doc) - 1;
        DocAddress::new(seg as u32, doc - self.segment_bases[seg])
    }
}
impl Drop for Snapshot { fn drop(&mut self) { unsafe { sys::fg_index_release(self.raw) } } }

/// One query in the flat plan form that crosses the ABI.
#[derive(Default, Clone)]
pub struct Plan { pub k: u32, pub clauses: Vec<sys::fg_clause>, pub leaves: Vec<sys::fg_leaf> }

/// What replaces `searcher.search(&base_query, &TopDocs::with_limit(search_limit))` (src/db/search.rs:162) for one
/// request. Blocking: call it under `tokio::task::spawn_blocking` / `block_in_place`, or go through the batcher.
pub fn search_top_docs(snap: &Snapshot, plan: &Plan) -> Result<Vec<(Score, DocAddress)>, GpuError> {
    let q = sys::fg_query { k: plan.k, clause_begin: 0, n_clauses: plan.clauses.len() as u32 };
    let b = sys::fg_query_batch { n_queries: 1, n_clauses: q.n_clauses, n_leaves: plan.leaves.len() as u32, reserved: 0,
                                  queries: &q, clauses: plan.clauses.as_ptr(), leaves: plan.leaves.as_ptr() };
    let mut hits = vec![sys::fg_hit::default(); plan.k as usize];
    let mut n = 0u32;
    // out_match_count = NULL: TopDocs does not count matches, and only this form lets the kernels prune
    check(unsafe { sys::fg_search_batch(snap.raw(), &b, plan.k, hits.as_mut_ptr(), &mut n, ptr::null_mut()) })?;
    Ok(hits[..n as usize].iter().map(|h| (h.score, snap.doc_address(h.doc))).collect())
}
