//! Micro-batcher. The HTTP API is one query per request
//! (/root/reference/src/server/handlers/search.rs:152, :210); the GPU wants thousands of queries per launch. Requests
//! that arrive within `window` (or until `max_batch`) are lowered into ONE `fg_query_batch`, submitted with the
//! asynchronous host-buffer calls (`fg_batch_prepare_ex` -> `fg_batch_submit` -> `fg_batch_collect`), and every
//! caller gets its own slice of the result. A query the device cannot take (`FG_PREP_PER_QUERY_STATUS`) fails alone:
//! its caller falls back to tantivy, its siblings are answered. NOT compiled in the build image (no toolchain).
use crate::{check, GpuError, Plan, Snapshot};
use arc_swap::ArcSwap;
use fugu_gpu_sys as sys;
use std::{ptr, sync::Arc, time::Duration};
use tantivy::{DocAddress, Score};
use tokio::sync::{mpsc, oneshot};

type Reply = oneshot::Sender<Result<Vec<(Score, DocAddress)>, GpuError>>;
struct Request { plan: Plan, reply: Reply }

#[derive(Clone)]
pub struct Batcher { tx: mpsc::Sender<Request> }

impl Batcher {
    /// `snapshot` is swapped by the commit hook; a batch runs on the snapshot it started on.
    pub fn spawn(snapshot: Arc<ArcSwap<Snapshot>>, window: Duration, max_batch: usize) -> Self {
        let (tx, mut rx) = mpsc::channel::<Request>(max_batch * 4);
        tokio::spawn(async move {
            while let Some(first) = rx.recv().await {
                let mut reqs = vec![first];
                let deadline = tokio::time::sleep(window);
                tokio::pin!(deadline);
                while reqs.len() < max_batch {
                    tokio::select! {
                        _ = &mut deadline => break,
                        r = rx.recv() => match r { Some(r) => reqs.push(r), None => break },
                    }
                }
                let snap = snapshot.load_full();
                // the device call blocks: keep it off the async workers
                let _ = tokio::task::spawn_blocking(move || run_batch(&snap, reqs)).await;
            }
        });
        Self { tx }
    }

    /// What `Dataset::search` awaits instead of calling `searcher.search` (src/db/search.rs:162).
    pub async fn search(&self, plan: Plan) -> Result<Vec<(Score, DocAddress)>, GpuError> {
        let (reply, rx) = oneshot::channel();
        self.tx.send(Request { plan, reply }).await.map_err(|_| GpuError { code: sys::FG_ERR_INVALID, message: "batcher stopped".into() })?;
        rx.await.map_err(|_| GpuError { code: sys::FG_ERR_INVALID, message: "batcher dropped the request".into() })?
    }
}

fn run_batch(snap: &Snapshot, reqs: Vec<Request>) {
    // concatenate the plans: clause / leaf indices become batch-global
    let (mut queries, mut clauses, mut leaves) = (Vec::new(), Vec::new(), Vec::new());
    let mut kmax = 1u32;
    for r in &reqs {
        let (c0, l0) = (clauses.len() as u32, leaves.len() as u32);
        queries.push(sys::fg_query { k: r.plan.k, clause_begin: c0, n_clauses: r.plan.clauses.len() as u32 });
        clauses.extend(r.plan.clauses.iter().map(|c| sys::fg_clause { leaf_begin: c.leaf_begin + l0, ..*c }));
        leaves.extend_from_slice(&r.plan.leaves);
        kmax = kmax.max(r.plan.k);
    }
    let qb = sys::fg_query_batch { n_queries: queries.len() as u32, n_clauses: clauses.len() as u32, n_leaves: leaves.len() as u32,
                                   reserved: 0, queries: queries.as_ptr(), clauses: clauses.as_ptr(), leaves: leaves.as_ptr() };
    let n = reqs.len();
    let result = (|| -> Result<(Vec<sys::fg_hit>, Vec<u32>, Vec<i32>), GpuError> {
        let mut b = ptr::null_mut();
        check(unsafe { sys::fg_batch_prepare_ex(snap.raw(), &qb, sys::FG_PREP_PER_QUERY_STATUS, &mut b) })?;
        struct Guard(*mut sys::fg_batch);
        impl Drop for Guard { fn drop(&mut self) { unsafe { sys::fg_batch_release(self.0) } } }
        let _g = Guard(b);
        let mut status = vec![0i32; n];
        check(unsafe { sys::fg_batch_query_status(b, status.as_mut_ptr()) })?;
        check(unsafe { sys::fg_batch_submit(b, 0, kmax, 0) })?; // no match counts: the TopDocs form
        let mut hits = vec![sys::fg_hit::default(); n * kmax as usize];
        let mut nh = vec![0u32; n];
        check(unsafe { sys::fg_batch_collect(b, hits.as_mut_ptr(), nh.as_mut_ptr(), ptr::null_mut()) })?;
        Ok((hits, nh, status))
    })();
    match result {
        Ok((hits, nh, status)) => for (i, r) in reqs.into_iter().enumerate() {
            let out = if status[i] != sys::FG_OK {
                Err(GpuError { code: status[i], message: "query not evaluated on the device".into() }) // caller: tantivy
            } else {
                let row = &hits[i * kmax as usize..][..nh[i] as usize];
                Ok(row.iter().map(|h| (h.score, snap.doc_address(h.doc))).collect())
            };
            let _ = r.reply.send(out);
        },
        Err(e) => for r in reqs { let _ = r.reply.send(Err(GpuError { code: e.code, message: e.message.clone() })); },
    }
}
