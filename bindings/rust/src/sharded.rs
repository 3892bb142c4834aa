//! Segments sharded over the GPUs of one box: the cross-segment merge of `BatchResultAggregator`
//! (lib/shard/src/search_result_aggregator.rs:50-117) happens on the devices.  SOURCE ONLY (see ffi.rs).
//!
//! One `B200Shard` per GPU; `SegmentsSearcher` keeps spawning one blocking task per segment (segments_searcher.rs:255) and every task
//! calls `search` with the same queries — the call is a collective: local fused scan, lists exchanged through peer-mapped buffers over
//! NVLink, merged on every GPU, every task returns the same global top-k (so the host-side aggregation becomes a no-op).
use common::types::ScoredPointOffset;

use super::ffi::*;
use super::raw_scorer::{last_error, B200Storage};
use crate::common::operation_error::{OperationError, OperationResult};

pub struct B200Shard { storage: B200Storage, comm: *mut qb_comm }
unsafe impl Send for B200Shard {}
unsafe impl Sync for B200Shard {}
impl Drop for B200Shard { fn drop(&mut self) { unsafe { qb_comm_destroy(self.comm) } } }

/// All shards live in this process: create the communicators and wire them up with direct peer access.
pub fn connect_shards(storages: Vec<(i32 /* device */, B200Storage, u32 /* id_base */)>, max_queries: u32, max_top: u32) -> OperationResult<Vec<B200Shard>> {
    let world = storages.len() as i32;
    let mut shards = Vec::with_capacity(storages.len());
    for (rank, (device, storage, id_base)) in storages.into_iter().enumerate() {
        let mut comm = std::ptr::null_mut();
        if unsafe { qb_comm_create(device, rank as i32, world, max_queries, max_top, &mut comm) } != QB_OK { return Err(OperationError::service_error(last_error())); }
        unsafe { qb_storage_set_id_base(storage.raw, id_base) };
        shards.push(B200Shard { storage, comm });
    }
    let comms: Vec<*mut qb_comm> = shards.iter().map(|s| s.comm).collect();
    if unsafe { qb_comm_connect_local(comms.as_ptr(), world) } != QB_OK { return Err(OperationError::service_error(last_error())); }
    Ok(shards)
}

impl B200Shard {
    /// Called by every shard's task with the same `queries` / `top`.
    pub fn search(&self, queries: &[f32], n_queries: usize, top: usize) -> Vec<Vec<ScoredPointOffset>> {
        let mut out = vec![qb_scored_point::default(); n_queries * top];
        let mut counts = vec![0u32; n_queries];
        let st = unsafe {
            qb_multi_search_batch(self.comm, self.storage.raw, queries.as_ptr(), n_queries as u32, top as u32, std::ptr::null(), std::ptr::null(), out.as_mut_ptr(),
                                  counts.as_mut_ptr(), std::ptr::null_mut())
        };
        assert!(st == QB_OK, "{}", last_error());
        (0..n_queries).map(|q| out[q * top..q * top + counts[q] as usize].iter().map(|p| ScoredPointOffset { idx: p.idx, score: p.score }).collect()).collect()
    }

    /// Device-resident, PIPELINED step (queries and results stay in HBM): the scan is enqueued on the storage's stream, the exchange + merge
    /// on the communicator's; the next step's scan does not wait for this step's merge.  `drain()` before reading `dev_out`.
    pub unsafe fn search_device(&self, dev_queries: *const f32, n_queries: usize, top: usize, dev_out: *mut qb_scored_point, dev_counts: *mut u32) -> OperationResult<()> {
        let st = qb_multi_search_batch_device(self.comm, self.storage.raw, dev_queries, n_queries as u32, top as u32, std::ptr::null_mut(), std::ptr::null_mut(), dev_out,
                                              dev_counts);
        if st != QB_OK { return Err(OperationError::service_error(last_error())); }
        Ok(())
    }

    /// Waits for every exchange + merge enqueued so far; a peer that never made the matching call surfaces here as an error.
    pub fn drain(&self) -> OperationResult<()> {
        if unsafe { qb_comm_check(self.comm) } != QB_OK { return Err(OperationError::service_error(last_error())); }
        Ok(())
    }
}
