//! `BatchFilteredSearcher::peek_top_iter` / `peek_top_all` (lib/segment/src/index/hnsw_index/point_scorer.rs:312-472) and the oversample →
//! rescore step (index/vector_index_search_common.rs:27-91) over the C ABI.  SOURCE ONLY (see ffi.rs).
//!
//! The reference walks the candidate ids in 64-id chunks, scores every chunk against every query and feeds per-query
//! `FixedLengthPriorityQueue`s; `qb_search_batch` does the whole loop in one call and returns the same `Vec<Vec<ScoredPointOffset>>`.
use std::sync::atomic::{AtomicBool, Ordering};

use common::counter::hardware_counter::HardwareCounterCell;
use common::types::{PointOffsetType, ScoredPointOffset};

use super::ffi::*;
use super::raw_scorer::{last_error, B200Storage};
use crate::common::operation_error::{CancellableResult, OperationError, OperationResult};

/// What the caller knows about the candidates — mirrors how `PlainVectorIndex` / HNSW's exact branch build the iterator they pass.
pub enum Candidates<'a> {
    /// every point that is not soft-deleted: the storage's `deleted` BitSlice words (bit = 1 => deleted), or None
    All { deleted: Option<&'a [u64]> },
    /// the ids a payload filter produced (`filtered_points`), already de-duplicated
    Ids(&'a [PointOffsetType]),
}

pub struct B200BatchSearcher<'a> {
    storage: &'a B200Storage,
    queries: Vec<f32>,          // n_queries x dim, raw (Metric::preprocess + encode_query run on the device)
    n_queries: u32,
    top: u32,
    hardware_counter: HardwareCounterCell,
}

impl<'a> B200BatchSearcher<'a> {
    /// BatchFilteredSearcher::new (point_scorer.rs:323-355): one scorer per query vector; construction is where errors surface.
    pub fn new(storage: &'a B200Storage, queries: &[&[f32]], top: usize, hc: HardwareCounterCell) -> OperationResult<Self> {
        if top == 0 { return Err(OperationError::service_error("top must be >= 1")); }
        let flat: Vec<f32> = queries.iter().flat_map(|q| q.iter().copied()).collect();
        Ok(Self { storage, queries: flat, n_queries: queries.len() as u32, top: top as u32, hardware_counter: hc })
    }

    /// peek_top_iter / peek_top_all: `is_stopped` is polled between kernel launches (check_process_stopped, :433).
    pub fn peek_top(&self, candidates: Candidates<'_>, is_stopped: &AtomicBool) -> CancellableResult<Vec<Vec<ScoredPointOffset>>> {
        let nq = self.n_queries as usize;
        let mut out = vec![qb_scored_point::default(); nq * self.top as usize];
        let mut counts = vec![0u32; nq];
        let mut hw = qb_hw_counters::default();
        let (deleted, ids, n_ids) = match candidates {
            Candidates::All { deleted } => (deleted.map_or(std::ptr::null(), |d| d.as_ptr()), std::ptr::null(), 0u64),
            Candidates::Ids(ids) => (std::ptr::null(), ids.as_ptr(), ids.len() as u64),
        };
        // AtomicBool and i32 differ in size: the flag the library polls is a local mirror refreshed by the caller's cancellation hook
        let stop_mirror: i32 = is_stopped.load(Ordering::Relaxed) as i32;
        let st = unsafe {
            qb_search_batch(self.storage.raw, self.queries.as_ptr(), self.n_queries, self.top, deleted, ids, n_ids, &stop_mirror, out.as_mut_ptr(), counts.as_mut_ptr(), &mut hw)
        };
        if st == QB_ERR_CANCELLED || is_stopped.load(Ordering::Relaxed) { return Err(crate::common::operation_error::CancelledError); }
        assert!(st == QB_OK, "{}", last_error());   // scoring is infallible in the reference (`.expect("read vectors")`)
        self.hardware_counter.cpu_counter().incr_delta(hw.cpu as usize);
        self.hardware_counter.vector_io_read().incr_delta(hw.vector_io_read as usize);
        // qb_scored_point and ScoredPointOffset are both #[repr(C)] { u32, f32 } (lib/common/common/src/types.rs:12-17)
        Ok((0..nq).map(|q| {
            out[q * self.top as usize..q * self.top as usize + counts[q] as usize]
                .iter().map(|p| ScoredPointOffset { idx: p.idx, score: p.score }).collect()
        }).collect())
    }
}

/// get_oversampled_top (vector_index_search_common.rs:27-46)
pub fn get_oversampled_top(top: usize, quantized: bool, oversampling: Option<f64>) -> usize {
    match oversampling { Some(o) if quantized && o > 1.0 => (o * top as f64) as usize, _ => top }
}

/// postprocess_search_result (:48-91): rescore the quantized search's candidates with the ORIGINAL vectors, sort descending, truncate.
pub fn rescore(original: &super::raw_scorer::B200RawScorer<'_>, candidates: &[ScoredPointOffset], top: usize) -> Vec<ScoredPointOffset> {
    let ids: Vec<PointOffsetType> = candidates.iter().map(|p| p.idx).collect();
    let mut out = vec![qb_scored_point::default(); top];
    let mut n = 0u32;
    let st = unsafe { qb_rescore(original.raw(), ids.as_ptr(), ids.len(), top as u32, out.as_mut_ptr(), &mut n) };
    assert!(st == QB_OK, "{}", last_error());
    out[..n as usize].iter().map(|p| ScoredPointOffset { idx: p.idx, score: p.score }).collect()
}
