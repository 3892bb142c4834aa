//! `impl RawScorer` over the C ABI — what `lib/segment/src/vector_storage/raw_scorer.rs:39-54` would dispatch to for
//! a GPU-resident storage.  SOURCE ONLY (see ffi.rs).  HNSW traversal (`GraphLayers::search`, graph_layers.rs:530-561)
//! and the plain index call this object exactly like the CPU `RawScorerImpl`.
use std::ffi::CStr;

use common::counter::hardware_counter::HardwareCounterCell;
use common::types::{PointOffsetType, ScoreType};

use super::ffi::*;
use crate::common::operation_error::{OperationError, OperationResult};
use crate::vector_storage::raw_scorer::RawScorer;
use crate::vector_storage::query_scorer::QueryScorerBytes;

pub(crate) fn last_error() -> String {
    unsafe { CStr::from_ptr(qb_last_error()).to_string_lossy().into_owned() }
}

/// Owns a `qb_storage`: one segment's vectors resident in HBM.
pub struct B200Storage { pub(crate) raw: *mut qb_storage }
unsafe impl Send for B200Storage {}
unsafe impl Sync for B200Storage {} // the library is thread-safe across handles (include/qb200.h, conventions)
impl Drop for B200Storage { fn drop(&mut self) { unsafe { qb_storage_destroy(self.raw) } } }

pub struct B200RawScorer<'a> {
    raw: *mut qb_scorer,
    hardware_counter: HardwareCounterCell,
    _storage: std::marker::PhantomData<&'a B200Storage>, // the `'a` borrow of `Box<dyn RawScorer + 'a>`
}

impl B200Storage {
    /// RawScorerBuilder::build_raw_scorer (raw_scorer.rs:122-128): errors surface here, never while scoring.
    pub fn build_raw_scorer<'a>(&'a self, query: &[f32], hc: HardwareCounterCell) -> OperationResult<Box<dyn RawScorer + 'a>> {
        let mut raw = std::ptr::null_mut();
        let st = unsafe { qb_scorer_create(self.raw, query.as_ptr(), &mut raw) };
        if st != QB_OK { return Err(OperationError::service_error(last_error())); }
        Ok(Box::new(B200RawScorer { raw, hardware_counter: hc, _storage: std::marker::PhantomData }))
    }
    /// QuantizedVectorsRead::raw_internal_scorer: `Err(InternalScorerUnsupported)` maps from QB_ERR_UNSUPPORTED (PQ).
    pub fn raw_internal_scorer<'a>(&'a self, point: PointOffsetType, hc: HardwareCounterCell)
        -> Result<Box<dyn RawScorer + 'a>, HardwareCounterCell> {
        let mut raw = std::ptr::null_mut();
        let st = unsafe { qb_scorer_create_internal(self.raw, point, &mut raw) };
        if st != QB_OK { return Err(hc); }
        Ok(Box::new(B200RawScorer { raw, hardware_counter: hc, _storage: std::marker::PhantomData }))
    }
}

impl B200RawScorer<'_> {
    pub(crate) fn raw(&self) -> *mut qb_scorer { self.raw }
    fn sync_counters(&self) {
        let mut c = qb_hw_counters::default();
        unsafe { qb_scorer_take_counters(self.raw, &mut c) };
        self.hardware_counter.cpu_counter().incr_delta(c.cpu as usize);
        self.hardware_counter.vector_io_read().incr_delta(c.vector_io_read as usize);
    }
}

impl RawScorer for B200RawScorer<'_> {
    fn score_points(&self, points: &[PointOffsetType], scores: &mut [ScoreType]) {
        assert_eq!(points.len(), scores.len()); // raw_scorer.rs:562
        let st = unsafe { qb_score_points(self.raw, points.as_ptr(), points.len(), scores.as_mut_ptr()) };
        assert!(st == QB_OK, "{}", last_error()); // same contract as `.expect("read vectors")`, metric_query_scorer.rs:91
        self.sync_counters();
    }
    fn score_point(&self, point: PointOffsetType) -> ScoreType {
        let mut s = 0.0;
        let st = unsafe { qb_score_point(self.raw, point, &mut s) };
        assert!(st == QB_OK, "{}", last_error());
        self.sync_counters();
        s
    }
    fn score_internal(&self, a: PointOffsetType, b: PointOffsetType) -> ScoreType {
        let mut s = 0.0;
        let st = unsafe { qb_score_internal(self.raw, a, b, &mut s) };
        assert!(st == QB_OK, "{}", last_error()); // "Panics if any id is out of range"
        self.sync_counters();
        s
    }
    fn scorer_bytes(&self) -> Option<&dyn QueryScorerBytes> { None }
}

impl Drop for B200RawScorer<'_> { fn drop(&mut self) { unsafe { qb_scorer_destroy(self.raw) } } }
