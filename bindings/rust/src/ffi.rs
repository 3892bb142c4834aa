//! Raw FFI declarations for libqdrant_b200.so (include/qb200.h).  SOURCE ONLY: there is no Rust toolchain in the
//! build image, so this file has never been compiled there; it is the binding a Qdrant maintainer would add under
//! `lib/segment/src/vector_storage/b200/ffi.rs`.  Every signature mirrors include/qb200.h one to one.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub type qb_status = i32;
pub const QB_OK: qb_status = 0;
pub const QB_ERR_UNSUPPORTED: qb_status = -3;
pub const QB_ERR_CANCELLED: qb_status = -5;

#[repr(C)]
pub struct qb_storage { _private: [u8; 0] }
#[repr(C)]
pub struct qb_scorer { _private: [u8; 0] }

/// Same layout as `common::types::ScoredPointOffset` (`#[repr(C)] { idx: u32, score: f32 }`).
#[repr(C)]
#[derive(Copy, Clone, Default)]
pub struct qb_scored_point { pub idx: u32, pub score: f32 }

#[repr(C)]
#[derive(Copy, Clone, Default)]
pub struct qb_hw_counters { pub cpu: u64, pub vector_io_read: u64 }

#[link(name = "qdrant_b200")]
extern "C" {
    pub fn qb_last_error() -> *const c_char;
    pub fn qb_device_count(out: *mut i32) -> qb_status;
    pub fn qb_storage_create_dense(device: i32, dt: i32, distance: i32, dim: u32, count: u64,
                                   host_rows: *const c_void, row_stride_bytes: u64, out: *mut *mut qb_storage) -> qb_status;
    pub fn qb_storage_write_rows(s: *mut qb_storage, first_row: u64, n_rows: u64, host_rows: *const c_void, row_stride_bytes: u64) -> qb_status;
    pub fn qb_storage_create_sq8(device: i32, dim: u32, count: u64, rows: *const u8, row_bytes: u32, alpha: f32, offset: f32,
                                 multiplier: f32, dt: i32, invert: i32, metric: i32, out: *mut *mut qb_storage) -> qb_status;
    pub fn qb_storage_create_pq(device: i32, dim: u32, m: u32, div_start_end: *const u32, centroids: *const f32, n_centroids: u32,
                                codes: *const u8, count: u64, dt: i32, invert: i32, metric: i32, out: *mut *mut qb_storage) -> qb_status;
    pub fn qb_storage_create_bq(device: i32, dim: u32, enc: i32, qenc: i32, rows: *const u8, row_bytes: u32, count: u64, dt: i32,
                                invert: i32, mean_std: *const f32, metric: i32, out: *mut *mut qb_storage) -> qb_status;
    pub fn qb_storage_destroy(s: *mut qb_storage);
    pub fn qb_storage_set_deleted(s: *mut qb_storage, bitmap_words: *const u64, n_words: u64) -> qb_status;
    pub fn qb_scorer_create(s: *mut qb_storage, query: *const f32, out: *mut *mut qb_scorer) -> qb_status;
    pub fn qb_scorer_create_internal(s: *mut qb_storage, point_id: u32, out: *mut *mut qb_scorer) -> qb_status;
    pub fn qb_scorer_destroy(sc: *mut qb_scorer);
    pub fn qb_score_points(sc: *mut qb_scorer, ids: *const u32, n: usize, scores: *mut f32) -> qb_status;
    pub fn qb_score_point(sc: *mut qb_scorer, id: u32, score: *mut f32) -> qb_status;
    pub fn qb_score_internal(sc: *mut qb_scorer, a: u32, b: u32, score: *mut f32) -> qb_status;
    pub fn qb_scorer_take_counters(sc: *mut qb_scorer, out: *mut qb_hw_counters) -> qb_status;
    pub fn qb_search_batch(s: *mut qb_storage, queries: *const f32, n_queries: u32, top: u32, deleted_bitmap: *const u64,
                           id_list: *const u32, n_ids: u64, is_stopped: *const i32, out: *mut qb_scored_point,
                           out_counts: *mut u32, counters: *mut qb_hw_counters) -> qb_status;
    /// QueryVector::{RecommendBestScore = 1, RecommendSumScores = 2, Discover = 3, Context = 4}: `vectors` holds the flattened
    /// example vectors (positives then negatives / target then pairs / pairs), see include/qb200.h.
    pub fn qb_scorer_create_custom(s: *mut qb_storage, kind: i32, vectors: *const f32, n_a: u32, n_b: u32, out: *mut *mut qb_scorer) -> qb_status;
    pub fn qb_search_custom(s: *mut qb_storage, kind: i32, vectors: *const f32, n_a: u32, n_b: u32, top: u32, deleted_bitmap: *const u64,
                            id_list: *const u32, n_ids: u64, is_stopped: *const i32, out: *mut qb_scored_point, out_count: *mut u32,
                            counters: *mut qb_hw_counters) -> qb_status;
    /// Multivector MaxSim (score_max_similarity): point p = rows [point_offsets[p], point_offsets[p+1]) of a token-level storage.
    pub fn qb_search_maxsim(s: *mut qb_storage, point_offsets: *const u32, n_points: u32, query_vectors: *const f32, n_query_vectors: u32, top: u32,
                            deleted_points: *const u64, out: *mut qb_scored_point, out_count: *mut u32, counters: *mut qb_hw_counters) -> qb_status;
    pub fn qb_score_maxsim(s: *mut qb_storage, point_offsets: *const u32, n_points: u32, query_vectors: *const f32, n_query_vectors: u32,
                           point_ids: *const u32, n: usize, scores: *mut f32) -> qb_status;
    /// Quantizer encode on rows already resident in HBM (device pointers); outputs are the reference's row formats.
    pub fn qb_sq8_find_alpha_offset_device(device: i32, dim: u32, count: u64, dev_rows: *const f32, row_stride_bytes: u64, alpha: *mut f32,
                                           offset: *mut f32) -> qb_status;
    pub fn qb_sq8_encode_rows_device(device: i32, dim: u32, count: u64, dev_rows: *const f32, row_stride_bytes: u64, alpha: f32, offset: f32, dt: i32,
                                     invert: i32, dev_out: *mut u8, stream: *mut c_void) -> qb_status;
    pub fn qb_bq_row_bytes(dim: u32, encoding: i32) -> u32;
    pub fn qb_bq_encode_rows_device(device: i32, dim: u32, count: u64, dev_rows: *const f32, row_stride_bytes: u64, encoding: i32, mean_std: *const f32,
                                    dev_out: *mut u8, stream: *mut c_void) -> qb_status;
    pub fn qb_pq_encode_rows_device(device: i32, dim: u32, chunk: u32, n_centroids: u32, centroids: *const f32, count: u64, dev_rows: *const f32,
                                    row_stride_bytes: u64, dev_codes: *mut u8, stream: *mut c_void) -> qb_status;
    pub fn qb_rescore(orig: *mut qb_scorer, ids: *const u32, n: usize, top: u32, out: *mut qb_scored_point, out_count: *mut u32) -> qb_status;
}
