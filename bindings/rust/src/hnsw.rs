//! `GraphLayers::search` (lib/segment/src/index/hnsw_index/graph_layers.rs:530-561) for a batch of queries with the traversal on the GPU.
//! SOURCE ONLY (see ffi.rs).  The graph is handed over as the bytes of `links.bin` in GraphLinksFormat::Plain (graph_links/view.rs:121-135);
//! compressed graphs are converted once with `GraphLinks::to_edges` + `serialize_graph_links(.., GraphLinksFormatParam::Plain, ..)`.
use common::types::{PointOffsetType, ScoredPointOffset};

use super::ffi::*;
use super::raw_scorer::{last_error, B200Storage};
use crate::common::operation_error::{OperationError, OperationResult};

pub struct B200Hnsw<'a> { raw: *mut qb_hnsw, _storage: std::marker::PhantomData<&'a B200Storage> }
unsafe impl Send for B200Hnsw<'_> {}
unsafe impl Sync for B200Hnsw<'_> {}
impl Drop for B200Hnsw<'_> { fn drop(&mut self) { unsafe { qb_hnsw_destroy(self.raw) } } }

impl<'a> B200Hnsw<'a> {
    pub fn from_plain_links(storage: &'a B200Storage, links_bin: &[u8], m: usize, m0: usize) -> OperationResult<Self> {
        let mut raw = std::ptr::null_mut();
        let st = unsafe { qb_hnsw_create_plain(storage.raw, links_bin.as_ptr(), links_bin.len() as u64, m as u32, m0 as u32, &mut raw) };
        if st != QB_OK { return Err(OperationError::service_error(last_error())); }
        Ok(Self { raw, _storage: std::marker::PhantomData })
    }

    /// `entry` = GraphLayers::get_entry_point(filters, custom_entry_points) (it depends on the filter, so it stays host logic);
    /// `deleted` = the filter as a bitmap (bit = 1: check_vector fails), or None.
    pub fn search_batch(&self, queries: &[f32], n_queries: usize, top: usize, ef: usize, entry: (PointOffsetType, usize), deleted: Option<&[u64]>)
        -> Vec<Vec<ScoredPointOffset>> {
        let mut out = vec![qb_scored_point::default(); n_queries * top];
        let mut counts = vec![0u32; n_queries];
        let st = unsafe {
            qb_hnsw_search_batch(self.raw, queries.as_ptr(), n_queries as u32, top as u32, ef as u32, entry.0, entry.1 as u32,
                                 deleted.map_or(std::ptr::null(), |d| d.as_ptr()), std::ptr::null(), out.as_mut_ptr(), counts.as_mut_ptr(), std::ptr::null_mut())
        };
        assert!(st == QB_OK, "{}", last_error());
        (0..n_queries).map(|q| out[q * top..q * top + counts[q] as usize].iter().map(|p| ScoredPointOffset { idx: p.idx, score: p.score }).collect()).collect()
    }
}
