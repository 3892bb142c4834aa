"""Host-side mirror of the reference's scorer interface, over the C ABI (include/qb200.h).

Same names, argument meaning and error behaviour as lib/segment:
    Distance                         lib/segment/src/types.rs:313-370
    RawScorer                        lib/segment/src/vector_storage/raw_scorer.rs:39-54
    RawScorerBuilder.build_raw_scorer                               raw_scorer.rs:122-128
    QuantizedVectorsRead.raw_scorer / raw_internal_scorer           quantized/quantized_vectors/read_access.rs:26-34
    FilteredScorer                   lib/segment/src/index/hnsw_index/point_scorer.rs:53-58,160-304
    BatchFilteredSearcher            point_scorer.rs:312-472
    get_oversampled_top / postprocess_search_result                 index/vector_index_search_common.rs:27-91
All scoring happens in libqdrant_b200.so on the GPU; this file holds no arithmetic (numpy is used for buffers).
"""
from __future__ import annotations

import ctypes as C
import enum
import sys
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _capi
from ._capi import HwCounters, QbError, ScoredPoint, check, f32p, i32p, lib, u8p, u32p, u64p, vp

SCORED_POINT_OFFSET = np.dtype([("idx", np.uint32), ("score", np.float32)])  # #[repr(C)] ScoredPointOffset
VECTOR_READ_BATCH_SIZE = 64  # lib/segment/src/vector_storage/common.rs:20


class Distance(enum.IntEnum):
    Cosine = 0
    Euclid = 1
    Dot = 2
    Manhattan = 3

    def postprocess_score(self, score: float) -> float:
        """Distance::postprocess_score (types.rs:349-358) — applied at shard level for Nearest queries."""
        return float(lib().qb_metric_postprocess(int(self), C.c_float(score)))


class VectorStorageDatatype(enum.IntEnum):
    Float32 = 0
    Float16 = 1
    Uint8 = 2


class DistanceType(enum.IntEnum):  # quantization::DistanceType
    Cosine = 0
    Dot = 1
    L1 = 2
    L2 = 3


class BQEncoding(enum.IntEnum):
    OneBit = 0
    TwoBits = 1
    OneAndHalfBits = 2


class BQQueryEncoding(enum.IntEnum):
    SameAsStorage = 0
    Scalar4bits = 1
    Scalar8bits = 2


def construct_vector_parameters(distance: Distance) -> tuple[DistanceType, bool]:
    """construct_vector_parameters (quantized_vectors.rs:205-234): Cosine -> Dot; invert = Euclid | Manhattan."""
    dt = {Distance.Cosine: DistanceType.Dot, Distance.Dot: DistanceType.Dot, Distance.Euclid: DistanceType.L2,
          Distance.Manhattan: DistanceType.L1}[distance]
    return dt, distance in (Distance.Euclid, Distance.Manhattan)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ids(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def _bitmap(point_deleted, count: int) -> Optional[np.ndarray]:
    """bool mask / BitSlice -> u64 words, bit i set = point i deleted."""
    if point_deleted is None:
        return None
    m = np.asarray(point_deleted)
    if m.dtype == np.uint64:
        if m.size < (count + 63) // 64:   # the library reads ceil(count / 64) words
            raise ValueError(f"deleted bitmap has {m.size} words, need {(count + 63) // 64}")
        return np.ascontiguousarray(m)
    bits = np.zeros(((count + 63) // 64) * 64, dtype=bool)
    bits[: m.size] = m.astype(bool)
    return np.packbits(bits, bitorder="little").view(np.uint64).copy()


def metric_preprocess(distance: Distance, vectors, device: int = 0) -> np.ndarray:
    """Metric::preprocess for a batch (cosine normalisation at insert time, types.rs:334-347)."""
    v = np.atleast_2d(_f32(vectors))
    out = np.empty_like(v)
    check(lib().qb_metric_preprocess(device, int(distance), v.shape[1], v.shape[0], v.ctypes.data_as(f32p), out.ctypes.data_as(f32p)))
    return out.reshape(np.shape(vectors))


class QueryKind(enum.IntEnum):  # QueryVector variants beyond Nearest (data_types/vectors.rs)
    RecommendBestScore = 1
    RecommendSumScores = 2
    Discover = 3
    Context = 4
    FeedbackNaive = 5


class RecoQuery:
    """vector_storage/query/reco_query.rs:12-29 — positives and negatives; wrap in RecoBestScoreQuery / RecoSumScoresQuery."""

    def __init__(self, positives, negatives):
        self.positives = [np.asarray(v, dtype=np.float32) for v in positives]
        self.negatives = [np.asarray(v, dtype=np.float32) for v in negatives]


class RecoBestScoreQuery:
    kind = QueryKind.RecommendBestScore

    def __init__(self, query: RecoQuery):
        self.query = query

    def flat(self):
        return self.query.positives + self.query.negatives, len(self.query.positives), len(self.query.negatives)


class RecoSumScoresQuery(RecoBestScoreQuery):
    kind = QueryKind.RecommendSumScores


class ContextPair:
    """vector_storage/query/context_query.rs:13-17"""

    def __init__(self, positive, negative):
        self.positive = np.asarray(positive, dtype=np.float32)
        self.negative = np.asarray(negative, dtype=np.float32)


class DiscoverQuery:
    """vector_storage/query/discover_query.rs:27-36 — a target and context pairs."""
    kind = QueryKind.Discover

    def __init__(self, target, pairs: Sequence[ContextPair]):
        self.target = np.asarray(target, dtype=np.float32)
        self.pairs = list(pairs)

    def flat(self):
        return [self.target] + [v for p in self.pairs for v in (p.positive, p.negative)], len(self.pairs), 0


class ContextQuery:
    """vector_storage/query/context_query.rs:87-99"""
    kind = QueryKind.Context

    def __init__(self, pairs: Sequence[ContextPair]):
        self.pairs = list(pairs)

    def flat(self):
        return [v for p in self.pairs for v in (p.positive, p.negative)], len(self.pairs), 0


class FeedbackQuery:
    """vector_storage/query/feedback_query.rs:150-226 — the scoring form of NaiveFeedbackQuery: a target, context pairs with their
    partial_computation (confidence^b * c, derived from the feedback scores by the host exactly as FeedbackQuery::new does) and the
    coefficient `a`."""
    kind = QueryKind.FeedbackNaive

    def __init__(self, target, pairs: Sequence[ContextPair], partial_computations, a: float):
        self.target = np.asarray(target, dtype=np.float32)
        self.pairs = list(pairs)
        self.partial = np.ascontiguousarray(partial_computations, dtype=np.float32).reshape(-1)
        self.a = float(a)
        assert self.partial.size == len(self.pairs)

    def flat(self):
        return [self.target] + [v for p in self.pairs for v in (p.positive, p.negative)], len(self.pairs), 0


class RawScorer:
    """Box<dyn RawScorer>.  Scoring calls are infallible in the reference; here a CUDA failure raises QbError."""

    def __init__(self, storage: "_Storage", handle: int):
        self._storage = storage  # keeps the borrow alive ('a)
        self._h = vp(handle)

    def score_points(self, points: Sequence[int], scores: Optional[np.ndarray] = None) -> np.ndarray:
        ids = _ids(points)
        if scores is None:
            scores = np.empty(ids.size, dtype=np.float32)
        assert scores.size == ids.size  # raw_scorer.rs:562
        check(lib().qb_score_points(self._h, ids.ctypes.data_as(u32p), ids.size, scores.ctypes.data_as(f32p)))
        return scores

    def score_point(self, point: int) -> float:
        s = C.c_float()
        check(lib().qb_score_point(self._h, int(point), C.byref(s)))
        return np.float32(s.value)

    def score_internal(self, point_a: int, point_b: int) -> float:
        s = C.c_float()
        st = lib().qb_score_internal(self._h, int(point_a), int(point_b), C.byref(s))
        if st == _capi.QB_ERR_INVALID:
            raise IndexError(lib().qb_last_error().decode())  # "Panics if any id is out of range"
        check(st)
        return np.float32(s.value)

    def scorer_bytes(self):
        return None  # a GPU scorer has no QueryScorerBytes view

    def take_hardware_counters(self) -> tuple[int, int]:
        hc = HwCounters()
        check(lib().qb_scorer_take_counters(self._h, C.byref(hc)))
        return int(hc.cpu), int(hc.vector_io_read)

    def close(self):
        if self._h:
            lib().qb_scorer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if sys is None or sys.is_finalizing():  # the CUDA runtime / library may already be torn down at interpreter exit
                return
            self.close()
        except Exception:
            pass


class _Storage:
    def __init__(self):
        self._h = vp()
        self.count = 0
        self.dim = 0
        self.device = 0

    def _raw_scorer(self, query) -> RawScorer:
        q = _f32(query)
        if q.size != self.dim:
            raise ValueError(f"query has dim {q.size}, storage has {self.dim}")  # OperationError at construction
        h = vp()
        check(lib().qb_scorer_create(self._h, q.ctypes.data_as(f32p), C.byref(h)))
        return RawScorer(self, h.value)

    def _flat_custom(self, query):
        vecs, n_a, n_b = query.flat()
        if not vecs:
            raise ValueError("custom query without example vectors")
        m = np.ascontiguousarray(np.stack([_f32(v).reshape(-1) for v in vecs]))
        if m.shape[1] != self.dim:
            raise ValueError(f"query vectors have dim {m.shape[1]}, storage has {self.dim}")
        return m, n_a, n_b

    def raw_scorer_custom(self, query) -> RawScorer:
        """new_raw_scorer for QueryVector::{RecommendBestScore, RecommendSumScores, Discover, Context} (raw_scorer.rs:228-333)."""
        m, n_a, n_b = self._flat_custom(query)
        h = vp()
        if query.kind == QueryKind.FeedbackNaive:
            check(lib().qb_scorer_create_feedback(self._h, m.ctypes.data_as(f32p), n_a, C.c_float(query.a),
                                                  query.partial.ctypes.data_as(f32p) if n_a else None, C.byref(h)))
        else:
            check(lib().qb_scorer_create_custom(self._h, int(query.kind), m.ctypes.data_as(f32p), n_a, n_b, C.byref(h)))
        return RawScorer(self, h.value)

    def search_custom(self, query, top: int, point_deleted=None, id_list=None, counters: Optional[HwCounters] = None):
        """peek_top_iter driven by a custom-query scorer -> SCORED_POINT_OFFSET array, descending."""
        m, n_a, n_b = self._flat_custom(query)
        out = np.zeros(max(top, 1), dtype=SCORED_POINT_OFFSET)
        count = C.c_uint32()
        bm = _bitmap(point_deleted, self.count)
        ids = None if id_list is None else _ids(id_list)
        tail = (None if bm is None else bm.ctypes.data_as(u64p), None if ids is None else ids.ctypes.data_as(u32p), 0 if ids is None else ids.size, None,
                out.ctypes.data_as(C.POINTER(ScoredPoint)), C.byref(count), None if counters is None else C.byref(counters))
        if query.kind == QueryKind.FeedbackNaive:
            check(lib().qb_search_feedback(self._h, m.ctypes.data_as(f32p), n_a, C.c_float(query.a), query.partial.ctypes.data_as(f32p) if n_a else None,
                                           int(top), *tail))
        else:
            check(lib().qb_search_custom(self._h, int(query.kind), m.ctypes.data_as(f32p), n_a, n_b, int(top), *tail))
        return out[: count.value].copy()

    def _raw_internal_scorer(self, point_id: int) -> RawScorer:
        h = vp()
        check(lib().qb_scorer_create_internal(self._h, int(point_id), C.byref(h)))
        return RawScorer(self, h.value)

    def set_deleted(self, point_deleted) -> None:
        bm = _bitmap(point_deleted, self.count)
        if bm is None:
            check(lib().qb_storage_set_deleted(self._h, None, 0))
        else:
            check(lib().qb_storage_set_deleted(self._h, bm.ctypes.data_as(u64p), bm.size))

    def hbm_bytes(self) -> int:
        b = C.c_uint64()
        check(lib().qb_storage_info(self._h, None, None, C.byref(b)))
        return int(b.value)

    def stream_ptr(self) -> int:
        return int(lib().qb_storage_stream(self._h) or 0)

    def set_on_disk(self, on_disk: bool) -> None:
        """VectorStorage::is_on_disk of the storage this copy caches (decides vector_io_read metering)."""
        check(lib().qb_storage_set_on_disk(self._h, 1 if on_disk else 0))

    def search_stats(self, reset: bool = False) -> tuple[int, int]:
        """(fused searches, reruns after a broken fast-path assumption)"""
        a, b = C.c_uint64(), C.c_uint64()
        check(lib().qb_search_stats(self._h, C.byref(a), C.byref(b), 1 if reset else 0))
        return int(a.value), int(b.value)

    def profile(self, on: bool) -> None:
        check(lib().qb_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, reset: bool = True) -> tuple[int, float]:
        n, ms = C.c_uint64(), C.c_double()
        check(lib().qb_profile_read(self._h, C.byref(n), C.byref(ms), 1 if reset else 0))
        return int(n.value), float(ms.value)

    def search_batch(self, queries, top: int, point_deleted=None, id_list=None, is_stopped=None, counters: Optional[HwCounters] = None):
        """Fused BatchFilteredSearcher scan -> list (one per query) of SCORED_POINT_OFFSET arrays, descending."""
        q = np.atleast_2d(_f32(queries))
        if q.shape[1] != self.dim:
            raise ValueError(f"queries have dim {q.shape[1]}, storage has {self.dim}")
        nq = q.shape[0]
        out = np.zeros((nq, max(top, 1)), dtype=SCORED_POINT_OFFSET)
        counts = np.zeros(nq, dtype=np.uint32)
        bm = _bitmap(point_deleted, self.count)
        ids = None if id_list is None else _ids(id_list)
        stop = None
        if is_stopped is not None:
            stop = is_stopped if isinstance(is_stopped, C.c_int32) else C.c_int32(int(bool(is_stopped)))
        check(lib().qb_search_batch(
            self._h, q.ctypes.data_as(f32p), nq, int(top),
            None if bm is None else bm.ctypes.data_as(u64p),
            None if ids is None else ids.ctypes.data_as(u32p), 0 if ids is None else ids.size,
            None if stop is None else C.byref(stop),
            out.ctypes.data_as(C.POINTER(ScoredPoint)), counts.ctypes.data_as(u32p),
            None if counters is None else C.byref(counters)))
        return [out[i, : counts[i]].copy() for i in range(nq)]

    def close(self):
        if self._h:
            lib().qb_storage_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            if sys is None or sys.is_finalizing():     # interpreter shutdown: module globals may already be gone
                return
            self.close()
        except Exception:
            pass


class DenseVectorStorage(_Storage):
    """A segment's dense vectors resident in HBM (the GPU copy is a cache of VectorStorageEnum's rows)."""

    def __init__(self, vectors, distance: Distance, datatype: VectorStorageDatatype = VectorStorageDatatype.Float32, device: int = 0,
                 count: Optional[int] = None, dim: Optional[int] = None):
        super().__init__()
        self.distance, self.datatype, self.device = Distance(distance), VectorStorageDatatype(datatype), device
        np_dt = {VectorStorageDatatype.Float32: np.float32, VectorStorageDatatype.Float16: np.float16, VectorStorageDatatype.Uint8: np.uint8}[self.datatype]
        if vectors is None:
            self.count, self.dim = int(count), int(dim)
            ptr, stride = None, 0
        else:
            v = np.ascontiguousarray(vectors, dtype=np_dt)
            assert v.ndim == 2
            self.count, self.dim = v.shape
            ptr, stride = v.ctypes.data_as(vp), v.strides[0]
        check(lib().qb_storage_create_dense(device, int(self.datatype), int(self.distance), self.dim, self.count, ptr, stride, C.byref(self._h)))

    # RawScorerBuilder
    def build_raw_scorer(self, query) -> RawScorer:
        return self._raw_scorer(query)

    def raw_internal_scorer(self, point_id: int) -> RawScorer:
        return self._raw_internal_scorer(point_id)

    def write_rows_device(self, first_row: int, n_rows: int, dev_ptr: int, row_stride_bytes: int = 0) -> None:
        check(lib().qb_storage_write_rows_device(self._h, first_row, n_rows, vp(dev_ptr), row_stride_bytes))

    def write_rows(self, first_row: int, rows) -> None:
        r = np.ascontiguousarray(rows)
        check(lib().qb_storage_write_rows(self._h, first_row, r.shape[0], r.ctypes.data_as(vp), r.strides[0]))

    def get_dense(self, ids) -> np.ndarray:
        ids = _ids(ids)
        np_dt = {VectorStorageDatatype.Float32: np.float32, VectorStorageDatatype.Float16: np.float16, VectorStorageDatatype.Uint8: np.uint8}[self.datatype]
        out = np.empty((ids.size, self.dim), dtype=np_dt)
        check(lib().qb_storage_read_rows(self._h, ids.ctypes.data_as(u32p), ids.size, out.ctypes.data_as(vp)))
        return out


class ScalarQuantizedVectors(_Storage):
    """QuantizedVectors backed by EncodedVectorsU8: `rows` is quantized.data ([f32 v_off][actual_dim u8] per row)."""

    def __init__(self, rows, dim: int, alpha: float, offset: float, multiplier: float, distance: Distance, device: int = 0,
                 rows_ptr: Optional[int] = None, count: Optional[int] = None):
        """`rows`: numpy [count, 4 + actual_dim] u8, or None with `rows_ptr` = address (host or device) of such rows."""
        super().__init__()
        self.dim, self.device, self.distance = int(dim), device, Distance(distance)
        actual_dim = self.dim + (16 - self.dim % 16) % 16
        if rows is not None:
            r = np.ascontiguousarray(rows, dtype=np.uint8)
            self.count, ptr, row_bytes = r.shape[0], r.ctypes.data_as(u8p), (r.shape[1] if r.ndim == 2 else 0)
        else:
            self.count, ptr, row_bytes = int(count), C.cast(vp(int(rows_ptr)), u8p), 4 + actual_dim
        dt, invert = construct_vector_parameters(self.distance)
        check(lib().qb_storage_create_sq8(device, self.dim, self.count, ptr, row_bytes, C.c_float(alpha),
                                          C.c_float(offset), C.c_float(multiplier), int(dt), int(invert), int(self.distance), C.byref(self._h)))

    def raw_scorer(self, query) -> RawScorer:
        return self._raw_scorer(query)

    def raw_internal_scorer(self, point_id: int) -> RawScorer:
        return self._raw_internal_scorer(point_id)


class ProductQuantizedVectors(_Storage):
    def __init__(self, codes, centroids, chunk: int, dim: int, distance: Distance, device: int = 0):
        super().__init__()
        c = np.ascontiguousarray(codes, dtype=np.uint8)
        cent = _f32(centroids)
        self.count, self.dim, self.device, self.distance = c.shape[0], int(dim), device, Distance(distance)
        starts = np.arange(0, dim, chunk, dtype=np.uint32)  # get_vector_division, encoded_vectors_pq.rs:164-169
        div = np.stack([starts, np.minimum(starts + chunk, dim).astype(np.uint32)], axis=1).astype(np.uint32).copy()
        self.m = div.shape[0]
        assert c.shape[1] == self.m
        dt, invert = construct_vector_parameters(self.distance)
        check(lib().qb_storage_create_pq(device, self.dim, self.m, div.ctypes.data_as(u32p), cent.ctypes.data_as(f32p), cent.shape[0],
                                         c.ctypes.data_as(u8p), self.count, int(dt), int(invert), int(self.distance), C.byref(self._h)))

    def raw_scorer(self, query) -> RawScorer:
        return self._raw_scorer(query)

    def raw_internal_scorer(self, point_id: int) -> RawScorer:
        """InternalScorerUnsupported for PQ (quantized_query_scorer.rs:60-66): raises QbError(QB_ERR_UNSUPPORTED)."""
        return self._raw_internal_scorer(point_id)


class BinaryQuantizedVectors(_Storage):
    def __init__(self, rows, dim: int, distance: Distance, encoding: BQEncoding = BQEncoding.OneBit,
                 query_encoding: BQQueryEncoding = BQQueryEncoding.SameAsStorage, mean_std=None, device: int = 0):
        super().__init__()
        r = np.ascontiguousarray(rows, dtype=np.uint8)
        self.count, self.dim, self.device, self.distance = r.shape[0], int(dim), device, Distance(distance)
        dt, invert = construct_vector_parameters(self.distance)
        ms = None if mean_std is None else _f32(mean_std)
        check(lib().qb_storage_create_bq(device, self.dim, int(encoding), int(query_encoding), r.ctypes.data_as(u8p), r.shape[1], self.count, int(dt),
                                         int(invert), None if ms is None else ms.ctypes.data_as(f32p), int(self.distance), C.byref(self._h)))

    def raw_scorer(self, query) -> RawScorer:
        return self._raw_scorer(query)

    def raw_internal_scorer(self, point_id: int) -> RawScorer:
        return self._raw_internal_scorer(point_id)


class FilteredScorer:
    """point_scorer.rs:53-58: a RawScorer plus the deleted-points filter; score_points clobbers and truncates `ids`."""

    def __init__(self, raw_scorer: RawScorer, point_deleted=None, filter_fn=None):
        self.raw_scorer = raw_scorer
        self.point_deleted = None if point_deleted is None else np.asarray(point_deleted, dtype=bool)
        self.filter_fn = filter_fn

    @classmethod
    def new(cls, query, vectors: _Storage, quantized_vectors: Optional[_Storage] = None, point_deleted=None, filter_fn=None):
        # point_scorer.rs:172-175
        rs = quantized_vectors.raw_scorer(query) if quantized_vectors is not None else vectors.build_raw_scorer(query)
        return cls(rs, point_deleted, filter_fn)

    @classmethod
    def new_internal(cls, point_id: int, vectors: _Storage, quantized_vectors: Optional[_Storage] = None, point_deleted=None):
        # point_scorer.rs:183-218: fall back to the original vectors when the quantizer cannot build an internal scorer (PQ)
        rs = None
        if quantized_vectors is not None:
            try:
                rs = quantized_vectors.raw_internal_scorer(point_id)
            except QbError as e:
                if e.status != _capi.QB_ERR_UNSUPPORTED:
                    raise
        if rs is None:
            rs = vectors.raw_internal_scorer(point_id)
        return cls(rs, point_deleted)

    def check_vector(self, point_id: int) -> bool:
        if self.point_deleted is not None and point_id < self.point_deleted.size and self.point_deleted[point_id]:
            return False
        return self.filter_fn is None or bool(self.filter_fn(point_id))

    def score_points(self, point_ids: list, limit: int = 0) -> np.ndarray:
        """point_scorer.rs:265-295 -> array of ScoredPointOffset for the ids that passed the filters."""
        kept = [p for p in point_ids if self.check_vector(p)]
        if limit:
            kept = kept[:limit]
        point_ids[:] = kept
        out = np.zeros(len(kept), dtype=SCORED_POINT_OFFSET)
        if kept:
            out["idx"] = kept
            out["score"] = self.raw_scorer.score_points(kept)
        return out

    def score_point(self, point_id: int) -> float:
        return self.raw_scorer.score_point(point_id)

    def score_internal(self, a: int, b: int) -> float:
        return self.raw_scorer.score_internal(a, b)


class BatchFilteredSearcher:
    """point_scorer.rs:312-472.  The reference keeps one RawScorer + one heap per query and walks 64-id chunks;
    here the whole scan + top-k is ONE fused call into the library (qb_search_batch)."""

    def __init__(self, queries, storage: _Storage, top: int, point_deleted=None):
        self.queries = np.atleast_2d(_f32(queries))
        self.storage = storage
        self.top = int(top)
        self.point_deleted = point_deleted
        if self.top == 0:
            raise ValueError("length must be greater than zero")  # FixedLengthPriorityQueue::new expect()

    @classmethod
    def new(cls, queries, vectors: _Storage, quantized_vectors: Optional[_Storage], top: int, point_deleted=None):
        return cls(queries, quantized_vectors if quantized_vectors is not None else vectors, top, point_deleted)

    def peek_top_all(self, is_stopped=None):
        return self.storage.search_batch(self.queries, self.top, point_deleted=self.point_deleted, is_stopped=is_stopped)

    def peek_top_iter(self, points: Iterable[int], is_stopped=None):
        ids = _ids(list(points))
        return self.storage.search_batch(self.queries, self.top, point_deleted=self.point_deleted, id_list=ids, is_stopped=is_stopped)


def rescore(original_scorer: RawScorer, ids, top: int) -> np.ndarray:
    """The rescoring step of postprocess_search_result (vector_index_search_common.rs:74-91) for a scorer that already exists:
    score `ids` with the original-vector scorer, sort descending, truncate to `top`."""
    ids = _ids(ids)
    out = np.zeros(max(int(top), 1), dtype=SCORED_POINT_OFFSET)
    n = C.c_uint32()
    check(lib().qb_rescore(original_scorer._h, ids.ctypes.data_as(u32p), ids.size, int(top), out.ctypes.data_as(C.POINTER(ScoredPoint)), C.byref(n)))
    return out[: n.value].copy()


def get_oversampled_top(top: int, quantized: bool, oversampling: Optional[float]) -> int:
    """vector_index_search_common.rs:27-45."""
    if quantized and oversampling is not None and oversampling > 1.0:
        return int(oversampling * top)
    return top


def postprocess_search_result(search_result: np.ndarray, original: DenseVectorStorage, query, top: int, rescore: bool) -> np.ndarray:
    """vector_index_search_common.rs:48-91: optionally rescore the candidates with the original vectors, sort desc, truncate."""
    if not rescore:
        return search_result[:top].copy()
    sc = original.build_raw_scorer(query)
    ids = _ids(search_result["idx"])
    out = np.zeros(max(top, 1), dtype=SCORED_POINT_OFFSET)
    n = C.c_uint32()
    check(lib().qb_rescore(sc._h, ids.ctypes.data_as(u32p), ids.size, int(top), out.ctypes.data_as(C.POINTER(ScoredPoint)), C.byref(n)))
    sc.close()
    return out[: n.value].copy()


# ------------------------------------------------------------------------------------------------ multivectors
class MultiVectorView:
    """A multivector collection over a token-level storage: point p = rows [offsets[p], offsets[p+1]) (the flattened layout of
    vector_storage/multi_dense).  Scores are ColBERT MaxSim (score_max_similarity, query_scorer/mod.rs:77-98)."""

    def __init__(self, storage: _Storage, offsets):
        self.storage = storage
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        assert self.offsets.ndim == 1 and self.offsets.size >= 1
        self.n_points = self.offsets.size - 1

    def _query(self, query_vectors) -> np.ndarray:
        q = np.atleast_2d(_f32(query_vectors))
        if q.shape[1] != self.storage.dim:
            raise ValueError(f"query vectors have dim {q.shape[1]}, storage has {self.storage.dim}")
        return np.ascontiguousarray(q)

    def search(self, query_vectors, top: int, point_deleted=None, counters: Optional[HwCounters] = None) -> np.ndarray:
        q = self._query(query_vectors)
        out = np.zeros(max(top, 1), dtype=SCORED_POINT_OFFSET)
        count = C.c_uint32()
        bm = _bitmap(point_deleted, self.n_points)
        check(lib().qb_search_maxsim(self.storage._h, self.offsets.ctypes.data_as(u32p), self.n_points, q.ctypes.data_as(f32p), q.shape[0], int(top),
                                     None if bm is None else bm.ctypes.data_as(u64p), out.ctypes.data_as(C.POINTER(ScoredPoint)), C.byref(count),
                                     None if counters is None else C.byref(counters)))
        return out[: count.value].copy()

    def score_points(self, query_vectors, points: Sequence[int]) -> np.ndarray:
        q = self._query(query_vectors)
        ids = _ids(points)
        scores = np.empty(ids.size, dtype=np.float32)
        check(lib().qb_score_maxsim(self.storage._h, self.offsets.ctypes.data_as(u32p), self.n_points, q.ctypes.data_as(f32p), q.shape[0],
                                    ids.ctypes.data_as(u32p), ids.size, scores.ctypes.data_as(f32p)))
        return scores

    # ---- custom queries whose examples are multivectors (MultiCustomQueryScorer, multi_custom_query_scorer.rs:88-104).  `query` is one of the
    # query classes above with 2-D arrays (vectors x dim) in place of vectors.
    def _flat_multi(self, query):
        examples, n_a, n_b = query.flat()
        mats = [self._query(e) for e in examples]
        off = np.concatenate([[0], np.cumsum([m.shape[0] for m in mats])]).astype(np.uint32)
        coef = None
        if query.kind == QueryKind.FeedbackNaive:
            coef = np.concatenate([[np.float32(query.a)], query.partial]).astype(np.float32)
        return np.ascontiguousarray(np.concatenate(mats)), off, n_a, n_b, coef

    def search_custom(self, query, top: int, point_deleted=None) -> np.ndarray:
        vecs, off, n_a, n_b, coef = self._flat_multi(query)
        out = np.zeros(max(top, 1), dtype=SCORED_POINT_OFFSET)
        count = C.c_uint32()
        bm = _bitmap(point_deleted, self.n_points)
        check(lib().qb_search_maxsim_custom(self.storage._h, self.offsets.ctypes.data_as(u32p), self.n_points, int(query.kind), vecs.ctypes.data_as(f32p),
                                            off.ctypes.data_as(u32p), n_a, n_b, None if coef is None else coef.ctypes.data_as(f32p), int(top),
                                            None if bm is None else bm.ctypes.data_as(u64p), out.ctypes.data_as(C.POINTER(ScoredPoint)), C.byref(count), None))
        return out[: count.value].copy()

    def score_points_custom(self, query, points: Sequence[int]) -> np.ndarray:
        vecs, off, n_a, n_b, coef = self._flat_multi(query)
        ids = _ids(points)
        scores = np.empty(ids.size, dtype=np.float32)
        check(lib().qb_score_maxsim_custom(self.storage._h, self.offsets.ctypes.data_as(u32p), self.n_points, int(query.kind), vecs.ctypes.data_as(f32p),
                                           off.ctypes.data_as(u32p), n_a, n_b, None if coef is None else coef.ctypes.data_as(f32p), ids.ctypes.data_as(u32p), ids.size,
                                           scores.ctypes.data_as(f32p)))
        return scores


# ------------------------------------------------------------------------------------------------ quantizer encode on the device
# Thin wrappers over the C ABI (include/qb200.h, "quantizer encode on the device"): every pointer is a raw device address.
def sq8_multiplier(alpha: float, distance: Distance) -> np.float32:
    """EncodedVectorsU8::encode, encoded_vectors_u8.rs:210-225: Dot a^2, L1 a, L2 -2a^2; negated when `invert`."""
    dt, inv = construct_vector_parameters(distance)
    a = np.float32(alpha)
    m = {DistanceType.Dot: a * a, DistanceType.L1: a, DistanceType.L2: np.float32(-2.0) * a * a}[dt]
    return np.float32(-m if inv else m)


def sq8_find_alpha_offset(rows_ptr: int, count: int, dim: int, row_stride_bytes: int = 0, device: int = 0) -> tuple[np.float32, np.float32]:
    a, o = C.c_float(), C.c_float()
    check(lib().qb_sq8_find_alpha_offset_device(device, dim, count, vp(rows_ptr), row_stride_bytes, C.byref(a), C.byref(o)))
    return np.float32(a.value), np.float32(o.value)


def sq8_encode_rows(rows_ptr: int, count: int, dim: int, alpha: float, offset: float, distance: Distance, out_ptr: int, row_stride_bytes: int = 0,
                    device: int = 0, stream: int = 0) -> None:
    """out_ptr: count x (4 + actual_dim) bytes, the row format qb_storage_create_sq8 / ScalarQuantizedVectors(rows_ptr=...) takes."""
    dt, inv = construct_vector_parameters(distance)
    check(lib().qb_sq8_encode_rows_device(device, dim, count, vp(rows_ptr), row_stride_bytes, np.float32(alpha), np.float32(offset), int(dt), int(inv), vp(out_ptr),
                                          vp(stream)))


def bq_row_bytes(dim: int, encoding: BQEncoding) -> int:
    return int(lib().qb_bq_row_bytes(dim, int(encoding)))


def bq_encode_rows(rows_ptr: int, count: int, dim: int, encoding: BQEncoding, out_ptr: int, mean_std=None, row_stride_bytes: int = 0, device: int = 0,
                   stream: int = 0) -> None:
    ms = None if mean_std is None else _f32(mean_std)
    check(lib().qb_bq_encode_rows_device(device, dim, count, vp(rows_ptr), row_stride_bytes, int(encoding), None if ms is None else ms.ctypes.data_as(f32p),
                                         vp(out_ptr), vp(stream)))


def pq_encode_rows(rows_ptr: int, count: int, dim: int, chunk: int, centroids, out_ptr: int, row_stride_bytes: int = 0, device: int = 0, stream: int = 0) -> None:
    c = _f32(centroids)
    assert c.ndim == 2 and c.shape[1] == dim
    check(lib().qb_pq_encode_rows_device(device, dim, chunk, c.shape[0], c.ctypes.data_as(f32p), count, vp(rows_ptr), row_stride_bytes, vp(out_ptr), vp(stream)))


class _LoadedStorage(_Storage):
    def __init__(self, handle, device: int, distance: Distance):
        super().__init__()
        self._h, self.device, self.distance = handle, device, Distance(distance)
        d, n = C.c_uint32(), C.c_uint64()
        check(lib().qb_storage_info(self._h, C.byref(d), C.byref(n), None))
        self.dim, self.count = int(d.value), int(n.value)

    def raw_scorer(self, query) -> RawScorer:
        return self._raw_scorer(query)

    def raw_internal_scorer(self, point_id: int) -> RawScorer:
        return self._raw_internal_scorer(point_id)


def load_dense_file(file_bytes, distance: Distance, dim: int, datatype: VectorStorageDatatype = VectorStorageDatatype.Float32, device: int = 0) -> _Storage:
    """A segment's `matrix.dat` (b"data" header + rows, dense/immutable_dense_vectors.rs:100-113) as it lies on disk."""
    blob = np.frombuffer(bytes(file_bytes), dtype=np.uint8) if not isinstance(file_bytes, np.ndarray) else np.ascontiguousarray(file_bytes, dtype=np.uint8)
    h = vp()
    check(lib().qb_storage_load_dense_file(device, int(datatype), int(distance), int(dim), blob.ctypes.data_as(u8p), blob.size, C.byref(h)))
    return _LoadedStorage(h, device, distance)


def load_quantized(meta_json, data, metric: Distance, count: int = 0, device: int = 0) -> _Storage:
    """`quantized.meta.json` + `quantized.data` of a segment, unchanged (quantized/quantized_storage.rs:63-69)."""
    meta = meta_json.encode() if isinstance(meta_json, str) else bytes(meta_json)
    blob = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    h = vp()
    check(lib().qb_storage_load_quantized(device, int(metric), meta, len(meta), blob.ctypes.data_as(u8p), blob.size, int(count), C.byref(h)))
    return _LoadedStorage(h, device, metric)


def set_option(name: str, value: int) -> None:
    """Debugging / experiment switches of the library (qb_set_option)."""
    check(lib().qb_set_option(name.encode(), int(value)))


class HnswGraph:
    """GraphLayers::search with the traversal on the device (qb_hnsw_*): a graph in the reference's plain links.bin layout
    bound to a storage (dense f32 or SQ8); search() answers a batch of queries in one call."""

    def __init__(self, storage: _Storage, links_bin, m: int, m0: int):
        self._storage = storage
        self._h = vp()
        blob = np.ascontiguousarray(links_bin, dtype=np.uint8)
        h = vp()
        check(lib().qb_hnsw_create_plain(storage._h, blob.ctypes.data_as(u8p), blob.size, int(m), int(m0), C.byref(h)))
        self._h = h

    def search(self, queries, top: int, ef: int, entry_point: int, entry_level: int, point_deleted=None, counters: Optional[HwCounters] = None):
        q = np.atleast_2d(_f32(queries))
        if q.shape[1] != self._storage.dim:
            raise ValueError(f"queries have dim {q.shape[1]}, storage has {self._storage.dim}")
        nq = q.shape[0]
        out = np.zeros((nq, max(top, 1)), dtype=SCORED_POINT_OFFSET)
        counts = np.zeros(nq, dtype=np.uint32)
        bm = _bitmap(point_deleted, self._storage.count)
        check(lib().qb_hnsw_search_batch(self._h, q.ctypes.data_as(f32p), nq, int(top), int(ef), int(entry_point), int(entry_level),
                                         None if bm is None else bm.ctypes.data_as(u64p), None, out.ctypes.data_as(C.POINTER(ScoredPoint)),
                                         counts.ctypes.data_as(u32p), None if counters is None else C.byref(counters)))
        return [out[i, : counts[i]].copy() for i in range(nq)]

    def stats(self, reset: bool = True) -> tuple[int, int]:
        a, b = C.c_uint64(), C.c_uint64()
        check(lib().qb_hnsw_stats(self._h, C.byref(a), C.byref(b), 1 if reset else 0))
        return int(a.value), int(b.value)

    def close(self):
        if self._h:
            lib().qb_hnsw_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            if sys is None or sys.is_finalizing():     # interpreter shutdown: module globals may already be gone
                return
            self.close()
        except Exception:
            pass
