"""Edge cases of the boundary the reference defines but round 1 left untested: OrderedFloat's NaN order, the bit pattern of
-0.0, tops above 4096 (oversampling), sharded ids through every id-taking entry point, device-pointer uploads, bitmap length,
vector_io_read metering for on-disk storages."""
import numpy as np
import pytest

from tests.util import assert_topk_equal, pack_bitmap

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qb():
    from qdrant_b200 import scorer

    return scorer


def test_nan_scores_rank_highest_like_ordered_float(qb, oracle):
    """ScoredPointOffset orders by OrderedFloat(score): NaN is the greatest value (lib/common/common/src/types.rs:21-25)."""
    rng = np.random.default_rng(2)
    for n in (3_000, 200_000):                       # direct path and the single-query in-kernel top-k path
        base = rng.standard_normal((n, 64)).astype(np.float32)
        base[[17, n // 2, n - 5], 3] = np.nan
        q = rng.standard_normal(64).astype(np.float32)
        st = qb.DenseVectorStorage(base, qb.Distance.Dot)
        got = st.search_batch(q, 10)[0]
        want = oracle.scan_f32(oracle.DOT, base, q[None], 10)[0]
        assert np.isnan(got["score"][:3]).all() and sorted(got["idx"][:3].tolist()) == [17, n // 2, n - 5]
        assert_topk_equal(got, want, what=f"NaN n={n}")
        # two queries -> threshold + filter path: NaN must pass the filter
        got2 = st.search_batch(np.stack([q, -q]), 10)
        want2 = oracle.scan_f32(oracle.DOT, base, np.stack([q, -q]), 10)
        for a, b in zip(got2, want2):
            assert_topk_equal(a, b, what=f"NaN filter path n={n}")
        st.close()


def test_negative_zero_keeps_its_bits(qb, oracle):
    base = np.zeros((5_000, 32), dtype=np.float32)
    base[:, 0] = -1.0
    base[100:110, 0] = 0.0                           # dot with q = (-0.0 ...) gives exact zeros of both signs
    q = np.zeros(32, dtype=np.float32); q[0] = 1.0
    base[100:105, 0] = -0.0                          # 1 * -0.0 + 0 ... = -0.0 only through the unfused tail; check bits against the oracle
    st = qb.DenseVectorStorage(base, qb.Distance.Dot)
    sc = st.build_raw_scorer(q)
    ids = np.arange(95, 115, dtype=np.uint32)
    np.testing.assert_array_equal(sc.score_points(ids).view(np.uint32), oracle.score_points_f32(oracle.DOT, base, q, ids).view(np.uint32))
    got = st.search_batch(q, 10)[0]
    want = oracle.scan_f32(oracle.DOT, base, q[None], 10)[0]
    np.testing.assert_array_equal(np.sort(got["score"].view(np.uint32)), np.sort(want["score"].view(np.uint32)))
    # Euclid of identical vectors is -0.0 in the reference (negated +0.0): the sign must survive the top-k keys
    e = qb.DenseVectorStorage(base, qb.Distance.Euclid)
    g = e.search_batch(base[100], 3)[0]
    w = oracle.scan_f32(oracle.EUCLID, base, base[100][None], 3)[0]
    np.testing.assert_array_equal(np.sort(g["score"].view(np.uint32)), np.sort(w["score"].view(np.uint32)))
    assert (g["score"].view(np.uint32) == 0x80000000).any()
    sc.close(); st.close(); e.close()


@pytest.mark.parametrize("top", [4097, 5000, 20000])
def test_top_above_4096(qb, oracle, top):
    """get_oversampled_top (vector_index_search_common.rs:27-46) easily exceeds 4096: the scan must answer, not fail."""
    rng = np.random.default_rng(8)
    n = 30_000 if top < 20000 else 100_000
    base = rng.standard_normal((n, 48)).astype(np.float32)
    q = rng.standard_normal((2, 48)).astype(np.float32)
    st = qb.DenseVectorStorage(base, qb.Distance.Dot)
    got = st.search_batch(q, top)
    want = oracle.scan_f32(oracle.DOT, base, q, top)
    for a, b in zip(got, want):
        assert a.size == top
        assert_topk_equal(a, b, what=f"top {top}")
    orig = st.build_raw_scorer(q[0])
    cand = got[0]["idx"]
    res = qb.rescore(orig, cand, top - 7)
    assert_topk_equal(res, want[0][: top - 7], what="rescore large top")
    orig.close(); st.close()


def test_sharded_ids_through_every_entry_point(qb, oracle):
    """ids reported by a shard are local row + id_base; the id-taking entry points take the same numbering."""
    from qdrant_b200._capi import check, lib

    rng = np.random.default_rng(6)
    n, dim, base_id = 20_000, 64, 1_000_000
    base = oracle.preprocess_rows_f32(oracle.COSINE, rng.standard_normal((n, dim)).astype(np.float32))
    q = rng.standard_normal(dim).astype(np.float32)
    qp = oracle.preprocess_f32(oracle.COSINE, q)
    st = qb.DenseVectorStorage(base, qb.Distance.Cosine)
    check(lib().qb_storage_set_id_base(st._h, base_id))
    got = st.search_batch(q, 50)[0]
    want = oracle.scan_f32(oracle.COSINE, base, qp[None], 50)[0]
    assert np.array_equal(got["idx"], want["idx"] + base_id) and np.array_equal(got["score"], want["score"])
    sc = st.build_raw_scorer(q)
    np.testing.assert_array_equal(sc.score_points(got["idx"]), want["score"])           # global ids accepted
    with pytest.raises(qb.QbError):
        sc.score_points(np.array([5], np.uint32))                                          # a local id is out of range now
    # oversample -> rescore on the shard: SQ8 shard reports global ids, the f32 scorer of the same shard rescores them
    dt, inv = qb.construct_vector_parameters(qb.Distance.Cosine)
    sq = oracle.SQ8.encode(base, int(dt), bool(inv))
    qs = qb.ScalarQuantizedVectors(sq.rows, dim, sq.meta.alpha, sq.meta.offset, sq.meta.multiplier, qb.Distance.Cosine)
    check(lib().qb_storage_set_id_base(qs._h, base_id))
    over = qs.search_batch(q, 200)[0]
    assert over["idx"].min() >= base_id
    res = qb.rescore(sc, over["idx"], 10)
    exact = oracle.score_points_f32(oracle.COSINE, base, qp, (over["idx"] - base_id).astype(np.uint32))
    order = np.lexsort((over["idx"], -exact.astype(np.float64)))[:10]
    np.testing.assert_array_equal(res["idx"], over["idx"][order])
    np.testing.assert_array_equal(res["score"], exact[order])
    # id_list filter with global ids
    ids = (rng.choice(n, 3000, replace=False).astype(np.uint32) + base_id)
    f = st.search_batch(q, 10, id_list=ids)[0]
    w = oracle.scan_f32(oracle.COSINE, base[ids - base_id], qp[None], 10)[0]
    np.testing.assert_array_equal(f["score"], w["score"])
    assert set(f["idx"].tolist()) <= set(ids.tolist())
    assert sc.score_internal(base_id + 3, base_id + 9) == oracle.score_points_f32(oracle.COSINE, base, base[3], np.array([9], np.uint32))[0]
    sc.close(); st.close(); qs.close()


def test_pq_bq_create_from_device_pointers(qb, oracle):
    import torch
    import ctypes as C
    from qdrant_b200._capi import check, f32p, lib, u8p, u32p, vp

    rng = np.random.default_rng(4)
    n, dim, chunk = 4_000, 64, 4
    base = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal(dim).astype(np.float32)
    cents = rng.standard_normal((256, dim)).astype(np.float32)
    pq = oracle.PQ.encode(base, chunk, cents, oracle.QD_DOT, False)
    host = qb.ProductQuantizedVectors(pq.codes, cents, chunk, dim, qb.Distance.Dot)
    want = host.search_batch(q, 10)[0]
    m = dim // chunk
    div = np.array([[j * chunk, (j + 1) * chunk] for j in range(m)], np.uint32)
    d_codes = torch.from_numpy(pq.codes).cuda(); d_cents = torch.from_numpy(cents).cuda()
    h = vp()
    check(lib().qb_storage_create_pq(0, dim, m, div.ctypes.data_as(u32p), C.cast(d_cents.data_ptr(), f32p), 256, C.cast(d_codes.data_ptr(), u8p), n, int(qb.DistanceType.Dot), 0,
                                     int(qb.Distance.Dot), C.byref(h)))
    out = np.zeros(10, dtype=qb.SCORED_POINT_OFFSET); cnt = np.zeros(1, np.uint32)
    check(lib().qb_search_batch(h, q.ctypes.data_as(f32p), 1, 10, None, None, 0, None, out.ctypes.data_as(C.POINTER(qb.ScoredPoint)), cnt.ctypes.data_as(u32p), None))
    np.testing.assert_array_equal(out, want)
    lib().qb_storage_destroy(h)
    host.close()
    # BQ rows from device memory
    rows = oracle.BQ.encode(base, oracle.BQ_ONE, oracle.BQQ_SAME, oracle.QD_DOT, False).rows
    if True:
        hb = qb.BinaryQuantizedVectors(rows, dim, qb.Distance.Dot)
        wb = hb.search_batch(q, 10)[0]
        d_rows = torch.from_numpy(rows).cuda()
        h2 = vp()
        check(lib().qb_storage_create_bq(0, dim, 0, 0, C.cast(d_rows.data_ptr(), u8p), rows.shape[1], n, int(qb.DistanceType.Dot), 0, None, int(qb.Distance.Dot), C.byref(h2)))
        check(lib().qb_search_batch(h2, q.ctypes.data_as(f32p), 1, 10, None, None, 0, None, out.ctypes.data_as(C.POINTER(qb.ScoredPoint)), cnt.ctypes.data_as(u32p), None))
        np.testing.assert_array_equal(out, wb)
        lib().qb_storage_destroy(h2); hb.close()


def test_short_bitmap_is_rejected_and_io_counter_follows_on_disk(qb, oracle):
    base = np.random.default_rng(1).standard_normal((1000, 32)).astype(np.float32)
    st = qb.DenseVectorStorage(base, qb.Distance.Dot)
    with pytest.raises(ValueError):
        st.search_batch(base[0], 5, point_deleted=np.zeros(3, np.uint64))   # needs ceil(1000 / 64) = 16 words
    cnt = qb.HwCounters()
    st.search_batch(base[0], 5, counters=cnt)
    assert (cnt.cpu, cnt.vector_io_read) == (1000 * 32 * 4, 0)              # RAM storage: io multiplier 0 (metric_query_scorer.rs:44-48)
    st.set_on_disk(True)
    cnt = qb.HwCounters()
    st.search_batch(base[0], 5, counters=cnt)
    assert (cnt.cpu, cnt.vector_io_read) == (1000 * 32 * 4, 1000 * 32 * 4)
    sc = st.build_raw_scorer(base[1])
    sc.score_points(np.arange(7, dtype=np.uint32))
    assert sc.take_hardware_counters() == (7 * 32 * 4, 7 * 32 * 4)
    sc.close(); st.close()


@pytest.mark.parametrize("enc", ["OneBit", "TwoBits", "OneAndHalfBits"])
def test_bq_u8_word_rows_of_multivector_storages(qb, oracle, enc):
    """EncodedVectorsBin<u8> rows (ceil(bits / 8) bytes, quantized_vectors.rs:270-282) score exactly like the u128 rows they are a prefix of."""
    rng = np.random.default_rng(7)
    n, dim = 2000, 100                                   # 100 bits: 13 bytes as u8 words, 16 bytes as one u128 word
    base = rng.standard_normal((n, dim)).astype(np.float32)
    e = int(getattr(qb.BQEncoding, enc))
    ms = oracle.bq_mean_std(base) if e != oracle.BQ_ONE else None
    bq = oracle.BQ.encode(base, e, oracle.BQQ_SCALAR8, oracle.QD_DOT, False, ms)
    ext = {0: dim, 1: 2 * dim, 2: (3 * dim + 1) // 2}[e]
    rows8 = np.ascontiguousarray(bq.rows[:, : (ext + 7) // 8])
    assert not bq.rows[:, (ext + 7) // 8 :].any()
    a = qb.BinaryQuantizedVectors(bq.rows, dim, qb.Distance.Dot, qb.BQEncoding(e), qb.BQQueryEncoding.Scalar8bits, ms)
    b = qb.BinaryQuantizedVectors(rows8, dim, qb.Distance.Dot, qb.BQEncoding(e), qb.BQQueryEncoding.Scalar8bits, ms)
    q = rng.standard_normal((3, dim)).astype(np.float32)
    for x, y in zip(a.search_batch(q, 10), b.search_batch(q, 10)):
        np.testing.assert_array_equal(x, y)
    a.close(); b.close()
