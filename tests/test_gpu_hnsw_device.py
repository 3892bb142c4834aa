"""Device-resident HNSW traversal (qb_hnsw_search_batch) vs the CPU traversal of the SAME graph with the CPU scorer
(oracle/hnsw.c restating graph_layers.rs:108-148,247-316,530-561).  Gate: identical result lists (tie-aware, score bits
equal) and identical scorer-call / scored-point counts => recall difference 0 (north_star: |delta recall@10| <= 1e-4)."""
import numpy as np
import pytest

from tests.util import assert_topk_equal, pack_bitmap

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qb():
    from qdrant_b200 import scorer

    return scorer


def _graph(oracle, base, dist, threads=4, m=16):
    g = oracle.HNSW(base, dist, m=m, ef_construct=64, seed=11, threads=threads)
    entry, entry_level, gm, gm0 = g.entry()
    return g, g.export_plain(), entry, entry_level, gm, gm0


@pytest.mark.parametrize("dist,dim,n", [("Cosine", 96, 20_000), ("Euclid", 100, 6_000), ("Dot", 8, 3_000), ("Manhattan", 40, 3_000), ("Cosine", 768, 4_000)])
def test_device_traversal_equals_cpu_traversal_f32(qb, oracle, dist, dim, n):
    d = getattr(qb.Distance, dist)
    rng = np.random.default_rng(3)
    base = rng.standard_normal((n, dim)).astype(np.float32)
    if d == qb.Distance.Cosine:
        base = oracle.preprocess_rows_f32(oracle.COSINE, base)
    queries = rng.standard_normal((70, dim)).astype(np.float32)
    qp = np.stack([oracle.preprocess_f32(int(d), q) for q in queries])
    g, blob, entry, lvl, m, m0 = _graph(oracle, base, int(d))
    st = qb.DenseVectorStorage(base, d)
    hg = qb.HnswGraph(st, blob, m, m0)
    for top, ef in ((10, 128), (5, 16), (40, 20)):
        g.stats(reset=True); hg.stats(reset=True)
        want = g.search_batch(qp, top, ef, threads=2)
        cnt = qb.HwCounters()
        got = hg.search(queries, top, ef, entry, lvl, counters=cnt)
        for i, (a, b) in enumerate(zip(got, want)):
            assert_topk_equal(a, b, what=f"{dist} dim {dim} top {top} ef {ef} query {i}")
        calls, scored = g.stats(reset=True)
        assert cnt.cpu == scored * dim * 4            # same number of scored points as the CPU traversal ...
        assert hg.stats(reset=True) == (calls, scored)  # ... in the same number of scorer calls (hops)
    # a second identical batch must leave the visited bitmaps clean (same answers again)
    again = hg.search(queries, 10, 128, entry, lvl)
    for a, b in zip(again, g.search_batch(qp, 10, 128)):
        assert_topk_equal(a, b, what="second batch")
    hg.close(); st.close(); g.close()


def test_device_traversal_with_filter(qb, oracle):
    n, dim = 8_000, 64
    rng = np.random.default_rng(5)
    base = oracle.preprocess_rows_f32(oracle.COSINE, rng.standard_normal((n, dim)).astype(np.float32))
    queries = rng.standard_normal((40, dim)).astype(np.float32)
    qp = np.stack([oracle.preprocess_f32(oracle.COSINE, q) for q in queries])
    g, blob, entry, lvl, m, m0 = _graph(oracle, base, oracle.COSINE)
    deleted = rng.random(n) < 0.3
    deleted[entry] = False                    # get_entry_point would pick another entry for a filtered-out one (host logic)
    st = qb.DenseVectorStorage(base, qb.Distance.Cosine)
    hg = qb.HnswGraph(st, blob, m, m0)
    want = g.search_batch(qp, 10, 64, deleted=pack_bitmap(deleted))
    got = hg.search(queries, 10, 64, entry, lvl, point_deleted=deleted)
    for a, b in zip(got, want):
        assert_topk_equal(a, b, what="filtered")
        assert not deleted[a["idx"]].any()
    # resident soft-deletes behave the same
    st.set_deleted(deleted)
    for a, b in zip(hg.search(queries, 10, 64, entry, lvl), want):
        assert_topk_equal(a, b, what="resident deleted flags")
    hg.close(); st.close(); g.close()


def test_device_traversal_sq8(qb, oracle):
    n, dim = 6_000, 96
    d = qb.Distance.Cosine
    rng = np.random.default_rng(9)
    base = oracle.preprocess_rows_f32(oracle.COSINE, rng.standard_normal((n, dim)).astype(np.float32))
    queries = rng.standard_normal((24, dim)).astype(np.float32)
    g, blob, entry, lvl, m, m0 = _graph(oracle, base, int(d))
    dt, inv = qb.construct_vector_parameters(d)
    sq = oracle.SQ8.encode(base, int(dt), bool(inv))
    qst = qb.ScalarQuantizedVectors(sq.rows, dim, sq.meta.alpha, sq.meta.offset, sq.meta.multiplier, d)
    hg = qb.HnswGraph(qst, blob, m, m0)
    got = hg.search(queries, 10, 64, entry, lvl)
    for q, a in zip(queries, got):
        qp = oracle.preprocess_f32(int(d), q)
        code, off = sq.encode_query(qp)
        want = g.search(qp, 10, 64, score_points=lambda ids, code=code, off=off: np.array([sq.score(code, off, int(i)) for i in ids], np.float32))
        assert_topk_equal(a, want, what="sq8 traversal")
    hg.close(); qst.close(); g.close()


def test_hnsw_create_rejects_bad_blobs(qb, oracle):
    base = np.random.default_rng(1).standard_normal((500, 32)).astype(np.float32)
    g, blob, entry, lvl, m, m0 = _graph(oracle, base, oracle.DOT, threads=1)
    st = qb.DenseVectorStorage(base, qb.Distance.Dot)
    with pytest.raises(qb.QbError):
        qb.HnswGraph(st, blob[:100], m, m0)                     # truncated
    bad = blob.copy(); bad[:8].view(np.uint64)[0] = 499          # point count != storage count
    with pytest.raises(qb.QbError):
        qb.HnswGraph(st, bad, m, m0)
    hg = qb.HnswGraph(st, blob, m, m0)
    with pytest.raises(qb.QbError):
        hg.search(base[:2], 10, 64, 10_000, 0)                   # entry point out of range
    st2 = qb.DenseVectorStorage(base.astype(np.float16), qb.Distance.Dot, qb.VectorStorageDatatype.Float16)
    hg2 = qb.HnswGraph(st2, blob, m, m0)
    with pytest.raises(qb.QbError):
        hg2.search(base[:2], 10, 64, entry, lvl)                 # f16 storages go through qb_score_points per hop
    hg.close(); hg2.close(); st.close(); st2.close(); g.close()
