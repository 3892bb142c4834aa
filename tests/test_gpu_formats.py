"""A segment's files as they lie on disk -> HBM storages: matrix.dat, quantized.data + quantized.meta.json written
the way the reference writes them (serde_json field names / enum spellings, headerless fixed-stride rows), loaded through
qb_storage_load_*, must search exactly like storages created from the decoded fields."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qb():
    from qdrant_b200 import scorer

    return scorer


def f32(x):
    return float(np.float32(x))   # rendered as the f64 value of the f32: parses back to the same f32 (serde prints the shortest form of it)


def vector_parameters(dim, dt, invert):
    return {"dim": dim, "distance_type": {0: "Cosine", 1: "Dot", 2: "L1", 3: "L2"}[int(dt)], "invert": bool(invert)}


def same_search(a, b, queries, top=10):
    for x, y in zip(a.search_batch(queries, top), b.search_batch(queries, top)):
        np.testing.assert_array_equal(x, y)


def test_matrix_dat(qb, oracle):
    rng = np.random.default_rng(1)
    base = oracle.preprocess_rows_f32(oracle.COSINE, rng.standard_normal((3000, 70)).astype(np.float32))
    blob = b"data" + base.tobytes() + b"\0" * 100          # mmap files are over-allocated: a trailing partial row is ignored
    st = qb.load_dense_file(blob, qb.Distance.Cosine, 70)
    ref = qb.DenseVectorStorage(base, qb.Distance.Cosine)
    assert (st.count, st.dim) == (3000, 70)
    same_search(st, ref, rng.standard_normal((3, 70)).astype(np.float32))
    with pytest.raises(qb.QbError):
        qb.load_dense_file(b"drop" + base.tobytes(), qb.Distance.Cosine, 70)   # deleted.dat's header, not matrix.dat's
    h16 = base.astype(np.float16)
    st16 = qb.load_dense_file(b"data" + h16.tobytes(), qb.Distance.Dot, 70, qb.VectorStorageDatatype.Float16)
    ref16 = qb.DenseVectorStorage(h16, qb.Distance.Dot, qb.VectorStorageDatatype.Float16)
    same_search(st16, ref16, rng.standard_normal((2, 70)).astype(np.float32))
    for s in (st, ref, st16, ref16):
        s.close()


@pytest.mark.parametrize("dist", ["Cosine", "Euclid"])
def test_scalar_quantized_files(qb, oracle, dist):
    d = getattr(qb.Distance, dist)
    dt, inv = qb.construct_vector_parameters(d)
    rng = np.random.default_rng(2)
    base = rng.standard_normal((5000, 100)).astype(np.float32)
    if d == qb.Distance.Cosine:
        base = oracle.preprocess_rows_f32(oracle.COSINE, base)
    sq = oracle.SQ8.encode(base, int(dt), bool(inv))
    meta = json.dumps({"actual_dim": sq.meta.actual_dim, "alpha": f32(sq.meta.alpha), "offset": f32(sq.meta.offset), "multiplier": f32(sq.meta.multiplier),
                       "vector_parameters": vector_parameters(100, dt, inv)})
    st = qb.load_quantized(meta, sq.rows, d, count=5000)
    ref = qb.ScalarQuantizedVectors(sq.rows, 100, sq.meta.alpha, sq.meta.offset, sq.meta.multiplier, d)
    assert st.count == 5000
    same_search(st, ref, rng.standard_normal((4, 100)).astype(np.float32))
    st.close(); ref.close()


def test_product_quantized_files(qb, oracle):
    rng = np.random.default_rng(3)
    n, dim, chunk = 4000, 60, 8      # 60 / 8: last chunk is short (vector_division ranges, encoded_vectors_pq.rs:164-169)
    base = rng.standard_normal((n, dim)).astype(np.float32)
    cents = (rng.standard_normal((256, dim)) * 0.5).astype(np.float32)
    pq = oracle.PQ.encode(base, chunk, cents, oracle.QD_DOT, False)
    starts = list(range(0, dim, chunk))
    meta = json.dumps({"centroids": [[f32(v) for v in row] for row in cents], "vector_division": [{"start": s, "end": min(s + chunk, dim)} for s in starts],
                       "vector_parameters": dict(vector_parameters(dim, 1, False), count=n)})   # the deprecated `count` field may still be present
    st = qb.load_quantized(meta, pq.codes, qb.Distance.Dot)
    ref = qb.ProductQuantizedVectors(pq.codes, cents, chunk, dim, qb.Distance.Dot)
    assert st.count == n
    same_search(st, ref, rng.standard_normal((3, dim)).astype(np.float32))
    st.close(); ref.close()


@pytest.mark.parametrize("enc,qenc", [("OneBit", "SameAsStorage"), ("TwoBits", "Scalar8bits"), ("OneAndHalfBits", "Scalar4bits")])
def test_binary_quantized_files(qb, oracle, enc, qenc):
    rng = np.random.default_rng(4)
    n, dim = 3000, 200
    base = rng.standard_normal((n, dim)).astype(np.float32)
    e, qe = int(getattr(qb.BQEncoding, enc)), int(getattr(qb.BQQueryEncoding, qenc))
    ms = oracle.bq_mean_std(base) if e != oracle.BQ_ONE else None
    bq = oracle.BQ.encode(base, e, qe, oracle.QD_DOT, False, ms)
    meta = {"vector_parameters": vector_parameters(dim, 1, False)}
    if enc != "OneBit":
        meta["encoding"] = enc                      # skip_serializing_if = is_one
    if qenc != "SameAsStorage":
        meta["query_encoding"] = qenc
    if ms is not None:
        m2 = np.asarray(ms, np.float32).reshape(dim, 2)
        meta["vector_stats"] = {"elements_stats": [{"min": -1.0, "max": 1.0, "mean": f32(a), "stddev": f32(b)} for a, b in m2]}
    st = qb.load_quantized(json.dumps(meta), bq.rows, qb.Distance.Dot)
    ref = qb.BinaryQuantizedVectors(bq.rows, dim, qb.Distance.Dot, qb.BQEncoding(e), qb.BQQueryEncoding(qe), ms)
    same_search(st, ref, rng.standard_normal((3, dim)).astype(np.float32))
    st.close(); ref.close()


def test_malformed_metadata_is_rejected(qb):
    rows = np.zeros((10, 20), np.uint8)
    for bad in ("", "[]", "{", '{"vector_parameters": {"dim": 16}}', '{"actual_dim": 16, "alpha": 1, "vector_parameters": {"dim": 16, "distance_type": "Dot", "invert": false}}',
                '{"vector_parameters": {"dim": 16, "distance_type": "Hamming", "invert": false}}'):
        with pytest.raises((qb.QbError, ValueError)):
            qb.load_quantized(bad, rows, qb.Distance.Dot)
