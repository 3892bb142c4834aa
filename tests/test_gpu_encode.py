"""GPU parity: the quantizers' encode step on the device vs the CPU oracle's restatement of the reference encoders —
bit-exact rows (encoded_vectors_u8.rs:143-316, encoded_vectors_binary.rs:531-671, encoded_vectors_pq.rs:301-329).
Shapes follow lib/quantization/tests/integration (dim 65 exercises the alignment tail) plus the BASELINE dims."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qb():
    from qdrant_b200 import scorer

    return scorer


@pytest.fixture(scope="module")
def torch():
    import torch

    return torch


def on_device(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("dist", ["Cosine", "Dot", "Euclid", "Manhattan"])
@pytest.mark.parametrize("n,dim", [(129, 65), (500, 16), (300, 768), (64, 1536), (50, 2000)])
def test_sq8_encode_rows_bit_exact(qb, oracle, torch, dist, n, dim):
    d = getattr(qb.Distance, dist)
    dt, inv = qb.construct_vector_parameters(d)
    rng = np.random.default_rng(42)
    base = rng.standard_normal((n, dim)).astype(np.float32)
    if d == qb.Distance.Cosine:
        base = oracle.preprocess_rows_f32(oracle.COSINE, base)
    base[3, 5] = np.float32(0.0)
    want = oracle.SQ8.encode(base, int(dt), bool(inv))
    x = on_device(torch, base)
    alpha, offset = qb.sq8_find_alpha_offset(x.data_ptr(), n, dim)
    assert alpha == np.float32(want.meta.alpha) and offset == np.float32(want.meta.offset)
    assert qb.sq8_multiplier(alpha, d) == np.float32(want.meta.multiplier)
    ad = dim + (16 - dim % 16) % 16
    out = torch.zeros((n, 4 + ad), dtype=torch.uint8, device="cuda")
    qb.sq8_encode_rows(x.data_ptr(), n, dim, alpha, offset, d, out.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), np.asarray(want.rows).reshape(n, 4 + ad))
    # the encoded rows feed the storage directly (device pointer) and score like the oracle's
    st = qb.ScalarQuantizedVectors(None, dim, alpha, offset, qb.sq8_multiplier(alpha, d), d, rows_ptr=out.data_ptr(), count=n)
    q = rng.standard_normal(dim).astype(np.float32)
    code, off = want.encode_query(oracle.preprocess_f32(int(d), q))
    ref = np.array([want.score(code, off, i) for i in range(n)], np.float32)
    np.testing.assert_array_equal(st.raw_scorer(q).score_points(np.arange(n, dtype=np.uint32)), ref)
    st.close()


def test_sq8_encode_rows_strided_and_extreme_values(qb, oracle, torch):
    n, dim = 40, 100
    rng = np.random.default_rng(1)
    base = rng.standard_normal((n, dim)).astype(np.float32) * 5
    want = oracle.SQ8.encode(base, oracle.QD_DOT, False)
    padded = np.zeros((n, dim + 28), np.float32)
    padded[:, :dim] = base
    padded[:, dim:] = 1e9  # must be ignored: outside the row
    x = on_device(torch, padded)
    alpha, offset = qb.sq8_find_alpha_offset(x.data_ptr(), n, dim, row_stride_bytes=(dim + 28) * 4)
    assert alpha == np.float32(want.meta.alpha) and offset == np.float32(want.meta.offset)
    out = torch.zeros((n, 4 + 112), dtype=torch.uint8, device="cuda")
    # clamp paths: encode with a narrower range than the data (what a quantile-clipped alpha/offset does, :194-208)
    a2, o2 = np.float32(alpha / 2), np.float32(offset / 2)
    meta2 = oracle.SQ8.encode(base, oracle.QD_DOT, False, alpha=a2, offset=o2) if "alpha" in oracle.SQ8.encode.__code__.co_varnames else None
    qb.sq8_encode_rows(x.data_ptr(), n, dim, alpha, offset, qb.Distance.Dot, out.data_ptr(), row_stride_bytes=(dim + 28) * 4)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), np.asarray(want.rows).reshape(n, 116))
    if meta2 is not None:
        qb.sq8_encode_rows(x.data_ptr(), n, dim, a2, o2, qb.Distance.Dot, out.data_ptr(), row_stride_bytes=(dim + 28) * 4)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(out.cpu().numpy(), np.asarray(meta2.rows).reshape(n, 116))


@pytest.mark.parametrize("enc", ["OneBit", "TwoBits", "OneAndHalfBits"])
@pytest.mark.parametrize("n,dim", [(129, 65), (200, 768), (77, 1000), (10, 1)])
def test_bq_encode_rows_bit_exact(qb, oracle, torch, enc, n, dim):
    e = getattr(qb.BQEncoding, enc)
    rng = np.random.default_rng(7)
    base = rng.standard_normal((n, dim)).astype(np.float32)
    base[:, 0] = 0.25 if dim > 1 else base[:, 0]  # a constant coordinate: stddev < EPSILON branch (:645-652)
    ms = oracle.bq_mean_std(base) if int(e) != oracle.BQ_ONE else None
    want = oracle.BQ.encode(base, int(e), oracle.BQQ_SAME, oracle.QD_DOT, False, ms)
    rb = qb.bq_row_bytes(dim, e)
    assert rb == np.asarray(want.rows).reshape(n, -1).shape[1]
    x = on_device(torch, base)
    out = torch.full((n, rb), 0xAA, dtype=torch.uint8, device="cuda")
    qb.bq_encode_rows(x.data_ptr(), n, dim, e, out.data_ptr(), ms)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), np.asarray(want.rows).reshape(n, rb))
    if int(e) != oracle.BQ_ONE:  # without statistics both bits are `v > 0` (:640-643)
        want0 = oracle.BQ.encode(base, int(e), oracle.BQQ_SAME, oracle.QD_DOT, False, None)
        qb.bq_encode_rows(x.data_ptr(), n, dim, e, out.data_ptr(), None)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(out.cpu().numpy(), np.asarray(want0.rows).reshape(n, rb))


@pytest.mark.parametrize("n,dim,chunk,k", [(129, 65, 2, 64), (300, 128, 8, 256), (100, 1536, 16, 256), (150, 70, 16, 100), (90, 96, 32, 256), (60, 100, 64, 32)])
def test_pq_encode_rows_bit_exact(qb, oracle, torch, n, dim, chunk, k):
    rng = np.random.default_rng(9)
    base = rng.standard_normal((n, dim)).astype(np.float32)
    cents = rng.standard_normal((k, dim)).astype(np.float32)
    cents[5] = cents[3]  # duplicate centroid: the first minimum must win (:318-326)
    want = oracle.PQ.encode(base, chunk, cents, oracle.QD_DOT, False)
    m = (dim + chunk - 1) // chunk
    x = on_device(torch, base)
    out = torch.full((n, m), 255, dtype=torch.uint8, device="cuda")
    qb.pq_encode_rows(x.data_ptr(), n, dim, chunk, cents, out.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), np.asarray(want.codes).reshape(n, m))


def test_encode_rejects_bad_arguments(qb, torch):
    from qdrant_b200._capi import QbError

    x = torch.zeros((4, 8), device="cuda")
    out = torch.zeros((4, 64), dtype=torch.uint8, device="cuda")
    with pytest.raises(QbError):
        qb.sq8_encode_rows(x.data_ptr(), 4, 8, 1.0, 0.0, qb.Distance.Dot, out.data_ptr(), row_stride_bytes=6)
    with pytest.raises(QbError):
        qb.pq_encode_rows(x.data_ptr(), 4, 8, 128, np.zeros((4, 8), np.float32), out.data_ptr())
