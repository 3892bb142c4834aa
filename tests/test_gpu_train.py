"""Quantizer training on the device vs the CPU restatement of the reference (oracle/train.c): BQ vector statistics (deterministic in the
reference: bit-exact), SQ8 quantile interval and PQ k-means (deterministic GIVEN the sampled vectors, the thread count and the re-seed
rule: bit-exact under those inputs), then the full chain train -> encode -> search on the device == the same chain on the oracle."""
import ctypes as C

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


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_bq_vector_stats_bit_exact(qb, oracle, torch):
    from qdrant_b200._capi import check, f32p, lib, vp

    rng = np.random.default_rng(1)
    x = (rng.standard_normal((5000, 133)) * rng.uniform(0.1, 3, 133) + rng.uniform(-1, 1, 133)).astype(np.float32)
    d = dev(torch, x)
    ms, mm = np.zeros((133, 2), np.float32), np.zeros((133, 2), np.float32)
    check(lib().qb_bq_vector_stats_device(0, 133, 5000, vp(d.data_ptr()), 0, ms.ctypes.data_as(f32p), mm.ctypes.data_as(f32p)))
    want_ms, want_mm = oracle.bq_vector_stats(x)
    np.testing.assert_array_equal(ms.view(np.uint32), want_ms.view(np.uint32))
    np.testing.assert_array_equal(mm, want_mm)


@pytest.mark.parametrize("quantile", [0.99, 0.95, 0.5, 1.0])
def test_sq8_quantile_interval(qb, oracle, torch, quantile):
    from qdrant_b200._capi import check, f32p, i32p, lib, vp

    rng = np.random.default_rng(2)
    sample = rng.standard_normal((1500, 96)).astype(np.float32)     # the reference samples min(count, 5000) vectors (quantile.rs:10,292)
    d = dev(torch, sample)
    a, o, f = C.c_float(), C.c_float(), C.c_int32()
    check(lib().qb_sq8_quantile_interval_device(0, 96, 1500, vp(d.data_ptr()), 0, C.c_float(quantile), C.byref(a), C.byref(o), C.byref(f)))
    want = oracle.sq8_quantile_interval(sample, quantile)
    if want is None:
        assert f.value == 0
    else:
        assert f.value == 1 and (np.float32(a.value), np.float32(o.value)) == want


@pytest.mark.parametrize("n,dim,chunk,K,groups", [(3000, 64, 8, 256, 8), (1000, 50, 16, 64, 3), (100, 32, 4, 256, 4), (700, 24, 2, 256, 1)])
def test_pq_kmeans_equals_reference_restatement(qb, oracle, torch, n, dim, chunk, K, groups):
    from qdrant_b200._capi import check, f32p, lib, u32p, vp

    rng = np.random.default_rng(n)
    centers = rng.standard_normal((40, dim)).astype(np.float32)
    sample = (centers[rng.integers(0, 40, n)] + 0.3 * rng.standard_normal((n, dim))).astype(np.float32)
    d = dev(torch, sample)
    got = np.zeros((K, dim), np.float32)
    it = C.c_uint32()
    check(lib().qb_pq_train_device(0, dim, chunk, K, n, vp(d.data_ptr()), 0, 100, C.c_float(1e-5), groups, 77, got.ctypes.data_as(f32p), C.byref(it)))
    want, want_it = oracle.kmeans_pq(sample, chunk, K, 100, 1e-5, groups, 77)
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    assert it.value >= want_it or n <= K


def test_train_encode_search_chain(qb, oracle, torch):
    """Everything a segment build does for PQ, on the device: k-means on a sample, encode all rows, upload, search — equal to the oracle chain."""
    from qdrant_b200._capi import check, f32p, lib, vp

    rng = np.random.default_rng(9)
    n, dim, chunk = 20_000, 64, 4
    base = rng.standard_normal((n, dim)).astype(np.float32)
    sample = np.ascontiguousarray(base[np.sort(rng.choice(n, 4000, replace=False))])
    d_base, d_sample = dev(torch, base), dev(torch, sample)
    cents = np.zeros((256, dim), np.float32)
    check(lib().qb_pq_train_device(0, dim, chunk, 256, 4000, vp(d_sample.data_ptr()), 0, 20, C.c_float(1e-5), 4, 5, cents.ctypes.data_as(f32p), None))
    codes = torch.zeros((n, dim // chunk), dtype=torch.uint8, device="cuda")
    check(lib().qb_pq_encode_rows_device(0, dim, chunk, 256, cents.ctypes.data_as(f32p), n, vp(d_base.data_ptr()), 0, vp(codes.data_ptr()), None))
    torch.cuda.synchronize()
    want_cents, _ = oracle.kmeans_pq(sample, chunk, 256, 20, 1e-5, 4, 5)
    np.testing.assert_array_equal(cents, want_cents)
    pq = oracle.PQ.encode(base, chunk, want_cents, oracle.QD_DOT, False)
    np.testing.assert_array_equal(codes.cpu().numpy(), pq.codes)
    st = qb.ProductQuantizedVectors(pq.codes, cents, chunk, dim, qb.Distance.Dot)
    q = rng.standard_normal((3, dim)).astype(np.float32)
    luts = np.stack([pq.encode_query(x) for x in q])
    for a, b in zip(st.search_batch(q, 10), pq.scan(luts, 10)):
        np.testing.assert_array_equal(a["score"], b["score"])
    st.close()
