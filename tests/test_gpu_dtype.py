"""GPU parity: Float16 and Uint8 dense storages vs the CPU oracle.

u8: bit-exact against the restated avx2 kernels (metric_uint/avx2/*.rs) incl. the reference's KAT vectors.
f16: the oracle restates metric_f16/avx/*.rs with F16C intrinsics; parity is bit-exact as well (the reference's own
     f16 test only asks for rel 5e-4, metric_f16/avx/dot.rs:124)."""
import numpy as np
import pytest

from tests.util import assert_topk_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qb():
    from qdrant_b200 import scorer

    return scorer


@pytest.mark.parametrize("dist", ["Cosine", "Dot", "Euclid", "Manhattan"])
@pytest.mark.parametrize("dim", [8, 24, 32, 100, 300, 768, 1030])
def test_u8_bit_exact(qb, oracle, dist, dim):
    d = getattr(qb.Distance, dist)
    rng = np.random.default_rng(dim)
    base = rng.integers(0, 256, (400, dim), dtype=np.uint8)
    base[0] = 0
    base[1] = 255
    st = qb.DenseVectorStorage(base, d, qb.VectorStorageDatatype.Uint8)
    ids = np.arange(400, dtype=np.uint32)
    for q in (rng.uniform(0, 255, dim).astype(np.float32), np.full(dim, 255.0, np.float32), np.full(dim, 300.7, np.float32),
              rng.uniform(-20, 40, dim).astype(np.float32)):
        qu8 = np.clip(np.trunc(q), 0, 255).astype(np.uint8)  # `x as u8` (data_types/primitive.rs:127-129)
        want = np.array([oracle.similarity_u8(int(d), qu8, base[i]) for i in range(400)], np.float32)
        sc = st.build_raw_scorer(q)
        np.testing.assert_array_equal(sc.score_points(ids), want)
        sc.close()
    isc = st.raw_internal_scorer(1)
    want = np.array([oracle.similarity_u8(int(d), base[1], base[i]) for i in range(400)], np.float32)
    np.testing.assert_array_equal(isc.score_points(ids), want)
    st.close()


def test_u8_reference_kat(qb, oracle):
    from tests.test_oracle_kat import u8_kat

    v1, v2 = u8_kat()
    for dist, name in ((qb.Distance.Dot, "dot"), (qb.Distance.Cosine, "cosine"), (qb.Distance.Euclid, "euclid"), (qb.Distance.Manhattan, "manhattan")):
        st = qb.DenseVectorStorage(v2[None, :], dist, qb.VectorStorageDatatype.Uint8)
        got = st.build_raw_scorer(v1.astype(np.float32)).score_point(0)
        assert got == oracle.raw_u8(name, "scalar", v1, v2) == oracle.raw_u8(name, "avx", v1, v2)
        st.close()


def test_u8_search(qb, oracle):
    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, (70_000, 64), dtype=np.uint8)
    st = qb.DenseVectorStorage(base, qb.Distance.Euclid, qb.VectorStorageDatatype.Uint8)
    q = rng.uniform(0, 255, 64).astype(np.float32)
    got = st.search_batch(q, 10)[0]
    qu8 = np.trunc(q).astype(np.uint8)
    sc = np.array([oracle.similarity_u8(oracle.EUCLID, qu8, base[i]) for i in range(base.shape[0])], np.float32)
    assert_topk_equal(got, oracle.topk(sc, 10), sc)
    st.close()


@pytest.mark.parametrize("dist", ["Cosine", "Dot", "Euclid", "Manhattan"])
@pytest.mark.parametrize("dim", [8, 20, 32, 70, 256, 768])
def test_f16_bit_exact(qb, oracle, dist, dim):
    d = getattr(qb.Distance, dist)
    rng = np.random.default_rng(dim + 1)
    base32 = rng.standard_normal((300, dim)).astype(np.float32)
    if d == qb.Distance.Cosine:
        base32 = oracle.preprocess_rows_f32(oracle.COSINE, base32)
    base = base32.astype(np.float16)
    st = qb.DenseVectorStorage(base, d, qb.VectorStorageDatatype.Float16)
    ids = np.arange(300, dtype=np.uint32)
    for _ in range(3):
        q = rng.standard_normal(dim).astype(np.float32)
        qh = oracle.preprocess_f32(int(d), q).astype(np.float16)  # f16::from_f32 after Metric::preprocess
        want = np.array([oracle.similarity_f16(int(d), qh, base[i]) for i in range(300)], np.float32)
        sc = st.build_raw_scorer(q)
        np.testing.assert_array_equal(sc.score_points(ids), want)
        sc.close()
    st.close()
