"""Full BASELINE sizes on the GPU, checked through size-independent properties (the oracle cannot scan 10M x 768 per
test in seconds): sortedness, planted needles, exact re-scoring of every returned id with the oracle on rows read back
from HBM, no better row in a large random sample, and shard-merge consistency (top-k of the whole == merge of top-k of
the halves)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, DIM, TOP = 10_000_000, 768, 10


@pytest.fixture(scope="module")
def big():
    import torch

    from qdrant_b200 import scorer as qb
    from qdrant_b200._capi import check, lib, vp

    if torch.cuda.mem_get_info()[1] < 60e9:
        pytest.skip("needs > 60 GB of HBM")
    dev = torch.device("cuda", 0)
    st = qb.DenseVectorStorage(None, qb.Distance.Cosine, count=N, dim=DIM)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    chunk = 500_000
    for r0 in range(0, N, chunk):
        x = torch.randn((chunk, DIM), generator=g, device=dev)
        check(lib().qb_metric_preprocess_device(0, int(qb.Distance.Cosine), DIM, chunk, vp(x.data_ptr()), DIM * 4))
        st.write_rows_device(r0, chunk, x.data_ptr(), DIM * 4)
        del x
    torch.cuda.synchronize()
    yield qb, st
    st.close()


def test_c2_properties(big, oracle):
    qb, st = big
    rng = np.random.default_rng(5)
    queries = rng.standard_normal((4, DIM)).astype(np.float32)
    # plant needles: the (normalised) stored rows 123 and 9_999_999 themselves are queries -> they must come back first with score ~1
    needles = st.get_dense(np.array([123, N - 1], np.uint32))
    queries[0], queries[1] = needles[0], needles[1]
    res = st.search_batch(queries, TOP)
    assert res[0]["idx"][0] == 123 and res[1]["idx"][0] == N - 1
    sample_ids = rng.choice(N, 200_000, replace=False).astype(np.uint32)
    for qi in (0, 3):
        r = res[qi]
        qp = oracle.preprocess_f32(oracle.COSINE, queries[qi])
        assert r.size == TOP and np.all(r["score"][:-1] >= r["score"][1:])
        rows = st.get_dense(r["idx"])
        np.testing.assert_array_equal(r["score"], oracle.score_rows_f32(oracle.COSINE, rows, qp))      # bit-exact re-score
        srows = st.get_dense(sample_ids)
        ssc = oracle.score_rows_f32(oracle.COSINE, srows, qp)
        better = sample_ids[ssc > r["score"][-1]]
        assert set(better.tolist()) <= set(r["idx"].tolist())                                          # nothing better was missed
    # batched == per query, and merge of halves == whole
    single = st.search_batch(queries[3], TOP)[0]
    np.testing.assert_array_equal(single, res[3])
    from qdrant_b200.sharded import merge_topk_host

    lo = st.search_batch(queries[3], TOP, id_list=None, point_deleted=None)[0]
    h1 = st.search_batch(queries[3], TOP, id_list=np.arange(0, 3_000_000, dtype=np.uint32))[0]
    mask = np.zeros(N, bool); mask[:3_000_000] = True
    h2 = st.search_batch(queries[3], TOP, point_deleted=mask)[0]
    np.testing.assert_array_equal(merge_topk_host([h1, h2], TOP), lo)
