"""Config #5 harness: HNSW (M=16, ef=128) graph search driven by the GPU RawScorer vs the same traversal driven by the
CPU scorer.  Traversal stays on the CPU exactly as in the reference (GraphLayers::search -> FilteredScorer::score_points,
<= 2M ids per hop); only the scorer behind the trait changes.  Gate: identical result lists => recall@10 difference 0
(north_star asks for |delta recall| <= 1e-4, tie-aware)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qb():
    from qdrant_b200 import scorer

    return scorer


def recall(res, exact):
    hit = tot = 0
    for r, e in zip(res, exact):
        kth = e["score"][-1]
        hit += sum(1 for s in r["score"] if s >= kth)  # tie-aware: anything at least as good as the exact k-th counts
        tot += e.size
    return hit / tot


@pytest.mark.parametrize("dist,quant", [("Cosine", None), ("Euclid", None), ("Cosine", "sq8")])
def test_hnsw_gpu_scorer_matches_cpu_scorer(qb, oracle, dist, quant):
    n, dim, nq, top, ef = 30_000, 96, 60, 10, 128
    d = getattr(qb.Distance, dist)
    rng = np.random.default_rng(7)
    base = rng.standard_normal((n, dim)).astype(np.float32)
    if d == qb.Distance.Cosine:
        base = oracle.preprocess_rows_f32(oracle.COSINE, base)
    queries = rng.standard_normal((nq, dim)).astype(np.float32)
    graph = oracle.HNSW(base, int(d), m=16, ef_construct=100, seed=42)   # built once on the CPU (single-threaded, deterministic)
    dense = qb.DenseVectorStorage(base, d)
    sq = qst = None
    if quant == "sq8":
        dt, inv = qb.construct_vector_parameters(d)
        sq = oracle.SQ8.encode(base, int(dt), bool(inv))
        qst = qb.ScalarQuantizedVectors(sq.rows, dim, sq.meta.alpha, sq.meta.offset, sq.meta.multiplier, d)
    cpu_res, gpu_res, exact = [], [], []
    calls = []
    for q in queries:
        qp = oracle.preprocess_f32(int(d), q)
        if quant == "sq8":
            code, off = sq.encode_query(qp)
            cpu_scorer = lambda ids, code=code, off=off: np.array([sq.score(code, off, int(i)) for i in ids], np.float32)
            fs = qb.FilteredScorer.new(q, dense, qst)       # quantized scorer preferred (point_scorer.rs:172-175)
        else:
            cpu_scorer = None                                # oracle's built-in f32 scorer
            fs = qb.FilteredScorer.new(q, dense, None)
        graph.stats(reset=True)
        cpu_res.append(graph.search(qp, top, ef, score_points=cpu_scorer))
        gpu_res.append(graph.search(qp, top, ef, score_points=lambda ids, fs=fs: fs.raw_scorer.score_points(ids)))
        calls.append(graph.stats(reset=True))
        exact.append(oracle.scan_f32(int(d), base, qp, top)[0])
    for a, b in zip(cpu_res, gpu_res):
        np.testing.assert_array_equal(a, b)                 # same traversal, same scores, same result
    r_cpu, r_gpu = recall(cpu_res, exact), recall(gpu_res, exact)
    assert abs(r_cpu - r_gpu) <= 1e-4
    assert r_cpu > (0.6 if quant is None else 0.4)  # graph quality on random 96-d data; the gate is CPU == GPU above
    hops = np.mean([c[0] for c in calls]); evals = np.mean([c[1] for c in calls])
    print(f"{dist} {quant}: recall@10 cpu={r_cpu:.4f} gpu={r_gpu:.4f}; {hops:.0f} score_points calls, {evals:.0f} evaluations per query")
    graph.close(); dense.close()
    if qst:
        qst.close()
