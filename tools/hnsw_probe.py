#!/usr/bin/env python
"""HNSW probe: device-resident traversal (qb_hnsw_search_batch) vs the CPU traversal on all host cores, same graph.
    python tools/hnsw_probe.py [rows] [dim] [queries] [ef]"""
import json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as o
from qdrant_b200 import scorer as qb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
ef = int(sys.argv[4]) if len(sys.argv) > 4 else 128
threads = os.cpu_count() or 1
rng = np.random.default_rng(42)
centers = rng.standard_normal((1024, dim)).astype(np.float32)
base = centers[rng.integers(0, 1024, n)] + 0.5 * rng.standard_normal((n, dim)).astype(np.float32)
base = o.preprocess_rows_f32(o.COSINE, base)
qr = np.random.default_rng(43)
queries = (centers[qr.integers(0, 1024, nq)] + 0.5 * qr.standard_normal((nq, dim))).astype(np.float32)
qp = o.preprocess_rows_f32(o.COSINE, queries)
t0 = time.perf_counter(); g = o.HNSW(base, o.COSINE, m=16, ef_construct=100, seed=42, threads=threads); build_s = time.perf_counter() - t0
entry, lvl, m, m0 = g.entry()
blob = g.export_plain()
st = qb.DenseVectorStorage(base, qb.Distance.Cosine)
hg = qb.HnswGraph(st, blob, m, m0)
variants = {}
for nt in (256, 128, 64):
    for nopf in (0, 1):
        qb.set_option("hnsw_threads", nt); qb.set_option("hnsw_no_prefetch", nopf)
        hg.search(queries, 10, ef, entry, lvl)      # same batch once untimed: scratch (visited bitmaps) is sized on first use
        t0 = time.perf_counter(); hg.search(queries, 10, ef, entry, lvl); dt = time.perf_counter() - t0
        variants[f"threads{nt}_{'noprefetch' if nopf else 'prefetch'}"] = nq / dt
qb.set_option("hnsw_threads", 0); qb.set_option("hnsw_no_prefetch", 0)
hg.search(queries[:256], 10, ef, entry, lvl)
hg.stats(reset=True)
t0 = time.perf_counter(); got = hg.search(queries, 10, ef, entry, lvl); gpu_s = time.perf_counter() - t0
hops, evals = hg.stats()
t0 = time.perf_counter(); want = g.search_batch(qp, 10, ef, threads=threads); cpu_s = time.perf_counter() - t0
t0 = time.perf_counter(); g.search_batch(qp[:256], 10, ef, threads=1); cpu1_s = (time.perf_counter() - t0) * nq / 256
same = sum(int(np.array_equal(a["score"], b["score"])) for a, b in zip(got, want))
exact = st.search_batch(queries[:200], 10)
rec = float(np.mean([np.mean(r["score"] >= e["score"][-1]) for r, e in zip(got[:200], exact)]))
print(json.dumps({"rows": n, "dim": dim, "queries": nq, "ef": ef, "build_s": build_s, "threads": threads, "gpu_qps_e2e": nq / gpu_s, "cpu_qps_all_threads": nq / cpu_s,
                  "cpu_qps_1_thread": nq / cpu1_s, "identical_lists": same, "recall_at_10": rec, "hops_per_query": hops / nq, "evals_per_query": evals / nq,
                  "gpu_ms_per_query_serial_equiv": gpu_s / nq * 1e3, "variants_qps_e2e": variants}))
