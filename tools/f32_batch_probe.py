#!/usr/bin/env python
"""Batched dense-f32 search probe (J1): N x 768 cosine, B-query batches through qb_search_batch_device; tensor-core prefilter + exact
rescoring vs the CUDA-core multi-query path.    python tools/f32_batch_probe.py [rows] [batch]"""
import json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qdrant_b200 import scorer as qb
from qdrant_b200._capi import check, lib, vp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dim, top = 768, 10
dev = torch.device("cuda", 0)
st = qb.DenseVectorStorage(None, qb.Distance.Cosine, count=n, dim=dim)
g = torch.Generator(device=dev); g.manual_seed(42)
for r0 in range(0, n, 500_000):
    cn = min(500_000, n - r0)
    x = torch.randn((cn, dim), generator=g, device=dev)
    check(lib().qb_metric_preprocess_device(0, int(qb.Distance.Cosine), dim, cn, vp(x.data_ptr()), dim * 4))
    st.write_rows_device(r0, cn, x.data_ptr(), dim * 4)
    del x
queries = np.random.default_rng(43).standard_normal((nq, dim)).astype(np.float32)
d_q = torch.from_numpy(queries).to(dev)
d_out = torch.empty((nq, top), dtype=torch.int64, device=dev); d_cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
stream = torch.cuda.ExternalStream(st.stream_ptr(), device=dev)


def run(k, q):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(k):
        check(lib().qb_search_batch_device(st._h, vp(d_q.data_ptr()), q, top, vp(d_out.data_ptr()), vp(d_cnt.data_ptr())))
    ev1.record(stream)
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / k


run(2, nq)
st.search_stats(reset=True)
st.profile(True)
ms = run(5, nq)
n_prof, prof_ms = st.profile_read(reset=True)
st.profile(False)
searches, reruns = st.search_stats(reset=True)
res = d_out.cpu().numpy().view(qb.SCORED_POINT_OFFSET).reshape(nq, top)[:64].copy()
qb.set_option("disable_mma", 1)
run(1, 64)
ms_cc = run(2, 64)
res_cc = d_out.cpu().numpy().view(qb.SCORED_POINT_OFFSET).reshape(nq, top)[:64].copy()
qb.set_option("disable_mma", 0)
print(json.dumps({"rows": n, "dim": dim, "batch": nq, "ms_per_batch": ms, "qps": nq / ms * 1e3, "main_pass_ms": prof_ms / max(n_prof, 1), "reruns": reruns,
                  "tflops_main_pass": 2.0 * nq * n * dim / (prof_ms / max(n_prof, 1) / 1e3) / 1e12 if n_prof else None,
                  "cuda_core_ms_per_64_queries": ms_cc, "cuda_core_qps": 64 / ms_cc * 1e3, "identical_to_cuda_core_first_64": bool(np.array_equal(res, res_cc))}))
