#!/usr/bin/env python
"""prefilter_probe.py [rows] [dim] — single-query searches over one synthetic cosine storage (generated on the device), device-timed, for several
ring-slot sizes / producer-warp counts of the shadow-plane filter kernels and both planes; prints one JSON line.  Results of every variant are compared with the exact scan."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from qdrant_b200 import scorer as qb
from qdrant_b200._capi import check, lib, vp

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
dev = torch.device("cuda", 0)
st = qb.DenseVectorStorage(None, qb.Distance.Cosine, count=rows, dim=dim, device=0)
gen = torch.Generator(device=dev); gen.manual_seed(42)
for r0 in range(0, rows, 500_000):
    n = min(500_000, rows - r0)
    x = torch.randn((n, dim), generator=gen, device=dev, dtype=torch.float32)
    check(lib().qb_metric_preprocess_device(0, int(qb.Distance.Cosine), dim, n, vp(x.data_ptr()), dim * 4))
    st.write_rows_device(r0, n, x.data_ptr(), dim * 4)
    del x
torch.cuda.synchronize()
queries = np.random.default_rng(43).standard_normal((16, dim)).astype(np.float32)
d_q = torch.from_numpy(queries).to(dev)
d_out = torch.empty((1, 10), dtype=torch.int64, device=dev); d_cnt = torch.empty((1,), dtype=torch.int32, device=dev)
stream = torch.cuda.ExternalStream(st.stream_ptr(), device=dev)


def run(K=30):
    def step(i):
        check(lib().qb_search_batch_device(st._h, vp(d_q[i % 16].data_ptr()), 1, 10, vp(d_out.data_ptr()), vp(d_cnt.data_ptr())))
    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    st.profile_read(reset=True); st.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(K):
        step(i)
    e1.record(stream)
    torch.cuda.synchronize()
    n_prof, prof_ms = st.profile_read(reset=True); st.profile(False)
    return {"ms_per_step": e0.elapsed_time(e1) / K, "qps": K / (e0.elapsed_time(e1) / 1e3), "kernel_ms": prof_ms / max(n_prof, 1)}


qb.set_option("disable_prefilter", 1)
exact = [st.search_batch(queries[i], 10)[0] for i in range(4)]
out = {"rows": rows, "dim": dim, "exact_f32_scan": run(10)}
qb.set_option("disable_prefilter", 0)
for plane, name in ((0, "int8"), (1, "bf16")):
    qb.set_option("prefilter_plane", plane)
    for prod in (1, 2, 4):
        qb.set_option("prefilter_producers", prod)
        for slot in (12288, 8192, 6144):
            qb.set_option("prefilter_slot_bytes", slot)
            got = [st.search_batch(queries[i], 10)[0] for i in range(4)]
            same = all(np.array_equal(a["idx"], b["idx"]) and np.array_equal(a["score"].view(np.uint32), b["score"].view(np.uint32)) for a, b in zip(got, exact))
            r = run()
            r["identical_to_exact_scan"] = bool(same)
            out[f"{name}_producers{prod}_slot{slot}"] = r
qb.set_option("prefilter_producers", 0)
qb.set_option("prefilter_slot_bytes", 0); qb.set_option("prefilter_plane", 0)
out["fallbacks"] = int(st.search_stats()[1])
print(json.dumps(out))
