import os, sys, json, time
sys.path.insert(0, '.')
import numpy as np, torch
from qdrant_b200 import scorer as qb
from qdrant_b200._capi import check, lib, vp
dev = torch.device('cuda', 0)
n, dim, nq = 4_000_000, 768, 1024
g = torch.Generator(device=dev); g.manual_seed(1)
rows = torch.empty((n, 772), dtype=torch.uint8, device=dev)
codes = torch.clamp(torch.randn((n, dim), generator=g, device=dev) * 12 + 64, 0, 127).to(torch.uint8)
rows[:, 4:] = codes; rows[:, :4] = 0
del codes
st = qb.ScalarQuantizedVectors(None, dim, 0.0035, -0.22, 0.0035 * 0.0035, qb.Distance.Cosine, rows_ptr=rows.data_ptr(), count=n)
del rows
q = np.random.default_rng(0).standard_normal((nq, dim)).astype(np.float32)
for i in range(3):
    t0 = time.perf_counter(); r = st.search_batch(q, 10); print("e2e ms", (time.perf_counter() - t0) * 1e3, file=sys.stderr)
st.close()
