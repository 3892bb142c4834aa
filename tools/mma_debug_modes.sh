# Timing experiments on the batched SQ8 tensor-core kernel (qb_sq8_mma.cu): kernel time of the main pass over 4M x 768 codes x 1024
# queries with QB_MMA_DEBUG = 0 (normal), 1 (epilogue skipped), 4 (prefilter never passes), for cta_group::1 and ::2.
# Results with a non-zero debug mode are INVALID as search results; this is a profiling aid (see profiles/README_r01.md).
mkdir -p gpurun_out
for mode in 0 1 4; do for cta in 1 2; do
  if [ $cta = 1 ]; then export QB_MMA_1CTA=1; else unset QB_MMA_1CTA; fi
  QB_MMA_DEBUG=$mode timeout 300 python - <<'PY'
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
dq = torch.from_numpy(q).to(dev)
out = torch.empty((nq, 10), dtype=torch.int64, device=dev); cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
for _ in range(2):
    check(lib().qb_search_batch_device(st._h, vp(dq.data_ptr()), nq, 10, vp(out.data_ptr()), vp(cnt.data_ptr())))
torch.cuda.synchronize()
st.profile(True)
for _ in range(5):
    check(lib().qb_search_batch_device(st._h, vp(dq.data_ptr()), nq, 10, vp(out.data_ptr()), vp(cnt.data_ptr())))
torch.cuda.synchronize()
k, ms = st.profile_read()
print(json.dumps({"debug": os.environ.get("QB_MMA_DEBUG"), "one_cta": os.environ.get("QB_MMA_1CTA"), "kernel_ms_4M_rows": ms / k}))
PY
done; done
