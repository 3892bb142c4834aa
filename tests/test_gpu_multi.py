"""Multi-GPU (>= 2 devices, NCCL): the row-sharded search (one process per GPU, all-gather of local top-k, device merge)
returns exactly what one GPU returns over the whole data set.  Skipped on single-GPU boxes."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["QB_ROOT"])
from qdrant_b200 import scorer as qb
from qdrant_b200.sharded import ShardedSegmentSearcher, shard_ranges
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
rng = np.random.default_rng(3)
base = rng.standard_normal((200_003, 128)).astype(np.float32)
queries = rng.standard_normal((5, 128)).astype(np.float32)
b, e = shard_ranges(base.shape[0], world)[rank]
st = qb.DenseVectorStorage(base[b:e], qb.Distance.Dot, device=lr)
s = ShardedSegmentSearcher(st, id_base=b, top=10, max_queries=5, device=dev)
res = s.search(queries)
if rank == 0:
    np.save(os.environ["QB_OUT"], np.stack(res))
del s; torch.cuda.synchronize(); st.close()
dist.barrier(); dist.destroy_process_group()
'''


def test_sharded_equals_single(tmp_path, oracle):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from qdrant_b200 import scorer as qb

    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = tmp_path / "res.npy"
    env = dict(os.environ, QB_ROOT=ROOT, QB_OUT=str(out))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    got = np.load(out)
    rng = np.random.default_rng(3)
    base = rng.standard_normal((200_003, 128)).astype(np.float32)
    queries = rng.standard_normal((5, 128)).astype(np.float32)
    st = qb.DenseVectorStorage(base, qb.Distance.Dot)
    want = st.search_batch(queries, 10)
    for i in range(5):
        np.testing.assert_array_equal(got[i], want[i])
    st.close()
