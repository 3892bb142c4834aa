"""N>1 host logic on CPU: world_size-2 gloo run of the shard plan + top-k exchange + merge.

No GPU here, so each rank scores its shard with the CPU oracle (test infrastructure) — what is under test is the
partitioning (shard_ranges / id_base), the all-gather protocol and merge_topk_host, against a single-process scan."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import assert_topk_equal


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    from oracle import oracle as o
    from qdrant_b200.sharded import merge_topk_host, shard_ranges
    from qdrant_b200.scorer import SCORED_POINT_OFFSET

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(42)
    base = rng.standard_normal((5001, 48)).astype(np.float32)
    queries = rng.standard_normal((3, 48)).astype(np.float32)
    top = 10
    b, e = shard_ranges(base.shape[0], world)[rank]
    local = o.scan_f32(o.DOT, base[b:e], queries, top)
    rec = np.zeros((len(queries), top), dtype=SCORED_POINT_OFFSET)
    cnt = np.zeros(len(queries), dtype=np.int32)
    for i, l in enumerate(local):
        l = l.copy()
        l["idx"] += b  # id_base
        rec[i, : l.size] = l
        cnt[i] = l.size
    t = torch.from_numpy(rec.view(np.int64).reshape(-1).copy())
    c = torch.from_numpy(cnt)
    all_t = torch.empty(world * t.numel(), dtype=torch.int64)
    all_c = torch.empty(world * c.numel(), dtype=torch.int32)
    dist.all_gather_into_tensor(all_t, t)
    dist.all_gather_into_tensor(all_c, c)
    all_rec = all_t.numpy().view(SCORED_POINT_OFFSET).reshape(world, len(queries), top)
    all_cnt = all_c.numpy().reshape(world, len(queries))
    merged = [merge_topk_host([all_rec[r, i, : all_cnt[r, i]] for r in range(world)], top) for i in range(len(queries))]
    np.save(os.path.join(out_dir, f"merged_{rank}.npy"), np.stack(merged))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_search_matches_single(tmp_path, oracle):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(42)
    base = rng.standard_normal((5001, 48)).astype(np.float32)
    queries = rng.standard_normal((3, 48)).astype(np.float32)
    want = oracle.scan_f32(oracle.DOT, base, queries, 10)
    m0 = np.load(tmp_path / "merged_0.npy")
    m1 = np.load(tmp_path / "merged_1.npy")
    np.testing.assert_array_equal(m0, m1)  # every rank ends with the same merged result
    for i in range(3):
        assert_topk_equal(m0[i], want[i])


def test_shard_ranges():
    from qdrant_b200.sharded import shard_ranges

    assert shard_ranges(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_ranges(0, 2) == [(0, 0), (0, 0)]
    r = shard_ranges(10_000_000, 8)
    assert r[0] == (0, 1_250_000) and r[-1][1] == 10_000_000
