"""Multi-GPU brute-force search: one process per GPU, rows sharded by contiguous ranges.

The reference searches independent segments on separate blocking tasks and merges their top-k lists on the host
(SegmentsSearcher, lib/collection/src/collection_manager/segments_searcher.rs:212-345 -> BatchResultAggregator,
lib/shard/src/search_result_aggregator.rs:50-117).  Here every rank holds one shard in HBM, runs the fused scan
locally, and the ONLY exchange is `n_queries x top x 8 B` per rank: written by the library's exchange kernel straight into
every peer's mapped buffer over NVLink and merged there (qb_comm_* / qb_multi_search_batch*, qdrant_b200/csrc/qb_comm.cu).
torch.distributed only carries the 64-byte IPC handles at start-up; torch is otherwise used for device buffers and streams.
`exchange="nccl"` keeps the round-1 path (two NCCL all-gathers + qb_topk_merge_device) for comparison.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from ._capi import check, lib, vp
from .scorer import SCORED_POINT_OFFSET, _Storage


def shard_ranges(n_rows: int, world: int) -> list[tuple[int, int]]:
    """Contiguous row ranges per rank (whole-segment granularity in the reference); the first n_rows % world ranks
    get one extra row.  Range r is [begin, end); its id_base is begin."""
    base, extra = divmod(int(n_rows), int(world))
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < extra else 0)
        out.append((b, e))
        b = e
    return out


def merge_topk_host(lists, top: int) -> np.ndarray:
    """Host-side BatchResultAggregator for one query: k-way merge of per-shard descending lists, ordered like the
    device merge (score desc, id asc).  Used when results are already on the host (qb_search_batch per shard)."""
    allp = np.concatenate([np.asarray(l, dtype=SCORED_POINT_OFFSET) for l in lists]) if len(lists) else np.zeros(0, SCORED_POINT_OFFSET)
    if allp.size == 0:
        return allp
    order = np.lexsort((allp["idx"], -allp["score"].astype(np.float64)))
    return allp[order][:top].copy()


class ShardedSegmentSearcher:
    def __init__(self, storage: _Storage, id_base: int, top: int, max_queries: int, device: torch.device, exchange: str = "peer"):
        self.storage, self.top, self.max_queries, self.device = storage, int(top), int(max_queries), device
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        check(lib().qb_storage_set_id_base(storage._h, int(id_base)))
        self.comm = vp()
        self.exchange = exchange if self.world > 1 else "none"
        if self.exchange == "peer":
            # every step is agreed on by all ranks: if CUDA IPC is not available on this box (container policy), ALL ranks take the NCCL exchange
            ok, why = 1, ""
            h = vp()
            if lib().qb_comm_create(device.index, self.rank, self.world, self.max_queries, self.top, C.byref(h)) != 0:
                ok, why = 0, lib().qb_last_error().decode("utf-8", "replace")
            mine = (C.c_uint8 * 64)()
            if ok and lib().qb_comm_local_handle(h, mine) != 0:
                ok, why = 0, lib().qb_last_error().decode("utf-8", "replace")
            handles = [None] * self.world
            dist.all_gather_object(handles, (ok, bytes(mine)))      # 64 bytes per rank, once
            ok = int(all(x[0] for x in handles))
            if ok:
                blob = (C.c_uint8 * (64 * self.world)).from_buffer_copy(b"".join(x[1] for x in handles))
                if lib().qb_comm_connect(h, blob) != 0:
                    ok, why = 0, lib().qb_last_error().decode("utf-8", "replace")
            flags = [None] * self.world
            dist.all_gather_object(flags, ok)
            if all(flags):
                self.comm = h
            else:
                if h:
                    lib().qb_comm_destroy(h)
                self.exchange = "nccl"
                if self.rank == 0:
                    import sys
                    print(f"[qdrant_b200] peer-memory exchange unavailable ({why or 'a peer rank failed'}): using the NCCL all-gather exchange", file=sys.stderr)
            dist.barrier()
        self.stream = torch.cuda.ExternalStream(storage.stream_ptr(), device=device)
        # the pipelined exchange + merge kernels run on the communicator's own stream (qb_comm_stream)
        self.xstream = torch.cuda.ExternalStream(lib().qb_comm_stream(self.comm), device=device) if self.comm else None
        nq, k, w = self.max_queries, self.top, self.world
        # ScoredPointOffset = 8 bytes -> int64 tensors as opaque 8-byte records
        self.d_queries = torch.empty((nq, storage.dim), dtype=torch.float32, device=device)
        self.d_local = torch.empty((nq, k), dtype=torch.int64, device=device)
        self.d_local_cnt = torch.empty((nq,), dtype=torch.int32, device=device)
        self.d_all = torch.empty((w, nq, k), dtype=torch.int64, device=device)
        self.d_all_cnt = torch.empty((w, nq), dtype=torch.int32, device=device)
        self.d_out = torch.empty((nq, k), dtype=torch.int64, device=device)
        self.d_out_cnt = torch.empty((nq,), dtype=torch.int32, device=device)
        self.d_scratch = torch.empty((nq * w * k,), dtype=torch.int64, device=device)
        self.h_queries = torch.empty((nq, storage.dim), dtype=torch.float32).pin_memory()
        self.h_out = torch.empty((nq, k), dtype=torch.int64).pin_memory()
        self.h_out_cnt = torch.empty((nq,), dtype=torch.int32).pin_memory()

    # ---- device-resident step: queries already in self.d_queries[:nq]; results land in self.d_out / d_out_cnt
    def search_device(self, nq: int) -> None:
        if self.exchange == "peer":
            # local fused scan + select, then ONE kernel: push this shard's lists into every peer's buffer, wait for theirs, merge.
            # Pipelined (dev_local = NULL): the exchange + merge runs on self.xstream while the next call's scan already streams rows;
            # drain() orders self.stream after the merges.
            check(lib().qb_multi_search_batch_device(self.comm, self.storage._h, vp(self.d_queries.data_ptr()), nq, self.top, None, None,
                                                     vp(self.d_out.data_ptr()), vp(self.d_out_cnt.data_ptr())))
            return
        with torch.cuda.stream(self.stream):
            check(lib().qb_search_batch_device(self.storage._h, vp(self.d_queries.data_ptr()), nq, self.top,
                                               vp(self.d_local.data_ptr()), vp(self.d_local_cnt.data_ptr())))
            if self.world == 1:
                self.d_out[:nq].copy_(self.d_local[:nq], non_blocking=True)
                self.d_out_cnt[:nq].copy_(self.d_local_cnt[:nq], non_blocking=True)
                return
            # one collective per batch: gather every rank's local top-k (k x 8 B per query) over NVLink
            dist.all_gather_into_tensor(self.d_all.view(-1), self.d_local.view(-1))
            dist.all_gather_into_tensor(self.d_all_cnt.view(-1), self.d_local_cnt.view(-1))
            check(lib().qb_topk_merge_device(self.device.index, vp(self.d_all.data_ptr()), vp(self.d_all_cnt.data_ptr()), self.world,
                                             self.max_queries, self.top, vp(self.d_out.data_ptr()), vp(self.d_out_cnt.data_ptr()),
                                             vp(self.d_scratch.data_ptr()), self.d_scratch.numel() * 8, vp(self.stream.cuda_stream)))

    # ---- end-to-end call: host queries in, host results out (H2D + D2H inside)
    def search(self, queries: np.ndarray):
        q = np.atleast_2d(np.ascontiguousarray(queries, dtype=np.float32))
        nq = q.shape[0]
        assert nq <= self.max_queries
        if self.world == 1:
            return self.storage.search_batch(q, self.top)  # the plain C-ABI call (qb_search_batch)
        if self.exchange == "peer":
            # the C-ABI collective with HOST buffers: H2D of the queries, scan, exchange + merge, D2H of the merged lists, all inside
            from ._capi import ScoredPoint, f32p, u32p
            out = np.zeros((nq, self.top), dtype=SCORED_POINT_OFFSET)
            counts = np.zeros(nq, dtype=np.uint32)
            check(lib().qb_multi_search_batch(self.comm, self.storage._h, q.ctypes.data_as(f32p), nq, self.top, None, None,
                                              out.ctypes.data_as(C.POINTER(ScoredPoint)), counts.ctypes.data_as(u32p), None))
            return [out[i, : counts[i]].copy() for i in range(nq)]
        self.h_queries[:nq].copy_(torch.from_numpy(q))
        with torch.cuda.stream(self.stream):
            self.d_queries[:nq].copy_(self.h_queries[:nq], non_blocking=True)
        self.search_device(nq)
        with torch.cuda.stream(self.stream):
            self.h_out.copy_(self.d_out, non_blocking=True)
            self.h_out_cnt.copy_(self.d_out_cnt, non_blocking=True)
        self.stream.synchronize()
        rec = self.h_out.numpy().view(SCORED_POINT_OFFSET).reshape(self.max_queries, self.top)
        cnt = self.h_out_cnt.numpy()
        return [rec[i, : cnt[i]].copy() for i in range(nq)]

    def close(self):
        if self.comm:
            torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier()          # no rank unmaps its buffer while a peer may still write into it
            lib().qb_comm_destroy(self.comm)
            self.comm = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def drain(self) -> None:
        """Order self.stream after every exchange + merge enqueued so far (no host synchronisation)."""
        if self.xstream is not None:
            self.stream.wait_stream(self.xstream)

    def results_host(self, nq: int):
        """Copy the device results of the last search_device() to the host (synchronises)."""
        self.drain()
        if self.comm:
            check(lib().qb_comm_check(self.comm))      # a step whose exchange timed out must not read as an empty result
        with torch.cuda.stream(self.stream):
            self.h_out.copy_(self.d_out, non_blocking=True)
            self.h_out_cnt.copy_(self.d_out_cnt, non_blocking=True)
        self.stream.synchronize()
        rec = self.h_out.numpy().view(SCORED_POINT_OFFSET).reshape(self.max_queries, self.top)
        cnt = self.h_out_cnt.numpy()
        return [rec[i, : cnt[i]].copy() for i in range(nq)]
