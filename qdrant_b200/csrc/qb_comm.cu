// qb_comm.cu — the cross-GPU step of a sharded search: exchange of per-shard top-k lists over NVLink peer memory + merge.
//
// Replaces BatchResultAggregator (lib/shard/src/search_result_aggregator.rs:50-117), which the reference runs on the host over
// the per-segment lists SegmentsSearcher collected (lib/collection/src/collection_manager/segments_searcher.rs:212-345), for
// segments that live on different GPUs of one box.  Rows shard naturally; the only data that crosses GPUs is
// `n_queries x top x 8 B` per shard.  Round 1 moved it with two NCCL all-gathers launched from Python (~45 us of launch latency
// on an 84-byte payload, 0.865 scaling at 8 GPUs).  Here the exchange is part of the scoring tail:
//
//   every rank owns an exchange buffer in HBM that all peers map (cudaIpc handles between processes, direct peer access inside
//   one process).  After the local fused scan + select, ONE kernel (one CTA per query)
//     1. stores this shard's list for its query straight into slot [parity][my rank][query] of EVERY peer's buffer (remote
//        stores through NVLink / NVSwitch), __threadfence_system(), then raises flag [parity][my rank][query] = seq on the peer;
//     2. spins (bounded) until its OWN buffer shows flag == seq from every rank for that query, merges the `world` sorted lists
//        (<= world x top keys: one bitonic sort in shared memory, keys = (score desc, id asc) as everywhere) and writes the result.
//   No NCCL call, no host round trip: the transfer overlaps the other queries' merges, and a step is preprocess + scan + this kernel.
//   `seq` increases by one per call and all ranks must call in the same order (it is a collective, like the aggregator it replaces
//   sees every segment's list).
//
// Waiting for the slowest of N GPUs once per query is what limits strong scaling of a 0.6 ms scan (max-of-8 of the scan times, not
// their mean, sets the step).  The device-resident entry point therefore PIPELINES independent steps: the exchange + merge kernel of step
// i runs on the communicator's own high-priority stream (it needs one small CTA and co-resides with the scan), while the scan of step
// i + 1 starts at once on the storage's stream.  Flow control: a rank starts scan(i) only after its own merge(i - 2) has completed (window
// W = 2), so at most two steps' lists wait in the shard's local ring (four buffers: a scan never overwrites lists that were not pushed).
// Remote slots (a ring of four, two would do): the exchange streams are in order — X pushes step i only after its merge(i - 1), which saw
// Y's push(i - 1), which Y issued after its merge(i - 2) — so a push never overwrites a slot its owner has not merged
// (tests/test_pipeline_protocol.py model-checks both hazards under random schedules).  Results of step i are complete when the
// communicator's stream has run its merge (qb_comm_stream).
#include <algorithm>
#include <functional>

#include "qb_internal.h"

namespace {

constexpr int XCHG_THREADS = 256;
constexpr uint32_t XCHG_MAX_KEYS = 4096;   // world * top keys merged per query in shared memory

struct XchgBuf {            // layout of one rank's exchange buffer (all offsets in bytes from the base)
    uint64_t flags_off;     // u32 [QB_XCHG_SLOTS][world][max_q]
    uint64_t counts_off;    // u32 [QB_XCHG_SLOTS][world][max_q]
    uint64_t lists_off;     // qb_scored_point [QB_XCHG_SLOTS][world][max_q][max_top]
    uint64_t total;
};
__host__ __device__ inline XchgBuf xchg_layout(uint32_t world, uint32_t max_q, uint32_t max_top) {
    XchgBuf b;
    const uint64_t nf = (uint64_t)QB_XCHG_SLOTS * world * max_q;
    b.flags_off = 0;
    b.counts_off = (nf * 4 + 255) & ~255ull;
    b.lists_off = (b.counts_off + nf * 4 + 255) & ~255ull;
    b.total = b.lists_off + nf * max_top * 8;
    return b;
}

struct XchgParams {
    uint8_t* peers[QB_MAX_WORLD];   // exchange buffers of all ranks (peers[rank] = own)
    uint32_t world, rank, max_q, max_top;
    uint32_t nq, top, seq, parity;
    const qb_scored_point* local; const uint32_t* local_cnt;   // this shard's lists [nq][top]
    qb_scored_point* out; uint32_t* out_cnt;
    unsigned int* error;            // device word: 1 = a peer never arrived (bounded spin expired)
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// The grid is capped at what is co-resident (the launcher sizes it), and every CTA pushes ALL its queries before it waits for any:
// a wait only depends on peers' pushes, pushes depend on nothing, so no schedule of CTAs on either side can deadlock.
__global__ void __launch_bounds__(XCHG_THREADS) xchg_merge_kernel(const XchgParams p) {
    extern __shared__ unsigned long long keys[];     // pow2 >= world * top keys: small, so the CTA co-resides with a running scan
    __shared__ unsigned int s_timeout, s_valid;
    const XchgBuf L = xchg_layout(p.world, p.max_q, p.max_top);
    // ---- 1. push my lists to every rank (own buffer included: one code path)
    for (uint32_t q = blockIdx.x; q < p.nq; q += gridDim.x) {
        const uint64_t slot = ((uint64_t)p.parity * p.world + p.rank) * p.max_q + q;   // [parity][my rank][q] in every peer's buffer
        const uint32_t cnt = min(p.local_cnt[q], p.top);
        for (uint32_t r = 0; r < p.world; ++r) {
            uint8_t* base = p.peers[(p.rank + 1 + r) % p.world];    // start with the next rank: spreads the traffic over the switch
            qb_scored_point* dst = reinterpret_cast<qb_scored_point*>(base + L.lists_off) + slot * p.max_top;
            for (uint32_t i = threadIdx.x; i < cnt; i += XCHG_THREADS) dst[i] = p.local[(uint64_t)q * p.top + i];
            if (threadIdx.x == 0) reinterpret_cast<uint32_t*>(base + L.counts_off)[slot] = cnt;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        s_timeout = 0u;
        __threadfence_system();    // the block's stores (ordered before this thread by the barrier) become visible system-wide before the flags
        for (uint32_t q = blockIdx.x; q < p.nq; q += gridDim.x) {
            const uint64_t slot = ((uint64_t)p.parity * p.world + p.rank) * p.max_q + q;
            for (uint32_t r = 0; r < p.world; ++r) st_release_sys(reinterpret_cast<uint32_t*>(p.peers[r] + L.flags_off) + slot, p.seq);
        }
    }
    __syncthreads();
    uint8_t* mine = p.peers[p.rank];
    for (uint32_t q = blockIdx.x; q < p.nq; q += gridDim.x) {
    // ---- 2. wait for every rank's list of query q in MY buffer
    if (threadIdx.x < p.world) {
        const uint32_t* f = reinterpret_cast<const uint32_t*>(mine + L.flags_off) + ((uint64_t)p.parity * p.world + threadIdx.x) * p.max_q + q;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) != p.seq) {
            __nanosleep(64);
            if (clock64() - t0 > 20000000000ll) { s_timeout = 1u; break; }   // ~10 s: a rank that never calls must not hang the GPU
        }
    }
    __syncthreads();
    if (s_timeout) { if (threadIdx.x == 0) atomicOr(p.error, 1u); return; }   // uniform: every thread reads the same shared word
    // ---- 3. merge: world sorted lists -> top
    const uint32_t per = p.top, total = p.world * per;
    uint32_t p2 = 32;
    while (p2 < total) p2 <<= 1;
    for (uint32_t i = threadIdx.x; i < p2; i += XCHG_THREADS) {
        unsigned long long k = 0ull;
        if (i < total) {
            const uint32_t r = i / per, j = i % per;
            const uint64_t s = ((uint64_t)p.parity * p.world + r) * p.max_q + q;
            const uint32_t c = __ldcg(reinterpret_cast<const uint32_t*>(mine + L.counts_off) + s);   // written by a peer over NVLink: read at L2, never a stale L1 line
            if (j < c) {
                const unsigned long long raw = __ldcg(reinterpret_cast<const unsigned long long*>(mine + L.lists_off) + s * p.max_top + j);
                k = qb_pack_key(__uint_as_float((uint32_t)(raw >> 32)), (uint32_t)raw);   // {idx, score} little-endian
            }
        }
        keys[i] = k;
    }
    __syncthreads();
    for (uint32_t k = 2; k <= p2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < p2; i += XCHG_THREADS) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool desc = ((i & k) == 0);
                    if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    unsigned int valid = 0;
    for (uint32_t i = threadIdx.x; i < p.top; i += XCHG_THREADS) {
        const unsigned long long k = keys[i];
        qb_scored_point sp;
        if (k != 0ull) { sp.idx = qb_key_id(k); sp.score = qb_key_score(k); valid++; }
        else { sp.idx = 0; sp.score = 0.0f; }
        p.out[(uint64_t)q * p.top + i] = sp;
    }
    if (threadIdx.x == 0) s_valid = 0u;
    __syncthreads();
    if (valid) atomicAdd(&s_valid, valid);
    __syncthreads();
    if (threadIdx.x == 0) p.out_cnt[q] = s_valid;
    __syncthreads();
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host side
extern "C" qb_status qb_comm_create(int32_t device, int32_t rank, int32_t world, uint32_t max_queries, uint32_t max_top, qb_comm** out) {
    QB_CHECK(out, QB_ERR_INVALID, "comm_create: null out");
    *out = nullptr;
    QB_CHECK(world >= 1 && world <= (int)QB_MAX_WORLD && rank >= 0 && rank < world, QB_ERR_INVALID, "comm_create: rank %d / world %d (max %u)", rank, world, QB_MAX_WORLD);
    QB_CHECK(max_queries >= 1 && max_top >= 1 && (uint64_t)world * max_top <= XCHG_MAX_KEYS, QB_ERR_INVALID,
             "comm_create: world x max_top = %llu exceeds the %u keys merged per query", (unsigned long long)world * max_top, XCHG_MAX_KEYS);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); qb_set_error("no CUDA device (this library has no CPU fallback)"); return QB_ERR_NO_DEVICE; }
    QB_CHECK(device >= 0 && device < n, QB_ERR_INVALID, "comm_create: device %d out of range", device);
    QB_CUDA(cudaSetDevice(device));
    qb_comm* c = new qb_comm();
    c->device = device; c->rank = rank; c->world = world; c->max_q = max_queries; c->max_top = max_top;
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    }
    const XchgBuf L = xchg_layout((uint32_t)world, max_queries, max_top);
    c->bytes = L.total;
    if (cudaMalloc(&c->d_buf, L.total) != cudaSuccess || cudaMalloc(&c->d_error, 256) != cudaSuccess) {
        qb_set_error("comm_create: cudaMalloc(%llu) failed: %s", (unsigned long long)L.total, cudaGetErrorString(cudaGetLastError()));
        qb_comm_destroy(c);
        return QB_ERR_OOM;
    }
    cudaMemset(c->d_buf, 0, L.total);
    cudaMemset(c->d_error, 0, 256);
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        bool ok = cudaStreamCreateWithPriority(&c->xstream, cudaStreamNonBlocking, hi) == cudaSuccess;
        for (uint32_t i = 0; i < QB_XCHG_SLOTS && ok; ++i) {
            ok = cudaEventCreateWithFlags(&c->ev_scan[i], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&c->ev_merge[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaMalloc(&c->d_ring[i], (size_t)max_queries * max_top * sizeof(qb_scored_point) + 256) == cudaSuccess &&
                 cudaMalloc(&c->d_ring_cnt[i], (size_t)max_queries * 4 + 256) == cudaSuccess;
        }
        if (!ok) { qb_set_error("comm_create: stream / event / ring allocation failed: %s", cudaGetErrorString(cudaGetLastError())); qb_comm_destroy(c); return QB_ERR_CUDA; }
    }
    cudaDeviceSynchronize();
    c->peers[rank] = reinterpret_cast<uint8_t*>(c->d_buf);
    c->connected = (world == 1);
    *out = c;
    return QB_OK;
}

extern "C" qb_status qb_comm_local_handle(qb_comm* c, uint8_t* handle_out) {
    QB_CHECK(c && handle_out, QB_ERR_INVALID, "comm_local_handle: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    QB_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    QB_CUDA(cudaIpcGetMemHandle(&h, c->d_buf));
    memcpy(handle_out, &h, 64);
    return QB_OK;
}

extern "C" qb_status qb_comm_connect(qb_comm* c, const uint8_t* handles) {
    QB_CHECK(c && handles, QB_ERR_INVALID, "comm_connect: null argument");
    QB_CUDA(cudaSetDevice(c->device));
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void* p = nullptr;
        QB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peers[r] = reinterpret_cast<uint8_t*>(p);
        c->ipc_opened[r] = true;
    }
    c->connected = true;
    return QB_OK;
}

extern "C" qb_status qb_comm_connect_local(qb_comm* const* comms, int32_t n) {
    QB_CHECK(comms && n >= 1, QB_ERR_INVALID, "comm_connect_local: bad arguments");
    for (int i = 0; i < n; ++i) QB_CHECK(comms[i] && comms[i]->world == n && comms[i]->rank == i, QB_ERR_INVALID, "comm_connect_local: comms[%d] is not rank %d of %d", i, i, n);
    for (int i = 0; i < n; ++i) {
        QB_CUDA(cudaSetDevice(comms[i]->device));
        for (int j = 0; j < n; ++j) {
            if (i == j) continue;
            if (comms[i]->device != comms[j]->device) {
                int can = 0;
                QB_CUDA(cudaDeviceCanAccessPeer(&can, comms[i]->device, comms[j]->device));
                QB_CHECK(can, QB_ERR_UNSUPPORTED, "comm_connect_local: device %d cannot access device %d", comms[i]->device, comms[j]->device);
                cudaError_t e = cudaDeviceEnablePeerAccess(comms[j]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { qb_set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); return QB_ERR_CUDA; }
                cudaGetLastError();
            }
            comms[i]->peers[j] = reinterpret_cast<uint8_t*>(comms[j]->d_buf);
        }
        comms[i]->connected = true;
    }
    return QB_OK;
}

extern "C" void qb_comm_destroy(qb_comm* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r) if (c->ipc_opened[r] && c->peers[r]) cudaIpcCloseMemHandle(c->peers[r]);
    cudaFree(c->d_buf); cudaFree(c->d_error); cudaFree(c->d_local); cudaFree(c->d_local_cnt);
    for (uint32_t i = 0; i < QB_XCHG_SLOTS; ++i) {
        cudaFree(c->d_ring[i]); cudaFree(c->d_ring_cnt[i]);
        if (c->ev_scan[i]) cudaEventDestroy(c->ev_scan[i]);
        if (c->ev_merge[i]) cudaEventDestroy(c->ev_merge[i]);
    }
    if (c->xstream) cudaStreamDestroy(c->xstream);
    cudaGetLastError();
    delete c;
}

// enqueue exchange + merge of this shard's device lists on `stream`; every rank calls with the same nq / top, in the same order
qb_status qb_comm_exchange_merge(qb_comm* c, const qb_scored_point* d_local, const uint32_t* d_local_cnt, uint32_t nq, uint32_t top, qb_scored_point* d_out,
                                 uint32_t* d_out_cnt, cudaStream_t stream) {
    QB_CHECK(c->connected, QB_ERR_INVALID, "multi_search: communicator not connected (qb_comm_connect / qb_comm_connect_local)");
    QB_CHECK(nq <= c->max_q && top <= c->max_top, QB_ERR_INVALID, "multi_search: %u queries x top %u exceed the communicator's %u x %u", nq, top, c->max_q, c->max_top);
    if (nq == 0) return QB_OK;
    XchgParams p{};
    for (int r = 0; r < c->world; ++r) p.peers[r] = c->peers[r];
    p.world = (uint32_t)c->world; p.rank = (uint32_t)c->rank; p.max_q = c->max_q; p.max_top = c->max_top;
    p.nq = nq; p.top = top;
    c->seq += 1;
    p.seq = c->seq; p.parity = c->seq % QB_XCHG_SLOTS;
    p.local = d_local; p.local_cnt = d_local_cnt; p.out = d_out; p.out_cnt = d_out_cnt; p.error = c->d_error;
    const unsigned grid = std::min<unsigned>(nq, (unsigned)c->sm_count);   // one CTA per SM at most: all co-resident
    uint32_t p2 = 32;
    while (p2 < (uint32_t)c->world * top) p2 <<= 1;
    xchg_merge_kernel<<<grid, XCHG_THREADS, (size_t)p2 * 8, stream>>>(p);
    QB_LAUNCHED();
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

extern "C" void* qb_comm_stream(qb_comm* c) { return c ? c->xstream : nullptr; }

// Drain the communicator's stream and report what the asynchronous (device-resident) steps could not: an exchange that gave up
// waiting for a peer (a rank that never made the matching call) leaves its results empty and sets the error word.
extern "C" qb_status qb_comm_check(qb_comm* c) {
    QB_CHECK(c, QB_ERR_INVALID, "comm_check: null communicator");
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != c->device) QB_CUDA(cudaSetDevice(c->device));
    QB_CUDA(cudaStreamSynchronize(c->xstream));
    unsigned int err = 0;
    QB_CUDA(cudaMemcpy(&err, c->d_error, 4, cudaMemcpyDeviceToHost));
    if (err) {
        cudaMemset(c->d_error, 0, 4);
        qb_set_error("sharded search: rank %d timed out waiting for a peer's lists (flags 0x%x); every rank must make the same sequence of calls", c->rank, err);
        return QB_ERR_CUDA;
    }
    return QB_OK;
}

// Pipelined collective for device-resident queries / results (see the header comment): scan on `scan_stream`, exchange + merge on the
// communicator's stream.  `launch_scan(d_local, d_local_cnt)` enqueues this shard's fused scan into the given list buffers.
qb_status qb_comm_pipelined_step(qb_comm* c, cudaStream_t scan_stream, uint32_t nq, uint32_t top, qb_scored_point* d_out, uint32_t* d_out_cnt,
                                 const std::function<qb_status(qb_scored_point*, uint32_t*)>& launch_scan) {
    QB_CHECK(nq <= c->max_q && top <= c->max_top, QB_ERR_INVALID, "multi_search: %u queries x top %u exceed the communicator's %u x %u", nq, top, c->max_q, c->max_top);
    const uint32_t slot = (c->seq + 1) % QB_XCHG_SLOTS;
    if (c->seq >= 2) QB_CUDA(cudaStreamWaitEvent(scan_stream, c->ev_merge[(c->seq + 1 - 2) % QB_XCHG_SLOTS], 0));   // window: scan(i) after own merge(i - 2)
    QB_TRY(launch_scan(c->d_ring[slot], c->d_ring_cnt[slot]));
    QB_CUDA(cudaEventRecord(c->ev_scan[slot], scan_stream));
    QB_CUDA(cudaStreamWaitEvent(c->xstream, c->ev_scan[slot], 0));
    QB_TRY(qb_comm_exchange_merge(c, c->d_ring[slot], c->d_ring_cnt[slot], nq, top, d_out, d_out_cnt, c->xstream));   // seq += 1 inside
    QB_CUDA(cudaEventRecord(c->ev_merge[slot], c->xstream));
    return QB_OK;
}
