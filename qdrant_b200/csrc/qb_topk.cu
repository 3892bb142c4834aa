// qb_topk.cu — per-query top-k selection over candidate key lists.
//
// Replaces the reference's FixedLengthPriorityQueue<ScoredPointOffset> (binary min-heap of size `top`,
// lib/common/common/src/fixed_length_priority_queue.rs:20-65) + into_sorted_vec (descending).  A heap is a
// serial structure; on the GPU the scan kernels emit 64-bit keys (score-major, id-minor, qb_common.cuh) and
// this file selects the k largest per query:
//     n <= 4096 : load into shared memory, bitonic sort, take the first k
//     n  > 4096 : 8-pass MSB radix select for the exact k-th key, gather keys >= it, bitonic sort those
// Keys are unique (ids are), so the result is deterministic: (score desc, id asc).
#include "qb_internal.h"

namespace {

constexpr int SORT_CAP = 4096;

__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* buf, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = buf[i], b = buf[ixj];
                    bool desc = ((i & k) == 0);
                    if (desc ? (a < b) : (a > b)) { buf[i] = b; buf[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(QB_SELECT_THREADS)
qb_select_kernel(const unsigned long long* cand, const unsigned int* __restrict__ cnt, unsigned long long cap,
                 unsigned long long fixed_n, uint32_t top, int mode, qb_scored_point* __restrict__ out,
                 uint32_t* __restrict__ out_counts, float* __restrict__ thr, unsigned int* __restrict__ overflow) {
    __shared__ unsigned long long buf[SORT_CAP];
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_kk, s_fill, s_short, s_done;

    const uint32_t q = blockIdx.x;
    unsigned long long n;
    if (fixed_n) {
        n = fixed_n;
    } else {
        unsigned int c = cnt[q];
        if (c > cap) {
            if (threadIdx.x == 0 && overflow) atomicOr(overflow, 1u);
            c = (unsigned int)cap;
        }
        n = c;
    }
    const unsigned long long* keys = cand + (unsigned long long)q * cap;
    int m = 0;  // number of keys staged in buf
    const bool select_only = (mode == 1);  // threshold mode never needs the sorted order: radix-select the staged keys instead of sorting them

    bool staged = false;
    if (n <= SORT_CAP) {
        for (int i = threadIdx.x; i < SORT_CAP; i += blockDim.x) buf[i] = (i < (int)n) ? keys[i] : 0ull;
        m = (int)n;
        staged = true;
        __syncthreads();
    } else if (fixed_n) {
        // sparse fixed-length list (per-CTA segments with empty slots): compact the non-empty keys; if they fit, sort them directly
        if (threadIdx.x == 0) s_fill = 0u;
        for (int i = threadIdx.x; i < SORT_CAP; i += blockDim.x) buf[i] = 0ull;
        __syncthreads();
        for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned long long key = keys[i];
            if (key != 0ull) {
                const unsigned int p = atomicAdd(&s_fill, 1u);
                if (p < SORT_CAP) buf[p] = key;
            }
        }
        __syncthreads();
        if (s_fill <= (unsigned int)SORT_CAP) { m = (int)s_fill; staged = true; }
        __syncthreads();
    }
    // staged keys: radix-select in shared memory instead of sorting all of them when only the threshold is wanted, or when the
    // wanted prefix is short (sorting 4096 keys to keep 10 costs 78 bitonic steps; selecting costs <= 8 histogram passes)
    const bool in_smem = staged && (select_only || (m >= 256 && (int)top * 4 <= m));
    if (in_smem) { keys = buf; n = (unsigned long long)m; }
    if (!staged || in_smem) {
        if (threadIdx.x == 0) { s_prefix = 0ull; s_kk = top; s_short = 0u; s_done = 0u; }
        unsigned long long mask = 0ull;
        __syncthreads();
        // threshold mode only needs the k-th SCORE (high 32 bits of the key): 4 passes instead of 8
        const int n_pass = (mode == 1) ? 4 : 8;
        for (int pass = 0; pass < n_pass; ++pass) {
            const int shift = 56 - 8 * pass;
            for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            for (unsigned long long i0 = 0; i0 < n; i0 += blockDim.x) {
                unsigned long long i = i0 + threadIdx.x;
                bool valid = false;
                unsigned int digit = 0xFFFFFFFFu;
                if (i < n) {
                    unsigned long long key = keys[i];
                    if ((key & mask) == prefix) { valid = true; digit = (unsigned int)((key >> shift) & 255ull); }
                }
                // warp-aggregated histogram update: keys share high bytes, so plain atomics would serialise
                unsigned int peers = __match_any_sync(0xFFFFFFFFu, digit);
                if (valid && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&hist[digit], (unsigned int)__popc(peers));
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                // warp-parallel descending scan of the 256 bins: lane l owns bins 255-8l .. 248-8l
                const int lane = threadIdx.x;
                unsigned int h[8], local = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) { h[k] = hist[255 - 8 * lane - k]; local += h[k]; }
                unsigned int incl = local;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { unsigned int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
                const unsigned int excl = incl - local;       // keys in strictly higher bins than this lane's group
                const unsigned int kk = s_kk;
                __syncwarp();
                const unsigned int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                const bool mine = (excl < kk) && (incl >= kk);  // the k-th key falls into one of my 8 bins
                if (total < kk) { if (lane == 0) s_short = 1u; }
                else if (mine) {
                    unsigned int cum = excl;
                    int k = 0;
                    for (; k < 8; ++k) { if (cum + h[k] >= kk) break; cum += h[k]; }
                    const int d = 255 - 8 * lane - k;
                    s_kk = kk - cum;
                    s_prefix = prefix | ((unsigned long long)d << shift);
                    // every key of the chosen bin is needed: the remaining low bits cannot change the selected set
                    if (h[k] == kk - cum && mode == 0) s_done = 1u;
                }
            }
            mask |= (255ull << shift);
            __syncthreads();
            if (s_short || s_done) break;
        }
        if (mode == 1) {
            // high word 0 == the empty key (fewer than `top` valid candidates): no threshold
            if (threadIdx.x == 0) thr[q] = (s_short || (s_prefix >> 32) == 0ull) ? __int_as_float(0xff800000) : qb_unorderable((uint32_t)(s_prefix >> 32));
            return;
        }
        const unsigned long long kth = s_short ? 1ull : s_prefix;  // short: take every non-empty key
        if (threadIdx.x == 0) s_fill = 0u;
        if (in_smem) {
            // the source IS buf: pull this thread's keys into registers before the buffer is cleared and refilled
            constexpr int PER = SORT_CAP / QB_SELECT_THREADS;
            unsigned long long mine[PER];
#pragma unroll
            for (int k = 0; k < PER; ++k) { const int i = threadIdx.x + k * QB_SELECT_THREADS; mine[k] = (i < m) ? buf[i] : 0ull; }
            __syncthreads();
            for (int i = threadIdx.x; i < SORT_CAP; i += blockDim.x) buf[i] = 0ull;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < PER; ++k)
                if (mine[k] >= kth && mine[k] != 0ull) buf[atomicAdd(&s_fill, 1u)] = mine[k];  // at most m <= SORT_CAP keys
        } else {
            for (int i = threadIdx.x; i < SORT_CAP; i += blockDim.x) buf[i] = 0ull;
            __syncthreads();
            for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x) {
                unsigned long long key = keys[i];
                if (key >= kth && key != 0ull) {
                    unsigned int p = atomicAdd(&s_fill, 1u);
                    if (p < SORT_CAP) buf[p] = key;
                }
            }
        }
        __syncthreads();
        m = (int)min(s_fill, (unsigned int)SORT_CAP);
    }

    int p2 = 32;
    while (p2 < m) p2 <<= 1;
    bitonic_sort_desc(buf, p2);

    if (mode == 1) {
        if (threadIdx.x == 0) {
            unsigned long long k = (top <= (uint32_t)SORT_CAP && top >= 1 && (int)top <= m) ? buf[top - 1] : 0ull;
            thr[q] = (k == 0ull) ? __int_as_float(0xff800000) : qb_key_score(k);
        }
        return;
    }
    // mode 0: write results; count = number of non-empty keys among the first `top`
    unsigned int valid = 0;
    for (int i = threadIdx.x; i < (int)top; i += blockDim.x) {
        unsigned long long k = (i < m) ? buf[i] : 0ull;
        qb_scored_point sp;
        if (k != 0ull) { sp.idx = qb_key_id(k); sp.score = qb_key_score(k); valid++; }
        else { sp.idx = 0; sp.score = 0.0f; }
        out[(unsigned long long)q * top + i] = sp;
    }
    // block-reduce `valid`
    __shared__ unsigned int s_valid;
    if (threadIdx.x == 0) s_valid = 0u;
    __syncthreads();
    if (valid) atomicAdd(&s_valid, valid);
    __syncthreads();
    if (threadIdx.x == 0) out_counts[q] = s_valid;
}

// top > SORT_CAP (oversampled tops: limit x oversampling easily exceeds 4096, vector_index_search_common.rs:27-46).  Same exact
// MSB radix select for the top-th key, then the selected keys are gathered into a global scratch row of p2 >= top slots and
// bitonic-sorted there.  One CTA per query; off the headline paths, so simplicity over speed.
__global__ void __launch_bounds__(QB_SELECT_THREADS)
qb_select_large_kernel(const unsigned long long* cand, const unsigned int* __restrict__ cnt, unsigned long long cap, unsigned long long fixed_n, uint32_t top,
                       unsigned long long* __restrict__ scratch, uint32_t p2, qb_scored_point* __restrict__ out, uint32_t* __restrict__ out_counts,
                       unsigned int* __restrict__ overflow) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_kk, s_fill, s_short;
    const uint32_t q = blockIdx.x;
    unsigned long long n;
    if (fixed_n) n = fixed_n;
    else {
        unsigned int c = cnt[q];
        if (c > cap) { if (threadIdx.x == 0 && overflow) atomicOr(overflow, 1u); c = (unsigned int)cap; }
        n = c;
    }
    const unsigned long long* keys = cand + (unsigned long long)q * cap;
    unsigned long long* buf = scratch + (unsigned long long)q * p2;
    if (threadIdx.x == 0) { s_prefix = 0ull; s_kk = top; s_short = 0u; s_fill = 0u; }
    unsigned long long mask = 0ull;
    __syncthreads();
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned long long key = keys[i];
            if (key != 0ull && (key & mask) == prefix) atomicAdd(&hist[(unsigned int)((key >> shift) & 255ull)], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int kk = s_kk, cum = 0;
            int d = 255;
            for (; d >= 0; --d) { if (cum + hist[d] >= kk) break; cum += hist[d]; }
            if (d < 0) s_short = 1u;                      // fewer than `top` non-empty keys: keep them all
            else { s_kk = kk - cum; s_prefix = prefix | ((unsigned long long)d << shift); }
        }
        mask |= (255ull << shift);
        __syncthreads();
        if (s_short) break;
    }
    const unsigned long long kth = s_short ? 1ull : s_prefix;
    for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x) buf[i] = 0ull;
    __syncthreads();
    for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long key = keys[i];
        if (key >= kth && key != 0ull) { const unsigned int p = atomicAdd(&s_fill, 1u); if (p < p2) buf[p] = key; }
    }
    __syncthreads();
    bitonic_sort_desc(buf, (int)p2);
    unsigned int valid = 0;
    for (uint32_t i = threadIdx.x; i < top; i += blockDim.x) {
        const unsigned long long k = buf[i];
        qb_scored_point sp;
        if (k != 0ull) { sp.idx = qb_key_id(k); sp.score = qb_key_score(k); valid++; }
        else { sp.idx = 0; sp.score = 0.0f; }
        out[(unsigned long long)q * top + i] = sp;
    }
    __shared__ unsigned int s_valid;
    if (threadIdx.x == 0) s_valid = 0u;
    __syncthreads();
    if (valid) atomicAdd(&s_valid, valid);
    __syncthreads();
    if (threadIdx.x == 0) out_counts[q] = s_valid;
}

__global__ void qb_fill_u32_kernel(unsigned int* p, unsigned int v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace

qb_status qb_launch_select(const unsigned long long* d_cand, const unsigned int* d_cnt, unsigned long long cap,
                           unsigned long long fixed_n, uint32_t nq, uint32_t top, int mode, qb_scored_point* d_out,
                           uint32_t* d_out_counts, float* d_thr, unsigned int* d_overflow, cudaStream_t stream) {
    QB_CHECK(top >= 1 && top <= QB_MAX_TOP, QB_ERR_UNSUPPORTED, "top=%u outside [1,%u]", top, QB_MAX_TOP);
    if (nq == 0) return QB_OK;
    if (mode == 0 && top > (uint32_t)SORT_CAP) {
        uint32_t p2 = 1;
        while (p2 < top) p2 <<= 1;
        unsigned long long* scratch = nullptr;
        QB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), (size_t)nq * p2 * 8, stream));
        qb_select_large_kernel<<<nq, QB_SELECT_THREADS, 0, stream>>>(d_cand, d_cnt, cap, fixed_n, top, scratch, p2, d_out, d_out_counts, d_overflow);
        QB_LAUNCHED();
        cudaError_t e = cudaGetLastError();
        cudaFreeAsync(scratch, stream);
        if (e != cudaSuccess) { qb_set_error("select (large top): %s", cudaGetErrorString(e)); return QB_ERR_CUDA; }
        return QB_OK;
    }
    qb_select_kernel<<<nq, QB_SELECT_THREADS, 0, stream>>>(d_cand, d_cnt, cap, fixed_n, top, mode, d_out, d_out_counts, d_thr,
                                                          d_overflow);
    QB_LAUNCHED();
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

qb_status qb_launch_fill_u32(unsigned int* p, unsigned int v, size_t n, cudaStream_t stream) {
    if (n == 0) return QB_OK;
    qb_fill_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, v, n);
    QB_LAUNCHED();
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}
