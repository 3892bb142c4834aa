// qb_prefilter.cu — single-query dense f32 (dot / cosine) scans at half the HBM bytes, results unchanged.
//
// The single-query scan of qb_dense.cu runs at the HBM copy rate: every query reads dim * 4 bytes per row.  The only way to more
// queries per second is fewer bytes per row — and the scan only has to FIND the rows whose exact score can reach the top-k; it does
// not have to produce their scores.  With a bf16 shadow plane of the rows (RN-even, the same plane the batched tensor-core path of
// qb_sq8_mma.cu keeps) and the query kept in f32:
//     | sum_i q_i bf16(x_i) - sum_i q_i x_i |  <=  2^-9 ||q|| ||x||            (Cauchy-Schwarz on the element-wise rounding errors)
//   + f32 evaluation of either sum, any order:  <=  dim 2^-24 ||q|| ||x|| each
//   =>  |approx - exact|  <=  eps_q := (2^-9 + dim 2^-22) ||q|| max_rows ||x||  (1 + 2^-10)
// Four launches per query, no host synchronisation:
//   1. the EXACT in-kernel-top-k scan (dense_f32_stream_kernel<., LOCALK>) over a prefix of the rows (1/64 of them, 2^14..2^17) -> thr_q = the k-th best exact score
//      of the sample (a lower bound of the final k-th score);
//   2. dense_bf16_filter_kernel: the whole bf16 plane streams through the same TMA-bulk ring (dim * 2 bytes per row); the query sits in
//      REGISTERS (a lane always meets the same dimensions); rows with approx >= thr_q - eps_q are appended to a candidate list — a
//      superset of the rows whose exact score reaches thr_q, hence of the true top-k;
//   3. f32_prefilter_finish_kernel: exact AVX-order scores of the candidates (a few hundred rows, 32 CTAs), top-k by the usual keys in the
//      last CTA to finish;
//   4. if the list overflowed (mass ties, a NaN query, a sample without k live rows), step 3 raises a device flag and the exact scan of
//      the whole storage — always enqueued, it exits at once while the flag is down — produces the answer instead.
// (RawScorer results are bit-identical to the exact path: tests/test_gpu_dense.py::test_single_query_prefilter_*.)
#include <algorithm>

#include "qb_internal.h"
#include "qb_score.cuh"

qb_status qb_dense_f32_scan_localk(const qb_storage* s, const QbScanArgs& a, uint32_t top, uint64_t* n_slots, cudaStream_t stream, uint64_t min_rows = 65536);   // qb_dense.cu

namespace {

constexpr int PF_CONSUMER_WARPS = 8;
constexpr int PF_MAX_PRODUCERS = 4;
constexpr int PF_PRODUCERS = 2;             // producer warps (one lane each): a bulk copy costs its issuing thread ~500 clocks (wait for the slot, arm the
                                            // barrier, issue), so ONE producer caps a CTA at one 12-KB slot per ~500 clocks — below the HBM rate for these planes
constexpr int PF_THREADS = 32 * (PF_CONSUMER_WARPS + PF_MAX_PRODUCERS);     // launch bound; the launch uses 32 * (consumers + producers)
constexpr uint32_t PF_SLOT_BYTES = 12288;  // target bytes per ring slot (a consumer warp holds one slot while the others are in flight)
constexpr uint32_t PF_CAP = 16384;         // candidate rows per query (a few hundred to a few thousand expected)

static int pf_producers() { const int o = qb_opt().prefilter_producers; return (o >= 1 && o <= PF_MAX_PRODUCERS) ? o : PF_PRODUCERS; }

struct PfParams {
    const uint8_t* rows;        // bf16 plane
    uint32_t stride;            // bytes per row = row_h * 2 (multiple of 16)
    uint32_t row_h;             // halfs per row
    uint32_t dim;
    uint64_t n_rows;
    const float* q;             // preprocessed f32 query
    uint32_t rows_per_slot, n_slots, slot_bytes;
    const qb_scored_point* samp_out; const uint32_t* samp_cnt; uint32_t top;   // exact top-k of the sample prefix
    const unsigned int* max_norm_bits;                                         // max row norm of the storage (f32 bits)
    uint32_t* cand; unsigned int* cnt;
    const uint32_t* deleted; const uint32_t* deleted2;
    int l2_keep;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// NCH = 16-byte chunks of a row per lane (row_h <= NCH * 256)
template <int NCH>
__global__ void __launch_bounds__(PF_THREADS, 1) dense_bf16_filter_kernel(const PfParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* slots = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(slots + (size_t)p.n_slots * p.slot_bytes);
    uint64_t* empty = full + p.n_slots;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t n_tiles = (p.n_rows + p.rows_per_slot - 1) / p.rows_per_slot;
    const uint64_t n_local = (blockIdx.x < n_tiles) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.n_slots; ++s) { qb_mbar_init(&full[s], 1); qb_mbar_init(&empty[s], 1); }
        qb_fence_barrier_init();
    }
    __syncthreads();
    const int n_prod = (int)(blockDim.x >> 5) - PF_CONSUMER_WARPS;
    if (warp < n_prod) {
        if (lane == 0) {
            const uint64_t policy = p.l2_keep ? qb_policy_evict_last() : qb_policy_evict_first();
            // slot / phase / first row advance incrementally: 64-bit divisions in this loop cost more than the copy they issue
            uint32_t s = (uint32_t)warp, ph = 0;                 // n_slots is a multiple of 8, n_prod of 1 / 2 / 4: s wraps exactly
            uint64_t r0 = ((uint64_t)blockIdx.x + (uint64_t)warp * gridDim.x) * p.rows_per_slot;
            const uint64_t r_step = (uint64_t)n_prod * gridDim.x * p.rows_per_slot;
            for (uint64_t i = warp; i < n_local; i += n_prod, r0 += r_step) {
                qb_mbar_wait(&empty[s], ph ^ 1u);
                const uint64_t left = p.n_rows - r0;
                const uint32_t nr = (uint32_t)(left < p.rows_per_slot ? left : p.rows_per_slot);
                const uint32_t bytes = nr * p.stride;
                qb_mbar_arrive_expect_tx(&full[s], bytes);
                qb_bulk_g2s(slots + (size_t)s * p.slot_bytes, p.rows + r0 * p.stride, bytes, &full[s], policy);
                s += (uint32_t)n_prod;
                if (s >= p.n_slots) { s -= p.n_slots; ph ^= 1u; }
            }
        }
        return;
    }
    const int cw = warp - n_prod;
    // this lane's slice of the query: dimensions [ (c * 32 + lane) * 8, + 8 ) for c < NCH, zero past dim
    float qr[NCH][8];
    uint32_t off[NCH];
    float qq = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const uint32_t d0 = (uint32_t)(c * 32 + lane) * 8;
        off[c] = (d0 < p.row_h) ? d0 * 2 : 0;                  // past the row: any valid address, the query slice is zero
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float v = (d0 < p.row_h && d0 + k < p.dim) ? p.q[d0 + k] : 0.f;
            qr[c][k] = v;
            qq = fmaf(v, v, qq);
        }
    }
    qq = warp_sum(qq);
    // thr_q - eps_q, rounded towards "pass"
    float thr_adj;
    {
        const float qn = __fmul_ru(__fsqrt_ru(qq), 1.0001f);
        const float mx = __uint_as_float(*p.max_norm_bits);
        const float c = __fadd_ru(0x1.004p-9f, __fmul_ru((float)p.dim, 0x1p-22f));       // (2^-9)(1 + 2^-10) + dim 2^-22
        const float eps = __fadd_ru(__fmul_ru(__fmul_ru(qn, mx), c), 1.0e-37f);
        const float thr = (*p.samp_cnt >= p.top) ? p.samp_out[p.top - 1].score : __int_as_float(0xff800000);
        thr_adj = __fsub_rd(thr, eps);                           // NaN query -> NaN: `approx < NaN` is false, every row passes (-> fallback)
    }
    uint32_t s = (uint32_t)cw, ph = 0;                            // as in the producers: no 64-bit division per slot
    uint64_t r0 = ((uint64_t)blockIdx.x + (uint64_t)cw * gridDim.x) * p.rows_per_slot;
    const uint64_t r_step = (uint64_t)PF_CONSUMER_WARPS * gridDim.x * p.rows_per_slot;
    for (uint64_t i = cw; i < n_local; i += PF_CONSUMER_WARPS, r0 += r_step) {
        const uint64_t left = p.n_rows - r0;
        const uint32_t nr = (uint32_t)(left < p.rows_per_slot ? left : p.rows_per_slot);
        qb_mbar_wait(&full[s], ph);
        const uint8_t* slot = slots + (size_t)s * p.slot_bytes;
        for (uint32_t r = 0; r < nr; r += 2) {
            const bool two = r + 1 < nr;
            const uint8_t* ra = slot + (size_t)r * p.stride;
            const uint8_t* rb = slot + (size_t)(two ? r + 1 : r) * p.stride;
            float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const uint4 va = *reinterpret_cast<const uint4*>(ra + off[c]);
                const uint4 vb = *reinterpret_cast<const uint4*>(rb + off[c]);
                const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    a0 = fmaf(qr[c][2 * k], __uint_as_float(wa[k] << 16), a0);
                    a1 = fmaf(qr[c][2 * k + 1], __uint_as_float(wa[k] & 0xFFFF0000u), a1);
                    b0 = fmaf(qr[c][2 * k], __uint_as_float(wb[k] << 16), b0);
                    b1 = fmaf(qr[c][2 * k + 1], __uint_as_float(wb[k] & 0xFFFF0000u), b1);
                }
            }
            float sa = a0 + a1, sb = b0 + b1;
            // both sums in one butterfly: lanes < 16 end up with row a, lanes >= 16 with row b
            {
                const float send = (lane & 16) ? sa : sb, keep = (lane & 16) ? sb : sa;
                float v = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 16);
#pragma unroll
                for (int o = 8; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
                sa = v;
            }
            if ((lane == 0 || (lane == 16 && two)) && !(sa < thr_adj)) {
                const uint32_t id = (uint32_t)(r0 + r + (lane >> 4));
                bool dead = false;
                if (p.deleted) dead = (p.deleted[id >> 5] >> (id & 31)) & 1u;
                if (p.deleted2) dead = dead || ((p.deleted2[id >> 5] >> (id & 31)) & 1u);
                if (!dead) {
                    const unsigned int pos = atomicAdd(p.cnt, 1u);
                    if (pos < PF_CAP) p.cand[pos] = id;
                }
            }
        }
        __syncwarp();
        if (lane == 0) qb_mbar_arrive(&empty[s]);
        s += PF_CONSUMER_WARPS;
        if (s >= p.n_slots) { s -= p.n_slots; ph ^= 1u; }
    }
}

// ------------------------------------------------------------------------------------------------ int8 shadow plane (a quarter of the bytes)
// x_i ~ s_r c_i, c_i = rint(x_i / s_r) in [-127, 127], s_r = max_i |x_i| / 127 (one f32 scale per row, stored right behind the row's codes);
// the QUERY is split into two int8 levels, q_i ~ s_q (h_i + l_i / 254), so that its own error is negligible and the scan is integer only:
//   H = sum_i h_i c_i, L = sum_i l_i c_i   (dp4a, exact: dim * 127^2 < 2^24 for dim <= 1040),   approx = s_r s_q (H + L / 254)
//   | exact - approx |  <=  s_r (1/2 + 2^-13) ||q||_1          (row rounding:   |x_i - s_r c_i| <= s_r (1/2 + 1e-4))
//                        +  s_r s_q (127 dim / 508)(1 + 0.08)   (query rounding: |q_i - q^_i| <= s_q (1/508 + 2.3e-5), |c_i| <= 127)
//                        +  slack_q = (dim 2^-22 + 2^-17)(1 + sqrt(dim) / 127) ||q|| max||x||     (every f32 evaluation on either side)
// so a row passes iff  s_r * (s_q (H + L / 254) + E_q) >= thr_q - slack_q  with the per-query constant
// E_q = (1/2 + 2^-13) ||q||_1 + 0.27 s_q dim: one FMA, one multiply and one compare per row after two warp-wide integer sums.
__global__ void __launch_bounds__(256) f32_to_q8_rows_kernel(const float* __restrict__ rows, uint64_t stride_f, uint32_t dim, uint64_t n, int8_t* __restrict__ out,
                                                              uint32_t out_stride_b /* codes + 16: the row's f32 scale follows its codes */, unsigned int* __restrict__ max_norm_bits,
                                                              unsigned int* __restrict__ nonfinite) {
    const int t = threadIdx.x & 7;
    const uint64_t groups = (uint64_t)gridDim.x * (blockDim.x >> 3), g0 = (uint64_t)blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);
    const uint64_t n_iter = (n + groups - 1) / groups;
    for (uint64_t it = 0; it < n_iter; ++it) {
        const uint64_t r = g0 + it * groups;
        const bool valid = r < n;
        const float* src = rows + (valid ? r : 0) * stride_f;
        int8_t* dst = out + (valid ? r : 0) * out_stride_b;
        float ss = 0.f, mx = 0.f;
        bool bad = false;
        for (uint32_t i = t; i < dim; i += 8) {
            const float v = src[i];
            bad |= !(fabsf(v) <= 3.0e38f);
            ss = fmaf(v, v, ss);
            mx = fmaxf(mx, fabsf(v));
        }
        ss += __shfl_xor_sync(0xFFFFFFFFu, ss, 1); ss += __shfl_xor_sync(0xFFFFFFFFu, ss, 2); ss += __shfl_xor_sync(0xFFFFFFFFu, ss, 4);
        mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, 1)); mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, 2)); mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, 4));
        bad = __shfl_xor_sync(0xFFFFFFFFu, (int)bad, 1) | __shfl_xor_sync(0xFFFFFFFFu, (int)bad, 2) | __shfl_xor_sync(0xFFFFFFFFu, (int)bad, 4) | (int)bad;
        const bool tiny = !(mx >= 1.0e-30f);                      // zero (or denormal-only) rows: all codes 0, scale = max so that the error bound still holds
        const float sr = tiny ? __fmul_ru(mx, 2.0f) : __fdiv_rn(mx, 127.f);
        const float inv = tiny ? 0.f : __fdiv_rn(127.f, mx);
        for (uint32_t i = t; i < out_stride_b - 16; i += 8) {
            const float v = (i < dim) ? src[i] : 0.f;
            const float c = fminf(fmaxf(rintf(__fmul_rn(v, inv)), -127.f), 127.f);
            if (valid) dst[i] = (int8_t)(int)c;
        }
        if (valid && t == 0) {
            *reinterpret_cast<float4*>(dst + (out_stride_b - 16)) = make_float4(sr, 0.f, 0.f, 0.f);
            if (bad || !(ss <= 3.0e38f)) atomicOr(nonfinite, 1u);
            else atomicMax(max_norm_bits, __float_as_uint(sqrtf(ss) * 1.000001f));
        }
    }
}

struct Pf8Params {
    const uint8_t* rows;        // int8 plane: per row `stride - 16` code bytes, then the row's f32 scale (+ 12 bytes of padding)
    uint32_t stride;            // bytes per row (multiple of 16)
    uint32_t dim;
    uint64_t n_rows;
    const float* q;
    uint32_t rows_per_slot, n_slots, slot_bytes;
    const qb_scored_point* samp_out; const uint32_t* samp_cnt; uint32_t top;
    const unsigned int* max_norm_bits;
    uint32_t* cand; unsigned int* cnt;
    const uint32_t* deleted; const uint32_t* deleted2;
    int l2_keep;
};

__device__ __forceinline__ uint32_t pack_s8x4(int a, int b, int c, int d) { return (uint32_t)(a & 255) | ((uint32_t)(b & 255) << 8) | ((uint32_t)(c & 255) << 16) | ((uint32_t)(d & 255) << 24); }

// NCH = 8-byte chunks of a row's codes per lane (stride - 16 <= NCH * 256)
template <int NCH>
__global__ void __launch_bounds__(PF_THREADS, 1) dense_q8_filter_kernel(const Pf8Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* slots = smem;                                       // [n_slots][slot_bytes]
    uint64_t* full = reinterpret_cast<uint64_t*>(slots + (size_t)p.n_slots * p.slot_bytes);
    uint64_t* empty = full + p.n_slots;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t n_tiles = (p.n_rows + p.rows_per_slot - 1) / p.rows_per_slot;
    const uint64_t n_local = (blockIdx.x < n_tiles) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t code_b = p.stride - 16;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.n_slots; ++s) { qb_mbar_init(&full[s], 1); qb_mbar_init(&empty[s], 1); }
        qb_fence_barrier_init();
    }
    __syncthreads();
    const int n_prod = (int)(blockDim.x >> 5) - PF_CONSUMER_WARPS;
    if (warp < n_prod) {
        if (lane == 0) {
            const uint64_t policy = p.l2_keep ? qb_policy_evict_last() : qb_policy_evict_first();
            // slot / phase / first row advance incrementally: 64-bit divisions in this loop cost more than the copy they issue
            uint32_t s = (uint32_t)warp, ph = 0;                 // n_slots is a multiple of 8, n_prod of 1 / 2 / 4: s wraps exactly
            uint64_t r0 = ((uint64_t)blockIdx.x + (uint64_t)warp * gridDim.x) * p.rows_per_slot;
            const uint64_t r_step = (uint64_t)n_prod * gridDim.x * p.rows_per_slot;
            for (uint64_t i = warp; i < n_local; i += n_prod, r0 += r_step) {
                qb_mbar_wait(&empty[s], ph ^ 1u);
                const uint64_t left = p.n_rows - r0;
                const uint32_t nr = (uint32_t)(left < p.rows_per_slot ? left : p.rows_per_slot);
                const uint32_t bytes = nr * p.stride;
                qb_mbar_arrive_expect_tx(&full[s], bytes);
                qb_bulk_g2s(slots + (size_t)s * p.slot_bytes, p.rows + r0 * p.stride, bytes, &full[s], policy);
                s += (uint32_t)n_prod;
                if (s >= p.n_slots) { s -= p.n_slots; ph ^= 1u; }
            }
        }
        return;
    }
    const int cw = warp - n_prod;
    // query statistics over the whole vector, then this lane's slice quantised to two int8 levels
    float qv[NCH][8];
    uint32_t off[NCH];
    float qmax = 0.f, q1 = 0.f, q2 = 0.f;
    bool qbad = false;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const uint32_t d0 = (uint32_t)(c * 32 + lane) * 8;
        off[c] = (d0 < code_b) ? d0 : 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float v = (d0 < code_b && d0 + k < p.dim) ? p.q[d0 + k] : 0.f;
            qv[c][k] = v;
            qbad |= !(fabsf(v) <= 3.0e38f);
            qmax = fmaxf(qmax, fabsf(v)); q1 = __fadd_ru(q1, fabsf(v)); q2 = fmaf(v, v, q2);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        qmax = fmaxf(qmax, __shfl_xor_sync(0xFFFFFFFFu, qmax, o)); q1 = __fadd_ru(q1, __shfl_xor_sync(0xFFFFFFFFu, q1, o)); q2 += __shfl_xor_sync(0xFFFFFFFFu, q2, o);
        qbad |= __shfl_xor_sync(0xFFFFFFFFu, (int)qbad, o) != 0;
    }
    qbad |= (qmax > 0.f && qmax < 1.0e-30f);                     // 127 / qmax would overflow: leave such a query to the exact scan
    const float sq = (qmax > 0.f) ? __fdiv_rn(qmax, 127.f) : 0.f;
    const float inv_sq = (qmax > 0.f && !qbad) ? __fdiv_rn(127.f, qmax) : 0.f;
    uint32_t hq[NCH][2], lq[NCH][2];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        int h[8], l[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float y = __fmul_rn(qv[c][k], inv_sq);
            const float hf = fminf(fmaxf(rintf(y), -127.f), 127.f);
            h[k] = (int)hf;
            l[k] = (int)fminf(fmaxf(rintf(__fmul_rn(__fsub_rn(y, hf), 254.f)), -127.f), 127.f);
        }
        hq[c][0] = pack_s8x4(h[0], h[1], h[2], h[3]); hq[c][1] = pack_s8x4(h[4], h[5], h[6], h[7]);
        lq[c][0] = pack_s8x4(l[0], l[1], l[2], l[3]); lq[c][1] = pack_s8x4(l[4], l[5], l[6], l[7]);
    }
    // E_q and thr_q - slack_q (rounded towards "pass"); a non-finite query makes every row pass (-> fallback to the exact scan)
    const float qn = __fmul_ru(__fsqrt_ru(q2), 1.0001f);
    const float mxn = __uint_as_float(*p.max_norm_bits);
    const float e_q = __fadd_ru(__fmul_ru(q1, 0x1.001p-1f), __fmul_ru(__fmul_ru(sq, (float)p.dim), 0.27f));      // (1/2 + 2^-13)||q||_1 + s_q dim (127/508)(1 + 0.08)
    const float slack = __fadd_ru(__fmul_ru(__fmul_ru(__fadd_ru(__fmul_ru((float)p.dim, 0x1p-22f), 0x1p-17f), __fadd_ru(1.f, __fdiv_ru(__fsqrt_ru((float)p.dim), 127.f))),
                                            __fmul_ru(qn, mxn)), 1.0e-37f);
    const float thr = (*p.samp_cnt >= p.top) ? p.samp_out[p.top - 1].score : __int_as_float(0xff800000);
    const float thr_adj = qbad ? __int_as_float(0x7fc00000) : __fsub_rd(thr, slack);
    const float k254 = 1.0f / 254.0f;
    uint32_t s = (uint32_t)cw, ph = 0;                            // as in the producers: no 64-bit division per slot
    uint64_t r0 = ((uint64_t)blockIdx.x + (uint64_t)cw * gridDim.x) * p.rows_per_slot;
    const uint64_t r_step = (uint64_t)PF_CONSUMER_WARPS * gridDim.x * p.rows_per_slot;
    for (uint64_t i = cw; i < n_local; i += PF_CONSUMER_WARPS, r0 += r_step) {
        const uint64_t left = p.n_rows - r0;
        const uint32_t nr = (uint32_t)(left < p.rows_per_slot ? left : p.rows_per_slot);
        qb_mbar_wait(&full[s], ph);
        const uint8_t* slot = slots + (size_t)s * p.slot_bytes;
        for (uint32_t r = 0; r < nr; r += 2) {
            const bool two = r + 1 < nr;
            const uint8_t* ra = slot + (size_t)r * p.stride;
            const uint8_t* rb = slot + (size_t)(two ? r + 1 : r) * p.stride;
            int ha = 0, la = 0, hb = 0, lb = 0;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const uint2 va = *reinterpret_cast<const uint2*>(ra + off[c]);
                const uint2 vb = *reinterpret_cast<const uint2*>(rb + off[c]);
                ha = __dp4a((int)va.x, (int)hq[c][0], ha); ha = __dp4a((int)va.y, (int)hq[c][1], ha);
                la = __dp4a((int)va.x, (int)lq[c][0], la); la = __dp4a((int)va.y, (int)lq[c][1], la);
                hb = __dp4a((int)vb.x, (int)hq[c][0], hb); hb = __dp4a((int)vb.y, (int)hq[c][1], hb);
                lb = __dp4a((int)vb.x, (int)lq[c][0], lb); lb = __dp4a((int)vb.y, (int)lq[c][1], lb);
            }
            ha = __reduce_add_sync(0xFFFFFFFFu, ha); la = __reduce_add_sync(0xFFFFFFFFu, la);
            hb = __reduce_add_sync(0xFFFFFFFFu, hb); lb = __reduce_add_sync(0xFFFFFFFFu, lb);
            if (lane < 2 && (lane == 0 || two)) {
                const int H = lane ? hb : ha, L = lane ? lb : la;
                const float v = fmaf(sq, fmaf((float)L, k254, (float)H), e_q);
                const float up = __fmul_ru(*reinterpret_cast<const float*>((lane ? rb : ra) + code_b), v);    // an upper bound of the exact score (up to slack_q)
                if (!(up < thr_adj)) {
                    const uint32_t id = (uint32_t)(r0 + r + lane);
                    bool dead = false;
                    if (p.deleted) dead = (p.deleted[id >> 5] >> (id & 31)) & 1u;
                    if (p.deleted2) dead = dead || ((p.deleted2[id >> 5] >> (id & 31)) & 1u);
                    if (!dead) {
                        const unsigned int pos = atomicAdd(p.cnt, 1u);
                        if (pos < PF_CAP) p.cand[pos] = id;
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) qb_mbar_arrive(&empty[s]);
        s += PF_CONSUMER_WARPS;
        if (s >= p.n_slots) { s -= p.n_slots; ph ^= 1u; }
    }
}

template <int NCH>
qb_status launch_filter_q8(Pf8Params& p, int sm_count, cudaStream_t stream) {
    const uint32_t kMaxSmem = 227 * 1024;
    const uint32_t target = qb_opt().prefilter_slot_bytes ? qb_opt().prefilter_slot_bytes : PF_SLOT_BYTES;
    uint32_t rps = (target / p.stride) & ~1u;
    if (rps < 2) rps = 2;
    p.rows_per_slot = rps;
    p.slot_bytes = rps * p.stride;
    uint32_t n_slots = (kMaxSmem - 2048) / p.slot_bytes;
    if (n_slots > 64) n_slots = 64;
    n_slots = (n_slots / PF_CONSUMER_WARPS) * PF_CONSUMER_WARPS;
    QB_CHECK(n_slots >= (uint32_t)PF_CONSUMER_WARPS, QB_ERR_INVALID, "prefilter: rows too wide for the ring");
    p.n_slots = n_slots;
    const size_t smem = (size_t)n_slots * p.slot_bytes + (size_t)n_slots * 16;
    QB_CUDA(cudaFuncSetAttribute(dense_q8_filter_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    const uint64_t n_tiles = ceil_div_u64(p.n_rows, p.rows_per_slot);
    const unsigned grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)sm_count);
    dense_q8_filter_kernel<NCH><<<grid, 32 * (PF_CONSUMER_WARPS + pf_producers()), smem, stream>>>(p);
    QB_LAUNCHED();
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

// exact scores of the candidates in score_avx_group8's order (all CTAs, one 8-lane group per candidate), then the LAST CTA to finish picks
// the top-k by (score desc, id asc) — or raises the fallback flag when the list overflowed / the sample gave no threshold
constexpr int PF_FINISH_CTAS = 32;
__global__ void __launch_bounds__(256) f32_prefilter_finish_kernel(const uint8_t* __restrict__ rows, uint32_t stride, uint32_t dim, const float* __restrict__ q,
                                                                   const uint32_t* __restrict__ cand, unsigned int* __restrict__ cnt, uint32_t top, uint32_t id_base,
                                                                   unsigned long long* __restrict__ keys, unsigned int* __restrict__ ticket, qb_scored_point* __restrict__ out,
                                                                   uint32_t* __restrict__ out_cnt, unsigned int* __restrict__ fallback, unsigned int* __restrict__ n_fallbacks) {
    __shared__ unsigned long long s_best[8];
    __shared__ unsigned int s_ticket;
    const unsigned int c = *cnt;                                  // nobody resets it before every CTA has drawn its ticket
    const bool bad = c > PF_CAP || c < top;                       // overflow, or a sample that could not give a threshold
    const int t = threadIdx.x & 7, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (!bad) {
        const uint32_t groups = gridDim.x * (blockDim.x >> 3), g = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);
        const uint32_t n_iter = (c + groups - 1) / groups;        // uniform: the 8-lane scorer shuffles with the full warp mask
        for (uint32_t it = 0; it < n_iter; ++it) {
            const uint32_t i = g + it * groups;
            const bool valid = i < c;
            const uint32_t row = valid ? cand[i] : 0u;
            const float sc = qbs::score_avx_group8<qbs::M_DOT>(reinterpret_cast<const float*>(rows + (size_t)row * stride), q, dim, t);
            if (valid && t == 0) keys[i] = qb_pack_key(sc, row + id_base);
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    if (threadIdx.x == 0) { *cnt = 0u; *ticket = 0u; *fallback = bad ? 1u : 0u; if (bad) atomicAdd(n_fallbacks, 1u); }   // ready for the next query on this context
    if (bad) return;
    // keys are unique (the id is part of the key): round r takes the largest key below round r-1's winner
    unsigned long long prev = ~0ull;
    for (uint32_t r = 0; r < top; ++r) {
        unsigned long long best = 0ull;
        for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) { const unsigned long long k = __ldcg(keys + i); if (k < prev && k > best) best = k; }
#pragma unroll
        for (int o = 16; o; o >>= 1) { const unsigned long long w = __shfl_xor_sync(0xFFFFFFFFu, best, o); best = w > best ? w : best; }
        if (lane == 0) s_best[warp] = best;
        __syncthreads();
        best = s_best[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) best = s_best[w] > best ? s_best[w] : best;
        __syncthreads();
        if (threadIdx.x == 0) { qb_scored_point sp; sp.idx = qb_key_id(best); sp.score = qb_key_score(best); out[r] = sp; }
        prev = best;
    }
    if (threadIdx.x == 0) *out_cnt = top;
}

template <int NCH>
qb_status launch_filter(PfParams& p, int sm_count, cudaStream_t stream) {
    const uint32_t kMaxSmem = 227 * 1024;
    const uint32_t target = qb_opt().prefilter_slot_bytes ? qb_opt().prefilter_slot_bytes : PF_SLOT_BYTES;
    uint32_t rps = (target / p.stride) & ~1u;
    if (rps < 2) rps = 2;
    p.rows_per_slot = rps;
    p.slot_bytes = rps * p.stride;
    uint32_t n_slots = (kMaxSmem - 2048) / p.slot_bytes;
    if (n_slots > 64) n_slots = 64;
    n_slots = (n_slots / PF_CONSUMER_WARPS) * PF_CONSUMER_WARPS;
    QB_CHECK(n_slots >= (uint32_t)PF_CONSUMER_WARPS, QB_ERR_INVALID, "prefilter: rows too wide for the ring");
    p.n_slots = n_slots;
    const size_t smem = (size_t)n_slots * p.slot_bytes + (size_t)n_slots * 16;
    QB_CUDA(cudaFuncSetAttribute(dense_bf16_filter_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    const uint64_t n_tiles = ceil_div_u64(p.n_rows, p.rows_per_slot);
    const unsigned grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)sm_count);
    dense_bf16_filter_kernel<NCH><<<grid, 32 * (PF_CONSUMER_WARPS + pf_producers()), smem, stream>>>(p);
    QB_LAUNCHED();
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

}  // namespace

// int8 shadow plane (codes + the row's scale) of a dense f32 storage, built on first use and rebuilt after rows were rewritten: +26 % HBM
static qb_status q8_shadow_ensure(qb_storage* s, cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->q8_ready) return QB_OK;
    const uint32_t row_b = (uint32_t)round_up_u64(s->dim, 16) + 16;      // codes, then the row's f32 scale in its own 16 bytes: one bulk copy per tile
    if (!s->d_q8) {
        QB_CUDA(cudaMalloc(&s->d_q8, std::max<size_t>((size_t)s->count * row_b, 256)));
        QB_CUDA(cudaMalloc(&s->d_q8_meta, 256));
        s->hbm_bytes += (uint64_t)s->count * row_b;
    }
    s->q8_row_b = row_b;
    QB_CUDA(cudaMemsetAsync(s->d_q8_meta, 0, 256, stream));
    const uint64_t blocks = std::min<uint64_t>(ceil_div_u64(std::max<uint64_t>(s->count, 1), 32), (uint64_t)s->sm_count * 16);
    f32_to_q8_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const float*>(s->d_rows), s->row_stride / 4, s->dim, s->count, s->d_q8, row_b, s->d_q8_meta,
                                                                s->d_q8_meta + 1);
    QB_LAUNCHED();
    QB_CUDA(cudaGetLastError());
    unsigned int meta[2] = {0, 0};
    QB_CUDA(cudaMemcpyAsync(meta, s->d_q8_meta, 8, cudaMemcpyDeviceToHost, stream));
    QB_CUDA(cudaStreamSynchronize(stream));
    s->q8_ready = true;
    s->q8_usable = meta[1] == 0;
    return QB_OK;
}
static bool use_q8_plane(const qb_storage* s) { return qb_opt().prefilter_plane != 1 && s->dim <= 1024; }   // dim * 127^2 < 2^24 and <= 4 chunks per lane

// Can this storage answer single-query top-k searches through a shadow-plane prefilter?  Builds the plane on first use.
bool qb_f32_prefilter_usable(qb_storage* s, uint64_t n_rows, uint32_t top, cudaStream_t stream) {
    if (s->kind != QB_KIND_DENSE || s->dtype != QB_DT_F32) return false;
    if (s->distance != QB_DIST_DOT && s->distance != QB_DIST_COSINE) return false;       // the bound is on a dot product
    if (qb_opt().disable_prefilter || n_rows != s->count || n_rows < (1ull << 19) || top > 16 || s->dim < 32 || round_up_u64(s->dim, 8) > 1024) return false;
    if (use_q8_plane(s)) {
        if (q8_shadow_ensure(s, stream) == QB_OK) return s->q8_usable;
        cudaGetLastError();                                                               // e.g. no room for the plane: try the other one / stay exact
    }
    if (qb_f32_shadow_ensure(s, stream) != QB_OK) { cudaGetLastError(); return false; }
    return s->bf16_usable;
}

// d_q = preprocessed query; the exact top-`top` of the storage lands in d_out / d_out_cnt.  `a` = the scan arguments of the exact
// in-kernel-top-k path (emit.cand / final_out / done_counter set up by the caller); scratch = c->d_pf (see qb_f32_prefilter_scratch_bytes).
size_t qb_f32_prefilter_scratch_bytes() { return (size_t)PF_CAP * 12 + 16 * sizeof(qb_scored_point) + 256; }

qb_status qb_f32_prefilter_search(qb_storage* s, const QbScanArgs& a, uint32_t top, void* d_scratch, unsigned int* d_n_fallbacks, qb_scored_point* d_out, uint32_t* d_out_cnt,
                                  cudaEvent_t prof0, cudaEvent_t prof1, cudaStream_t stream) {
    uint8_t* sc = reinterpret_cast<uint8_t*>(d_scratch);
    // scratch: [0,16) cnt | [16,32) fallback flag | [32,48) finish ticket | [48,64) sample count | [64, 64+16*8) sample top-k | candidate keys | candidate rows
    unsigned int* d_cnt = reinterpret_cast<unsigned int*>(sc);
    unsigned int* d_fallback = reinterpret_cast<unsigned int*>(sc + 16);
    uint32_t* d_samp_cnt = reinterpret_cast<uint32_t*>(sc + 48);
    qb_scored_point* d_samp = reinterpret_cast<qb_scored_point*>(sc + 64);
    unsigned int* d_ticket = reinterpret_cast<unsigned int*>(sc + 32);
    unsigned long long* d_keys = reinterpret_cast<unsigned long long*>(sc + 64 + 16 * sizeof(qb_scored_point));
    uint32_t* d_cand = reinterpret_cast<uint32_t*>(d_keys + PF_CAP);
    const uint64_t n = s->count;
    // 1. exact top-k of a prefix
    // 1/64 of the rows, 2^14..2^17: a shard of a sharded search (1.25M rows at N = 8) should not spend a fifth of its step on the sample
    uint64_t sample = std::min<uint64_t>(131072, std::max<uint64_t>(16384, (n / 64) & ~(uint64_t)3));
    if (qb_opt().sample_rows) sample = std::min<uint64_t>(n, std::max<uint64_t>(16384, qb_opt().sample_rows & ~(uint64_t)3));   // experiments
    QbScanArgs as = a;
    as.row_begin = 0; as.row_end = sample;
    as.emit.final_out = d_samp; as.emit.final_count = d_samp_cnt; as.emit.run_if = nullptr;
    uint64_t n_slots = 0;
    QB_TRY(qb_dense_f32_scan_localk(s, as, top, &n_slots, stream, 4096));
    QB_CHECK(n_slots != 0 && n_slots <= 4096, QB_ERR_CUDA, "prefilter: the sample scan did not launch (%llu slots)", (unsigned long long)n_slots);
    // 2. filter pass over the whole shadow plane
    const float* d_q = reinterpret_cast<const float*>(a.d_q_enc);
    if (use_q8_plane(s) && s->q8_ready && s->q8_usable) {
        Pf8Params p8{};
        p8.rows = reinterpret_cast<const uint8_t*>(s->d_q8); p8.stride = s->q8_row_b; p8.dim = s->dim; p8.n_rows = n;
        p8.q = d_q; p8.samp_out = d_samp; p8.samp_cnt = d_samp_cnt; p8.top = top; p8.max_norm_bits = s->d_q8_meta;
        p8.cand = d_cand; p8.cnt = d_cnt; p8.deleted = a.emit.deleted; p8.deleted2 = a.emit.deleted2;
        p8.l2_keep = ((uint64_t)n * p8.stride <= (64ull << 20)) ? 1 : 0;
        if (prof0) cudaEventRecord(prof0, stream);
        switch ((p8.stride - 16 + 255) / 256) {
            case 1: QB_TRY(launch_filter_q8<1>(p8, s->sm_count, stream)); break;
            case 2: QB_TRY(launch_filter_q8<2>(p8, s->sm_count, stream)); break;
            case 3: QB_TRY(launch_filter_q8<3>(p8, s->sm_count, stream)); break;
            default: QB_TRY(launch_filter_q8<4>(p8, s->sm_count, stream)); break;
        }
        if (prof1) cudaEventRecord(prof1, stream);
    } else {
    PfParams p{};
    p.rows = reinterpret_cast<const uint8_t*>(s->d_bf16); p.row_h = s->bf16_row_h; p.stride = s->bf16_row_h * 2; p.dim = s->dim; p.n_rows = n;
    p.q = reinterpret_cast<const float*>(a.d_q_enc);
    p.samp_out = d_samp; p.samp_cnt = d_samp_cnt; p.top = top; p.max_norm_bits = s->d_bf16_meta;
    p.cand = d_cand; p.cnt = d_cnt; p.deleted = a.emit.deleted; p.deleted2 = a.emit.deleted2;
    p.l2_keep = ((uint64_t)n * p.stride <= (64ull << 20)) ? 1 : 0;
    if (prof0) cudaEventRecord(prof0, stream);
    const uint32_t nch = (p.row_h + 255) / 256;
    switch (nch) {
        case 1: QB_TRY(launch_filter<1>(p, s->sm_count, stream)); break;
        case 2: QB_TRY(launch_filter<2>(p, s->sm_count, stream)); break;
        case 3: QB_TRY(launch_filter<3>(p, s->sm_count, stream)); break;
        default: QB_TRY(launch_filter<4>(p, s->sm_count, stream)); break;
    }
    if (prof1) cudaEventRecord(prof1, stream);
    }
    // 3. exact scores + top-k of the survivors (or the fallback flag)
    f32_prefilter_finish_kernel<<<PF_FINISH_CTAS, 256, 0, stream>>>(reinterpret_cast<const uint8_t*>(s->d_rows), s->row_stride, s->dim, d_q, d_cand, d_cnt, top, a.emit.id_base,
                                                                    d_keys, d_ticket, d_out, d_out_cnt, d_fallback, d_n_fallbacks);
    QB_LAUNCHED();
    QB_CUDA(cudaGetLastError());
    // 4. the exact scan of everything: its CTAs return at once unless the flag is up
    QbScanArgs af = a;
    af.row_begin = 0; af.row_end = n;
    af.emit.final_out = d_out; af.emit.final_count = d_out_cnt; af.emit.run_if = d_fallback;
    QB_TRY(qb_dense_f32_scan_localk(s, af, top, &n_slots, stream));
    QB_CHECK(n_slots != 0 && n_slots <= 4096, QB_ERR_CUDA, "prefilter: the fallback scan did not launch (%llu slots)", (unsigned long long)n_slots);
    return QB_OK;
}

