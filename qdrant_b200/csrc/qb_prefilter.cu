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
//   1. the EXACT in-kernel-top-k scan (dense_f32_stream_kernel<., LOCALK>) over a prefix of the rows -> thr_q = the k-th best exact score
//      of the sample (a lower bound of the final k-th score);
//   2. dense_bf16_filter_kernel: the whole bf16 plane streams through the same TMA-bulk ring (dim * 2 bytes per row); the query sits in
//      REGISTERS (a lane always meets the same dimensions); rows with approx >= thr_q - eps_q are appended to a candidate list — a
//      superset of the rows whose exact score reaches thr_q, hence of the true top-k;
//   3. f32_prefilter_finish_kernel (one CTA): exact AVX-order scores of the candidates (a few hundred rows), top-k by the usual keys;
//   4. if the list overflowed (mass ties, a NaN query, a sample without k live rows), step 3 raises a device flag and the exact scan of
//      the whole storage — always enqueued, it exits at once while the flag is down — produces the answer instead.
// (RawScorer results are bit-identical to the exact path: tests/test_gpu_dense.py::test_single_query_prefilter_*.)
#include <algorithm>

#include "qb_internal.h"
#include "qb_score.cuh"

qb_status qb_dense_f32_scan_localk(const qb_storage* s, const QbScanArgs& a, uint32_t top, uint64_t* n_slots, cudaStream_t stream);   // qb_dense.cu

namespace {

constexpr int PF_CONSUMER_WARPS = 8;
constexpr int PF_THREADS = 32 * (PF_CONSUMER_WARPS + 1);
constexpr uint32_t PF_CAP = 4096;          // candidate rows per query (a few hundred expected)

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
    if (warp == 0) {
        if (lane == 0) {
            const uint64_t policy = p.l2_keep ? qb_policy_evict_last() : qb_policy_evict_first();
            for (uint64_t i = 0; i < n_local; ++i) {
                const uint32_t s = (uint32_t)(i % p.n_slots);
                const uint32_t ph = (uint32_t)((i / p.n_slots) & 1);
                qb_mbar_wait(&empty[s], ph ^ 1u);
                const uint64_t r0 = (blockIdx.x + i * gridDim.x) * p.rows_per_slot;
                const uint64_t left = p.n_rows - r0;
                const uint32_t nr = (uint32_t)(left < p.rows_per_slot ? left : p.rows_per_slot);
                const uint32_t bytes = nr * p.stride;
                qb_mbar_arrive_expect_tx(&full[s], bytes);
                qb_bulk_g2s(slots + (size_t)s * p.slot_bytes, p.rows + r0 * p.stride, bytes, &full[s], policy);
            }
        }
        return;
    }
    const int cw = warp - 1;
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
    for (uint64_t i = cw; i < n_local; i += PF_CONSUMER_WARPS) {
        const uint32_t s = (uint32_t)(i % p.n_slots);
        const uint32_t ph = (uint32_t)((i / p.n_slots) & 1);
        const uint64_t r0 = (blockIdx.x + i * gridDim.x) * p.rows_per_slot;
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
    }
}

// one CTA: exact scores of the candidates in score_avx_group8's order, top-k by (score desc, id asc), or the fallback flag
__global__ void __launch_bounds__(1024) f32_prefilter_finish_kernel(const uint8_t* __restrict__ rows, uint32_t stride, uint32_t dim, const float* __restrict__ q,
                                                                    const uint32_t* __restrict__ cand, unsigned int* __restrict__ cnt, uint32_t top, uint32_t id_base,
                                                                    qb_scored_point* __restrict__ out, uint32_t* __restrict__ out_cnt, unsigned int* __restrict__ fallback,
                                                                    unsigned int* __restrict__ n_fallbacks) {
    __shared__ unsigned long long keys[PF_CAP];
    const unsigned int c = *cnt;
    __syncthreads();
    if (threadIdx.x == 0) *cnt = 0u;                              // ready for the next query on this context
    if (c > PF_CAP || c < top) {                                  // overflow, or a sample that could not give a threshold
        if (threadIdx.x == 0) { *fallback = 1u; atomicAdd(n_fallbacks, 1u); }
        return;
    }
    if (threadIdx.x == 0) *fallback = 0u;
    uint32_t n2 = 32;
    while (n2 < c) n2 <<= 1;
    const int t = threadIdx.x & 7;
    for (uint32_t i = threadIdx.x >> 3; i < n2; i += blockDim.x >> 3) {   // warp-uniform trip count: n2 is a multiple of 4 groups... (32 | n2)
        const bool valid = i < c;
        const uint32_t row = valid ? cand[i] : 0u;
        const float sc = qbs::score_avx_group8<qbs::M_DOT>(reinterpret_cast<const float*>(rows + (size_t)row * stride), q, dim, t);
        if (t == 0) keys[i] = valid ? qb_pack_key(sc, row + id_base) : 0ull;
    }
    __syncthreads();
    for (uint32_t k = 2; k <= n2; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool desc = ((i & k) == 0);
                    if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    if (threadIdx.x < top) {
        qb_scored_point sp;
        sp.idx = qb_key_id(keys[threadIdx.x]); sp.score = qb_key_score(keys[threadIdx.x]);
        out[threadIdx.x] = sp;
    }
    if (threadIdx.x == 0) *out_cnt = top;
}

template <int NCH>
qb_status launch_filter(PfParams& p, int sm_count, cudaStream_t stream) {
    const uint32_t kMaxSmem = 227 * 1024;
    uint32_t rps = (12288 / p.stride) & ~1u;      // 12-KB slots, sixteen of them: the shape the f32 stream kernel saturates HBM with
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
    dense_bf16_filter_kernel<NCH><<<grid, PF_THREADS, smem, stream>>>(p);
    QB_LAUNCHED();
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

}  // namespace

// Can this storage answer single-query top-k searches through the bf16 prefilter?  Builds the shadow plane on first use.
bool qb_f32_prefilter_usable(qb_storage* s, uint64_t n_rows, uint32_t top, cudaStream_t stream) {
    if (s->kind != QB_KIND_DENSE || s->dtype != QB_DT_F32) return false;
    if (s->distance != QB_DIST_DOT && s->distance != QB_DIST_COSINE) return false;       // the bound is on a dot product
    if (qb_opt().disable_prefilter || n_rows != s->count || n_rows < (1ull << 19) || top > 16 || s->dim < 32 || round_up_u64(s->dim, 8) > 1024) return false;
    if (qb_f32_shadow_ensure(s, stream) != QB_OK) { cudaGetLastError(); return false; }   // e.g. no room for the shadow plane: stay on the exact kernel
    return s->bf16_usable;
}

// d_q = preprocessed query; the exact top-`top` of the storage lands in d_out / d_out_cnt.  `a` = the scan arguments of the exact
// in-kernel-top-k path (emit.cand / final_out / done_counter set up by the caller); scratch = c->d_pf (see qb_f32_prefilter_scratch_bytes).
size_t qb_f32_prefilter_scratch_bytes() { return (size_t)PF_CAP * 4 + 16 * sizeof(qb_scored_point) + 256; }

qb_status qb_f32_prefilter_search(qb_storage* s, const QbScanArgs& a, uint32_t top, void* d_scratch, unsigned int* d_n_fallbacks, qb_scored_point* d_out, uint32_t* d_out_cnt,
                                  cudaEvent_t prof0, cudaEvent_t prof1, cudaStream_t stream) {
    uint8_t* sc = reinterpret_cast<uint8_t*>(d_scratch);
    // scratch: [0,16) cnt | [16,32) fallback flag | [48,64) sample count | [64, 64+16*8) sample top-k | candidates
    unsigned int* d_cnt = reinterpret_cast<unsigned int*>(sc);
    unsigned int* d_fallback = reinterpret_cast<unsigned int*>(sc + 16);
    uint32_t* d_samp_cnt = reinterpret_cast<uint32_t*>(sc + 48);
    qb_scored_point* d_samp = reinterpret_cast<qb_scored_point*>(sc + 64);
    uint32_t* d_cand = reinterpret_cast<uint32_t*>(sc + 64 + 16 * sizeof(qb_scored_point));
    const uint64_t n = s->count;
    // 1. exact top-k of a prefix
    uint64_t sample = std::min<uint64_t>(131072, std::max<uint64_t>(65536, (n / 64) & ~(uint64_t)3));
    QbScanArgs as = a;
    as.row_begin = 0; as.row_end = sample;
    as.emit.final_out = d_samp; as.emit.final_count = d_samp_cnt; as.emit.run_if = nullptr;
    uint64_t n_slots = 0;
    QB_TRY(qb_dense_f32_scan_localk(s, as, top, &n_slots, stream));
    QB_CHECK(n_slots != 0 && n_slots <= 4096, QB_ERR_CUDA, "prefilter: the sample scan did not launch (%llu slots)", (unsigned long long)n_slots);
    // 2. bf16 filter pass over everything
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
    // 3. exact scores + top-k of the survivors (or the fallback flag)
    f32_prefilter_finish_kernel<<<1, 1024, 0, stream>>>(reinterpret_cast<const uint8_t*>(s->d_rows), s->row_stride, s->dim, p.q, d_cand, d_cnt, top, a.emit.id_base, d_out,
                                                        d_out_cnt, d_fallback, d_n_fallbacks);
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

