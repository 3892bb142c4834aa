// qb_score.cuh — device-side per-pair scoring primitives shared by the scan / gather kernels (qb_dense.cu, qb_quant.cu) and the
// device-resident HNSW traversal (qb_hnsw.cu).  Each function restates one reference routine with its accumulation order; see the
// headers of qb_dense.cu / qb_quant.cu for the bit-exactness argument and the reference citations.
#pragma once
#include "qb_common.cuh"

namespace qbs {

enum { M_DOT = 0, M_EUCLID = 1, M_MANHATTAN = 2 };

__device__ __forceinline__ float4 shfl_xor4(float4 v, int m) {
    v.x = __shfl_xor_sync(0xFFFFFFFFu, v.x, m);
    v.y = __shfl_xor_sync(0xFFFFFFFFu, v.y, m);
    v.z = __shfl_xor_sync(0xFFFFFFFFu, v.z, m);
    v.w = __shfl_xor_sync(0xFFFFFFFFu, v.w, m);
    return v;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
    return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}

template <int METRIC>
__device__ __forceinline__ float elem_step(float q, float v, float acc) {
    if (METRIC == M_DOT) return __fmaf_rn(q, v, acc);
    float d = __fsub_rn(q, v);
    if (METRIC == M_EUCLID) return __fmaf_rn(d, d, acc);
    return __fadd_rn(fabsf(d), acc);
}
template <int METRIC>
__device__ __forceinline__ float tail_step(float q, float v, float r) {
    if (METRIC == M_DOT) return __fadd_rn(r, __fmul_rn(q, v));
    float d = __fsub_rn(q, v);
    if (METRIC == M_EUCLID) return __fadd_rn(r, __fmul_rn(d, d));
    return __fadd_rn(r, fabsf(d));
}

// AVX tier (dim >= 32).  `row` and `qry` are 16-B aligned; t = lane & 7.  All 8 lanes of the group return r.
template <int METRIC>
__device__ __forceinline__ float score_avx_group8(const float* __restrict__ row, const float* __restrict__ qry, uint32_t dim, int t) {
    const uint32_t nblk = dim >> 5;
    const float4* r4 = reinterpret_cast<const float4*>(row) + t;
    const float4* q4 = reinterpret_cast<const float4*>(qry) + t;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (uint32_t b = 0; b < nblk; ++b) {
        float4 v = r4[b * 8];
        float4 q = q4[b * 8];
        acc.x = elem_step<METRIC>(q.x, v.x, acc.x);
        acc.y = elem_step<METRIC>(q.y, v.y, acc.y);
        acc.z = elem_step<METRIC>(q.z, v.z, acc.z);
        acc.w = elem_step<METRIC>(q.w, v.w, acc.w);
    }
    acc = add4(acc, shfl_xor4(acc, 2));  // (P0+P1), (P2+P3)        four_way_hsum, simple_avx.rs:21-28
    acc = add4(acc, shfl_xor4(acc, 4));  // (P0+P1)+(P2+P3) = T[l]
    acc = add4(acc, shfl_xor4(acc, 1));  // T[i+4]+T[i] = L[i]       hsum256_ps_avx, simple_avx.rs:10-16
    float r = __fadd_rn(__fadd_rn(acc.x, acc.y), __fadd_rn(acc.z, acc.w));
    for (uint32_t i = nblk << 5; i < dim; ++i) r = tail_step<METRIC>(qry[i], row[i], r);
    return (METRIC == M_DOT) ? r : -r;
}

// NQ queries against ONE row: the row's float4 of every 32-block is loaded once and FMA-ed into NQ independent accumulator sets, so a
// batch (or the examples of a custom query) costs one pass over the rows and one row read from shared memory instead of NQ.  Each
// (query, lane) chain is the very chain score_avx_group8 runs, so every result is bit-identical to the single-query function.
// qry + q * q_stride_f = query q (16-B aligned).
template <int METRIC, int NQ>
__device__ __forceinline__ void score_avx_group8_multi(const float* __restrict__ row, const float* __restrict__ qry, uint32_t q_stride_f, uint32_t dim, int t,
                                                       float (&out)[NQ]) {
    const uint32_t nblk = dim >> 5;
    const float4* r4 = reinterpret_cast<const float4*>(row) + t;
    float4 acc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (uint32_t b = 0; b < nblk; ++b) {
        const float4 v = r4[b * 8];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const float4 w = (reinterpret_cast<const float4*>(qry + (size_t)q * q_stride_f) + t)[b * 8];
            acc[q].x = elem_step<METRIC>(w.x, v.x, acc[q].x);
            acc[q].y = elem_step<METRIC>(w.y, v.y, acc[q].y);
            acc[q].z = elem_step<METRIC>(w.z, v.z, acc[q].z);
            acc[q].w = elem_step<METRIC>(w.w, v.w, acc[q].w);
        }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        float4 a = acc[q];
        a = add4(a, shfl_xor4(a, 2));
        a = add4(a, shfl_xor4(a, 4));
        a = add4(a, shfl_xor4(a, 1));
        float r = __fadd_rn(__fadd_rn(a.x, a.y), __fadd_rn(a.z, a.w));
        const float* qq = qry + (size_t)q * q_stride_f;
        for (uint32_t i = nblk << 5; i < dim; ++i) r = tail_step<METRIC>(qq[i], row[i], r);
        out[q] = (METRIC == M_DOT) ? r : -r;
    }
}

// SSE tier (16 <= dim < 32, one 16-block, unfused mul+add) and scalar tier (dim < 16); one thread per pair.
template <int METRIC>
__device__ __forceinline__ float score_small(const float* __restrict__ row, const float* __restrict__ qry, uint32_t dim) {
    float r;
    uint32_t start;
    if (dim >= 16) {
        float p[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float q = qry[i], v = row[i];
            if (METRIC == M_DOT) p[i] = __fadd_rn(__fmul_rn(q, v), 0.0f);
            else {
                float d = __fsub_rn(q, v);
                p[i] = (METRIC == M_EUCLID) ? __fadd_rn(__fmul_rn(d, d), 0.0f) : __fadd_rn(fabsf(d), 0.0f);
            }
        }
        float h[4];
#pragma unroll
        for (int a = 0; a < 4; ++a)  // hsum128_ps_sse: (x0+x2)+(x1+x3), simple_sse.rs:13-17
            h[a] = __fadd_rn(__fadd_rn(p[4 * a], p[4 * a + 2]), __fadd_rn(p[4 * a + 1], p[4 * a + 3]));
        r = __fadd_rn(__fadd_rn(__fadd_rn(h[0], h[1]), h[2]), h[3]);
        start = 16;
    } else {
        r = -0.0f;  // Rust's f32 Sum folds from -0.0
        start = 0;
    }
    for (uint32_t i = start; i < dim; ++i) r = tail_step<METRIC>(qry[i], row[i], r);
    return (METRIC == M_DOT) ? r : -r;
}


// SQ8 (EncodedVectorsU8): raw integer score of one stored code row against one query code, 8 lanes per row, 16 B per lane per
// step; every lane of the group returns the value.  l1 = Manhattan on codes (impl_score_l1_avx, cpp/avx2.c:65-122); LANEX =
// the 8-lane partition + HSUM256_PS tree of impl_score_dot_avx (cpp/avx2.c:7-63) for totals that can leave the f32-exact window.
template <bool LANEX>
__device__ __forceinline__ float sq8_raw_group8(const uint4* __restrict__ rp, const uint4* __restrict__ qp, uint32_t n_chunks, int t, int l1) {
    float score;
    if (l1) {
        unsigned int acc = 0;
        for (uint32_t c = t; c < n_chunks; c += 8) {
            uint4 v = __ldg(rp + c), w = qp[c];
            acc += __vsadu4(v.x, w.x) + __vsadu4(v.y, w.y) + __vsadu4(v.z, w.z) + __vsadu4(v.w, w.w);
        }
        acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 1);
        acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 2);
        acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 4);
        score = (float)acc;  // impl_score_l1_avx returns (float)sum, cpp/avx2.c:117-121
    } else if (!LANEX) {
        int acc = 0;
        for (uint32_t c = t; c < n_chunks; c += 8) {
            uint4 v = __ldg(rp + c), w = qp[c];
            acc = __dp4a((int)v.x, (int)w.x, acc);
            acc = __dp4a((int)v.y, (int)w.y, acc);
            acc = __dp4a((int)v.z, (int)w.z, acc);
            acc = __dp4a((int)v.w, (int)w.w, acc);
        }
        acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 1);
        acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 2);
        acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 4);
        score = (float)acc;  // exact: total < 2^24
    } else {
        // lane partition of impl_score_dot_avx: byte pair j of every 16-B chunk accumulates into i32 lane j
        int ln[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (uint32_t c = t; c < n_chunks; c += 8) {
            uint4 v = __ldg(rp + c), w = qp[c];
            ln[0] = __dp4a((int)v.x, (int)(w.x & 0x0000FFFFu), ln[0]); ln[1] = __dp4a((int)v.x, (int)(w.x & 0xFFFF0000u), ln[1]);
            ln[2] = __dp4a((int)v.y, (int)(w.y & 0x0000FFFFu), ln[2]); ln[3] = __dp4a((int)v.y, (int)(w.y & 0xFFFF0000u), ln[3]);
            ln[4] = __dp4a((int)v.z, (int)(w.z & 0x0000FFFFu), ln[4]); ln[5] = __dp4a((int)v.z, (int)(w.z & 0xFFFF0000u), ln[5]);
            ln[6] = __dp4a((int)v.w, (int)(w.w & 0x0000FFFFu), ln[6]); ln[7] = __dp4a((int)v.w, (int)(w.w & 0xFFFF0000u), ln[7]);
        }
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            ln[l] += __shfl_xor_sync(0xFFFFFFFFu, ln[l], 1);
            ln[l] += __shfl_xor_sync(0xFFFFFFFFu, ln[l], 2);
            ln[l] += __shfl_xor_sync(0xFFFFFFFFu, ln[l], 4);
        }
        // HSUM256_PS (cpp/avx2.c:7-14): ((l0+l4)+(l2+l6)) + ((l1+l5)+(l3+l7))
        float x0 = __fadd_rn((float)ln[4], (float)ln[0]), x1 = __fadd_rn((float)ln[5], (float)ln[1]);
        float x2 = __fadd_rn((float)ln[6], (float)ln[2]), x3 = __fadd_rn((float)ln[7], (float)ln[3]);
        score = __fadd_rn(__fadd_rn(x0, x2), __fadd_rn(x1, x3));
    }
    return score;
}

}  // namespace qbs
