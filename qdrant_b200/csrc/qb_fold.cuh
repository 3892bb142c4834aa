// qb_fold.cuh — Query::score_by of the custom queries (recommend / discover / context / feedback) as ONE device function over an
// accessor `sim(e)` = the candidate's similarity to example e, so that the matrix fold (qb_custom.cu) and the fold fused into the
// streaming scan (qb_dense.cu) share the reference's f32 operation order:
//   RecoBestScore  query/reco_query.rs:64-90     max over positives / negatives (total_cmp), scaled_fast_sigmoid of the winner
//   RecoSumScores  query/reco_query.rs:116-133   sequential f32 sums, positives minus negatives
//   Discover       query/discover_query.rs:16-76 rank = sum of total_cmp(positive, negative) per pair, + sigmoid(target)
//   Context        query/context_query.rs:52-62,111-119   sum over pairs of fast_sigmoid(min(p - n - EPSILON, 0))
//   Feedback       query/feedback_query.rs:204-226 a * sim(target) + sum over pairs of partial_computation * (sim(pos) - sim(neg))
// fast_sigmoid = x / (1 + |x|), scaled_fast_sigmoid = 0.5 * (fast_sigmoid(x) + 1)   (lib/common/common/src/math.rs:7-18)
#pragma once
#include "qb_common.cuh"

namespace qbf {

__device__ __forceinline__ int total_cmp(float a, float b) {  // f32::total_cmp as -1 / 0 / 1
    int x = __float_as_int(a), y = __float_as_int(b);
    x ^= (int)((unsigned int)(x >> 31) >> 1);
    y ^= (int)((unsigned int)(y >> 31) >> 1);
    return (x > y) - (x < y);
}
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdiv_rn(x, __fadd_rn(1.0f, fabsf(x))); }
__device__ __forceinline__ float scaled_fast_sigmoid(float x) { return __fmul_rn(0.5f, __fadd_rn(fast_sigmoid(x), 1.0f)); }

template <class Sim>
__device__ __forceinline__ float fold(int kind, uint32_t n_a, uint32_t n_b, const float* __restrict__ coef, Sim sim) {
    switch (kind) {
        case QB_QUERY_FEEDBACK_NAIVE: {  // coef = [a, partial_computation of pair 0, 1, ...]; `score += partial * delta` is a multiply then an add in Rust
            float score = __fmul_rn(coef[0], sim(0));
            for (uint32_t e = 0; e < n_a; ++e) {
                const float delta = __fsub_rn(sim(1 + 2 * e), sim(2 + 2 * e));
                score = __fadd_rn(score, __fmul_rn(coef[1 + e], delta));
            }
            return score;
        }
        case QB_QUERY_RECO_BEST_SCORE: {
            float max_p = __int_as_float(0xff800000), max_n = __int_as_float(0xff800000);
            for (uint32_t e = 0; e < n_a; ++e) { const float s = sim(e); if (total_cmp(s, max_p) > 0) max_p = s; }
            for (uint32_t e = 0; e < n_b; ++e) { const float s = sim(n_a + e); if (total_cmp(s, max_n) > 0) max_n = s; }
            return (max_p > max_n) ? scaled_fast_sigmoid(max_p) : -scaled_fast_sigmoid(max_n);
        }
        case QB_QUERY_RECO_SUM_SCORES: {
            float p = 0.0f, n = 0.0f;
            for (uint32_t e = 0; e < n_a; ++e) p = __fadd_rn(p, sim(e));
            for (uint32_t e = 0; e < n_b; ++e) n = __fadd_rn(n, sim(n_a + e));
            return __fsub_rn(p, n);
        }
        case QB_QUERY_DISCOVER: {
            int rank = 0;
            for (uint32_t e = 0; e < n_a; ++e) rank += total_cmp(sim(1 + 2 * e), sim(2 + 2 * e));
            return __fadd_rn((float)rank, scaled_fast_sigmoid(sim(0)));
        }
        default: {  // QB_QUERY_CONTEXT
            float sum = 0.0f;
            for (uint32_t e = 0; e < n_a; ++e) {
                const float d = __fsub_rn(__fsub_rn(sim(2 * e), sim(2 * e + 1)), 1.1920929e-7f);
                sum = __fadd_rn(sum, fast_sigmoid(fminf(d, 0.0f)));
            }
            return sum;
        }
    }
}

}  // namespace qbf
