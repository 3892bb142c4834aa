/*
 * train.c — CPU restatement of the reference's quantizer TRAINING steps.  TEST INFRASTRUCTURE (see oracle.c header).
 *
 *   qo_bq_vector_stats         VectorStatsBuilder::{add, build}   lib/quantization/src/vector_stats.rs:48-117 (Welford in f64, deterministic)
 *   qo_sq8_quantile_interval   find_quantile_interval             lib/quantization/src/quantile.rs:35-88, given the sampled vectors
 *   qo_kmeans_pq               find_centroids -> kmeans           encoded_vectors_pq.rs:342-407, kmeans.rs:9-167, given the sampled vectors
 *
 * The reference samples with an unseeded Permutor and re-seeds empty k-means clusters from rand::rng(): those two draws are inputs
 * here (the sample; `seed` through the same mixing function the device uses), everything else follows the Rust line by line —
 * including update_centroids' per-thread-range f64 partial sums merged in range order (`groups` = the reference's max_threads).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define API __attribute__((visibility("default")))

API void qo_bq_vector_stats(const float* rows, uint64_t count, uint32_t dim, float* mean_std, float* min_max) {
    double* mean = (double*)calloc(dim, sizeof(double)); double* m2 = (double*)calloc(dim, sizeof(double));
    for (uint32_t k = 0; k < dim; k++) if (min_max) { min_max[2 * k] = FLT_MAX; min_max[2 * k + 1] = -FLT_MAX; }
    for (uint64_t i = 0; i < count; i++) {
        const double cnt = (double)(i + 1);
        for (uint32_t k = 0; k < dim; k++) {
            const float v = rows[i * dim + k];
            if (min_max) { if (v < min_max[2 * k]) min_max[2 * k] = v; if (v > min_max[2 * k + 1]) min_max[2 * k + 1] = v; }
            const double x = (double)v, delta = x - mean[k];
            mean[k] += delta / cnt;
            m2[k] += delta * (x - mean[k]);
        }
    }
    for (uint32_t k = 0; k < dim; k++) { mean_std[2 * k] = (float)mean[k]; mean_std[2 * k + 1] = count > 1 ? (float)sqrt(m2[k] / (double)(count - 1)) : 0.0f; }
    free(mean); free(m2);
}

static int cmp_f32(const void* a, const void* b) { float x = *(const float*)a, y = *(const float*)b; return (x > y) - (x < y); }

/* returns 1 and (alpha, offset) = alpha_offset_from_min_max(min, max) of the values strictly between the two cut positions; 0 = None */
API int qo_sq8_quantile_interval(const float* sample, uint64_t n_vectors, uint32_t dim, float quantile, float* alpha, float* offset) {
    const uint64_t len = n_vectors * dim;
    if (quantile >= 1.0f || len < 4) return 0;
    uint64_t cut = (uint64_t)((float)n_vectors * (1.0f - quantile) / 2.0f);
    if (cut > (len - 1) / 2) cut = (len - 1) / 2;
    if (cut < 1) cut = 1;
    if (len - cut < cut + 1 + 2) return 0;
    float* s = (float*)malloc(len * 4);
    memcpy(s, sample, len * 4);
    qsort(s, len, 4, cmp_f32);          /* select_nth_unstable twice == these two order statistics of the fully sorted slice */
    const float mn = s[cut + 1], mx = s[len - cut - 1];
    free(s);
    *alpha = (mx - mn) / 127.0f; *offset = mn;
    return 1;
}

static uint64_t km_mix(uint64_t z) { z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

/* centroids_out: K full-dim vectors (Metadata.centroids).  Returns the largest number of Lloyd iterations any chunk ran. */
API uint32_t qo_kmeans_pq(const float* sample, uint32_t n, uint32_t dim, uint32_t chunk, uint32_t K, uint32_t max_iter, float accuracy, uint32_t groups, uint64_t seed,
                          float* centroids_out) {
    memset(centroids_out, 0, (size_t)K * dim * 4);
    if (n <= K) { memcpy(centroids_out, sample, (size_t)n * dim * 4); return 0; }
    if (groups > n) groups = n;
    const uint32_t m = (dim + chunk - 1) / chunk;
    uint32_t worst = 0;
    uint32_t* idx = (uint32_t*)malloc((size_t)n * 4);
    for (uint32_t j = 0; j < m; j++) {
        const uint32_t s = j * chunk, e = (j + 1) * chunk < dim ? (j + 1) * chunk : dim, clen = e - s;
        float* cent = (float*)malloc((size_t)K * clen * 4);
        double* acc = (double*)malloc((size_t)K * clen * 8); double* part = (double*)malloc((size_t)K * clen * 8);
        uint64_t* cnt = (uint64_t*)malloc((size_t)K * 8);
        for (uint32_t c = 0; c < K; c++) memcpy(cent + (size_t)c * clen, sample + (size_t)c * dim + s, clen * 4);   /* data[0..K*dim] */
        for (uint32_t it = 0; it < max_iter; it++) {
            /* update_indexes */
            for (uint32_t i = 0; i < n; i++) {
                const float* v = sample + (size_t)i * dim + s;
                float best = FLT_MAX; uint32_t bi = 0;
                for (uint32_t c = 0; c < K; c++) {
                    float d2 = -0.0f;
                    for (uint32_t k = 0; k < clen; k++) { float d = v[k] - cent[(size_t)c * clen + k]; d2 += d * d; }
                    if (d2 < best) { best = d2; bi = c; }
                }
                idx[i] = bi;
            }
            /* update_centroids: per-range f64 partial sums, merged in range order */
            memset(acc, 0, (size_t)K * clen * 8); memset(cnt, 0, (size_t)K * 8);
            const uint32_t cs = n / groups;
            for (uint32_t g = 0; g < groups; g++) {
                const uint32_t b = cs * g, en = (g + 1 == groups) ? n : cs * (g + 1);
                memset(part, 0, (size_t)K * clen * 8);
                for (uint32_t i = b; i < en; i++) {
                    cnt[idx[i]]++;
                    for (uint32_t k = 0; k < clen; k++) part[(size_t)idx[i] * clen + k] += (double)sample[(size_t)i * dim + s + k];
                }
                for (size_t t = 0; t < (size_t)K * clen; t++) acc[t] += part[t];
            }
            for (uint32_t c = 0; c < K; c++) {
                if (cnt[c] == 0) {
                    const uint32_t di = (uint32_t)(km_mix(seed ^ km_mix(((uint64_t)it << 40) ^ ((uint64_t)j << 20) ^ c)) % n);
                    for (uint32_t k = 0; k < clen; k++) acc[(size_t)c * clen + k] = (double)sample[(size_t)di * dim + s + k];
                } else {
                    const double count = (double)cnt[c];
                    for (uint32_t k = 0; k < clen; k++) acc[(size_t)c * clen + k] /= count;
                }
            }
            float diff = -0.0f;
            for (size_t t = 0; t < (size_t)K * clen; t++) { const float ca = (float)acc[t]; diff += fabsf(cent[t] - ca); cent[t] = ca; }
            if (it + 1 > worst) worst = it + 1;
            if (diff < accuracy) break;
        }
        for (uint32_t c = 0; c < K; c++) memcpy(centroids_out + (size_t)c * dim + s, cent + (size_t)c * clen, clen * 4);
        free(cent); free(acc); free(part); free(cnt);
    }
    free(idx);
    return worst;
}
