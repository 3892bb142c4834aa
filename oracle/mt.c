/*
 * mt.c — the reference's CPU path run the way the reference runs it on a many-core host: one blocking task per segment
 * (lib/collection/src/collection_manager/segments_searcher.rs:255), each task the BatchFilteredSearcher::peek_top_iter loop
 * of oracle.c (point_scorer.rs:423-472), lists merged on the host (BatchResultAggregator, search_result_aggregator.rs:50-117).
 *
 * TEST INFRASTRUCTURE (see oracle.c header): bench.py's cpu_baseline / --impl reference legs only.
 *
 * A pool of T threads pinned to T CPUs; every thread OWNS one contiguous segment of the data set, allocated and first-touched
 * by that thread (so the pages sit on the thread's NUMA node) and scanned by that thread.  Nothing is dispatched from Python
 * per segment: one call = one condition-variable broadcast + T scans + one T-way merge.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#define API __attribute__((visibility("default")))

typedef struct { uint32_t idx; float score; } qo_scored;
typedef struct { uint32_t dim, actual_dim; float alpha, offset, multiplier; int32_t distance_type, invert; } qo_sq8_meta;

void qo_scan_f32(int distance, const float* base, uint64_t row_begin, uint64_t row_end, uint32_t dim, const float* queries_preprocessed, uint32_t n_queries,
                 uint32_t top, const uint64_t* deleted, qo_scored* out, uint32_t* out_counts);
void qo_scan_sq8(const qo_sq8_meta* m, const uint8_t* rows, uint64_t row_begin, uint64_t row_end, const uint8_t* q_codes, const float* q_offs,
                 uint32_t n_queries, uint32_t top, const uint64_t* deleted, qo_scored* out, uint32_t* out_counts);
void qo_scan_pq(const uint8_t* codes, uint64_t row_begin, uint64_t row_end, uint32_t m_chunks, uint32_t n_centroids, const float* luts, uint32_t n_queries,
                uint32_t top, const uint64_t* deleted, qo_scored* out, uint32_t* out_counts);
void qo_preprocess_f32(int distance, const float* v, float* out, size_t n);
uint32_t qo_topk(const uint32_t* ids, const float* scores, uint64_t n, uint32_t top, qo_scored* out);

enum { JOB_NONE = 0, JOB_GEN_F32, JOB_COPY_F32, JOB_SCAN_F32, JOB_SCAN_SQ8, JOB_SCAN_PQ, JOB_EXIT };

typedef struct qo_pool qo_pool;
typedef struct {
    qo_pool* pool; uint32_t tid; pthread_t th;
    /* the f32 segment this thread owns */
    float* seg; uint64_t seg_begin, seg_rows; size_t seg_bytes;
    qo_scored* res; uint32_t* cnt; size_t res_cap;
} worker_t;

struct qo_pool {
    uint32_t T;
    worker_t* w;
    pthread_mutex_t mu; pthread_cond_t cv_job, cv_done;
    uint64_t gen; uint32_t done; int job;
    /* job arguments */
    uint64_t rows; uint32_t dim; int distance; uint64_t seed; const float* src;
    const float* queries; uint32_t nq, top;
    const qo_sq8_meta* sq_meta; const uint8_t* sq_rows; const uint8_t* q_codes; const float* q_offs;
    const uint8_t* pq_codes; uint32_t pq_m, pq_centroids; const float* luts;
};

static void pin_to_cpu(uint32_t tid) {
    cpu_set_t all, one;
    if (sched_getaffinity(0, sizeof(all), &all) != 0) return;
    int ncpu = CPU_COUNT(&all), k = (int)(tid % (uint32_t)(ncpu > 0 ? ncpu : 1)), seen = 0;
    for (int c = 0; c < CPU_SETSIZE; c++)
        if (CPU_ISSET(c, &all)) { if (seen++ == k) { CPU_ZERO(&one); CPU_SET(c, &one); pthread_setaffinity_np(pthread_self(), sizeof(one), &one); return; } }
}

static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t xoshiro(uint64_t s[4]) {
    const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
}
static uint64_t splitmix(uint64_t* s) { uint64_t z = (*s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

static void seg_range(uint64_t rows, uint32_t T, uint32_t t, uint64_t* b, uint64_t* e) {   /* same split as qdrant_b200.sharded.shard_ranges */
    const uint64_t base = rows / T, rem = rows % T;
    *b = (uint64_t)t * base + (t < rem ? t : rem);
    *e = *b + base + (t < rem ? 1 : 0);
}

static void ensure_res(worker_t* w, uint32_t nq, uint32_t top) {
    const size_t need = (size_t)nq * (top ? top : 1);
    if (need > w->res_cap) { free(w->res); free(w->cnt); w->res = (qo_scored*)malloc(need * sizeof(qo_scored)); w->cnt = (uint32_t*)malloc(((size_t)nq + 1) * 4); w->res_cap = need; }
}

static void run_job(worker_t* w) {
    qo_pool* p = w->pool;
    uint64_t b, e;
    switch (p->job) {
        case JOB_GEN_F32:
        case JOB_COPY_F32: {
            seg_range(p->rows, p->T, w->tid, &b, &e);
            if (w->seg) munmap(w->seg, w->seg_bytes);
            w->seg_begin = b; w->seg_rows = e - b;
            w->seg_bytes = ((size_t)(e - b) * p->dim * 4 + 4095) & ~(size_t)4095;
            if (w->seg_bytes == 0) w->seg_bytes = 4096;
            w->seg = (float*)mmap(NULL, w->seg_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (w->seg == MAP_FAILED) { w->seg = NULL; w->seg_rows = 0; break; }
            if (p->job == JOB_COPY_F32) { memcpy(w->seg, p->src + b * p->dim, (size_t)(e - b) * p->dim * 4); break; }
            /* standard-normal rows (Box-Muller on xoshiro256**), then Distance::preprocess_vector like an insert would */
            uint64_t sm = p->seed ^ (0xD1B54A32D192ED03ull * (w->tid + 1)), s[4];
            for (int i = 0; i < 4; i++) s[i] = splitmix(&sm);
            for (uint64_t r = 0; r < e - b; r++) {
                float* row = w->seg + r * p->dim;
                for (uint32_t i = 0; i < p->dim; i += 2) {
                    const float u1 = ((xoshiro(s) >> 40) + 1.0f) * (1.0f / 16777217.0f), u2 = (xoshiro(s) >> 40) * (1.0f / 16777216.0f);
                    const float rad = sqrtf(-2.0f * logf(u1)), ang = 6.2831853f * u2;
                    row[i] = rad * cosf(ang);
                    if (i + 1 < p->dim) row[i + 1] = rad * sinf(ang);
                }
                qo_preprocess_f32(p->distance, row, row, p->dim);
            }
            break;
        }
        case JOB_SCAN_F32: {
            ensure_res(w, p->nq, p->top);
            if (!w->seg) { memset(w->cnt, 0, (size_t)p->nq * 4); break; }
            qo_scan_f32(p->distance, w->seg, 0, w->seg_rows, p->dim, p->queries, p->nq, p->top, NULL, w->res, w->cnt);
            for (uint32_t q = 0; q < p->nq; q++) for (uint32_t i = 0; i < w->cnt[q]; i++) w->res[(size_t)q * p->top + i].idx += (uint32_t)w->seg_begin;
            break;
        }
        case JOB_SCAN_SQ8:
            ensure_res(w, p->nq, p->top);
            seg_range(p->rows, p->T, w->tid, &b, &e);
            qo_scan_sq8(p->sq_meta, p->sq_rows, b, e, p->q_codes, p->q_offs, p->nq, p->top, NULL, w->res, w->cnt);
            break;
        case JOB_SCAN_PQ:
            ensure_res(w, p->nq, p->top);
            seg_range(p->rows, p->T, w->tid, &b, &e);
            qo_scan_pq(p->pq_codes, b, e, p->pq_m, p->pq_centroids, p->luts, p->nq, p->top, NULL, w->res, w->cnt);
            break;
        default: break;
    }
}

static void* worker_main(void* ap) {
    worker_t* w = (worker_t*)ap;
    qo_pool* p = w->pool;
    pin_to_cpu(w->tid);
    uint64_t seen = 0;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        while (p->gen == seen) pthread_cond_wait(&p->cv_job, &p->mu);
        seen = p->gen;
        const int job = p->job;
        pthread_mutex_unlock(&p->mu);
        if (job == JOB_EXIT) return NULL;
        run_job(w);
        pthread_mutex_lock(&p->mu);
        if (++p->done == p->T) pthread_cond_signal(&p->cv_done);
        pthread_mutex_unlock(&p->mu);
    }
}

static void dispatch(qo_pool* p, int job) {
    pthread_mutex_lock(&p->mu);
    p->job = job; p->done = 0; p->gen++;
    pthread_cond_broadcast(&p->cv_job);
    while (p->done != p->T) pthread_cond_wait(&p->cv_done, &p->mu);
    pthread_mutex_unlock(&p->mu);
}

API qo_pool* qo_pool_create(uint32_t threads) {
    if (threads < 1) threads = 1;
    qo_pool* p = (qo_pool*)calloc(1, sizeof(qo_pool));
    p->T = threads;
    p->w = (worker_t*)calloc(threads, sizeof(worker_t));
    pthread_mutex_init(&p->mu, NULL); pthread_cond_init(&p->cv_job, NULL); pthread_cond_init(&p->cv_done, NULL);
    for (uint32_t t = 0; t < threads; t++) { p->w[t].pool = p; p->w[t].tid = t; pthread_create(&p->w[t].th, NULL, worker_main, &p->w[t]); }
    return p;
}

API void qo_pool_destroy(qo_pool* p) {
    if (!p) return;
    pthread_mutex_lock(&p->mu);
    p->job = JOB_EXIT; p->gen++;
    pthread_cond_broadcast(&p->cv_job);
    pthread_mutex_unlock(&p->mu);
    for (uint32_t t = 0; t < p->T; t++) {
        pthread_join(p->w[t].th, NULL);
        if (p->w[t].seg) munmap(p->w[t].seg, p->w[t].seg_bytes);
        free(p->w[t].res); free(p->w[t].cnt);
    }
    free(p->w); free(p);
}

/* every thread allocates, first-touches and fills its own segment of a `rows` x `dim` f32 data set.
 * src != NULL: copy of those rows (the very rows the GPU scans); else seeded standard-normal rows, Metric::preprocess applied. */
API int qo_pool_load_f32(qo_pool* p, uint64_t rows, uint32_t dim, int distance, uint64_t seed, const float* src) {
    p->rows = rows; p->dim = dim; p->distance = distance; p->seed = seed; p->src = src;
    dispatch(p, src ? JOB_COPY_F32 : JOB_GEN_F32);
    for (uint32_t t = 0; t < p->T; t++) if (!p->w[t].seg) return -1;
    return 0;
}

/* k-way merge of the per-segment lists: heap push of every (score, id) like BatchResultAggregator */
static void merge(qo_pool* p, uint32_t nq, uint32_t top, qo_scored* out, uint32_t* out_counts) {
    const size_t cap = (size_t)p->T * top;
    uint32_t* ids = (uint32_t*)malloc(cap * 4); float* sc = (float*)malloc(cap * 4);
    for (uint32_t q = 0; q < nq; q++) {
        size_t n = 0;
        for (uint32_t t = 0; t < p->T; t++) for (uint32_t i = 0; i < p->w[t].cnt[q]; i++) { ids[n] = p->w[t].res[(size_t)q * top + i].idx; sc[n] = p->w[t].res[(size_t)q * top + i].score; n++; }
        out_counts[q] = qo_topk(ids, sc, n, top, out + (size_t)q * top);
    }
    free(ids); free(sc);
}

API void qo_pool_scan_f32(qo_pool* p, const float* queries_pre, uint32_t nq, uint32_t top, qo_scored* out, uint32_t* out_counts) {
    p->queries = queries_pre; p->nq = nq; p->top = top;
    dispatch(p, JOB_SCAN_F32);
    merge(p, nq, top, out, out_counts);
}

API void qo_pool_scan_sq8(qo_pool* p, const qo_sq8_meta* m, const uint8_t* rows, uint64_t n_rows, const uint8_t* q_codes, const float* q_offs, uint32_t nq,
                          uint32_t top, qo_scored* out, uint32_t* out_counts) {
    p->sq_meta = m; p->sq_rows = rows; p->rows = n_rows; p->q_codes = q_codes; p->q_offs = q_offs; p->nq = nq; p->top = top;
    dispatch(p, JOB_SCAN_SQ8);
    merge(p, nq, top, out, out_counts);
}

API void qo_pool_scan_pq(qo_pool* p, const uint8_t* codes, uint64_t n_rows, uint32_t m_chunks, uint32_t n_centroids, const float* luts, uint32_t nq,
                         uint32_t top, qo_scored* out, uint32_t* out_counts) {
    p->pq_codes = codes; p->rows = n_rows; p->pq_m = m_chunks; p->pq_centroids = n_centroids; p->luts = luts; p->nq = nq; p->top = top;
    dispatch(p, JOB_SCAN_PQ);
    merge(p, nq, top, out, out_counts);
}

API uint32_t qo_pool_threads(const qo_pool* p) { return p->T; }
