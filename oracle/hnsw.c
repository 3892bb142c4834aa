/*
 * hnsw.c — from-spec CPU HNSW (build + search) used ONLY as the checker / CPU baseline of the config-#5 harness.
 *
 * TEST INFRASTRUCTURE (see oracle.c header).  The reference keeps HNSW traversal on the CPU and calls the scorer
 * through FilteredScorer::score_points once per hop (lib/segment/src/index/hnsw_index/graph_layers.rs:108-148,
 * 247-316, 530-561; search_context.rs:8-41).  This file restates that traversal with the scorer behind a callback,
 * so the same traversal can be driven by the CPU oracle scorer and by the GPU RawScorer (qb_score_points), and it is
 * the reference list the device-resident traversal (qb_hnsw_search_batch) is compared with.
 *
 * The builder follows graph_layers_builder.rs:388-566 and links_container.rs:47-71,139-... (heuristic on).  Like the
 * reference (hnsw/build.rs:285-355: the first SINGLE_THREADED_HNSW_BUILD_THRESHOLD = 256 points serially, the rest on
 * a thread pool with per-point link locks) it can build with many threads; such graphs are not deterministic — in the
 * reference either — so parity tests always compare two traversals of the SAME graph.  The level RNG is ours: the
 * reference's rand stream is not reproducible without Rust.
 *
 * qo_hnsw_export_plain writes the graph in the reference's plain `links.bin` layout (graph_links/header.rs:9-20,
 * graph_links/serializer.rs:53-200): HeaderPlain, level offsets, reindex, neighbors, padding, offsets.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

float qo_similarity_f32(int distance, const float* q, const float* v, size_t n);

typedef struct { uint32_t idx; float score; } sp_t;
typedef void (*qo_score_cb)(void* user, const uint32_t* ids, uint32_t n, float* scores);

typedef struct {
    uint32_t n, dim, m, m0, ef_construct;
    int distance;
    const float* base;
    uint8_t* level;          /* per point */
    uint32_t*** links;       /* links[p][lvl] -> array: [count, ids...] with capacity level_m(lvl) */
    atomic_uchar* lock;      /* per-point spin lock (multi-threaded build only) */
    int locking;
    uint32_t entry, entry_level;
    int has_entry;
    pthread_mutex_t entry_mu;
    _Atomic uint64_t n_score_calls, n_scored;
} hnsw_t;

/* per-thread traversal state (the reference's VisitedPool hands out one list per search) */
typedef struct { sp_t* d; size_t len, cap; } heap_t;
typedef struct { sp_t* d; size_t len, cap; } flpq_t;
typedef struct {
    uint32_t* visited; uint32_t stamp;
    flpq_t nearest; heap_t cand;
    uint64_t calls, scored;
} tctx_t;

static inline uint32_t level_m(const hnsw_t* h, uint32_t lvl) { return lvl == 0 ? h->m0 : h->m; }  /* HnswM::level_m, mod.rs:34-40 */

static inline void plock(hnsw_t* h, uint32_t p) {
    if (!h->locking) return;
    unsigned char e = 0;
    while (!atomic_compare_exchange_weak_explicit(&h->lock[p], &e, 1, memory_order_acquire, memory_order_relaxed)) { e = 0; __builtin_ia32_pause(); }
}
static inline void punlock(hnsw_t* h, uint32_t p) { if (h->locking) atomic_store_explicit(&h->lock[p], 0, memory_order_release); }
/* copy of a point's links at a level (under its lock while other threads may be rewriting them) */
static inline uint32_t read_links(hnsw_t* h, uint32_t p, uint32_t lvl, uint32_t* out) {
    plock(h, p);
    const uint32_t* l = h->links[p][lvl];
    uint32_t n = l[0];
    memcpy(out, l + 1, n * sizeof(uint32_t));
    punlock(h, p);
    return n;
}

/* ---- heaps ------------------------------------------------------------------------------------------------- */
static void heap_reserve(heap_t* h, size_t n) { if (n > h->cap) { h->cap = n * 2 + 16; h->d = (sp_t*)realloc(h->d, h->cap * sizeof(sp_t)); } }
/* max-heap on score (BinaryHeap<ScoredPointOffset>) */
static void maxheap_push(heap_t* h, sp_t v) {
    heap_reserve(h, h->len + 1);
    size_t i = h->len++;
    while (i > 0) { size_t p = (i - 1) / 2; if (!(v.score > h->d[p].score)) break; h->d[i] = h->d[p]; i = p; }
    h->d[i] = v;
}
static sp_t maxheap_pop(heap_t* h) {
    sp_t top = h->d[0], v = h->d[--h->len];
    size_t i = 0;
    for (;;) {
        size_t c = 2 * i + 1;
        if (c >= h->len) break;
        if (c + 1 < h->len && h->d[c + 1].score > h->d[c].score) c++;
        if (!(h->d[c].score > v.score)) break;
        h->d[i] = h->d[c]; i = c;
    }
    if (h->len) h->d[i] = v;
    return top;
}
/* FixedLengthPriorityQueue: min-heap of the `cap` best; push returns 1 if the pushed element was kept */
static void minheap_up(sp_t* d, size_t i) { sp_t v = d[i]; while (i > 0) { size_t p = (i - 1) / 2; if (!(v.score < d[p].score)) break; d[i] = d[p]; i = p; } d[i] = v; }
static void minheap_down(sp_t* d, size_t len, size_t i) {
    sp_t v = d[i];
    for (;;) {
        size_t c = 2 * i + 1;
        if (c >= len) break;
        if (c + 1 < len && d[c + 1].score < d[c].score) c++;
        if (!(d[c].score < v.score)) break;
        d[i] = d[c]; i = c;
    }
    d[i] = v;
}
static int flpq_push(flpq_t* q, sp_t v) {  /* fixed_length_priority_queue.rs:47-59 */
    if (q->len < q->cap) { q->d[q->len] = v; minheap_up(q->d, q->len); q->len++; return 1; }
    if (q->d[0].score < v.score) { q->d[0] = v; minheap_down(q->d, q->len, 0); return 1; }
    return 0;
}
static int cmp_desc(const void* a, const void* b) {
    const sp_t* x = (const sp_t*)a; const sp_t* y = (const sp_t*)b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

static void tctx_init(tctx_t* t, uint32_t n, uint32_t ef_max) {
    memset(t, 0, sizeof(*t));
    t->visited = (uint32_t*)calloc(n ? n : 1, sizeof(uint32_t));
    t->nearest.d = (sp_t*)malloc(sizeof(sp_t) * (ef_max + 1));
    t->nearest.cap = ef_max;
}
static void tctx_free(tctx_t* t) { free(t->visited); free(t->nearest.d); free(t->cand.d); }

/* ---- scoring front-ends -------------------------------------------------------------------------------------- */
typedef struct { hnsw_t* h; tctx_t* t; qo_score_cb cb; void* user; const float* internal_query; const uint64_t* deleted; } scorer_t;
static void score_points(scorer_t* s, const uint32_t* ids, uint32_t n, float* out) {
    s->t->calls++; s->t->scored += n;
    if (s->cb) { s->cb(s->user, ids, n, out); return; }
    for (uint32_t i = 0; i < n; i++) out[i] = qo_similarity_f32(s->h->distance, s->internal_query, s->h->base + (size_t)ids[i] * s->h->dim, s->h->dim);
}
/* FilteredScorer::score_points (point_scorer.rs:265-281): drop ids the filter rejects, keep the first `limit` */
static uint32_t filter_ids(const scorer_t* s, uint32_t* ids, uint32_t n, uint32_t limit) {
    if (s->deleted) {
        uint32_t k = 0;
        for (uint32_t i = 0; i < n; i++) if (!((s->deleted[ids[i] >> 6] >> (ids[i] & 63)) & 1)) ids[k++] = ids[i];
        n = k;
    }
    return (limit && n > limit) ? limit : n;
}
static float score_internal(hnsw_t* h, uint32_t a, uint32_t b) {
    return qo_similarity_f32(h->distance, h->base + (size_t)a * h->dim, h->base + (size_t)b * h->dim, h->dim);
}

/* ---- traversal ----------------------------------------------------------------------------------------------- */
/* search_entry_on_level, graph_layers.rs:279-316 */
static sp_t search_entry_on_level(hnsw_t* h, scorer_t* s, uint32_t entry, uint32_t lvl) {
    uint32_t limit = level_m(h, lvl);
    uint32_t ids[512]; float sc[512];
    sp_t cur; cur.idx = entry; score_points(s, &entry, 1, &cur.score);
    int changed = 1;
    while (changed) {
        changed = 0;
        uint32_t n = read_links(h, cur.idx, lvl, ids);
        n = filter_ids(s, ids, n, limit);                  /* score_points(links, limit) filters, then truncates */
        if (n) score_points(s, ids, n, sc);
        for (uint32_t i = 0; i < n; i++) if (sc[i] > cur.score) { changed = 1; cur.idx = ids[i]; cur.score = sc[i]; }
    }
    return cur;
}
/* search_entry, graph_layers.rs:247-277: greedy from top_level down to target_level + 1 */
static sp_t search_entry(hnsw_t* h, scorer_t* s, uint32_t entry, uint32_t top_level, uint32_t target_level) {
    sp_t r; int have = 0; uint32_t e = entry;
    for (uint32_t lvl = top_level; lvl > target_level; lvl--) { r = search_entry_on_level(h, s, e, lvl); e = r.idx; have = 1; }
    if (!have) { r.idx = entry; score_points(s, &entry, 1, &r.score); }
    return r;
}
/* search_on_level, graph_layers.rs:108-148 + SearchContext (search_context.rs:8-41); result left in t->nearest */
static void search_on_level(hnsw_t* h, scorer_t* s, sp_t level_entry, uint32_t lvl, uint32_t ef) {
    tctx_t* t = s->t;
    t->stamp++;
    if (t->stamp == 0) { memset(t->visited, 0, sizeof(uint32_t) * h->n); t->stamp = 1; }
    t->visited[level_entry.idx] = t->stamp;
    flpq_t* nearest = &t->nearest; heap_t* cand = &t->cand;
    nearest->len = 0; nearest->cap = ef; cand->len = 0;
    if (flpq_push(nearest, level_entry)) maxheap_push(cand, level_entry);
    uint32_t limit = level_m(h, lvl);
    uint32_t ids[512], lk[512]; float sc[512];
    while (cand->len) {
        sp_t c = maxheap_pop(cand);
        float lower = nearest->len ? nearest->d[0].score : -INFINITY;
        if (c.score < lower) break;
        uint32_t nl = read_links(h, c.idx, lvl, lk), n = 0;
        for (uint32_t i = 0; i < nl; i++) if (t->visited[lk[i]] != t->stamp) ids[n++] = lk[i];
        n = filter_ids(s, ids, n, limit);
        if (n) score_points(s, ids, n, sc);
        for (uint32_t i = 0; i < n; i++) {
            sp_t p = { ids[i], sc[i] };
            if (flpq_push(nearest, p)) maxheap_push(cand, p);
            t->visited[ids[i]] = t->stamp;
        }
    }
}

/* ---- builder ------------------------------------------------------------------------------------------------- */
/* fill_from_sorted_with_heuristic, links_container.rs:47-71 */
static void fill_with_heuristic(hnsw_t* h, uint32_t* links, const sp_t* sorted, size_t n, uint32_t lm) {
    links[0] = 0;
    for (size_t i = 0; i < n && links[0] < lm; i++) {
        int ok = 1;
        for (uint32_t j = 0; j < links[0]; j++) if (score_internal(h, sorted[i].idx, links[1 + j]) > sorted[i].score) { ok = 0; break; }
        if (ok) links[1 + links[0]++] = sorted[i].idx;
    }
}
/* connect_with_heuristic (== connect_with_heuristic_simple by the reference's own comment, links_container.rs:115-139) */
static void connect_with_heuristic(hnsw_t* h, uint32_t* links, uint32_t new_point, uint32_t target, uint32_t lm) {
    if (links[0] < lm) { links[1 + links[0]++] = new_point; return; }
    sp_t c[513];
    uint32_t n = links[0];
    for (uint32_t i = 0; i < n; i++) { c[i].idx = links[1 + i]; c[i].score = score_internal(h, target, links[1 + i]); }
    c[n].idx = new_point; c[n].score = score_internal(h, target, new_point);
    qsort(c, n + 1, sizeof(sp_t), cmp_desc);
    fill_with_heuristic(h, links, c, n + 1, lm);
}

static uint64_t splitmix(uint64_t* s) { uint64_t z = (*s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

/* link_new_point, graph_layers_builder.rs:417-475 */
static void link_new_point(hnsw_t* h, tctx_t* t, uint32_t p, sp_t* sorted) {
    const uint32_t level = h->level[p];
    scorer_t s = { h, t, NULL, NULL, h->base + (size_t)p * h->dim, NULL };   /* FilteredScorer::new_internal(point) */
    pthread_mutex_lock(&h->entry_mu);
    const int has_entry = h->has_entry; const uint32_t entry = h->entry, entry_level = h->entry_level;
    if (!has_entry) { h->entry = p; h->entry_level = level; h->has_entry = 1; }
    pthread_mutex_unlock(&h->entry_mu);
    if (!has_entry) return;
    sp_t level_entry;
    if (entry_level > level) level_entry = search_entry(h, &s, entry, entry_level, level);
    else { level_entry.idx = entry; level_entry.score = score_internal(h, p, entry); }
    const uint32_t linking = level < entry_level ? level : entry_level;
    for (int cl = (int)linking; cl >= 0; cl--) {          /* link_new_point_on_level :502-530 */
        search_on_level(h, &s, level_entry, (uint32_t)cl, h->ef_construct);
        memcpy(sorted, t->nearest.d, t->nearest.len * sizeof(sp_t));
        qsort(sorted, t->nearest.len, sizeof(sp_t), cmp_desc);
        if (t->nearest.len) level_entry = sorted[0];
        const uint32_t lm = level_m(h, (uint32_t)cl);
        uint32_t mine[514];
        fill_with_heuristic(h, mine, sorted, t->nearest.len, lm);   /* link_with_heuristic :532-553 */
        plock(h, p);
        memcpy(h->links[p][cl], mine, (mine[0] + 1) * sizeof(uint32_t));
        punlock(h, p);
        for (uint32_t i = 0; i < mine[0]; i++) {
            const uint32_t o = mine[1 + i];
            plock(h, o);
            connect_with_heuristic(h, h->links[o][cl], p, o, lm);
            punlock(h, o);
        }
    }
    if (level > entry_level) {                                      /* entry_points.rs new_point */
        pthread_mutex_lock(&h->entry_mu);
        if (level > h->entry_level) { h->entry = p; h->entry_level = level; }
        pthread_mutex_unlock(&h->entry_mu);
    }
}

typedef struct { hnsw_t* h; _Atomic uint32_t* next; uint32_t tid; } build_arg_t;
static void pin_to_cpu(uint32_t tid) {
    cpu_set_t all, one;
    if (sched_getaffinity(0, sizeof(all), &all) != 0) return;
    int ncpu = CPU_COUNT(&all), k = (int)(tid % (uint32_t)(ncpu > 0 ? ncpu : 1)), seen = 0;
    for (int c = 0; c < CPU_SETSIZE; c++) if (CPU_ISSET(c, &all)) { if (seen++ == k) { CPU_ZERO(&one); CPU_SET(c, &one); sched_setaffinity(0, sizeof(one), &one); return; } }
}
static void* build_worker(void* ap) {
    build_arg_t* a = (build_arg_t*)ap;
    hnsw_t* h = a->h;
    pin_to_cpu(a->tid);
    tctx_t t; tctx_init(&t, h->n, h->ef_construct);
    sp_t* sorted = (sp_t*)malloc(sizeof(sp_t) * (h->ef_construct + 1));
    for (;;) {
        uint32_t p = atomic_fetch_add(a->next, 1);
        if (p >= h->n) break;
        link_new_point(h, &t, p, sorted);
    }
    free(sorted); tctx_free(&t);
    return NULL;
}

API void* qo_hnsw_build_mt(const float* base, uint32_t n, uint32_t dim, int distance, uint32_t m, uint32_t ef_construct, uint64_t seed, uint32_t threads) {
    hnsw_t* h = (hnsw_t*)calloc(1, sizeof(hnsw_t));
    h->n = n; h->dim = dim; h->m = m; h->m0 = 2 * m; h->ef_construct = ef_construct; h->distance = distance; h->base = base;
    pthread_mutex_init(&h->entry_mu, NULL);
    h->level = (uint8_t*)calloc(n ? n : 1, 1);
    h->lock = (atomic_uchar*)calloc(n ? n : 1, 1);
    const double level_factor = 1.0 / log((double)(m > 2 ? m : 2));   /* graph_layers_builder.rs:319 */
    uint64_t rs = seed;
    uint32_t*** lk = (uint32_t***)calloc(n ? n : 1, sizeof(uint32_t**));
    for (uint32_t p = 0; p < n; p++) {
        double u = ((splitmix(&rs) >> 11) + 1.0) * (1.0 / 9007199254740993.0);
        double lv = -log(u) * level_factor;
        uint32_t level = (uint32_t)floor(lv + 0.5);                    /* get_random_layer :388-396 */
        if (level > 30) level = 30;
        h->level[p] = (uint8_t)level;
        lk[p] = (uint32_t**)calloc(level + 1, sizeof(uint32_t*));
        for (uint32_t l = 0; l <= level; l++) lk[p][l] = (uint32_t*)calloc(level_m(h, l) + 2, sizeof(uint32_t));
    }
    h->links = lk;  /* links[p][lvl] */
    /* hnsw/build.rs:285-355: the first 256 points one by one, the rest in parallel */
    const uint32_t serial = (threads <= 1) ? n : (n < 256 ? n : 256);
    {
        tctx_t t; tctx_init(&t, n, ef_construct);
        sp_t* sorted = (sp_t*)malloc(sizeof(sp_t) * (ef_construct + 1));
        for (uint32_t p = 0; p < serial; p++) link_new_point(h, &t, p, sorted);
        free(sorted); tctx_free(&t);
    }
    if (serial < n) {
        h->locking = 1;
        _Atomic uint32_t next = serial;
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
        build_arg_t* args = (build_arg_t*)malloc(sizeof(build_arg_t) * threads);
        for (uint32_t i = 0; i < threads; i++) { args[i].h = h; args[i].next = &next; args[i].tid = i; pthread_create(&th[i], NULL, build_worker, &args[i]); }
        for (uint32_t i = 0; i < threads; i++) pthread_join(th[i], NULL);
        free(th); free(args);
        h->locking = 0;
    }
    h->n_score_calls = h->n_scored = 0;
    return h;
}

API void* qo_hnsw_build(const float* base, uint32_t n, uint32_t dim, int distance, uint32_t m, uint32_t ef_construct, uint64_t seed) {
    return qo_hnsw_build_mt(base, n, dim, distance, m, ef_construct, seed, 1);
}

/* GraphLayers::search, graph_layers.rs:530-561 */
static uint32_t search_one(hnsw_t* h, tctx_t* t, const float* query_pre, qo_score_cb cb, void* user, const uint64_t* deleted, uint32_t top, uint32_t ef, sp_t* out) {
    if (!h->has_entry) return 0;
    scorer_t s = { h, t, cb, user, query_pre, deleted };
    sp_t zero = search_entry(h, &s, h->entry, h->entry_level, 0);
    const uint32_t e = ef > top ? ef : top;
    if (e > t->nearest.cap) { t->nearest.d = (sp_t*)realloc(t->nearest.d, sizeof(sp_t) * (e + 1)); }
    search_on_level(h, &s, zero, 0, e);
    qsort(t->nearest.d, t->nearest.len, sizeof(sp_t), cmp_desc);
    const uint32_t n = t->nearest.len < top ? (uint32_t)t->nearest.len : top;
    memcpy(out, t->nearest.d, n * sizeof(sp_t));
    atomic_fetch_add(&h->n_score_calls, t->calls); atomic_fetch_add(&h->n_scored, t->scored);
    t->calls = t->scored = 0;
    return n;
}

/* cb == NULL -> CPU scoring with `query_pre` (preprocessed f32 query).  deleted: optional bitmap (bit = 1 -> filtered out). */
API uint32_t qo_hnsw_search(void* hp, const float* query_pre, qo_score_cb cb, void* user, uint32_t top, uint32_t ef, sp_t* out) {
    hnsw_t* h = (hnsw_t*)hp;
    tctx_t t; tctx_init(&t, h->n, ef > top ? ef : top);
    uint32_t n = search_one(h, &t, query_pre, cb, user, NULL, top, ef, out);
    tctx_free(&t);
    return n;
}

typedef struct { hnsw_t* h; const float* q; uint32_t nq, top, ef, tid; const uint64_t* deleted; sp_t* out; uint32_t* counts; _Atomic uint32_t* next; } search_arg_t;
static void* search_worker(void* ap) {
    search_arg_t* a = (search_arg_t*)ap;
    pin_to_cpu(a->tid);
    tctx_t t; tctx_init(&t, a->h->n, a->ef > a->top ? a->ef : a->top);
    for (;;) {
        uint32_t i = atomic_fetch_add(a->next, 1);
        if (i >= a->nq) break;
        a->counts[i] = search_one(a->h, &t, a->q + (size_t)i * a->h->dim, NULL, NULL, a->deleted, a->top, a->ef, a->out + (size_t)i * a->top);
    }
    tctx_free(&t);
    return NULL;
}
/* many searches at once, one per thread at a time (the reference runs one search per blocking task, segments_searcher.rs:255) */
API void qo_hnsw_search_batch(void* hp, const float* queries_pre, uint32_t nq, uint32_t top, uint32_t ef, const uint64_t* deleted, uint32_t threads,
                              sp_t* out, uint32_t* counts) {
    hnsw_t* h = (hnsw_t*)hp;
    if (threads < 1) threads = 1;
    if (threads > nq) threads = nq ? nq : 1;
    _Atomic uint32_t next = 0;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    search_arg_t* args = (search_arg_t*)malloc(sizeof(search_arg_t) * threads);
    for (uint32_t i = 0; i < threads; i++) {
        search_arg_t a = { h, queries_pre, nq, top, ef, i, deleted, out, counts, &next };
        args[i] = a;
        pthread_create(&th[i], NULL, search_worker, &args[i]);
    }
    for (uint32_t i = 0; i < threads; i++) pthread_join(th[i], NULL);
    free(th); free(args);
}

API void qo_hnsw_stats(void* hp, uint64_t* calls, uint64_t* scored, int reset) {
    hnsw_t* h = (hnsw_t*)hp;
    *calls = h->n_score_calls; *scored = h->n_scored;
    if (reset) h->n_score_calls = h->n_scored = 0;
}

API void qo_hnsw_entry(void* hp, uint32_t* entry, uint32_t* entry_level, uint32_t* m, uint32_t* m0) {
    hnsw_t* h = (hnsw_t*)hp;
    *entry = h->entry; *entry_level = h->entry_level; *m = h->m; *m0 = h->m0;
}

/* Plain `links.bin` (graph_links/serializer.rs:53-200, header.rs:9-20).  Returns the byte size; writes when out != NULL. */
API uint64_t qo_hnsw_export_plain(void* hp, uint8_t* out) {
    hnsw_t* h = (hnsw_t*)hp;
    const uint32_t n = h->n;
    uint32_t levels_count = 0;
    for (uint32_t p = 0; p < n; p++) if ((uint32_t)h->level[p] + 1 > levels_count) levels_count = (uint32_t)h->level[p] + 1;
    uint64_t* by_level = (uint64_t*)calloc(levels_count ? levels_count : 1, sizeof(uint64_t));
    for (uint32_t p = 0; p < n; p++) by_level[h->level[p]]++;
    /* back_index: points sorted by descending level count (counting sort, stable; the reference's sort_unstable_by_key leaves the
       order inside a level unspecified) */
    uint32_t* back = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    {
        uint64_t* start = (uint64_t*)calloc(levels_count + 1, sizeof(uint64_t));
        uint64_t acc = 0;
        for (int l = (int)levels_count - 1; l >= 0; l--) { start[l] = acc; acc += by_level[l]; }
        for (uint32_t p = 0; p < n; p++) back[start[h->level[p]]++] = p;
        free(start);
    }
    uint64_t total_neighbors = 0, total_offsets = 1;
    for (uint32_t p = 0; p < n; p++) for (uint32_t l = 0; l <= h->level[p]; l++) { total_neighbors += h->links[p][l][0]; total_offsets++; }
    uint64_t pos = 64 + 8ull * levels_count + 4ull * n + 4ull * total_neighbors;
    const uint64_t pad = (8 - pos % 8) % 8;
    const uint64_t total = pos + pad + 8ull * total_offsets;
    if (out) {
        memset(out, 0, 64);
        uint64_t hdr[5] = { n, levels_count, total_neighbors, total_offsets, pad };
        memcpy(out, hdr, sizeof(hdr));
        uint64_t* lo = (uint64_t*)(out + 64);
        uint64_t tot = 0, suffix = n;
        for (uint32_t l = 0; l < levels_count; l++) { lo[l] = tot; tot += suffix; suffix -= by_level[l]; }
        uint32_t* reindex = (uint32_t*)(out + 64 + 8ull * levels_count);
        for (uint32_t i = 0; i < n; i++) reindex[back[i]] = i;
        uint32_t* nb = reindex + n;
        uint64_t* offs = (uint64_t*)(out + pos + pad);
        memset(out + pos, 0, pad);
        uint64_t off = 0, oi = 0;
        offs[oi++] = 0;
        for (uint32_t l = 0; l < levels_count; l++) {
            uint64_t count = 0;
            for (uint32_t k = l; k < levels_count; k++) count += by_level[k];
            for (uint64_t i = 0; i < count; i++) {
                const uint32_t id = (l == 0) ? (uint32_t)i : back[i];
                const uint32_t* lk = h->links[id][l];
                memcpy(nb + off, lk + 1, lk[0] * sizeof(uint32_t));
                off += lk[0];
                offs[oi++] = off;
            }
        }
    }
    free(by_level); free(back);
    return total;
}

API void qo_hnsw_free(void* hp) {
    hnsw_t* h = (hnsw_t*)hp;
    uint32_t*** lk = h->links;
    for (uint32_t p = 0; p < h->n; p++) { for (uint32_t l = 0; l <= h->level[p]; l++) free(lk[p][l]); free(lk[p]); }
    free(lk); free(h->level); free(h->lock); pthread_mutex_destroy(&h->entry_mu); free(h);
}
