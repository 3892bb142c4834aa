/*
 * hnsw.c — from-spec CPU HNSW (build + search) used ONLY as the traversal driver of the config-#5 harness.
 *
 * TEST INFRASTRUCTURE (see oracle.c header).  The reference keeps HNSW traversal on the CPU and calls the scorer
 * through FilteredScorer::score_points once per hop (lib/segment/src/index/hnsw_index/graph_layers.rs:108-148,
 * 247-316, 530-561; search_context.rs:8-41).  This file restates that traversal with the scorer behind a callback,
 * so the same traversal can be driven by the CPU oracle scorer and by the GPU RawScorer (qb_score_points) and the
 * two result lists compared.  The builder follows graph_layers_builder.rs:388-566 and links_container.rs:47-71,139-...
 * (heuristic on, single-threaded => deterministic; level RNG is ours: the reference's rand stream is not reproducible
 * without Rust).
 */
#include <math.h>
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
    uint8_t* level;      /* per point */
    uint32_t*** links;   /* links[p][lvl] -> array: [count, ids...] with capacity level_m(lvl) */
    uint32_t entry, entry_level;
    int has_entry;
    uint32_t* visited;   /* visit stamps */
    uint32_t stamp;
    uint64_t n_score_calls, n_scored;
} hnsw_t;

static inline uint32_t level_m(const hnsw_t* h, uint32_t lvl) { return lvl == 0 ? h->m0 : h->m; }  /* HnswM::level_m, mod.rs:34-40 */

/* ---- heaps ------------------------------------------------------------------------------------------------- */
typedef struct { sp_t* d; size_t len, cap; } heap_t;
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
typedef struct { sp_t* d; size_t len, cap; } flpq_t;
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

/* ---- scoring front-ends -------------------------------------------------------------------------------------- */
typedef struct { hnsw_t* h; qo_score_cb cb; void* user; const float* internal_query; } scorer_t;
static void score_points(scorer_t* s, const uint32_t* ids, uint32_t n, float* out) {
    s->h->n_score_calls++; s->h->n_scored += n;
    if (s->cb) { s->cb(s->user, ids, n, out); return; }
    for (uint32_t i = 0; i < n; i++) out[i] = qo_similarity_f32(s->h->distance, s->internal_query, s->h->base + (size_t)ids[i] * s->h->dim, s->h->dim);
}
static float score_internal(hnsw_t* h, uint32_t a, uint32_t b) {
    return qo_similarity_f32(h->distance, h->base + (size_t)a * h->dim, h->base + (size_t)b * h->dim, h->dim);
}

/* ---- traversal ----------------------------------------------------------------------------------------------- */
/* search_entry_on_level, graph_layers.rs:279-316 */
static sp_t search_entry_on_level(hnsw_t* h, scorer_t* s, uint32_t entry, uint32_t lvl) {
    uint32_t limit = level_m(h, lvl);
    uint32_t ids[256]; float sc[256];
    sp_t cur; cur.idx = entry; score_points(s, &entry, 1, &cur.score);
    int changed = 1;
    while (changed) {
        changed = 0;
        const uint32_t* l = h->links[cur.idx][lvl];
        uint32_t n = l[0] < limit ? l[0] : limit;          /* score_points(links, limit) truncates */
        memcpy(ids, l + 1, n * sizeof(uint32_t));
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
/* search_on_level, graph_layers.rs:108-148 + SearchContext (search_context.rs:8-41); result left in `nearest` */
static void search_on_level(hnsw_t* h, scorer_t* s, sp_t level_entry, uint32_t lvl, uint32_t ef, flpq_t* nearest, heap_t* cand) {
    h->stamp++;
    if (h->stamp == 0) { memset(h->visited, 0, sizeof(uint32_t) * h->n); h->stamp = 1; }
    h->visited[level_entry.idx] = h->stamp;
    nearest->len = 0; nearest->cap = ef; cand->len = 0;
    if (flpq_push(nearest, level_entry)) maxheap_push(cand, level_entry);
    uint32_t limit = level_m(h, lvl);
    uint32_t ids[256]; float sc[256];
    while (cand->len) {
        sp_t c = maxheap_pop(cand);
        float lower = nearest->len ? nearest->d[0].score : -INFINITY;
        if (c.score < lower) break;
        const uint32_t* l = h->links[c.idx][lvl];
        uint32_t n = 0;
        for (uint32_t i = 0; i < l[0]; i++) if (h->visited[l[1 + i]] != h->stamp) ids[n++] = l[1 + i];
        if (n > limit) n = limit;
        if (n) score_points(s, ids, n, sc);
        for (uint32_t i = 0; i < n; i++) {
            sp_t p = { ids[i], sc[i] };
            if (flpq_push(nearest, p)) maxheap_push(cand, p);
            h->visited[ids[i]] = h->stamp;
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
    sp_t c[257];
    uint32_t n = links[0];
    for (uint32_t i = 0; i < n; i++) { c[i].idx = links[1 + i]; c[i].score = score_internal(h, target, links[1 + i]); }
    c[n].idx = new_point; c[n].score = score_internal(h, target, new_point);
    qsort(c, n + 1, sizeof(sp_t), cmp_desc);
    fill_with_heuristic(h, links, c, n + 1, lm);
}

static uint64_t splitmix(uint64_t* s) { uint64_t z = (*s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

API void* qo_hnsw_build(const float* base, uint32_t n, uint32_t dim, int distance, uint32_t m, uint32_t ef_construct, uint64_t seed) {
    hnsw_t* h = (hnsw_t*)calloc(1, sizeof(hnsw_t));
    h->n = n; h->dim = dim; h->m = m; h->m0 = 2 * m; h->ef_construct = ef_construct; h->distance = distance; h->base = base;
    h->level = (uint8_t*)calloc(n, 1);
    h->visited = (uint32_t*)calloc(n, sizeof(uint32_t));
    const double level_factor = 1.0 / log((double)(m > 2 ? m : 2));   /* graph_layers_builder.rs:319 */
    uint64_t rs = seed;
    uint32_t*** lk = (uint32_t***)calloc(n, sizeof(uint32_t**));
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
    flpq_t nearest = { (sp_t*)malloc(sizeof(sp_t) * (ef_construct + 1)), 0, ef_construct };
    heap_t cand = { NULL, 0, 0 };
    sp_t* sorted = (sp_t*)malloc(sizeof(sp_t) * (ef_construct + 1));
    for (uint32_t p = 0; p < n; p++) {
        uint32_t level = h->level[p];
        scorer_t s = { h, NULL, NULL, base + (size_t)p * dim };   /* FilteredScorer::new_internal(point) */
        if (h->has_entry) {                                       /* link_new_point :417-475 */
            sp_t level_entry;
            if (h->entry_level > level) level_entry = search_entry(h, &s, h->entry, h->entry_level, level);
            else { level_entry.idx = h->entry; level_entry.score = score_internal(h, p, h->entry); }
            uint32_t linking = level < h->entry_level ? level : h->entry_level;
            for (int cl = (int)linking; cl >= 0; cl--) {          /* link_new_point_on_level :502-530 */
                search_on_level(h, &s, level_entry, (uint32_t)cl, ef_construct, &nearest, &cand);
                memcpy(sorted, nearest.d, nearest.len * sizeof(sp_t));
                qsort(sorted, nearest.len, sizeof(sp_t), cmp_desc);
                if (nearest.len) level_entry = sorted[0];
                uint32_t lm = level_m(h, (uint32_t)cl);
                uint32_t* mine = lk[p][cl];
                fill_with_heuristic(h, mine, sorted, nearest.len, lm);   /* link_with_heuristic :532-553 */
                for (uint32_t i = 0; i < mine[0]; i++) connect_with_heuristic(h, lk[mine[1 + i]][cl], p, mine[1 + i], lm);
            }
        }
        if (!h->has_entry || level > h->entry_level) { h->entry = p; h->entry_level = level; h->has_entry = 1; }   /* entry_points.rs new_point */
    }
    free(nearest.d); free(cand.d); free(sorted);
    h->n_score_calls = h->n_scored = 0;
    return h;
}

/* GraphLayers::search, graph_layers.rs:530-561.  cb == NULL -> CPU scoring with `query_pre` (preprocessed f32 query). */
API uint32_t qo_hnsw_search(void* hp, const float* query_pre, qo_score_cb cb, void* user, uint32_t top, uint32_t ef, sp_t* out) {
    hnsw_t* h = (hnsw_t*)hp;
    if (!h->has_entry) return 0;
    scorer_t s = { h, cb, user, query_pre };
    sp_t zero = search_entry(h, &s, h->entry, h->entry_level, 0);
    uint32_t e = ef > top ? ef : top;
    flpq_t nearest = { (sp_t*)malloc(sizeof(sp_t) * (e + 1)), 0, e };
    heap_t cand = { NULL, 0, 0 };
    search_on_level(h, &s, zero, 0, e, &nearest, &cand);
    qsort(nearest.d, nearest.len, sizeof(sp_t), cmp_desc);
    uint32_t n = nearest.len < top ? (uint32_t)nearest.len : top;
    memcpy(out, nearest.d, n * sizeof(sp_t));
    free(nearest.d); free(cand.d);
    return n;
}

API void qo_hnsw_stats(void* hp, uint64_t* calls, uint64_t* scored, int reset) {
    hnsw_t* h = (hnsw_t*)hp;
    *calls = h->n_score_calls; *scored = h->n_scored;
    if (reset) h->n_score_calls = h->n_scored = 0;
}

API void qo_hnsw_free(void* hp) {
    hnsw_t* h = (hnsw_t*)hp;
    uint32_t*** lk = h->links;
    for (uint32_t p = 0; p < h->n; p++) { for (uint32_t l = 0; l <= h->level[p]; l++) free(lk[p][l]); free(lk[p]); }
    free(lk); free(h->level); free(h->visited); free(h);
}
