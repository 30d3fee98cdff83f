/*
 * tracegen.c -- deterministic synthetic trace + pool-state + index-plan generator.
 *
 * Shared by the tests, bench.py's GPU arm and its CPU-baseline arm, so that the engine and the oracle see the
 * SAME inputs.  It contains no hashing and no scoring: it only draws tokens, pool metrics and the plan of which
 * endpoint caches which prefix family to which depth.  Counter-based SplitMix64 => output is independent of
 * the number of generator threads.
 *
 *   prompts : uint32 token ids < 128000, T tokens each (hashed as little-endian bytes, 4 bytes/token)
 *   trace   : G = max(64, E/4) prefix families with canonical random prompts; request r is, w.p. 0.7, family
 *             g ~ Zipf(1.1) sharing its first L_r ~ U{0..B} blocks then random tokens; w.p. 0.3 fully random
 *   pool    : kv_usage = (u % 1001)/1000, waiting: 70 % zero else U{1..64}, running = U{0..256}
 *   index   : family g cached on c_g in {1,2,4,8} random endpoints to depth d ~ U{B/4..B} blocks; for 5 % of
 *             (g,e) the first k ~ U{1..8} blocks are missing (LRU-head eviction) so the global-stop rule fires
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
/* independent stream `s`, counter `i` */
static inline uint64_t rnd(uint64_t seed, uint64_t s, uint64_t i) { return sm64(sm64(seed ^ (s * 0xD1B54A32D192ED03ULL)) + i); }
static inline double unif(uint64_t x) { return (double)(x >> 11) * (1.0 / 9007199254740992.0); }

enum { S_FAMILY = 1, S_REQ_KIND = 2, S_REQ_FAM = 3, S_REQ_LEN = 4, S_REQ_TOK = 5, S_POOL = 6, S_PLAN = 7 };
#define VOCAB 128000u

typedef struct {
    uint64_t seed;
    int32_t E;           /* endpoints */
    int32_t T;           /* tokens per prompt */
    int32_t bst;         /* block size in tokens */
    int32_t G;           /* families (0 => max(64, E/4)) */
    int64_t R;           /* requests */
    int32_t n_prefill;   /* first n_prefill endpoints get role prefill (config 4), rest decode */
    int32_t _pad;
} tg_cfg;

int32_t tg_families(const tg_cfg *c) { return c->G > 0 ? c->G : (c->E / 4 > 64 ? c->E / 4 : 64); }
int32_t tg_blocks(const tg_cfg *c) { return c->T / c->bst; }

/* canonical family prompts: out[G][T] */
void tg_family_tokens(const tg_cfg *c, uint32_t *out) {
    int32_t G = tg_families(c);
    for (int64_t g = 0; g < G; g++)
        for (int64_t t = 0; t < c->T; t++)
            out[g * c->T + t] = (uint32_t)(rnd(c->seed, S_FAMILY, (uint64_t)(g * c->T + t)) % VOCAB);
}

typedef struct {
    const tg_cfg *c; const uint32_t *fam; const double *cdf; uint32_t *out; int32_t *fam_of; int32_t *shared;
    int64_t lo, hi;
} req_job;

static int32_t zipf_draw(const double *cdf, int32_t G, double u) {
    int32_t lo = 0, hi = G - 1;
    while (lo < hi) {
        int32_t mid = (lo + hi) / 2;
        if (cdf[mid] < u) lo = mid + 1; else hi = mid;
    }
    return lo;
}

static void *req_worker(void *arg) {
    req_job *j = (req_job *)arg;
    const tg_cfg *c = j->c;
    int32_t G = tg_families(c), B = tg_blocks(c);
    for (int64_t r = j->lo; r < j->hi; r++) {
        uint32_t *row = j->out + r * (int64_t)c->T;
        int32_t g = -1, L = 0;
        if (unif(rnd(c->seed, S_REQ_KIND, (uint64_t)r)) < 0.7) {
            g = zipf_draw(j->cdf, G, unif(rnd(c->seed, S_REQ_FAM, (uint64_t)r)));
            L = (int32_t)(rnd(c->seed, S_REQ_LEN, (uint64_t)r) % (uint64_t)(B + 1));
        }
        int64_t nshared = (int64_t)L * c->bst;
        if (g >= 0 && nshared > 0) memcpy(row, j->fam + (int64_t)g * c->T, (size_t)nshared * sizeof(uint32_t));
        for (int64_t t = nshared; t < c->T; t++)
            row[t] = (uint32_t)(rnd(c->seed, S_REQ_TOK, (uint64_t)(r * (int64_t)c->T + t)) % VOCAB);
        if (j->fam_of) j->fam_of[r] = g;
        if (j->shared) j->shared[r] = L;
    }
    return NULL;
}

/* requests [r0, r0+n) -> out[n][T]; fam_of / shared (optional) [n].  fam = tg_family_tokens output. */
void tg_requests(const tg_cfg *c, const uint32_t *fam, int64_t r0, int64_t n, uint32_t *out, int32_t *fam_of,
                 int32_t *shared, int n_threads) {
    int32_t G = tg_families(c);
    double *cdf = (double *)malloc(sizeof(double) * (size_t)G);
    double tot = 0;
    for (int32_t k = 0; k < G; k++) { tot += 1.0 / pow((double)(k + 1), 1.1); cdf[k] = tot; }
    for (int32_t k = 0; k < G; k++) cdf[k] /= tot;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    req_job *jobs = (req_job *)malloc(sizeof(req_job) * (size_t)n_threads);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; t++) {
        /* rows are written relative to `out`, request ids are absolute: shift the pointers */
        jobs[t] = (req_job){c, fam, cdf, out - r0 * (int64_t)c->T, fam_of ? fam_of - r0 : NULL, shared ? shared - r0 : NULL,
                            r0 + n * t / n_threads, r0 + n * (t + 1) / n_threads};
        pthread_create(&th[t], NULL, req_worker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    free(jobs); free(th); free(cdf);
}

/* role codes follow include/epp_engine.h epp_role: 1 = decode, 2 = prefill */
void tg_pool(const tg_cfg *c, double *kv, int32_t *waiting, int32_t *running, uint8_t *role) {
    for (int64_t e = 0; e < c->E; e++) {
        kv[e] = (double)(rnd(c->seed, S_POOL, (uint64_t)(4 * e)) % 1001u) / 1000.0;
        waiting[e] = unif(rnd(c->seed, S_POOL, (uint64_t)(4 * e + 1))) < 0.7
                         ? 0 : 1 + (int32_t)(rnd(c->seed, S_POOL, (uint64_t)(4 * e + 2)) % 64u);
        running[e] = (int32_t)(rnd(c->seed, S_POOL, (uint64_t)(4 * e + 3)) % 257u);
        role[e] = e < c->n_prefill ? 2 : 1;
    }
}

/* index plan: per family up to 8 (endpoint, depth, hole) triples.  Arrays [G][8]; n_per_family [G].
 * The (hash, endpoint) pairs are blocks [hole, depth) of family g on endpoint e. */
void tg_index_plan(const tg_cfg *c, int32_t *n_per_family, uint32_t *ep, int32_t *depth, int32_t *hole) {
    int32_t G = tg_families(c), B = tg_blocks(c);
    static const int32_t fan[4] = {1, 2, 4, 8};
    for (int64_t g = 0; g < G; g++) {
        int32_t cg = fan[rnd(c->seed, S_PLAN, (uint64_t)(g * 64)) % 4u];
        if (cg > c->E) cg = c->E;
        int32_t n = 0;
        for (int32_t k = 0; k < cg; k++) {
            uint64_t base = (uint64_t)(g * 64 + 1 + 4 * k);
            uint32_t e = (uint32_t)(rnd(c->seed, S_PLAN, base) % (uint64_t)c->E);
            int dup = 0;
            for (int32_t q = 0; q < n; q++) if (ep[g * 8 + q] == e) dup = 1;
            if (dup) continue;                       /* keep endpoints distinct within a family */
            int32_t lo = B / 4 > 0 ? B / 4 : (B > 0 ? 1 : 0);
            int32_t d = lo + (int32_t)(rnd(c->seed, S_PLAN, base + 1) % (uint64_t)(B - lo + 1));
            int32_t h = 0;
            if (unif(rnd(c->seed, S_PLAN, base + 2)) < 0.05) h = 1 + (int32_t)(rnd(c->seed, S_PLAN, base + 3) % 8u);
            if (h > d) h = d;
            ep[g * 8 + n] = e; depth[g * 8 + n] = d; hole[g * 8 + n] = h;
            n++;
        }
        n_per_family[g] = n;
    }
}
