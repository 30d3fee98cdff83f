/*
 * epp_oracle.c -- CPU ORACLE (test infrastructure only; see epp_oracle.h).
 *
 * Plain-C restatement of the reference Go loops.  Compile with -ffp-contract=off: Go on amd64
 * never fuses x*y+z (SURVEY.md section 7 "Bit-exact fp64"), so neither may this file.
 */
#include "epp_oracle.h"

#include <float.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* =====================================================================================
 * A.1  XXH64 -- what cespare/xxhash/v2 New/Write/Sum64 computes (call sites
 * approximateprefix/hashing.go:71-96).  Public XXH64 specification, seed parameterised.
 * ===================================================================================== */
#define P1 0x9E3779B185EBCA87ULL
#define P2 0xC2B2AE3D27D4EB4FULL
#define P3 0x165667B19E3779F9ULL
#define P4 0x85EBCA77C2B2AE63ULL
#define P5 0x27D4EB2F165667C5ULL

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t rd64(const uint8_t *p) {
    return (uint64_t)p[0] | ((uint64_t)p[1] << 8) | ((uint64_t)p[2] << 16) | ((uint64_t)p[3] << 24) |
           ((uint64_t)p[4] << 32) | ((uint64_t)p[5] << 40) | ((uint64_t)p[6] << 48) | ((uint64_t)p[7] << 56);
}
static inline uint32_t rd32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline uint64_t xxh_round(uint64_t acc, uint64_t x) { return rotl64(acc + x * P2, 31) * P1; }
static inline uint64_t xxh_merge(uint64_t h, uint64_t v) { return (h ^ xxh_round(0, v)) * P1 + P4; }

uint64_t orc_xxh64(const void *data, size_t len, uint64_t seed) {
    const uint8_t *p = (const uint8_t *)data;
    const uint8_t *end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        const uint8_t *limit = end - 32;
        do {
            v1 = xxh_round(v1, rd64(p));
            v2 = xxh_round(v2, rd64(p + 8));
            v3 = xxh_round(v3, rd64(p + 16));
            v4 = xxh_round(v4, rd64(p + 24));
            p += 32;
        } while (p <= limit);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xxh_merge(h, v1);
        h = xxh_merge(h, v2);
        h = xxh_merge(h, v3);
        h = xxh_merge(h, v4);
    } else {
        h = seed + P5;
    }
    h += (uint64_t)len;
    while (p + 8 <= end) {
        h ^= xxh_round(0, rd64(p));
        h = rotl64(h, 27) * P1 + P4;
        p += 8;
    }
    if (p + 4 <= end) {
        h ^= (uint64_t)rd32(p) * P1;
        h = rotl64(h, 23) * P2 + P3;
        p += 4;
    }
    while (p < end) {
        h ^= (uint64_t)(*p) * P5;
        h = rotl64(h, 11) * P1;
        p++;
    }
    h ^= h >> 33;
    h *= P2;
    h ^= h >> 29;
    h *= P3;
    h ^= h >> 32;
    return h;
}

/* Streaming digest of two concatenated pieces (h.Write(a); h.Write(b); h.Sum64()) without
 * assuming a bound on either piece. */
static uint64_t xxh64_concat2(const uint8_t *a, size_t na, const uint8_t *b, size_t nb) {
    size_t n = na + nb;
    uint8_t stackbuf[512];
    stackbuf[0] = 0;
    uint8_t *buf = n <= sizeof stackbuf ? stackbuf : (uint8_t *)malloc(n);
    if (na) memcpy(buf, a, na);
    if (nb) memcpy(buf + na, b, nb);
    uint64_t h = orc_xxh64(buf, n, 0);
    if (buf != stackbuf) free(buf);
    return h;
}

/* =====================================================================================
 * A.2  hashPrompt -- approximateprefix/hashing.go:35-99
 * ===================================================================================== */
int orc_hash_prompt(const uint8_t *data, size_t len, const uint8_t *model, size_t model_len,
                    const uint8_t *salt, size_t salt_len, int block_size_tokens, int max_prefix_blocks,
                    uint64_t *out, int out_cap) {
    /* hashing.go:49: cacheBlockSizeChars := blockSizeTokens * averageCharactersPerToken (types.go:113) */
    long long bs = (long long)block_size_tokens * 4;
    if (bs <= 0) return 0;                               /* hashing.go:51-56 */
    if ((long long)len < bs) return 0;                   /* hashing.go:58-61 */
    if ((long long)len > bs * (long long)max_prefix_blocks) {
        /* hashing.go:63-66; a negative/zero cap truncates to the empty string */
        long long cap = bs * (long long)max_prefix_blocks;
        len = cap > 0 ? (size_t)cap : 0;
    }
    /* hashing.go:71-78: h.Write(model); if salt != "" h.Write(salt); prev = h.Sum64() */
    uint64_t prev = xxh64_concat2(model, model_len, salt, salt_len);
    uint8_t le[8];
    int n = 0;
    size_t i = 0;
    size_t ubs = (size_t)bs;
    for (; i + ubs <= len; i += ubs) {                   /* hashing.go:80-87 */
        for (int k = 0; k < 8; k++) le[k] = (uint8_t)(prev >> (8 * k));   /* toBytes, :101-105 */
        prev = xxh64_concat2(data + i, ubs, le, 8);
        if (n < out_cap) out[n] = prev;
        n++;
    }
    if (i < len) {                                       /* hashing.go:90-96 partial block */
        for (int k = 0; k < 8; k++) le[k] = (uint8_t)(prev >> (8 * k));
        prev = xxh64_concat2(data + i, len - i, le, 8);
        if (n < out_cap) out[n] = prev;
        n++;
    }
    return n;
}

/* =====================================================================================
 * small open-addressed u64 -> u64 map (linear probing, backward-shift delete)
 * ===================================================================================== */
typedef struct {
    uint64_t *keys;
    uint64_t *vals;
    uint8_t *used;
    size_t cap;   /* power of two */
    size_t n;
} umap;

static inline size_t umix(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33;
    return (size_t)k;
}
static void umap_init(umap *m, size_t cap) {
    size_t c = 16;
    while (c < cap) c <<= 1;
    m->cap = c; m->n = 0;
    m->keys = (uint64_t *)malloc(c * sizeof(uint64_t));
    m->vals = (uint64_t *)malloc(c * sizeof(uint64_t));
    m->used = (uint8_t *)calloc(c, 1);
}
static void umap_free(umap *m) { free(m->keys); free(m->vals); free(m->used); m->keys = m->vals = NULL; m->used = NULL; }
static int umap_find(const umap *m, uint64_t k, uint64_t *v) {
    size_t mask = m->cap - 1, i = umix(k) & mask;
    while (m->used[i]) {
        if (m->keys[i] == k) { if (v) *v = m->vals[i]; return 1; }
        i = (i + 1) & mask;
    }
    return 0;
}
static void umap_put(umap *m, uint64_t k, uint64_t v);
static void umap_grow(umap *m) {
    umap o = *m;
    umap_init(m, o.cap * 2);
    for (size_t i = 0; i < o.cap; i++) if (o.used[i]) umap_put(m, o.keys[i], o.vals[i]);
    umap_free(&o);
}
static void umap_put(umap *m, uint64_t k, uint64_t v) {
    if ((m->n + 1) * 10 > m->cap * 7) umap_grow(m);
    size_t mask = m->cap - 1, i = umix(k) & mask;
    while (m->used[i]) {
        if (m->keys[i] == k) { m->vals[i] = v; return; }
        i = (i + 1) & mask;
    }
    m->used[i] = 1; m->keys[i] = k; m->vals[i] = v; m->n++;
}
static void umap_del(umap *m, uint64_t k) {
    size_t mask = m->cap - 1, i = umix(k) & mask;
    while (m->used[i]) {
        if (m->keys[i] == k) break;
        i = (i + 1) & mask;
    }
    if (!m->used[i]) return;
    size_t j = i;
    for (;;) {
        j = (j + 1) & mask;
        if (!m->used[j]) break;
        size_t h = umix(m->keys[j]) & mask;
        /* can entry j move into hole i?  yes iff h is cyclically outside (i, j] */
        int between = (i <= j) ? (h > i && h <= j) : (h > i || h <= j);
        if (!between) { m->keys[i] = m->keys[j]; m->vals[i] = m->vals[j]; i = j; }
    }
    m->used[i] = 0; m->n--;
}

/* =====================================================================================
 * per-server LRU -- hashicorp/golang-lru v2.0.7 simplelru semantics as used by indexer.go:
 * Add(k): existing -> move to front (no eviction); new -> push front, evict the oldest when
 * len > size (eviction callback fires).  Remove(k) also fires the callback.
 * ===================================================================================== */
typedef struct {
    uint64_t key;
    int32_t prev, next;   /* toward newer / toward older */
} lru_node;

typedef struct {
    uint32_t server;
    int32_t size;
    int32_t len;
    int32_t head, tail;   /* head = most recent, tail = oldest */
    int32_t free_head;
    lru_node *nodes;
    int32_t nodes_cap;
    umap pos;             /* key -> node index */
} lru;

static lru *lru_new(uint32_t server, int size) {
    lru *l = (lru *)calloc(1, sizeof(lru));
    l->server = server; l->size = size; l->head = l->tail = -1; l->free_head = -1;
    l->nodes_cap = 64; l->nodes = (lru_node *)malloc(sizeof(lru_node) * (size_t)l->nodes_cap);
    for (int i = 0; i < l->nodes_cap; i++) { l->nodes[i].next = l->free_head; l->free_head = i; }
    umap_init(&l->pos, 64);
    return l;
}
static void lru_free(lru *l) { umap_free(&l->pos); free(l->nodes); free(l); }
static int32_t lru_alloc_node(lru *l) {
    if (l->free_head < 0) {
        int32_t oc = l->nodes_cap;
        l->nodes_cap *= 2;
        l->nodes = (lru_node *)realloc(l->nodes, sizeof(lru_node) * (size_t)l->nodes_cap);
        for (int32_t i = oc; i < l->nodes_cap; i++) { l->nodes[i].next = l->free_head; l->free_head = i; }
    }
    int32_t i = l->free_head;
    l->free_head = l->nodes[i].next;
    return i;
}
static void lru_unlink(lru *l, int32_t i) {
    lru_node *nd = &l->nodes[i];
    if (nd->prev >= 0) l->nodes[nd->prev].next = nd->next; else l->head = nd->next;
    if (nd->next >= 0) l->nodes[nd->next].prev = nd->prev; else l->tail = nd->prev;
}
static void lru_push_front(lru *l, int32_t i) {
    lru_node *nd = &l->nodes[i];
    nd->prev = -1; nd->next = l->head;
    if (l->head >= 0) l->nodes[l->head].prev = i;
    l->head = i;
    if (l->tail < 0) l->tail = i;
}

/* =====================================================================================
 * A.3  indexer -- approximateprefix/indexer.go:32-37
 * ===================================================================================== */
typedef struct {
    uint32_t *ids;
    int32_t n, cap;
} podset;

struct orc_indexer {
    umap hash_to_pods;      /* blockHash -> index into sets[]       (indexer.go:34) */
    podset *sets;
    int64_t sets_n, sets_cap;
    int64_t *free_sets;
    int64_t free_n, free_cap;
    umap pod_to_lru;        /* ServerID -> index into lrus[]        (indexer.go:35) */
    lru **lrus;
    int32_t lrus_n, lrus_cap;
    int default_lru_size;   /* indexer.go:36 */
};

orc_indexer *orc_indexer_new(int default_lru_size) {
    orc_indexer *ix = (orc_indexer *)calloc(1, sizeof(orc_indexer));
    umap_init(&ix->hash_to_pods, 1024);
    umap_init(&ix->pod_to_lru, 64);
    ix->default_lru_size = default_lru_size;
    return ix;
}
void orc_indexer_free(orc_indexer *ix) {
    if (!ix) return;
    for (int64_t i = 0; i < ix->sets_n; i++) free(ix->sets[i].ids);
    free(ix->sets); free(ix->free_sets);
    for (int32_t i = 0; i < ix->lrus_n; i++) if (ix->lrus[i]) lru_free(ix->lrus[i]);
    free(ix->lrus);
    umap_free(&ix->hash_to_pods); umap_free(&ix->pod_to_lru);
    free(ix);
}
static int64_t ix_new_set(orc_indexer *ix) {
    if (ix->free_n > 0) return ix->free_sets[--ix->free_n];
    if (ix->sets_n == ix->sets_cap) {
        ix->sets_cap = ix->sets_cap ? ix->sets_cap * 2 : 1024;
        ix->sets = (podset *)realloc(ix->sets, sizeof(podset) * (size_t)ix->sets_cap);
    }
    podset *s = &ix->sets[ix->sets_n];
    s->ids = NULL; s->n = 0; s->cap = 0;
    return ix->sets_n++;
}
static void ix_release_set(orc_indexer *ix, int64_t si) {
    ix->sets[si].n = 0;
    if (ix->free_n == ix->free_cap) {
        ix->free_cap = ix->free_cap ? ix->free_cap * 2 : 256;
        ix->free_sets = (int64_t *)realloc(ix->free_sets, sizeof(int64_t) * (size_t)ix->free_cap);
    }
    ix->free_sets[ix->free_n++] = si;
}
/* hashToPods[hash][server] = struct{}{}  (indexer.go:74-82) */
static void ix_set_insert(orc_indexer *ix, uint64_t hash, uint32_t server) {
    uint64_t si;
    if (!umap_find(&ix->hash_to_pods, hash, &si)) {
        si = (uint64_t)ix_new_set(ix);
        umap_put(&ix->hash_to_pods, hash, si);
    }
    podset *s = &ix->sets[si];
    for (int32_t i = 0; i < s->n; i++) if (s->ids[i] == server) return;
    if (s->n == s->cap) {
        s->cap = s->cap ? s->cap * 2 : 4;
        s->ids = (uint32_t *)realloc(s->ids, sizeof(uint32_t) * (size_t)s->cap);
    }
    s->ids[s->n++] = server;
}
/* makeEvictionFn: indexer.go:105-115 */
static void ix_evict(orc_indexer *ix, uint64_t hash, uint32_t server) {
    uint64_t si;
    if (!umap_find(&ix->hash_to_pods, hash, &si)) return;
    podset *s = &ix->sets[si];
    for (int32_t i = 0; i < s->n; i++) {
        if (s->ids[i] == server) { s->ids[i] = s->ids[--s->n]; break; }
    }
    if (s->n == 0) {
        umap_del(&ix->hash_to_pods, hash);
        ix_release_set(ix, (int64_t)si);
    }
}
static void ix_lru_add(orc_indexer *ix, lru *l, uint64_t key) {
    uint64_t ni;
    if (umap_find(&l->pos, key, &ni)) {             /* existing: refresh recency only */
        lru_unlink(l, (int32_t)ni);
        lru_push_front(l, (int32_t)ni);
        return;
    }
    int32_t i = lru_alloc_node(l);
    l->nodes[i].key = key;
    lru_push_front(l, i);
    umap_put(&l->pos, key, (uint64_t)i);
    l->len++;
    if (l->len > l->size) {                         /* evict oldest */
        int32_t t = l->tail;
        uint64_t old = l->nodes[t].key;
        lru_unlink(l, t);
        umap_del(&l->pos, old);
        l->nodes[t].next = l->free_head; l->free_head = t;
        l->len--;
        ix_evict(ix, old, l->server);
    }
}

void orc_indexer_add(orc_indexer *ix, const uint64_t *hashes, int n, uint32_t server, int num_gpu_blocks) {
    uint64_t li;
    lru *l;
    if (!umap_find(&ix->pod_to_lru, server, &li)) {  /* indexer.go:57-69 */
        int size = num_gpu_blocks;
        if (size <= 0) size = ix->default_lru_size;
        if (size <= 0) size = 1;                     /* lru.NewWithEvict errors on <= 0; keep usable */
        l = lru_new(server, size);
        if (ix->lrus_n == ix->lrus_cap) {
            ix->lrus_cap = ix->lrus_cap ? ix->lrus_cap * 2 : 64;
            ix->lrus = (lru **)realloc(ix->lrus, sizeof(lru *) * (size_t)ix->lrus_cap);
        }
        ix->lrus[ix->lrus_n] = l;
        umap_put(&ix->pod_to_lru, server, (uint64_t)ix->lrus_n);
        ix->lrus_n++;
    } else {
        l = ix->lrus[li];
    }
    for (int i = 0; i < n; i++) ix_lru_add(ix, l, hashes[i]);        /* indexer.go:71-74 */
    for (int i = 0; i < n; i++) ix_set_insert(ix, hashes[i], server); /* indexer.go:76-83 */
}

int orc_indexer_get(const orc_indexer *ix, uint64_t hash, uint32_t *out, int out_cap) {
    uint64_t si;
    if (!umap_find(&ix->hash_to_pods, hash, &si)) return 0;           /* indexer.go:90-93 */
    const podset *s = &ix->sets[si];
    for (int32_t i = 0; i < s->n && i < out_cap; i++) out[i] = s->ids[i];   /* deep copy, :95-99 */
    return s->n;
}

void orc_indexer_remove_pod(orc_indexer *ix, uint32_t server) {
    uint64_t li;
    if (!umap_find(&ix->pod_to_lru, server, &li)) return;             /* indexer.go:171-174 */
    lru *l = ix->lrus[li];
    /* for _, hash := range lruCache.Keys() { lruCache.Remove(hash) }   indexer.go:177-179 */
    for (int32_t i = l->tail; i >= 0;) {
        int32_t nx = l->nodes[i].prev;
        ix_evict(ix, l->nodes[i].key, server);
        i = nx;
    }
    lru_free(l);
    ix->lrus[li] = NULL;
    umap_del(&ix->pod_to_lru, server);                                /* indexer.go:181 */
}

int orc_indexer_pods(const orc_indexer *ix, uint32_t *out, int out_cap) {
    int n = 0;
    for (int32_t i = 0; i < ix->lrus_n; i++) {
        if (!ix->lrus[i]) continue;
        if (n < out_cap) out[n] = ix->lrus[i]->server;
        n++;
    }
    return n;
}

int orc_indexer_lru_len(const orc_indexer *ix, uint32_t server) {
    uint64_t li;
    if (!umap_find(&ix->pod_to_lru, server, &li)) return -1;
    return ix->lrus[li]->len;
}

size_t orc_indexer_export(const orc_indexer *ix, uint64_t *hashes, uint32_t *servers, size_t cap) {
    size_t n = 0;
    const umap *m = &ix->hash_to_pods;
    for (size_t i = 0; i < m->cap; i++) {
        if (!m->used[i]) continue;
        const podset *s = &ix->sets[m->vals[i]];
        for (int32_t k = 0; k < s->n; k++) {
            if (n < cap) { hashes[n] = m->keys[i]; servers[n] = s->ids[k]; }
            n++;
        }
    }
    return n;
}

void orc_indexer_load_pairs(orc_indexer *ix, const uint64_t *hashes, const uint32_t *servers, size_t n) {
    for (size_t i = 0; i < n; i++) ix_set_insert(ix, hashes[i], servers[i]);
}

/* matchLongestPrefix -- approximateprefix/plugin.go:214-230 */
int orc_match_longest_prefix(const orc_indexer *ix, const uint64_t *hashes, int n,
                             int32_t *counts, int n_servers) {
    uint32_t stackbuf[256];
    uint32_t *buf = stackbuf;
    int buf_cap = 256;
    int walked = 0;
    for (int i = 0; i < n; i++) {
        int c = orc_indexer_get(ix, hashes[i], buf, buf_cap);      /* cachedServers := Get(hash) */
        if (c > buf_cap) {
            if (buf != stackbuf) free(buf);
            buf_cap = c;
            buf = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)c);
            c = orc_indexer_get(ix, hashes[i], buf, buf_cap);
        }
        if (c == 0) break;                                          /* plugin.go:221-223 */
        for (int k = 0; k < c; k++)                                 /* res[server]++, :225-227 */
            if ((int64_t)buf[k] < (int64_t)n_servers) counts[buf[k]]++;
        walked++;
    }
    if (buf != stackbuf) free(buf);
    return walked;
}

/* =====================================================================================
 * A.6 role filters -- filter/bylabel/roles.go:46-70, filter.go:104-117
 * ===================================================================================== */
int orc_role_filter_keeps(int filter, int role) {
    if (role == 0xFF) return 0;                     /* slot not in the pool */
    switch (filter) {
    case ORC_FILTER_NONE: return 1;
    case ORC_FILTER_DECODE:   /* roles.go:46-48: allowsNoLabel=true; decode, prefill-decode, both, e-p-d */
        return role == ORC_ROLE_NONE || role == ORC_ROLE_DECODE || role == ORC_ROLE_PREFILL_DECODE ||
               role == ORC_ROLE_BOTH || role == ORC_ROLE_ENCODE_PREFILL_DECODE;
    case ORC_FILTER_PREFILL:  /* roles.go:56-58: label required; prefill, encode-prefill, prefill-decode, both, e-p-d */
        return role == ORC_ROLE_PREFILL || role == ORC_ROLE_ENCODE_PREFILL || role == ORC_ROLE_PREFILL_DECODE ||
               role == ORC_ROLE_BOTH || role == ORC_ROLE_ENCODE_PREFILL_DECODE;
    case ORC_FILTER_ENCODE:   /* roles.go:68-70: label required; encode, encode-prefill, e-p-d */
        return role == ORC_ROLE_ENCODE || role == ORC_ROLE_ENCODE_PREFILL || role == ORC_ROLE_ENCODE_PREFILL_DECODE;
    default: return 0;
    }
}

/* =====================================================================================
 * A.4 scorers
 * ===================================================================================== */
void orc_score_column(const orc_scorer *s, const orc_pool *pool, const uint8_t *cand,
                      const int32_t *match, int32_t total, double *out) {
    int n = pool->n;
    switch (s->kind) {
    case ORC_SCORER_PREFIX:        /* scorer/prefix/plugin.go:100-111 */
        for (int e = 0; e < n; e++) {
            out[e] = 0.0;
            if (!cand[e]) continue;
            if (total != 0) out[e] = (double)match[e] / (double)total;
        }
        break;
    case ORC_SCORER_KV_UTIL:       /* kvcache_utilization.go:79 */
        for (int e = 0; e < n; e++) out[e] = cand[e] ? 1 - pool->kv_usage[e] : 0.0;
        break;
    case ORC_SCORER_QUEUE:         /* queue.go:79-107 */
    case ORC_SCORER_RUNNING: {     /* runningrequest.go:79-107 (same shape on RunningRequestsSize) */
        const int32_t *q = s->kind == ORC_SCORER_QUEUE ? pool->waiting : pool->running;
        int64_t mn = INT64_MAX, mx = INT64_MIN;
        for (int e = 0; e < n; e++) {
            if (!cand[e]) continue;
            if (q[e] < mn) mn = q[e];
            if (q[e] > mx) mx = q[e];
        }
        for (int e = 0; e < n; e++) {
            if (!cand[e]) { out[e] = 0.0; continue; }
            if (mx == mn) out[e] = 1.0;
            else out[e] = (double)(mx - (int64_t)q[e]) / (double)(mx - mn);
        }
        break;
    }
    case ORC_SCORER_LOAD_AWARE: {  /* load_aware.go:43-52, 87-97 */
        double thr = s->param;
        if (!(thr > 0)) thr = 128.0;   /* NewLoadAware: queueThreshold <= 0 -> default */
        for (int e = 0; e < n; e++) {
            if (!cand[e]) { out[e] = 0.0; continue; }
            double w = (double)pool->waiting[e];
            if (w == 0) out[e] = 0.5;
            else {
                if (w > thr) w = thr;
                out[e] = 0.5 * (1.0 - (w / thr));
            }
        }
        break;
    }
    case ORC_SCORER_EXTERNAL: {
        int col = (int)s->param;
        for (int e = 0; e < n; e++)
            out[e] = (cand[e] && col >= 0 && col < pool->n_ext_cols) ? pool->ext[(size_t)col * (size_t)n + (size_t)e] : 0.0;
        break;
    }
    case ORC_SCORER_TOKEN_LOAD: {  /* token_load.go:84-112 */
        double thr = s->param;
        if (!(thr > 0)) thr = 4194304.0;            /* TokenLoadScorerFactory: <= 0 -> tokenQueueThresholdDefault */
        int col = s->column;
        for (int e = 0; e < n; e++) {
            if (!cand[e]) { out[e] = 0.0; continue; }
            double load = (col >= 0 && col < pool->n_ext_cols) ? pool->ext[(size_t)col * (size_t)n + (size_t)e] : 0.0;
            if (load <= 0) out[e] = 1.0;
            else {
                if (load > thr) load = thr;
                out[e] = 1.0 - (load / thr);
            }
        }
        break;
    }
    case ORC_SCORER_ACTIVE_REQUEST: {  /* active_request.go:140-173; NewActiveRequest :83-93 */
        int64_t idle = s->param2 >= 0 ? (int64_t)s->param2 : 0;
        double max_busy = (s->param >= 0 && s->param <= 1.0) ? s->param : 1.0;
        int col = s->column;
        int64_t max_count = 0;
        for (int e = 0; e < n; e++) {
            if (!cand[e]) continue;
            int64_t c = (col >= 0 && col < pool->n_ext_cols) ? (int64_t)pool->ext[(size_t)col * (size_t)n + (size_t)e] : 0;
            if (c > max_count) max_count = c;
        }
        for (int e = 0; e < n; e++) {
            if (!cand[e]) { out[e] = 0.0; continue; }
            int64_t c = (col >= 0 && col < pool->n_ext_cols) ? (int64_t)pool->ext[(size_t)col * (size_t)n + (size_t)e] : 0;
            if (c <= idle) out[e] = 1.0;
            else out[e] = (double)(max_count - c) / (double)max_count * max_busy;
        }
        break;
    }
    case ORC_SCORER_LORA_AFFINITY:  /* lora_affinity.go:76-100: the switch, in its order */
        for (int e = 0; e < n; e++) {
            if (!cand[e]) { out[e] = 0.0; continue; }
            int st = pool->lora_state ? pool->lora_state[e] : 0;
            int room = pool->lora_max && pool->lora_loaded && pool->lora_loaded[e] < pool->lora_max[e];
            if (st == 1) out[e] = 1.0;            /* active */
            else if (room) out[e] = 0.8;          /* capacity for one more adapter */
            else if (st == 2) out[e] = 0.6;       /* waiting */
            else out[e] = 0.0;
        }
        break;
    default:
        for (int e = 0; e < n; e++) out[e] = 0.0;
    }
}

/* enforceScoreRange -- scheduling/scheduler_profile.go:194-202 */
static inline double enforce_score_range(double s) {
    if (s < 0) return 0;
    if (s > 1) return 1;
    return s;
}

/* Tie rule of the BUILD (not of the reference): MaxScorePicker.Pick shuffles with math/rand before its stable sort
 * (maxscore/picker.go:91-102), i.e. the winner is a uniformly random member of the arg-max set.  With tie_seed == 0
 * engine and oracle report the LOWEST slot of the set; with tie_seed != 0 both pick the member of rank
 * orc_tie_rank(seed, key, |set|) in ascending slot order, key = 4 * (request ordinal) + profile index
 * (0 primary / decode, 1 prefill, 2 encode) -- uniform over the set, reproducible on both sides. */
static inline uint64_t mix64(uint64_t z) {            /* SplitMix64 output function */
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
uint32_t orc_tie_rank(uint64_t seed, uint64_t key, uint32_t n) {
    uint64_t m = mix64(seed ^ mix64(key));
    return (uint32_t)(((m >> 32) * (uint64_t)n) >> 32);
}

/* =====================================================================================
 * A.5 SchedulerProfile.Run -- scheduling/scheduler_profile.go:117-192 + maxscore picker
 * ===================================================================================== */
typedef struct { int k; int32_t *picks; double *scores; int32_t n; } topk_req;   /* the first k of the picker, optional */
static int profile_run_scratch(const orc_profile *p, const orc_pool *pool, const int32_t *match, int32_t total,
                               double *out_scores, double *out_max, int32_t *out_pick, int32_t *argmax_set,
                               uint8_t *cand, double *col, uint64_t tie_seed, uint64_t tie_key, topk_req *tk);

int orc_profile_run(const orc_profile *p, const orc_pool *pool, const int32_t *match, int32_t total,
                    double *out_scores, double *out_max, int32_t *out_pick, int32_t *argmax_set) {
    int n = pool->n;
    uint8_t *cand = (uint8_t *)malloc((size_t)n + 1);
    double *col = (double *)malloc(sizeof(double) * ((size_t)n + 1));
    int r = profile_run_scratch(p, pool, match, total, out_scores, out_max, out_pick, argmax_set, cand, col, 0, 0, NULL);
    free(cand); free(col);
    return r;
}

int orc_profile_run_topk(const orc_profile *p, const orc_pool *pool, const int32_t *match, int32_t total,
                         uint64_t tie_seed, uint64_t tie_key, int k, int32_t *topk_picks, double *topk_scores,
                         int32_t *out_n_picks, double *out_scores) {
    int n = pool->n;
    uint8_t *cand = (uint8_t *)malloc((size_t)n + 1);
    double *col = (double *)malloc(sizeof(double) * ((size_t)n + 1));
    double *scores = out_scores ? out_scores : (double *)malloc(sizeof(double) * ((size_t)n + 1));
    topk_req tk = {k, topk_picks, topk_scores, 0};
    double mx; int32_t pick;
    int r = profile_run_scratch(p, pool, match, total, scores, &mx, &pick, NULL, cand, col, tie_seed, tie_key, &tk);
    if (out_n_picks) *out_n_picks = tk.n;
    if (!out_scores) free(scores);
    free(cand); free(col);
    return r;
}

double orc_explore_u(uint64_t seed, uint64_t key) {
    return (double)(mix64((seed ^ 0xA0761D6478BD642FULL) ^ mix64(key)) >> 11) * 0x1.0p-53;
}

/* Plugin.Filter -- filter/prefixcacheaffinity/plugin.go:105-151: narrows cand[] to the sticky endpoints. */
static void affinity_filter(const orc_profile *p, const orc_pool *pool, const int32_t *match, int32_t total,
                            uint8_t *cand, int n_cand, uint64_t tie_seed, uint64_t tie_key) {
    int n = pool->n;
    if (n_cand <= 1 || !(p->affinity_threshold > 0)) return;                           /* :108-110 */
    if (tie_seed && orc_explore_u(tie_seed, tie_key) < p->exploration_probability) return;   /* :113-117 */
    const double *ttft = (p->ttft_column >= 0 && p->ttft_column < pool->n_ext_cols)
                             ? pool->ext + (size_t)p->ttft_column * (size_t)n : NULL;
    int n_sticky = 0, n_non = 0;
    double best_sticky = DBL_MAX, best_non = DBL_MAX;                                   /* bestTTFT, :173-184 */
    for (int e = 0; e < n; e++) {
        if (!cand[e]) continue;
        double score = 0;                                                              /* prefixCacheScore, :160-171 */
        if (total > 0) score = (double)match[e] / (double)total;
        double t = ttft ? ttft[e] : DBL_MAX;
        if (score >= p->affinity_threshold) { n_sticky++; if (t < best_sticky) best_sticky = t; }
        else { n_non++; if (t < best_non) best_non = t; }
    }
    if (n_sticky == 0) return;                                                          /* :130-134 */
    if (p->max_ttft_penalty_ms > 0 && n_non > 0 && best_sticky - best_non > p->max_ttft_penalty_ms) return;   /* :137-146 */
    for (int e = 0; e < n; e++) {
        if (!cand[e]) continue;
        double score = 0;
        if (total > 0) score = (double)match[e] / (double)total;
        if (!(score >= p->affinity_threshold)) cand[e] = 0;
    }
}

/* The Go code allocates fresh maps per call; the timed CPU baseline reuses per-thread scratch instead, which only
 * makes the baseline FASTER than the reference. */
static int profile_run_scratch(const orc_profile *p, const orc_pool *pool, const int32_t *match, int32_t total,
                               double *out_scores, double *out_max, int32_t *out_pick, int32_t *argmax_set,
                               uint8_t *cand, double *col, uint64_t tie_seed, uint64_t tie_key, topk_req *tk) {
    int n = pool->n;
    int n_cand = 0;
    if (tk) {
        tk->n = 0;
        for (int j = 0; j < tk->k; j++) { if (tk->picks) tk->picks[j] = -1; if (tk->scores) tk->scores[j] = 0; }
    }
    for (int e = 0; e < n; e++) {                       /* runFilterPlugins, :130-149: the role filter ... */
        cand[e] = (uint8_t)orc_role_filter_keeps(p->filter, pool->role[e]);
        n_cand += cand[e];
    }
    affinity_filter(p, pool, match, total, cand, n_cand, tie_seed, tie_key);   /* ... then the affinity filter */
    if (n_cand == 0) {                                  /* :119-121 */
        for (int e = 0; e < n; e++) out_scores[e] = -1.0;
        if (out_max) *out_max = 0;
        if (out_pick) *out_pick = -1;
        return 0;
    }
    for (int e = 0; e < n; e++) out_scores[e] = cand[e] ? 0.0 : -1.0;    /* :155-158 */
    for (int s = 0; s < p->n_scorers; s++) {            /* :160-170, in profile order */
        orc_score_column(&p->scorers[s], pool, cand, match, total, col);
        double w = p->scorers[s].weight;
        for (int e = 0; e < n; e++) {
            if (!cand[e]) continue;
            double prod = enforce_score_range(col[e]) * w;
            out_scores[e] = out_scores[e] + prod;       /* += : rounded product, rounded sum */
        }
    }
    /* MaxScorePicker.Pick (maxscore/picker.go:87-115): shuffle + stable sort desc + take 1 ==
     * a uniformly random member of the arg-max set; the oracle reports the whole set. */
    double mx = 0; int have = 0;
    for (int e = 0; e < n; e++) {
        if (!cand[e]) continue;
        if (!have || out_scores[e] > mx) { mx = out_scores[e]; have = 1; }
    }
    int cnt = 0; int first = -1;
    for (int e = 0; e < n; e++) {
        if (!cand[e]) continue;
        if (!(out_scores[e] < mx) && !(out_scores[e] > mx)) {
            if (first < 0) first = e;
            if (argmax_set) argmax_set[cnt] = e;
            cnt++;
        }
    }
    if (tie_seed && cnt > 1) {                           /* the build's reproducible stand-in for the shuffle */
        uint32_t k = orc_tie_rank(tie_seed, tie_key, (uint32_t)cnt);
        for (int e = 0; e < n; e++) {
            if (!cand[e]) continue;
            if (!(out_scores[e] < mx) && !(out_scores[e] > mx)) {
                if (k == 0) { first = e; break; }
                k--;
            }
        }
    }
    if (out_max) *out_max = mx;
    if (out_pick) *out_pick = first;
    if (tk && tk->k > 0) {
        /* the first k of shuffle + stable sort (maxscore/picker.go:91-110): group after group in descending score
         * order; col[] is free now and marks the endpoints already emitted */
        for (int e = 0; e < n; e++) col[e] = 0;
        double v = 0; int have_v = 0; int rem = 0;
        for (int j = 0; j < tk->k; j++) {
            if (rem == 0) {
                double best = 0; int hb = 0;
                for (int e = 0; e < n; e++) {
                    if (!cand[e] || col[e] != 0) continue;
                    if (have_v && !(out_scores[e] < v)) continue;
                    if (!hb || out_scores[e] > best) { best = out_scores[e]; hb = 1; }
                }
                if (!hb) break;
                v = best; have_v = 1;
                for (int e = 0; e < n; e++)
                    if (cand[e] && col[e] == 0 && !(out_scores[e] < v) && !(out_scores[e] > v)) rem++;
            }
            uint32_t rank = tie_seed ? orc_tie_rank(tie_seed + (uint64_t)j * 0x9E3779B97F4A7C15ULL, tie_key, (uint32_t)rem) : 0;
            int pick = -1;
            for (int e = 0; e < n; e++) {
                if (!cand[e] || col[e] != 0) continue;
                if (out_scores[e] < v || out_scores[e] > v) continue;
                if (rank == 0) { pick = e; break; }
                rank--;
            }
            if (pick < 0) break;
            col[pick] = 1;
            rem--;
            if (tk->picks) tk->picks[j] = pick;
            if (tk->scores) tk->scores[j] = v;
            tk->n = j + 1;
        }
    }
    return cnt;
}

/* PrefixBasedPDDecider.disaggregate -- prefix_based_pd_decider.go:99-149, :152-167 */
int orc_pd_decide(int64_t non_cached_tokens, int64_t input_len_bytes, int32_t match_blocks,
                  int32_t block_size_tokens) {
    if (non_cached_tokens == 0) return 0;                       /* :105-107 */
    int64_t input_tokens = input_len_bytes / 4;                 /* :158 len(raw)/AverageCharactersPerToken */
    if (input_tokens < non_cached_tokens) return 0;             /* :117-120 */
    int64_t hit = (int64_t)match_blocks * (int64_t)block_size_tokens;   /* :135 */
    int64_t non_cached = input_tokens - hit;                    /* :137 */
    if (non_cached < non_cached_tokens) return 0;               /* :143-146 */
    return 1;
}

/* Scheduler.Schedule + single / disagg profile handlers */
static void schedule_scratch(const orc_profile *primary, const orc_profile *prefill, const orc_pool *pool,
                             const int32_t *match, int32_t total, int32_t block_size_tokens, int64_t input_len_bytes,
                             int64_t non_cached_tokens, int always_disagg, double *scratch_scores, orc_decision *out,
                             uint8_t *cand, double *col, uint64_t tie_seed, uint64_t tie_req,
                             const orc_profile *encode, int multimodal, topk_req *tk);

void orc_schedule(const orc_profile *primary, const orc_profile *prefill, const orc_pool *pool,
                  const int32_t *match, int32_t total, int32_t block_size_tokens,
                  int64_t input_len_bytes, int64_t non_cached_tokens, int always_disagg,
                  double *scratch_scores, orc_decision *out) {
    uint8_t *cand = (uint8_t *)malloc((size_t)pool->n + 1);
    double *col = (double *)malloc(sizeof(double) * ((size_t)pool->n + 1));
    schedule_scratch(primary, prefill, pool, match, total, block_size_tokens, input_len_bytes, non_cached_tokens,
                     always_disagg, scratch_scores, out, cand, col, 0, 0, NULL, 0, NULL);
    free(cand); free(col);
}

static void schedule_scratch(const orc_profile *primary, const orc_profile *prefill, const orc_pool *pool,
                             const int32_t *match, int32_t total, int32_t block_size_tokens, int64_t input_len_bytes,
                             int64_t non_cached_tokens, int always_disagg, double *scratch_scores, orc_decision *out,
                             uint8_t *cand, double *col, uint64_t tie_seed, uint64_t tie_req,
                             const orc_profile *encode, int multimodal, topk_req *tk) {
    memset(out, 0, sizeof *out);
    if (tk) for (int q = 1; q < 3; q++)          /* stages that do not run leave empty lists */
        for (int j = 0; j < tk[q].k; j++) { if (tk[q].picks) tk[q].picks[j] = -1; if (tk[q].scores) tk[q].scores[j] = 0; }
    out->pick = -1; out->prefill_pick = -1; out->encode_pick = -1;
    double mx; int32_t pick;
    int ties = profile_run_scratch(primary, pool, match, total, scratch_scores, &mx, &pick, NULL, cand, col, tie_seed, 4 * tie_req, tk ? &tk[0] : NULL);
    if (ties == 0) {            /* disagg ProcessResults :335-338 / single ProcessResults: error */
        out->status = -1;
        return;
    }
    out->pick = pick; out->tie_count = ties; out->score = mx;
    out->encode_pick = -1;
    if (encode && multimodal) { /* disagg_profile_handler.go:284-295 + always_disagg_mm_decider.go:47-49 */
        double emx; int32_t epick;
        out->encode_ran = 1;
        int et = profile_run_scratch(encode, pool, match, total, scratch_scores, &emx, &epick, NULL, cand, col, tie_seed, 4 * tie_req + 2, tk ? &tk[2] : NULL);
        if (et > 0) { out->encode_pick = epick; out->encode_tie_count = et; out->encode_score = emx; }
    }
    if (prefill) {              /* disagg_profile_handler.go:296-308 */
        int go = always_disagg ? 1
                               : orc_pd_decide(non_cached_tokens, input_len_bytes, match[pick], block_size_tokens);
        out->prefill_ran = go;
        if (go) {
            double pmx; int32_t ppick;
            int pt = profile_run_scratch(prefill, pool, match, total, scratch_scores, &pmx, &ppick, NULL, cand, col, tie_seed, 4 * tie_req + 1, tk ? &tk[1] : NULL);
            if (pt > 0) { out->prefill_pick = ppick; out->prefill_tie_count = pt; out->prefill_score = pmx; }
        }
    }
}

static void cycle_scratch(const orc_cycle_cfg *cfg, const orc_indexer *ix, const orc_profile *primary,
                          const orc_profile *prefill, const orc_pool *pool, const uint8_t *prompt, size_t prompt_len,
                          uint64_t *scratch_hashes, int32_t *scratch_match, double *scratch_scores, orc_decision *out,
                          int32_t *out_total, uint8_t *cand, double *col, uint64_t tie_req) {
    int cap = cfg->max_prefix_blocks + 1;
    int total = orc_hash_prompt(prompt, prompt_len, cfg->model, cfg->model_len, NULL, 0,
                                cfg->block_size_tokens, cfg->max_prefix_blocks, scratch_hashes, cap);
    memset(scratch_match, 0, sizeof(int32_t) * (size_t)pool->n);
    orc_match_longest_prefix(ix, scratch_hashes, total, scratch_match, pool->n);
    const int64_t row = (int64_t)(tie_req - cfg->tie_base);
    topk_req tk[3];
    const int k = cfg->topk > 1 ? cfg->topk : 0;
    tk[0] = (topk_req){k, cfg->topk_primary ? cfg->topk_primary + row * k : NULL,
                       cfg->topk_primary_scores ? cfg->topk_primary_scores + row * k : NULL, 0};
    tk[1] = (topk_req){k, cfg->topk_prefill ? cfg->topk_prefill + row * k : NULL, NULL, 0};
    tk[2] = (topk_req){k, cfg->topk_encode ? cfg->topk_encode + row * k : NULL, NULL, 0};
    schedule_scratch(primary, prefill, pool, scratch_match, total, cfg->block_size_tokens, (int64_t)prompt_len,
                     cfg->non_cached_tokens, cfg->always_disagg, scratch_scores, out, cand, col, cfg->tie_seed, tie_req,
                     cfg->encode, cfg->multimodal ? cfg->multimodal[row] : 0, k ? tk : NULL);
    if (out_total) *out_total = total;
}

void orc_cycle(const orc_cycle_cfg *cfg, const orc_indexer *ix, const orc_profile *primary,
               const orc_profile *prefill, const orc_pool *pool, const uint8_t *prompt, size_t prompt_len,
               uint64_t *scratch_hashes, int32_t *scratch_match, double *scratch_scores,
               orc_decision *out, int32_t *out_total) {
    int cap = cfg->max_prefix_blocks + 1;
    int total = orc_hash_prompt(prompt, prompt_len, cfg->model, cfg->model_len, NULL, 0,
                                cfg->block_size_tokens, cfg->max_prefix_blocks, scratch_hashes, cap);
    memset(scratch_match, 0, sizeof(int32_t) * (size_t)pool->n);
    orc_match_longest_prefix(ix, scratch_hashes, total, scratch_match, pool->n);
    orc_schedule(primary, prefill, pool, scratch_match, total, cfg->block_size_tokens,
                 (int64_t)prompt_len, cfg->non_cached_tokens, cfg->always_disagg, scratch_scores, out);
    if (out_total) *out_total = total;
}

typedef struct {
    const orc_cycle_cfg *cfg; const orc_indexer *ix; const orc_profile *primary; const orc_profile *prefill;
    const orc_pool *pool; const uint8_t *data; const uint64_t *offsets; int64_t lo, hi;
    orc_decision *out; int32_t *out_total;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    int n = j->pool->n;
    uint64_t *hashes = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)j->cfg->max_prefix_blocks + 2));
    int32_t *match = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 1));
    double *scores = (double *)malloc(sizeof(double) * ((size_t)n + 1));
    uint8_t *cand = (uint8_t *)malloc((size_t)n + 1);
    double *col = (double *)malloc(sizeof(double) * ((size_t)n + 1));
    for (int64_t r = j->lo; r < j->hi; r++) {
        cycle_scratch(j->cfg, j->ix, j->primary, j->prefill, j->pool, j->data + j->offsets[r],
                      (size_t)(j->offsets[r + 1] - j->offsets[r]), hashes, match, scores, &j->out[r],
                      j->out_total ? &j->out_total[r] : NULL, cand, col, j->cfg->tie_base + (uint64_t)r);
    }
    free(hashes); free(match); free(scores); free(cand); free(col);
    return NULL;
}

void orc_cycle_batch(const orc_cycle_cfg *cfg, const orc_indexer *ix, const orc_profile *primary,
                     const orc_profile *prefill, const orc_pool *pool, const uint8_t *data,
                     const uint64_t *offsets, int64_t R, int n_threads, orc_decision *out,
                     int32_t *out_total) {
    if (n_threads <= 0) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    if ((int64_t)n_threads > R) n_threads = R > 0 ? (int)R : 1;
    batch_job *jobs = (batch_job *)malloc(sizeof(batch_job) * (size_t)n_threads);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; t++) {
        jobs[t] = (batch_job){cfg, ix, primary, prefill, pool, data, offsets,
                              R * t / n_threads, R * (t + 1) / n_threads, out, out_total};
        if (n_threads == 1) batch_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    if (n_threads > 1) for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    free(jobs); free(th);
}
