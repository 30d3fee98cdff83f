/*
 * epp_oracle.h -- CPU ORACLE for the EPP scheduling-cycle hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It is a plain-C restatement of
 * the reference's Go algorithm (llm-d/llm-d-inference-scheduler @ 520af478) for
 * the path named by BASELINE.json:north_star.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * library (libepp_engine.so) never links, loads or calls anything in oracle/.
 *
 * Parity pinning: the reference cannot be built here (no Go toolchain), so the
 * oracle is pinned against (a) every known-answer test the reference's own unit
 * tests hold for this path (SURVEY.md App. B.1 -> tests/test_oracle_kat.py) and
 * (b) XXH64 vectors from the independent python-xxhash 3.7.0 implementation
 * (SURVEY.md App. B.2 -> tests/golden/xxh64_vectors.json).  Hash VALUES are not
 * pinned by the reference's tests (they only pin counts/equalities); they are
 * pinned by the public XXH64 specification that cespare/xxhash v2.3.0 implements.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference checkout root).
 */
#ifndef EPP_ORACLE_H
#define EPP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- A.1 XXH64 (github.com/cespare/xxhash/v2 v2.3.0, go.mod:10; seed 0 in the reference) ---- */
uint64_t orc_xxh64(const void *data, size_t len, uint64_t seed);

/* ---- A.2 hashPrompt: approximateprefix/hashing.go:35-99 ----
 * data/len        user-input bytes (getUserInputBytes, hashing.go:107-136)
 * model, salt     request.TargetModel, Body.CacheSalt() (salt_len 0 => not written, hashing.go:75-77)
 * returns number of hashes written to out (<= out_cap); 0 when the prompt is shorter than a block
 * (hashing.go:58-61) or block_size_tokens*4 <= 0 (hashing.go:51-56). */
int orc_hash_prompt(const uint8_t *data, size_t len, const uint8_t *model, size_t model_len,
                    const uint8_t *salt, size_t salt_len, int block_size_tokens, int max_prefix_blocks,
                    uint64_t *out, int out_cap);

/* ---- A.3 indexer: approximateprefix/indexer.go:32-182 ---- */
typedef struct orc_indexer orc_indexer;
orc_indexer *orc_indexer_new(int default_lru_size);                 /* indexer.go:40-49 */
void orc_indexer_free(orc_indexer *ix);
/* indexer.go:52-83.  num_gpu_blocks<=0 => default size (indexer.go:59-62). */
void orc_indexer_add(orc_indexer *ix, const uint64_t *hashes, int n, uint32_t server, int num_gpu_blocks);
/* indexer.go:86-102: copies the pod set (like the Go deep copy); returns its size. */
int orc_indexer_get(const orc_indexer *ix, uint64_t hash, uint32_t *out, int out_cap);
void orc_indexer_remove_pod(orc_indexer *ix, uint32_t server);      /* indexer.go:167-182 */
int orc_indexer_pods(const orc_indexer *ix, uint32_t *out, int out_cap); /* indexer.go:185-195 */
int orc_indexer_lru_len(const orc_indexer *ix, uint32_t server);    /* lruCache.Len() */
/* Export every (hash, server) pair of hashToPods; returns the pair count (call with cap 0 to size). */
size_t orc_indexer_export(const orc_indexer *ix, uint64_t *hashes, uint32_t *servers, size_t cap);
/* Bulk-load a frozen snapshot of (hash, server) pairs without LRU bookkeeping (tests / bench only). */
void orc_indexer_load_pairs(orc_indexer *ix, const uint64_t *hashes, const uint32_t *servers, size_t n);

/* matchLongestPrefix: approximateprefix/plugin.go:214-230.
 * Dense form: counts[s] for s < n_servers (servers with id >= n_servers keep the walk alive but are
 * not reported -- SURVEY App. C.5).  Returns the number of blocks walked before the stop. */
int orc_match_longest_prefix(const orc_indexer *ix, const uint64_t *hashes, int n,
                             int32_t *counts, int n_servers);

/* ---- A.4 scorers (all float64, unfused) ---- */
enum {
    ORC_SCORER_PREFIX = 0,       /* scorer/prefix/plugin.go:95-117                       */
    ORC_SCORER_KV_UTIL = 1,      /* scorer/kvcacheutilization/kvcache_utilization.go:76-82 */
    ORC_SCORER_QUEUE = 2,        /* scorer/queuedepth/queue.go:78-108                    */
    ORC_SCORER_LOAD_AWARE = 3,   /* scorer/loadaware/load_aware.go:84-100 (param = threshold) */
    ORC_SCORER_EXTERNAL = 4,     /* host-computed column (param = column index); used for the
                                    lora-affinity column of scheduler_test.go:77-141      */
    ORC_SCORER_RUNNING = 5,      /* scorer/runningrequests/runningrequest.go:78-108      */
    ORC_SCORER_TOKEN_LOAD = 6,   /* scorer/tokenload/token_load.go:84-112: column = ext column of InFlightLoad.Tokens,
                                    param = queueThresholdTokens (<= 0: 4194304, :27-28, :60-62) */
    ORC_SCORER_ACTIVE_REQUEST = 7, /* scorer/activerequest/active_request.go:140-173: column = ext column of
                                    InFlightLoad.Requests, param = maxBusyScore, param2 = idleThreshold (:83-93) */
    ORC_SCORER_LORA_AFFINITY = 8  /* scorer/loraaffinity/lora_affinity.go:76-100; reads orc_pool.lora_* (the pool as
                                    seen by a request whose TargetModel is the adapter lora_state describes) */
};

/* Role filters: filter/bylabel/roles.go:46-70, filter.go:104-117. */
enum {
    ORC_ROLE_NONE = 0,           /* no llm-d.ai/role label */
    ORC_ROLE_DECODE = 1,
    ORC_ROLE_PREFILL = 2,
    ORC_ROLE_PREFILL_DECODE = 3,
    ORC_ROLE_BOTH = 4,
    ORC_ROLE_ENCODE = 5,
    ORC_ROLE_ENCODE_PREFILL = 6,
    ORC_ROLE_ENCODE_PREFILL_DECODE = 7,
    ORC_ROLE_OTHER = 8           /* label present with an unknown value */
};
enum { ORC_FILTER_NONE = 0, ORC_FILTER_DECODE = 1, ORC_FILTER_PREFILL = 2, ORC_FILTER_ENCODE = 3 };
int orc_role_filter_keeps(int filter, int role);

typedef struct {
    int32_t kind;
    int32_t column;              /* ext column read by TOKEN_LOAD / ACTIVE_REQUEST */
    double weight;               /* WeightedScorer.weight, weighted_scorer.go:24-40 */
    double param;
    double param2;
} orc_scorer;

#define ORC_MAX_SCORERS 8
typedef struct {
    int32_t filter;              /* ORC_FILTER_* role filter, first in the profile's filter chain */
    int32_t n_scorers;
    orc_scorer scorers[ORC_MAX_SCORERS];
    /* prefix-cache-affinity-filter after the role filter (filter/prefixcacheaffinity/plugin.go:105-151);
     * affinity_threshold <= 0: not configured (:108). */
    double affinity_threshold;
    double exploration_probability;
    double max_ttft_penalty_ms;
    int32_t ttft_column;         /* ext column with LatencyPredictionInfo.TTFT per endpoint; < 0: attribute absent */
    int32_t _pad;
} orc_profile;

/* Pool-state snapshot (fwkdl.Metrics + role label), struct-of-arrays, n endpoints. */
typedef struct {
    int32_t n;
    int32_t n_ext_cols;
    const uint8_t *role;         /* ORC_ROLE_*; 0xFF = slot not in the pool                 */
    const double *kv_usage;      /* KVCacheUsagePercent (a fraction, extractor.go:136)      */
    const int32_t *waiting;      /* WaitingQueueSize                                        */
    const int32_t *running;      /* RunningRequestsSize                                     */
    const double *ext;           /* [n_ext_cols][n] host-computed score columns             */
    /* LoRA residency for the request's TargetModel (NULL = no adapter information):        */
    const uint8_t *lora_state;   /* [n] 0 = not resident, 1 = in ActiveModels, 2 = in WaitingModels */
    const int32_t *lora_max;     /* [n] MaxActiveModels                                     */
    const int32_t *lora_loaded;  /* [n] len(ActiveModels) + len(WaitingModels)              */
} orc_pool;

/* SchedulerProfile.Run: scheduling/scheduler_profile.go:117-128 (filters -> scorers -> picker).
 * match[e], total: PrefixCacheMatchInfo of the request (data_types.go:27-54).
 * out_scores[e]: weighted sum for candidates, NaN-free sentinel -1 for filtered-out endpoints.
 * Returns the size of the arg-max set (0 => "no endpoints available", scheduler_profile.go:119-121);
 * *out_max = max score; *out_pick = LOWEST-index member of the arg-max set (the deterministic
 * representative of maxscore/picker.go:87-115's random tie-break); argmax_set (cap n) optional. */
int orc_profile_run(const orc_profile *p, const orc_pool *pool, const int32_t *match, int32_t total,
                    double *out_scores, double *out_max, int32_t *out_pick, int32_t *argmax_set);

/* The same run with the build's reproducible random draws (tie_seed != 0; key = 4 * request ordinal + profile index) and
 * the first k endpoints of MaxScorePicker.Pick (maxscore/picker.go:87-115: shuffle, stable sort by score descending,
 * first k): groups of equal score in descending order, inside a group pick j is the member of rank
 * orc_tie_rank(tie_seed + j * 0x9E3779B97F4A7C15, key, members left) among the members not emitted yet (ascending slot
 * order), or the lowest remaining slot when tie_seed == 0.  topk_picks / topk_scores: k entries, padded with -1 / 0.
 * Returns the size of the arg-max set; *out_n_picks = number of endpoints emitted. */
int orc_profile_run_topk(const orc_profile *p, const orc_pool *pool, const int32_t *match, int32_t total,
                         uint64_t tie_seed, uint64_t tie_key, int k, int32_t *topk_picks, double *topk_scores,
                         int32_t *out_n_picks, double *out_scores);
/* Exploration draw of the affinity filter: u in [0, 1) = (mix64((seed ^ 0xA0761D6478BD642F) ^ mix64(key)) >> 11) * 2^-53. */
double orc_explore_u(uint64_t seed, uint64_t key);

/* Individual scorer columns (plugin-parity mode for Scorer.Score).  cand[e]!=0 marks the filtered
 * candidate list the scorer sees; non-candidates get 0. */
void orc_score_column(const orc_scorer *s, const orc_pool *pool, const uint8_t *cand,
                      const int32_t *match, int32_t total, double *out);

/* PrefixBasedPDDecider.disaggregate: profilehandler/disagg/prefix_based_pd_decider.go:99-149. */
int orc_pd_decide(int64_t non_cached_tokens, int64_t input_len_bytes, int32_t match_blocks,
                  int32_t block_size_tokens);

typedef struct {
    int32_t status;              /* 0 ok; -1 no decode/primary endpoint (Schedule error)   */
    int32_t pick;                /* primary (decode) pick, lowest index of the arg-max set  */
    int32_t tie_count;
    int32_t prefill_pick;        /* -1 = no prefill stage result                            */
    int32_t prefill_tie_count;
    int32_t prefill_ran;         /* decider said disaggregate                               */
    double score;
    double prefill_score;
    int32_t encode_pick;         /* -1 = no encode stage result (disagg_profile_handler.go:284-295)    */
    int32_t encode_tie_count;
    int32_t encode_ran;          /* the encode decider said disaggregate (multimodal request)           */
    int32_t _pad;
    double encode_score;
} orc_decision;

/* Scheduler.Schedule with the single-profile handler (single_profile_handler.go:66-99) when
 * prefill==NULL, else the disagg handler decode -> decider -> prefill
 * (disagg_profile_handler.go:246-354).  non_cached_tokens: decider parameter (0 disables).
 * always_disagg!=0 models always-disagg-pd-decider (always_disagg_pd_decider.go:48-50). */
void orc_schedule(const orc_profile *primary, const orc_profile *prefill, const orc_pool *pool,
                  const int32_t *match, int32_t total, int32_t block_size_tokens,
                  int64_t input_len_bytes, int64_t non_cached_tokens, int always_disagg,
                  double *scratch_scores, orc_decision *out);

/* Whole-cycle driver for one request against a frozen index: hashPrompt ->
 * matchLongestPrefix -> schedule.  scratch_* sized n_servers.  out_hashes (cap >= max blocks) optional. */
typedef struct {
    int32_t block_size_tokens;
    int32_t max_prefix_blocks;
    int64_t non_cached_tokens;
    int32_t always_disagg;
    int32_t _pad;
    const uint8_t *model;
    size_t model_len;
    uint64_t tie_seed;           /* 0: lowest-slot representative; else the build's reproducible random tie rule   */
    uint64_t tie_base;           /* ordinal of request 0 of the batch (orc_cycle_batch: request r has ordinal base+r) */
    const orc_profile *encode;   /* the "encode" profile of the disagg handler, or NULL                             */
    const uint8_t *multimodal;   /* [R] hasMultimodalContent per request (multimodal_helpers.go), or NULL; orc_cycle_batch only */
    int32_t topk;                /* maxNumOfEndpoints of the pickers (<= 1: one endpoint); orc_cycle_batch only       */
    int32_t _pad2;
    int32_t *topk_primary;       /* [R][topk] (-1 padded) or NULL                                                    */
    double *topk_primary_scores; /* [R][topk] or NULL                                                                */
    int32_t *topk_prefill;       /* [R][topk] or NULL                                                                */
    int32_t *topk_encode;        /* [R][topk] or NULL                                                                */
} orc_cycle_cfg;
/* Rank (in ascending slot order) of the arg-max-set member the build picks when tie_seed != 0:
 * ((mix64(seed ^ mix64(key)) >> 32) * n) >> 32 with the SplitMix64 output function, key = 4 * ordinal + profile index. */
uint32_t orc_tie_rank(uint64_t seed, uint64_t key, uint32_t n);
void orc_cycle(const orc_cycle_cfg *cfg, const orc_indexer *ix, const orc_profile *primary,
               const orc_profile *prefill, const orc_pool *pool, const uint8_t *prompt, size_t prompt_len,
               uint64_t *scratch_hashes, int32_t *scratch_match, double *scratch_scores,
               orc_decision *out, int32_t *out_total);

/* Multi-threaded batch of orc_cycle over R prompts (offsets[R+1] into data) -- used only as the
 * timed CPU baseline of bench.py.  n_threads<=0 => 1. */
void orc_cycle_batch(const orc_cycle_cfg *cfg, const orc_indexer *ix, const orc_profile *primary,
                     const orc_profile *prefill, const orc_pool *pool, const uint8_t *data,
                     const uint64_t *offsets, int64_t R, int n_threads, orc_decision *out,
                     int32_t *out_total);

#ifdef __cplusplus
}
#endif
#endif
