/*
 * epp_engine.h -- C ABI of libepp_engine.so, the B200-native Endpoint-Picker scoring engine.
 *
 * The reference (llm-d/llm-d-inference-scheduler @ 520af478, pure Go, CGO_ENABLED=0) has NO FFI; this
 * header is the boundary the build introduces.  Every entry point names the reference
 * interface it replaces (file:line relative to the reference root).  A thin cgo shim (go/eppcuda/,
 * INTEGRATION.md) binds exactly these symbols and keeps the Go plugin interfaces unchanged:
 *
 *   requestcontrol.DataProducer.Produce   pkg/epp/framework/interface/requestcontrol/plugins.go:69-72
 *   scheduling.Filter / Scorer / Picker    pkg/epp/framework/interface/scheduling/plugins.go:59-78
 *   requestcontrol.PreRequest              pkg/epp/framework/interface/requestcontrol/plugins.go:36-39
 *   Scheduler.Schedule                     pkg/epp/scheduling/scheduler.go:54-102
 *
 * Conventions: plain C, no exceptions cross the boundary.  Every function returns int32_t: 0 = EPP_OK,
 * negative = epp_status error; epp_last_error() returns a thread-local message.  The caller owns every
 * host buffer and may free it on return (cgo forbids retaining Go pointers); the engine owns all device
 * memory.  Every entry point is thread-safe (calls on one engine are serialised internally; goroutines
 * should batch requests before calling).  Handles are opaque.
 *
 * Streams: the engine launches on its own CUDA streams.  With EPP_BATCH_DEVICE_PTRS (and in the epp_shard_* calls)
 * every device buffer passed IN must be complete before the call (the engine does not wait on the caller's
 * streams); every output is complete when the call returns (EPP_BATCH_ASYNC: when epp_synchronize returns).
 *
 * Endpoint identity: the reference keys servers by NamespacedName strings.  The shim maps each endpoint
 * to a dense SLOT id in [0, max_endpoints); all arrays below are indexed by slot id.
 *
 * Batch semantics: the reference schedules one request at a time; a batch evaluates
 * all R requests against ONE frozen snapshot (pool state + prefix index).  Index updates (PreRequest)
 * are applied between batches.
 *
 * Tie rule: the reference's max-score picker shuffles with math/rand before its stable sort, i.e. it returns a
 * uniformly random member of the arg-max set (picker/maxscore/picker.go:91-102).  The engine always returns the bit
 * pattern of the max score and the size of the set, and as the pick
 *   epp_config.tie_seed == 0: the LOWEST slot id of the set (deterministic; what the parity tests compare), or
 *   epp_config.tie_seed != 0: the member of rank ((mix64(tie_seed ^ mix64(key)) >> 32) * |set|) >> 32 of the set in
 *       ascending slot order, mix64 = the SplitMix64 output function, key = 4 * (ordinal of the request since the
 *       engine was created = epp_stats.n_decisions before the call + index in the batch) + profile index (0 primary /
 *       decode, 1 prefill, 2 encode): uniform over the set like the reference, reproducible by the oracle
 *       (oracle/epp_oracle.c: orc_tie_rank).  Without it every tied request of a batch (cold prompts on a balanced
 *       pool) would land on the same endpoint.  The P/D decider reads the match count of the member picked, as the
 *       reference does (disagg_profile_handler.go:300).  Endpoint-sharded mode always uses the lowest slot.
 *
 * Top-k (epp_config.pick_k > 1, epp_schedule_topk): maxscore/picker.go:91-110 shuffles, stable-sorts by score descending
 * and keeps the first k, i.e. the endpoints come out group by group in descending score order and every group of equal
 * scores in uniformly random order.  The engine emits pick j (j = 0 .. k-1) as the member of rank
 *   tie_seed == 0: 0 (the lowest remaining slot), or
 *   tie_seed != 0: ((mix64((tie_seed + j * 0x9E3779B97F4A7C15) ^ mix64(key)) >> 32) * n_j) >> 32
 * among the n_j members of the current group not emitted yet (ascending slot order); pick 0 is the pick of the rule above.
 *
 * Exploration draw of the prefix-cache-affinity-filter (rand.Float64() < explorationProbability,
 * filter/prefixcacheaffinity/plugin.go:113): u = (mix64((tie_seed ^ 0xA0761D6478BD642F) ^ mix64(key)) >> 11) * 2^-53 when
 * tie_seed != 0; with tie_seed == 0 (the deterministic mode) the filter never explores.
 */
#ifndef EPP_ENGINE_H
#define EPP_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define EPP_API __attribute__((visibility("default")))
#else
#define EPP_API
#endif

#define EPP_ABI_VERSION 5
#define EPP_MAX_SCORERS 8
#define EPP_NO_ENDPOINT 0xFFFFFFFFu

typedef struct epp_engine epp_engine;

typedef enum {
    EPP_OK = 0,
    EPP_ERR_INVALID = -1,     /* bad argument / configuration                                       */
    EPP_ERR_CUDA = -2,        /* a CUDA runtime call failed (message has the CUDA error string)     */
    EPP_ERR_NO_DEVICE = -3,   /* no usable CUDA device: the engine NEVER falls back to a CPU path   */
    EPP_ERR_CAPACITY = -4,    /* batch / index larger than the configured or available capacity     */
    EPP_ERR_STATE = -5,       /* call order violation (e.g. schedule before pool_set)               */
    EPP_ERR_NCCL = -6
} epp_status;

/* Scorer kinds (plugin type strings of the reference in comments). */
typedef enum {
    EPP_SCORER_PREFIX = 0,     /* "prefix-cache-scorer"          scorer/prefix/plugin.go:95-117               */
    EPP_SCORER_KV_UTIL = 1,    /* "kv-cache-utilization-scorer"  scorer/kvcacheutilization/kvcache_utilization.go:76-82 */
    EPP_SCORER_QUEUE = 2,      /* "queue-scorer"                 scorer/queuedepth/queue.go:78-108            */
    EPP_SCORER_LOAD_AWARE = 3, /* "load-aware-scorer"            scorer/loadaware/load_aware.go:84-100; param = threshold */
    EPP_SCORER_EXTERNAL = 4,   /* host-computed per-endpoint column (param = column index): lets the Go
                                  side keep any other Scorer (lora-affinity, session-affinity, ...) */
    EPP_SCORER_RUNNING = 5,    /* "running-requests-size-scorer" scorer/runningrequests/runningrequest.go:78-108 */
    EPP_SCORER_TOKEN_LOAD = 6, /* "token-load-scorer"            scorer/tokenload/token_load.go:84-112; column = ext column
                                  holding InFlightLoad.Tokens, param = queueThresholdTokens (<= 0: 4194304)       */
    EPP_SCORER_ACTIVE_REQUEST = 7, /* "active-request-scorer"    scorer/activerequest/active_request.go:140-173; column = ext
                                  column holding InFlightLoad.Requests, param = maxBusyScore (outside [0,1]: 1.0),
                                  param2 = idleThreshold (< 0: 0)                                                 */
    EPP_SCORER_LORA_AFFINITY = 8  /* "lora-affinity-scorer"      scorer/loraaffinity/lora_affinity.go:76-100; the adapter is the
                                  request's model id; residency comes from epp_pool_set_lora                     */
} epp_scorer_kind;

/* Value of the llm-d.ai/role label (filter/bylabel/roles.go:25-44). */
typedef enum {
    EPP_ROLE_NONE = 0,         /* label absent */
    EPP_ROLE_DECODE = 1,
    EPP_ROLE_PREFILL = 2,
    EPP_ROLE_PREFILL_DECODE = 3,
    EPP_ROLE_BOTH = 4,
    EPP_ROLE_ENCODE = 5,
    EPP_ROLE_ENCODE_PREFILL = 6,
    EPP_ROLE_ENCODE_PREFILL_DECODE = 7,
    EPP_ROLE_OTHER = 8         /* label present, value not in the tables */
} epp_role;

/* Role filters (filter/bylabel/roles.go:46-70, filter.go:104-117). */
typedef enum {
    EPP_FILTER_NONE = 0,
    EPP_FILTER_DECODE = 1,     /* "decode-filter": decode, prefill-decode, both, encode-prefill-decode, or no label */
    EPP_FILTER_PREFILL = 2,    /* "prefill-filter": prefill, encode-prefill, prefill-decode, both, e-p-d; label required */
    EPP_FILTER_ENCODE = 3      /* "encode-filter": encode, encode-prefill, e-p-d; label required */
} epp_filter_kind;

/* Profile handlers. */
typedef enum {
    EPP_HANDLER_SINGLE = 0,    /* profilehandler/single/single_profile_handler.go:66-99     */
    EPP_HANDLER_DISAGG = 1     /* profilehandler/disagg/disagg_profile_handler.go:246-354 (decode -> decider -> prefill) */
} epp_handler_kind;

typedef struct {
    int32_t kind;              /* epp_scorer_kind */
    int32_t column;            /* ext column read by TOKEN_LOAD / ACTIVE_REQUEST (raw attribute values, not scores) */
    double weight;             /* WeightedScorer.weight, scheduling/weighted_scorer.go:24-40 */
    double param;
    double param2;
} epp_scorer_cfg;

/* One SchedulerProfile (scheduling/scheduler_profile.go:117-128): role filter -> [prefix-cache-affinity-filter] ->
 * scorers IN ORDER -> max-score picker. */
typedef struct {
    int32_t filter;            /* epp_filter_kind */
    int32_t n_scorers;
    epp_scorer_cfg scorers[EPP_MAX_SCORERS];
    /* "prefix-cache-affinity-filter" after the role filter (filter/prefixcacheaffinity/plugin.go:105-151): with
     * probability 1 - exploration_probability the candidates are narrowed to the STICKY endpoints (matchBlocks /
     * totalBlocks >= affinity_threshold) unless there is none, or the best predicted TTFT among them exceeds the best
     * among the others by more than max_ttft_penalty_ms.  The scorers then normalise over the narrowed set
     * (queue / running min-max, active-request max).  affinity_threshold <= 0: no such filter (plugin.go:108).
     * The reference's defaults are 0.80 / 0.01 / 5000 (plugin.go:62-66). */
    double affinity_threshold;
    double exploration_probability;    /* [0, 1]                                                              */
    double max_ttft_penalty_ms;        /* >= 0; 0 = always stick                                              */
    int32_t ttft_column;               /* ext column holding LatencyPredictionInfo.TTFT per endpoint; < 0 = no
                                          prediction attached (the gate then never breaks stickiness, :169-180) */
    int32_t reserved;
} epp_profile_cfg;

/* Mirrors the EndpointPickerConfig fields that define this path (apix/config/v1alpha1, and
 * approximateprefix/types.go:78-142 for the prefix producer defaults). */
typedef struct {
    uint32_t struct_size;          /* sizeof(epp_config), for ABI evolution                              */
    int32_t device;                /* CUDA device ordinal                                                */
    int32_t max_endpoints;         /* slot-id capacity                                                   */
    int32_t block_size_tokens;     /* blockSizeTokens (default 16); block bytes = 4 * this (hashing.go:49) */
    int32_t max_prefix_blocks;     /* maxPrefixBlocksToMatch (default 256)                               */
    int32_t lru_capacity_per_server; /* lruCapacityPerServer (default 31250)                             */
    int32_t handler;               /* epp_handler_kind                                                   */
    int32_t always_disagg;         /* always-disagg-pd-decider (always_disagg_pd_decider.go:48-50)        */
    int64_t non_cached_tokens;     /* prefix-based-pd-decider nonCachedTokens (0 disables)               */
    int32_t n_ext_cols;            /* number of EPP_SCORER_EXTERNAL columns                              */
    int32_t pick_k;                /* max-score-picker maxNumOfEndpoints (picker/common.go:36, maxscore/picker.go:104-115);
                                      0 or 1 = one endpoint; 2..64: epp_schedule_topk / epp_schedule_with_match_topk also
                                      return the k best of every profile that ran                                     */
    epp_profile_cfg primary;       /* the single profile, or the decode profile under EPP_HANDLER_DISAGG */
    epp_profile_cfg prefill;       /* the prefill profile (EPP_HANDLER_DISAGG only)                      */
    epp_profile_cfg encode;        /* the encode profile (EPP_HANDLER_DISAGG with encode_enabled != 0):
                                      disagg_profile_handler.go:284-295                                   */
    uint64_t tie_seed;             /* seed of every reproducible random draw (ties, top-k order, exploration); 0 = the
                                      deterministic mode: lowest slot of the arg-max set, no exploration              */
    int32_t encode_enabled;        /* an "encode" profile + always-disagg-multimodal-decider are configured */
    int32_t index_commit_interval_us; /* 0: every scheduling call first makes all earlier index updates (epp_index_add,
                                      epp_index_add_picked) visible -- exact sequential semantics, what the parity tests use;
                                      > 0: a scheduling call does that only if the last one that did is at least this long
                                      ago, i.e. picks become visible to later requests with a bounded delay, like the
                                      reference's asynchronous PreRequest (approximateprefix/plugin.go:189-194).
                                      epp_index_commit / epp_index_get always commit.                              */
    uint64_t reserved1[2];
} epp_config;

/* One routing decision (32 bytes).  status: 0 = ok, -1 = no endpoint available for the primary profile
 * (Schedule error: scheduler_profile.go:119-121 / disagg_profile_handler.go:335-338). */
typedef struct {
    int32_t status;
    uint32_t pick;             /* primary (decode) pick: lowest slot id of the arg-max set, or EPP_NO_ENDPOINT */
    double score;              /* max weighted score (bit-exact vs the reference arithmetic)        */
    uint32_t prefill_pick;     /* prefill pick, EPP_NO_ENDPOINT when the prefill stage did not run / found nobody */
    uint32_t tie_count;        /* size of the primary arg-max set                                   */
    int32_t total_blocks;      /* PrefixCacheMatchInfo.totalBlocks (= number of prefix hashes)      */
    int32_t match_blocks;      /* PrefixCacheMatchInfo.matchBlocks of the primary pick              */
} epp_decision;

/* Optional second record per decision (40 bytes): the other stages of the disagg handler, for the shim's
 * SchedulingResult.ProfileResults, tests and metrics. */
typedef struct {
    double prefill_score;
    uint32_t prefill_tie_count;
    uint32_t prefill_ran;      /* the P/D decider asked for the prefill stage                        */
    double encode_score;
    uint32_t encode_pick;      /* encode pick, EPP_NO_ENDPOINT when the stage did not run / found nobody */
    uint32_t encode_tie_count;
    uint32_t encode_ran;       /* the encode decider asked for the encode stage (multimodal request) */
    uint32_t reserved;
} epp_decision_detail;

#define EPP_BATCH_DEVICE_PTRS 1u   /* data / offsets / model_ids and all outputs are DEVICE pointers */
#define EPP_BATCH_ASYNC 2u         /* epp_schedule + DEVICE_PTRS: enqueue and return; outputs (and epp_stats.last_*)
                                    * are complete after epp_synchronize().  Batches enqueue in call order.       */
#define EPP_BATCH_LENGTHS_EXCEED_ROWS 4u /* lengths[r] may exceed the row offsets[r+1] - offsets[r]: a row then holds only the
                                    * first min(lengths[r], max_prefix_blocks * block_size_tokens * 4) bytes of its prompt --
                                    * everything hashPrompt reads (hashing.go:63-66) -- while lengths[r] stays the prompt's
                                    * true length (input of the P/D decider).  Saves shipping the tail of long prompts.  */

/* A batch of prompts.  Prompt r is data[offsets[r] .. offsets[r+1]) (bytes), or data[offsets[r] .. offsets[r] +
 * lengths[r]) when `lengths` is given (lets ragged prompts START on 32-byte boundaries, which selects the 256-bit-load
 * kernel; offsets[r] + lengths[r] <= offsets[r+1] must hold).  A uint32 token array viewed as little-endian bytes is
 * a prompt with 4 bytes/token (hashing.go:49, types.go:113). */
typedef struct {
    int64_t n_requests;
    const void *data;
    const uint64_t *offsets;       /* [n_requests+1]; NULL => uniform: prompt r = data[r*uniform_len ..) */
    uint64_t uniform_len;
    const uint32_t *model_ids;     /* [n_requests] ids from epp_model_register; NULL => model 0       */
    uint32_t flags;                /* EPP_BATCH_* */
    uint32_t reserved;
    const uint64_t *lengths;       /* [n_requests] or NULL */
    const uint8_t *multimodal;     /* [n_requests] or NULL: != 0 = the request carries image / video / audio content blocks
                                      (hasMultimodalContent, disagg/multimodal_helpers.go): input of the encode decider */
} epp_batch;

typedef struct {
    uint64_t n_batches, n_decisions;
    uint64_t index_pairs, index_hashes, index_slots;   /* device table occupancy */
    double last_h2d_ms, last_kernels_ms, last_d2h_ms;  /* CUDA-event times of the last batch */
    double last_hash_ms, last_match_pick_ms;
    uint64_t last_kernel_launches;
    uint64_t device_bytes;
    /* device-pointer batches only: CUDA-event time of each kernel of the last batch on the launch stream.
     * [0] prompt lengths  [1] block digests (dominant: reads every token once)  [2] chain  [3] match+score+pick */
    double last_kernel_ms[8];
    /* algorithmic work of the last batch: table probes P = sum_r min(stop_r+1, B_r) and
     * postings consumed M = sum_r sum_{i<stop_r} |servers(h_i)| */
    uint64_t last_probes, last_postings;
    /* write side (index_store.cu): the last epp_index_add_picked / queued-Add flush and the last read-table build */
    double last_index_apply_ms;        /* Add + eviction kernels incl. their two small read-backs (host clock)   */
    double last_index_build_ms;        /* export of the inverted map + bulk build of the read table (host clock) */
    uint64_t last_index_items;         /* hashes added by the last applied batch                               */
    uint64_t last_index_launches;      /* kernels launched by it                                               */
    uint64_t last_index_patched;       /* 1: the last commit patched the read table from the change log; 0: bulk build */
} epp_stats;

/* ---- lifecycle --------------------------------------------------------------------------------- */
EPP_API int32_t epp_abi_version(void);
EPP_API const char *epp_last_error(void);
/* plugin factories + EndpointPickerConfig (framework/interface/plugin/registry.go:25-30). */
EPP_API int32_t epp_engine_create(const epp_config *cfg, epp_engine **out);
EPP_API int32_t epp_engine_destroy(epp_engine *h);
EPP_API void epp_config_default(epp_config *cfg);   /* types.go:136-142 + config/loader/defaults.go:47-49,78-87 */

/* Pinned host staging buffers for the shim (C memory, so cgo may keep them). */
EPP_API int32_t epp_host_alloc(size_t bytes, void **out);
EPP_API int32_t epp_host_free(void *p);

/* request.TargetModel (+ Body.CacheSalt()): registers the seed h_{-1} = XXH64(model || salt)
 * (hashing.go:71-78), computed on the device.  Returns a small id used in epp_batch.model_ids. */
EPP_API int32_t epp_model_register(epp_engine *h, const uint8_t *model, size_t model_len, const uint8_t *salt,
                           size_t salt_len, uint32_t *out_model_id);
EPP_API int32_t epp_model_seed(epp_engine *h, uint32_t model_id, uint64_t *out_seed);

/* ---- pool state (fwkdl.Metrics framework/interface/datalayer/metrics.go:26-42, EndpointMetadata.Labels) --
 * Replaces the snapshot (called by the 50 ms scrape loop).  n entries; ids[i] is the slot id of entry i;
 * slots not listed are absent from the pool.  role: epp_role.  ext may be NULL when n_ext_cols == 0, else
 * [n_ext_cols][n].  Also derives, ON THE DEVICE, every request-independent quantity of the scorers
 * (queue min/max over each profile's candidates, per-scorer contributions, ordered base sums). */
EPP_API int32_t epp_pool_set(epp_engine *h, int32_t n, const uint32_t *ids, const uint8_t *role, const double *kv_usage,
                     const int32_t *waiting, const int32_t *running, const double *ext);

/* LoRA adapter residency for EPP_SCORER_LORA_AFFINITY (fwkdl.Metrics.ActiveModels / WaitingModels / MaxActiveModels,
 * scorer/loraaffinity/lora_affinity.go:76-100).  n entries by slot id: MaxActiveModels and len(ActiveModels) +
 * len(WaitingModels); n_members triples (endpoint slot, model id from epp_model_register, state 1 = active /
 * 2 = waiting) list where each adapter is resident.  Replaces the previous LoRA state and re-derives the scorer terms;
 * call it after epp_pool_set of the same scrape (slots absent here have capacity 0 and nothing resident). */
EPP_API int32_t epp_pool_set_lora(epp_engine *h, int32_t n, const uint32_t *ids, const int32_t *max_active_models,
                          const int32_t *n_models_loaded, int64_t n_members, const uint32_t *member_ep,
                          const uint32_t *member_model, const uint8_t *member_state);

/* ---- prefix index write side (approximateprefix/indexer.go:52-83, 105-115, 167-182; plugin.go:164-211) --
 * The LRU bookkeeping lives in HBM (index_store.cu): epp_index_add calls are queued and applied as ONE device batch,
 * in call order, by the next commit / lookup / remove; epp_index_add_picked never leaves the device.  The read table
 * is rebuilt on the device by epp_index_commit (also called implicitly by the next lookup when dirty). */
EPP_API int32_t epp_index_add(epp_engine *h, uint32_t ep, int32_t n, const uint64_t *hashes, int32_t num_gpu_blocks);
EPP_API int32_t epp_index_remove_endpoint(epp_engine *h, uint32_t ep);
/* CleanUpInactivePods (approximateprefix/plugin.go:99-122): RemovePod for every indexed endpoint NOT in active_ids. */
EPP_API int32_t epp_index_retain_endpoints(epp_engine *h, int32_t n, const uint32_t *active_ids);
/* Bulk-replace the index with a frozen snapshot of UNIQUE-or-not (hash, endpoint) pairs (host pointers);
 * bypasses LRU bookkeeping.  Duplicates are removed on the device. */
EPP_API int32_t epp_index_load_snapshot(epp_engine *h, uint64_t n_pairs, const uint64_t *hashes, const uint32_t *eps);
EPP_API int32_t epp_index_commit(epp_engine *h);
/* indexer.Get (indexer.go:86-102), read from the DEVICE table; returns the set size in *out_n. */
EPP_API int32_t epp_index_get(epp_engine *h, uint64_t hash, uint32_t *out_eps, int32_t cap, int32_t *out_n);

/* ---- a1: hashPrompt (approximateprefix/hashing.go:35-99) -----------------------------------------
 * out_hashes: [R][max_prefix_blocks] (row-major, unused tail undefined); out_nblocks: [R]. */
EPP_API int32_t epp_hash_prompts(epp_engine *h, const epp_batch *batch, uint64_t *out_hashes, int32_t *out_nblocks);

/* ---- a1-a4: Produce (approximateprefix/plugin.go:135-160) -- plugin-parity mode ---------------------
 * out_match: dense [R][max_endpoints] matchBlocks (0 for unmatched); out_total: [R] totalBlocks. */
EPP_API int32_t epp_prefix_match(epp_engine *h, const epp_batch *batch, int32_t *out_match, int32_t *out_total);

/* ---- a5-a9: Scorer.Score / runScorerPlugins (scheduler_profile.go:151-174) -- plugin-parity mode ----
 * match: [R][max_endpoints], total: [R] (as produced by epp_prefix_match or injected by a test).
 * profile: 0 = primary, 1 = prefill.  scorer_index >= 0: raw column of that scorer (what Scorer.Score
 * returns, 0 for non-candidates); -1: weighted, clamped, ordered sum (-1.0 for filtered-out endpoints).
 * out_scores: [R][max_endpoints].  flags: EPP_BATCH_DEVICE_PTRS.
 * model_ids: [R] model (= LoRA adapter) of each request, NULL = model 0 (only lora-affinity reads it). */
EPP_API int32_t epp_score(epp_engine *h, int64_t n_requests, const int32_t *match, const int32_t *total,
                  const uint32_t *model_ids, int32_t profile, int32_t scorer_index, double *out_scores, uint32_t flags);

/* ---- a1-a14 fused: Scheduler.Schedule for a batch (scheduling/scheduler.go:54-102) ---------------
 * out: [R]; detail: [R] or NULL; keep_hashes != 0 keeps the batch's prefix hashes resident on the device
 * for epp_index_add_picked (the PluginState stash of plugin.go:150-157). */
EPP_API int32_t epp_schedule(epp_engine *h, const epp_batch *batch, epp_decision *out, epp_decision_detail *detail,
                     int32_t keep_hashes);

/* Same decision logic with PrefixCacheMatchInfo injected by the caller (how the reference's own scheduler
 * tests drive it: disagg/scheduler_test.go:264-268): match [R][max_endpoints], total [R],
 * model_ids [R] or NULL, input_len_bytes [R] (prompt length for the P/D decider), block_size_tokens override
 * (0 = config). */
EPP_API int32_t epp_schedule_with_match(epp_engine *h, int64_t n_requests, const int32_t *match, const int32_t *total,
                                const uint32_t *model_ids, const int64_t *input_len_bytes, int32_t block_size_tokens, epp_decision *out,
                                epp_decision_detail *detail, uint32_t flags);

/* maxNumOfEndpoints > 1 (maxscore/picker.go:104-115): the k = epp_config.pick_k best endpoints of every profile that ran
 * for the request, in picker order (see "Top-k" above); rows are padded with EPP_NO_ENDPOINT / 0.0 when a profile has
 * fewer candidates than k or did not run.  primary[r*k] == out[r].pick, prefill[r*k] == out[r].prefill_pick,
 * encode[r*k] == detail[r].encode_pick.  Host pointers for host batches, device pointers with EPP_BATCH_DEVICE_PTRS;
 * any of the arrays may be NULL. */
typedef struct {
    uint32_t struct_size;          /* sizeof(epp_topk_out)                                              */
    int32_t k;                     /* row stride; must equal epp_config.pick_k                           */
    uint32_t *primary;             /* [R][k] slot ids                                                    */
    double *primary_scores;        /* [R][k] their weighted scores                                       */
    uint32_t *prefill;             /* [R][k]                                                             */
    uint32_t *encode;              /* [R][k]                                                             */
} epp_topk_out;
EPP_API int32_t epp_schedule_topk(epp_engine *h, const epp_batch *batch, epp_decision *out, epp_decision_detail *detail,
                          int32_t keep_hashes, const epp_topk_out *topk);
EPP_API int32_t epp_schedule_with_match_topk(epp_engine *h, int64_t n_requests, const int32_t *match, const int32_t *total,
                                     const uint32_t *model_ids, const int64_t *input_len_bytes, int32_t block_size_tokens,
                                     epp_decision *out, epp_decision_detail *detail, uint32_t flags,
                                     const epp_topk_out *topk);

/* PreRequest (approximateprefix/plugin.go:164-200): index the hashes of the LAST epp_schedule batch
 * (keep_hashes=1) for each request's primary pick and prefill pick. */
EPP_API int32_t epp_index_add_picked(epp_engine *h);

EPP_API int32_t epp_get_stats(epp_engine *h, epp_stats *out);

/* Waits for everything the engine has enqueued (EPP_BATCH_ASYNC batches). */
EPP_API int32_t epp_synchronize(epp_engine *h);
/* CUDA events on the engine's launch stream: record event `which` (0 = start, 1 = stop) after the work enqueued so
 * far; epp_event_elapsed_ms waits for the stop event and returns stop - start in milliseconds (device time). */
EPP_API int32_t epp_event_record(epp_engine *h, int32_t which);
EPP_API int32_t epp_event_elapsed_ms(epp_engine *h, double *out_ms);

/* ---- endpoint-sharded multi-GPU mode ----------------------------------------------
 * Each rank holds the postings of the endpoints of its shard and the full pool state.  Phase 1 probes
 * the local table and produces per-request block-presence masks (ceil(max_prefix_blocks/32) words per
 * request); the caller ORs the masks of all ranks (NCCL all-gather / bitwise OR); phase 2 applies the
 * global stop rule (plugin.go:219-223), counts local matches and returns the local best record; the caller
 * all-gathers the records and epp_shard_merge reduces them.  All pointers are DEVICE pointers. */
typedef struct {
    double score;
    uint32_t pick;
    uint32_t tie_count;
    int32_t match_blocks;
    int32_t status;            /* 0 = this shard has a candidate, -1 = none */
} epp_shard_best;              /* 24 bytes */
EPP_API int32_t epp_shard_set(epp_engine *h, uint32_t ep_begin, uint32_t ep_end);
EPP_API int32_t epp_shard_probe(epp_engine *h, const epp_batch *batch, uint32_t *out_masks);
EPP_API int32_t epp_shard_pick(epp_engine *h, int64_t n_requests, const uint32_t *global_masks, epp_shard_best *out_best);
EPP_API int32_t epp_shard_merge(epp_engine *h, int64_t n_requests, int32_t n_ranks, const epp_shard_best *all_best,
                        epp_decision *out);

/* The same protocol with both exchanges done by the ranks' own kernels over NVLink PEER MEMORY (no NCCL): every rank
 * exports one exchange buffer (CUDA IPC handle, 64 bytes; out_ptr is the raw device pointer for ranks that share a
 * process), connects to all peers' buffers (peers: n_ranks x 64-byte handles when ipc_handles != 0, else n_ranks
 * uint64 pointers), then calls epp_shard_schedule_p2p with the SAME device batch on every rank: presence masks are
 * OR-reduced and best records merged straight out of the peers' memory, ordered by release/acquire flags (the two halves
 * of a batch pipelined through the exchange on the engine's two streams); `out`
 * ([R] device) is identical on every rank.  A peer that does not show up within 20 s (EPP_P2P_TIMEOUT_MS) yields EPP_ERR_NCCL. */
EPP_API int32_t epp_shard_p2p_export(epp_engine *h, int64_t max_requests, uint8_t *out_handle, uint64_t *out_ptr);
EPP_API int32_t epp_shard_p2p_connect(epp_engine *h, int32_t n_ranks, int32_t rank, const void *peers, int32_t ipc_handles);
EPP_API int32_t epp_shard_schedule_p2p(epp_engine *h, const epp_batch *batch, epp_decision *out);
/* The same batch ONE phase at a time (0: hashes + local masks + flag 1; 1: OR of the peers' masks + local best records +
 * flag 2; 2: gather + merge -> out), each call synchronising before it returns: for ranks that share a GPU, whose
 * device-side waits could starve each other -- step all ranks through phase 0, then all through 1, then 2.
 * A peer that times out (or whose flags are poisoned) breaks the exchange for good: every later call fails with
 * EPP_ERR_STATE until epp_shard_p2p_export / _connect are called again on every rank. */
EPP_API int32_t epp_shard_p2p_phase(epp_engine *h, const epp_batch *batch, epp_decision *out, int32_t phase);

/* The configuration the engine was created with. */
EPP_API int32_t epp_get_config(epp_engine *h, epp_config *out);

/* ---- micro-batcher: one request in, one decision out ------------------------------------------
 * The reference calls Scheduler.Schedule(ctx, *InferenceRequest, []Endpoint) once per in-flight request, from one
 * goroutine each (requestcontrol/director.go:69-71, 243; handlers/server.go:168).  epp_submit / epp_wait give a shim
 * exactly that shape: any number of threads submit single prompts; a flusher thread inside the library closes a batch
 * when it holds max_batch requests or its oldest request has waited max_delay_us, evaluates it against ONE frozen
 * snapshot (epp_schedule), hands out the decisions and -- with index_picks -- applies PreRequest (epp_index_add_picked,
 * approximateprefix/plugin.go:164-200; off the request path like the reference's, :189-194) before the next batch;
 * epp_wait blocks until the ticket's batch is done.
 * A batcher must be destroyed before its engine; other threads may keep calling epp_pool_set etc. on the engine
 * (the scrape loop) -- every flush sees the snapshot current at its start. */
typedef struct epp_batcher epp_batcher;
typedef struct {
    uint32_t struct_size;
    int32_t max_batch;             /* requests per flush at most                                          */
    int32_t max_delay_us;          /* a batch closes when its oldest request has waited this long (0 = flush at once) */
    int32_t index_picks;           /* != 0: PreRequest after every flush (the picks' prefix hashes enter the index) */
} epp_batcher_cfg;
typedef struct {
    uint64_t n_flushes, n_requests, n_full_flushes, n_pending;
    uint64_t n_index_errors;       /* flushes whose PreRequest (epp_index_add_picked) failed: it runs after the decisions
                                      were handed out, so its error reaches epp_batcher_stats / the log, not the waiters */
} epp_batcher_stats_t;
EPP_API int32_t epp_batcher_create(epp_engine *h, const epp_batcher_cfg *cfg, epp_batcher **out);
EPP_API int32_t epp_batcher_destroy(epp_batcher *b);     /* flushes what is pending, then stops the flusher   */
/* prompt: the request's bytes (a uint32 token array = 4 bytes per token, hashing.go:49), copied before the call returns
 * (only the first max_prefix_blocks blocks are kept); multimodal: hasMultimodalContent (encode decider).  A ticket can
 * be waited for (any number of times, from any thread) until 4096 later batches have been flushed; after that
 * epp_wait returns EPP_ERR_STATE.  Every epp_wait must have returned before epp_batcher_destroy.
 * Neither call takes a lock: a submit is one compare-and-swap + the copy, a wait polls / sleeps on the flush counter. */
EPP_API int32_t epp_submit(epp_batcher *b, uint32_t model_id, const void *prompt, uint64_t prompt_len,
                           uint32_t multimodal, uint64_t *out_ticket);
EPP_API int32_t epp_wait(epp_batcher *b, uint64_t ticket, epp_decision *out, epp_decision_detail *out_detail);
/* The same on an engine created with pick_k > 1, plus the request's first-k lists (ProfileRunResult.TargetEndpoints of
 * each profile, see epp_topk_out): primary / prefill / encode receive pick_k slot ids each (EPP_NO_ENDPOINT padded); any
 * of them may be NULL. */
EPP_API int32_t epp_wait_topk(epp_batcher *b, uint64_t ticket, epp_decision *out, epp_decision_detail *out_detail,
                              uint32_t *primary, uint32_t *prefill, uint32_t *encode);
EPP_API int32_t epp_batcher_stats(epp_batcher *b, epp_batcher_stats_t *out);
EPP_API const char *epp_batcher_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
