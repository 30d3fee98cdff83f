// engine.cu -- the C ABI of libepp_engine.so (include/epp_engine.h) and the host logic above the kernels:
// configuration, pool-state snapshot, prefix-index store + device table, batch staging (H2D / D2H pipelined
// against the kernels on two streams) and the per-batch launch sequence.
//
// There is NO CPU compute path in this file: without a CUDA device epp_engine_create fails with
// EPP_ERR_NO_DEVICE, and every data-path entry point only stages buffers and launches kernels.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "devbuf.h"
#include "index_store.h"
#include "kernels.h"

using namespace epp;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int32_t fail(int32_t code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(EPP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define EPP_TRY(expr)            \
    do {                         \
        int32_t _r = (expr);     \
        if (_r != EPP_OK) return _r; \
    } while (0)

struct ProfileState {
    DevBuf cand, contrib, base, order, grp_size, sort_key, n_cand, qminmax;
};

struct Slot {                       // per-stream staging for host-pointer batches
    cudaStream_t stream = nullptr;
    DevBuf data;
    DevBuf overflow_list, overflow_n;   // requests handed from a sparse pass to the dense-counter pass (per stream)
    DevBuf pick_scratch;                // zeroed global match counters of the dense-counter kernel when max_endpoints is too
                                        // large for shared memory: per stream, two streams may run that kernel at once
    cudaEvent_t done = nullptr;
};

struct epp_engine {
    std::mutex mu;
    epp_config cfg;
    int n_profiles = 1;             // profiles of the P/D handler: 1 (single) or 2 (decode + prefill)
    int n_alloc_profiles = 1;       // + the encode profile ([2]) when configured
    int sm_count = 148;
    int32_t Epad = 0;
    size_t dev_bytes = 0;
    size_t max_smem_optin = 0;

    Slot slot[2];
    cudaEvent_t ev[8] = {};
    cudaEvent_t user_ev[2] = {};    // epp_event_record / epp_event_elapsed_ms
    unsigned long long *wc_host = nullptr;   // pinned: work counters of an async batch
    bool async_pending = false;     // an EPP_BATCH_ASYNC batch whose stats have not been read back yet
    bool s1_unjoined = false;       // chunk-pipelined async batches left work on stream 1 that stream 0 has not waited for
    int64_t pipe_R = 0, pipe_per = 0;   // chunk layout of the pipelined batches in flight (row ranges are per stream)
    int async_launches = 0;
    DevBuf work_counters;           // u64[2] probes, postings

    // models
    DevBuf seeds;                   // u64[kMaxModels]
    DevBuf seed_msg;
    int n_models = 0;
    std::vector<uint64_t> host_seeds;

    // pool state
    bool pool_ready = false;
    DevBuf role, kv, waiting, running, ext;
    DevBuf lora_max, lora_cnt, lora_ptr, lora_ep, lora_state, score_models;   // LoRA residency (epp_pool_set_lora)
    bool lora_set = false, lora_enabled = false;
    int32_t lora_models = 0;
    ProfileState prof[kMaxProfiles];

    // index
    std::unique_ptr<IndexStore> store;          // write side: LRU bookkeeping in HBM (index_store.cu)
    // indexer.Add calls queued by epp_index_add until the next commit / read (applied as ONE device batch)
    std::vector<uint32_t> q_ep, q_n;
    std::vector<int32_t> q_nb;
    std::vector<uint64_t> q_src, q_hashes;
    DevBuf q_ep_dev, q_n_dev, q_nb_dev, q_src_dev, q_hashes_dev, kept_dec;
    bool snapshot_mode = false;
    DevBuf slots, postings, idx_scratch, idx_cursor, idx_special, pair_hash, pair_ep, get_out, intern_keys, intern_vals;
    uint64_t idx_capacity = 0, idx_pairs = 0;
    // incremental maintenance of the read table (IndexStore::patch_read_table)
    bool rt_from_store = false;     // the table mirrors the store's inverted map (bulk-built from its export)
    uint64_t post_cap = 0, post_used = 0, rt_slots_used = 0;
    int patches_since_build = 0;
    int patch_enabled = 1;          // EPP_INDEX_PATCH=0: always rebuild
    IndexSlot idx_special_host{};
    uint32_t shard_begin = 0, shard_end = 0xFFFFFFFFu;

    // batch buffers (sized for the largest batch seen)
    DevBuf offsets, lengths, model_ids, multimodal, hashes, nblocks, eff_len, in_len, decisions, details, flag;
    DevBuf dense_match, dense_total, dense_scores;
    int index_load = 0;             // EPP_INDEX_LOAD=n: read-table slots per distinct hash (2 = load factor <= 0.5, 4 = <= 0.25);
                                    // 0 = automatic: 4 while the table stays L2-sized (<= 48 MiB), else 2
    int dev_chunks = 2;             // EPP_DEV_CHUNKS=N: async device batches run as N chunks over both streams (1 = off)
    int staged = 0;                 // EPP_HASH_STAGED: 0 = hash_fused.cu (default), 1 = hash_staged.cu default shape, else a shape
    int pick_grid = 0;
    bool pick_global = false;
    size_t pick_smem = 0;
    // small host batches: one zero-copy kernel, no copy engine, no stream synchronisation (cycle_small.cu)
    int small_max = 1024;           // EPP_SMALL_BATCH=n: largest batch that takes the path (0 = off)
    int small_zc_max = 8;           // EPP_SMALL_ZEROCOPY=n: largest batch whose prompts the kernel reads over PCIe itself
    int64_t small_cap = 0;          // requests the pinned staging below is sized for
    uint8_t *small_host = nullptr;  // one pinned + mapped allocation: 2 x {offsets, lengths, model_ids, multimodal}, dec, det, flags
    uint64_t *small_offsets[2] = {}, *small_lengths[2] = {};
    uint32_t *small_models[2] = {};
    uint8_t *small_mm[2] = {};
    epp_decision *small_dec = nullptr;
    epp_decision_detail *small_det = nullptr;
    uint32_t *small_flags = nullptr;
    uint32_t small_epoch = 0;
    DevBuf small_overflow_n;        // stays zero between batches (reset by the overflow pass)
    int small_pipe_min = 1;         // EPP_SMALL_PIPELINE=n: DMA-copied batches of >= n requests start their kernel BEFORE the copy
                                    // lands (copy on the second stream, one stream-ordered flag write behind it); 0 = off
    DevBuf small_arrive;            // epoch word: the prompts of the batch have arrived
    uint32_t *small_epoch_host = nullptr;   // pinned source of that word when the driver has no stream write-value
    bool small_no_write_value = false;
    bool small_stats_pending = false;   // ev[0] / ev[1] bracket the last small batch; read lazily by epp_get_stats
    std::chrono::steady_clock::time_point last_commit{};   // last commit done by a scheduling call (index_commit_interval_us)
    bool committed_once = false;
    bool general = false;           // a profile configures the prefix-cache-affinity-filter or pick_k > 1: every batch takes
                                    // the dense-counter kernel with the full-scan evaluation of pick_general.cuh
    const epp_topk_out *cur_topk = nullptr;   // top-k destination of the call in progress (under mu)
    DevBuf topk_picks[kMaxProfiles], topk_scores;   // device staging of the lists for host batches
    int64_t kept_R = 0;             // rows of `hashes` valid for epp_index_add_picked
    int64_t shard_R = 0;            // rows of `hashes` valid for epp_shard_pick / epp_shard_merge
    // endpoint-sharded mode over NVLink peer memory (shard_p2p.cu)
    void *p2p_buf = nullptr;        // this rank's exchange buffer {flags, masks, records} (cudaMalloc, IPC-exported)
    size_t p2p_bytes = 0, p2p_masks_off = 0, p2p_best_off = 0;
    int64_t p2p_R = 0;              // requests the buffer was sized for
    int p2p_ranks = 0, p2p_rank = 0;
    bool p2p_ipc = false;
    std::vector<void *> p2p_peer;   // peers' buffers in this process's address space (own buffer at [p2p_rank])
    DevBuf p2p_peer_dev, p2p_gmasks, p2p_allbest, p2p_err;
    unsigned long long p2p_epoch = 0;
    uint64_t p2p_or_bits = 0;       // offsets alignment of the batch being stepped through epp_shard_p2p_phase
    bool p2p_broken = false;        // a wait timed out: the ranks are out of step until the buffers are connected again
    int p2p_launches = 0;

    epp_stats stats{};
};

static constexpr int kMaxModels = 4096;

static int32_t set_device(epp_engine *e) {
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    return EPP_OK;
}

// Orders stream 0 after whatever chunk-pipelined async batches left on stream 1 (no host synchronisation).  Every entry
// point that enqueues on stream 0 calls it first, so everything outside the pipelined path stays in call order.
static int32_t join_streams(epp_engine *e) {
    if (!e->s1_unjoined) return EPP_OK;
    CUDA_TRY(cudaEventRecord(e->slot[1].done, e->slot[1].stream));
    CUDA_TRY(cudaStreamWaitEvent(e->slot[0].stream, e->slot[1].done, 0));
    e->s1_unjoined = false;
    return EPP_OK;
}

// ------------------------------------------------------------------------------------------------
// config
// ------------------------------------------------------------------------------------------------
extern "C" int32_t epp_abi_version(void) { return EPP_ABI_VERSION; }
extern "C" const char *epp_last_error(void) { return g_last_error.c_str(); }

extern "C" void epp_config_default(epp_config *cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    cfg->struct_size = sizeof(epp_config);
    cfg->device = 0;
    cfg->max_endpoints = 4096;
    cfg->block_size_tokens = 16;            // types.go:92
    cfg->max_prefix_blocks = 256;           // types.go:99
    cfg->lru_capacity_per_server = 31250;   // types.go:110
    cfg->handler = EPP_HANDLER_SINGLE;
    // config/loader/defaults.go:47-49, 78-87: queue 2, kv-cache-utilization 2, prefix 3 -- in that order
    cfg->primary.filter = EPP_FILTER_NONE;
    cfg->primary.n_scorers = 3;
    cfg->primary.scorers[0] = {EPP_SCORER_QUEUE, 0, 2.0, 0.0, 0.0};
    cfg->primary.scorers[1] = {EPP_SCORER_KV_UTIL, 0, 2.0, 0.0, 0.0};
    cfg->primary.scorers[2] = {EPP_SCORER_PREFIX, 0, 3.0, 0.0, 0.0};
    cfg->primary.ttft_column = cfg->prefill.ttft_column = cfg->encode.ttft_column = -1;   // no affinity filter, no TTFT column
}

static int32_t validate_profile(const epp_profile_cfg &p, int n_ext, const char *name) {
    if (p.filter < EPP_FILTER_NONE || p.filter > EPP_FILTER_ENCODE) return fail(EPP_ERR_INVALID, "%s: bad filter %d", name, p.filter);
    if (p.n_scorers < 0 || p.n_scorers > EPP_MAX_SCORERS) return fail(EPP_ERR_INVALID, "%s: n_scorers %d out of range", name, p.n_scorers);
    for (int s = 0; s < p.n_scorers; s++) {
        const epp_scorer_cfg &sc = p.scorers[s];
        if (sc.kind < EPP_SCORER_PREFIX || sc.kind > EPP_SCORER_LORA_AFFINITY) return fail(EPP_ERR_INVALID, "%s: scorer %d has unknown kind %d", name, s, sc.kind);
        if (!std::isfinite(sc.weight)) return fail(EPP_ERR_INVALID, "%s: scorer %d weight is not finite", name, s);
        if ((sc.kind == EPP_SCORER_TOKEN_LOAD || sc.kind == EPP_SCORER_ACTIVE_REQUEST) && (sc.column < 0 || sc.column >= n_ext))
            return fail(EPP_ERR_INVALID, "%s: scorer %d reads ext column %d, out of range [0,%d)", name, s, sc.column, n_ext);
        if (sc.kind == EPP_SCORER_EXTERNAL && (sc.param < 0 || sc.param >= n_ext)) return fail(EPP_ERR_INVALID, "%s: scorer %d external column %g out of range [0,%d)", name, s, sc.param, n_ext);
    }
    // prefix-cache-affinity-filter: Config.validate, filter/prefixcacheaffinity/plugin.go:87-98
    if (std::isnan(p.affinity_threshold) || p.affinity_threshold > 1.0) return fail(EPP_ERR_INVALID, "%s: affinityThreshold must be <= 1.0, got %f", name, p.affinity_threshold);
    if (!(p.exploration_probability >= 0.0 && p.exploration_probability <= 1.0)) return fail(EPP_ERR_INVALID, "%s: explorationProbability must be in [0, 1], got %f", name, p.exploration_probability);
    if (!(p.max_ttft_penalty_ms >= 0.0)) return fail(EPP_ERR_INVALID, "%s: maxTTFTPenaltyMs must be >= 0, got %f", name, p.max_ttft_penalty_ms);
    if (p.affinity_threshold > 0.0 && p.ttft_column >= n_ext) return fail(EPP_ERR_INVALID, "%s: ttft_column %d out of range [0,%d)", name, p.ttft_column, n_ext);
    return EPP_OK;
}

static int32_t alloc_profile(epp_engine *e, ProfileState &ps) {
    size_t E = (size_t)e->cfg.max_endpoints, Ep = (size_t)e->Epad;
    CUDA_TRY(ps.cand.reserve(E, &e->dev_bytes));
    CUDA_TRY(ps.contrib.reserve(sizeof(double) * EPP_MAX_SCORERS * E, &e->dev_bytes));
    CUDA_TRY(ps.base.reserve(sizeof(double) * E, &e->dev_bytes));
    CUDA_TRY(ps.order.reserve(sizeof(uint32_t) * Ep, &e->dev_bytes));
    CUDA_TRY(ps.grp_size.reserve(sizeof(uint32_t) * Ep, &e->dev_bytes));
    CUDA_TRY(ps.sort_key.reserve(sizeof(uint64_t) * Ep, &e->dev_bytes));
    CUDA_TRY(ps.n_cand.reserve(sizeof(int32_t) * 4, &e->dev_bytes));
    CUDA_TRY(ps.qminmax.reserve(sizeof(int64_t) * (4 + EPP_MAX_SCORERS), &e->dev_bytes));
    return EPP_OK;
}

extern "C" int32_t epp_engine_create(const epp_config *cfg, epp_engine **out) {
    if (!cfg || !out) return fail(EPP_ERR_INVALID, "cfg/out is NULL");
    *out = nullptr;
    if (cfg->struct_size != sizeof(epp_config)) return fail(EPP_ERR_INVALID, "epp_config.struct_size %u != %zu (ABI mismatch)", cfg->struct_size, sizeof(epp_config));
    if (cfg->max_endpoints <= 0 || cfg->max_endpoints > (1 << 24)) return fail(EPP_ERR_INVALID, "max_endpoints %d out of range", cfg->max_endpoints);
    if (cfg->block_size_tokens <= 0 || cfg->block_size_tokens > (1 << 20)) return fail(EPP_ERR_INVALID, "block_size_tokens %d must be > 0 (plugin.go:75-77)", cfg->block_size_tokens);
    if (cfg->max_prefix_blocks <= 0 || cfg->max_prefix_blocks > 65535) return fail(EPP_ERR_INVALID, "max_prefix_blocks %d out of range [1,65535]", cfg->max_prefix_blocks);
    if (cfg->lru_capacity_per_server <= 0) return fail(EPP_ERR_INVALID, "lru_capacity_per_server must be > 0");
    if (cfg->handler != EPP_HANDLER_SINGLE && cfg->handler != EPP_HANDLER_DISAGG) return fail(EPP_ERR_INVALID, "bad handler %d", cfg->handler);
    if (cfg->n_ext_cols < 0 || cfg->n_ext_cols > 64) return fail(EPP_ERR_INVALID, "n_ext_cols %d out of range", cfg->n_ext_cols);
    if (cfg->non_cached_tokens < 0) return fail(EPP_ERR_INVALID, "non_cached_tokens must be >= 0");
    EPP_TRY(validate_profile(cfg->primary, cfg->n_ext_cols, "primary profile"));
    if (cfg->handler == EPP_HANDLER_DISAGG) EPP_TRY(validate_profile(cfg->prefill, cfg->n_ext_cols, "prefill profile"));
    if (cfg->encode_enabled) {
        if (cfg->handler != EPP_HANDLER_DISAGG) return fail(EPP_ERR_INVALID, "encode_enabled needs the disagg handler");
        EPP_TRY(validate_profile(cfg->encode, cfg->n_ext_cols, "encode profile"));
    }
    if (cfg->index_commit_interval_us < 0) return fail(EPP_ERR_INVALID, "index_commit_interval_us must be >= 0");
    if (cfg->pick_k < 0 || cfg->pick_k > 64) return fail(EPP_ERR_INVALID, "pick_k %d out of range [0,64]", cfg->pick_k);

    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(EPP_ERR_NO_DEVICE, "no CUDA device available (%s); the engine has no CPU fallback", ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(EPP_ERR_INVALID, "device %d out of range [0,%d)", cfg->device, ndev);

    std::unique_ptr<epp_engine> e(new epp_engine());
    e->cfg = *cfg;
    e->n_profiles = cfg->handler == EPP_HANDLER_DISAGG ? 2 : 1;
    e->n_alloc_profiles = cfg->encode_enabled ? 3 : e->n_profiles;
    CUDA_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    e->sm_count = prop.multiProcessorCount;
    e->max_smem_optin = prop.sharedMemPerBlockOptin;
    int32_t Ep = 1;
    while (Ep < cfg->max_endpoints) Ep <<= 1;
    e->Epad = Ep;
    for (int i = 0; i < 2; i++) {
        CUDA_TRY(cudaStreamCreateWithFlags(&e->slot[i].stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&e->slot[i].done, cudaEventDisableTiming));
    }
    for (auto &ev : e->ev) CUDA_TRY(cudaEventCreate(&ev));
    for (auto &ev : e->user_ev) CUDA_TRY(cudaEventCreate(&ev));
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&e->wc_host), sizeof(unsigned long long) * 2, cudaHostAllocDefault));
    size_t E = (size_t)cfg->max_endpoints;
    CUDA_TRY(e->seeds.reserve(sizeof(uint64_t) * kMaxModels, &e->dev_bytes));
    CUDA_TRY(e->seed_msg.reserve(4096, &e->dev_bytes));
    CUDA_TRY(e->role.reserve(E, &e->dev_bytes));
    CUDA_TRY(e->kv.reserve(sizeof(double) * E, &e->dev_bytes));
    CUDA_TRY(e->waiting.reserve(sizeof(int32_t) * E, &e->dev_bytes));
    CUDA_TRY(e->running.reserve(sizeof(int32_t) * E, &e->dev_bytes));
    if (cfg->n_ext_cols) CUDA_TRY(e->ext.reserve(sizeof(double) * E * (size_t)cfg->n_ext_cols, &e->dev_bytes));
    for (int p = 0; p < e->n_alloc_profiles; p++) EPP_TRY(alloc_profile(e.get(), e->prof[p]));
    CUDA_TRY(e->idx_cursor.reserve(sizeof(uint32_t) * 4, &e->dev_bytes));
    CUDA_TRY(e->idx_special.reserve(sizeof(IndexSlot), &e->dev_bytes));
    CUDA_TRY(e->get_out.reserve(sizeof(uint32_t) * 4096 + 16, &e->dev_bytes));
    CUDA_TRY(e->flag.reserve(sizeof(int) * 4, &e->dev_bytes));
    CUDA_TRY(e->work_counters.reserve(sizeof(unsigned long long) * 2, &e->dev_bytes));
    memset(&e->idx_special_host, 0, sizeof(IndexSlot));
    e->idx_special_host.key = kEmptyKey;
    for (int pi = 0; pi < e->n_alloc_profiles; pi++) {
        const epp_profile_cfg &pc = pi == 0 ? cfg->primary : (pi == 1 ? cfg->prefill : cfg->encode);
        for (int si = 0; si < pc.n_scorers; si++) if (pc.scorers[si].kind == EPP_SCORER_LORA_AFFINITY) e->lora_enabled = true;
    }
    e->general = cfg->pick_k > 1;
    for (int pi = 0; pi < e->n_alloc_profiles; pi++) {
        const epp_profile_cfg &pc = pi == 0 ? cfg->primary : (pi == 1 ? cfg->prefill : cfg->encode);
        if (pc.affinity_threshold > 0.0) e->general = true;
    }
    e->store.reset(new IndexStore((uint32_t)cfg->max_endpoints, cfg->lru_capacity_per_server));
    { const char *v1 = getenv("EPP_INDEX_PATCH"); e->patch_enabled = v1 ? atoi(v1) : 1; }
    { const char *v1 = getenv("EPP_INDEX_LOAD"); e->index_load = v1 ? std::max(2, atoi(v1)) : 0; }
    { const char *v1 = getenv("EPP_DEV_CHUNKS"); e->dev_chunks = v1 ? std::max(1, atoi(v1)) : 2; }
    { const char *v1 = getenv("EPP_HASH_STAGED"); e->staged = v1 ? atoi(v1) : 0; }
    { const char *v1 = getenv("EPP_SMALL_BATCH"); e->small_max = v1 ? std::max(0, atoi(v1)) : 1024; }
    { const char *v1 = getenv("EPP_SMALL_ZEROCOPY"); e->small_zc_max = v1 ? std::max(0, atoi(v1)) : 8; }
    { const char *v1 = getenv("EPP_SMALL_PIPELINE"); e->small_pipe_min = v1 ? std::max(0, atoi(v1)) : 1; }
    CUDA_TRY(e->small_arrive.reserve(sizeof(uint32_t) * 16, &e->dev_bytes));
    CUDA_TRY(cudaMemset(e->small_arrive.p, 0, sizeof(uint32_t) * 16));
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&e->small_epoch_host), 64, cudaHostAllocDefault));
    CUDA_TRY(e->small_overflow_n.reserve(sizeof(int32_t) * 4, &e->dev_bytes));
    CUDA_TRY(cudaMemset(e->small_overflow_n.p, 0, sizeof(int32_t) * 4));
    for (int i = 0; i < 2; i++) CUDA_TRY(e->slot[i].overflow_n.reserve(sizeof(int32_t) * 4, &e->dev_bytes));

    // match/pick launch geometry: counters in shared memory when they fit, else zeroed global scratch
    size_t smem_local = match_pick_smem_bytes(cfg->max_endpoints, false);
    e->pick_global = smem_local > std::min<size_t>(e->max_smem_optin, 200 * 1024);
    e->pick_smem = match_pick_smem_bytes(cfg->max_endpoints, e->pick_global);
    int ctas_per_sm = e->pick_global ? 4 : std::max<int>(1, (int)((size_t)(220 * 1024) / (e->pick_smem + 1024)));
    ctas_per_sm = std::min(ctas_per_sm, 8);
    e->pick_grid = e->sm_count * ctas_per_sm;
    if (e->pick_global) {
        size_t words = (size_t)e->pick_grid * match_pick_warps_per_cta() * (size_t)((cfg->max_endpoints + 1) / 2);
        for (int i = 0; i < 2; i++) {
            CUDA_TRY(e->slot[i].pick_scratch.reserve(words * sizeof(uint32_t), &e->dev_bytes));
            CUDA_TRY(cudaMemset(e->slot[i].pick_scratch.p, 0, words * sizeof(uint32_t)));
        }
    }
    CUDA_TRY(cudaDeviceSynchronize());
    *out = e.release();
    return EPP_OK;
}

extern "C" int32_t epp_engine_destroy(epp_engine *h) {
    if (!h) return EPP_OK;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 2; i++) {
        if (h->slot[i].stream) cudaStreamDestroy(h->slot[i].stream);
        if (h->slot[i].done) cudaEventDestroy(h->slot[i].done);
    }
    for (auto &ev : h->ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : h->user_ev) if (ev) cudaEventDestroy(ev);
    if (h->wc_host) cudaFreeHost(h->wc_host);
    if (h->small_host) cudaFreeHost(h->small_host);
    if (h->small_epoch_host) cudaFreeHost(h->small_epoch_host);
    for (int g = 0; g < (int)h->p2p_peer.size(); g++)
        if (h->p2p_ipc && g != h->p2p_rank && h->p2p_peer[g]) cudaIpcCloseMemHandle(h->p2p_peer[g]);
    if (h->p2p_buf) cudaFree(h->p2p_buf);
    delete h;
    return EPP_OK;
}

extern "C" int32_t epp_host_alloc(size_t bytes, void **out) {
    if (!out) return fail(EPP_ERR_INVALID, "out is NULL");
    CUDA_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return EPP_OK;
}
extern "C" int32_t epp_host_free(void *p) {
    if (p) CUDA_TRY(cudaFreeHost(p));
    return EPP_OK;
}

// ------------------------------------------------------------------------------------------------
// models: h_{-1} = XXH64(model || salt) on the device (hashing.go:71-78)
// ------------------------------------------------------------------------------------------------
extern "C" int32_t epp_model_register(epp_engine *h, const uint8_t *model, size_t model_len, const uint8_t *salt,
                                      size_t salt_len, uint32_t *out_model_id) {
    if (!h || !out_model_id) return fail(EPP_ERR_INVALID, "NULL argument");
    if ((model_len && !model) || (salt_len && !salt)) return fail(EPP_ERR_INVALID, "NULL model/salt with non-zero length");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    if (h->n_models >= kMaxModels) return fail(EPP_ERR_CAPACITY, "too many registered models (max %d)", kMaxModels);
    size_t n = model_len + salt_len;
    std::vector<uint8_t> msg(n ? n : 1);
    if (model_len) memcpy(msg.data(), model, model_len);
    if (salt_len) memcpy(msg.data() + model_len, salt, salt_len);
    CUDA_TRY(h->seed_msg.reserve(n + 16, &h->dev_bytes));
    cudaStream_t s = h->slot[0].stream;
    if (n) CUDA_TRY(cudaMemcpyAsync(h->seed_msg.p, msg.data(), n, cudaMemcpyHostToDevice, s));
    CUDA_TRY(launch_hash_bytes(h->seed_msg.as<uint8_t>(), n, h->seeds.as<uint64_t>() + h->n_models, s));
    uint64_t seed = 0;
    CUDA_TRY(cudaMemcpyAsync(&seed, h->seeds.as<uint64_t>() + h->n_models, sizeof seed, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    h->host_seeds.push_back(seed);
    *out_model_id = (uint32_t)h->n_models++;
    return EPP_OK;
}

extern "C" int32_t epp_model_seed(epp_engine *h, uint32_t model_id, uint64_t *out_seed) {
    if (!h || !out_seed) return fail(EPP_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(h->mu);
    if ((int)model_id >= h->n_models) return fail(EPP_ERR_INVALID, "unknown model id %u", model_id);
    *out_seed = h->host_seeds[model_id];
    return EPP_OK;
}

// ------------------------------------------------------------------------------------------------
// pool state
// ------------------------------------------------------------------------------------------------
static LoraDev lora_dev(epp_engine *h) {
    LoraDev L{};
    L.enabled = h->lora_enabled ? 1 : 0;
    if (h->lora_set) {
        L.max_active = h->lora_max.as<int32_t>();
        L.n_loaded = h->lora_cnt.as<int32_t>();
        L.ptr = h->lora_ptr.as<uint32_t>();
        L.ep = h->lora_ep.as<uint32_t>();
        L.state = h->lora_state.as<uint8_t>();
        L.n_models = h->lora_models;
    }
    return L;
}

static PoolArrays pool_arrays(epp_engine *h) {
    PoolArrays pa;
    pa.E = h->cfg.max_endpoints;
    pa.n_ext_cols = h->cfg.n_ext_cols;
    pa.role = h->role.as<uint8_t>();
    pa.kv_usage = h->kv.as<double>();
    pa.waiting = h->waiting.as<int32_t>();
    pa.running = h->running.as<int32_t>();
    pa.ext = h->ext.as<double>();
    pa.lora = lora_dev(h);
    return pa;
}

static ProfileDev profile_dev(epp_engine *h, int p) {
    ProfileDev pd;
    pd.cfg = p == 0 ? h->cfg.primary : (p == 1 ? h->cfg.prefill : h->cfg.encode);
    pd.cand = h->prof[p].cand.as<uint8_t>();
    pd.contrib = h->prof[p].contrib.as<double>();
    pd.base = h->prof[p].base.as<double>();
    pd.order = h->prof[p].order.as<uint32_t>();
    pd.grp_size = h->prof[p].grp_size.as<uint32_t>();
    pd.n_cand = h->prof[p].n_cand.as<int32_t>();
    return pd;
}

// Request-independent scorer terms of every profile from the current pool (+ LoRA) snapshot, on the device.
static int32_t derive_pool(epp_engine *h) {
    cudaStream_t s = h->slot[0].stream;
    PoolArrays pa = pool_arrays(h);
    int launches = 0;
    for (int p = 0; p < h->n_alloc_profiles; p++) {
        ProfileDerived d;
        d.cand = h->prof[p].cand.as<uint8_t>();
        d.contrib = h->prof[p].contrib.as<double>();
        d.base = h->prof[p].base.as<double>();
        d.order = h->prof[p].order.as<uint32_t>();
        d.grp_size = h->prof[p].grp_size.as<uint32_t>();
        d.sort_key = h->prof[p].sort_key.as<uint64_t>();
        d.n_cand = h->prof[p].n_cand.as<int32_t>();
        d.qminmax = h->prof[p].qminmax.as<int64_t>();
        CUDA_TRY(launch_pool_prepare(pa, p == 0 ? h->cfg.primary : (p == 1 ? h->cfg.prefill : h->cfg.encode), d, h->Epad, h->shard_begin, h->shard_end, s, &launches));
    }
    return EPP_OK;
}

extern "C" int32_t epp_pool_set(epp_engine *h, int32_t n, const uint32_t *ids, const uint8_t *role,
                                const double *kv_usage, const int32_t *waiting, const int32_t *running,
                                const double *ext) {
    if (!h) return fail(EPP_ERR_INVALID, "NULL engine");
    if (n < 0 || (n > 0 && (!ids || !role || !kv_usage || !waiting))) return fail(EPP_ERR_INVALID, "NULL pool arrays");
    if (h->cfg.n_ext_cols > 0 && n > 0 && !ext) return fail(EPP_ERR_INVALID, "ext columns configured but ext is NULL");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    const int32_t E = h->cfg.max_endpoints;
    std::vector<uint8_t> r(E, 0xFF);
    std::vector<double> kv(E, 0.0);
    std::vector<int32_t> w(E, 0), run(E, 0);
    std::vector<double> ex((size_t)E * (size_t)std::max(1, h->cfg.n_ext_cols), 0.0);
    for (int32_t i = 0; i < n; i++) {
        uint32_t id = ids[i];
        if (id >= (uint32_t)E) return fail(EPP_ERR_INVALID, "pool entry %d: slot id %u >= max_endpoints %d", i, id, E);
        if (r[id] != 0xFF) return fail(EPP_ERR_INVALID, "pool entry %d: duplicate slot id %u", i, id);
        if (role[i] > EPP_ROLE_OTHER) return fail(EPP_ERR_INVALID, "pool entry %d: bad role %u", i, role[i]);
        if (std::isnan(kv_usage[i])) return fail(EPP_ERR_INVALID, "pool entry %d: kv_usage is NaN", i);
        r[id] = role[i];
        kv[id] = kv_usage[i];
        w[id] = waiting[i];
        run[id] = running ? running[i] : 0;
        for (int c = 0; c < h->cfg.n_ext_cols; c++) {
            double v = ext[(size_t)c * (size_t)n + (size_t)i];
            if (std::isnan(v)) return fail(EPP_ERR_INVALID, "pool entry %d: ext column %d is NaN", i, c);
            ex[(size_t)c * (size_t)E + id] = v;
        }
    }
    cudaStream_t s = h->slot[0].stream;
    CUDA_TRY(cudaMemcpyAsync(h->role.p, r.data(), E, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->kv.p, kv.data(), sizeof(double) * E, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->waiting.p, w.data(), sizeof(int32_t) * E, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->running.p, run.data(), sizeof(int32_t) * E, cudaMemcpyHostToDevice, s));
    if (h->cfg.n_ext_cols)
        CUDA_TRY(cudaMemcpyAsync(h->ext.p, ex.data(), sizeof(double) * E * (size_t)h->cfg.n_ext_cols, cudaMemcpyHostToDevice, s));
    EPP_TRY(derive_pool(h));
    CUDA_TRY(cudaStreamSynchronize(s));
    h->pool_ready = true;
    return EPP_OK;
}

extern "C" int32_t epp_pool_set_lora(epp_engine *h, int32_t n, const uint32_t *ids, const int32_t *max_active_models,
                                     const int32_t *n_models_loaded, int64_t n_members, const uint32_t *member_ep,
                                     const uint32_t *member_model, const uint8_t *member_state) {
    if (!h || n < 0 || n_members < 0) return fail(EPP_ERR_INVALID, "bad arguments");
    if (n > 0 && (!ids || !max_active_models || !n_models_loaded)) return fail(EPP_ERR_INVALID, "NULL LoRA arrays");
    if (n_members > 0 && (!member_ep || !member_model || !member_state)) return fail(EPP_ERR_INVALID, "NULL LoRA member arrays");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    const uint32_t E = (uint32_t)h->cfg.max_endpoints;
    std::vector<int32_t> mx(E, 0), cnt(E, 0);
    for (int32_t i = 0; i < n; i++) {
        if (ids[i] >= E) return fail(EPP_ERR_INVALID, "LoRA entry %d: slot id %u >= max_endpoints %u", i, ids[i], E);
        mx[ids[i]] = max_active_models[i];
        cnt[ids[i]] = n_models_loaded[i];
    }
    const int32_t n_models = std::max(1, h->n_models);
    struct Mem { uint32_t model, ep; uint8_t st; };
    std::vector<Mem> mem;
    mem.reserve((size_t)n_members);
    for (int64_t i = 0; i < n_members; i++) {
        if (member_ep[i] >= E) return fail(EPP_ERR_INVALID, "LoRA member %lld: slot id %u >= max_endpoints %u", (long long)i, member_ep[i], E);
        if (member_model[i] >= (uint32_t)n_models) return fail(EPP_ERR_INVALID, "LoRA member %lld: model id %u is not registered", (long long)i, member_model[i]);
        if (member_state[i] != 1 && member_state[i] != 2) return fail(EPP_ERR_INVALID, "LoRA member %lld: state %u (1 = active, 2 = waiting)", (long long)i, member_state[i]);
        mem.push_back({member_model[i], member_ep[i], member_state[i]});
    }
    // CSR by adapter, endpoints ascending; an adapter both active and waiting on an endpoint counts as active
    // (the `case active` branch comes first, lora_affinity.go:85-87)
    std::sort(mem.begin(), mem.end(), [](const Mem &a, const Mem &b) {
        return a.model != b.model ? a.model < b.model : (a.ep != b.ep ? a.ep < b.ep : a.st < b.st);
    });
    std::vector<uint32_t> ptr((size_t)n_models + 1, 0), eps;
    std::vector<uint8_t> sts;
    for (size_t i = 0; i < mem.size(); i++) {
        if (i && mem[i].model == mem[i - 1].model && mem[i].ep == mem[i - 1].ep) continue;
        eps.push_back(mem[i].ep);
        sts.push_back(mem[i].st);
        ptr[mem[i].model + 1]++;
    }
    for (int32_t a = 0; a < n_models; a++) ptr[a + 1] += ptr[a];
    cudaStream_t s = h->slot[0].stream;
    CUDA_TRY(h->lora_max.reserve(sizeof(int32_t) * E, &h->dev_bytes));
    CUDA_TRY(h->lora_cnt.reserve(sizeof(int32_t) * E, &h->dev_bytes));
    CUDA_TRY(h->lora_ptr.reserve(sizeof(uint32_t) * ptr.size(), &h->dev_bytes));
    CUDA_TRY(h->lora_ep.reserve(sizeof(uint32_t) * std::max<size_t>(1, eps.size()), &h->dev_bytes));
    CUDA_TRY(h->lora_state.reserve(std::max<size_t>(1, sts.size()), &h->dev_bytes));
    CUDA_TRY(cudaMemcpyAsync(h->lora_max.p, mx.data(), sizeof(int32_t) * E, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->lora_cnt.p, cnt.data(), sizeof(int32_t) * E, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->lora_ptr.p, ptr.data(), sizeof(uint32_t) * ptr.size(), cudaMemcpyHostToDevice, s));
    if (!eps.empty()) {
        CUDA_TRY(cudaMemcpyAsync(h->lora_ep.p, eps.data(), sizeof(uint32_t) * eps.size(), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(h->lora_state.p, sts.data(), sts.size(), cudaMemcpyHostToDevice, s));
    }
    h->lora_models = n_models;
    h->lora_set = true;
    if (h->pool_ready) EPP_TRY(derive_pool(h));
    CUDA_TRY(cudaStreamSynchronize(s));
    return EPP_OK;
}

// ------------------------------------------------------------------------------------------------
// prefix index
// ------------------------------------------------------------------------------------------------
// Bulk build of the read table from n (hash, endpoint) pairs already in h->pair_hash / h->pair_ep.
static int32_t build_index_from_device_pairs(epp_engine *h, uint64_t n, uint64_t n_distinct_hint, bool from_store = false) {
    cudaStream_t s = h->slot[0].stream;
    if (n >= 0xFFFFFFF0ull) return fail(EPP_ERR_CAPACITY, "index snapshot of %llu pairs exceeds the u32 posting space", (unsigned long long)n);
    // posting space: the lists of this build + room for the copy-on-write lists of later patches
    uint64_t post_entries = std::max<uint64_t>(n, 1);
    if (from_store) post_entries = std::min<uint64_t>(0xFFFFFFF0ull, n + std::max<uint64_t>(n / 2, 1ull << 20));
    CUDA_TRY(h->postings.reserve(sizeof(uint32_t) * post_entries, &h->dev_bytes));
    uint64_t distinct = std::max<uint64_t>(1, n_distinct_hint);
    uint32_t cursor[4] = {0, 0, 0, 0};
    uint64_t cap = 16;
    for (int pass = 0; pass < 2; pass++) {
        // load factor <= 0.5 over the distinct hashes (<= 0.25 while the table stays L2-sized: shorter probe chains, measured
        // 10 % faster match kernel); the first pass only knows an upper bound (the pair count)
        const uint64_t load = h->index_load ? (uint64_t)h->index_load : (4 * distinct * sizeof(IndexSlot) <= (48ull << 20) ? 4 : 2);
        uint64_t want = std::max<uint64_t>(16, load * distinct);
        cap = 16;
        while (cap < want) cap <<= 1;
        CUDA_TRY(h->slots.reserve(sizeof(IndexSlot) * cap, &h->dev_bytes));
        CUDA_TRY(h->idx_scratch.reserve(sizeof(uint32_t) * cap, &h->dev_bytes));
        CUDA_TRY(h->intern_keys.reserve(sizeof(uint64_t) * cap, &h->dev_bytes));
        CUDA_TRY(h->intern_vals.reserve(sizeof(uint32_t) * cap, &h->dev_bytes));
        int launches = 0;
        CUDA_TRY(launch_index_build(h->pair_hash.as<uint64_t>(), h->pair_ep.as<uint32_t>(), n, h->slots.as<IndexSlot>(), cap,
                                    h->postings.as<uint32_t>(), h->idx_scratch.as<uint32_t>(), h->idx_cursor.as<uint32_t>(),
                                    h->idx_special.as<IndexSlot>(), h->intern_keys.as<uint64_t>(), h->intern_vals.as<uint32_t>(), s, &launches));
        CUDA_TRY(cudaMemcpyAsync(&h->idx_special_host, h->idx_special.p, sizeof(IndexSlot), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(cursor, h->idx_cursor.p, sizeof cursor, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        uint64_t true_distinct = std::max<uint64_t>(1, cursor[2]);
        const uint64_t load2 = h->index_load ? (uint64_t)h->index_load : (4 * true_distinct * sizeof(IndexSlot) <= (48ull << 20) ? 4 : 2);
        uint64_t tight = 16;
        while (tight < load2 * true_distinct) tight <<= 1;
        if (tight >= cap) break;            // already as small as the load factor allows
        distinct = true_distinct;           // rebuild once into a table a quarter (or less) of the size
    }
    h->idx_capacity = cap;
    h->idx_pairs = n;
    h->stats.index_pairs = n;
    h->stats.index_hashes = cursor[2];
    h->stats.index_slots = cap;
    // patch bookkeeping: pending-list heads idle, posting cursor, used slots
    CUDA_TRY(cudaMemsetAsync(h->idx_scratch.p, 0xFF, sizeof(uint32_t) * cap, s));
    h->post_cap = h->postings.cap / sizeof(uint32_t);
    h->post_used = cursor[0];
    h->rt_slots_used = cursor[2];
    h->rt_from_store = from_store;
    h->patches_since_build = 0;
    return EPP_OK;
}

static int32_t build_device_index(epp_engine *h, const uint64_t *hashes, const uint32_t *eps, uint64_t n,
                                  uint64_t n_distinct_hint) {
    cudaStream_t s = h->slot[0].stream;
    CUDA_TRY(h->pair_hash.reserve(sizeof(uint64_t) * std::max<uint64_t>(n, 1), &h->dev_bytes));
    CUDA_TRY(h->pair_ep.reserve(sizeof(uint32_t) * std::max<uint64_t>(n, 1), &h->dev_bytes));
    if (n) {
        CUDA_TRY(cudaMemcpyAsync(h->pair_hash.p, hashes, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(h->pair_ep.p, eps, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, s));
    }
    return build_index_from_device_pairs(h, n, n_distinct_hint);
}

// Applies the queued epp_index_add calls to the device store as one batch (sub-batches bound the ordering scratch).
static int32_t flush_add_queue(epp_engine *h) {
    const size_t M = h->q_ep.size();
    if (M == 0) return EPP_OK;
    cudaStream_t s = h->slot[0].stream;
    CUDA_TRY(h->q_ep_dev.reserve(sizeof(uint32_t) * M, &h->dev_bytes));
    CUDA_TRY(h->q_n_dev.reserve(sizeof(uint32_t) * M, &h->dev_bytes));
    CUDA_TRY(h->q_nb_dev.reserve(sizeof(int32_t) * M, &h->dev_bytes));
    CUDA_TRY(h->q_src_dev.reserve(sizeof(uint64_t) * M, &h->dev_bytes));
    CUDA_TRY(h->q_hashes_dev.reserve(sizeof(uint64_t) * std::max<size_t>(1, h->q_hashes.size()), &h->dev_bytes));
    CUDA_TRY(cudaMemcpyAsync(h->q_ep_dev.p, h->q_ep.data(), sizeof(uint32_t) * M, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->q_n_dev.p, h->q_n.data(), sizeof(uint32_t) * M, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->q_nb_dev.p, h->q_nb.data(), sizeof(int32_t) * M, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->q_src_dev.p, h->q_src.data(), sizeof(uint64_t) * M, cudaMemcpyHostToDevice, s));
    if (!h->q_hashes.empty())
        CUDA_TRY(cudaMemcpyAsync(h->q_hashes_dev.p, h->q_hashes.data(), sizeof(uint64_t) * h->q_hashes.size(), cudaMemcpyHostToDevice, s));
    const size_t kSub = 1u << 18;
    for (size_t c0 = 0; c0 < M; c0 += kSub) {
        StoreCalls c;
        c.M = (uint32_t)std::min(kSub, M - c0);
        c.ep = h->q_ep_dev.as<uint32_t>() + c0;
        c.n = h->q_n_dev.as<uint32_t>() + c0;
        c.nb = h->q_nb_dev.as<int32_t>() + c0;
        c.src = h->q_src_dev.as<uint64_t>() + c0;
        c.hashes = h->q_hashes_dev.as<uint64_t>();
        CUDA_TRY(h->store->apply(c, s));
    }
    h->q_ep.clear();
    h->q_n.clear();
    h->q_nb.clear();
    h->q_src.clear();
    h->q_hashes.clear();
    return EPP_OK;
}

static double ms_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

static int32_t commit_locked(epp_engine *h) {
    if (h->snapshot_mode) return EPP_OK;
    if (!h->q_ep.empty() || h->store->dirty()) EPP_TRY(join_streams(h));     // the table is about to change under stream 1
    EPP_TRY(flush_add_queue(h));
    if (!h->store->dirty()) return EPP_OK;
    auto t0 = std::chrono::steady_clock::now();
    cudaStream_t s = h->slot[0].stream;
    bool patched = false;
    if (h->patch_enabled && h->rt_from_store && h->idx_capacity && h->patches_since_build < 256) {
        ReadTableRef rt;
        rt.slots = h->slots.as<IndexSlot>();
        rt.capacity = h->idx_capacity;
        rt.slots_used = h->rt_slots_used;
        rt.postings = h->postings.as<uint32_t>();
        rt.post_cap = h->post_cap;
        rt.post_used = h->post_used;
        rt.head = h->idx_scratch.as<uint32_t>();
        rt.intern_keys = h->intern_keys.as<uint64_t>();
        rt.intern_vals = h->intern_vals.as<uint32_t>();
        uint64_t np = 0, nu = 0, npost = 0;
        CUDA_TRY(h->store->patch_read_table(rt, h->idx_pairs, s, &patched, &np, &nu, &npost));
        if (patched) {
            h->idx_pairs = np;
            h->rt_slots_used = nu;
            h->post_used = npost;
            h->stats.index_pairs = np;
            h->stats.index_hashes = nu;
            h->patches_since_build++;
            h->stats.last_index_patched = 1;
        }
    }
    if (!patched) {
        uint64_t n = 0;
        CUDA_TRY(h->store->export_pairs(h->pair_hash, h->pair_ep, &n, &h->dev_bytes, s));
        EPP_TRY(build_index_from_device_pairs(h, n, n, true));
        h->stats.last_index_patched = 0;
    }
    CUDA_TRY(h->store->touch_reset(s));
    h->store->mark_clean();
    h->stats.last_index_build_ms = ms_since(t0);
    return EPP_OK;
}

static IndexView index_view(epp_engine *h) {
    IndexView v;
    v.slots = h->idx_capacity ? h->slots.as<IndexSlot>() : nullptr;
    v.postings = h->postings.as<uint32_t>();
    v.mask = h->idx_capacity ? h->idx_capacity - 1 : 0;
    v.special = h->idx_special_host;
    if (!h->idx_capacity) v.special.cnt = 0;
    v.ep_begin = h->shard_begin;
    v.ep_end = h->shard_end;
    return v;
}

extern "C" int32_t epp_index_add(epp_engine *h, uint32_t ep, int32_t n, const uint64_t *hashes, int32_t num_gpu_blocks) {
    if (!h || n < 0 || (n > 0 && !hashes)) return fail(EPP_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->snapshot_mode) return fail(EPP_ERR_STATE, "index holds a bulk snapshot; incremental adds need an empty or incrementally built index");
    if (ep >= (uint32_t)h->cfg.max_endpoints) return fail(EPP_ERR_INVALID, "endpoint slot %u out of range [0,%d)", ep, h->cfg.max_endpoints);
    h->q_ep.push_back(ep);
    h->q_n.push_back((uint32_t)n);
    h->q_nb.push_back(num_gpu_blocks);
    h->q_src.push_back((uint64_t)h->q_hashes.size());
    h->q_hashes.insert(h->q_hashes.end(), hashes, hashes + n);
    if (h->q_hashes.size() >= (1u << 26) || h->q_ep.size() >= (1u << 22)) {      // bound the host queue
        EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
        EPP_TRY(flush_add_queue(h));
    }
    return EPP_OK;
}

extern "C" int32_t epp_index_remove_endpoint(epp_engine *h, uint32_t ep) {
    if (!h) return fail(EPP_ERR_INVALID, "NULL engine");
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->snapshot_mode) return fail(EPP_ERR_STATE, "index holds a bulk snapshot; reload it without the endpoint instead");
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    EPP_TRY(flush_add_queue(h));                    // calls apply in the order they were made
    CUDA_TRY(h->store->remove_endpoint(ep, h->slot[0].stream));
    return EPP_OK;
}

// CleanUpInactivePods (approximateprefix/plugin.go:99-122).
extern "C" int32_t epp_index_retain_endpoints(epp_engine *h, int32_t n, const uint32_t *active_ids) {
    if (!h || n < 0 || (n > 0 && !active_ids)) return fail(EPP_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->snapshot_mode) return fail(EPP_ERR_STATE, "index holds a bulk snapshot; reload it without the endpoints instead");
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    EPP_TRY(flush_add_queue(h));
    const size_t E = (size_t)h->cfg.max_endpoints;
    std::vector<uint8_t> active(E, 0);
    for (int32_t i = 0; i < n; i++) {
        if (active_ids[i] >= E) return fail(EPP_ERR_INVALID, "endpoint slot %u out of range [0,%zu)", active_ids[i], E);
        active[active_ids[i]] = 1;
    }
    cudaStream_t s = h->slot[0].stream;
    CUDA_TRY(h->q_ep_dev.reserve(E, &h->dev_bytes));
    CUDA_TRY(cudaMemcpyAsync(h->q_ep_dev.p, active.data(), E, cudaMemcpyHostToDevice, s));
    CUDA_TRY(h->store->retain_endpoints(h->q_ep_dev.as<uint8_t>(), s));
    return EPP_OK;
}

extern "C" int32_t epp_index_load_snapshot(epp_engine *h, uint64_t n_pairs, const uint64_t *hashes, const uint32_t *eps) {
    if (!h || (n_pairs && (!hashes || !eps))) return fail(EPP_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    h->q_ep.clear(); h->q_n.clear(); h->q_nb.clear(); h->q_src.clear(); h->q_hashes.clear();
    CUDA_TRY(h->store->clear(h->slot[0].stream));
    h->store->mark_clean();
    EPP_TRY(build_device_index(h, hashes, eps, n_pairs, n_pairs));
    h->snapshot_mode = n_pairs > 0;
    return EPP_OK;
}

extern "C" int32_t epp_index_commit(epp_engine *h) {
    if (!h) return fail(EPP_ERR_INVALID, "NULL engine");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    return commit_locked(h);
}

extern "C" int32_t epp_index_get(epp_engine *h, uint64_t hash, uint32_t *out_eps, int32_t cap, int32_t *out_n) {
    if (!h || !out_n || cap < 0 || (cap > 0 && !out_eps)) return fail(EPP_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    EPP_TRY(commit_locked(h));
    cudaStream_t s = h->slot[0].stream;
    int32_t dcap = std::min(cap, 4096);
    uint32_t *d_eps = h->get_out.as<uint32_t>();
    int32_t *d_n = reinterpret_cast<int32_t *>(d_eps + 4096);
    CUDA_TRY(launch_index_get(index_view(h), hash, d_eps, dcap, d_n, s));
    CUDA_TRY(cudaMemcpyAsync(out_n, d_n, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    int32_t ncopy = std::min(*out_n, dcap);
    if (ncopy > 0) CUDA_TRY(cudaMemcpy(out_eps, d_eps, sizeof(uint32_t) * ncopy, cudaMemcpyDeviceToHost));
    return EPP_OK;
}

// ------------------------------------------------------------------------------------------------
// batches
// ------------------------------------------------------------------------------------------------
struct BatchView {
    int64_t R = 0;
    bool device = false;
    bool async = false;            // EPP_BATCH_ASYNC (device batches, epp_schedule only)
    const uint8_t *data = nullptr;
    const uint64_t *offsets = nullptr;
    const uint64_t *lengths = nullptr;
    uint64_t uniform_len = 0;
    const uint32_t *model_ids = nullptr;
    const uint8_t *multimodal = nullptr;
    uint64_t total_bytes = 0;      // host batches only
    uint64_t offsets_or_bits = 0;
};

static int32_t check_batch(epp_engine *h, const epp_batch *b, BatchView &v) {
    if (!b) return fail(EPP_ERR_INVALID, "batch is NULL");
    if (b->n_requests < 0) return fail(EPP_ERR_INVALID, "n_requests < 0");
    if (h->n_models == 0) return fail(EPP_ERR_STATE, "no model registered (epp_model_register)");
    v.R = b->n_requests;
    v.device = (b->flags & EPP_BATCH_DEVICE_PTRS) != 0;
    v.async = (b->flags & EPP_BATCH_ASYNC) != 0;
    if (v.async && !v.device) return fail(EPP_ERR_INVALID, "EPP_BATCH_ASYNC needs EPP_BATCH_DEVICE_PTRS");
    v.data = reinterpret_cast<const uint8_t *>(b->data);
    v.offsets = b->offsets;
    v.lengths = b->lengths;
    if (v.lengths && !v.offsets) return fail(EPP_ERR_INVALID, "lengths needs offsets");
    v.uniform_len = b->uniform_len;
    v.model_ids = b->model_ids;
    v.multimodal = b->multimodal;
    if (v.R == 0) return EPP_OK;
    if (!v.offsets && v.uniform_len > 0 && !v.data) return fail(EPP_ERR_INVALID, "data is NULL");
    if (!v.device) {
        if (v.offsets) {
            uint64_t bits = 0;
            for (int64_t r = 0; r < v.R; r++) {
                if (v.offsets[r + 1] < v.offsets[r]) return fail(EPP_ERR_INVALID, "offsets not monotonic at request %lld", (long long)r);
                if (v.lengths && v.offsets[r] + v.lengths[r] > v.offsets[r + 1]) {
                    // EPP_BATCH_LENGTHS_EXCEED_ROWS: the row must still hold every byte hashPrompt reads
                    const uint64_t cap = (uint64_t)h->cfg.max_prefix_blocks * (uint64_t)h->cfg.block_size_tokens * 4;
                    const uint64_t need = std::min<uint64_t>(v.lengths[r], cap);
                    if (!(b->flags & EPP_BATCH_LENGTHS_EXCEED_ROWS) || v.offsets[r] + need > v.offsets[r + 1])
                        return fail(EPP_ERR_INVALID, "request %lld: offsets[r] + lengths[r] exceeds offsets[r+1]", (long long)r);
                }
                bits |= v.offsets[r];
            }
            v.offsets_or_bits = bits;
            v.total_bytes = v.offsets[v.R] - v.offsets[0];
            if (v.total_bytes && !v.data) return fail(EPP_ERR_INVALID, "data is NULL");
        } else {
            v.total_bytes = (uint64_t)v.R * v.uniform_len;
        }
        if (v.model_ids)
            for (int64_t r = 0; r < v.R; r++)
                if ((int)v.model_ids[r] >= h->n_models) return fail(EPP_ERR_INVALID, "request %lld: unknown model id %u", (long long)r, v.model_ids[r]);
    }
    return EPP_OK;
}

static int32_t reserve_batch(epp_engine *h, int64_t R) {
    size_t B = (size_t)h->cfg.max_prefix_blocks;
    size_t n = (size_t)std::max<int64_t>(R, 1);
    CUDA_TRY(h->hashes.reserve(sizeof(uint64_t) * n * B, &h->dev_bytes));
    CUDA_TRY(h->nblocks.reserve(sizeof(int32_t) * n, &h->dev_bytes));
    CUDA_TRY(h->eff_len.reserve(sizeof(int64_t) * n, &h->dev_bytes));
    CUDA_TRY(h->in_len.reserve(sizeof(int64_t) * n, &h->dev_bytes));
    CUDA_TRY(h->decisions.reserve(sizeof(epp_decision) * n, &h->dev_bytes));
    CUDA_TRY(h->details.reserve(sizeof(epp_decision_detail) * n, &h->dev_bytes));
    CUDA_TRY(h->offsets.reserve(sizeof(uint64_t) * (n + 1), &h->dev_bytes));
    CUDA_TRY(h->lengths.reserve(sizeof(uint64_t) * n, &h->dev_bytes));
    CUDA_TRY(h->model_ids.reserve(sizeof(uint32_t) * n, &h->dev_bytes));
    for (int i = 0; i < 2; i++) CUDA_TRY(h->slot[i].overflow_list.reserve(sizeof(int32_t) * n, &h->dev_bytes));
    return EPP_OK;
}

// What one launch sequence works on: requests [r0, r1) whose bytes are resident at `data_base + offsets[r]`.
struct Work {
    int64_t r0, r1;
    const uint8_t *data_base;
    const uint64_t *offsets_dev;   // indexed by absolute request id, or nullptr
    const uint64_t *lengths_dev;   // absolute, or nullptr
    uint64_t uniform_len;
    const uint32_t *model_ids_dev; // absolute, or nullptr
    uint64_t offsets_or_bits;      // OR of all offsets (alignment of the batch layout)
    uint64_t *hashes_out;          // row r0 of the destination
    int32_t *nblocks_out;
    const uint8_t *multimodal_dev = nullptr;   // absolute, or nullptr
};

static HashParams hash_params(epp_engine *h, const Work &w) {
    HashParams p;
    int64_t n = w.r1 - w.r0;
    p.data = w.data_base;
    p.offsets = w.offsets_dev ? w.offsets_dev + w.r0 : nullptr;
    p.lengths = w.lengths_dev ? w.lengths_dev + w.r0 : nullptr;
    p.uniform_len = w.uniform_len;
    p.model_ids = w.model_ids_dev ? w.model_ids_dev + w.r0 : nullptr;
    p.seeds = h->seeds.as<uint64_t>();
    p.R = n;
    p.block_bytes = h->cfg.block_size_tokens * 4;
    p.max_blocks = h->cfg.max_prefix_blocks;
    p.hashes = w.hashes_out;
    p.nblocks = w.nblocks_out;
    p.eff_len = h->eff_len.as<int64_t>() + w.r0;
    p.in_len = h->in_len.as<int64_t>() + w.r0;
    p.offsets_or_bits = w.offsets_or_bits;
    p.sm_count = h->sm_count;
    p.staged = h->staged;
    return p;
}

// General-evaluation view; the top-k rows are attached by the caller (they depend on where the batch's rows live).
static GenView gen_view(epp_engine *h) {
    GenView g;
    memset(&g, 0, sizeof g);
    g.on = h->general ? 1 : 0;
    g.pool = pool_arrays(h);
    return g;
}
// Rows [r0, ...) of the top-k lists of the call in progress, in `base` arrays of row stride k.
static void attach_topk(epp_engine *h, GenView &g, uint32_t *const picks[kMaxProfiles], double *scores, int64_t r0) {
    const int32_t k = h->cfg.pick_k;
    g.topk = k;
    for (int i = 0; i < kMaxProfiles; i++) g.topk_picks[i] = picks[i] ? picks[i] + (size_t)r0 * (size_t)k : nullptr;
    g.topk_scores = scores ? scores + (size_t)r0 * (size_t)k : nullptr;
}

static PickParams pick_params(epp_engine *h, const Work &w, epp_decision *out, epp_decision_detail *detail,
                              int32_t *out_match) {
    PickParams p;
    p.R = w.r1 - w.r0;
    p.E = h->cfg.max_endpoints;
    p.max_blocks = h->cfg.max_prefix_blocks;
    p.block_size_tokens = h->cfg.block_size_tokens;
    p.n_profiles = h->n_profiles;
    p.encode_on = h->cfg.encode_enabled ? 1 : 0;
    p.multimodal = w.multimodal_dev ? w.multimodal_dev + w.r0 : nullptr;
    p.always_disagg = h->cfg.always_disagg;
    p.non_cached_tokens = h->cfg.non_cached_tokens;
    for (int i = 0; i < h->n_alloc_profiles; i++) p.prof[i] = profile_dev(h, i);
    p.tie_seed = h->shard_end != 0xFFFFFFFFu || h->shard_begin != 0 ? 0 : h->cfg.tie_seed;      // sharded mode: lowest slot
    p.tie_base = h->stats.n_decisions + (uint64_t)w.r0;
    p.hashes = w.hashes_out;
    p.nblocks = w.nblocks_out;
    p.in_len = h->in_len.as<int64_t>() + w.r0;
    p.model_ids = w.model_ids_dev ? w.model_ids_dev + w.r0 : nullptr;
    p.lora = lora_dev(h);
    p.index = index_view(h);
    p.out = out;
    p.detail = detail;
    p.out_match = out_match;
    p.work_counters = nullptr;
    p.global_masks = nullptr;
    p.mask_words = 0;
    p.shard_out = nullptr;
    p.req_list = nullptr;
    p.req_list_n = nullptr;
    p.overflow_list = nullptr;
    p.overflow_n = nullptr;
    p.gen = gen_view(h);
    return p;
}

// Device-side offsets alignment probe for device-pointer batches.
static int32_t device_offsets_or_bits(epp_engine *h, const uint64_t *offsets_dev, int64_t n, cudaStream_t s, uint64_t *out) {
    CUDA_TRY(cudaMemsetAsync(h->flag.p, 0, sizeof(int), s));
    CUDA_TRY(launch_check_offsets_aligned(offsets_dev, n, h->flag.as<int>(), s));
    int res = 0;
    CUDA_TRY(cudaMemcpyAsync(&res, h->flag.p, sizeof res, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    *out = (uint64_t)res;
    return EPP_OK;
}

enum class Mode { HashOnly, Match, Schedule };

// Dense-counter kernel over the requests a sparse kernel could not finish (their matched-endpoint set overflowed its
// per-request map).  No host round trip: the kernel reads the list length on the device.
static int32_t launch_overflow_pass(epp_engine *h, Slot &sl, PickParams pp, int *launches) {
    cudaStream_t s = sl.stream;
    pp.req_list = sl.overflow_list.as<int32_t>();
    pp.req_list_n = sl.overflow_n.as<int32_t>();
    pp.overflow_list = nullptr;
    pp.overflow_n = nullptr;
    pp.work_counters = nullptr;            // already counted by the sparse pass
    CUDA_TRY(launch_match_pick(pp, h->pick_global ? sl.pick_scratch.as<uint32_t>() : nullptr, h->pick_grid, h->pick_smem, s, launches));
    return EPP_OK;
}

// Standalone match/pick over hashes already in HBM: warp-per-request sparse kernel + overflow pass (or the dense
// kernel alone for Produce-parity rows / A-B runs).
static int32_t launch_match(epp_engine *h, Slot &sl, PickParams pp, int *launches) {
    cudaStream_t s = sl.stream;
    uint32_t *gs = h->pick_global ? sl.pick_scratch.as<uint32_t>() : nullptr;
    if (pp.out_match || pp.gen.on) {
        CUDA_TRY(launch_match_pick(pp, gs, h->pick_grid, h->pick_smem, s, launches));
        return EPP_OK;
    }
    CUDA_TRY(cudaMemsetAsync(sl.overflow_n.p, 0, sizeof(int32_t), s));
    pp.overflow_list = sl.overflow_list.as<int32_t>();
    pp.overflow_n = sl.overflow_n.as<int32_t>();
    CUDA_TRY(launch_match_pick_sparse(pp, h->sm_count, s, launches));
    return launch_overflow_pass(h, sl, pp, launches);
}

// The whole cycle for one Work item: hash kernel, then the standalone match / score / pick kernel (every fused or
// concurrently scheduled arrangement of the two was measured slower).
static int32_t launch_cycle(epp_engine *h, Slot &sl, const HashParams &hp, PickParams pp, int *launches, cudaEvent_t *ev) {
    CUDA_TRY(launch_hash_prompts(hp, sl.stream, launches, ev));
    return launch_match(h, sl, pp, launches);
}

// Completes the device batch in flight on stream 0 and reads its per-kernel CUDA-event times and work counters.
static int32_t finish_async(epp_engine *h) {
    EPP_TRY(join_streams(h));
    CUDA_TRY(cudaStreamSynchronize(h->slot[0].stream));
    if (!h->async_pending) return EPP_OK;
    h->async_pending = false;
    float t[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; i++) cudaEventElapsedTime(&t[i], h->ev[i], h->ev[i + 1]);
    for (int i = 0; i < 8; i++) h->stats.last_kernel_ms[i] = i < 4 ? t[i] : 0.0;
    h->stats.last_hash_ms = t[0] + t[1] + t[2];
    h->stats.last_match_pick_ms = t[3];
    h->stats.last_kernels_ms = t[0] + t[1] + t[2] + t[3];
    h->stats.last_h2d_ms = h->stats.last_d2h_ms = 0;
    h->stats.last_probes = h->wc_host[0];
    h->stats.last_postings = h->wc_host[1];
    h->stats.last_kernel_launches = (uint64_t)h->async_launches;
    return EPP_OK;
}

// ---- small host batches (cycle_small.cu) -------------------------------------------------------------------------
static int32_t small_reserve(epp_engine *h, int64_t R) {
    if (R <= h->small_cap) return EPP_OK;
    const int64_t cap = std::max<int64_t>(64, std::min<int64_t>(h->small_max, R * 2));
    if (h->small_host) { CUDA_TRY(cudaStreamSynchronize(h->slot[0].stream)); CUDA_TRY(cudaFreeHost(h->small_host)); h->small_host = nullptr; h->small_cap = 0; }
    const size_t n = (size_t)cap;
    const size_t per_set = sizeof(uint64_t) * (n + 1) + sizeof(uint64_t) * n + sizeof(uint32_t) * n + ((n + 7) & ~(size_t)7);
    const size_t total = 2 * per_set + sizeof(epp_decision) * n + sizeof(epp_decision_detail) * n + sizeof(uint32_t) * n + 64;
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&h->small_host), total, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(h->small_host, 0, total);
    uint8_t *q = h->small_host;
    for (int i = 0; i < 2; i++) {
        h->small_offsets[i] = reinterpret_cast<uint64_t *>(q); q += sizeof(uint64_t) * (n + 1);
        h->small_lengths[i] = reinterpret_cast<uint64_t *>(q); q += sizeof(uint64_t) * n;
        h->small_models[i] = reinterpret_cast<uint32_t *>(q); q += sizeof(uint32_t) * n;
        h->small_mm[i] = q; q += (n + 7) & ~(size_t)7;
    }
    q = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(q) + 15) & ~(uintptr_t)15);
    h->small_dec = reinterpret_cast<epp_decision *>(q); q += sizeof(epp_decision) * n;
    h->small_det = reinterpret_cast<epp_decision_detail *>(q); q += sizeof(epp_decision_detail) * n;
    h->small_flags = reinterpret_cast<uint32_t *>(q);
    h->small_cap = cap;
    return EPP_OK;
}

// Device view of a pointer into pinned host memory (cudaHostAlloc / cudaHostRegister), or nullptr.
static const uint8_t *device_view_of_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    return reinterpret_cast<const uint8_t *>(a.devicePointer);
}

// Is this host batch one for the single-launch path?  Returns 0 = no; 1 = the kernel reads the prompts from the
// caller's pinned memory over PCIe (*data_dev = their device view); 2 = one DMA copy into HBM first (SM-issued PCIe
// reads top out near 20 GB/s, the copy engine reaches 55: beyond a handful of prompts the copy wins).
static int small_batch_mode(epp_engine *h, const BatchView &v, const uint8_t **data_dev, int *align) {
    if (v.device || h->small_max <= 0 || v.R > h->small_max || h->general) return 0;
    if (h->shard_begin != 0 || h->shard_end != 0xFFFFFFFFu) return 0;
    const int64_t bs = (int64_t)h->cfg.block_size_tokens * 4;
    if ((bs & 31) || bs > 64 * 1024 || (size_t)h->cfg.max_prefix_blocks > cycle_small_max_blocks()) return 0;
    const uint64_t layout_bits = v.offsets ? v.offsets_or_bits : v.uniform_len;
    if (v.R <= h->small_zc_max || !v.total_bytes) {
        *data_dev = v.total_bytes ? device_view_of_pinned(v.data) : nullptr;
        if (*data_dev || !v.total_bytes) {
            const uint64_t bits = reinterpret_cast<uintptr_t>(*data_dev) | layout_bits;
            *align = (bits & 31) == 0 ? 32 : ((bits & 15) == 0 ? 16 : 0);
            if (*align >= 16) return 1;
        }
    }
    // the staged copy keeps the caller's layout modulo 32 (see run_batch), so only the layout decides the alignment
    *align = (layout_bits & 31) == 0 ? 32 : ((layout_bits & 15) == 0 ? 16 : 0);
    return *align >= 16 ? 2 : 0;
}

// cuStreamWriteValue32 through the runtime's driver-entry-point lookup (no link against libcuda); nullptr when absent.
typedef CUresult (*StreamWriteValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static StreamWriteValue32 stream_write_value32() {
    static const StreamWriteValue32 fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<StreamWriteValue32>(p);
    }();
    return fn;
}

static int32_t run_small(epp_engine *h, const BatchView &v, int mode, const uint8_t *data_dev, int align,
                         epp_decision *out_dec, epp_decision_detail *out_detail) {
    const int64_t R = v.R;
    cudaStream_t s0 = h->slot[0].stream;
    EPP_TRY(reserve_batch(h, R));
    EPP_TRY(small_reserve(h, R));
    uint32_t epoch = h->small_epoch + 1;
    if (epoch > 0x7fffffffu) {
        epoch = 1;
        CUDA_TRY(cudaMemsetAsync(h->small_arrive.p, 0, sizeof(uint32_t), s0));
        CUDA_TRY(cudaStreamSynchronize(s0));
    }
    h->small_epoch = epoch;
    const uint32_t *arrive = nullptr;
    if (mode == 2) {
        const uint64_t start = v.offsets ? v.offsets[0] : 0;
        CUDA_TRY(h->slot[0].data.reserve(v.total_bytes + 64, &h->dev_bytes));
        uint8_t *stage = h->slot[0].data.as<uint8_t>() + (start & 31);
        data_dev = reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(stage) - (uintptr_t)start);   // + offsets[r]
        if (h->small_pipe_min > 0 && R >= h->small_pipe_min &&
            cycle_small_stages_prompt(h->cfg.max_prefix_blocks, h->cfg.block_size_tokens * 4)) {
            // The copy goes to the second stream with a stream-ordered 32-bit write of the epoch behind it; the kernel is
            // launched on s0 at once and thread 0 of every CTA waits for that word, so the launch latency (and the stream's
            // copy -> kernel hand-over) hides under the transfer.  The staging buffer is free although the kernel of the batch
            // before may still be walking chains: it keeps a whole prompt in shared memory (cycle_small_stages_prompt),
            // so every request raised its flag after its last read of these bytes.
            cudaStream_t s1 = h->slot[1].stream;
            uint32_t *arr = h->small_arrive.as<uint32_t>();
            CUDA_TRY(cudaMemcpyAsync(stage, v.data + start, v.total_bytes, cudaMemcpyHostToDevice, s1));
            const StreamWriteValue32 wv = h->small_no_write_value ? nullptr : stream_write_value32();
            if (!wv || wv(s1, reinterpret_cast<CUdeviceptr>(arr), epoch, 0) != CUDA_SUCCESS) {
                h->small_no_write_value = true;    // no stream memory operations in this driver / context: a 4-byte copy does it
                *h->small_epoch_host = epoch;      // read by the copy below before this (synchronous) call returns
                CUDA_TRY(cudaMemcpyAsync(arr, h->small_epoch_host, sizeof(uint32_t), cudaMemcpyHostToDevice, s1));
            }
            arrive = arr;
            h->s1_unjoined = true;
        } else {
            CUDA_TRY(cudaMemcpyAsync(stage, v.data + start, v.total_bytes, cudaMemcpyHostToDevice, s0));
        }
    }
    const int set = (int)(epoch & 1);              // the kernel of the batch before may still be finishing its chains
    if (v.offsets) memcpy(h->small_offsets[set], v.offsets, sizeof(uint64_t) * (size_t)(R + 1));
    if (v.lengths) memcpy(h->small_lengths[set], v.lengths, sizeof(uint64_t) * (size_t)R);
    if (v.model_ids) memcpy(h->small_models[set], v.model_ids, sizeof(uint32_t) * (size_t)R);
    if (v.multimodal) memcpy(h->small_mm[set], v.multimodal, (size_t)R);
    Work w;
    w.r0 = 0;
    w.r1 = R;
    w.data_base = data_dev;
    w.offsets_dev = v.offsets ? h->small_offsets[set] : nullptr;       // pinned + mapped: host address == device address (UVA)
    w.lengths_dev = v.lengths ? h->small_lengths[set] : nullptr;
    w.uniform_len = v.uniform_len;
    w.model_ids_dev = v.model_ids ? h->small_models[set] : nullptr;
    w.offsets_or_bits = v.offsets_or_bits;
    w.hashes_out = h->hashes.as<uint64_t>();
    w.nblocks_out = h->nblocks.as<int32_t>();
    w.multimodal_dev = v.multimodal ? h->small_mm[set] : nullptr;
    PickParams pp = pick_params(h, w, h->decisions.as<epp_decision>(), h->details.as<epp_decision_detail>(), nullptr);
    pp.overflow_list = h->slot[0].overflow_list.as<int32_t>();
    pp.overflow_n = h->small_overflow_n.as<int32_t>();
    SmallOut so{h->small_dec, h->small_det, h->small_flags, epoch, arrive};
    int launches = 0;
    CUDA_TRY(cudaEventRecord(h->ev[0], s0));
    CUDA_TRY(launch_cycle_small(hash_params(h, w), pp, so, align, s0, &launches));
    CUDA_TRY(cudaEventRecord(h->ev[1], s0));

    // wait for the flag word of every request (written after a system-scope fence behind its decision)
    volatile uint32_t *flags = h->small_flags;
    uint32_t overflowed = 0, spins = 0;
    bool finished_once = false;
    for (int64_t next = 0; next < R;) {
        const uint32_t x = flags[next];
        if ((x & 0x7fffffffu) == epoch) { overflowed |= x >> 31; next++; continue; }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if ((++spins & 0x3fff) == 0) {             // a faulted launch never raises its flags
            const cudaError_t q = cudaStreamQuery(s0);
            if (q == cudaSuccess) {
                if (finished_once) return fail(EPP_ERR_CUDA, "small-batch kernel finished without publishing request %lld", (long long)next);
                finished_once = true;
            } else if (q != cudaErrorNotReady) {
                return fail(EPP_ERR_CUDA, "small-batch kernel failed: %s", cudaGetErrorString(q));
            }
        }
    }
    if (overflowed) {
        // > 32 endpoints hold a part of some prompt: those requests go through the dense-counter kernel, then the whole
        // batch is read back the ordinary way
        Slot &sl = h->slot[0];
        PickParams po = pp;
        po.req_list = sl.overflow_list.as<int32_t>();
        po.req_list_n = h->small_overflow_n.as<int32_t>();
        po.overflow_list = nullptr;
        po.overflow_n = nullptr;
        CUDA_TRY(launch_match_pick(po, h->pick_global ? sl.pick_scratch.as<uint32_t>() : nullptr, h->pick_grid, h->pick_smem, s0, &launches));
        CUDA_TRY(cudaMemsetAsync(h->small_overflow_n.p, 0, sizeof(int32_t), s0));
        CUDA_TRY(cudaMemcpyAsync(h->small_dec, h->decisions.p, sizeof(epp_decision) * (size_t)R, cudaMemcpyDeviceToHost, s0));
        CUDA_TRY(cudaMemcpyAsync(h->small_det, h->details.p, sizeof(epp_decision_detail) * (size_t)R, cudaMemcpyDeviceToHost, s0));
        CUDA_TRY(cudaStreamSynchronize(s0));
    }
    memcpy(out_dec, h->small_dec, sizeof(epp_decision) * (size_t)R);
    if (out_detail) memcpy(out_detail, h->small_det, sizeof(epp_decision_detail) * (size_t)R);
    h->small_stats_pending = true;
    h->stats.last_hash_ms = h->stats.last_match_pick_ms = 0;
    h->stats.last_kernel_launches = (uint64_t)launches;
    return EPP_OK;
}

// Runs hashing (+ match/pick) for a batch.  Host batches are split into chunks whose H2D copy overlaps the
// kernels of the previous chunk (two streams, two staging buffers); device batches run in one pass.
static int32_t run_batch(epp_engine *h, const BatchView &v, Mode mode, uint64_t *out_hashes, int32_t *out_nblocks,
                         epp_decision *out_dec, epp_decision_detail *out_detail, int32_t *out_match,
                         int32_t *out_total) {
    const int64_t R = v.R;
    const size_t B = (size_t)h->cfg.max_prefix_blocks;
    if (R == 0) return EPP_OK;
    if (mode != Mode::HashOnly) {
        if (!h->pool_ready) return fail(EPP_ERR_STATE, "epp_pool_set has not been called");
        // index updates become visible here -- or, with index_commit_interval_us, at most that much later: the reference
        // applies PreRequest in a goroutine of its own "to avoid blocking the request path" (approximateprefix/plugin.go:189-194),
        // so a scheduling cycle may well run against an index that lacks the last few picks
        const bool due = h->cfg.index_commit_interval_us <= 0 || !h->committed_once ||
                         ms_since(h->last_commit) * 1000.0 >= (double)h->cfg.index_commit_interval_us;
        if (due) {
            EPP_TRY(commit_locked(h));
            h->last_commit = std::chrono::steady_clock::now();
            h->committed_once = true;
        }
    }
    if (h->async_pending && !(v.device && v.async)) EPP_TRY(finish_async(h));
    // top-k lists of the call in progress: the caller's device arrays, or device staging for a host batch
    const epp_topk_out *tk = mode == Mode::Schedule ? h->cur_topk : nullptr;
    const size_t K = (size_t)h->cfg.pick_k;
    uint32_t *tk_picks[kMaxProfiles] = {nullptr, nullptr, nullptr};
    uint32_t *tk_host[kMaxProfiles] = {nullptr, nullptr, nullptr};
    double *tk_scores = nullptr;
    if (tk) {
        tk_host[0] = tk->primary; tk_host[1] = tk->prefill; tk_host[2] = tk->encode;
        if (v.device) {
            for (int i = 0; i < kMaxProfiles; i++) tk_picks[i] = tk_host[i];
            tk_scores = tk->primary_scores;
        } else {
            for (int i = 0; i < kMaxProfiles; i++)
                if (tk_host[i]) {
                    CUDA_TRY(h->topk_picks[i].reserve(sizeof(uint32_t) * (size_t)R * K, &h->dev_bytes));
                    tk_picks[i] = h->topk_picks[i].as<uint32_t>();
                }
            if (tk->primary_scores) {
                CUDA_TRY(h->topk_scores.reserve(sizeof(double) * (size_t)R * K, &h->dev_bytes));
                tk_scores = h->topk_scores.as<double>();
            }
        }
    }
    const bool pipelined = v.device && v.async && mode == Mode::Schedule && h->dev_chunks > 1 && R >= 2 * 4096 && !h->pick_global;
    if (!pipelined) EPP_TRY(join_streams(h));
    EPP_TRY(reserve_batch(h, R));
    int launches = 0;
    cudaStream_t s0 = h->slot[0].stream;

    if (v.device) {
        uint64_t or_bits = 0;
        if (v.offsets) EPP_TRY(device_offsets_or_bits(h, v.offsets, R, s0, &or_bits));
        Work w{0, R, v.data, v.offsets, v.lengths, v.uniform_len, v.model_ids, or_bits,
               (mode == Mode::HashOnly && out_hashes) ? out_hashes : h->hashes.as<uint64_t>(),
               (mode == Mode::HashOnly && out_nblocks) ? out_nblocks : (mode == Mode::Match && out_total ? out_total : h->nblocks.as<int32_t>()),
               v.multimodal};
        CUDA_TRY(cudaMemsetAsync(h->work_counters.p, 0, sizeof(unsigned long long) * 2, s0));
        const int64_t min_chunk = 4096;
        if (pipelined) {
            // Throughput mode: the batch runs as C chunks alternating between the engine's two streams, so the hash
            // kernel of one chunk (HBM / integer bound) shares the SMs with the match kernel of another (latency bound).
            // Consecutive pipelined batches of the same shape are not joined: every row range of the engine's buffers is
            // only ever touched by one stream, and the next entry point of any other kind joins the streams first.
            const int64_t C = std::min<int64_t>(h->dev_chunks, R / min_chunk);
            const int64_t per = ((R + C - 1) / C + 31) & ~(int64_t)31;
            cudaStream_t s1 = h->slot[1].stream;
            if (h->s1_unjoined && (h->pipe_R != R || h->pipe_per != per)) EPP_TRY(join_streams(h));
            for (int i = 0; i < 4; i++) CUDA_TRY(cudaEventRecord(h->ev[i], s0));
            if (!h->s1_unjoined) {
                CUDA_TRY(cudaEventRecord(h->slot[0].done, s0));
                CUDA_TRY(cudaStreamWaitEvent(s1, h->slot[0].done, 0));
            }
            epp_decision *dec_base = out_dec ? out_dec : h->decisions.as<epp_decision>();
            int k = 0;
            for (int64_t r0 = 0; r0 < R; r0 += per, k++) {
                const int64_t r1 = std::min(R, r0 + per);
                Work w{r0, r1, v.offsets ? v.data : v.data + (uint64_t)r0 * v.uniform_len, v.offsets, v.lengths, v.uniform_len,
                       v.model_ids, or_bits, h->hashes.as<uint64_t>() + (size_t)r0 * B, h->nblocks.as<int32_t>() + r0, v.multimodal};
                PickParams pp = pick_params(h, w, dec_base + r0, out_detail ? out_detail + r0 : nullptr, nullptr);
                if (tk) attach_topk(h, pp.gen, tk_picks, tk_scores, r0);
                Slot &sl = h->slot[k & 1];
                EPP_TRY(launch_cycle(h, sl, hash_params(h, w), pp, &launches, nullptr));
            }
            h->s1_unjoined = true;
            h->pipe_R = R;
            h->pipe_per = per;
        } else if (mode == Mode::HashOnly) {
            CUDA_TRY(launch_hash_prompts(hash_params(h, w), s0, &launches, h->ev));
        } else {
            epp_decision *dec = (mode == Mode::Schedule && out_dec) ? out_dec : h->decisions.as<epp_decision>();
            PickParams pp = pick_params(h, w, dec, mode == Mode::Schedule ? out_detail : nullptr, mode == Mode::Match ? out_match : nullptr);
            pp.work_counters = h->work_counters.as<unsigned long long>();
            if (tk) attach_topk(h, pp.gen, tk_picks, tk_scores, 0);
            EPP_TRY(launch_cycle(h, h->slot[0], hash_params(h, w), pp, &launches, h->ev));
        }
        CUDA_TRY(cudaEventRecord(h->ev[4], s0));
        CUDA_TRY(cudaMemcpyAsync(h->wc_host, h->work_counters.p, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost, s0));
        h->async_pending = true;
        h->async_launches = launches;
        if (v.async && mode != Mode::Match) return EPP_OK;         // epp_synchronize() (or the next call that needs the
                                                                   // results on the host) completes it
        return finish_async(h);
    }

    // ---- small host batch in pinned memory: one zero-copy launch, no copy engine, no stream synchronisation
    if (mode == Mode::Schedule && !tk && out_dec) {
        const uint8_t *data_dev = nullptr;
        int align = 0;
        const int sm = small_batch_mode(h, v, &data_dev, &align);
        if (sm) return run_small(h, v, sm, data_dev, align, out_dec, out_detail);
    }

    // ---- host batch: upload the small per-request arrays once, then pipeline the prompt bytes
    if (v.offsets) CUDA_TRY(cudaMemcpyAsync(h->offsets.p, v.offsets, sizeof(uint64_t) * (size_t)(R + 1), cudaMemcpyHostToDevice, s0));
    if (v.lengths) CUDA_TRY(cudaMemcpyAsync(h->lengths.p, v.lengths, sizeof(uint64_t) * (size_t)R, cudaMemcpyHostToDevice, s0));
    if (v.model_ids) CUDA_TRY(cudaMemcpyAsync(h->model_ids.p, v.model_ids, sizeof(uint32_t) * (size_t)R, cudaMemcpyHostToDevice, s0));
    if (v.multimodal) {
        CUDA_TRY(h->multimodal.reserve((size_t)R, &h->dev_bytes));
        CUDA_TRY(cudaMemcpyAsync(h->multimodal.p, v.multimodal, (size_t)R, cudaMemcpyHostToDevice, s0));
    }
    CUDA_TRY(cudaEventRecord(h->slot[0].done, s0));
    CUDA_TRY(cudaStreamWaitEvent(h->slot[1].stream, h->slot[0].done, 0));

    // chunk plan: at least one request per chunk, about kChunkBytes of prompt bytes
    const uint64_t kChunkBytes = 48ull << 20;
    struct Chunk { int64_t r0, r1; uint64_t start, bytes; };
    std::vector<Chunk> chunks;
    uint64_t max_bytes = 0;
    for (int64_t r = 0; r < R;) {
        int64_t r_end = r;
        uint64_t start, bytes;
        if (v.offsets) {
            while (r_end < R && (r_end == r || (v.offsets[r_end + 1] - v.offsets[r]) <= kChunkBytes)) r_end++;
            start = v.offsets[r];
            bytes = v.offsets[r_end] - start;
        } else {
            int64_t per = v.uniform_len ? (int64_t)std::max<uint64_t>(1, kChunkBytes / v.uniform_len) : R;
            r_end = std::min(R, r + per);
            start = (uint64_t)r * v.uniform_len;
            bytes = (uint64_t)(r_end - r) * v.uniform_len;
        }
        chunks.push_back({r, r_end, start, bytes});
        max_bytes = std::max(max_bytes, bytes);
        r = r_end;
    }
    for (int i = 0; i < 2; i++) CUDA_TRY(h->slot[i].data.reserve(max_bytes + 64, &h->dev_bytes));
    const size_t E = (size_t)h->cfg.max_endpoints;
    if (mode == Mode::Match) CUDA_TRY(h->dense_match.reserve(sizeof(int32_t) * (size_t)R * E, &h->dev_bytes));

    CUDA_TRY(cudaEventRecord(h->ev[0], s0));
    for (size_t k = 0; k < chunks.size(); k++) {
        const Chunk &c = chunks[k];
        Slot &sl = h->slot[k & 1];
        cudaStream_t s = sl.stream;
        const size_t nreq = (size_t)(c.r1 - c.r0);
        // keep the staged copy aligned (mod 32) like the caller's layout
        uint8_t *stage = sl.data.as<uint8_t>() + (c.start & 31);
        if (c.bytes) CUDA_TRY(cudaMemcpyAsync(stage, v.data + c.start, c.bytes, cudaMemcpyHostToDevice, s));
        Work w;
        w.r0 = c.r0;
        w.r1 = c.r1;
        if (v.offsets) {
            // the kernels address  data_base + offsets[r]  with ABSOLUTE offsets
            w.data_base = reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(stage) - (uintptr_t)c.start);
            w.offsets_dev = h->offsets.as<uint64_t>();
            w.lengths_dev = v.lengths ? h->lengths.as<uint64_t>() : nullptr;
            w.offsets_or_bits = v.offsets_or_bits;
        } else {
            w.data_base = stage;                    // chunk-local request index * uniform_len
            w.offsets_dev = nullptr;
            w.lengths_dev = nullptr;
            w.offsets_or_bits = 0;
        }
        w.uniform_len = v.uniform_len;
        w.model_ids_dev = v.model_ids ? h->model_ids.as<uint32_t>() : nullptr;
        w.multimodal_dev = v.multimodal ? h->multimodal.as<uint8_t>() : nullptr;
        w.hashes_out = h->hashes.as<uint64_t>() + (size_t)c.r0 * B;
        w.nblocks_out = h->nblocks.as<int32_t>() + c.r0;
        if (mode == Mode::HashOnly) {
            CUDA_TRY(launch_hash_prompts(hash_params(h, w), s, &launches));
            if (out_hashes) CUDA_TRY(cudaMemcpyAsync(out_hashes + (size_t)c.r0 * B, w.hashes_out, sizeof(uint64_t) * nreq * B, cudaMemcpyDeviceToHost, s));
            if (out_nblocks) CUDA_TRY(cudaMemcpyAsync(out_nblocks + c.r0, w.nblocks_out, sizeof(int32_t) * nreq, cudaMemcpyDeviceToHost, s));
        } else {
            int32_t *dm = mode == Mode::Match ? h->dense_match.as<int32_t>() + (size_t)c.r0 * E : nullptr;
            epp_decision *dec = h->decisions.as<epp_decision>() + c.r0;
            epp_decision_detail *det = h->details.as<epp_decision_detail>() + c.r0;
            PickParams pp = pick_params(h, w, dec, det, dm);
            if (tk) attach_topk(h, pp.gen, tk_picks, tk_scores, c.r0);
            EPP_TRY(launch_cycle(h, sl, hash_params(h, w), pp, &launches, nullptr));
            if (mode == Mode::Schedule) {
                if (out_dec) CUDA_TRY(cudaMemcpyAsync(out_dec + c.r0, dec, sizeof(epp_decision) * nreq, cudaMemcpyDeviceToHost, s));
                if (out_detail) CUDA_TRY(cudaMemcpyAsync(out_detail + c.r0, det, sizeof(epp_decision_detail) * nreq, cudaMemcpyDeviceToHost, s));
                if (tk) {
                    const size_t o = (size_t)c.r0 * K;
                    for (int i = 0; i < kMaxProfiles; i++)
                        if (tk_host[i]) CUDA_TRY(cudaMemcpyAsync(tk_host[i] + o, tk_picks[i] + o, sizeof(uint32_t) * nreq * K, cudaMemcpyDeviceToHost, s));
                    if (tk_scores) CUDA_TRY(cudaMemcpyAsync(tk->primary_scores + o, tk_scores + o, sizeof(double) * nreq * K, cudaMemcpyDeviceToHost, s));
                }
            } else {
                if (out_match) CUDA_TRY(cudaMemcpyAsync(out_match + (size_t)c.r0 * E, dm, sizeof(int32_t) * nreq * E, cudaMemcpyDeviceToHost, s));
                if (out_total) CUDA_TRY(cudaMemcpyAsync(out_total + c.r0, w.nblocks_out, sizeof(int32_t) * nreq, cudaMemcpyDeviceToHost, s));
            }
        }
    }
    CUDA_TRY(cudaStreamSynchronize(h->slot[1].stream));
    CUDA_TRY(cudaEventRecord(h->ev[1], s0));
    CUDA_TRY(cudaStreamSynchronize(s0));
    float t_all = 0;
    cudaEventElapsedTime(&t_all, h->ev[0], h->ev[1]);
    h->stats.last_kernels_ms = t_all;
    h->stats.last_hash_ms = h->stats.last_match_pick_ms = 0;
    h->stats.last_kernel_launches = (uint64_t)launches;
    return EPP_OK;
}

extern "C" int32_t epp_hash_prompts(epp_engine *h, const epp_batch *batch, uint64_t *out_hashes, int32_t *out_nblocks) {
    if (!h) return fail(EPP_ERR_INVALID, "NULL engine");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    BatchView v;
    EPP_TRY(check_batch(h, batch, v));
    v.async = false;                               // EPP_BATCH_ASYNC is an epp_schedule flag: hashes are complete on return
    h->kept_R = 0;
    return run_batch(h, v, Mode::HashOnly, out_hashes, out_nblocks, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int32_t epp_prefix_match(epp_engine *h, const epp_batch *batch, int32_t *out_match, int32_t *out_total) {
    if (!h) return fail(EPP_ERR_INVALID, "NULL engine");
    if (!out_match) return fail(EPP_ERR_INVALID, "out_match is NULL");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    BatchView v;
    EPP_TRY(check_batch(h, batch, v));
    h->kept_R = 0;
    return run_batch(h, v, Mode::Match, nullptr, nullptr, nullptr, nullptr, out_match, out_total);
}

static int32_t check_topk(epp_engine *h, const epp_topk_out *topk) {
    if (!topk) return fail(EPP_ERR_INVALID, "topk is NULL");
    if (topk->struct_size != sizeof(epp_topk_out)) return fail(EPP_ERR_INVALID, "epp_topk_out.struct_size %u != %zu", topk->struct_size, sizeof(epp_topk_out));
    if (h->cfg.pick_k <= 1) return fail(EPP_ERR_STATE, "the engine was created with pick_k = %d: nothing but the pick to return", h->cfg.pick_k);
    if (topk->k != h->cfg.pick_k) return fail(EPP_ERR_INVALID, "epp_topk_out.k %d != epp_config.pick_k %d", topk->k, h->cfg.pick_k);
    return EPP_OK;
}

static int32_t schedule_impl(epp_engine *h, const epp_batch *batch, epp_decision *out, epp_decision_detail *detail,
                             int32_t keep_hashes, const epp_topk_out *topk) {
    if (!h) return fail(EPP_ERR_INVALID, "NULL engine");
    if (!out) return fail(EPP_ERR_INVALID, "out is NULL");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    BatchView v;
    EPP_TRY(check_batch(h, batch, v));
    if (topk) EPP_TRY(check_topk(h, topk));
    if (h->cfg.encode_enabled && v.multimodal && !detail) return fail(EPP_ERR_INVALID, "an encode profile is configured and the batch flags multimodal requests: `detail` (which carries the encode pick) must not be NULL");
    h->kept_R = 0;
    h->cur_topk = topk;
    const int32_t rc = run_batch(h, v, Mode::Schedule, nullptr, nullptr, out, detail, nullptr, nullptr);
    h->cur_topk = nullptr;
    EPP_TRY(rc);
    h->stats.n_batches++;
    h->stats.n_decisions += (uint64_t)v.R;
    if (keep_hashes && v.R) {
        // the decisions stay on the device next to the hashes; host batches already have them in h->decisions
        cudaStream_t s0 = h->slot[0].stream;
        EPP_TRY(join_streams(h));
        CUDA_TRY(h->kept_dec.reserve(sizeof(epp_decision) * (size_t)v.R, &h->dev_bytes));
        const void *src = v.device ? (const void *)out : (const void *)h->decisions.p;
        CUDA_TRY(cudaMemcpyAsync(h->kept_dec.p, src, sizeof(epp_decision) * (size_t)v.R, cudaMemcpyDeviceToDevice, s0));
        if (v.device && !v.async) CUDA_TRY(cudaStreamSynchronize(s0));     // `out` is the caller's again when we return
        h->kept_R = v.R;
    }
    return EPP_OK;
}

extern "C" int32_t epp_schedule(epp_engine *h, const epp_batch *batch, epp_decision *out, epp_decision_detail *detail,
                                int32_t keep_hashes) {
    return schedule_impl(h, batch, out, detail, keep_hashes, nullptr);
}

extern "C" int32_t epp_schedule_topk(epp_engine *h, const epp_batch *batch, epp_decision *out, epp_decision_detail *detail,
                                     int32_t keep_hashes, const epp_topk_out *topk) {
    if (!topk) return fail(EPP_ERR_INVALID, "topk is NULL");
    return schedule_impl(h, batch, out, detail, keep_hashes, topk);
}

// PreRequest: approximateprefix/plugin.go:164-200 (primary target + "prefill" profile target).
extern "C" int32_t epp_index_add_picked(epp_engine *h) {
    if (!h) return fail(EPP_ERR_INVALID, "NULL engine");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    if (h->kept_R == 0) return fail(EPP_ERR_STATE, "no batch kept (call epp_schedule with keep_hashes=1 first)");
    if (h->snapshot_mode) return fail(EPP_ERR_STATE, "index holds a bulk snapshot; incremental adds need an incrementally built index");
    EPP_TRY(flush_add_queue(h));
    auto t0 = std::chrono::steady_clock::now();
    CUDA_TRY(h->store->apply_picks(h->kept_dec.as<epp_decision>(), h->hashes.as<uint64_t>(), h->nblocks.as<int32_t>(), h->kept_R,
                                   h->cfg.max_prefix_blocks, h->slot[0].stream));
    h->stats.last_index_apply_ms = ms_since(t0);
    h->stats.last_index_items = h->store->last_items;
    h->stats.last_index_launches = (uint64_t)h->store->last_launches;
    h->kept_R = 0;
    return EPP_OK;
}

// ------------------------------------------------------------------------------------------------
// plugin-parity entry points on injected match info
// ------------------------------------------------------------------------------------------------
extern "C" int32_t epp_score(epp_engine *h, int64_t n_requests, const int32_t *match, const int32_t *total,
                             const uint32_t *model_ids, int32_t profile, int32_t scorer_index, double *out_scores,
                             uint32_t flags) {
    if (!h || n_requests < 0 || (n_requests && (!match || !total || !out_scores))) return fail(EPP_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    if (!h->pool_ready) return fail(EPP_ERR_STATE, "epp_pool_set has not been called");
    if (profile < 0 || profile >= h->n_alloc_profiles) return fail(EPP_ERR_INVALID, "profile %d out of range", profile);
    const epp_profile_cfg &pc = profile == 0 ? h->cfg.primary : (profile == 1 ? h->cfg.prefill : h->cfg.encode);
    if (scorer_index < -1 || scorer_index >= pc.n_scorers) return fail(EPP_ERR_INVALID, "scorer_index %d out of range", scorer_index);
    if (n_requests == 0) return EPP_OK;
    const size_t E = (size_t)h->cfg.max_endpoints, n = (size_t)n_requests * E;
    cudaStream_t s = h->slot[0].stream;
    const bool dev = (flags & EPP_BATCH_DEVICE_PTRS) != 0;
    const int32_t *d_match = match, *d_total = total;
    double *d_out = out_scores;
    if (!dev) {
        CUDA_TRY(h->dense_match.reserve(sizeof(int32_t) * n, &h->dev_bytes));
        CUDA_TRY(h->dense_total.reserve(sizeof(int32_t) * (size_t)n_requests, &h->dev_bytes));
        CUDA_TRY(h->dense_scores.reserve(sizeof(double) * n, &h->dev_bytes));
        CUDA_TRY(cudaMemcpyAsync(h->dense_match.p, match, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(h->dense_total.p, total, sizeof(int32_t) * (size_t)n_requests, cudaMemcpyHostToDevice, s));
        d_match = h->dense_match.as<int32_t>();
        d_total = h->dense_total.as<int32_t>();
        d_out = h->dense_scores.as<double>();
    }
    const uint32_t *d_models = model_ids;
    if (!dev && model_ids) {
        CUDA_TRY(h->score_models.reserve(sizeof(uint32_t) * (size_t)n_requests, &h->dev_bytes));
        CUDA_TRY(cudaMemcpyAsync(h->score_models.p, model_ids, sizeof(uint32_t) * (size_t)n_requests, cudaMemcpyHostToDevice, s));
        d_models = h->score_models.as<uint32_t>();
    }
    int launches = 0;
    CUDA_TRY(launch_score_dense(n_requests, h->cfg.max_endpoints, profile_dev(h, profile), pool_arrays(h),
                                h->prof[profile].qminmax.as<int64_t>(), d_match, d_total, d_models, scorer_index, d_out, s, &launches));
    if (!dev) CUDA_TRY(cudaMemcpyAsync(out_scores, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return EPP_OK;
}

static int32_t with_match_impl(epp_engine *h, int64_t n_requests, const int32_t *match, const int32_t *total,
                               const uint32_t *model_ids, const int64_t *input_len_bytes, int32_t block_size_tokens, epp_decision *out,
                               epp_decision_detail *detail, uint32_t flags, const epp_topk_out *topk) {
    if (!h || n_requests < 0 || (n_requests && (!match || !total || !out))) return fail(EPP_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    if (topk) EPP_TRY(check_topk(h, topk));
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    if (!h->pool_ready) return fail(EPP_ERR_STATE, "epp_pool_set has not been called");
    if (n_requests == 0) return EPP_OK;
    const size_t E = (size_t)h->cfg.max_endpoints, n = (size_t)n_requests * E, R = (size_t)n_requests;
    cudaStream_t s = h->slot[0].stream;
    const bool dev = (flags & EPP_BATCH_DEVICE_PTRS) != 0;
    DensePickParams p;
    p.R = n_requests;
    p.E = h->cfg.max_endpoints;
    p.block_size_tokens = block_size_tokens > 0 ? block_size_tokens : h->cfg.block_size_tokens;
    p.n_profiles = h->n_profiles;
    p.encode_on = 0;                              // the injected-match entry point has no multimodal flags
    p.multimodal = nullptr;
    p.always_disagg = h->cfg.always_disagg;
    p.non_cached_tokens = h->cfg.non_cached_tokens;
    for (int i = 0; i < h->n_alloc_profiles; i++) p.prof[i] = profile_dev(h, i);
    p.lora = lora_dev(h);
    p.model_ids = model_ids;
    p.tie_seed = h->cfg.tie_seed;
    p.tie_base = h->stats.n_decisions;
    p.gen = gen_view(h);
    const size_t K = (size_t)h->cfg.pick_k;
    uint32_t *tk_host[kMaxProfiles] = {nullptr, nullptr, nullptr};
    uint32_t *tk_picks[kMaxProfiles] = {nullptr, nullptr, nullptr};
    double *tk_scores = nullptr;
    if (topk) {
        tk_host[0] = topk->primary; tk_host[1] = topk->prefill; tk_host[2] = topk->encode;
        for (int i = 0; i < kMaxProfiles; i++) {
            tk_picks[i] = tk_host[i];
            if (!dev && tk_host[i]) {
                CUDA_TRY(h->topk_picks[i].reserve(sizeof(uint32_t) * R * K, &h->dev_bytes));
                tk_picks[i] = h->topk_picks[i].as<uint32_t>();
            }
        }
        tk_scores = topk->primary_scores;
        if (!dev && tk_scores) {
            CUDA_TRY(h->topk_scores.reserve(sizeof(double) * R * K, &h->dev_bytes));
            tk_scores = h->topk_scores.as<double>();
        }
        attach_topk(h, p.gen, tk_picks, tk_scores, 0);
    }
    if (dev) {
        p.match = match; p.total = total; p.in_len = input_len_bytes; p.out = out; p.detail = detail;
    } else {
        if (model_ids) {
            CUDA_TRY(h->score_models.reserve(sizeof(uint32_t) * R, &h->dev_bytes));
            CUDA_TRY(cudaMemcpyAsync(h->score_models.p, model_ids, sizeof(uint32_t) * R, cudaMemcpyHostToDevice, s));
            p.model_ids = h->score_models.as<uint32_t>();
        }
        EPP_TRY(reserve_batch(h, n_requests));
        CUDA_TRY(h->dense_match.reserve(sizeof(int32_t) * n, &h->dev_bytes));
        CUDA_TRY(h->dense_total.reserve(sizeof(int32_t) * R, &h->dev_bytes));
        CUDA_TRY(cudaMemcpyAsync(h->dense_match.p, match, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(h->dense_total.p, total, sizeof(int32_t) * R, cudaMemcpyHostToDevice, s));
        if (input_len_bytes) CUDA_TRY(cudaMemcpyAsync(h->in_len.p, input_len_bytes, sizeof(int64_t) * R, cudaMemcpyHostToDevice, s));
        else CUDA_TRY(cudaMemsetAsync(h->in_len.p, 0, sizeof(int64_t) * R, s));
        p.match = h->dense_match.as<int32_t>();
        p.total = h->dense_total.as<int32_t>();
        p.in_len = h->in_len.as<int64_t>();
        p.out = h->decisions.as<epp_decision>();
        p.detail = h->details.as<epp_decision_detail>();
    }
    int launches = 0;
    CUDA_TRY(launch_dense_pick(p, s, &launches));
    if (!dev) {
        CUDA_TRY(cudaMemcpyAsync(out, p.out, sizeof(epp_decision) * R, cudaMemcpyDeviceToHost, s));
        if (detail) CUDA_TRY(cudaMemcpyAsync(detail, p.detail, sizeof(epp_decision_detail) * R, cudaMemcpyDeviceToHost, s));
        for (int i = 0; i < kMaxProfiles; i++)
            if (tk_host[i]) CUDA_TRY(cudaMemcpyAsync(tk_host[i], tk_picks[i], sizeof(uint32_t) * R * K, cudaMemcpyDeviceToHost, s));
        if (topk && topk->primary_scores) CUDA_TRY(cudaMemcpyAsync(topk->primary_scores, tk_scores, sizeof(double) * R * K, cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    h->kept_R = 0;
    h->stats.n_decisions += (uint64_t)n_requests;          // request ordinals (tie rule) advance here too
    return EPP_OK;
}

extern "C" int32_t epp_schedule_with_match(epp_engine *h, int64_t n_requests, const int32_t *match, const int32_t *total,
                                           const uint32_t *model_ids, const int64_t *input_len_bytes, int32_t block_size_tokens, epp_decision *out,
                                           epp_decision_detail *detail, uint32_t flags) {
    return with_match_impl(h, n_requests, match, total, model_ids, input_len_bytes, block_size_tokens, out, detail, flags, nullptr);
}

extern "C" int32_t epp_schedule_with_match_topk(epp_engine *h, int64_t n_requests, const int32_t *match, const int32_t *total,
                                                const uint32_t *model_ids, const int64_t *input_len_bytes, int32_t block_size_tokens,
                                                epp_decision *out, epp_decision_detail *detail, uint32_t flags,
                                                const epp_topk_out *topk) {
    if (!topk) return fail(EPP_ERR_INVALID, "topk is NULL");
    return with_match_impl(h, n_requests, match, total, model_ids, input_len_bytes, block_size_tokens, out, detail, flags, topk);
}

extern "C" int32_t epp_get_config(epp_engine *h, epp_config *out) {
    if (!h || !out) return fail(EPP_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(h->mu);
    *out = h->cfg;
    return EPP_OK;
}

extern "C" int32_t epp_get_stats(epp_engine *h, epp_stats *out) {
    if (!h || !out) return fail(EPP_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->async_pending) { EPP_TRY(set_device(h)); EPP_TRY(finish_async(h)); }
    if (h->small_stats_pending) {
        EPP_TRY(set_device(h));
        float t = 0;
        CUDA_TRY(cudaEventSynchronize(h->ev[1]));
        cudaEventElapsedTime(&t, h->ev[0], h->ev[1]);
        h->stats.last_kernels_ms = t;
        for (int i = 0; i < 8; i++) h->stats.last_kernel_ms[i] = i == 0 ? t : 0.0;
        h->small_stats_pending = false;
    }
    h->stats.device_bytes = h->dev_bytes + h->store->device_bytes();
    *out = h->stats;
    return EPP_OK;
}

extern "C" int32_t epp_synchronize(epp_engine *h) {
    if (!h) return fail(EPP_ERR_INVALID, "NULL engine");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    EPP_TRY(finish_async(h));
    CUDA_TRY(cudaStreamSynchronize(h->slot[1].stream));
    return EPP_OK;
}

extern "C" int32_t epp_event_record(epp_engine *h, int32_t which) {
    if (!h || which < 0 || which > 1) return fail(EPP_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    CUDA_TRY(cudaEventRecord(h->user_ev[which], h->slot[0].stream));
    return EPP_OK;
}

extern "C" int32_t epp_event_elapsed_ms(epp_engine *h, double *out_ms) {
    if (!h || !out_ms) return fail(EPP_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    CUDA_TRY(cudaEventSynchronize(h->user_ev[1]));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->user_ev[0], h->user_ev[1]));
    *out_ms = ms;
    return EPP_OK;
}

// ------------------------------------------------------------------------------------------------
// endpoint-sharded mode (kernels: pick_kernels.cu k_shard_*, shard_p2p.cu)
// ------------------------------------------------------------------------------------------------
extern "C" int32_t epp_shard_set(epp_engine *h, uint32_t ep_begin, uint32_t ep_end) {
    if (!h || ep_begin > ep_end) return fail(EPP_ERR_INVALID, "bad shard range");
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->cfg.handler != EPP_HANDLER_SINGLE) return fail(EPP_ERR_INVALID, "endpoint-sharded mode supports the single-profile handler only");
    if (h->general) return fail(EPP_ERR_INVALID, "endpoint-sharded mode supports neither pick_k > 1 nor the prefix-cache-affinity-filter (both need every candidate's score on one rank)");
    h->shard_begin = ep_begin;
    h->shard_end = ep_end;
    h->pool_ready = false;          // candidates are derived per shard: epp_pool_set must follow
    return EPP_OK;
}

static int32_t mask_words_of(const epp_engine *h) { return (h->cfg.max_prefix_blocks + 31) / 32; }

extern "C" int32_t epp_shard_probe(epp_engine *h, const epp_batch *batch, uint32_t *out_masks) {
    if (!h || !out_masks) return fail(EPP_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    BatchView v;
    EPP_TRY(check_batch(h, batch, v));
    if (!v.device) return fail(EPP_ERR_INVALID, "epp_shard_probe takes device-pointer batches (EPP_BATCH_DEVICE_PTRS)");
    h->kept_R = 0;
    h->shard_R = 0;
    EPP_TRY(commit_locked(h));
    EPP_TRY(run_batch(h, v, Mode::HashOnly, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
    int launches = 0;
    cudaStream_t s = h->slot[0].stream;
    CUDA_TRY(launch_shard_probe(v.R, h->cfg.max_prefix_blocks, h->hashes.as<uint64_t>(), h->nblocks.as<int32_t>(),
                                index_view(h), out_masks, mask_words_of(h), s, &launches));
    CUDA_TRY(cudaStreamSynchronize(s));
    h->shard_R = v.R;
    return EPP_OK;
}

extern "C" int32_t epp_shard_pick(epp_engine *h, int64_t n_requests, const uint32_t *global_masks, epp_shard_best *out_best) {
    if (!h || !global_masks || !out_best) return fail(EPP_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    if (!h->pool_ready) return fail(EPP_ERR_STATE, "epp_pool_set has not been called (after epp_shard_set)");
    if (n_requests != h->shard_R) return fail(EPP_ERR_STATE, "epp_shard_pick: n_requests %lld does not match the probed batch (%lld)", (long long)n_requests, (long long)h->shard_R);
    if (n_requests == 0) return EPP_OK;
    Work w{0, n_requests, nullptr, nullptr, nullptr, 0, nullptr, 0, h->hashes.as<uint64_t>(), h->nblocks.as<int32_t>()};
    PickParams pp = pick_params(h, w, h->decisions.as<epp_decision>(), nullptr, nullptr);
    pp.global_masks = global_masks;
    pp.mask_words = mask_words_of(h);
    pp.shard_out = out_best;
    int launches = 0;
    cudaStream_t s = h->slot[0].stream;
    EPP_TRY(reserve_batch(h, n_requests));
    EPP_TRY(launch_match(h, h->slot[0], pp, &launches));
    CUDA_TRY(cudaStreamSynchronize(s));
    return EPP_OK;
}

extern "C" int32_t epp_shard_merge(epp_engine *h, int64_t n_requests, int32_t n_ranks, const epp_shard_best *all_best, epp_decision *out) {
    if (!h || !all_best || !out || n_ranks <= 0) return fail(EPP_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    if (n_requests != h->shard_R) return fail(EPP_ERR_STATE, "epp_shard_merge: n_requests does not match the probed batch");
    int launches = 0;
    cudaStream_t s = h->slot[0].stream;
    CUDA_TRY(launch_shard_merge(n_requests, n_ranks, all_best, h->nblocks.as<int32_t>(), out, s, &launches));
    CUDA_TRY(cudaStreamSynchronize(s));
    return EPP_OK;
}

// ---- endpoint-sharded mode with the exchanges over NVLink peer memory (shard_p2p.cu) ------------------------------
// flags of chunk c of a batch: flag 1 at c * 256, flag 2 at c * 256 + 128 (two chunks: the halves of a batch travel through
// the exchange on the engine's two streams, so that a rank computes on one half while it waits for its peers on the other)
static constexpr size_t kP2pFlag1 = 0, kP2pFlag2 = 128, kP2pChunkFlags = 256, kP2pChunks = 2, kP2pHeader = kP2pChunkFlags * kP2pChunks;

extern "C" int32_t epp_shard_p2p_export(epp_engine *h, int64_t max_requests, uint8_t *out_handle, uint64_t *out_ptr) {
    if (!h || max_requests <= 0 || !out_handle || !out_ptr) return fail(EPP_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    if (h->p2p_buf && !h->p2p_broken) return fail(EPP_ERR_STATE, "the exchange buffer is already exported");
    if (h->p2p_buf) {                                           // broken exchange: start over with a fresh buffer
        for (int g = 0; g < (int)h->p2p_peer.size(); g++)
            if (h->p2p_ipc && g != h->p2p_rank && h->p2p_peer[g]) cudaIpcCloseMemHandle(h->p2p_peer[g]);
        h->p2p_peer.clear();
        cudaFree(h->p2p_buf);
        h->dev_bytes -= h->p2p_bytes;
        h->p2p_buf = nullptr;
        h->p2p_broken = false;
        h->p2p_epoch = 0;
    }
    const size_t W = (size_t)mask_words_of(h);
    h->p2p_masks_off = kP2pHeader;
    h->p2p_best_off = (kP2pHeader + (size_t)max_requests * W * sizeof(uint32_t) + 255) & ~(size_t)255;
    h->p2p_bytes = (h->p2p_best_off + (size_t)max_requests * sizeof(epp_shard_best) + 255) & ~(size_t)255;
    CUDA_TRY(cudaMalloc(&h->p2p_buf, h->p2p_bytes));
    CUDA_TRY(cudaMemset(h->p2p_buf, 0, h->p2p_bytes));
    h->dev_bytes += h->p2p_bytes;
    h->p2p_R = max_requests;
    cudaIpcMemHandle_t hd;
    CUDA_TRY(cudaIpcGetMemHandle(&hd, h->p2p_buf));
    static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(out_handle, &hd, sizeof hd);
    *out_ptr = (uint64_t)(uintptr_t)h->p2p_buf;
    return EPP_OK;
}

extern "C" int32_t epp_shard_p2p_connect(epp_engine *h, int32_t n_ranks, int32_t rank, const void *peers, int32_t ipc_handles) {
    if (!h || n_ranks <= 0 || n_ranks > 64 || rank < 0 || rank >= n_ranks || !peers) return fail(EPP_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    if (!h->p2p_buf) return fail(EPP_ERR_STATE, "epp_shard_p2p_export has not been called");
    if (!h->p2p_peer.empty()) return fail(EPP_ERR_STATE, "peers are already connected");
    h->p2p_peer.assign((size_t)n_ranks, nullptr);
    h->p2p_ipc = ipc_handles != 0;
    for (int g = 0; g < n_ranks; g++) {
        if (g == rank) { h->p2p_peer[g] = h->p2p_buf; continue; }
        if (ipc_handles) {
            cudaIpcMemHandle_t hd;
            memcpy(&hd, static_cast<const uint8_t *>(peers) + (size_t)g * 64, sizeof hd);
            CUDA_TRY(cudaIpcOpenMemHandle(&h->p2p_peer[g], hd, cudaIpcMemLazyEnablePeerAccess));
        } else {
            h->p2p_peer[g] = (void *)(uintptr_t)static_cast<const uint64_t *>(peers)[g];    // same process: plain pointers
        }
    }
    h->p2p_ranks = n_ranks;
    h->p2p_rank = rank;
    CUDA_TRY(h->p2p_peer_dev.reserve(sizeof(void *) * (size_t)n_ranks, &h->dev_bytes));
    CUDA_TRY(cudaMemcpy(h->p2p_peer_dev.p, h->p2p_peer.data(), sizeof(void *) * (size_t)n_ranks, cudaMemcpyHostToDevice));
    CUDA_TRY(h->p2p_err.reserve(sizeof(int), &h->dev_bytes));
    CUDA_TRY(cudaMemset(h->p2p_err.p, 0, sizeof(int)));
    // everything the batches need is allocated now: no cudaMalloc / cudaFree (device-wide synchronisation) may happen
    // while a peer-wait kernel is spinning
    const size_t W = (size_t)mask_words_of(h);
    CUDA_TRY(h->p2p_gmasks.reserve(((size_t)h->p2p_R * W * sizeof(uint32_t) + 15) & ~(size_t)15, &h->dev_bytes));
    CUDA_TRY(h->p2p_allbest.reserve(sizeof(epp_shard_best) * (size_t)h->p2p_R * (size_t)n_ranks, &h->dev_bytes));
    EPP_TRY(reserve_batch(h, h->p2p_R));
    return EPP_OK;
}

// One batch of the sharded protocol with both exchanges done by this rank's own kernels over peer memory.  Every rank
// calls it with the same batch; the decisions (identical on every rank) land in `out` (device pointer).
// The protocol is three phases; epp_shard_schedule_p2p runs them back to back (the waits happen ON THE DEVICE, the host
// only waits at the end), epp_shard_p2p_phase runs ONE phase and synchronises, so that ranks which share a GPU -- whose
// spin-wait kernels could starve each other -- can be stepped phase by phase from one host thread (tests, small setups).
static constexpr unsigned long long kP2pPoison = ~0ull;

static unsigned long long p2p_timeout_ns() {
    unsigned long long timeout_ns = 20ull * 1000 * 1000 * 1000;
    if (const char *tv = getenv("EPP_P2P_TIMEOUT_MS")) timeout_ns = (unsigned long long)std::max(1, atoi(tv)) * 1000ull * 1000ull;
    return timeout_ns;
}

// A wait that timed out (or met a poisoned flag) leaves the ranks out of step for good: the exchange is marked broken,
// this rank's flags are poisoned so that the peers fail too instead of reading half-written buffers, and every later
// call fails until the buffers are exported / connected again.
static int32_t p2p_check(epp_engine *h, cudaStream_t s, unsigned long long epoch) {
    int err = 0;
    CUDA_TRY(cudaMemcpyAsync(&err, h->p2p_err.p, sizeof err, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (!err) return EPP_OK;
    h->p2p_broken = true;
    uint8_t *own = static_cast<uint8_t *>(h->p2p_buf);
    for (size_t c = 0; c < kP2pChunks; c++) {
        CUDA_TRY(launch_p2p_signal(reinterpret_cast<unsigned long long *>(own + c * kP2pChunkFlags + kP2pFlag1), kP2pPoison, s));
        CUDA_TRY(launch_p2p_signal(reinterpret_cast<unsigned long long *>(own + c * kP2pChunkFlags + kP2pFlag2), kP2pPoison, s));
    }
    CUDA_TRY(cudaMemsetAsync(h->p2p_err.p, 0, sizeof(int), s));
    CUDA_TRY(cudaStreamSynchronize(s));
    const bool poisoned = err > 64;
    return fail(EPP_ERR_NCCL, poisoned ? "peer rank %d reported a broken sharded exchange (batch %llu); re-export and re-connect the exchange buffers"
                                       : "peer rank %d did not reach batch %llu of the sharded exchange within %llu ms; the exchange is broken until the buffers are exported and connected again",
                (poisoned ? err - 64 : err) - 1, epoch, p2p_timeout_ns() / 1000000ull);
}

// One phase of the exchange for rows [r0, r1) of the batch (chunk c: its own flag words, stream of slot c).  Phase 0
// expects ++p2p_epoch and the commit to have happened; the caller checks p2p_err after phase 2 of every chunk.
static int32_t p2p_phase_rows(epp_engine *h, const BatchView &v, epp_decision *out, int phase, int c, int64_t r0, int64_t r1,
                              uint64_t or_bits) {
    const int64_t R = r1 - r0;
    const int n = h->p2p_ranks;
    const size_t W = (size_t)mask_words_of(h), B = (size_t)h->cfg.max_prefix_blocks;
    Slot &sl = h->slot[c];
    cudaStream_t s = sl.stream;
    uint8_t *own = static_cast<uint8_t *>(h->p2p_buf);
    unsigned char *const *peers = h->p2p_peer_dev.as<unsigned char *>();
    const size_t flag1 = (size_t)c * kP2pChunkFlags + kP2pFlag1, flag2 = (size_t)c * kP2pChunkFlags + kP2pFlag2;
    const size_t mask_off = h->p2p_masks_off + (size_t)r0 * W * sizeof(uint32_t);
    const size_t mask_bytes = ((size_t)R * W * sizeof(uint32_t) + 15) & ~(size_t)15;
    const size_t best_off = h->p2p_best_off + (size_t)r0 * sizeof(epp_shard_best);
    const unsigned long long timeout_ns = p2p_timeout_ns();
    const unsigned long long epoch = h->p2p_epoch;
    uint64_t *hashes = h->hashes.as<uint64_t>() + (size_t)r0 * B;
    int32_t *nbs = h->nblocks.as<int32_t>() + r0;
    uint32_t *gmasks = h->p2p_gmasks.as<uint32_t>() + (size_t)r0 * W;
    epp_shard_best *allbest = h->p2p_allbest.as<epp_shard_best>() + (size_t)r0 * (size_t)n;   // this chunk's own [n][R] matrix
    int launches = 0;
    if (phase == 0) {
        // hashes + local presence masks -> own exchange buffer, raise flag 1
        Work w{r0, r1, v.offsets ? v.data : v.data + (uint64_t)r0 * v.uniform_len, v.offsets, v.lengths, v.uniform_len, v.model_ids,
               or_bits, hashes, nbs, nullptr};
        CUDA_TRY(launch_hash_prompts(hash_params(h, w), s, &launches));
        CUDA_TRY(launch_shard_probe(R, h->cfg.max_prefix_blocks, hashes, nbs, index_view(h),
                                    reinterpret_cast<uint32_t *>(own + mask_off), (int32_t)W, s, &launches));
        CUDA_TRY(launch_p2p_signal(reinterpret_cast<unsigned long long *>(own + flag1), epoch, s));
        h->p2p_launches += launches + 1;
        return EPP_OK;
    }
    if (phase == 1) {
        // exchange 1: OR of every rank's masks, read straight from the peers; then the global stop rule on the local counts ->
        // local best record in the own exchange buffer, raise flag 2
        CUDA_TRY(launch_p2p_wait(peers, n, flag1, epoch, h->p2p_err.as<int>(), timeout_ns, s));
        CUDA_TRY(launch_p2p_or_masks(peers, n, mask_off, mask_bytes, gmasks, s));
        Work w{r0, r1, nullptr, nullptr, nullptr, 0, nullptr, 0, hashes, nbs};
        PickParams pp = pick_params(h, w, h->decisions.as<epp_decision>() + r0, nullptr, nullptr);
        pp.model_ids = v.model_ids ? v.model_ids + r0 : nullptr;
        pp.global_masks = gmasks;
        pp.mask_words = (int32_t)W;
        pp.shard_out = reinterpret_cast<epp_shard_best *>(own + best_off);
        EPP_TRY(launch_match(h, sl, pp, &launches));
        // a rank whose wait failed must NOT publish records computed from an incomplete OR: the signal kernel checks
        CUDA_TRY(launch_p2p_signal_unless(reinterpret_cast<unsigned long long *>(own + flag2), epoch, h->p2p_err.as<int>(), s));
        h->p2p_launches += launches + 3;
        return EPP_OK;
    }
    // exchange 2: gather every rank's records from the peers and merge
    CUDA_TRY(launch_p2p_wait(peers, n, flag2, epoch, h->p2p_err.as<int>(), timeout_ns, s));
    CUDA_TRY(launch_p2p_gather(peers, n, best_off, sizeof(epp_shard_best) * (size_t)R, allbest, s));
    CUDA_TRY(launch_shard_merge(R, n, allbest, nbs, out + r0, s, &launches));
    h->p2p_launches += launches + 2;
    return EPP_OK;
}

// Start / end of a batch of the exchange.
static int32_t p2p_begin(epp_engine *h, const BatchView &v, uint64_t *or_bits) {
    h->kept_R = 0;
    h->shard_R = 0;
    EPP_TRY(commit_locked(h));
    if (h->async_pending) EPP_TRY(finish_async(h));
    EPP_TRY(join_streams(h));
    EPP_TRY(reserve_batch(h, v.R));
    *or_bits = 0;
    if (v.offsets) EPP_TRY(device_offsets_or_bits(h, v.offsets, v.R, h->slot[0].stream, or_bits));
    ++h->p2p_epoch;
    h->p2p_launches = 0;
    return EPP_OK;
}
static int32_t p2p_end(epp_engine *h, int64_t R) {
    h->stats.last_kernel_launches = (uint64_t)h->p2p_launches;
    EPP_TRY(p2p_check(h, h->slot[0].stream, h->p2p_epoch));
    h->stats.n_batches++;
    h->stats.n_decisions += (uint64_t)R;
    return EPP_OK;
}
static int32_t p2p_prepare(epp_engine *h, const epp_batch *batch, epp_decision *out, BatchView &v) {
    if (!h || !out) return fail(EPP_ERR_INVALID, "NULL argument");
    EPP_TRY(set_device(h));
    EPP_TRY(join_streams(h));
    if (h->p2p_peer.empty()) return fail(EPP_ERR_STATE, "epp_shard_p2p_connect has not been called");
    if (h->p2p_broken) return fail(EPP_ERR_STATE, "the sharded exchange is broken (an earlier batch timed out): export and connect the buffers again");
    if (h->shard_begin == 0 && h->shard_end == 0xFFFFFFFFu) return fail(EPP_ERR_STATE, "epp_shard_set has not been called: every rank would count every endpoint");
    if (!h->pool_ready) return fail(EPP_ERR_STATE, "epp_pool_set has not been called (after epp_shard_set)");
    EPP_TRY(check_batch(h, batch, v));
    if (!v.device) return fail(EPP_ERR_INVALID, "the sharded peer-memory entry points take device-pointer batches (EPP_BATCH_DEVICE_PTRS)");
    if (v.R > h->p2p_R) return fail(EPP_ERR_CAPACITY, "batch of %lld requests exceeds the exported exchange buffer (%lld)", (long long)v.R, (long long)h->p2p_R);
    v.async = false;
    return EPP_OK;
}

extern "C" int32_t epp_shard_schedule_p2p(epp_engine *h, const epp_batch *batch, epp_decision *out) {
    if (!h) return fail(EPP_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(h->mu);
    BatchView v;
    EPP_TRY(p2p_prepare(h, batch, out, v));
    if (v.R == 0) return EPP_OK;
    uint64_t or_bits = 0;
    EPP_TRY(p2p_begin(h, v, &or_bits));
    // two halves on the engine's two streams: while one half waits for the peers' flags, the other half computes
    const bool split = h->dev_chunks > 1 && v.R >= 2 * 4096;       // (the dense-counter scratch is per stream)
    const int64_t half = split ? (((v.R + 1) / 2 + 31) & ~(int64_t)31) : v.R;
    if (split) {
        CUDA_TRY(cudaEventRecord(h->slot[0].done, h->slot[0].stream));
        CUDA_TRY(cudaStreamWaitEvent(h->slot[1].stream, h->slot[0].done, 0));
    }
    for (int phase = 0; phase < 3; phase++) {
        EPP_TRY(p2p_phase_rows(h, v, out, phase, 0, 0, half, or_bits));
        if (split) EPP_TRY(p2p_phase_rows(h, v, out, phase, 1, half, v.R, or_bits));
    }
    if (split) {
        h->s1_unjoined = true;
        EPP_TRY(join_streams(h));
    }
    return p2p_end(h, v.R);
}
extern "C" int32_t epp_shard_p2p_phase(epp_engine *h, const epp_batch *batch, epp_decision *out, int32_t phase) {
    if (!h) return fail(EPP_ERR_INVALID, "NULL argument");
    if (phase < 0 || phase > 2) return fail(EPP_ERR_INVALID, "phase %d out of range [0,2]", phase);
    std::lock_guard<std::mutex> lk(h->mu);
    BatchView v;
    EPP_TRY(p2p_prepare(h, batch, out, v));
    if (v.R == 0) return EPP_OK;
    if (phase == 0) EPP_TRY(p2p_begin(h, v, &h->p2p_or_bits));
    EPP_TRY(p2p_phase_rows(h, v, out, phase, 0, 0, v.R, h->p2p_or_bits));
    if (phase == 2) return p2p_end(h, v.R);
    CUDA_TRY(cudaStreamSynchronize(h->slot[0].stream));
    return EPP_OK;
}
