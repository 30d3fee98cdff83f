// kernels.h -- host-callable launchers of the sm_100a kernels (internal to libepp_engine.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/epp_engine.h"

namespace epp {

// ------------------------------------------------------------------------------------------------
// hashing (a1)
// ------------------------------------------------------------------------------------------------
struct HashParams {
    const uint8_t *data;        // prompt bytes (device)
    const uint64_t *offsets;    // [R+1] device, or nullptr => uniform_len
    const uint64_t *lengths;    // [R] device or nullptr => offsets[r+1] - offsets[r]
    uint64_t uniform_len;
    const uint32_t *model_ids;  // [R] device or nullptr
    const uint64_t *seeds;      // device table of h_{-1} per registered model
    int64_t R;
    int32_t block_bytes;        // blockSizeTokens * 4
    int32_t max_blocks;         // row pitch of `hashes`
    uint64_t *hashes;           // [R][max_blocks]
    int32_t *nblocks;           // [R]
    int64_t *eff_len;           // [R] bytes hashed after truncation (hashing.go:63-66)
    int64_t *in_len;            // [R] untruncated prompt length in bytes (P/D decider), may be nullptr
    uint64_t offsets_or_bits;   // OR of every offsets[r] (low bits decide the common alignment)
    int32_t sm_count;
    int32_t staged;             // hash_staged.cu (cp.async-staged, warp-per-task) kernel: 0 off, 1 default shape, else a launch shape (A/B)
};
// Common alignment (0, 16, 32) of every block start; >= 16 (and block_bytes % 32 == 0) enables the fused kernel.
int hash_batch_alignment(const HashParams &p);
// Generic single-message XXH64 (model || salt seeds).  msg on device.
cudaError_t launch_hash_bytes(const uint8_t *msg, size_t len, uint64_t *out, cudaStream_t s);
cudaError_t launch_check_offsets_aligned(const uint64_t *offsets, int64_t n, int *flag_dev, cudaStream_t s);
// Whole hashPrompt for a batch.  Returns number of kernels launched via *launches.
// ev (optional, 4 events): [1]..[2] brackets the kernel(s) that read the prompt bytes.
cudaError_t launch_hash_prompts(const HashParams &p, cudaStream_t s, int *launches, cudaEvent_t *ev = nullptr);

// hash_staged.cu: a1 with the prompt bytes staged through shared memory by cp.async, one warp per task of 16 / 32
// requests, no cross-warp synchronisation.  Needs 64-byte blocks and 16-byte aligned prompts.
bool hash_staged_supported(const HashParams &p);
cudaError_t launch_hash_staged(const HashParams &p, int shape, cudaStream_t s, int *launches);

// ------------------------------------------------------------------------------------------------
// prefix index (a2): open-addressed table  hash -> (posting offset, count)
// ------------------------------------------------------------------------------------------------
constexpr int kInlineIds = 5;
// One 32-byte slot = one DRAM/L2 sector, fetched by a single 256-bit load.  Posting lists of up to kInlineIds
// endpoints live IN the slot (no dependent load); longer lists live in `postings` at offset ids[0].  Lists are
// sorted by endpoint id and duplicate-free.
struct __align__(32) IndexSlot {
    uint64_t key;
    uint32_t cnt;                 // endpoints holding the block; 0 with a real key = every holder was evicted
                                  // (tombstone left by the incremental patch); a FREE slot has key == kEmptyKey
    uint32_t ids[kInlineIds];     // cnt <= kInlineIds: the endpoints; else ids[0] = offset into postings
};
constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFULL;   // free-slot sentinel (ends probe chains); a real key equal to it
                                                        // lives in IndexView::special
struct IndexView {
    const IndexSlot *slots;
    const uint32_t *postings;
    uint64_t mask;            // capacity - 1 (capacity is a power of two); slots == nullptr => empty index
    IndexSlot special;        // record for key == kEmptyKey (cnt 0 when absent)
    uint32_t ep_begin, ep_end;  // shard range (postings outside are ignored for counting but keep the walk alive)
};
// Builds table + postings from n pairs (device arrays).  slots must hold `capacity` entries, postings n,
// scratch `capacity` uint32, cursor 4 uint32 (cursor[2] returns the number of distinct hashes).  *special_dev
// receives the record of key == kEmptyKey.  Deduplicates pairs, sorts every posting list and interns equal spilled
// lists (intern_keys / intern_vals: `capacity` entries each).
cudaError_t launch_index_build(const uint64_t *pair_hash, const uint32_t *pair_ep, uint64_t n, IndexSlot *slots,
                               uint64_t capacity, uint32_t *postings, uint32_t *scratch, uint32_t *cursor,
                               IndexSlot *special_dev, uint64_t *intern_keys, uint32_t *intern_vals,
                               cudaStream_t s, int *launches);
cudaError_t launch_index_get(const IndexView &ix, uint64_t hash, uint32_t *out_eps, int32_t cap, int32_t *out_n,
                             cudaStream_t s);

// ------------------------------------------------------------------------------------------------
// pool state -> request-independent scorer terms (a5-a9 precompute)
// ------------------------------------------------------------------------------------------------
constexpr int kMaxProfiles = 3;            // [0] primary / decode, [1] prefill, [2] encode
// LoRA adapter residency (fwkdl.Metrics.ActiveModels / WaitingModels / MaxActiveModels) for lora-affinity-scorer:
// per endpoint the adapter capacity, per adapter (= registered model id) the endpoints where it is active (1) or
// waiting (2), sorted by endpoint.  ptr == nullptr: no LoRA state (every endpoint scores 0.0 / its capacity tier).
struct LoraDev {
    const int32_t *max_active;   // [E] MaxActiveModels
    const int32_t *n_loaded;     // [E] len(ActiveModels) + len(WaitingModels)
    const uint32_t *ptr;         // [n_models + 1]
    const uint32_t *ep;          // [ptr[n_models]] endpoint slot ids, ascending within an adapter
    const uint8_t *state;        // 1 = active, 2 = waiting
    int32_t n_models;
    int32_t enabled;             // some profile has a lora-affinity scorer
};
struct PoolArrays {
    int32_t E;                 // slot capacity (max_endpoints)
    int32_t n_ext_cols;
    const uint8_t *role;       // [E] epp_role, 0xFF = absent
    const double *kv_usage;    // [E]
    const int32_t *waiting;    // [E]
    const int32_t *running;    // [E]
    const double *ext;         // [n_ext_cols][E]
    LoraDev lora;
};
struct ProfileDerived {
    // per profile, device arrays
    uint8_t *cand;             // [E] 1 = passes the role filter
    double *contrib;           // [EPP_MAX_SCORERS][E] clamp(score)*weight of every request-independent scorer
    double *base;              // [E] ordered weighted sum with every prefix term = clamp(0)*w
    uint32_t *order;           // [Epad] candidates sorted by (base desc, slot asc); then non-candidates
    uint32_t *grp_size;        // [Epad] size of the equal-base group of order[k]
    uint64_t *sort_key;        // [Epad] scratch
    int32_t *n_cand;           // [1]
    int64_t *qminmax;          // [4] waiting min,max, running min,max over candidates
};
// [shard_begin, shard_end): only these slots become candidates (queue min/max still span the whole pool).
cudaError_t launch_pool_prepare(const PoolArrays &pool, const epp_profile_cfg &prof, const ProfileDerived &d,
                                int32_t Epad, uint32_t shard_begin, uint32_t shard_end, cudaStream_t s, int *launches);

// ------------------------------------------------------------------------------------------------
// match + score + pick (a3-a14)
// ------------------------------------------------------------------------------------------------
struct ProfileDev {            // what the pick kernels read (device pointers + config by value)
    epp_profile_cfg cfg;
    const uint8_t *cand;
    const double *contrib;
    const double *base;
    const uint32_t *order;
    const uint32_t *grp_size;
    const int32_t *n_cand;
};
// General evaluation (pick_general.cuh): set when a profile configures the prefix-cache-affinity-filter or pick_k > 1.
struct GenView {
    int32_t on;                // every profile evaluation of the launch goes through gen::eval
    int32_t topk;              // k of the lists below (<= 1: no lists)
    PoolArrays pool;           // raw pool arrays (scores over a narrowed candidate set are not in the snapshot columns)
    uint32_t *topk_picks[kMaxProfiles];   // [R][topk] per profile, or nullptr
    double *topk_scores;       // [R][topk] scores of the primary list, or nullptr
};
struct PickParams {
    int64_t R;
    int32_t E;
    int32_t max_blocks;
    int32_t block_size_tokens;
    int32_t n_profiles;        // 1 or 2 (2 => disagg: [0]=decode, [1]=prefill)
    int32_t encode_on;         // disagg with an encode profile ([2]): runs for requests flagged in `multimodal`
    const uint8_t *multimodal; // [R] or nullptr (encode decider input: hasMultimodalContent)
    int32_t always_disagg;
    int64_t non_cached_tokens;
    ProfileDev prof[kMaxProfiles];
    const uint64_t *hashes;    // [R][max_blocks]
    const int32_t *nblocks;    // [R]
    const int64_t *in_len;     // [R] prompt bytes (decider)
    const uint32_t *model_ids; // [R] registered model (= LoRA adapter) of each request, or nullptr => model 0
    LoraDev lora;
    IndexView index;
    epp_decision *out;         // [R]
    epp_decision_detail *detail;  // [R] or nullptr
    int32_t *out_match;        // dense [R][E] or nullptr (Produce parity mode)
    unsigned long long *work_counters;  // [2] += (probes, postings) of this launch, or nullptr
    // endpoint-sharded mode (nullptr otherwise): the stop rule comes from the OR of all ranks' presence masks
    const uint32_t *global_masks;       // [R][mask_words]
    int32_t mask_words;
    epp_shard_best *shard_out;          // [R] local best record instead of `out`
    // dense-counter kernel (k_match_pick) only: process the requests listed in req_list[0 .. *req_list_n) (the sparse kernel's overflows)
    const int32_t *req_list;
    const int32_t *req_list_n;
    // sparse kernels only: requests whose matched-endpoint set overflowed the per-warp map are appended here
    int32_t *overflow_list;
    int32_t *overflow_n;
    // tie rule: 0 = lowest slot of the arg-max set; else the member of rank tie_rank(seed, 4 * (tie_base + r) + profile)
    uint64_t tie_seed;
    uint64_t tie_base;         // ordinal of request 0 of this launch
    GenView gen;
};
// Fused lookup + match + score + pick, one warp per request.  Per-warp match counters live in shared memory
// (smem = match_pick_smem_bytes(E, false)) or, when E is too large for that, in a zero-initialised global
// scratch of grid * warps * ceil(E/2) words (smem = match_pick_smem_bytes(E, true)).
size_t match_pick_smem_bytes(int32_t E, bool global_counts);
int match_pick_warps_per_cta();
cudaError_t launch_match_pick(const PickParams &p, uint32_t *gscratch, int grid, size_t smem, cudaStream_t s,
                              int *launches);
// The same decision per request with a lane-distributed register map of the matched endpoints (requests whose set
// overflows it are appended to p.overflow_list for the dense-counter kernel).  8 warps per CTA, 6-8 CTAs per SM.
cudaError_t launch_match_pick_sparse(const PickParams &p, int sm_count, cudaStream_t s, int *launches);
// cycle_small.cu: a1-a14 of a small host batch in ONE launch, one CTA per request, prompts read from and decisions
// written to PINNED HOST memory (zero-copy); flags[r] = epoch (| 0x80000000 when the request overflowed the sparse map and
// was appended to pp.overflow_list instead of being decided) is written last, after a system-scope fence.
struct SmallOut {
    epp_decision *dec;             // [R] device view of pinned host memory
    epp_decision_detail *det;      // [R]
    uint32_t *flags;               // [R]
    uint32_t epoch;                // 1 .. 0x7fffffff
    // Prompts still on their way when the kernel starts (the copy runs on another stream): the rows may be read once
    // *arrive == epoch (a stream-ordered 32-bit write behind the copy).  nullptr: resident at launch.
    const uint32_t *arrive;        // device
};
size_t cycle_small_max_blocks();
// Does the kernel bring a whole prompt into shared memory (one bulk copy) before it hashes?  Then a request's flag is
// raised only after its last read of the prompt bytes, and the staging buffer may be refilled as soon as all flags are up.
bool cycle_small_stages_prompt(int32_t max_blocks, int32_t block_bytes);
cudaError_t launch_cycle_small(const HashParams &hp, const PickParams &pp, const SmallOut &so, int align, cudaStream_t s,
                               int *launches);
// Decision logic on injected dense match info (plugin parity / KAT mode).
struct DensePickParams {
    int64_t R;
    int32_t E;
    int32_t block_size_tokens;
    int32_t n_profiles;
    int32_t encode_on;
    const uint8_t *multimodal;
    int32_t always_disagg;
    int64_t non_cached_tokens;
    ProfileDev prof[kMaxProfiles];
    const int32_t *match;      // [R][E]
    const int32_t *total;      // [R]
    const int64_t *in_len;     // [R]
    const uint32_t *model_ids; // [R] or nullptr
    LoraDev lora;
    epp_decision *out;
    epp_decision_detail *detail;
    uint64_t tie_seed;
    uint64_t tie_base;
    GenView gen;
};
cudaError_t launch_dense_pick(const DensePickParams &p, cudaStream_t s, int *launches);
// Scorer.Score parity: out[R][E].  scorer_index -1 => weighted ordered sum (-1.0 for non-candidates).
cudaError_t launch_score_dense(int64_t R, int32_t E, const ProfileDev &prof, const PoolArrays &pool,
                               const int64_t *qminmax, const int32_t *match, const int32_t *total,
                               const uint32_t *model_ids, int32_t scorer_index, double *out, cudaStream_t s,
                               int *launches);

// endpoint-sharded mode: per-request block-presence masks of the LOCAL table (bit i of word i/32 = block i held
// by some endpoint of this shard), and the cross-rank reduction of the per-shard best records.
cudaError_t launch_shard_probe(int64_t R, int32_t max_blocks, const uint64_t *hashes, const int32_t *nblocks,
                               const IndexView &index, uint32_t *out_masks, int32_t mask_words, cudaStream_t s,
                               int *launches);
cudaError_t launch_shard_merge(int64_t R, int32_t n_ranks, const epp_shard_best *all_best, const int32_t *nblocks,
                               epp_decision *out, cudaStream_t s, int *launches);

// shard_p2p.cu: exchanges of the endpoint-sharded mode over NVLink peer memory (no NCCL)
cudaError_t launch_p2p_signal(unsigned long long *flag, unsigned long long epoch, cudaStream_t s);
cudaError_t launch_p2p_signal_unless(unsigned long long *flag, unsigned long long epoch, const int *err_dev, cudaStream_t s);
cudaError_t launch_p2p_wait(unsigned char *const *peers_dev, int n, size_t flag_off, unsigned long long epoch, int *err_dev,
                            unsigned long long timeout_ns, cudaStream_t s);
cudaError_t launch_p2p_or_masks(unsigned char *const *peers_dev, int n, size_t masks_off, unsigned long long n_bytes,
                                void *out, cudaStream_t s);
cudaError_t launch_p2p_gather(unsigned char *const *peers_dev, int n, size_t best_off, unsigned long long n_bytes,
                              void *out, cudaStream_t s);

}  // namespace epp
