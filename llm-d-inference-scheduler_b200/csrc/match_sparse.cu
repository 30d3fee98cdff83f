// match_sparse.cu -- a2-a14 as a STANDALONE kernel over precomputed hashes (plugin-parity modes, endpoint-sharded
// mode, generic-hash-path batches): index lookup with the global-stop rule, per-endpoint match counts, ordered
// weighted sum, arg-max pick and the decode -> decider -> prefill second stage.  ONE WARP PER REQUEST, no shared
// memory, no atomics.  (The main path fuses the same steps into the hash kernel's chain warp: hash_fused.cu.)
//
//   * 32 blocks are probed per step (cold prompts stop after the first); the kernel is kept at 32 registers so that
//     64 warps = 64 requests are resident per SM -- thread-level parallelism hides the L2 latency of the probes;
//   * matched endpoints live in a LANE-DISTRIBUTED register map: lane j holds (endpoint E_j, count C_j), j <
//     n_distinct <= 32; membership tests are ballots;
//   * counting is run-length based: posting lists are sorted (and long lists interned), and blocks of one cached
//     prefix carry the same endpoint set, so the 32 lanes of a chunk fall into a few runs of identical lists; each
//     run adds its length to each of its endpoints.
//
// Exactness: a request whose matched-endpoint set exceeds 32 distinct endpoints is appended to
// PickParams::overflow_list and handled by the dense-counter kernel (pick_kernels.cu), never approximated.
#include <cstdlib>

#include "index.cuh"
#include "score.cuh"

namespace epp {

namespace {
constexpr int kWarps = 8;
#ifndef EPP_MATCH_MIN_CTAS
#define EPP_MATCH_MIN_CTAS 8
#endif
constexpr int kMinCtas = EPP_MATCH_MIN_CTAS;
constexpr uint32_t kNoKey = 0xFFFFFFFFu;
constexpr uint32_t kFull = 0xffffffffu;

struct LaneMap {                   // one entry per lane
    uint32_t e;                    // endpoint held by this lane (kNoKey = none)
    uint32_t c;                    // its match count
    uint32_t n;                    // distinct endpoints so far (warp-uniform)
    bool overflow;                 // warp-uniform
};

// count[e] += c for a warp-uniform (e, c).
__device__ __forceinline__ void map_add(LaneMap &m, uint32_t e, uint32_t c, int lane) {
    uint32_t holder = __ballot_sync(kFull, m.e == e);
    if (holder) {
        if (lane == __ffs(holder) - 1) m.c += c;
    } else if (m.n < 32) {
        if (lane == (int)m.n) { m.e = e; m.c = c; }
        m.n++;
    } else {
        m.overflow = true;
    }
}

__device__ __forceinline__ uint32_t map_get(const LaneMap &m, uint32_t e) {   // warp-uniform e
    uint32_t holder = __ballot_sync(kFull, m.e == e);
    uint32_t c = __shfl_sync(kFull, m.c, holder ? __ffs(holder) - 1 : 0);
    return holder ? c : 0;
}

// Lane-parallel membership: is MY candidate e (different per lane) one of the matched endpoints?
__device__ __forceinline__ bool map_contains_any(const LaneMap &m, uint32_t e) {
    bool hit = false;
    for (uint32_t j = 0; j < m.n; j++) hit |= (e == __shfl_sync(kFull, m.e, (int)j));
    return hit;
}

// One profile for the current request (SchedulerProfile.Run): matched candidates from the lane map, everyone else
// from the (base desc, slot asc) order.
__device__ inline Best eval_profile_lanes(const ProfileDev &pf, int32_t E, const LaneMap &m, int32_t total, int lane,
                                         const LoraDev &lora, int lora_st) {
    Best b;
    best_init(b);
    const int32_t ncand = *pf.n_cand;
    if (ncand == 0) return b;
    const bool mine = (uint32_t)lane < m.n && pf.cand[m.e];
    if (mine) best_add(b, weighted_sum(pf, E, m.e, (int32_t)m.c, total, lora, lora_st), m.e);
    if (m.n) b = best_warp_reduce(b);
    for (int32_t k0 = 0; k0 < ncand; k0 += 32) {
        int32_t k = k0 + lane;
        uint32_t e = k < ncand ? pf.order[k] : kNoKey;
        const bool matched = map_contains_any(m, e);   // shuffles inside: EVERY lane must call it (no short-circuit)
        bool un = k < ncand && !matched;
        uint32_t bal = __ballot_sync(kFull, un);
        if (bal) {
            int first = __ffs(bal) - 1;
            uint32_t ue = __shfl_sync(kFull, e, first);
            double ubase = pf.base[ue];
            uint32_t gsz = pf.grp_size[k0 + first];
            uint32_t same = __popc(__ballot_sync(kFull, mine && pf.base[m.e] == ubase));
            best_add(b, ubase, ue, gsz - same);
            break;
        }
    }
    return b;
}

// Adds one chunk's postings to the lane map.  cnt = this lane's list length (0 when the block is beyond the stop or
// absent), hit.w = the slot's id words.  Runs of lanes with identical lists are added at once.
__device__ __forceinline__ void count_chunk(LaneMap &m, const IndexView &ix, const Hit &hit, uint32_t cnt,
                                            uint32_t shard_lo, uint32_t shard_hi, int lane) {
    const bool in = cnt != 0 && cnt <= (uint32_t)kInlineIds;
    uint32_t w0 = cnt ? hit.w[0] : 0, w1 = in && cnt > 1 ? hit.w[1] : 0;
    uint32_t w2 = in && cnt > 2 ? hit.w[2] : 0, w3 = in && cnt > 3 ? hit.w[3] : 0;
    uint32_t w4 = in && cnt > 4 ? hit.w[4] : 0;
    // run boundaries: a lane starts a run when its (cnt, ids) differ from the previous lane's
    bool diff = lane == 0;
    diff |= cnt != __shfl_up_sync(kFull, cnt, 1);
    diff |= w0 != __shfl_up_sync(kFull, w0, 1);
    diff |= w1 != __shfl_up_sync(kFull, w1, 1);
    diff |= w2 != __shfl_up_sync(kFull, w2, 1);
    diff |= w3 != __shfl_up_sync(kFull, w3, 1);
    diff |= w4 != __shfl_up_sync(kFull, w4, 1);
    uint32_t heads = __ballot_sync(kFull, diff);
    uint32_t nonempty = __ballot_sync(kFull, cnt != 0);
    while (heads) {
        const int s = __ffs(heads) - 1;
        heads &= heads - 1;
        const int end = heads ? __ffs(heads) - 1 : 32;
        if (!((nonempty >> s) & 1u)) continue;                       // run of absent / out-of-range blocks
        const uint32_t len = (uint32_t)(end - s);
        const uint32_t rc = __shfl_sync(kFull, cnt, s);
        const uint32_t r0 = __shfl_sync(kFull, w0, s);
        if (rc <= (uint32_t)kInlineIds) {
            if (r0 >= shard_lo && r0 < shard_hi) map_add(m, r0, len, lane);
            if (rc > 1) { uint32_t e = __shfl_sync(kFull, w1, s); if (e >= shard_lo && e < shard_hi) map_add(m, e, len, lane); }
            if (rc > 2) { uint32_t e = __shfl_sync(kFull, w2, s); if (e >= shard_lo && e < shard_hi) map_add(m, e, len, lane); }
            if (rc > 3) { uint32_t e = __shfl_sync(kFull, w3, s); if (e >= shard_lo && e < shard_hi) map_add(m, e, len, lane); }
            if (rc > 4) { uint32_t e = __shfl_sync(kFull, w4, s); if (e >= shard_lo && e < shard_hi) map_add(m, e, len, lane); }
        } else {
            // spilled (interned) list: r0 is the offset into the postings array
            for (uint32_t k0 = 0; k0 < rc && !m.overflow; k0 += 32) {
                uint32_t mine = (k0 + lane < rc) ? ix.postings[r0 + k0 + lane] : kNoKey;
                uint32_t nk = min(32u, rc - k0);
                for (uint32_t k = 0; k < nk && !m.overflow; k++) {
                    uint32_t e = __shfl_sync(kFull, mine, (int)k);
                    if (e >= shard_lo && e < shard_hi) map_add(m, e, len, lane);
                }
            }
        }
    }
}
// indexer.Get for the hot loop: 32-bit slot arithmetic (the table never exceeds 2^32 slots), no per-call checks.  A
// hash equal to the free-slot sentinel ends at the first free slot like any absent hash; its side record is consulted
// by the caller on that (rare) miss.
__device__ __forceinline__ bool probe_fast(const IndexSlot *slots, uint32_t mask, uint64_t hash, Hit &h) {
    uint32_t i = (uint32_t)hash & mask;
    for (;;) {
        uint64_t key;
        load_slot(slots + i, key, h);
        if (key == hash) return h.cnt != 0;
        if (key == kEmptyKey) { h.cnt = 0; return false; }
        i = (i + 1) & mask;
    }
}
}  // namespace

__global__ void __launch_bounds__(kWarps * 32, kMinCtas) k_match_pick_sparse(PickParams p) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t gwarp = (int64_t)blockIdx.x * kWarps + warp;
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    unsigned long long w_probes = 0, w_postings = 0;
    const uint32_t shard_lo = p.index.ep_begin, shard_hi = min(p.index.ep_end, (uint32_t)p.E);
    const IndexSlot *slots = p.index.slots;
    const uint32_t mask32 = (uint32_t)p.index.mask;
    const bool counting = p.work_counters != nullptr;

    for (int64_t r = gwarp; r < p.R; r += nwarps) {
        const int32_t total = p.nblocks[r];
        const uint64_t *row = p.hashes + r * (int64_t)p.max_blocks;
        LaneMap m;
        m.e = kNoKey; m.c = 0; m.n = 0; m.overflow = false;
        // ---- a2/a3: probe in block order, 32 blocks per step; global stop at the first block nobody holds
        //      (plugin.go:214-230).  One chunk at a time keeps the kernel at 32 registers = 64 resident warps per SM:
        //      measured faster than probing four chunks per step at half the occupancy (0.173 ms vs 0.221 ms).
        bool stopped = false;
        uint64_t hnext = lane < total ? row[lane] : 0;
        for (int32_t cq = 0; cq < total && !stopped; cq += 32) {
            const int32_t i = cq + lane;
            const uint64_t hcur = hnext;
            if (i + 32 < total) hnext = row[i + 32];                  // next chunk's hashes in flight during this one
            Hit hit;
            hit.cnt = 0;
            if (i < total && slots) {
                if (!probe_fast(slots, mask32, hcur, hit)) {
                    hit.cnt = 0;
                    if (hcur == kEmptyKey && !probe(p.index, hcur, hit)) hit.cnt = 0;     // the sentinel hash's side record
                }
            } else if (i < total && hcur == kEmptyKey) {
                if (!probe(p.index, hcur, hit)) hit.cnt = 0;
            }
            uint32_t miss;
            if (p.global_masks) {   // sharded: a block is missing only if NO rank holds it
                uint32_t word = p.global_masks[r * (int64_t)p.mask_words + (cq >> 5)];
                uint32_t valid = (total - cq) >= 32 ? kFull : ((1u << (total - cq)) - 1u);
                miss = ~word & valid;
            } else {
                miss = __ballot_sync(kFull, i < total && hit.cnt == 0);
            }
            const int32_t limit = miss ? cq + (__ffs(miss) - 1) : total;
            const uint32_t cnt = (i < limit) ? hit.cnt : 0;
            if (counting) {
                if (lane == 0) w_probes += (unsigned long long)((miss ? limit + 1 : min(total, cq + 32)) - cq);
                w_postings += cnt;
            }
            if (limit > cq) count_chunk(m, p.index, hit, cnt, shard_lo, shard_hi, lane);
            if (miss) stopped = true;
        }
        // ---- lora-affinity: the endpoints where the request's adapter is active or waiting get a request-dependent
        //      score, so they join the map (count 0) and are evaluated one by one like the prefix holders
        int lora_st = 0;
        if (p.lora.enabled && p.lora.ptr && !m.overflow) {
            const uint32_t a = p.model_ids ? p.model_ids[r] : 0u;
            if (a < (uint32_t)p.lora.n_models) {
                const uint32_t lo = p.lora.ptr[a], hi = p.lora.ptr[a + 1];
                for (uint32_t k0 = lo; k0 < hi && !m.overflow; k0 += 32) {
                    const uint32_t mine_e = (k0 + lane < hi) ? p.lora.ep[k0 + lane] : kNoKey;
                    const uint32_t nk = min(32u, hi - k0);
                    for (uint32_t k = 0; k < nk && !m.overflow; k++) {
                        const uint32_t e = __shfl_sync(kFull, mine_e, (int)k);
                        if (e >= shard_lo && e < shard_hi) map_add(m, e, 0, lane);
                    }
                }
                if ((uint32_t)lane < m.n) lora_st = lora_lookup(p.lora, a, m.e);
            }
        }
        if (m.overflow) {
            // hand the request to the dense-counter kernel
            if (lane == 0 && p.overflow_list) p.overflow_list[atomicAdd(p.overflow_n, 1)] = (int32_t)r;
            continue;
        }
        // ---- a5-a10: primary profile
        Best b0 = eval_profile_lanes(p.prof[0], p.E, m, total, lane, p.lora, lora_st);
        epp_decision d;
        d.status = b0.ties ? 0 : -1;
        d.pick = b0.ties ? b0.pick : EPP_NO_ENDPOINT;
        d.score = b0.ties ? b0.val : 0.0;
        d.prefill_pick = EPP_NO_ENDPOINT;
        d.tie_count = b0.ties;
        d.total_blocks = total;
        d.match_blocks = (b0.ties && m.n) ? (int32_t)map_get(m, b0.pick) : 0;
        epp_decision_detail dd;
        dd.prefill_score = 0.0; dd.prefill_tie_count = 0; dd.prefill_ran = 0;
        // ---- a13: decode -> decider -> prefill (disagg_profile_handler.go:264-308)
        if (p.n_profiles == 2 && b0.ties) {
            bool go = p.always_disagg || pd_decide(p.non_cached_tokens, p.in_len[r], d.match_blocks, p.block_size_tokens);
            if (go) {
                dd.prefill_ran = 1;
                Best b1 = eval_profile_lanes(p.prof[1], p.E, m, total, lane, p.lora, lora_st);
                if (b1.ties) { d.prefill_pick = b1.pick; dd.prefill_score = b1.val; dd.prefill_tie_count = b1.ties; }
            }
        }
        if (lane == 0) {
            if (p.shard_out) {
                epp_shard_best sb;
                sb.score = d.score; sb.pick = d.pick; sb.tie_count = d.tie_count;
                sb.match_blocks = d.match_blocks; sb.status = d.status;
                p.shard_out[r] = sb;
            } else {
                p.out[r] = d;
                if (p.detail) p.detail[r] = dd;
            }
        }
    }
    if (p.work_counters) {
        for (int o = 16; o; o >>= 1) w_postings += __shfl_xor_sync(kFull, w_postings, o);
        if (lane == 0) {
            atomicAdd(&p.work_counters[0], w_probes);
            atomicAdd(&p.work_counters[1], w_postings);
        }
    }
}

cudaError_t launch_match_pick_sparse(const PickParams &p, int sm_count, cudaStream_t s, int *launches) {
    if (p.R <= 0) return cudaSuccess;
    static int occ = 0;
    if (!occ) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_match_pick_sparse, kWarps * 32, 0);
        if (occ < 1) occ = 1;
    }
    if (sm_count <= 0) sm_count = 148;
    int64_t need = (p.R + kWarps - 1) / kWarps;
    int grid = (int)(need < (int64_t)sm_count * occ ? need : (int64_t)sm_count * occ);
    static int full_grid = -1;          // EPP_MATCH_FULL_GRID=1: one warp per request, the block scheduler balances
    if (full_grid < 0) { const char *v = getenv("EPP_MATCH_FULL_GRID"); full_grid = (v && v[0] == '1') ? 1 : 0; }
    if (full_grid) grid = (int)need;
    k_match_pick_sparse<<<grid, kWarps * 32, 0, s>>>(p);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace epp
