// lane_match.cuh -- the lane = request form of a2-a14 (index walk with the global-stop rule, run-length counting
// into a small per-thread endpoint map, single-thread scoring of one SchedulerProfile and the decode -> decider ->
// prefill decision).  Used by the chain warp of the fused kernel (hash_fused.cu), where a warp advances 32
// requests per instruction and the block hashes never leave the SM.
#pragma once
#include "index.cuh"
#include "score.cuh"

namespace epp {
namespace lane {

constexpr int kMaxMatched = 16;
constexpr int kGroup = 4;                 // blocks per step (one 32-byte hash load)
constexpr uint32_t kNone = 0xFFFFFFFFu;

struct Slot8 {                            // raw 32-byte slot
    uint64_t key;
    uint32_t cnt, w0, w1, w2, w3, w4;
};

__device__ __forceinline__ Slot8 ld_slot(const IndexSlot *p) {
    uint64_t a, b, c, d;
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    Slot8 s;
    s.key = a;
    s.cnt = (uint32_t)b; s.w0 = (uint32_t)(b >> 32);
    s.w1 = (uint32_t)c;  s.w2 = (uint32_t)(c >> 32);
    s.w3 = (uint32_t)d;  s.w4 = (uint32_t)(d >> 32);
    return s;
}

struct Matched {                          // per-thread endpoint map (local memory, a few entries)
    uint32_t e[kMaxMatched];
    uint32_t c[kMaxMatched];
    int n;
    bool overflow;
};

__device__ __forceinline__ void matched_add(Matched &m, uint32_t e, uint32_t n, uint32_t lo, uint32_t hi) {
    if (e < lo || e >= hi) return;        // outside this shard / the slot range: keeps the walk alive only
    for (int j = 0; j < m.n; j++)
        if (m.e[j] == e) { m.c[j] += n; return; }
    if (m.n < kMaxMatched) { m.e[m.n] = e; m.c[m.n] = n; m.n++; }
    else m.overflow = true;
}

__device__ __forceinline__ uint32_t matched_get(const Matched &m, uint32_t e) {
    for (int j = 0; j < m.n; j++)
        if (m.e[j] == e) return m.c[j];
    return 0;
}

// Adds a run of `len` blocks that all carry the posting set described by (cnt, w0..w4).
__device__ __forceinline__ void flush_run(Matched &m, const IndexView &ix, uint32_t len, uint32_t cnt, uint32_t w0,
                                          uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, uint32_t lo, uint32_t hi) {
    if (len == 0 || cnt == 0) return;
    if (cnt <= (uint32_t)kInlineIds) {
        matched_add(m, w0, len, lo, hi);
        if (cnt > 1) matched_add(m, w1, len, lo, hi);
        if (cnt > 2) matched_add(m, w2, len, lo, hi);
        if (cnt > 3) matched_add(m, w3, len, lo, hi);
        if (cnt > 4) matched_add(m, w4, len, lo, hi);
    } else {
        for (uint32_t k = 0; k < cnt && !m.overflow; k++) matched_add(m, ix.postings[w0 + k], len, lo, hi);
    }
}

// One profile (SchedulerProfile.Run) for one request, single thread.
__device__ inline Best eval_profile_thread(const ProfileDev &pf, int32_t E, const Matched &m, int32_t total) {
    Best b;
    best_init(b);
    const int32_t ncand = *pf.n_cand;
    if (ncand == 0) return b;
    for (int j = 0; j < m.n; j++) {
        uint32_t e = m.e[j];
        if (pf.cand[e]) best_add(b, weighted_sum(pf, E, e, (int32_t)m.c[j], total), e);
    }
    // best of everyone else = first unmatched entry of the (base desc, slot asc) order
    for (int32_t k = 0; k < ncand; k++) {
        uint32_t e = pf.order[k];
        bool matched = false;
        for (int j = 0; j < m.n; j++) matched |= (m.e[j] == e);
        if (matched) continue;
        double ubase = pf.base[e];
        uint32_t same = 0;
        for (int j = 0; j < m.n; j++) {
            uint32_t me = m.e[j];
            if (pf.cand[me] && pf.base[me] == ubase) same++;
        }
        best_add(b, ubase, e, pf.grp_size[k] - same);
        break;
    }
    return b;
}

// Per-lane state of the walk over one request's blocks (plugin.go:214-230).
struct Walk {
    uint32_t rl, rc, r0, r1, r2, r3, r4;      // current run: length and posting-set signature
    uint32_t n_probes, n_postings;
    bool stopped;
    __device__ __forceinline__ void init() {
        rl = rc = r0 = r1 = r2 = r3 = r4 = 0;
        n_probes = n_postings = 0;
        stopped = false;
    }
    // Consume block b's slot (already fetched into s for the home position of hash h).  `present_elsewhere`: in the
    // endpoint-sharded mode, the OR of all ranks' presence bits for this block; otherwise ignored.
    __device__ __forceinline__ void consume(Matched &m, const IndexView &ix, Slot8 s, uint64_t h, bool sharded,
                                            bool present_anywhere, uint32_t lo, uint32_t hi) {
        if (h == kEmptyKey) {                                    // the all-ones key lives in a side record
            s.key = h; s.cnt = ix.special.cnt;
            s.w0 = ix.special.ids[0]; s.w1 = ix.special.ids[1]; s.w2 = ix.special.ids[2];
            s.w3 = ix.special.ids[3]; s.w4 = ix.special.ids[4];
        } else if (ix.slots) {
            uint64_t i = h & ix.mask;
            while (s.key != h && s.key != kEmptyKey) {           // linear probing past colliding keys (rare)
                i = (i + 1) & ix.mask;
                s = ld_slot(ix.slots + i);
            }
            if (s.key != h) s.cnt = 0;
        } else {
            s.cnt = 0;
        }
        n_probes++;
        bool missing = sharded ? !present_anywhere : (s.cnt == 0);
        if (missing) {                                           // plugin.go:221-223: nobody holds it -> stop
            stopped = true;
            return;
        }
        n_postings += s.cnt;
        if (s.cnt == rc && s.w0 == r0 && s.w1 == r1 && s.w2 == r2 && s.w3 == r3 && s.w4 == r4) {
            rl++;
        } else {
            flush_run(m, ix, rl, rc, r0, r1, r2, r3, r4, lo, hi);
            rl = 1; rc = s.cnt; r0 = s.w0; r1 = s.w1; r2 = s.w2; r3 = s.w3; r4 = s.w4;
        }
    }
    __device__ __forceinline__ void finish(Matched &m, const IndexView &ix, uint32_t lo, uint32_t hi) {
        flush_run(m, ix, rl, rc, r0, r1, r2, r3, r4, lo, hi);
        rl = 0;
    }
};

// Scores the profiles and writes the decision of request r (or hands it to the dense-counter kernel on overflow).
__device__ inline void decide(const PickParams &p, int64_t r, const Matched &m, int32_t total) {
    if (m.overflow) {
        if (p.overflow_list) p.overflow_list[atomicAdd(p.overflow_n, 1)] = (int32_t)r;
        return;
    }
    Best b0 = eval_profile_thread(p.prof[0], p.E, m, total);
    epp_decision d;
    d.status = b0.ties ? 0 : -1;
    d.pick = b0.ties ? b0.pick : EPP_NO_ENDPOINT;
    d.score = b0.ties ? b0.val : 0.0;
    d.prefill_pick = EPP_NO_ENDPOINT;
    d.tie_count = b0.ties;
    d.total_blocks = total;
    d.match_blocks = b0.ties ? (int32_t)matched_get(m, b0.pick) : 0;
    epp_decision_detail dd;
    dd.prefill_score = 0.0; dd.prefill_tie_count = 0; dd.prefill_ran = 0;
    if (p.n_profiles == 2 && b0.ties) {      // decode -> decider -> prefill (disagg_profile_handler.go:264-308)
        bool go = p.always_disagg || pd_decide(p.non_cached_tokens, p.in_len[r], d.match_blocks, p.block_size_tokens);
        if (go) {
            dd.prefill_ran = 1;
            Best b1 = eval_profile_thread(p.prof[1], p.E, m, total);
            if (b1.ties) { d.prefill_pick = b1.pick; dd.prefill_score = b1.val; dd.prefill_tie_count = b1.ties; }
        }
    }
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

}  // namespace lane
}  // namespace epp
