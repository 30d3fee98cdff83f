// hash_fused.cu -- the fused hot-path kernel: a1 (prompt lengths + stripe digests + hash chain) and, when kMatch,
// a2-a14 (index walk with the global-stop rule, match counts, scoring, arg-max pick, P/D second stage) in ONE
// kernel.  Fast path: block_bytes % 32 == 0 and 16- or 32-byte aligned prompts.
//
// HBM traffic per request: every prompt byte is read exactly once (coalesced 128/256-bit loads), the 8-byte
// pre-chain digests never leave the SM (shared-memory ring), the block hashes are written once (the PluginState
// stash that PreRequest needs, plugin.go:150-157) and, when kMatch, never read back: the chain warp probes the
// L2-resident index with each hash while it is still in a register and writes the 32-byte decision.
//
// CTA = 9 warps, tile = 32 requests:
//   warps 0-7 (digest): thread (r = t/8, j = t%8) owns block k*8+j of request r in window k: 4 independent XXH64
//       accumulator chains over the block's 32-byte stripes, merge -> m, stored to the stage's [r][j] cell.
//   warp 8 (chain): lane = request.  Walks the window's 8 blocks in order (the only serial part of the digest:
//       m + len, one 8-byte round with h_{i-1}, avalanche), in place.  32 chains advance per instruction, so the
//       serial dependency costs issue slots for one warp only while the 8 digest warps keep the memory pipe busy.
//   warp 9 (match, kMatch only): lane = request.  Per window it cp.asyncs the 8 table slots of its request's new
//       hashes into shared memory, then walks them in order: stop test (plugin.go:221-223) and run-length encoding
//       of the posting sets (a run record is pushed to a per-lane shared-memory queue only when the set changes).
//       Everything divergent -- turning runs into per-endpoint counts, scoring, picking -- is deferred to an
//       epilogue that runs once per tile with all 32 lanes busy.
//   Stages are handed over with named barriers (bar.arrive / bar.sync), 4-deep ring:
//       full (digest -> chain), hashed (chain -> match), empty (match or chain -> digest).
#include "kernels.h"
#include "lane_match.cuh"
#include "xxh64.cuh"

namespace epp {

namespace {
constexpr int kTileR = 32;
constexpr int kWin = 8;
constexpr int kDigestWarps = 8;
constexpr int kDigestThreads = kDigestWarps * 32;
constexpr int kStages = 4;
constexpr int kPitch = kWin + 1;          // u64 cells per request row (odd pitch: conflict-free lane = request reads)
constexpr int kBarFull = 1;               // named barrier ids 1..4
constexpr int kBarEmpty = 1 + kStages;    // 5..8
constexpr int kBarHashed = 1 + 2 * kStages;   // 9..12
constexpr int kBarTail = 1 + 3 * kStages;     // 13
constexpr int kMaxRuns = 24;              // run records per request before the dense-counter fallback takes over
template <bool kMatch> struct Cta { static constexpr int kThreads = kDigestThreads + 32 + (kMatch ? 32 : 0); };

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One 32-byte XXH64 stripe -> four little-endian u64 lanes.  kAlign32: a single 256-bit load (LDG.E.256, exactly one
// DRAM sector per instruction per lane); else two 128-bit loads.
template <bool kAlign32>
__device__ __forceinline__ void load_stripe(const uint8_t *p, uint64_t x[4]) {
    if (kAlign32) {
        asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x[0]), "=l"(x[1]), "=l"(x[2]), "=l"(x[3]) : "l"(p));
    } else {
        uint4 a = __ldg(reinterpret_cast<const uint4 *>(p)), c = __ldg(reinterpret_cast<const uint4 *>(p) + 1);
        x[0] = ((uint64_t)a.y << 32) | a.x; x[1] = ((uint64_t)a.w << 32) | a.z;
        x[2] = ((uint64_t)c.y << 32) | c.x; x[3] = ((uint64_t)c.w << 32) | c.z;
    }
}

struct MatchSmem {                        // match warp state (kMatch only); every array is [..][lane]
    uint4 slot[kWin][2][32];              // the window's table slots: [block in window][half of the 32 bytes][lane]
    uint32_t run[kMaxRuns][7][32];        // run queue: (length, count, w0..w4)
    uint64_t tail_hash[32];
    uint32_t has_tail[32];
};

__device__ __forceinline__ lane::Slot8 ring_read(const MatchSmem &ms, int q, int lane) {
    uint4 a = ms.slot[q][0][lane], b = ms.slot[q][1][lane];
    lane::Slot8 s;
    s.key = ((uint64_t)a.y << 32) | a.x;
    s.cnt = a.z; s.w0 = a.w;
    s.w1 = b.x; s.w2 = b.y; s.w3 = b.z; s.w4 = b.w;
    return s;
}
}  // namespace

template <bool kAlign32, bool kMatch>
__global__ void __launch_bounds__(Cta<kMatch>::kThreads) k_hash_fused(HashParams p, PickParams pk, int n_tiles) {
    constexpr int kProducers = kDigestThreads + 32;       // threads on the full/empty barriers
    __shared__ uint64_t s_m[kStages][kTileR][kPitch];
    __shared__ uint64_t s_off[kTileR];
    __shared__ int64_t s_eff[kTileR];
    __shared__ int32_t s_nfull[kTileR];
    __shared__ int32_t s_maxfull;
    __shared__ __align__(16) unsigned char s_match_raw[kMatch ? sizeof(MatchSmem) : 16];
    MatchSmem &ms = *reinterpret_cast<MatchSmem *>(s_match_raw);

    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    const int64_t bs = p.block_bytes;
    const int n_stripes = (int)(bs >> 5);
    const uint64_t lenp8 = (uint64_t)bs + 8;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t r0 = (int64_t)tile * kTileR;
        // ---- per-request lengths (hashing.go:58-66) by the first warp
        if (t < kTileR) {
            int64_t r = r0 + t;
            uint64_t off = 0;
            int64_t eff = 0;
            int32_t nfull = 0;
            if (r < p.R) {
                uint64_t len;
                if (p.offsets) { off = p.offsets[r]; len = p.lengths ? p.lengths[r] : p.offsets[r + 1] - off; }
                else { off = (uint64_t)r * p.uniform_len; len = p.uniform_len; }
                if (p.in_len) p.in_len[r] = (int64_t)len;
                eff = (int64_t)len;
                int32_t nb = 0;
                if (eff < bs) {
                    eff = 0;
                } else {
                    int64_t cap = bs * (int64_t)p.max_blocks;
                    if (eff > cap) eff = cap;
                    nfull = (int32_t)(eff / bs);
                    nb = nfull + ((eff % bs) ? 1 : 0);
                }
                p.nblocks[r] = nb;
                p.eff_len[r] = eff;
            }
            s_off[t] = off;
            s_eff[t] = eff;
            s_nfull[t] = nfull;
            int mx = nfull;
            for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (t == 0) s_maxfull = mx;
        }
        __syncthreads();
        const int n_win = (s_maxfull + kWin - 1) / kWin;

        if (warp < kDigestWarps) {
            // ================= digest warps =================
            const int r = t / kWin, j = t % kWin;
            const int32_t nfull = s_nfull[r];
            const uint8_t *base = p.data + s_off[r] + (uint64_t)j * (uint64_t)bs;
            for (int k = 0; k < n_win; k++) {
                const int s = k % kStages;
                if (k >= kStages) bar_sync(kBarEmpty + s, kProducers);
                const int b = k * kWin + j;
                if (b < nfull) {
                    const uint8_t *src = base + (uint64_t)k * (uint64_t)(kWin * bs);
                    uint64_t v[4];
                    xxh_init(v);
                    if (n_stripes == 2) {               // the default 64-byte block: both stripes in flight at once
                        uint64_t x0[4], x1[4];
                        load_stripe<kAlign32>(src, x0);
                        load_stripe<kAlign32>(src + 32, x1);
#pragma unroll
                        for (int q = 0; q < 4; q++) v[q] = xxh_round(v[q], x0[q]);
#pragma unroll
                        for (int q = 0; q < 4; q++) v[q] = xxh_round(v[q], x1[q]);
                    } else {
                        for (int st = 0; st < n_stripes; st++) {
                            uint64_t x[4];
                            load_stripe<kAlign32>(src + 32 * st, x);
#pragma unroll
                            for (int q = 0; q < 4; q++) v[q] = xxh_round(v[q], x[q]);
                        }
                    }
                    s_m[s][r][j] = xxh_merge_all(v);
                }
                __threadfence_block();
                bar_arrive(kBarFull + s, kProducers);
            }
            // drain: the empty-barrier arrivals of the last min(n_win, kStages) windows must be consumed before
            // the ring is reused by the next tile
            int first = n_win > kStages ? n_win - kStages : 0;
            for (int k = first; k < n_win; k++) bar_sync(kBarEmpty + (k % kStages), kProducers);
        } else if (warp == kDigestWarps) {
            // ================= chain warp: lane = request =================
            const int64_t r = r0 + lane;
            const int32_t nfull = s_nfull[lane];
            uint64_t prev = 0;
            if (r < p.R) prev = p.seeds[p.model_ids ? p.model_ids[r] : 0];
            for (int k = 0; k < n_win; k++) {
                const int s = k % kStages;
                bar_sync(kBarFull + s, kProducers);
#pragma unroll
                for (int j = 0; j < kWin; j++) {
                    if (k * kWin + j < nfull) {
                        prev = xxh_chain_step32(s_m[s][lane][j], lenp8, prev);
                        s_m[s][lane][j] = prev;
                    }
                }
                __syncwarp();
                // write-out: 8 lanes cover one request's 8 hashes = one 64-byte segment
#pragma unroll
                for (int it = 0; it < kTileR / 4; it++) {
                    int rr = it * 4 + (lane >> 3), jj = lane & 7;
                    int b = k * kWin + jj;
                    if (b < s_nfull[rr]) p.hashes[(r0 + rr) * (int64_t)p.max_blocks + b] = s_m[s][rr][jj];
                }
                __syncwarp();
                if (kMatch) {
                    __threadfence_block();
                    bar_arrive(kBarHashed + s, 64);                 // the match warp releases the stage
                } else {
                    bar_arrive(kBarEmpty + s, kProducers);
                }
            }
            // trailing partial block (hashing.go:90-96): generic tail, rare
            uint64_t tail_hash = 0;
            uint32_t has_tail = 0;
            if (r < p.R) {
                int64_t eff = s_eff[lane];
                if ((int64_t)nfull * bs < eff) {
                    tail_hash = hash_block_generic(p.data + s_off[lane] + (uint64_t)nfull * (uint64_t)bs,
                                                   eff - (int64_t)nfull * bs, prev);
                    p.hashes[r * (int64_t)p.max_blocks + nfull] = tail_hash;
                    has_tail = 1;
                }
            }
            if (kMatch) {
                ms.tail_hash[lane] = tail_hash;
                ms.has_tail[lane] = has_tail;
                __threadfence_block();
                bar_arrive(kBarTail, 64);
            }
        } else if (kMatch) {
            // ================= match warp: lane = request =================
            const int64_t r = r0 + lane;
            const int32_t nfull = s_nfull[lane];
            const IndexSlot *slots = pk.index.slots;
            const uint64_t mask = pk.index.mask;
            lane::Walk w;
            w.init();
            uint32_t nruns = 0;
            bool run_overflow = false;
            auto push_run = [&]() {
                if (w.rl == 0 || w.rc == 0) return;
                if (nruns < (uint32_t)kMaxRuns) {
                    ms.run[nruns][0][lane] = w.rl; ms.run[nruns][1][lane] = w.rc;
                    ms.run[nruns][2][lane] = w.r0; ms.run[nruns][3][lane] = w.r1;
                    ms.run[nruns][4][lane] = w.r2; ms.run[nruns][5][lane] = w.r3;
                    ms.run[nruns][6][lane] = w.r4;
                    nruns++;
                } else {
                    run_overflow = true;
                }
            };
            // Walks one block: stop test + run-length encoding of the posting set (plugin.go:214-230).
            auto step = [&](lane::Slot8 sl, uint64_t h) {
                if (h == kEmptyKey) {                                    // the all-ones key lives in a side record
                    sl.key = h; sl.cnt = pk.index.special.cnt;
                    sl.w0 = pk.index.special.ids[0]; sl.w1 = pk.index.special.ids[1]; sl.w2 = pk.index.special.ids[2];
                    sl.w3 = pk.index.special.ids[3]; sl.w4 = pk.index.special.ids[4];
                } else if (slots) {
                    uint64_t i = h & mask;
                    while (sl.cnt != 0 && sl.key != h) {                 // linear probing past colliding keys (rare)
                        i = (i + 1) & mask;
                        sl = lane::ld_slot(slots + i);
                    }
                } else {
                    sl.cnt = 0;
                }
                w.n_probes++;
                if (sl.cnt == 0) { w.stopped = true; return; }            // nobody holds it: stop (plugin.go:221-223)
                w.n_postings += sl.cnt;
                if (sl.cnt == w.rc && sl.w0 == w.r0 && sl.w1 == w.r1 && sl.w2 == w.r2 && sl.w3 == w.r3 && sl.w4 == w.r4) {
                    w.rl++;
                } else {
                    push_run();
                    w.rl = 1; w.rc = sl.cnt; w.r0 = sl.w0; w.r1 = sl.w1; w.r2 = sl.w2; w.r3 = sl.w3; w.r4 = sl.w4;
                }
            };
            for (int k = 0; k < n_win; k++) {
                const int s = k % kStages;
                bar_sync(kBarHashed + s, 64);
                const bool any = __any_sync(0xffffffffu, !w.stopped && k * kWin < nfull);
                if (any) {
                    uint64_t hh[kWin];
#pragma unroll
                    for (int j = 0; j < kWin; j++) {
                        hh[j] = s_m[s][lane][j];
                        if (k * kWin + j < nfull && !w.stopped && slots) {
                            const IndexSlot *home = slots + (hh[j] & mask);
                            cp_async16(&ms.slot[j][0][lane], home);
                            cp_async16(&ms.slot[j][1][lane], reinterpret_cast<const unsigned char *>(home) + 16);
                        }
                    }
                    cp_async_commit();
                    bar_arrive(kBarEmpty + s, kProducers);          // hashes are in registers: release the stage early
                    cp_async_wait<0>();
#pragma unroll
                    for (int j = 0; j < kWin; j++)
                        if (k * kWin + j < nfull && !w.stopped) step(ring_read(ms, j, lane), hh[j]);
                } else {
                    bar_arrive(kBarEmpty + s, kProducers);
                }
            }
            bar_sync(kBarTail, 64);
            if (ms.has_tail[lane] && !w.stopped) {
                const uint64_t th = ms.tail_hash[lane];
                lane::Slot8 sl;
                sl.key = 0; sl.cnt = 0; sl.w0 = sl.w1 = sl.w2 = sl.w3 = sl.w4 = 0;
                if (slots) sl = lane::ld_slot(slots + (th & mask));
                step(sl, th);
            }
            push_run();
            __syncwarp();
            // ---- epilogue, all 32 lanes together: runs -> per-endpoint counts -> score -> pick (a4-a14)
            if (r < p.R) {
                const int32_t total = nfull + (int32_t)ms.has_tail[lane];
                const uint32_t lo = pk.index.ep_begin, hi = min(pk.index.ep_end, (uint32_t)pk.E);
                lane::Matched m;
                m.n = 0;
                m.overflow = run_overflow;
                for (uint32_t i = 0; i < nruns && !m.overflow; i++)
                    lane::flush_run(m, pk.index, ms.run[i][0][lane], ms.run[i][1][lane], ms.run[i][2][lane],
                                    ms.run[i][3][lane], ms.run[i][4][lane], ms.run[i][5][lane], ms.run[i][6][lane], lo, hi);
                lane::decide(pk, r, m, total);
            }
            if (pk.work_counters) {
                unsigned long long pr = w.n_probes, po = w.n_postings;
                for (int o = 16; o; o >>= 1) {
                    pr += __shfl_xor_sync(0xffffffffu, pr, o);
                    po += __shfl_xor_sync(0xffffffffu, po, o);
                }
                if (lane == 0) {
                    atomicAdd(&pk.work_counters[0], pr);
                    atomicAdd(&pk.work_counters[1], po);
                }
            }
        }
        __syncthreads();
    }
}

template <bool kAlign32, bool kMatch>
static cudaError_t launch_fused_t(const HashParams &p, const PickParams &pk, int sm_count, cudaStream_t s) {
    constexpr int kThreads = Cta<kMatch>::kThreads;
    int n_tiles = (int)((p.R + kTileR - 1) / kTileR);
    static int occ = 0;
    if (!occ) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_hash_fused<kAlign32, kMatch>, kThreads, 0);
        if (occ < 1) occ = 1;
    }
    if (sm_count <= 0) sm_count = 148;
    int grid = n_tiles < sm_count * occ ? n_tiles : sm_count * occ;
    k_hash_fused<kAlign32, kMatch><<<grid, kThreads, 0, s>>>(p, pk, n_tiles);
    return cudaGetLastError();
}

// pick == nullptr: hashing only.  Otherwise the whole cycle; decisions go to pick->out, overflowing requests to
// pick->overflow_list (dense-counter kernel).
cudaError_t launch_hash_fused(const HashParams &p, const PickParams *pick, int align, int sm_count, cudaStream_t s,
                              int *launches) {
    if (p.R <= 0) return cudaSuccess;
    cudaError_t e;
    if (pick) {
        e = align >= 32 ? launch_fused_t<true, true>(p, *pick, sm_count, s) : launch_fused_t<false, true>(p, *pick, sm_count, s);
    } else {
        PickParams none{};
        e = align >= 32 ? launch_fused_t<true, false>(p, none, sm_count, s) : launch_fused_t<false, false>(p, none, sm_count, s);
    }
    if (launches) *launches += 1;
    return e;
}

}  // namespace epp
