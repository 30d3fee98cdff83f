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
constexpr int kMaxTileR = 32;             // tile = TR requests (16 or 32): chain-warp lanes = requests
constexpr int kWin = 8;
constexpr int kStages = 4;
constexpr int kPitch = kWin + 1;          // u64 cells per request row (odd pitch: conflict-free lane = request reads)
constexpr int kBarFull = 1;               // named barrier ids 1..4
constexpr int kBarEmpty = 1 + kStages;    // 5..8
constexpr int kBarHashed = 1 + 2 * kStages;   // 9..12
constexpr int kBarTail = 1 + 3 * kStages;     // 13
#ifndef EPP_STRIPE_AT_A_TIME
#define EPP_STRIPE_AT_A_TIME 0
#endif
#ifndef EPP_HASH_MIN_CTAS
#define EPP_HASH_MIN_CTAS 4
#endif
constexpr int kHashMinCtas = EPP_HASH_MIN_CTAS;
constexpr int kLag = 3;                   // digest warps match window k-kLag while hashing window k (kLag < kStages)
constexpr int kMaxRuns = 24;              // run records per request before the dense-counter fallback takes over

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

// Per-request lengths (hashing.go:58-66) for one tile; called by the first warp (threads t < 32; t >= TR idle).
__device__ __forceinline__ void tile_lengths(const HashParams &p, int64_t r0, int t, int tr, uint64_t *s_off,
                                             int64_t *s_eff, int32_t *s_nfull, int32_t *s_maxfull, int rows_alloc = 0) {
    const int64_t bs = p.block_bytes;
    int64_t r = r0 + t;
    uint64_t off = 0;
    int64_t eff = 0;
    int32_t nfull = 0;
    if (t < tr && r < p.R) {
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
    if (t < tr) {
        s_off[t] = off;
        s_eff[t] = eff;
        s_nfull[t] = nfull;
    } else if (t < rows_alloc) {          // rows of the CTA that this (smaller) tile does not use
        s_off[t] = 0;
        s_eff[t] = 0;
        s_nfull[t] = 0;
    }
    int mx = nfull;
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (t == 0) *s_maxfull = mx;
}

// Stripe rounds + merge of one full block (part A of xxh64.cuh).
template <bool kAlign32>
__device__ __forceinline__ uint64_t block_digest(const uint8_t *src, int n_stripes) {
    uint64_t v[4];
    xxh_init(v);
    if (n_stripes == 2 && !EPP_STRIPE_AT_A_TIME) {   // the default 64-byte block: both stripes in flight at once
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
    return xxh_merge_all(v);
}

// Same, for the default 64-byte block with the two stripes already in registers (software prefetch).
__device__ __forceinline__ uint64_t block_digest64(const uint64_t x0[4], const uint64_t x1[4]) {
    uint64_t v[4];
    xxh_init(v);
#pragma unroll
    for (int q = 0; q < 4; q++) v[q] = xxh_round(v[q], x0[q]);
#pragma unroll
    for (int q = 0; q < 4; q++) v[q] = xxh_round(v[q], x1[q]);
    return xxh_merge_all(v);
}

// Chain warp, one window: the serial part of the digest for 32 requests at once, in place, then the 8 hashes of
// each request written as one 64-byte segment.
template <int TR, int W = kWin>
__device__ __forceinline__ uint64_t chain_window(uint64_t (*sm)[W + 1], const int32_t *s_nfull, const HashParams &p,
                                                 int64_t r0, int k, int lane, int32_t nfull, uint64_t lenp8,
                                                 uint64_t prev) {
    if (lane < TR) {
#pragma unroll
        for (int j = 0; j < W; j++) {
            if (k * W + j < nfull) {
                prev = xxh_chain_step32(sm[lane][j], lenp8, prev);
                sm[lane][j] = prev;
            }
        }
    }
    __syncwarp();
#pragma unroll
    constexpr int kRows = 32 / W;                 // request rows written per warp instruction
    for (int it = 0; it < TR / kRows; it++) {
        int rr = it * kRows + lane / W, jj = lane % W;
        int b = k * W + jj;
        if (b < s_nfull[rr]) p.hashes[(r0 + rr) * (int64_t)p.max_blocks + b] = sm[rr][jj];
    }
    __syncwarp();
    return prev;
}
}  // namespace

// =====================================================================================================
// k_hash_fused: a1 only (epp_hash_prompts, Produce-parity and sharded modes, A/B runs)
// =====================================================================================================
template <bool kAlign32, int TR, int W = kWin, int MINCTA = kHashMinCtas>
__global__ void __launch_bounds__(TR * W + 32, MINCTA) k_hash_fused(HashParams p, int n_tiles) {
    constexpr int kTileR = TR;
    constexpr int kDigestThreads = TR * W, kDigestWarps = kDigestThreads / 32;
    constexpr int kProducers = kDigestThreads + 32;
    __shared__ uint64_t s_m[kStages][kTileR][(W + 1)];
    __shared__ uint64_t s_off[kTileR];
    __shared__ int64_t s_eff[kTileR];
    __shared__ int32_t s_nfull[kTileR];
    __shared__ int32_t s_maxfull;

    // Role placement: the chain warp is the serial critical path and warps are bound to the four SM sub-partitions by
    // warp index mod 4; with the chain always on the last warp every co-resident CTA puts it on the SAME scheduler.
    // p.chain_spread rotates it by the CTA's residency slot (blockIdx / SM count) so the chains sit on different ones.
    const int lane = threadIdx.x & 31;
    const int hw_warp = threadIdx.x >> 5;
    const int chain_at = p.chain_spread ? (int)((blockIdx.x / (unsigned)max(p.sm_count, 1)) & 3u) : kDigestWarps;
    const int warp = hw_warp == chain_at ? kDigestWarps : (hw_warp > chain_at ? hw_warp - 1 : hw_warp);
    const int t = warp * 32 + lane;
    const int64_t bs = p.block_bytes;
    const int n_stripes = (int)(bs >> 5);
    const uint64_t lenp8 = (uint64_t)bs + 8;

    // rows per tile: TR, or fewer when the launcher balances the waves of the persistent grid (p.tile_rows)
    const int tr = (p.tile_rows > 0 && p.tile_rows < TR) ? p.tile_rows : TR;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t r0 = (int64_t)tile * tr;
        if (t < 32) tile_lengths(p, r0, t, tr, s_off, s_eff, s_nfull, &s_maxfull, TR);
        __syncthreads();
        const int n_win = (s_maxfull + W - 1) / W;

        if (warp < kDigestWarps) {
            const int r = t / W, j = t % W;
            const int32_t nfull = s_nfull[r];
            const uint8_t *base = p.data + s_off[r] + (uint64_t)j * (uint64_t)bs;
            if (n_stripes == 2 && p.prefetch) {
                // default 64-byte blocks: the loads of window k+1 are issued before window k is hashed, so every
                // digest thread always has 64 bytes in flight
                uint64_t xa[4] = {0, 0, 0, 0}, xb[4] = {0, 0, 0, 0};
                if (j < nfull) { load_stripe<kAlign32>(base, xa); load_stripe<kAlign32>(base + 32, xb); }
                for (int k = 0; k < n_win; k++) {
                    const int s = k % kStages;
                    uint64_t x0[4], x1[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) { x0[q] = xa[q]; x1[q] = xb[q]; }
                    if ((k + 1) * W + j < nfull) {
                        const uint8_t *nx = base + (uint64_t)(k + 1) * (uint64_t)(W * bs);
                        load_stripe<kAlign32>(nx, xa);
                        load_stripe<kAlign32>(nx + 32, xb);
                    }
                    if (k >= kStages) bar_sync(kBarEmpty + s, kProducers);
                    if (k * W + j < nfull) s_m[s][r][j] = block_digest64(x0, x1);
                    __threadfence_block();
                    bar_arrive(kBarFull + s, kProducers);
                }
            } else {
                for (int k = 0; k < n_win; k++) {
                    const int s = k % kStages;
                    if (k >= kStages) bar_sync(kBarEmpty + s, kProducers);
                    if (k * W + j < nfull)
                        s_m[s][r][j] = block_digest<kAlign32>(base + (uint64_t)k * (uint64_t)(W * bs), n_stripes);
                    __threadfence_block();
                    bar_arrive(kBarFull + s, kProducers);
                }
            }
            int first = n_win > kStages ? n_win - kStages : 0;
            for (int k = first; k < n_win; k++) bar_sync(kBarEmpty + (k % kStages), kProducers);
        } else {
            const bool mine = lane < tr;
            const int64_t r = r0 + lane;
            const int32_t nfull = mine ? s_nfull[lane] : 0;
            uint64_t prev = 0;
            if (mine && r < p.R) prev = p.seeds[p.model_ids ? p.model_ids[r] : 0];
            for (int k = 0; k < n_win; k++) {
                const int s = k % kStages;
                bar_sync(kBarFull + s, kProducers);
                prev = chain_window<TR, W>(s_m[s], s_nfull, p, r0, k, lane, nfull, lenp8, prev);
                bar_arrive(kBarEmpty + s, kProducers);
            }
            if (mine && r < p.R) {                   // trailing partial block (hashing.go:90-96): generic tail, rare
                int64_t eff = s_eff[lane];
                if ((int64_t)nfull * bs < eff)
                    p.hashes[r * (int64_t)p.max_blocks + nfull] = hash_block_generic(
                        p.data + s_off[lane] + (uint64_t)nfull * (uint64_t)bs, eff - (int64_t)nfull * bs, prev);
            }
        }
        __syncthreads();
    }
}

// =====================================================================================================
// k_hash_wide: a1 with 32-block windows.  Same pipeline as k_hash_fused, but a digest warp reads 32 consecutive
// blocks (2 KiB contiguous) of ONE request per load instruction instead of 8 blocks of four requests -- DRAM sees
// 2-KiB bursts, which is what the streaming v1 kernel had (5.85 TB/s there vs 4.85 TB/s with 512-byte bursts).
// Each digest warp owns four requests of the tile and hashes their windows one after the other.
// =====================================================================================================
template <bool kAlign32>
__global__ void __launch_bounds__(288) k_hash_wide(HashParams p, int n_tiles) {
    constexpr int TR = 32, W = 32, PW = W + 1, kDigestWarps = 8, kProducers = kDigestWarps * 32 + 32;
    __shared__ uint64_t s_m[kStages][TR][PW];
    __shared__ uint64_t s_off[TR];
    __shared__ int64_t s_eff[TR];
    __shared__ int32_t s_nfull[TR];
    __shared__ int32_t s_maxfull;

    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    const int64_t bs = p.block_bytes;
    const int n_stripes = (int)(bs >> 5);
    const uint64_t lenp8 = (uint64_t)bs + 8;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t r0 = (int64_t)tile * TR;
        if (t < 32) tile_lengths(p, r0, t, TR, s_off, s_eff, s_nfull, &s_maxfull);
        __syncthreads();
        const int n_win = (s_maxfull + W - 1) / W;

        if (warp < kDigestWarps) {
            // ================= digest warps: warp w owns requests 4w .. 4w+3, lane = block in window =================
            const uint8_t *base[4];
            int32_t nf[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                base[q] = p.data + s_off[warp * 4 + q] + (uint64_t)lane * (uint64_t)bs;
                nf[q] = s_nfull[warp * 4 + q];
            }
            for (int k = 0; k < n_win; k++) {
                const int s = k % kStages;
                if (k >= kStages) bar_sync(kBarEmpty + s, kProducers);
                const int b = k * W + lane;
                const uint64_t woff = (uint64_t)k * (uint64_t)(W * bs);
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (b < nf[q]) s_m[s][warp * 4 + q][lane] = block_digest<kAlign32>(base[q] + woff, n_stripes);
                __threadfence_block();
                bar_arrive(kBarFull + s, kProducers);
            }
            int first = n_win > kStages ? n_win - kStages : 0;
            for (int k = first; k < n_win; k++) bar_sync(kBarEmpty + (k % kStages), kProducers);
        } else {
            // ================= chain warp: lane = request =================
            const int64_t r = r0 + lane;
            const int32_t nfull = s_nfull[lane];
            uint64_t prev = 0;
            if (r < p.R) prev = p.seeds[p.model_ids ? p.model_ids[r] : 0];
            for (int k = 0; k < n_win; k++) {
                const int s = k % kStages;
                bar_sync(kBarFull + s, kProducers);
#pragma unroll 8
                for (int j = 0; j < W; j++) {
                    if (k * W + j < nfull) {
                        prev = xxh_chain_step32(s_m[s][lane][j], lenp8, prev);
                        s_m[s][lane][j] = prev;
                    }
                }
                __syncwarp();
                // write-out: the 32 hashes of each request's window = one 256-byte segment
                const int b = k * W + lane;
#pragma unroll 4
                for (int rr = 0; rr < TR; rr++)
                    if (b < s_nfull[rr]) p.hashes[(r0 + rr) * (int64_t)p.max_blocks + b] = s_m[s][rr][lane];
                __syncwarp();
                bar_arrive(kBarEmpty + s, kProducers);
            }
            if (r < p.R) {                           // trailing partial block (hashing.go:90-96): generic tail, rare
                int64_t eff = s_eff[lane];
                if ((int64_t)nfull * bs < eff)
                    p.hashes[r * (int64_t)p.max_blocks + nfull] = hash_block_generic(
                        p.data + s_off[lane] + (uint64_t)nfull * (uint64_t)bs, eff - (int64_t)nfull * bs, prev);
            }
        }
        __syncthreads();
    }
}

// =====================================================================================================
// k_cycle_fused: a1-a14 in one kernel.  Same hashing pipeline; the index probes of window k-2 are spread over the
// 8 digest warps (thread (r, j) probes block j of that window for request r while it hashes window k), so the
// lookup costs no serial warp time.  Per-request walk state lives in shared memory and is touched only by the
// request's own 8 lanes; run records are closed and scored by the chain warp once per tile.
// =====================================================================================================
template <int TR>
struct CycleSmem {
    uint4 slot[2][TR * kWin];                 // each digest thread's in-flight table slot (cp.async target)
    int32_t stopped[TR];                      // walk reached a block nobody holds
    int32_t stop_b[TR];                       // ... at this block index
    int32_t nrec[TR];
    int32_t overflow[TR];
    uint32_t carry[6][TR];                    // posting-set signature (cnt, w0..w4) of the last valid block
    uint32_t rec[kMaxRuns][7][TR];            // run records: (start block, cnt, w0..w4)
};

template <bool kAlign32, int TR>
__global__ void __launch_bounds__(TR * kWin + 32) k_cycle_fused(HashParams p, PickParams pk, int n_tiles) {
    constexpr int kTileR = TR;
    constexpr int kDigestThreads = TR * kWin, kDigestWarps = kDigestThreads / 32;
    constexpr int kAll = kDigestThreads + 32;
    __shared__ uint64_t s_m[kStages][kTileR][kPitch];
    __shared__ uint64_t s_off[kTileR];
    __shared__ int64_t s_eff[kTileR];
    __shared__ int32_t s_nfull[kTileR];
    __shared__ int32_t s_maxfull;
    __shared__ __align__(16) CycleSmem<TR> cs;

    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    const int64_t bs = p.block_bytes;
    const int n_stripes = (int)(bs >> 5);
    const uint64_t lenp8 = (uint64_t)bs + 8;
    const IndexSlot *slots = pk.index.slots;
    const uint64_t mask = pk.index.mask;
    unsigned long long n_probes = 0, n_postings = 0;      // per-thread work counters

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t r0 = (int64_t)tile * kTileR;
        if (t < 32) {
            tile_lengths(p, r0, t, TR, s_off, s_eff, s_nfull, &s_maxfull);
            if (t < TR) {
                cs.stopped[t] = 0; cs.stop_b[t] = 0; cs.nrec[t] = 0; cs.overflow[t] = 0;
#pragma unroll
                for (int q = 0; q < 6; q++) cs.carry[q][t] = 0;
            }
        }
        __syncthreads();
        const int n_win = (s_maxfull + kWin - 1) / kWin;

        if (warp < kDigestWarps) {
            // ================= digest warps: hash window k, match window k-2 =================
            const int r = t / kWin, j = t % kWin;
            const int g8 = (lane >> 3) * 8;                       // first lane of this request's 8-lane group
            const int32_t nfull = s_nfull[r];
            const uint8_t *base = p.data + s_off[r] + (uint64_t)j * (uint64_t)bs;
            for (int k = 0; k < n_win + kLag; k++) {
                // ---- issue the probe of window kk = k-kLag (its hashes are final once the chain warp arrived)
                const int kk = k - kLag;
                bool active = false, any_active = false;
                uint64_t h = 0;
                if (kk >= 0) {
                    const int s2 = kk % kStages;
                    bar_sync(kBarHashed + s2, kAll);
                    active = (kk * kWin + j < nfull) && !cs.stopped[r];
                    any_active = __any_sync(0xffffffffu, active);   // most late windows: all 4 requests already stopped
                    if (active) {
                        h = s_m[s2][r][j];
                        if (slots) {
                            const unsigned char *home = reinterpret_cast<const unsigned char *>(slots + (h & mask));
                            cp_async16(&cs.slot[0][t], home);
                            cp_async16(&cs.slot[1][t], home + 16);
                        }
                    }
                    if (any_active) cp_async_commit();
                }
                // ---- hash window k (the stage was released when every digest thread passed hashed[k-4] two
                //      iterations ago; its cells are private to this thread until bar_arrive(full))
                if (k < n_win) {
                    const int s = k % kStages;
                    if (k * kWin + j < nfull)
                        s_m[s][r][j] = block_digest<kAlign32>(base + (uint64_t)k * (uint64_t)(kWin * bs), n_stripes);
                    __threadfence_block();
                    bar_arrive(kBarFull + s, kAll);
                }
                // ---- consume the probe: stop test + run-length encoding within the 8-lane group
                if (any_active) {
                    cp_async_wait<0>();
                    lane::Slot8 sl;
                    sl.key = 0; sl.cnt = 0; sl.w0 = sl.w1 = sl.w2 = sl.w3 = sl.w4 = 0;
                    if (active) {
                        if (h == kEmptyKey) {                      // the all-ones key lives in a side record
                            sl.key = h; sl.cnt = pk.index.special.cnt;
                            sl.w0 = pk.index.special.ids[0]; sl.w1 = pk.index.special.ids[1]; sl.w2 = pk.index.special.ids[2];
                            sl.w3 = pk.index.special.ids[3]; sl.w4 = pk.index.special.ids[4];
                        } else if (slots) {
                            uint4 a = cs.slot[0][t], b4 = cs.slot[1][t];
                            sl.key = ((uint64_t)a.y << 32) | a.x;
                            sl.cnt = a.z; sl.w0 = a.w; sl.w1 = b4.x; sl.w2 = b4.y; sl.w3 = b4.z; sl.w4 = b4.w;
                            uint64_t i = h & mask;
                            while (sl.key != h && sl.key != kEmptyKey) {   // linear probing past colliding keys (rare)
                                i = (i + 1) & mask;
                                sl = lane::ld_slot(slots + i);
                            }
                            if (sl.key != h) sl.cnt = 0;
                        }
                    }
                    const bool present = active && sl.cnt != 0;
                    const uint32_t amask = (__ballot_sync(0xffffffffu, active) >> g8) & 0xFFu;
                    const uint32_t pmask = (__ballot_sync(0xffffffffu, present) >> g8) & 0xFFu;
                    const uint32_t miss = amask & ~pmask;           // active blocks nobody holds
                    const int limit = miss ? __ffs(miss) - 1 : 8;  // blocks j < limit are walked (plugin.go:214-230)
                    const bool valid = active && j < limit;
                    if (active && j <= limit) n_probes++;
                    // signature of this block's posting set; unused words are zero by construction of the table
                    const uint32_t c = valid ? sl.cnt : 0, w0 = valid ? sl.w0 : 0, w1 = valid ? sl.w1 : 0,
                                   w2 = valid ? sl.w2 : 0, w3 = valid ? sl.w3 : 0, w4 = valid ? sl.w4 : 0;
                    n_postings += c;
                    uint32_t pc = __shfl_up_sync(0xffffffffu, c, 1), p0 = __shfl_up_sync(0xffffffffu, w0, 1),
                             p1 = __shfl_up_sync(0xffffffffu, w1, 1), p2 = __shfl_up_sync(0xffffffffu, w2, 1),
                             p3 = __shfl_up_sync(0xffffffffu, w3, 1), p4 = __shfl_up_sync(0xffffffffu, w4, 1);
                    if (j == 0) {                                  // previous block = last valid block of the previous window
                        pc = cs.carry[0][r]; p0 = cs.carry[1][r]; p1 = cs.carry[2][r];
                        p2 = cs.carry[3][r]; p3 = cs.carry[4][r]; p4 = cs.carry[5][r];
                    }
                    const bool boundary = valid && (c != pc || w0 != p0 || w1 != p1 || w2 != p2 || w3 != p3 || w4 != p4);
                    const uint32_t ball = __ballot_sync(0xffffffffu, boundary);
                    const uint32_t bmask = (ball >> g8) & 0xFFu;
                    if (boundary) {
                        const int idx = cs.nrec[r] + __popc(bmask & ((1u << j) - 1u));
                        if (idx < kMaxRuns) {
                            cs.rec[idx][0][r] = (uint32_t)(kk * kWin + j);
                            cs.rec[idx][1][r] = c;  cs.rec[idx][2][r] = w0; cs.rec[idx][3][r] = w1;
                            cs.rec[idx][4][r] = w2; cs.rec[idx][5][r] = w3; cs.rec[idx][6][r] = w4;
                        } else {
                            cs.overflow[r] = 1;
                        }
                    }
                    if (valid && j == limit - 1) {                 // last walked block of the window: carry its signature
                        cs.carry[0][r] = c;  cs.carry[1][r] = w0; cs.carry[2][r] = w1;
                        cs.carry[3][r] = w2; cs.carry[4][r] = w3; cs.carry[5][r] = w4;
                    }
                    __syncwarp();
                    if (j == 0 && amask) {
                        cs.nrec[r] += __popc(bmask);
                        if (miss) { cs.stopped[r] = 1; cs.stop_b[r] = kk * kWin + limit; }
                    }
                    __syncwarp();
                }
            }
            __threadfence_block();
            bar_arrive(kBarTail, kAll);                            // every window of this tile has been walked
        } else {
            // ================= chain warp: lane = request =================
            const bool mine = lane < TR;
            const int64_t r = r0 + lane;
            const int32_t nfull = mine ? s_nfull[lane] : 0;
            uint64_t prev = 0;
            if (mine && r < p.R) prev = p.seeds[p.model_ids ? p.model_ids[r] : 0];
            for (int k = 0; k < n_win; k++) {
                const int s = k % kStages;
                bar_sync(kBarFull + s, kAll);
                prev = chain_window<TR>(s_m[s], s_nfull, p, r0, k, lane, nfull, lenp8, prev);
                __threadfence_block();
                bar_arrive(kBarHashed + s, kAll);
            }
            // trailing partial block (hashing.go:90-96): generic tail, rare
            uint64_t tail_hash = 0;
            bool has_tail = false;
            if (mine && r < p.R) {
                int64_t eff = s_eff[lane];
                if ((int64_t)nfull * bs < eff) {
                    tail_hash = hash_block_generic(p.data + s_off[lane] + (uint64_t)nfull * (uint64_t)bs,
                                                   eff - (int64_t)nfull * bs, prev);
                    p.hashes[r * (int64_t)p.max_blocks + nfull] = tail_hash;
                    has_tail = true;
                }
            }
            bar_sync(kBarTail, kAll);
            // ---- epilogue, all lanes together: close the runs, count, score, pick (a3-a14)
            if (mine && r < p.R) {
                const uint32_t lo = pk.index.ep_begin, hi = min(pk.index.ep_end, (uint32_t)pk.E);
                lane::Matched m;
                m.n = 0;
                m.overflow = cs.overflow[lane] != 0;
                int32_t nrec = min(cs.nrec[lane], kMaxRuns);
                int32_t end = cs.stopped[lane] ? cs.stop_b[lane] : nfull;     // blocks [0, end) were walked
                int32_t total = nfull;
                // the partial tail block continues the walk when nothing stopped it
                uint32_t tc = 0, t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
                bool tail_valid = false;
                if (has_tail) {
                    total = nfull + 1;
                    if (!cs.stopped[lane]) {
                        n_probes++;
                        Hit th;
                        if (probe(pk.index, tail_hash, th)) {
                            tail_valid = true;
                            tc = th.cnt; t0 = th.w[0];
                            if (tc <= (uint32_t)kInlineIds) { t1 = th.w[1]; t2 = th.w[2]; t3 = th.w[3]; t4 = th.w[4]; }
                            n_postings += tc;
                        }
                    }
                }
                for (int32_t i = 0; i < nrec && !m.overflow; i++) {
                    int32_t start = (int32_t)cs.rec[i][0][lane];
                    int32_t next = i + 1 < nrec ? (int32_t)cs.rec[i + 1][0][lane] : end;
                    lane::flush_run(m, pk.index, (uint32_t)(next - start), cs.rec[i][1][lane], cs.rec[i][2][lane],
                                    cs.rec[i][3][lane], cs.rec[i][4][lane], cs.rec[i][5][lane], cs.rec[i][6][lane], lo, hi);
                }
                if (tail_valid && !m.overflow) lane::flush_run(m, pk.index, 1, tc, t0, t1, t2, t3, t4, lo, hi);
                lane::decide(pk, r, m, total);
            }
        }
        __syncthreads();
    }
    if (pk.work_counters) {
        for (int o = 16; o; o >>= 1) {
            n_probes += __shfl_xor_sync(0xffffffffu, n_probes, o);
            n_postings += __shfl_xor_sync(0xffffffffu, n_postings, o);
        }
        if (lane == 0 && (n_probes | n_postings)) {
            atomicAdd(&pk.work_counters[0], n_probes);
            atomicAdd(&pk.work_counters[1], n_postings);
        }
    }
}

template <typename K, typename... Args>
static cudaError_t launch_persistent_w(K kernel, int tile_r, int threads, int64_t R, int sm_count, cudaStream_t s,
                                       int *occ_cache, Args... args);

template <typename K, typename... Args>
static cudaError_t launch_persistent(K kernel, int tile_r, int64_t R, int sm_count, cudaStream_t s, int *occ_cache,
                                     Args... args) {
    return launch_persistent_w(kernel, tile_r, tile_r * kWin + 32, R, sm_count, s, occ_cache, args...);
}

template <typename K, typename... Args>
static cudaError_t launch_persistent_w(K kernel, int tile_r, int threads, int64_t R, int sm_count, cudaStream_t s,
                                       int *occ_cache, Args... args) {
    int n_tiles = (int)((R + tile_r - 1) / tile_r);
    if (!*occ_cache) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ_cache, kernel, threads, 0);
        if (*occ_cache < 1) *occ_cache = 1;
    }
    if (sm_count <= 0) sm_count = 148;
    int grid = n_tiles < sm_count * *occ_cache ? n_tiles : sm_count * *occ_cache;
    kernel<<<grid, threads, 0, s>>>(args..., n_tiles);
    return cudaGetLastError();
}

// pick == nullptr: hashing only.  Otherwise the whole cycle; decisions go to pick->out, overflowing requests to
// pick->overflow_list (dense-counter kernel).  tile_r: 32 (default) or 16 requests per CTA tile.  16-request tiles
// give a finer-grained last wave (2 048 tiles of 32 over 592 CTA slots quantise to 4 rounds for 3.46 rounds of
// work) but double the chain-warp overhead per digest thread; measured 0.323 ms vs 0.251 ms on config 3.
cudaError_t launch_hash_fused(const HashParams &p, const PickParams *pick, int align, int sm_count, cudaStream_t s,
                              int *launches) {
    if (p.R <= 0) return cudaSuccess;
    static int occ[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int tile_r = p.tile_r == 16 ? 16 : 32;
    const bool a32 = align >= 32;
    cudaError_t e;
    if (pick) {
        if (tile_r == 32)
            e = a32 ? launch_persistent(k_cycle_fused<true, 32>, 32, p.R, sm_count, s, &occ[0], p, *pick)
                    : launch_persistent(k_cycle_fused<false, 32>, 32, p.R, sm_count, s, &occ[1], p, *pick);
        else
            e = a32 ? launch_persistent(k_cycle_fused<true, 16>, 16, p.R, sm_count, s, &occ[2], p, *pick)
                    : launch_persistent(k_cycle_fused<false, 16>, 16, p.R, sm_count, s, &occ[3], p, *pick);
    } else {
        static int occw[2] = {0, 0};
        if (tile_r == 32 && p.wide)
            e = a32 ? launch_persistent(k_hash_wide<true>, 32, p.R, sm_count, s, &occw[0], p)
                    : launch_persistent(k_hash_wide<false>, 32, p.R, sm_count, s, &occw[1], p);
        else if (tile_r == 32 && p.win == 4) {
            static int occ4[2] = {0, 0};
            e = a32 ? launch_persistent_w(k_hash_fused<true, 32, 4, 8>, 32, 160, p.R, sm_count, s, &occ4[0], p)
                    : launch_persistent_w(k_hash_fused<false, 32, 4, 8>, 32, 160, p.R, sm_count, s, &occ4[1], p);
        } else if (tile_r == 32) {
            // Wave balancing: 2 048 tiles of 32 requests over 592 resident CTAs is 3.46 waves -- the last one 46 % full
            // but as long as the others.  Shrinking the tile to `rows` requests (rows of the CTA left idle) makes the
            // tile count a near multiple of the CTA slots: cost ~ waves x (rows + fixed per-tile overhead).
            HashParams q = p;
            int &oc = occ[a32 ? 4 : 5];
            if (!oc) {
                if (a32) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc, k_hash_fused<true, 32>, 288, 0);
                else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc, k_hash_fused<false, 32>, 288, 0);
                if (oc < 1) oc = 1;
            }
            int rows = 32;
            if (p.tile_rows > 0) {
                rows = p.tile_rows < 32 ? p.tile_rows : 32;
            } else if (p.tile_rows == 0) {
                const int64_t slots = (int64_t)(sm_count > 0 ? sm_count : 148) * oc;
                double best = 1e30;
                for (int tr = 32; tr >= 8; tr--) {
                    const int64_t tiles = (p.R + tr - 1) / tr, waves = (tiles + slots - 1) / slots;
                    const double cost = (double)waves * (tr + 4.0);
                    if (cost < best * 0.97) { best = cost; rows = tr; }     // prefer full tiles unless clearly better
                }
            }
            q.tile_rows = rows;
            e = a32 ? launch_persistent_w(k_hash_fused<true, 32>, rows, 288, p.R, sm_count, s, &oc, q)
                    : launch_persistent_w(k_hash_fused<false, 32>, rows, 288, p.R, sm_count, s, &oc, q);
        }
        else
            e = a32 ? launch_persistent(k_hash_fused<true, 16>, 16, p.R, sm_count, s, &occ[6], p)
                    : launch_persistent(k_hash_fused<false, 16>, 16, p.R, sm_count, s, &occ[7], p);
    }
    if (launches) *launches += 1;
    return e;
}

}  // namespace epp
