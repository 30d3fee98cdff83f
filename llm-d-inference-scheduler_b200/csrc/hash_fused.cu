// hash_fused.cu -- a1 fast path: prompt lengths + stripe digests + chain in ONE kernel (block_bytes % 32 == 0,
// 16-byte aligned prompts).  Every prompt byte is read from HBM exactly once (coalesced 128-bit loads), the
// 8-byte pre-chain digests never leave the SM (shared-memory ring), and the only global write is the final hash.
//
// CTA = 9 warps, tile = 32 requests:
//   warps 0-7 (digest): thread (r = t/8, j = t%8) owns block k*8+j of request r in window k: 4 independent XXH64
//       accumulator chains over the block's 32-byte stripes, merge -> m, stored to the stage's [r][j] cell.
//   warp 8 (chain): lane = request.  Walks the window's 8 blocks in order (the only serial part of the digest:
//       m + len, one 8-byte round with h_{i-1}, avalanche), in place, then writes the 8 hashes of each request as
//       one 64-byte segment.  32 chains advance per instruction, so the serial dependency costs issue slots for one
//       warp only while the 8 digest warps keep the memory pipe busy.
//   Stages are handed over with named barriers (bar.arrive / bar.sync), 4-deep ring.
#include "kernels.h"
#include "xxh64.cuh"

namespace epp {

namespace {
constexpr int kTileR = 32;
constexpr int kWin = 8;
constexpr int kDigestWarps = 8;
constexpr int kDigestThreads = kDigestWarps * 32;
constexpr int kThreads = kDigestThreads + 32;
constexpr int kStages = 4;
constexpr int kPitch = kWin + 1;          // u64 cells per request row (odd pitch: conflict-free lane = request reads)
constexpr int kBarFull = 1;               // named barrier ids 1..4
constexpr int kBarEmpty = 1 + kStages;    // 5..8

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

}  // namespace

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

template <bool kAlign32>
__global__ void __launch_bounds__(kThreads) k_hash_fused(HashParams p, int n_tiles) {
    __shared__ uint64_t s_m[kStages][kTileR][kPitch];
    __shared__ uint64_t s_off[kTileR];
    __shared__ int64_t s_eff[kTileR];
    __shared__ int32_t s_nfull[kTileR];
    __shared__ int32_t s_maxfull;

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
                if (p.offsets) { off = p.offsets[r]; len = p.offsets[r + 1] - off; }
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
                if (k >= kStages) bar_sync(kBarEmpty + s, kThreads);
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
                bar_arrive(kBarFull + s, kThreads);
            }
        } else {
            // ================= chain warp: lane = request =================
            const int64_t r = r0 + lane;
            const int32_t nfull = s_nfull[lane];
            uint64_t prev = 0;
            if (r < p.R) prev = p.seeds[p.model_ids ? p.model_ids[r] : 0];
            for (int k = 0; k < n_win; k++) {
                const int s = k % kStages;
                bar_sync(kBarFull + s, kThreads);
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
                bar_arrive(kBarEmpty + s, kThreads);
            }
            // trailing partial block (hashing.go:90-96): generic tail, rare
            if (r < p.R) {
                int64_t eff = s_eff[lane];
                if ((int64_t)nfull * bs < eff)
                    p.hashes[r * (int64_t)p.max_blocks + nfull] =
                        hash_block_generic(p.data + s_off[lane] + (uint64_t)nfull * (uint64_t)bs, eff - (int64_t)nfull * bs, prev);
            }
        }
        // drain: outstanding empty-barrier arrivals of the last min(n_win, kStages) windows must be consumed before
        // the ring is reused by the next tile
        if (warp < kDigestWarps) {
            int first = n_win > kStages ? n_win - kStages : 0;
            for (int k = first; k < n_win; k++) bar_sync(kBarEmpty + (k % kStages), kThreads);
        }
        __syncthreads();
    }
}

cudaError_t launch_hash_fused(const HashParams &p, int align, int sm_count, cudaStream_t s, int *launches) {
    if (p.R <= 0) return cudaSuccess;
    int n_tiles = (int)((p.R + kTileR - 1) / kTileR);
    static int occ = 0;
    if (!occ) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_hash_fused<true>, kThreads, 0);
        if (occ < 1) occ = 1;
    }
    if (sm_count <= 0) sm_count = 148;
    int grid = n_tiles < sm_count * occ ? n_tiles : sm_count * occ;
    if (align >= 32) k_hash_fused<true><<<grid, kThreads, 0, s>>>(p, n_tiles);
    else k_hash_fused<false><<<grid, kThreads, 0, s>>>(p, n_tiles);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace epp
