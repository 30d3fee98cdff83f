// hash_bulk.cu -- a1 (hashPrompt, approximateprefix/hashing.go:35-99) with the prompt bytes streamed into shared
// memory by the bulk-copy engine (cp.async.bulk + mbarrier complete_tx) instead of per-thread global loads.
//
// Why: k_hash_fused sits at ~53 % issue utilisation with `long_scoreboard` (the 256-bit global loads) as its top
// stall -- the digest warps both fetch and hash, so their registers and their place in the scheduler are what keeps
// bytes in flight.  Here a producer warp keeps DS stages of 16.5 KiB per CTA in flight with one instruction per
// request row, and the digest warps only ever wait on shared memory.
//
// CTA = 10 warps, tile = 32 requests, window = 8 blocks:
//   warp 9 (producer): lane = request.  Per window, one cp.async.bulk of the request's next 8 blocks (512 B for the
//       default 64-byte block) into row `lane` of the stage; rows are 16 bytes longer than the copy so that the
//       digest warps' 128-bit shared loads (lane = request, stride = row pitch) hit 8 different 16-byte bank groups
//       per quarter-warp -- conflict-free without swizzling.
//   warps 0-7 (digest): warp w = block w of the window, lane = request: four independent XXH64 accumulator chains
//       over the block's 32-byte stripes read from shared memory, merge -> m into the m-ring [request][block].
//   warp 8 (chain): lane = request; identical to k_hash_fused (serial m -> h chain, 64-byte hash segments out).
//   data ring: mbarrier full (tx bytes) / empty (8 digest warps); m ring: named barriers as in k_hash_fused.
#include "kernels.h"
#include "xxh64.cuh"

namespace epp {

namespace {
constexpr int kTR = 32;
constexpr int kWin = 8;
constexpr int kMStages = 4;
constexpr int kPitch = kWin + 1;
constexpr int kBarFull = 1;
constexpr int kBarEmpty = 1 + kMStages;
constexpr int kChainGroup = kWin * 32 + 32;     // digest warps + chain warp on the m-ring barriers
constexpr int kThreads = kWin * 32 + 64;        // + producer warp

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *b, uint32_t tx) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(tx) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(b)), "r"(parity)
            : "memory");
    } while (!ok);
}
// global -> shared bulk copy; completion is signalled on `bar` as `bytes` of transaction count.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void lds_stripe(uint32_t addr, uint64_t x[4]) {
    asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(x[0]), "=l"(x[1]) : "r"(addr));
    asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(x[2]), "=l"(x[3]) : "r"(addr + 16));
}
}  // namespace

template <int DS, int MINCTA>
__global__ void __launch_bounds__(kThreads, MINCTA) k_hash_bulk(HashParams p, int n_tiles, int row_pitch) {
    extern __shared__ __align__(128) uint8_t s_data[];            // [DS][kTR][row_pitch]
    __shared__ uint64_t s_m[kMStages][kTR][kPitch];
    __shared__ __align__(8) uint64_t s_full[DS], s_empty[DS];
    __shared__ uint64_t s_off[kTR];
    __shared__ int64_t s_eff[kTR];
    __shared__ int32_t s_nfull[kTR];
    __shared__ int32_t s_maxfull;

    const int lane = threadIdx.x & 31, hw_warp = threadIdx.x >> 5;
    // see k_hash_fused: chain warp rotated over the SM sub-partitions per resident CTA (roles: 0-7 digest, 8 chain, 9 producer)
    const int chain_at = p.chain_spread ? (int)((blockIdx.x / (unsigned)max(p.sm_count, 1)) & 3u) : kWin;
    const int warp = hw_warp == kWin + 1 ? kWin + 1 : (hw_warp == chain_at ? kWin : (hw_warp > chain_at ? hw_warp - 1 : hw_warp));
    const int t = threadIdx.x;
    const int64_t bs = p.block_bytes;
    const int n_stripes = (int)(bs >> 5);
    const uint64_t lenp8 = (uint64_t)bs + 8;
    const uint32_t stage_bytes = (uint32_t)kTR * (uint32_t)row_pitch;

    if (t == 0) {
        for (int s = 0; s < DS; s++) {
            mbar_init(&s_full[s], 1);          // the producer's arrive.expect_tx
            mbar_init(&s_empty[s], kWin);      // one arrive per digest warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    uint32_t it = 0;                            // windows this CTA has pushed through the data ring so far
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t r0 = (int64_t)tile * kTR;
        if (t < 32) {
            // per-request lengths (hashing.go:58-66)
            int64_t r = r0 + t;
            uint64_t off = 0;
            int64_t eff = 0;
            int32_t nfull = 0, nb = 0;
            if (r < p.R) {
                uint64_t len;
                if (p.offsets) { off = p.offsets[r]; len = p.lengths ? p.lengths[r] : p.offsets[r + 1] - off; }
                else { off = (uint64_t)r * p.uniform_len; len = p.uniform_len; }
                if (p.in_len) p.in_len[r] = (int64_t)len;
                eff = (int64_t)len;
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

        if (warp == kWin + 1) {
            // ================= producer: one bulk copy per request row per window =================
            const int32_t nfull = s_nfull[lane];
            const uint8_t *src = p.data + s_off[lane];
            for (int k = 0; k < n_win; k++) {
                const uint32_t w = it + (uint32_t)k, s = w % DS;
                if (w >= (uint32_t)DS) mbar_wait(&s_empty[s], ((w / DS) - 1) & 1);
                int nblk = nfull - k * kWin;
                nblk = nblk < 0 ? 0 : (nblk > kWin ? kWin : nblk);
                const uint32_t bytes = (uint32_t)nblk * (uint32_t)bs;
                uint32_t total = bytes;
                for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
                if (lane == 0) mbar_arrive_expect_tx(&s_full[s], total);
                __syncwarp();
                if (bytes)
                    bulk_g2s(s_data + (size_t)s * stage_bytes + (size_t)lane * (size_t)row_pitch,
                             src + (uint64_t)k * (uint64_t)(kWin * bs), bytes, &s_full[s]);
            }
        } else if (warp < kWin) {
            // ================= digest: warp = block of the window, lane = request =================
            const int32_t nfull = s_nfull[lane];
            const uint32_t row = smem_u32(s_data) + (uint32_t)lane * (uint32_t)row_pitch + (uint32_t)warp * (uint32_t)bs;
            for (int k = 0; k < n_win; k++) {
                const uint32_t w = it + (uint32_t)k, s = w % DS;
                const int s2 = k % kMStages;
                const bool active = k * kWin + warp < nfull;
                mbar_wait(&s_full[s], (w / DS) & 1);
                uint64_t m = 0;
                if (n_stripes == 2) {                              // the default 64-byte block
                    uint64_t x0[4], x1[4];
                    if (active) {
                        lds_stripe(row + s * stage_bytes, x0);
                        lds_stripe(row + s * stage_bytes + 32, x1);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_empty[s]);       // bytes are in registers: the slot can refill
                    if (active) {
                        uint64_t v[4];
                        xxh_init(v);
#pragma unroll
                        for (int q = 0; q < 4; q++) v[q] = xxh_round(v[q], x0[q]);
#pragma unroll
                        for (int q = 0; q < 4; q++) v[q] = xxh_round(v[q], x1[q]);
                        m = xxh_merge_all(v);
                    }
                } else {
                    if (active) {
                        uint64_t v[4];
                        xxh_init(v);
                        for (int st = 0; st < n_stripes; st++) {
                            uint64_t x[4];
                            lds_stripe(row + s * stage_bytes + 32 * st, x);
#pragma unroll
                            for (int q = 0; q < 4; q++) v[q] = xxh_round(v[q], x[q]);
                        }
                        m = xxh_merge_all(v);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_empty[s]);
                }
                if (k >= kMStages) bar_sync(kBarEmpty + s2, kChainGroup);
                if (active) s_m[s2][lane][warp] = m;
                __threadfence_block();
                bar_arrive(kBarFull + s2, kChainGroup);
            }
            int first = n_win > kMStages ? n_win - kMStages : 0;
            for (int k = first; k < n_win; k++) bar_sync(kBarEmpty + (k % kMStages), kChainGroup);
        } else {
            // ================= chain: lane = request (as in k_hash_fused) =================
            const int64_t r = r0 + lane;
            const int32_t nfull = s_nfull[lane];
            uint64_t prev = 0;
            if (r < p.R) prev = p.seeds[p.model_ids ? p.model_ids[r] : 0];
            for (int k = 0; k < n_win; k++) {
                const int s2 = k % kMStages;
                bar_sync(kBarFull + s2, kChainGroup);
#pragma unroll
                for (int j = 0; j < kWin; j++) {
                    if (k * kWin + j < nfull) {
                        prev = xxh_chain_step32(s_m[s2][lane][j], lenp8, prev);
                        s_m[s2][lane][j] = prev;
                    }
                }
                __syncwarp();
#pragma unroll
                for (int q = 0; q < kTR / 4; q++) {
                    const int rr = q * 4 + (lane >> 3), jj = lane & 7;
                    const int b = k * kWin + jj;
                    if (b < s_nfull[rr]) p.hashes[(r0 + rr) * (int64_t)p.max_blocks + b] = s_m[s2][rr][jj];
                }
                __syncwarp();
                bar_arrive(kBarEmpty + s2, kChainGroup);
            }
            if (r < p.R) {                           // trailing partial block (hashing.go:90-96): generic tail, rare
                int64_t eff = s_eff[lane];
                if ((int64_t)nfull * bs < eff)
                    p.hashes[r * (int64_t)p.max_blocks + nfull] = hash_block_generic(
                        p.data + s_off[lane] + (uint64_t)nfull * (uint64_t)bs, eff - (int64_t)nfull * bs, prev);
            }
        }
        it += (uint32_t)n_win;
        __syncthreads();
    }
}

template <int DS, int MINCTA>
static cudaError_t launch_variant(const HashParams &p, int sm_count, cudaStream_t s) {
    static int occ_for[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};            // per block_bytes / 32
    static size_t dyn_set = 0;
    const int row_pitch = kWin * p.block_bytes + 16;
    const size_t dyn = (size_t)DS * kTR * (size_t)row_pitch;
    auto kernel = k_hash_bulk<DS, MINCTA>;
    if (dyn > dyn_set) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return e;
        dyn_set = dyn;
    }
    int &occ = occ_for[p.block_bytes / 32];
    if (!occ) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, dyn);
        if (occ < 1) occ = 1;
    }
    const int n_tiles = (int)((p.R + kTR - 1) / kTR);
    if (sm_count <= 0) sm_count = 148;
    const int grid = n_tiles < sm_count * occ ? n_tiles : sm_count * occ;
    kernel<<<grid, kThreads, dyn, s>>>(p, n_tiles, row_pitch);
    return cudaGetLastError();
}

// p.bulk: 2 / 3 / 4 = data stages per CTA.  Needs 16-byte aligned prompts and block_bytes % 32 == 0 (the caller checks);
// blocks larger than 256 bytes would need more shared memory than the occupancy targets allow -> not handled here.
bool hash_bulk_supported(const HashParams &p) {
    // 4 stages of 32 rows of (8 blocks + 16 bytes) must fit the opt-in shared memory next to the static arrays
    return p.block_bytes % 32 == 0 && p.block_bytes <= 192;
}

cudaError_t launch_hash_bulk(const HashParams &p, int sm_count, cudaStream_t s, int *launches) {
    if (p.R <= 0) return cudaSuccess;
    cudaError_t e;
    switch (p.bulk) {
        case 2: e = launch_variant<2, 4>(p, sm_count, s); break;
        case 4: e = launch_variant<4, 2>(p, sm_count, s); break;
        case 5: e = launch_variant<2, 5>(p, sm_count, s); break;
        default: e = launch_variant<3, 3>(p, sm_count, s); break;
    }
    if (launches) *launches += 1;
    return e;
}

}  // namespace epp
