// cycle_small.cu -- the whole scheduling cycle (a1-a14) of a SMALL host batch as ONE kernel launch that reads the
// prompts straight from the caller's pinned host memory and writes the decisions straight into pinned host memory:
// no copy-engine transfer before or after, no stream synchronisation -- the host polls one flag word per request.
// This is the latency path of the micro-batcher (csrc/batcher.cu): a flush of 1 .. 1024 requests.
//
// One CTA per request, 9 warps:
//   warps 0-7  digest: every thread reads whole 32-byte stripes of "its" blocks over PCIe (zero-copy, 128/256-bit
//              loads) and leaves the merged stripe state of each FULL block in shared memory -- one round trip for a
//              4 096-token prompt (256 blocks, one per thread);
//   warp 8     chain: ONE lane walks the blocks in order (the serial part of hashPrompt, hashing.go:80-96; ~70 ns per
//              block: five dependent 64-bit multiplies), stores every hash to the request's row in HBM (the stash
//              PreRequest needs) and over the block's stripe state in shared memory, and publishes its progress there
//              every 32 blocks (no global fence on the critical path);
//   warp 1     match: after its digests, runs sparse::match_request in FOLLOW mode -- it probes the index 32 blocks at
//              a time as soon as the chain has produced them, so the table and posting-list latencies hide under the
//              chain, and the global-stop rule (plugin.go:219-223) lets it decide a cold prompt after the first chunk
//              while the chain is still hashing.  It then copies the decision to host memory and raises the flag.
// The critical path is the chain (256 blocks x ~75 ns for a 16 KiB prompt) + one PCIe round trip each way.
#include "hash_blocks.cuh"
#include "match_sparse.cuh"

namespace epp {

namespace {
constexpr int kDigestThreads = 256;
constexpr int kThreads = kDigestThreads + 32;

// Hand-over of the hashes from the chain lane to the match warp: plain shared-memory stores published by a release
// store of the progress word, consumed behind an acquire load (CTA scope).  compute-sanitizer's racecheck models
// barriers only and reports this flag protocol as hazards (profiles/r2_sanitizer_racecheck.log); memcheck is clean.
__device__ __forceinline__ void st_release_cta(int32_t *p, int32_t v) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int32_t ld_acquire_cta(const int32_t *p) {
    int32_t v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
struct ChainFollower {
    static constexpr bool kOn = true;
    const int32_t *progress;                   // hashes [0, *progress) are in `h`
    const uint64_t *h;                         // shared memory: the chain overwrites every stripe state with the block's hash
    __device__ __forceinline__ void wait(int32_t n) const {
        while (ld_acquire_cta(progress) < n) {}
    }
    __device__ __forceinline__ uint64_t hash(int32_t i) const { return h[i]; }
};

template <bool kAlign32>
__global__ void __launch_bounds__(kThreads) k_cycle_small(HashParams hp, PickParams pp, SmallOut so) {
    extern __shared__ uint64_t s_m[];              // [max_blocks] merged stripe state of every full block
    __shared__ uint64_t s_off;
    __shared__ int64_t s_eff;
    __shared__ int32_t s_nfull, s_nb;
    __shared__ int32_t s_progress;                 // hashes [0, s_progress) are in s_m (release / acquire, CTA scope)
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int64_t r = blockIdx.x;
    const int64_t bs = hp.block_bytes;

    if (t == 0) {                                  // hashing.go:58-66 (the same arithmetic as tile_lengths, hash_fused.cu)
        uint64_t off, len;
        if (hp.offsets) { off = hp.offsets[r]; len = hp.lengths ? hp.lengths[r] : hp.offsets[r + 1] - off; }
        else { off = (uint64_t)r * hp.uniform_len; len = hp.uniform_len; }
        int64_t eff = (int64_t)len;
        int32_t nfull = 0, nb = 0;
        if (eff < bs) {
            eff = 0;
        } else {
            const int64_t cap = bs * (int64_t)hp.max_blocks;
            if (eff > cap) eff = cap;
            nfull = (int32_t)(eff / bs);
            nb = nfull + ((eff % bs) ? 1 : 0);
        }
        if (hp.in_len) hp.in_len[r] = (int64_t)len;
        hp.nblocks[r] = nb;
        hp.eff_len[r] = eff;
        s_off = off; s_eff = eff; s_nfull = nfull; s_nb = nb;
        s_progress = 0;
        __threadfence();                           // the match warp reads nblocks / in_len at L2
    }
    __syncthreads();
    const int32_t nfull = s_nfull;
    if (warp < kDigestThreads / 32) {
        const uint8_t *base = hp.data + s_off;
        const int n_stripes = (int)(bs >> 5);
        for (int32_t b = t; b < nfull; b += kDigestThreads)
            s_m[b] = block_digest<kAlign32>(base + (uint64_t)b * (uint64_t)bs, n_stripes);
    }
    __syncthreads();                               // the prompt bytes are consumed: nothing below reads host input data

    if (warp == kDigestThreads / 32) {
        if (lane == 0 && s_nb > 0) {
            uint64_t prev = hp.seeds[hp.model_ids ? hp.model_ids[r] : 0];
            uint64_t *row = hp.hashes + r * (int64_t)hp.max_blocks;
            const uint64_t lenp8 = (uint64_t)bs + 8;
            uint64_t *sh = s_m;
            int32_t b = 0;
            for (; b + 32 <= nfull; b += 32) {     // 32 blocks per hand-over, nothing but the dependent multiplies inside
#pragma unroll 8
                for (int j = 0; j < 32; j++) {
                    prev = xxh_chain_step32_lat(s_m[b + j], lenp8, prev);
                    row[b + j] = prev;             // the stash PreRequest reads (stream-ordered, after this kernel)
                    sh[b + j] = prev;              // what the match warp of this CTA reads
                }
                st_release_cta(&s_progress, b + 32);
            }
            for (; b < nfull; b++) {
                prev = xxh_chain_step32_lat(s_m[b], lenp8, prev);
                row[b] = prev;
                sh[b] = prev;
            }
            if ((int64_t)nfull * bs < s_eff) {     // trailing partial block (hashing.go:90-96)
                prev = hash_block_generic(hp.data + s_off + (uint64_t)nfull * (uint64_t)bs, s_eff - (int64_t)nfull * bs, prev);
                row[nfull] = prev;
                sh[nfull] = prev;
            }
            st_release_cta(&s_progress, s_nb);
        }
        return;
    }
    if (warp != 1) return;                         // not warp 0: that one shares its scheduler with the chain warp (8 % 4)

    sparse::Work wk;
    const ChainFollower fw{&s_progress, s_m};
    const bool decided = sparse::match_request<true, false, true, 3, ChainFollower>(pp, r, lane, false, wk, fw);
    if (lane == 0) {
        if (decided) {
            so.dec[r] = pp.out[r];
            so.det[r] = pp.detail[r];
        }
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t *>(so.flags + r) = decided ? so.epoch : (so.epoch | 0x80000000u);
    }
}
}  // namespace

size_t cycle_small_max_blocks() { return 8192; }   // 64 KiB of stripe states per CTA

cudaError_t launch_cycle_small(const HashParams &hp, const PickParams &pp, const SmallOut &so, int align, cudaStream_t s,
                               int *launches) {
    if (hp.R <= 0) return cudaSuccess;
    const size_t smem = sizeof(uint64_t) * (size_t)hp.max_blocks;
    int dev = 0;
    cudaGetDevice(&dev);
    static bool attr_set[64] = {};                 // function attributes are per device: engines of one process may sit on several GPUs
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(k_cycle_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_cycle_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    if (align >= 32) k_cycle_small<true><<<(unsigned)hp.R, kThreads, smem, s>>>(hp, pp, so);
    else k_cycle_small<false><<<(unsigned)hp.R, kThreads, smem, s>>>(hp, pp, so);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace epp
