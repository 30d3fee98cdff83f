// hash_kernels.cu -- a1: batched chained XXH64 prefix-block hashing (approximateprefix/hashing.go:35-99).
//
// v1 layout (correctness-first; the TMA-pipelined fused kernel lives in hash_fused.cu):
//   k_prompt_lengths : per request, truncation + block count                     (hashing.go:58-66)
//   k_block_digests  : one thread per (request, full block), block_bytes % 32 == 0: stripe rounds + merge ->
//                      8-byte pre-chain digest m_b stored in hashes[r][b]         (part A of xxh64.cuh)
//   k_chain          : one thread per request walks the chain in place             (part B)
//   k_hash_generic   : any block size / alignment, one thread per request, fully serial
#include "kernels.h"
#include "xxh64.cuh"

namespace epp {

__device__ __forceinline__ void request_span(const HashParams &p, int64_t r, uint64_t &off, uint64_t &len) {
    if (p.offsets) {
        off = p.offsets[r];
        len = p.lengths ? p.lengths[r] : p.offsets[r + 1] - off;
    } else {
        off = (uint64_t)r * p.uniform_len;
        len = p.uniform_len;
    }
}

__global__ void k_hash_bytes(const uint8_t *msg, size_t len, uint64_t *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint64_t v[4];
    bool have_v = false;
    size_t i = 0;
    if (len >= 32) {
        xxh_init(v);
        have_v = true;
        for (; i + 32 <= len; i += 32) {
            v[0] = xxh_round(v[0], load_le64(msg + i));
            v[1] = xxh_round(v[1], load_le64(msg + i + 8));
            v[2] = xxh_round(v[2], load_le64(msg + i + 16));
            v[3] = xxh_round(v[3], load_le64(msg + i + 24));
        }
    }
    uint8_t tail[32];
    int t = 0;
    for (; i < len; i++) tail[t++] = msg[i];
    *out = xxh_finish(v, have_v, (uint64_t)len, tail, t);
}

cudaError_t launch_hash_bytes(const uint8_t *msg, size_t len, uint64_t *out, cudaStream_t s) {
    k_hash_bytes<<<1, 32, 0, s>>>(msg, len, out);
    return cudaGetLastError();
}

__global__ void k_prompt_lengths(HashParams p) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.R) return;
    uint64_t off, len;
    request_span(p, r, off, len);
    if (p.in_len) p.in_len[r] = (int64_t)len;
    int64_t bs = p.block_bytes;
    int64_t eff = (int64_t)len;
    int32_t nb = 0;
    if (bs <= 0 || eff < bs) {            // hashing.go:51-61 -> nil
        eff = 0;
    } else {
        int64_t cap = bs * (int64_t)p.max_blocks;
        if (eff > cap) eff = cap;         // hashing.go:63-66
        nb = (int32_t)(eff / bs) + ((eff % bs) ? 1 : 0);
    }
    p.nblocks[r] = nb;
    p.eff_len[r] = eff;
}

// One thread per (request, full block).  Requires block_bytes % 32 == 0 and 16-byte aligned block starts.
__global__ void __launch_bounds__(256) k_block_digests(HashParams p) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t r = idx / p.max_blocks;
    int32_t b = (int32_t)(idx % p.max_blocks);
    if (r >= p.R) return;
    int64_t eff = p.eff_len[r];
    int64_t bs = p.block_bytes;
    if ((int64_t)(b + 1) * bs > eff) return;          // not a full block
    uint64_t off, len;
    request_span(p, r, off, len);
    const uint4 *src = reinterpret_cast<const uint4 *>(p.data + off + (uint64_t)b * (uint64_t)bs);
    uint64_t v[4];
    xxh_init(v);
    int ns = (int)(bs / 32);
    for (int s = 0; s < ns; s++) {
        uint4 a = __ldg(src + 2 * s);
        uint4 c = __ldg(src + 2 * s + 1);
        v[0] = xxh_round(v[0], ((uint64_t)a.y << 32) | a.x);
        v[1] = xxh_round(v[1], ((uint64_t)a.w << 32) | a.z);
        v[2] = xxh_round(v[2], ((uint64_t)c.y << 32) | c.x);
        v[3] = xxh_round(v[3], ((uint64_t)c.w << 32) | c.z);
    }
    p.hashes[r * (int64_t)p.max_blocks + b] = xxh_merge_all(v);
}

__global__ void __launch_bounds__(128) k_chain(HashParams p) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.R) return;
    int32_t nb = p.nblocks[r];
    if (nb == 0) return;
    int64_t eff = p.eff_len[r];
    int64_t bs = p.block_bytes;
    int32_t nfull = (int32_t)(eff / bs);
    uint64_t prev = p.seeds[p.model_ids ? p.model_ids[r] : 0];
    uint64_t *row = p.hashes + r * (int64_t)p.max_blocks;
    uint64_t lenp8 = (uint64_t)bs + 8;
    for (int32_t b = 0; b < nfull; b++) {
        prev = xxh_chain_step32(row[b], lenp8, prev);
        row[b] = prev;
    }
    if (nfull < nb) {                                    // trailing partial block (hashing.go:90-96)
        uint64_t off, len;
        request_span(p, r, off, len);
        row[nfull] = hash_block_generic(p.data + off + (uint64_t)nfull * (uint64_t)bs, eff - (int64_t)nfull * bs, prev);
    }
}

__global__ void __launch_bounds__(128) k_hash_generic(HashParams p) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.R) return;
    int32_t nb = p.nblocks[r];
    if (nb == 0) return;
    int64_t eff = p.eff_len[r];
    int64_t bs = p.block_bytes;
    uint64_t off, len;
    request_span(p, r, off, len);
    const uint8_t *base = p.data + off;
    uint64_t prev = p.seeds[p.model_ids ? p.model_ids[r] : 0];
    uint64_t *row = p.hashes + r * (int64_t)p.max_blocks;
    for (int32_t b = 0; b < nb; b++) {
        int64_t start = (int64_t)b * bs;
        int64_t n = eff - start < bs ? eff - start : bs;
        prev = hash_block_generic(base + start, n, prev);
        row[b] = prev;
    }
}

// Common alignment (0, 16 or 32 bytes) of every block start of the batch; the vectorised paths need
// block_bytes % 32 == 0 and alignment >= 16 (128-bit loads) or 32 (256-bit loads, one per XXH64 stripe).
int hash_batch_alignment(const HashParams &p) {
    if (p.block_bytes <= 0 || (p.block_bytes % 32) != 0) return 0;
    uint64_t bits = reinterpret_cast<uintptr_t>(p.data);
    bits |= p.offsets ? p.offsets_or_bits : p.uniform_len;
    if ((bits & 31) == 0) return 32;
    if ((bits & 15) == 0) return 16;
    return 0;
}
static bool fast_path_ok(const HashParams &p) { return hash_batch_alignment(p) >= 16; }

// *flag |= low bits of every offset (device-pointer batches).
__global__ void k_offsets_aligned(const uint64_t *offsets, int64_t n, int *flag) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n && (offsets[r] & 31)) atomicOr(flag, (int)(offsets[r] & 31));
}
cudaError_t launch_check_offsets_aligned(const uint64_t *offsets, int64_t n, int *flag_dev, cudaStream_t s) {
    k_offsets_aligned<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(offsets, n, flag_dev);
    return cudaGetLastError();
}

cudaError_t launch_hash_fused(const HashParams &p, const PickParams *pick, int align, int sm_count, cudaStream_t s,
                              int *launches);

cudaError_t launch_hash_prompts(const HashParams &p, cudaStream_t s, int *launches, cudaEvent_t *ev) {
    if (p.R <= 0) return cudaSuccess;
    int n = 0;
    int align = hash_batch_alignment(p);
    if (align >= 16 && !p.force_v1) {          // one fused kernel: lengths + digests + chain
        if (ev) { cudaEventRecord(ev[0], s); cudaEventRecord(ev[1], s); }
        cudaError_t e = (p.bulk && !p.fused_pick && hash_bulk_supported(p))
                            ? launch_hash_bulk(p, p.sm_count, s, launches)
                            : launch_hash_fused(p, p.fused_pick, align, p.sm_count, s, launches);
        if (ev) { cudaEventRecord(ev[2], s); cudaEventRecord(ev[3], s); }
        return e;
    }
    unsigned gR = (unsigned)((p.R + 127) / 128);
    if (ev) cudaEventRecord(ev[0], s);
    k_prompt_lengths<<<gR, 128, 0, s>>>(p);
    if (ev) cudaEventRecord(ev[1], s);
    n++;
    if (fast_path_ok(p)) {
        int64_t items = p.R * (int64_t)p.max_blocks;
        k_block_digests<<<(unsigned)((items + 255) / 256), 256, 0, s>>>(p);
        if (ev) cudaEventRecord(ev[2], s);
        k_chain<<<gR, 128, 0, s>>>(p);
        if (ev) cudaEventRecord(ev[3], s);
        n += 2;
    } else {
        if (ev) cudaEventRecord(ev[2], s);
        k_hash_generic<<<gR, 128, 0, s>>>(p);
        if (ev) cudaEventRecord(ev[3], s);
        n++;
    }
    if (launches) *launches += n;
    return cudaGetLastError();
}

}  // namespace epp
