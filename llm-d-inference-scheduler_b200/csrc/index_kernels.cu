// index_kernels.cu -- a2: the GPU-resident prefix-block index (inverted map  blockHash -> {endpoints}),
// the device counterpart of indexer.hashToPods (approximateprefix/indexer.go:32-37).
//
// Layout in HBM: one open-addressed table of 16-byte slots {key u64, posting offset u32, count u32}
// (one LDG.128 per probe), linear probing from key & mask (XXH64 output is already uniformly mixed),
// load factor <= 0.5; postings are u32 endpoint slot ids, contiguous per key.
//
// Build (from a snapshot of (hash, endpoint) pairs, bulk, on the device):
//   k_index_clear -> k_index_insert (atomicCAS claim + count) -> k_index_alloc (posting ranges by atomic
//   cursor) -> k_index_fill (scatter endpoint ids) -> k_index_dedupe (set semantics of podSet).
#include "kernels.h"

namespace epp {

__global__ void k_index_clear(IndexSlot *slots, uint32_t *fill, uint64_t capacity, uint32_t *cursor,
                              IndexSlot *special) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < capacity) {
        slots[i].key = kEmptyKey;
        slots[i].off = 0;
        slots[i].cnt = 0;
        fill[i] = 0;
    }
    if (i == 0) {
        cursor[0] = 0;   // posting cursor
        cursor[1] = 0;   // special fill
        special->key = kEmptyKey;
        special->off = 0;
        special->cnt = 0;
    }
}

__device__ __forceinline__ uint64_t find_or_claim(IndexSlot *slots, uint64_t mask, uint64_t key) {
    uint64_t i = key & mask;
    for (;;) {
        unsigned long long *kp = reinterpret_cast<unsigned long long *>(&slots[i].key);
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(kp);
        if (cur == key) return i;
        if (cur == kEmptyKey) {
            unsigned long long old = atomicCAS(kp, (unsigned long long)kEmptyKey, (unsigned long long)key);
            if (old == kEmptyKey || old == key) return i;
        }
        i = (i + 1) & mask;
    }
}

__global__ void k_index_insert(const uint64_t *pair_hash, uint64_t n, IndexSlot *slots, uint64_t mask,
                               IndexSlot *special) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint64_t key = pair_hash[t];
    if (key == kEmptyKey) {
        atomicAdd(&special->cnt, 1u);
        return;
    }
    uint64_t i = find_or_claim(slots, mask, key);
    atomicAdd(&slots[i].cnt, 1u);
}

__global__ void k_index_alloc(IndexSlot *slots, uint64_t capacity, uint32_t *cursor, IndexSlot *special) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < capacity && slots[i].cnt) slots[i].off = atomicAdd(&cursor[0], slots[i].cnt);
    if (i == 0 && special->cnt) special->off = atomicAdd(&cursor[0], special->cnt);
}

__device__ __forceinline__ uint64_t find_slot(const IndexSlot *slots, uint64_t mask, uint64_t key) {
    uint64_t i = key & mask;
    while (slots[i].key != key) i = (i + 1) & mask;   // present by construction
    return i;
}

__global__ void k_index_fill(const uint64_t *pair_hash, const uint32_t *pair_ep, uint64_t n, IndexSlot *slots,
                             uint64_t mask, uint32_t *fill, uint32_t *postings, uint32_t *cursor,
                             IndexSlot *special) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint64_t key = pair_hash[t];
    if (key == kEmptyKey) {
        uint32_t pos = atomicAdd(&cursor[1], 1u);
        postings[special->off + pos] = pair_ep[t];
        return;
    }
    uint64_t i = find_slot(slots, mask, key);
    uint32_t pos = atomicAdd(&fill[i], 1u);
    postings[slots[i].off + pos] = pair_ep[t];
}

__device__ inline uint32_t dedupe_list(uint32_t *list, uint32_t cnt) {
    uint32_t w = 0;
    for (uint32_t a = 0; a < cnt; a++) {
        uint32_t e = list[a];
        bool dup = false;
        for (uint32_t b = 0; b < w; b++)
            if (list[b] == e) { dup = true; break; }
        if (!dup) list[w++] = e;
    }
    return w;
}

// podSet is a SET (indexer.go:78-82): drop duplicate (hash, endpoint) pairs; also drops ids >= max_endpoints
// is NOT done (servers outside the pool keep the walk alive, plugin.go:219-228) -- ids are kept verbatim.
__global__ void k_index_dedupe(IndexSlot *slots, uint64_t capacity, uint32_t *postings, IndexSlot *special) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < capacity && slots[i].cnt > 1) slots[i].cnt = dedupe_list(postings + slots[i].off, slots[i].cnt);
    if (i == 0 && special->cnt > 1) special->cnt = dedupe_list(postings + special->off, special->cnt);
}

cudaError_t launch_index_build(const uint64_t *pair_hash, const uint32_t *pair_ep, uint64_t n, IndexSlot *slots,
                               uint64_t capacity, uint32_t *postings, uint32_t *scratch, uint32_t *cursor,
                               IndexSlot *special_dev, uint32_t max_endpoints, cudaStream_t s, int *launches) {
    (void)max_endpoints;
    unsigned gc = (unsigned)((capacity + 255) / 256);
    if (gc == 0) gc = 1;
    k_index_clear<<<gc, 256, 0, s>>>(slots, scratch, capacity, cursor, special_dev);
    int nl = 1;
    if (n > 0) {
        unsigned gn = (unsigned)((n + 255) / 256);
        uint64_t mask = capacity - 1;
        k_index_insert<<<gn, 256, 0, s>>>(pair_hash, n, slots, mask, special_dev);
        k_index_alloc<<<gc, 256, 0, s>>>(slots, capacity, cursor, special_dev);
        k_index_fill<<<gn, 256, 0, s>>>(pair_hash, pair_ep, n, slots, mask, scratch, postings, cursor, special_dev);
        k_index_dedupe<<<gc, 256, 0, s>>>(slots, capacity, postings, special_dev);
        nl += 4;
    }
    if (launches) *launches += nl;
    return cudaGetLastError();
}

__global__ void k_index_get(IndexView ix, uint64_t hash, uint32_t *out_eps, int32_t cap, int32_t *out_n) {
    if (threadIdx.x || blockIdx.x) return;
    uint32_t off = 0, cnt = 0;
    if (hash == kEmptyKey) {
        off = ix.special.off;
        cnt = ix.special.cnt;
    } else if (ix.slots) {
        uint64_t i = hash & ix.mask;
        for (;;) {
            IndexSlot sl = ix.slots[i];
            if (sl.cnt == 0) break;
            if (sl.key == hash) { off = sl.off; cnt = sl.cnt; break; }
            i = (i + 1) & ix.mask;
        }
    }
    for (uint32_t k = 0; k < cnt && (int32_t)k < cap; k++) out_eps[k] = ix.postings[off + k];
    *out_n = (int32_t)cnt;
}

cudaError_t launch_index_get(const IndexView &ix, uint64_t hash, uint32_t *out_eps, int32_t cap, int32_t *out_n,
                             cudaStream_t s) {
    k_index_get<<<1, 32, 0, s>>>(ix, hash, out_eps, cap, out_n);
    return cudaGetLastError();
}

}  // namespace epp
