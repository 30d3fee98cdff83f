// xxh64.cuh -- device-side XXH64 primitives (seed 0 chain of approximateprefix/hashing.go:71-96).
//
// The message hashed for block i is  M_i = block_i_bytes || LE64(h_{i-1})  (hashing.go:82-83), so for a block of
// n bytes the digest splits into
//   (A) the stripe rounds over the floor(n/32) 32-byte stripes made of block bytes only -- independent of the
//       chain, embarrassingly parallel over (request, block);
//   (B) the serial part: [one more stripe if (n mod 32) + 8 >= 32] -> merge -> += len -> tail rounds over the
//       remaining (< 32) bytes, which contain LE64(h_{i-1}) -> avalanche.
// For block sizes that are a multiple of 32 bytes (every multiple of 8 tokens, incl. the default 16) part (B)
// is: m + len, one 8-byte tail round with h_{i-1}, avalanche  (SURVEY.md App. A.1).
#pragma once
#include <stdint.h>

namespace epp {

constexpr uint64_t XP1 = 0x9E3779B185EBCA87ULL;
constexpr uint64_t XP2 = 0xC2B2AE3D27D4EB4FULL;
constexpr uint64_t XP3 = 0x165667B19E3779F9ULL;
constexpr uint64_t XP4 = 0x85EBCA77C2B2AE63ULL;
constexpr uint64_t XP5 = 0x27D4EB2F165667C5ULL;

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) {
    // funnel shifts: 2 SHF per 64-bit rotate
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    uint32_t nlo, nhi;
    if (r < 32) {
        nhi = __funnelshift_l(lo, hi, r);
        nlo = __funnelshift_l(hi, lo, r);
    } else {
        nhi = __funnelshift_l(hi, lo, r - 32);
        nlo = __funnelshift_l(lo, hi, r - 32);
    }
    return ((uint64_t)nhi << 32) | nlo;
}

// a + x * C (mod 2^64) for a compile-time constant C, in exactly three integer multiply-adds: one IMAD.WIDE.U32
// carrying the 64-bit addend and two IMADs folded into the high word.  (The compiler's own expansion of the 64-bit
// multiply spends a fourth instruction on a separate add; XXH64 is ~2 such multiplies per 8 input bytes and the
// hash kernel is issue-bound, so the count matters.)
template <uint64_t C>
__device__ __forceinline__ uint64_t mad64c(uint64_t x, uint64_t a) {
    constexpr uint32_t cl = (uint32_t)C, ch = (uint32_t)(C >> 32);
    const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32);
    uint64_t t;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(xl), "n"(cl), "l"(a));
    uint32_t tl = (uint32_t)t, th = (uint32_t)(t >> 32);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(th) : "r"(xl), "n"(ch));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(th) : "r"(xh), "n"(cl));
    return ((uint64_t)th << 32) | tl;
}

// The same product with a dependency depth of TWO (one wide multiply-add and the two cross products side by side, then
// a three-input add) instead of three chained multiply-adds: one instruction more, a third less latency.  For the
// serial chain step, where one dependent instruction stream per warp is all there is.
template <uint64_t C>
__device__ __forceinline__ uint64_t mad64c_lat(uint64_t x, uint64_t a) {
    constexpr uint32_t cl = (uint32_t)C, ch = (uint32_t)(C >> 32);
    const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32);
    uint64_t t;
    uint32_t u, v;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(xl), "n"(cl), "l"(a));
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(u) : "r"(xl), "n"(ch));
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(v) : "r"(xh), "n"(cl));
    const uint32_t th = (uint32_t)(t >> 32) + u + v;
    return ((uint64_t)th << 32) | (uint32_t)t;
}

__device__ __forceinline__ uint64_t xxh_round(uint64_t acc, uint64_t x) {
    return mad64c<XP1>(rotl64(mad64c<XP2>(x, acc), 31), 0);
}
__device__ __forceinline__ uint64_t xxh_merge(uint64_t h, uint64_t v) {
    return mad64c<XP1>(h ^ xxh_round(0, v), XP4);
}
__device__ __forceinline__ uint64_t xxh_avalanche(uint64_t h) {
    h ^= h >> 33;
    h = mad64c<XP2>(h, 0);
    h ^= h >> 29;
    h = mad64c<XP3>(h, 0);
    h ^= h >> 32;
    return h;
}
__device__ __forceinline__ void xxh_init(uint64_t v[4]) {
    v[0] = XP1 + XP2;
    v[1] = XP2;
    v[2] = 0;
    v[3] = 0ULL - XP1;
}
__device__ __forceinline__ uint64_t xxh_merge_all(const uint64_t v[4]) {
    uint64_t h = rotl64(v[0], 1) + rotl64(v[1], 7) + rotl64(v[2], 12) + rotl64(v[3], 18);
    h = xxh_merge(h, v[0]);
    h = xxh_merge(h, v[1]);
    h = xxh_merge(h, v[2]);
    h = xxh_merge(h, v[3]);
    return h;
}

// Fast chain step for a block whose byte length n is a multiple of 32: m = merged stripe state.
__device__ __forceinline__ uint64_t xxh_chain_step32(uint64_t m, uint64_t len_plus8, uint64_t prev) {
    uint64_t h = m + len_plus8;
    h ^= xxh_round(0, prev);
    h = mad64c<XP1>(rotl64(h, 27), XP4);
    return xxh_avalanche(h);
}

// Latency-optimised variant of the chain step (same value).
__device__ __forceinline__ uint64_t xxh_chain_step32_lat(uint64_t m, uint64_t len_plus8, uint64_t prev) {
    uint64_t h = m + len_plus8;
    h ^= mad64c_lat<XP1>(rotl64(mad64c_lat<XP2>(prev, 0), 31), 0);
    h = mad64c_lat<XP1>(rotl64(h, 27), XP4);
    h ^= h >> 33;
    h = mad64c_lat<XP2>(h, 0);
    h ^= h >> 29;
    h = mad64c_lat<XP3>(h, 0);
    h ^= h >> 32;
    return h;
}

__device__ __forceinline__ uint64_t load_le64(const uint8_t *p) {
    uint64_t x = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) x |= (uint64_t)p[k] << (8 * k);
    return x;
}
__device__ __forceinline__ uint32_t load_le32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// Generic finish: the stripe state v (valid iff have_v) covers `consumed` bytes; `tail`/`tail_len` are the
// remaining message bytes (tail_len < 64), total_len = consumed + tail_len.  Any layout of XXH64 is covered:
// if total_len >= 32 and the tail still holds a full stripe it is consumed first.
__device__ inline uint64_t xxh_finish(uint64_t v[4], bool have_v, uint64_t total_len, const uint8_t *tail,
                                      int tail_len) {
    uint64_t h;
    int p = 0;
    if (total_len >= 32) {
        if (!have_v) xxh_init(v);
        while (p + 32 <= tail_len) {
            v[0] = xxh_round(v[0], load_le64(tail + p));
            v[1] = xxh_round(v[1], load_le64(tail + p + 8));
            v[2] = xxh_round(v[2], load_le64(tail + p + 16));
            v[3] = xxh_round(v[3], load_le64(tail + p + 24));
            p += 32;
        }
        h = xxh_merge_all(v);
    } else {
        h = XP5;
    }
    h += total_len;
    while (p + 8 <= tail_len) {
        h ^= xxh_round(0, load_le64(tail + p));
        h = rotl64(h, 27) * XP1 + XP4;
        p += 8;
    }
    if (p + 4 <= tail_len) {
        h ^= (uint64_t)load_le32(tail + p) * XP1;
        h = rotl64(h, 23) * XP2 + XP3;
        p += 4;
    }
    while (p < tail_len) {
        h ^= (uint64_t)tail[p] * XP5;
        h = rotl64(h, 11) * XP1;
        p++;
    }
    return xxh_avalanche(h);
}

// Generic hash of one chain block: message = bytes[0..n) || LE64(prev)  (hashing.go:82-83), any n >= 0, any
// alignment (byte loads).
__device__ inline uint64_t hash_block_generic(const uint8_t *b, int64_t n, uint64_t prev) {
    uint64_t v[4];
    bool have_v = false;
    int64_t i = 0;
    if (n >= 32) {
        xxh_init(v);
        have_v = true;
        for (; i + 32 <= n; i += 32) {
            v[0] = xxh_round(v[0], load_le64(b + i));
            v[1] = xxh_round(v[1], load_le64(b + i + 8));
            v[2] = xxh_round(v[2], load_le64(b + i + 16));
            v[3] = xxh_round(v[3], load_le64(b + i + 24));
        }
    }
    uint8_t tail[40];
    int t = 0;
    for (; i < n; i++) tail[t++] = b[i];
#pragma unroll
    for (int k = 0; k < 8; k++) tail[t++] = (uint8_t)(prev >> (8 * k));
    return xxh_finish(v, have_v, (uint64_t)n + 8, tail, t);
}

}  // namespace epp
