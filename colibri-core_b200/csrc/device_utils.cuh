// device_utils.cuh -- small device/host helpers shared by the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace colibri {

constexpr int kWarp = 32;

// ---------------------------------------------------------------------------------------------
// SpookyHash V2, "Short" path (messages < 192 bytes), written for registers: this is the function
// behind Pattern::hash (reference src/pattern.cpp:234-238 -> SpookyHash::Hash64 include/SpookyV2.h:59-66
// -> Hash128 src/SpookyV2.cpp:116-120 -> Short :21-113).  Bob Jenkins' algorithm is public domain.
// The device tables use it both on raw pattern bytes (parity row a5) and on the fixed-width
// (prefix-id, suffix-id) keys that stand in for the byte strings inside the device hash table.
constexpr uint64_t kSpookyConst = 0xdeadbeefdeadbeefULL;

__host__ __device__ __forceinline__ uint64_t rotl64(uint64_t x, int k) {
    return (x << k) | (x >> (64 - k));
}

__host__ __device__ __forceinline__ void spooky_short_mix(uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d) {
    c = rotl64(c, 50); c += d; a ^= c;
    d = rotl64(d, 52); d += a; b ^= d;
    a = rotl64(a, 30); a += b; c ^= a;
    b = rotl64(b, 41); b += c; d ^= b;
    c = rotl64(c, 54); c += d; a ^= c;
    d = rotl64(d, 48); d += a; b ^= d;
    a = rotl64(a, 38); a += b; c ^= a;
    b = rotl64(b, 37); b += c; d ^= b;
    c = rotl64(c, 62); c += d; a ^= c;
    d = rotl64(d, 34); d += a; b ^= d;
    a = rotl64(a, 5);  a += b; c ^= a;
    b = rotl64(b, 36); b += c; d ^= b;
}

__host__ __device__ __forceinline__ void spooky_short_end(uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d) {
    d ^= c; c = rotl64(c, 15); d += c;
    a ^= d; d = rotl64(d, 52); a += d;
    b ^= a; a = rotl64(a, 26); b += a;
    c ^= b; b = rotl64(b, 51); c += b;
    d ^= c; c = rotl64(c, 28); d += c;
    a ^= d; d = rotl64(d, 9);  a += d;
    b ^= a; a = rotl64(a, 47); b += a;
    c ^= b; b = rotl64(b, 54); c += b;
    d ^= c; c = rotl64(c, 32); d += c;
    a ^= d; d = rotl64(d, 25); a += d;
    b ^= a; a = rotl64(a, 63); b += a;
}

// little-endian 64-bit word from up to 8 bytes at p[0..n), zero padded
__host__ __device__ __forceinline__ uint64_t load_le_partial(const uint8_t* p, int n) {
    uint64_t v = 0;
    for (int i = 0; i < n && i < 8; ++i) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

// SpookyHash::Hash64 for len < 192 (byte-addressed, no alignment assumptions)
__host__ __device__ inline uint64_t spooky_hash64(const uint8_t* p, uint32_t len, uint64_t seed) {
    uint64_t a = seed, b = seed, c = kSpookyConst, d = kSpookyConst;
    uint32_t rem = len & 31;
    if (len > 15) {
        uint32_t blocks = len >> 5;
        for (uint32_t i = 0; i < blocks; ++i, p += 32) {
            c += load_le_partial(p, 8);
            d += load_le_partial(p + 8, 8);
            spooky_short_mix(a, b, c, d);
            a += load_le_partial(p + 16, 8);
            b += load_le_partial(p + 24, 8);
        }
        if (rem >= 16) {
            c += load_le_partial(p, 8);
            d += load_le_partial(p + 8, 8);
            spooky_short_mix(a, b, c, d);
            p += 16;
            rem -= 16;
        }
    }
    d += (uint64_t)len << 56;
    if (rem == 0) {
        c += kSpookyConst;
        d += kSpookyConst;
    } else {
        c += load_le_partial(p, rem > 8 ? 8 : (int)rem);
        if (rem > 8) d += load_le_partial(p + 8, (int)rem - 8);
    }
    spooky_short_end(a, b, c, d);
    return a;
}

// Hash64 of one 8-byte little-endian message: the table hash of a (prefix-id, suffix-id) key
__host__ __device__ __forceinline__ uint64_t spooky_hash64_u64(uint64_t msg, uint64_t seed) {
    uint64_t a = seed, b = seed, c = kSpookyConst + msg, d = kSpookyConst + ((uint64_t)8 << 56);
    spooky_short_end(a, b, c, d);
    return a;
}
// Hash64 of one 16-byte message (two little-endian words): the table hash of a skipgram key.
// len == 16 takes the "len > 15, remainder >= 16" branch, then the empty-tail rule.
__host__ __device__ __forceinline__ uint64_t spooky_hash64_u128(uint64_t lo, uint64_t hi, uint64_t seed) {
    uint64_t a = seed, b = seed, c = kSpookyConst + lo, d = kSpookyConst + hi;
    spooky_short_mix(a, b, c, d);
    d += (uint64_t)16 << 56;
    c += kSpookyConst;
    d += kSpookyConst;
    spooky_short_end(a, b, c, d);
    return a;
}

// ---------------------------------------------------------------------------------------------
// class codec (reference src/classencoder.cpp:22-42): little-endian base 128, bit 7 on all but the last byte
__host__ __device__ __forceinline__ uint32_t varint_len(uint32_t cls) {
    return cls < (1u << 7) ? 1u : cls < (1u << 14) ? 2u : cls < (1u << 21) ? 3u : cls < (1u << 28) ? 4u : 5u;
}
__host__ __device__ __forceinline__ uint32_t varint_put(uint8_t* out, uint32_t cls) {
    uint32_t n = 0;
    do {
        uint8_t digit = (uint8_t)(cls & 0x7F);
        cls >>= 7;
        out[n++] = cls ? (uint8_t)(digit | 0x80) : digit;
    } while (cls);
    return n;
}

// map a 64-bit hash onto [0, n) without a modulo (n need not be a power of two)
__device__ __forceinline__ uint64_t fast_range(uint64_t h, uint64_t n) {
    return __umul64hi(h, n);
}

// ---------------------------------------------------------------------------------------------
// warp / block primitives
__device__ __forceinline__ uint32_t lane_id() {
    return threadIdx.x & 31;
}
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= (uint32_t)d) v += t;
    }
    return v;
}
__device__ __forceinline__ uint64_t warp_reduce_sum(uint64_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ uint32_t warp_reduce_max(uint32_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}
// block-wide sum of a u64, result valid in thread 0; `scratch` holds >= blockDim/32 words
__device__ __forceinline__ uint64_t block_reduce_sum(uint64_t v, uint64_t* scratch) {
    v = warp_reduce_sum(v);
    __syncthreads();
    if (lane_id() == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    uint64_t r = 0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? scratch[threadIdx.x] : 0;
        r = warp_reduce_sum(r);
    }
    return r;
}
// one atomicAdd per warp for a compaction cursor; every participating lane gets its own output index
__device__ __forceinline__ uint64_t warp_aggregated_inc(unsigned long long* counter, bool active) {
    uint32_t mask = __ballot_sync(0xffffffffu, active);
    if (!active) return 0;
    int      leader = __ffs(mask) - 1;
    uint64_t base   = 0;
    if ((int)lane_id() == leader) base = atomicAdd(counter, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane_id()) - 1));
}

}  // namespace colibri
