// device_utils.cuh -- small device/host helpers shared by the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "spooky.h"

namespace colibri {

// map a 64-bit hash onto [0, n) without a modulo (n need not be a power of two)
__device__ __forceinline__ uint64_t fast_range(uint64_t h, uint64_t n) {
    return __umul64hi(h, n);
}

// ---------------------------------------------------------------------------------------------
// Placement hash of the fixed-width id keys (the (prefix-id, suffix-id) pair of an n-gram, the 16-byte skipgram key).  Its value is
// not observable in any result (only iteration order depends on the hash), so it only has to spread keys evenly.
// Round 1 used SpookyV2 Hash64 here as well: ~75 integer instructions per key, executed by whole warps even when only a few lanes
// hold a valid window -- the filter and the sparse levels 4/5 were ALU-bound on it (profiles/r02_ncu.md: 140 warp instructions per
// position in ngram_filter_kernel, ALU pipe 52 %).  This is the 64-bit finaliser of MurmurHash3 (public domain): two multiplies,
// three xor-shifts, ~18 instructions.  SpookyV2 stays where the key IS the pattern's bytes (pattern_index.cu: constrained training,
// load, lookups -- the hash the reference computes, Pattern::hash).  Build with -DCOLIBRI_TABLE_HASH_SPOOKY to get round 1's choice back.
__host__ __device__ __forceinline__ uint64_t fmix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xFF51AFD7ED558CCDull;
    x ^= x >> 33;
    x *= 0xC4CEB9FE1A85EC53ull;
    x ^= x >> 33;
    return x;
}
__host__ __device__ __forceinline__ uint64_t table_hash_u64(uint64_t key) {
#ifdef COLIBRI_TABLE_HASH_SPOOKY
    return spooky_hash64_u64(key, 0);
#else
    return fmix64(key);
#endif
}
__host__ __device__ __forceinline__ uint64_t table_hash_u128(uint64_t k0, uint64_t k1) {
#ifdef COLIBRI_TABLE_HASH_SPOOKY
    return spooky_hash64_u128(k0, k1, 0);
#else
    return fmix64(k0 + 0x9E3779B97F4A7C15ull * fmix64(k1 ^ 0xD6E8FEB86659FD93ull));
#endif
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
