// spooky.h -- SpookyHash V2 "Short" path and the class varint codec, usable from host C++ and from CUDA device code.
// (COLIBRI_HD expands to __host__ __device__ under nvcc and to nothing under a plain C++ compiler.)
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define COLIBRI_HD __host__ __device__
#define COLIBRI_FORCEINLINE __forceinline__
#else
#define COLIBRI_HD
#define COLIBRI_FORCEINLINE inline
#endif

namespace colibri {

// ---------------------------------------------------------------------------------------------
// SpookyHash V2, "Short" path (messages < 192 bytes), written for registers: this is the function
// behind Pattern::hash (reference src/pattern.cpp:234-238 -> SpookyHash::Hash64 include/SpookyV2.h:59-66
// -> Hash128 src/SpookyV2.cpp:116-120 -> Short :21-113).  Bob Jenkins' algorithm is public domain.
// The device tables use it both on raw pattern bytes (parity row a5) and on the fixed-width
// (prefix-id, suffix-id) keys that stand in for the byte strings inside the device hash table.
constexpr uint64_t kSpookyConst = 0xdeadbeefdeadbeefULL;

COLIBRI_HD COLIBRI_FORCEINLINE uint64_t rotl64(uint64_t x, int k) {
    return (x << k) | (x >> (64 - k));
}

COLIBRI_HD COLIBRI_FORCEINLINE void spooky_short_mix(uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d) {
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

COLIBRI_HD COLIBRI_FORCEINLINE void spooky_short_end(uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d) {
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
COLIBRI_HD COLIBRI_FORCEINLINE uint64_t load_le_partial(const uint8_t* p, int n) {
    uint64_t v = 0;
    for (int i = 0; i < n && i < 8; ++i) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

// SpookyHash::Hash64 for len < 192 (byte-addressed, no alignment assumptions)
COLIBRI_HD inline uint64_t spooky_hash64(const uint8_t* p, uint32_t len, uint64_t seed) {
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
COLIBRI_HD COLIBRI_FORCEINLINE uint64_t spooky_hash64_u64(uint64_t msg, uint64_t seed) {
    uint64_t a = seed, b = seed, c = kSpookyConst + msg, d = kSpookyConst + ((uint64_t)8 << 56);
    spooky_short_end(a, b, c, d);
    return a;
}
// Hash64 of one 16-byte message (two little-endian words): the table hash of a skipgram key.
// len == 16 takes the "len > 15, remainder >= 16" branch, then the empty-tail rule.
COLIBRI_HD COLIBRI_FORCEINLINE uint64_t spooky_hash64_u128(uint64_t lo, uint64_t hi, uint64_t seed) {
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
COLIBRI_HD COLIBRI_FORCEINLINE uint32_t varint_len(uint32_t cls) {
    return cls < (1u << 7) ? 1u : cls < (1u << 14) ? 2u : cls < (1u << 21) ? 3u : cls < (1u << 28) ? 4u : 5u;
}
COLIBRI_HD COLIBRI_FORCEINLINE uint32_t varint_put(uint8_t* out, uint32_t cls) {
    uint32_t n = 0;
    do {
        uint8_t digit = (uint8_t)(cls & 0x7F);
        cls >>= 7;
        out[n++] = cls ? (uint8_t)(digit | 0x80) : digit;
    } while (cls);
    return n;
}

}  // namespace colibri
