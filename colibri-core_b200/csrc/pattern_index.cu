// pattern_index.cu -- a pattern set resident in HBM, addressed by Pattern::hash of the pattern BYTES.
//
// Reference: every membership test of the reference is `unordered_map<Pattern,...>::find` keyed by SpookyV2 of the byte string
// (src/pattern.cpp:234-238, include/patternstore.h:954-959): constrainbymodel->has(window) during constrained training
// (include/patternmodel.h:1088-1089), constrainstore->has(p) while loading (include/patternstore.h:586), has()/occurrencecount()
// queries (:751-756, :1653-1669).  Here a model's patterns are a flat blob + offsets; the index over them is an open-addressing
// table of 32-byte slots = one HBM sector {pattern index + 1, counter, the key bytes themselves (up to 23), key length}, placed by
// the low bits of SpookyV2 Hash64 of the bytes (the same function as Pattern::hash, csrc/spooky.h).  A probe compares the key in
// full inside the sector it just fetched (longer keys: in the blob), and constrained training counts in that same sector, so a
// window that matches costs ONE random sector instead of four (slot, two offsets, key bytes, counter: 16.8 GB of DRAM reads for the
// 68 M matching bigram windows of the 100 M-token corpus before, profiles/r01b).
//   * constrained_match_kernel: one thread per corpus position and window length n: re-encode the n class ids to their
//     varint bytes in registers, hash, probe, compare, atomicAdd the pattern's counter; optionally remember the match per
//     position (the forward index of indexed models is built from those with index.cu's ordered pairs + stable radix sort).
//   * index_lookup_kernel: batch queries (has / occurrencecount / load-time constraint).
//   * compaction kernels: threshold -> survivors, ordered by pattern length so that per-length occurrence lists concatenate.
// All of it is HBM-bound integer/byte work (random 8-byte probes + short byte compares).
#include <algorithm>
#include <cstdlib>

#include "device_utils.cuh"
#include "kernels.h"

namespace colibri {

static inline unsigned pi_div_up(uint64_t a, uint64_t b) {
    return (unsigned)((a + b - 1) / b);
}

// ---------------------------------------------------------------------------------------------
// key bytes packed little-endian into 64-bit words (zero padded): the form SpookyV2 Short consumes
template <int NW>
__device__ __forceinline__ void words_clear(uint64_t (&w)[NW]) {
#pragma unroll
    for (int i = 0; i < NW; ++i) w[i] = 0;
}
template <int NW>
__device__ __forceinline__ void words_put(uint64_t (&w)[NW], uint32_t at, uint8_t byte) {
    if (at < 8u * NW) w[at >> 3] |= (uint64_t)byte << (8 * (at & 7));
}
template <int NW>
__device__ __forceinline__ uint8_t words_get(const uint64_t (&w)[NW], uint32_t at) {
    return (uint8_t)(w[at >> 3] >> (8 * (at & 7)));
}
// SpookyHash::Hash64(bytes, len, 0) for len < 192 over the packed words (same arithmetic as spooky_hash64 in spooky.h)
template <int NW>
__device__ __forceinline__ uint64_t spooky_words(const uint64_t (&w)[NW], uint32_t len) {
    uint64_t a = 0, b = 0, c = kSpookyConst, d = kSpookyConst;
    uint32_t rem = len & 31;
    int      wi  = 0;
    if (len > 15) {
        const uint32_t blocks = len >> 5;
        for (uint32_t i = 0; i < blocks; ++i, wi += 4) {
            c += w[wi];
            d += w[wi + 1];
            spooky_short_mix(a, b, c, d);
            a += w[wi + 2];
            b += w[wi + 3];
        }
        if (rem >= 16) {
            c += w[wi];
            d += w[wi + 1];
            spooky_short_mix(a, b, c, d);
            wi += 2;
            rem -= 16;
        }
    }
    d += (uint64_t)len << 56;
    if (rem == 0) {
        c += kSpookyConst;
        d += kSpookyConst;
    } else {
        c += w[wi];
        if (rem > 8) d += w[wi + 1];
    }
    spooky_short_end(a, b, c, d);
    return a;
}
template <int NW>
__device__ __forceinline__ void words_load(uint64_t (&w)[NW], const uint8_t* __restrict__ p, uint32_t len) {
    words_clear(w);
    for (uint32_t i = 0; i < len; ++i) words_put(w, i, p[i]);
}
// probe the index for the key held in w[0..len): pattern index + 1, or 0
// The presence bitmap: one bit per hash bucket, 16 buckets per pattern, small enough to stay in L2 (11 M patterns: 32 MB).  Most
// windows of a corpus are NOT in the set; they are turned away here by an L2 hit instead of a random HBM sector.
__device__ __forceinline__ uint64_t presence_bit(uint64_t h, uint64_t pmask) {
    return (h >> 13) & pmask;  // bits 13.. (the slot uses the low bits)
}
// the three key words a slot stores for a key of len bytes held in w: bytes 0..7, 8..15, 16..22 | len << 56 (len <= 23), else a marker
constexpr uint32_t           kInlineKeyBytes = 23;
constexpr unsigned long long kLongKey        = 0xFFull << 56;
template <int NW>
__device__ __forceinline__ void slot_key(const uint64_t (&w)[NW], uint32_t len, unsigned long long& k0, unsigned long long& k1, unsigned long long& k2) {
    k0 = w[0];
    k1 = NW > 1 ? w[1] : 0ull;
    k2 = len <= kInlineKeyBytes ? ((NW > 2 ? w[2] : 0ull) | ((unsigned long long)len << 56)) : kLongKey;
}
// probe for a key whose hash is h: pattern index + 1 (and the slot, for the counter), or 0
template <int NW>
__device__ __forceinline__ uint32_t index_probe(uint64_t h, const uint64_t (&w)[NW], uint32_t len, const uint8_t* __restrict__ keys, const uint64_t* __restrict__ off,
                                                const PatSlot* __restrict__ slots, uint64_t mask, uint64_t& slot_out) {
    unsigned long long k0, k1, k2;
    slot_key(w, len, k0, k1, k2);
    uint64_t s = h & mask;
    for (uint64_t step = 0; step <= mask; ++step, s = (s + 1) & mask) {
        const uint4 lo = __ldg(reinterpret_cast<const uint4*>(slots + s));  // {idx1, count, k0}; read-only path: the key words never change after the build
        if (lo.x == 0) return 0;
        if ((((unsigned long long)lo.w << 32) | lo.z) != k0) continue;
        const uint4 hi = __ldg(reinterpret_cast<const uint4*>(slots + s) + 1);  // {k1, k2}: same sector
        if ((((unsigned long long)hi.y << 32) | hi.x) != k1 || (((unsigned long long)hi.w << 32) | hi.z) != k2) continue;
        if (len > kInlineKeyBytes) {  // rare: the rest of a long key lives in the blob
            const uint64_t o = off[lo.x - 1];
            if (off[lo.x] - o != len) continue;
            bool same = true;
            for (uint32_t i = 16; i < len; ++i)
                if (keys[o + i] != words_get(w, i)) {
                    same = false;
                    break;
                }
            if (!same) continue;
        }
        slot_out = s;
        return lo.x;
    }
    return 0;
}
template <int NW>
__device__ __forceinline__ uint32_t index_find(const uint64_t (&w)[NW], uint32_t len, const uint8_t* __restrict__ keys, const uint64_t* __restrict__ off,
                                               const PatSlot* __restrict__ slots, uint64_t mask, const uint32_t* __restrict__ presence, uint64_t pmask, uint64_t& slot_out) {
    const uint64_t h = spooky_words(w, len);
    if (presence != nullptr) {
        const uint64_t b = presence_bit(h, pmask);
        if (((__ldg(presence + (b >> 5)) >> (b & 31)) & 1u) == 0) return 0;
    }
    return index_probe(h, w, len, keys, off, slots, mask, slot_out);
}

// ---------------------------------------------------------------------------------------------
// per-pattern shape: tokens (Pattern::n = one per byte < 128, src/pattern.cpp:74-97) and category (datacategory, :23-43:
// the first skip (3) / flex (4) token decides); model-wide maxima for postread (include/patternmodel.h:572-588)
__global__ void __launch_bounds__(256) pattern_meta_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ off, uint64_t np, uint16_t* __restrict__ pn,
                                                           uint8_t* __restrict__ pcat, PatternMetaStats* __restrict__ st) {
    __shared__ unsigned long long s_nhist[256];
    __shared__ uint32_t           s_maxn, s_minn, s_skip, s_flex, s_maxclass, s_malformed, s_uni;
    s_nhist[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        s_maxn = s_skip = s_flex = s_maxclass = s_malformed = s_uni = 0;
        s_minn = 0xFFFFFFFFu;
    }
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t a = off[i], b = off[i + 1];
        uint32_t       n = 0, cat = 0;
        uint64_t       cls = 0, maxcls = 0;
        uint32_t       shift = 0;
        bool           start = true, wide = false;
        for (uint64_t k = a; k < b; ++k) {
            const uint8_t c = keys[k];
            if (shift < 35) cls |= (uint64_t)(c & 0x7F) << shift; else wide = true;
            shift += 7;
            if (c < 128) {
                if (start && cat == 0 && c == 3) cat = 1;
                if (start && cat == 0 && c == 4) cat = 2;
                ++n;
                if (cls > maxcls) maxcls = cls;
                cls   = 0;
                shift = 0;
                start = true;
            } else {
                start = false;
            }
        }
        if (maxcls > 0xFFFFFFFFull) wide = true;
        pn[i]   = (uint16_t)min(n, 65535u);
        pcat[i] = (uint8_t)cat;
        atomicMax(&s_maxn, n);
        atomicMin(&s_minn, n);
        if (cat == 1) s_skip = 1;
        if (cat == 2) s_flex = 1;
        atomicMax(&s_maxclass, (unsigned)min(maxcls, (uint64_t)0xFFFFFFFFull));
        if (b - a > kMaxIndexedKeyBytes || b == a || wide || (b > a && keys[b - 1] >= 128)) atomicAdd(&s_malformed, 1u);
        atomicAdd(&s_nhist[min(n, 255u)], 1ull);
        if (n == 1 && cat == 0) atomicAdd(&s_uni, 1u);
    }
    __syncthreads();
    if (s_nhist[threadIdx.x]) atomicAdd(&st->nhist[threadIdx.x], s_nhist[threadIdx.x]);
    if (threadIdx.x == 0 && s_minn != 0xFFFFFFFFu) {
        atomicMax(&st->maxn, s_maxn);
        atomicMin(&st->minn, s_minn);
        if (s_skip) st->hasskip = 1;
        if (s_flex) st->hasflex = 1;
        atomicMax(&st->maxclass, s_maxclass);
        if (s_malformed) atomicAdd(&st->malformed, s_malformed);
        if (s_uni) atomicAdd(&st->unigram_ngrams, s_uni);
    }
}

// rep (optional): group-by mode -- equal keys are expected; rep[i] = the index of the copy that claimed the slot (rep[i] == i for that one)
__global__ void __launch_bounds__(256) index_build_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ off, uint64_t np, PatSlot* __restrict__ slots,
                                                          uint64_t mask, uint32_t* __restrict__ presence, uint64_t pmask, PatternMetaStats* __restrict__ st,
                                                          uint32_t* __restrict__ rep) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const uint64_t a   = off[i];
    const uint32_t len = (uint32_t)min(off[i + 1] - a, (uint64_t)kMaxIndexedKeyBytes);
    uint64_t       w[24];
    words_load(w, keys + a, len);
    const uint64_t h = spooky_words(w, len);
    {
        const uint64_t b = presence_bit(h, pmask);
        atomicOr(presence + (b >> 5), 1u << (b & 31));
    }
    unsigned long long k0, k1, k2;
    slot_key(w, len, k0, k1, k2);
    uint64_t s = h & mask;
    for (uint64_t step = 0; step <= mask; ++step, s = (s + 1) & mask) {
        const uint32_t prev = atomicCAS(&slots[s].idx1, 0u, (uint32_t)i + 1);
        if (prev == 0) {  // claimed: nobody reads the key words before this kernel has finished
            slots[s].k0 = k0;
            slots[s].k1 = k1;
            slots[s].k2 = k2;
            if (rep) rep[i] = (uint32_t)i;
            return;
        }
        // a second copy of the same pattern?  (compared in the blob: the other slot's key words may still be in flight)
        const uint32_t j = prev - 1;
        const uint64_t o = off[j];
        if (off[j + 1] - o == len) {
            bool same = true;
            for (uint32_t k = 0; k < len; ++k)
                if (keys[o + k] != keys[a + k]) {
                    same = false;
                    break;
                }
            if (same) {
                if (rep)
                    rep[i] = j;
                else
                    atomicAdd(&st->duplicates, 1u);
                return;
            }
        }
    }
    atomicAdd(&st->malformed, 1u);  // table full: cannot happen with cap >= 2 * np
}

// batch lookup: out_idx1[q] = pattern index + 1, or 0
__global__ void __launch_bounds__(256) index_lookup_kernel(const uint8_t* __restrict__ qkeys, const uint64_t* __restrict__ qoff, uint64_t nq, const uint8_t* __restrict__ keys,
                                                           const uint64_t* __restrict__ off, const PatSlot* __restrict__ slots, uint64_t mask,
                                                           const uint32_t* __restrict__ presence, uint64_t pmask, uint32_t* __restrict__ out_idx1) {
    uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const uint64_t a = qoff[q], len64 = qoff[q + 1] - a;
    if (len64 == 0 || len64 > kMaxIndexedKeyBytes) {
        out_idx1[q] = 0;
        return;
    }
    uint64_t w[24];
    words_load(w, qkeys + a, (uint32_t)len64);
    uint64_t slot;
    out_idx1[q] = index_find(w, (uint32_t)len64, keys, off, slots, mask, presence, pmask, slot);
}
__global__ void __launch_bounds__(256) gather_counts_kernel(const uint32_t* __restrict__ idx1, uint64_t nq, const uint32_t* __restrict__ counts, uint32_t* __restrict__ out) {
    uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) out[q] = idx1[q] ? counts[idx1[q] - 1] : 0u;
}

// ---------------------------------------------------------------------------------------------
// constrained training, one window length per launch: window (p, n) is counted iff the constraint set has its bytes
// (include/patternmodel.h:1064-1072 subngrams, :1088-1089 has(), :1155-1161 add)
// ---------------------------------------------------------------------------------------------
// keys of at most 31 bytes (every window of up to 6 tokens) held in four registers: no dynamically indexed array, hence no local memory
__device__ __forceinline__ void key_append4(uint64_t (&w)[4], uint32_t& len, uint32_t c) {
    uint64_t v;  // the varint bytes of class c as one little-endian value (src/classencoder.cpp:22-42)
    uint32_t nb;
    if (c < (1u << 7)) {
        v  = c;
        nb = 1;
    } else if (c < (1u << 14)) {
        v  = (uint64_t)((c & 0x7F) | 0x80) | ((uint64_t)(c >> 7) << 8);
        nb = 2;
    } else if (c < (1u << 21)) {
        v  = (uint64_t)((c & 0x7F) | 0x80) | ((uint64_t)(((c >> 7) & 0x7F) | 0x80) << 8) | ((uint64_t)(c >> 14) << 16);
        nb = 3;
    } else if (c < (1u << 28)) {
        v  = (uint64_t)((c & 0x7F) | 0x80) | ((uint64_t)(((c >> 7) & 0x7F) | 0x80) << 8) | ((uint64_t)(((c >> 14) & 0x7F) | 0x80) << 16) | ((uint64_t)(c >> 21) << 24);
        nb = 4;
    } else {
        v = (uint64_t)((c & 0x7F) | 0x80) | ((uint64_t)(((c >> 7) & 0x7F) | 0x80) << 8) | ((uint64_t)(((c >> 14) & 0x7F) | 0x80) << 16) |
            ((uint64_t)(((c >> 21) & 0x7F) | 0x80) << 24) | ((uint64_t)(c >> 28) << 32);
        nb = 5;
    }
    const uint32_t wi = len >> 3, bo = (len & 7) * 8;
    const uint64_t lo = v << bo;
    const uint64_t hi = bo ? (v >> (64 - bo)) : 0ull;  // the bytes that spill into the next word
    // selects on scalars (written out: an indexed form gets turned back into a local-memory array by the compiler)
    w[0] |= wi == 0 ? lo : 0ull;
    w[1] |= wi == 1 ? lo : (wi == 0 ? hi : 0ull);
    w[2] |= wi == 2 ? lo : (wi == 1 ? hi : 0ull);
    w[3] |= wi == 3 ? lo : (wi == 2 ? hi : 0ull);
    len += nb;
}
// SpookyHash::Hash64 for len <= 31 over four register words (static indices only)
__device__ __forceinline__ uint64_t spooky_words4(const uint64_t (&w)[4], uint32_t len) {
    uint64_t a = 0, b = 0, c = kSpookyConst, d = kSpookyConst;
    uint32_t rem = len;
    uint64_t t0 = w[0], t1 = w[1];
    if (len > 15) {
        c += w[0];
        d += w[1];
        spooky_short_mix(a, b, c, d);
        rem -= 16;
        t0 = w[2];
        t1 = w[3];
    }
    d += (uint64_t)len << 56;
    if (rem == 0) {
        c += kSpookyConst;
        d += kSpookyConst;
    } else {
        c += t0;
        if (rem > 8) d += t1;
    }
    spooky_short_end(a, b, c, d);
    return a;
}
__device__ __forceinline__ uint32_t index_find4(const uint64_t (&w)[4], uint32_t len, const uint8_t* __restrict__ keys, const uint64_t* __restrict__ off,
                                                const PatSlot* __restrict__ slots, uint64_t mask, const uint32_t* __restrict__ presence, uint64_t pmask, uint64_t& slot_out) {
    const uint64_t h = spooky_words4(w, len);
    {
        const uint64_t b = presence_bit(h, pmask);
        if (((__ldg(presence + (b >> 5)) >> (b & 31)) & 1u) == 0) return 0;
    }
    const unsigned long long k0 = w[0], k1 = w[1];
    const unsigned long long k2 = len <= kInlineKeyBytes ? (w[2] | ((unsigned long long)len << 56)) : kLongKey;
    uint64_t s = h & mask;
    for (uint64_t step = 0; step <= mask; ++step, s = (s + 1) & mask) {
        const uint4 lo = __ldg(reinterpret_cast<const uint4*>(slots + s));
        if (lo.x == 0) return 0;
        if ((((unsigned long long)lo.w << 32) | lo.z) != k0) continue;
        const uint4 hi = __ldg(reinterpret_cast<const uint4*>(slots + s) + 1);
        if ((((unsigned long long)hi.y << 32) | hi.x) != k1 || (((unsigned long long)hi.w << 32) | hi.z) != k2) continue;
        if (len > kInlineKeyBytes) {  // 24..31 bytes: bytes 16.. are compared in the blob (w[2], w[3] hold them)
            const uint64_t o = off[lo.x - 1];
            if (off[lo.x] - o != len) continue;
            uint64_t p2 = 0, p3 = 0;
            for (uint32_t i = 16; i < len; ++i) {
                const uint64_t byte = keys[o + i];
                if (i < 24) p2 |= byte << (8 * (i - 16)); else p3 |= byte << (8 * (i - 24));
            }
            if (p2 != w[2] || p3 != w[3]) continue;
        }
        slot_out = s;
        return lo.x;
    }
    return 0;
}

// Frequent patterns ("the", "of the") would put millions of atomics on one counter, where they serialise.  Each block keeps a small
// never-evicting cache in shared memory, pattern -> pending count: the first window of a pattern that finds its line empty claims it,
// later windows of that pattern in the block bump the shared counter, and the block adds the total to the pattern's counter once when it
// retires.  A line never changes owner, so the counts stay exact; patterns that lose the race for a line go to the global counter.
constexpr uint32_t kMatchLines = 2048;

// one count for pattern idx1 (> 0) through the block's cache; a pattern that lost the race for its cache line counts in `direct`
// (its own index slot -- the sector the probe just fetched -- or, for unigrams, its entry of counts[])
__device__ __forceinline__ void count_match(uint32_t idx1, uint32_t* line_key, uint32_t* line_cnt, uint32_t* direct) {
    const uint32_t line = (idx1 * 2654435761u) >> (32 - 11);  // kMatchLines = 2^11
    uint32_t       own  = *(volatile uint32_t*)&line_key[line];
    if (own == 0) {
        own = atomicCAS(&line_key[line], 0u, idx1);
        if (own == 0) own = idx1;
    }
    if (own == idx1)
        atomicAdd(&line_cnt[line], 1u);
    else
        atomicAdd(direct, 1u);
}

// prev (optional): match[] of length n-1.  When every pattern of length n has its (n-1)-token prefix in the set (use_prefix), a window
// whose prefix did not match cannot match either; likewise for the suffix.  This is the back-off rule of unconstrained training
// (include/patternmodel.h:1139-1152) re-derived for sets that happen to be closed -- every model train() itself produced is.
template <int NW>
__global__ void __launch_bounds__(256, NW == 4 ? 6 : 1) constrained_match_kernel(const uint32_t* __restrict__ tok, uint64_t npos, int n, const uint8_t* __restrict__ keys,
                                                                const uint64_t* __restrict__ off, PatSlot* __restrict__ slots, uint64_t mask,
                                                                const uint32_t* __restrict__ presence, uint64_t pmask, uint32_t* __restrict__ counts,
                                                                uint32_t* __restrict__ match, const uint32_t* __restrict__ prev, bool use_prefix, bool use_suffix,
                                                                DeviceStats* __restrict__ st) {
    __shared__ uint32_t line_key[kMatchLines];
    __shared__ uint32_t line_cnt[kMatchLines];
    for (uint32_t i = threadIdx.x; i < kMatchLines; i += blockDim.x) {
        line_key[i] = 0;
        line_cnt[i] = 0;
    }
    __syncthreads();
    unsigned long long windows = 0;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npos; p += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t idx1 = 0;
        bool     candidate = true;
        if (use_prefix && __ldcs(prev + p) == 0) candidate = false;
        if (candidate && use_suffix && __ldg(prev + p + 1) == 0) candidate = false;
        if (candidate) {
            uint64_t w[NW];
            words_clear(w);
            uint32_t len = 0;
            bool     ok  = true;
            for (int j = 0; j < n; ++j) {
                uint32_t t = tok[p + j];  // a delimiter (0) ends the scan, so the read never passes the end of its sentence
                if (t == 0) {
                    ok = false;
                    break;
                }
                if constexpr (NW == 4) {
                    key_append4(w, len, t);
                } else {
                    do {
                        const uint8_t digit = (uint8_t)(t & 0x7F);
                        t >>= 7;
                        words_put(w, len, t ? (uint8_t)(digit | 0x80) : digit);
                        ++len;
                    } while (t);
                }
            }
            if (ok) {
                ++windows;
                uint64_t slot = 0;
                if (len <= kMaxIndexedKeyBytes && len < 8u * NW) {
                    if constexpr (NW == 4)
                        idx1 = index_find4(w, len, keys, off, slots, mask, presence, pmask, slot);
                    else
                        idx1 = index_find(w, len, keys, off, slots, mask, presence, pmask, slot);
                }
                if (idx1) count_match(idx1, line_key, line_cnt, &slots[slot].count);
            }
        }
        if (match) __stcs(match + p, idx1);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < kMatchLines; i += blockDim.x) {
        const uint32_t c = line_cnt[i];
        if (c) atomicAdd(&counts[line_key[i] - 1], c);
    }
    windows = warp_reduce_sum(windows);
    if (lane_id() == 0 && windows) atomicAdd(&st->valid_windows, windows);
}

// length 1 needs no hashing: classes are dense (src/classencoder.cpp:213-226), so the unigram patterns are a class-indexed array
// (uni[class] = pattern index + 1), like the level-1 histogram of unconstrained training
__global__ void __launch_bounds__(256) unigram_table_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ off, const uint16_t* __restrict__ pn, uint64_t np,
                                                            uint32_t* __restrict__ uni, uint32_t nclasses) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np || pn[i] != 1) return;
    const uint64_t a = off[i], l = off[i + 1] - a;
    if (l == 0 || l > 5) return;
    uint64_t cls = 0;
    for (uint64_t k = 0; k < l; ++k) cls |= (uint64_t)(keys[a + k] & 0x7F) << (7 * k);
    // only the canonical encoding of a class can equal the bytes of a corpus token (the tokeniser refuses the others)
    if (cls >= nclasses || varint_len((uint32_t)cls) != l) return;
    uni[cls] = (uint32_t)i + 1;
}
// level 1 of a constrained run = the class histogram of unconstrained training (unigram_hist_kernel, kernels.cu) handed to the unigram
// patterns: counts[uni[c] - 1] += hist[c]; the per-position matches (only needed when level 2 chains on them) are one streaming pass
__global__ void __launch_bounds__(256) unigram_apply_kernel(const uint32_t* __restrict__ hist, const uint32_t* __restrict__ uni, uint32_t nclasses, uint32_t* __restrict__ counts) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nclasses) return;
    const uint32_t idx1 = uni[c], h = hist[c];
    if (idx1 != 0 && h != 0) counts[idx1 - 1] += h;  // a class has at most one unigram pattern
}
__global__ void __launch_bounds__(256) unigram_match_kernel(const uint32_t* __restrict__ tok, uint64_t npos, const uint32_t* __restrict__ uni, uint32_t nclasses,
                                                            uint32_t* __restrict__ match) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npos) return;
    const uint32_t t = __ldcs(tok + p);
    __stcs(match + p, (t != 0 && t < nclasses) ? __ldg(uni + t) : 0u);
}

// closure of the set per pattern length: how many patterns of n tokens lack their (n-1)-token prefix / suffix in the set
__global__ void __launch_bounds__(256) closure_check_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ off, const uint16_t* __restrict__ pn, uint64_t np,
                                                            const PatSlot* __restrict__ slots, uint64_t mask, const uint32_t* __restrict__ presence, uint64_t pmask,
                                                            unsigned long long* __restrict__ prefix_open, unsigned long long* __restrict__ suffix_open) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const uint32_t n = pn[i];
    if (n < 2) return;
    const uint64_t a = off[i], b = off[i + 1];
    if (b - a > kMaxIndexedKeyBytes) return;
    // end of the first token, start of the last token
    uint64_t first_end = a, last_start = a;
    {
        uint64_t k = a;
        while (k < b && keys[k] >= 128) ++k;
        first_end = k + 1;
        uint64_t start = a;
        for (k = a; k < b; ++k)
            if (keys[k] < 128 && k + 1 < b) start = k + 1;
        last_start = start;
    }
    uint64_t w[24];
    words_load(w, keys + a, (uint32_t)(last_start - a));
    uint64_t slot;
    if (index_find(w, (uint32_t)(last_start - a), keys, off, slots, mask, presence, pmask, slot) == 0) atomicAdd(&prefix_open[min(n, 255u)], 1ull);
    words_load(w, keys + first_end, (uint32_t)(b - first_end));
    if (index_find(w, (uint32_t)(b - first_end), keys, off, slots, mask, presence, pmask, slot) == 0) atomicAdd(&suffix_open[min(n, 255u)], 1ull);
}

// the counts taken in the index slots go to counts[] (every pattern has exactly one slot) and the slots are ready for the next run
__global__ void __launch_bounds__(256) collect_slot_counts_kernel(PatSlot* __restrict__ slots, uint64_t cap, uint32_t* __restrict__ counts) {
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += (uint64_t)gridDim.x * blockDim.x) {
        const uint2 v = __ldcs(reinterpret_cast<const uint2*>(slots + s));  // {idx1, count}
        if (v.x != 0 && v.y != 0) {
            counts[v.x - 1] += v.y;
            slots[s].count = 0;
        }
    }
}

// after the scan: found = patterns seen at least once, kept = count >= threshold, per-length sums of the kept ones
__global__ void __launch_bounds__(256) constrained_stats_kernel(const uint32_t* __restrict__ counts, const uint16_t* __restrict__ pn, uint64_t np, uint32_t threshold,
                                                                uint32_t* __restrict__ flags, PatternMetaStats* __restrict__ st, DeviceStats* __restrict__ ds, bool per_length) {
    __shared__ unsigned long long s_kept_n[256], s_occ_n[256];
    __shared__ uint64_t           scratch[8];
    if (per_length) {
        s_kept_n[threadIdx.x] = 0;
        s_occ_n[threadIdx.x]  = 0;
        __syncthreads();
    }
    // per-thread registers first, one block reduction at the end; the per-length sums (indexed models only) are aggregated per warp
    unsigned long long found = 0, kept_c = 0, occ = 0;
    uint32_t           maxn = 0, minn = 0xFFFFFFFFu;
    const uint64_t     stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t     rounds = (np + stride - 1) / stride;  // the same trip count for every thread: the warp collectives below stay converged
    for (uint64_t r = 0; r < rounds; ++r) {
        const uint64_t i = r * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        uint32_t       c = 0, n = 0;
        bool           kept = false;
        if (i < np) {
            c        = counts[i];
            n        = min((uint32_t)pn[i], 255u);
            kept     = c >= threshold;
            flags[i] = kept ? 1u : 0u;
            found += c > 0;
        }
        if (kept) {
            ++kept_c;
            occ += c;
            maxn = max(maxn, n);
            minn = min(minn, n);
        }
        if (per_length) {
            uint32_t todo = __ballot_sync(0xffffffffu, kept);
            while (todo) {  // one round per distinct length present in the warp (a handful)
                const uint32_t n0   = __shfl_sync(0xffffffffu, n, __ffs(todo) - 1);
                const bool     mine = kept && n == n0;
                const uint32_t grp  = __ballot_sync(0xffffffffu, mine);
                const uint64_t sum  = warp_reduce_sum(mine ? (uint64_t)c : 0ull);
                if (lane_id() == 0) {
                    atomicAdd(&s_kept_n[n0], (unsigned long long)__popc(grp));
                    atomicAdd(&s_occ_n[n0], (unsigned long long)sum);
                }
                todo &= ~grp;
            }
        }
    }
    const uint64_t f = block_reduce_sum(found, scratch);
    const uint64_t k = block_reduce_sum(kept_c, scratch);
    const uint64_t o = block_reduce_sum(occ, scratch);
    maxn = warp_reduce_max(maxn);
    minn = ~warp_reduce_max(~minn);
    if (lane_id() == 0 && minn != 0xFFFFFFFFu) {
        atomicMax(&st->kept_maxn, maxn);
        atomicMin(&st->kept_minn, minn);
    }
    if (threadIdx.x == 0) {
        if (f) atomicAdd(&ds->found, (unsigned long long)f);
        if (k) {
            atomicAdd(&ds->kept, (unsigned long long)k);
            atomicAdd(&ds->kept_occ, (unsigned long long)o);
        }
    }
    if (per_length) {
        __syncthreads();
        if (s_kept_n[threadIdx.x]) {
            atomicAdd(&st->kept_n[threadIdx.x], s_kept_n[threadIdx.x]);
            atomicAdd(&st->kept_occ_n[threadIdx.x], s_occ_n[threadIdx.x]);
        }
    }
}

// load-time filter (PatternMapStore::read, include/patternstore.h:574-586): category switches, length window, occurrence
// threshold, membership in the constraint store
__global__ void __launch_bounds__(256) load_filter_kernel(const uint16_t* __restrict__ pn, const uint8_t* __restrict__ pcat, const uint32_t* __restrict__ counts,
                                                          const uint32_t* __restrict__ constrain_idx1, uint64_t np, uint32_t mintokens, uint32_t minlength, uint32_t maxlength,
                                                          int dongrams, int doskipgrams, int doflexgrams, uint32_t* __restrict__ flags, PatternMetaStats* __restrict__ st) {
    __shared__ unsigned long long s_kept_n[256];
    __shared__ uint32_t           s_maxn, s_minn, s_skip, s_flex;
    s_kept_n[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        s_maxn = 0;
        s_minn = 0xFFFFFFFFu;
        s_skip = s_flex = 0;
    }
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t n = pn[i], cat = pcat[i];
        bool keep = !((!dongrams && cat == 0) || (!doskipgrams && cat == 1) || (!doflexgrams && cat == 2));
        keep      = keep && n >= minlength && n <= maxlength && counts[i] >= mintokens;
        if (keep && constrain_idx1 != nullptr) keep = constrain_idx1[i] != 0;
        flags[i] = keep ? 1u : 0u;
        if (keep) {
            atomicMax(&s_maxn, n);
            atomicMin(&s_minn, n);
            if (cat == 1) s_skip = 1;
            if (cat == 2) s_flex = 1;
            atomicAdd(&s_kept_n[min(n, 255u)], 1ull);
        }
    }
    __syncthreads();
    if (s_kept_n[threadIdx.x]) atomicAdd(&st->kept_n[threadIdx.x], s_kept_n[threadIdx.x]);
    if (threadIdx.x == 0 && s_minn != 0xFFFFFFFFu) {
        atomicMax(&st->kept_maxn, s_maxn);
        atomicMin(&st->kept_minn, s_minn);
        if (s_skip) st->kept_hasskip = 1;
        if (s_flex) st->kept_hasflex = 1;
    }
}

// ---------------------------------------------------------------------------------------------
// compaction: survivors in index order -> (optionally) ordered by length -> gathered into a new flat model
__global__ void __launch_bounds__(256) select_scatter_kernel(const uint32_t* __restrict__ flags, const uint64_t* __restrict__ newpos, const uint16_t* __restrict__ pn, uint64_t np,
                                                             uint32_t* __restrict__ sel_idx, uint32_t* __restrict__ sel_n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np && flags[i]) {
        sel_idx[newpos[i]] = (uint32_t)i;
        sel_n[newpos[i]]   = pn[i];
    }
}
// j-th survivor = old pattern sel_idx[j]: its key length, count, and the old -> new map (new index + 1)
__global__ void __launch_bounds__(256) gather_meta_kernel(const uint32_t* __restrict__ sel_idx, uint64_t k, const uint64_t* __restrict__ off, const uint32_t* __restrict__ counts,
                                                          uint32_t* __restrict__ kmap, uint32_t* __restrict__ lens, uint16_t* __restrict__ len16, uint32_t* __restrict__ counts_out) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    const uint32_t i = sel_idx[j];
    const uint32_t l = (uint32_t)(off[i + 1] - off[i]);
    if (kmap) kmap[i] = (uint32_t)j + 1;
    lens[j]       = l;
    len16[j]      = (uint16_t)min(l, 65535u);
    counts_out[j] = counts ? counts[i] : 0u;
}
__global__ void __launch_bounds__(256) gather_keys_kernel(const uint32_t* __restrict__ sel_idx, uint64_t k, const uint8_t* __restrict__ keys, const uint64_t* __restrict__ off,
                                                          const uint64_t* __restrict__ new_off, uint8_t* __restrict__ out) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    const uint32_t i = sel_idx[j];
    const uint64_t a = off[i], l = off[i + 1] - a, d = new_off[j];
    for (uint64_t b = 0; b < l; ++b) out[d + b] = keys[a + b];
}
// occurrence lists of the survivors (indexed load): one warp per survivor copies its run
__global__ void __launch_bounds__(256) gather_refs_kernel(const uint32_t* __restrict__ sel_idx, uint64_t k, const uint64_t* __restrict__ ref_off, const uint32_t* __restrict__ rs,
                                                          const uint16_t* __restrict__ rt, const uint64_t* __restrict__ new_ref_off, uint32_t* __restrict__ rs_out,
                                                          uint16_t* __restrict__ rt_out) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= k) return;
    const uint32_t i = sel_idx[warp];
    const uint64_t a = ref_off[i], l = ref_off[i + 1] - a, d = new_ref_off[warp];
    for (uint64_t b = lane_id(); b < l; b += 32) {
        rs_out[d + b] = rs[a + b];
        rt_out[d + b] = rt[a + b];
    }
}
__global__ void __launch_bounds__(256) iota_kernel(uint32_t* __restrict__ out, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)i;
}

// distinct word types covered by the patterns (totalwordtypesingroup(0, 0), include/patternmodel.h:1953-1975): one bit per class
__global__ void __launch_bounds__(256) token_bitmap_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ off, uint64_t np, uint32_t* __restrict__ bitmap) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    uint64_t cls   = 0;
    uint32_t shift = 0;
    for (uint64_t k = off[i]; k < off[i + 1]; ++k) {
        const uint8_t c = keys[k];
        if (shift < 35) cls |= (uint64_t)(c & 0x7F) << shift;
        shift += 7;
        if (c < 128) {
            const uint32_t v = (uint32_t)cls;
            atomicOr(&bitmap[v >> 5], 1u << (v & 31));
            cls   = 0;
            shift = 0;
        }
    }
}
__global__ void __launch_bounds__(256) popcount_kernel(const uint32_t* __restrict__ words, uint64_t n, unsigned long long* __restrict__ total) {
    unsigned long long c = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) c += __popc(words[i]);
    c = warp_reduce_sum(c);
    if (lane_id() == 0 && c) atomicAdd(total, c);
}

// ---------------------------------------------------------------------------------------------
// streaming kernels over the patterns: enough blocks to fill the machine, each aggregating its statistics in shared memory first
static inline unsigned pattern_grid(uint64_t np) {
    return (unsigned)std::min<uint64_t>(pi_div_up(np, 256), 148 * 16);
}
int launch_pattern_meta(cudaStream_t s, const uint8_t* keys, const uint64_t* off, uint64_t np, uint16_t* pn, uint8_t* pcat, PatternMetaStats* st) {
    if (!np) return 0;
    pattern_meta_kernel<<<pattern_grid(np), 256, 0, s>>>(keys, off, np, pn, pcat, st);
    return 1;
}
int launch_index_build(cudaStream_t s, const uint8_t* keys, const uint64_t* off, uint64_t np, PatSlot* slots, uint64_t cap_pow2, uint32_t* presence,
                       uint64_t presence_bits_pow2, PatternMetaStats* st, uint32_t* rep) {
    if (!np) return 0;
    index_build_kernel<<<pi_div_up(np, 256), 256, 0, s>>>(keys, off, np, slots, cap_pow2 - 1, presence, presence_bits_pow2 - 1, st, rep);
    return 1;
}
int launch_index_lookup(cudaStream_t s, const uint8_t* qkeys, const uint64_t* qoff, uint64_t nq, const uint8_t* keys, const uint64_t* off, const PatSlot* slots,
                        uint64_t cap_pow2, const uint32_t* presence, uint64_t presence_bits_pow2, uint32_t* out_idx1) {
    if (!nq) return 0;
    index_lookup_kernel<<<pi_div_up(nq, 256), 256, 0, s>>>(qkeys, qoff, nq, keys, off, slots, cap_pow2 - 1, presence, presence_bits_pow2 - 1, out_idx1);
    return 1;
}
int launch_gather_counts(cudaStream_t s, const uint32_t* idx1, uint64_t nq, const uint32_t* counts, uint32_t* out) {
    if (!nq) return 0;
    gather_counts_kernel<<<pi_div_up(nq, 256), 256, 0, s>>>(idx1, nq, counts, out);
    return 1;
}
static int match_blocks_per_sm(const void* fn) {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, 256, 0);
    return n > 0 ? n : 1;
}
int launch_constrained_match(cudaStream_t s, const uint32_t* tok, uint64_t npos, int n, const uint8_t* keys, const uint64_t* off, PatSlot* slots, uint64_t cap_pow2,
                             const uint32_t* presence, uint64_t presence_bits_pow2, uint32_t* counts, uint32_t* match, const uint32_t* prev, bool use_prefix, bool use_suffix,
                             DeviceStats* st, int sms) {
    if (!npos) return 0;
    if (prev == nullptr) use_prefix = use_suffix = false;
    // resident blocks only (each owns one count cache), exactly as many as fit: a partial second wave would idle half the machine
    if (n * 5 < 8 * 4) {  // every window of n tokens fits 31 bytes: registers only
        static int     bps  = getenv("COLIBRI_B200_MATCH_BPS") ? atoi(getenv("COLIBRI_B200_MATCH_BPS")) : match_blocks_per_sm((const void*)constrained_match_kernel<4>);
        const unsigned grid = (unsigned)std::min<uint64_t>(pi_div_up(npos, 256), (uint64_t)sms * bps);
        constrained_match_kernel<4><<<grid, 256, 0, s>>>(tok, npos, n, keys, off, slots, cap_pow2 - 1, presence, presence_bits_pow2 - 1, counts, match, prev, use_prefix, use_suffix, st);
    } else {
        static int     bps  = match_blocks_per_sm((const void*)constrained_match_kernel<24>);
        const unsigned grid = (unsigned)std::min<uint64_t>(pi_div_up(npos, 256), (uint64_t)sms * bps);
        constrained_match_kernel<24><<<grid, 256, 0, s>>>(tok, npos, n, keys, off, slots, cap_pow2 - 1, presence, presence_bits_pow2 - 1, counts, match, prev, use_prefix, use_suffix, st);
    }
    return 1;
}
int launch_collect_slot_counts(cudaStream_t s, PatSlot* slots, uint64_t cap_pow2, uint32_t* counts) {
    if (!cap_pow2) return 0;
    collect_slot_counts_kernel<<<(unsigned)std::min<uint64_t>(pi_div_up(cap_pow2, 256), 148 * 16), 256, 0, s>>>(slots, cap_pow2, counts);
    return 1;
}
int launch_unigram_table(cudaStream_t s, const uint8_t* keys, const uint64_t* off, const uint16_t* pn, uint64_t np, uint32_t* uni, uint32_t nclasses) {
    if (!np) return 0;
    unigram_table_kernel<<<pi_div_up(np, 256), 256, 0, s>>>(keys, off, pn, np, uni, nclasses);
    return 1;
}
int launch_unigram_apply(cudaStream_t s, const uint32_t* hist, const uint32_t* uni, uint32_t nclasses, uint32_t* counts) {
    if (!nclasses) return 0;
    unigram_apply_kernel<<<pi_div_up(nclasses, 256), 256, 0, s>>>(hist, uni, nclasses, counts);
    return 1;
}
int launch_unigram_match(cudaStream_t s, const uint32_t* tok, uint64_t npos, const uint32_t* uni, uint32_t nclasses, uint32_t* match) {
    if (!npos) return 0;
    unigram_match_kernel<<<pi_div_up(npos, 256), 256, 0, s>>>(tok, npos, uni, nclasses, match);
    return 1;
}
int launch_closure_check(cudaStream_t s, const uint8_t* keys, const uint64_t* off, const uint16_t* pn, uint64_t np, const PatSlot* slots, uint64_t cap_pow2,
                         const uint32_t* presence, uint64_t presence_bits_pow2, unsigned long long* prefix_open, unsigned long long* suffix_open) {
    if (!np) return 0;
    closure_check_kernel<<<pi_div_up(np, 256), 256, 0, s>>>(keys, off, pn, np, slots, cap_pow2 - 1, presence, presence_bits_pow2 - 1, prefix_open, suffix_open);
    return 1;
}
int launch_constrained_stats(cudaStream_t s, const uint32_t* counts, const uint16_t* pn, uint64_t np, uint32_t threshold, uint32_t* flags, PatternMetaStats* st, DeviceStats* ds,
                             bool per_length) {
    if (!np) return 0;
    constrained_stats_kernel<<<pattern_grid(np), 256, 0, s>>>(counts, pn, np, threshold, flags, st, ds, per_length);
    return 1;
}
int launch_load_filter(cudaStream_t s, const uint16_t* pn, const uint8_t* pcat, const uint32_t* counts, const uint32_t* constrain_idx1, uint64_t np, uint32_t mintokens,
                       uint32_t minlength, uint32_t maxlength, int dongrams, int doskipgrams, int doflexgrams, uint32_t* flags, PatternMetaStats* st) {
    if (!np) return 0;
    load_filter_kernel<<<pattern_grid(np), 256, 0, s>>>(pn, pcat, counts, constrain_idx1, np, mintokens, minlength, maxlength, dongrams, doskipgrams, doflexgrams, flags, st);
    return 1;
}
int launch_select_scatter(cudaStream_t s, const uint32_t* flags, const uint64_t* newpos, const uint16_t* pn, uint64_t np, uint32_t* sel_idx, uint32_t* sel_n) {
    if (!np) return 0;
    select_scatter_kernel<<<pi_div_up(np, 256), 256, 0, s>>>(flags, newpos, pn, np, sel_idx, sel_n);
    return 1;
}
int launch_gather_meta(cudaStream_t s, const uint32_t* sel_idx, uint64_t k, const uint64_t* off, const uint32_t* counts, uint32_t* kmap, uint32_t* lens, uint16_t* len16,
                       uint32_t* counts_out) {
    if (!k) return 0;
    gather_meta_kernel<<<pi_div_up(k, 256), 256, 0, s>>>(sel_idx, k, off, counts, kmap, lens, len16, counts_out);
    return 1;
}
int launch_gather_keys(cudaStream_t s, const uint32_t* sel_idx, uint64_t k, const uint8_t* keys, const uint64_t* off, const uint64_t* new_off, uint8_t* out) {
    if (!k) return 0;
    gather_keys_kernel<<<pi_div_up(k, 256), 256, 0, s>>>(sel_idx, k, keys, off, new_off, out);
    return 1;
}
int launch_gather_refs(cudaStream_t s, const uint32_t* sel_idx, uint64_t k, const uint64_t* ref_off, const uint32_t* rs, const uint16_t* rt, const uint64_t* new_ref_off,
                       uint32_t* rs_out, uint16_t* rt_out) {
    if (!k) return 0;
    gather_refs_kernel<<<pi_div_up(k * 32, 256), 256, 0, s>>>(sel_idx, k, ref_off, rs, rt, new_ref_off, rs_out, rt_out);
    return 1;
}
int launch_iota(cudaStream_t s, uint32_t* out, uint64_t n) {
    if (!n) return 0;
    iota_kernel<<<pi_div_up(n, 256), 256, 0, s>>>(out, n);
    return 1;
}
int launch_token_bitmap(cudaStream_t s, const uint8_t* keys, const uint64_t* off, uint64_t np, uint32_t* bitmap) {
    if (!np) return 0;
    token_bitmap_kernel<<<pi_div_up(np, 256), 256, 0, s>>>(keys, off, np, bitmap);
    return 1;
}
int launch_popcount(cudaStream_t s, const uint32_t* words, uint64_t n, unsigned long long* total) {
    if (!n) return 0;
    popcount_kernel<<<(unsigned)std::min<uint64_t>(pi_div_up(n, 256), 1184), 256, 0, s>>>(words, n, total);
    return 1;
}

// ---- order-independent checksum of a model (measurement / parity aid: equal models give equal sums whatever order the table scans,
// the ranks or the reference's unordered_map put the patterns in).  Per pattern: FNV-1a 64 of the key bytes, mixed with the count;
// per occurrence of an indexed model: the same key hash mixed with (sentence, token).  Sums wrap modulo 2^64.
__device__ __forceinline__ uint64_t fnv1a64(const uint8_t* __restrict__ p, uint64_t len) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint64_t i = 0; i < len; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
    return h;
}
__global__ void __launch_bounds__(256) model_checksum_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ off, const uint32_t* __restrict__ counts, uint64_t np,
                                                             uint64_t* __restrict__ keyhash /* may be NULL */, unsigned long long* __restrict__ out) {
    __shared__ uint64_t scratch[8];
    uint64_t sum = 0, x = 0, occ = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t b = off[i], e = off[i + 1];
        const uint64_t h = fnv1a64(keys + b, e - b);
        if (keyhash) keyhash[i] = h;
        const uint64_t v = fmix64(h ^ ((uint64_t)counts[i] * 0x9E3779B97F4A7C15ull));
        sum += v;
        x ^= v;
        occ += counts[i];
    }
    sum = block_reduce_sum(sum, scratch);
    occ = block_reduce_sum(occ, scratch);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) x ^= __shfl_xor_sync(0xffffffffu, x, d);
    if (threadIdx.x == 0) {
        atomicAdd(&out[0], (unsigned long long)sum);
        atomicAdd(&out[2], (unsigned long long)occ);
    }
    if (lane_id() == 0) atomicXor(&out[1], (unsigned long long)x);
}
__global__ void __launch_bounds__(256) refs_checksum_kernel(const uint64_t* __restrict__ keyhash, const uint64_t* __restrict__ ref_off, uint64_t np, const uint32_t* __restrict__ rs,
                                                            const uint16_t* __restrict__ rt, uint64_t nrefs, unsigned long long* __restrict__ out) {
    __shared__ uint64_t scratch[8];
    uint64_t sum = 0;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nrefs; j += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t lo = 0, hi = np;  // the pattern whose list holds reference j: last i with ref_off[i] <= j
        while (hi - lo > 1) {
            uint64_t mid = (lo + hi) / 2;
            if (__ldg(ref_off + mid) <= j) lo = mid; else hi = mid;
        }
        sum += fmix64(keyhash[lo] ^ ((((uint64_t)rs[j] << 16) | rt[j]) * 0xD6E8FEB86659FD93ull));
    }
    sum = block_reduce_sum(sum, scratch);
    if (threadIdx.x == 0) atomicAdd(&out[4], (unsigned long long)sum);
}
int launch_model_checksum(cudaStream_t s, const uint8_t* keys, const uint64_t* off, const uint32_t* counts, uint64_t np, uint64_t* keyhash, unsigned long long* out) {
    if (!np) return 0;
    model_checksum_kernel<<<(unsigned)std::min<uint64_t>(pi_div_up(np, 256), 148 * 16), 256, 0, s>>>(keys, off, counts, np, keyhash, out);
    return 1;
}
int launch_refs_checksum(cudaStream_t s, const uint64_t* keyhash, const uint64_t* ref_off, uint64_t np, const uint32_t* rs, const uint16_t* rt, uint64_t nrefs, unsigned long long* out) {
    if (!nrefs || !np) return 0;
    refs_checksum_kernel<<<(unsigned)std::min<uint64_t>(pi_div_up(nrefs, 256), 148 * 16), 256, 0, s>>>(keyhash, ref_off, np, rs, rt, nrefs, out);
    return 1;
}

}  // namespace colibri
