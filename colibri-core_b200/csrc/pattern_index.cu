// pattern_index.cu -- a pattern set resident in HBM, addressed by Pattern::hash of the pattern BYTES.
//
// Reference: every membership test of the reference is `unordered_map<Pattern,...>::find` keyed by SpookyV2 of the byte string
// (src/pattern.cpp:234-238, include/patternstore.h:954-959): constrainbymodel->has(window) during constrained training
// (include/patternmodel.h:1088-1089), constrainstore->has(p) while loading (include/patternstore.h:586), has()/occurrencecount()
// queries (:751-756, :1653-1669).  Here a model's patterns are a flat blob + offsets; the index over them is an open-addressing
// table of 8-byte slots {hash tag (high 32 bits) | pattern index + 1}, placed by the low bits of SpookyV2 Hash64 of the bytes
// (the same function as Pattern::hash, csrc/spooky.h).  Keys compare in full (tag first, then the bytes), so it is exact.
//   * constrained_match_kernel: one thread per corpus position and window length n: re-encode the n class ids to their
//     varint bytes in registers, hash, probe, compare, atomicAdd the pattern's counter; optionally remember the match per
//     position (the forward index of indexed models is built from those with index.cu's ordered pairs + stable radix sort).
//   * index_lookup_kernel: batch queries (has / occurrencecount / load-time constraint).
//   * compaction kernels: threshold -> survivors, ordered by pattern length so that per-length occurrence lists concatenate.
// All of it is HBM-bound integer/byte work (random 8-byte probes + short byte compares).
#include <algorithm>

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
template <int NW>
__device__ __forceinline__ uint32_t index_find(const uint64_t (&w)[NW], uint32_t len, const uint8_t* __restrict__ keys, const uint64_t* __restrict__ off,
                                               const unsigned long long* __restrict__ slots, uint64_t mask) {
    const uint64_t h   = spooky_words(w, len);
    const uint32_t tag = (uint32_t)(h >> 32);
    uint64_t       s   = h & mask;
    for (uint64_t step = 0; step <= mask; ++step, s = (s + 1) & mask) {
        const unsigned long long v = slots[s];
        if (v == 0) return 0;
        if ((uint32_t)(v >> 32) != tag) continue;
        const uint32_t idx1 = (uint32_t)v;
        const uint64_t o    = off[idx1 - 1];
        if (off[idx1] - o != len) continue;
        bool same = true;
        for (uint32_t i = 0; i < len; ++i)
            if (keys[o + i] != words_get(w, i)) {
                same = false;
                break;
            }
        if (same) return idx1;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// per-pattern shape: tokens (Pattern::n = one per byte < 128, src/pattern.cpp:74-97) and category (datacategory, :23-43:
// the first skip (3) / flex (4) token decides); model-wide maxima for postread (include/patternmodel.h:572-588)
__global__ void __launch_bounds__(256) pattern_meta_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ off, uint64_t np, uint16_t* __restrict__ pn,
                                                           uint8_t* __restrict__ pcat, PatternMetaStats* __restrict__ st) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
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
    atomicMax(&st->maxn, n);
    atomicMin(&st->minn, n);
    if (cat == 1) st->hasskip = 1;
    if (cat == 2) st->hasflex = 1;
    atomicMax(&st->maxclass, (unsigned)min(maxcls, (uint64_t)0xFFFFFFFFull));
    if (b - a > kMaxIndexedKeyBytes || b == a || wide || (b > a && keys[b - 1] >= 128)) atomicAdd(&st->malformed, 1u);
    atomicAdd(&st->nhist[min(n, 255u)], 1ull);
    if (n == 1 && cat == 0) atomicAdd(&st->unigram_ngrams, 1u);
}

__global__ void __launch_bounds__(256) index_build_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ off, uint64_t np, unsigned long long* __restrict__ slots,
                                                          uint64_t mask, PatternMetaStats* __restrict__ st) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const uint64_t a   = off[i];
    const uint32_t len = (uint32_t)min(off[i + 1] - a, (uint64_t)kMaxIndexedKeyBytes);
    uint64_t       w[24];
    words_load(w, keys + a, len);
    const uint64_t           h   = spooky_words(w, len);
    const unsigned long long val = ((unsigned long long)(h >> 32) << 32) | (unsigned long long)(i + 1);
    uint64_t                 s   = h & mask;
    for (uint64_t step = 0; step <= mask; ++step, s = (s + 1) & mask) {
        unsigned long long prev = atomicCAS(&slots[s], 0ull, val);
        if (prev == 0) return;
        if ((prev >> 32) == (val >> 32)) {  // same tag: a second copy of the same pattern?
            const uint32_t j = (uint32_t)prev - 1;
            const uint64_t o = off[j];
            if (off[j + 1] - o == len) {
                bool same = true;
                for (uint32_t k = 0; k < len; ++k)
                    if (keys[o + k] != keys[a + k]) {
                        same = false;
                        break;
                    }
                if (same) {
                    atomicAdd(&st->duplicates, 1u);
                    return;
                }
            }
        }
    }
    atomicAdd(&st->malformed, 1u);  // table full: cannot happen with cap >= 2 * np
}

// batch lookup: out_idx1[q] = pattern index + 1, or 0
__global__ void __launch_bounds__(256) index_lookup_kernel(const uint8_t* __restrict__ qkeys, const uint64_t* __restrict__ qoff, uint64_t nq, const uint8_t* __restrict__ keys,
                                                           const uint64_t* __restrict__ off, const unsigned long long* __restrict__ slots, uint64_t mask,
                                                           uint32_t* __restrict__ out_idx1) {
    uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const uint64_t a = qoff[q], len64 = qoff[q + 1] - a;
    if (len64 == 0 || len64 > kMaxIndexedKeyBytes) {
        out_idx1[q] = 0;
        return;
    }
    uint64_t w[24];
    words_load(w, qkeys + a, (uint32_t)len64);
    out_idx1[q] = index_find(w, (uint32_t)len64, keys, off, slots, mask);
}
__global__ void __launch_bounds__(256) gather_counts_kernel(const uint32_t* __restrict__ idx1, uint64_t nq, const uint32_t* __restrict__ counts, uint32_t* __restrict__ out) {
    uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) out[q] = idx1[q] ? counts[idx1[q] - 1] : 0u;
}

// ---------------------------------------------------------------------------------------------
// constrained training, one window length per launch: window (p, n) is counted iff the constraint set has its bytes
// (include/patternmodel.h:1064-1072 subngrams, :1088-1089 has(), :1155-1161 add)
template <int NW>
__global__ void __launch_bounds__(256) constrained_match_kernel(const uint32_t* __restrict__ tok, uint64_t npos, int n, const uint8_t* __restrict__ keys,
                                                                const uint64_t* __restrict__ off, const unsigned long long* __restrict__ slots, uint64_t mask,
                                                                uint32_t* __restrict__ counts, uint32_t* __restrict__ match, DeviceStats* __restrict__ st) {
    unsigned long long windows = 0;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npos; p += (uint64_t)gridDim.x * blockDim.x) {
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
            do {
                const uint8_t digit = (uint8_t)(t & 0x7F);
                t >>= 7;
                words_put(w, len, t ? (uint8_t)(digit | 0x80) : digit);
                ++len;
            } while (t);
        }
        uint32_t idx1 = 0;
        if (ok) {
            ++windows;
            if (len <= kMaxIndexedKeyBytes && len < 8u * NW) idx1 = index_find(w, len, keys, off, slots, mask);
            if (idx1) atomicAdd(&counts[idx1 - 1], 1u);
        }
        if (match) match[p] = idx1;
    }
    windows = warp_reduce_sum(windows);
    if (lane_id() == 0 && windows) atomicAdd(&st->valid_windows, windows);
}

// after the scan: found = patterns seen at least once, kept = count >= threshold, per-length sums of the kept ones
__global__ void __launch_bounds__(256) constrained_stats_kernel(const uint32_t* __restrict__ counts, const uint16_t* __restrict__ pn, uint64_t np, uint32_t threshold,
                                                                uint32_t* __restrict__ flags, PatternMetaStats* __restrict__ st, DeviceStats* __restrict__ ds) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool     seen = false, kept = false;
    uint32_t c = 0, n = 0;
    if (i < np) {
        c        = counts[i];
        n        = pn[i];
        seen     = c > 0;
        kept     = c >= threshold;
        flags[i] = kept ? 1u : 0u;
    }
    const uint32_t mseen = __ballot_sync(0xffffffffu, seen), mkept = __ballot_sync(0xffffffffu, kept);
    if (lane_id() == 0) {
        if (mseen) atomicAdd(&ds->found, (unsigned long long)__popc(mseen));
        if (mkept) atomicAdd(&ds->kept, (unsigned long long)__popc(mkept));
    }
    if (kept) {
        atomicAdd(&ds->kept_occ, (unsigned long long)c);
        atomicAdd(&st->kept_occ_n[min(n, 255u)], (unsigned long long)c);
        atomicAdd(&st->kept_n[min(n, 255u)], 1ull);
        atomicMax(&st->kept_maxn, n);
        atomicMin(&st->kept_minn, n);
    }
}

// load-time filter (PatternMapStore::read, include/patternstore.h:574-586): category switches, length window, occurrence
// threshold, membership in the constraint store
__global__ void __launch_bounds__(256) load_filter_kernel(const uint16_t* __restrict__ pn, const uint8_t* __restrict__ pcat, const uint32_t* __restrict__ counts,
                                                          const uint32_t* __restrict__ constrain_idx1, uint64_t np, uint32_t mintokens, uint32_t minlength, uint32_t maxlength,
                                                          int dongrams, int doskipgrams, int doflexgrams, uint32_t* __restrict__ flags, PatternMetaStats* __restrict__ st) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const uint32_t n = pn[i], cat = pcat[i];
    bool keep = !((!dongrams && cat == 0) || (!doskipgrams && cat == 1) || (!doflexgrams && cat == 2));
    keep      = keep && n >= minlength && n <= maxlength && counts[i] >= mintokens;
    if (keep && constrain_idx1 != nullptr) keep = constrain_idx1[i] != 0;
    flags[i] = keep ? 1u : 0u;
    if (keep) {
        atomicMax(&st->kept_maxn, n);
        atomicMin(&st->kept_minn, n);
        if (cat == 1) st->kept_hasskip = 1;
        if (cat == 2) st->kept_hasflex = 1;
        atomicAdd(&st->kept_n[min(n, 255u)], 1ull);
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
int launch_pattern_meta(cudaStream_t s, const uint8_t* keys, const uint64_t* off, uint64_t np, uint16_t* pn, uint8_t* pcat, PatternMetaStats* st) {
    if (!np) return 0;
    pattern_meta_kernel<<<pi_div_up(np, 256), 256, 0, s>>>(keys, off, np, pn, pcat, st);
    return 1;
}
int launch_index_build(cudaStream_t s, const uint8_t* keys, const uint64_t* off, uint64_t np, unsigned long long* slots, uint64_t cap_pow2, PatternMetaStats* st) {
    if (!np) return 0;
    index_build_kernel<<<pi_div_up(np, 256), 256, 0, s>>>(keys, off, np, slots, cap_pow2 - 1, st);
    return 1;
}
int launch_index_lookup(cudaStream_t s, const uint8_t* qkeys, const uint64_t* qoff, uint64_t nq, const uint8_t* keys, const uint64_t* off, const unsigned long long* slots,
                        uint64_t cap_pow2, uint32_t* out_idx1) {
    if (!nq) return 0;
    index_lookup_kernel<<<pi_div_up(nq, 256), 256, 0, s>>>(qkeys, qoff, nq, keys, off, slots, cap_pow2 - 1, out_idx1);
    return 1;
}
int launch_gather_counts(cudaStream_t s, const uint32_t* idx1, uint64_t nq, const uint32_t* counts, uint32_t* out) {
    if (!nq) return 0;
    gather_counts_kernel<<<pi_div_up(nq, 256), 256, 0, s>>>(idx1, nq, counts, out);
    return 1;
}
int launch_constrained_match(cudaStream_t s, const uint32_t* tok, uint64_t npos, int n, const uint8_t* keys, const uint64_t* off, const unsigned long long* slots, uint64_t cap_pow2,
                             uint32_t* counts, uint32_t* match, DeviceStats* st, int sms) {
    if (!npos) return 0;
    const unsigned grid = (unsigned)std::min<uint64_t>(pi_div_up(npos, 256), (uint64_t)sms * 8 * 4);
    if (n * 5 < 8 * 4)  // every window of n tokens fits 31 bytes: registers only
        constrained_match_kernel<4><<<grid, 256, 0, s>>>(tok, npos, n, keys, off, slots, cap_pow2 - 1, counts, match, st);
    else
        constrained_match_kernel<24><<<grid, 256, 0, s>>>(tok, npos, n, keys, off, slots, cap_pow2 - 1, counts, match, st);
    return 1;
}
int launch_constrained_stats(cudaStream_t s, const uint32_t* counts, const uint16_t* pn, uint64_t np, uint32_t threshold, uint32_t* flags, PatternMetaStats* st, DeviceStats* ds) {
    if (!np) return 0;
    constrained_stats_kernel<<<pi_div_up(np, 256), 256, 0, s>>>(counts, pn, np, threshold, flags, st, ds);
    return 1;
}
int launch_load_filter(cudaStream_t s, const uint16_t* pn, const uint8_t* pcat, const uint32_t* counts, const uint32_t* constrain_idx1, uint64_t np, uint32_t mintokens,
                       uint32_t minlength, uint32_t maxlength, int dongrams, int doskipgrams, int doflexgrams, uint32_t* flags, PatternMetaStats* st) {
    if (!np) return 0;
    load_filter_kernel<<<pi_div_up(np, 256), 256, 0, s>>>(pn, pcat, counts, constrain_idx1, np, mintokens, minlength, maxlength, dongrams, doskipgrams, doflexgrams, flags, st);
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

}  // namespace colibri
