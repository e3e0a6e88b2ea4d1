// kernels.cu -- hand-written sm_100a kernels of the PatternModel::train path.
//
// The reference (include/patternmodel.h:880-1345) walks the corpus once per n, builds every n-token window as a
// heap-allocated byte string, hashes it with SpookyV2 and bumps a std::unordered_map entry after looking up the two
// (n-1)-sub-windows in the same map.  On the device the same result is produced with fixed-width work:
//
//   K0  tokenise      bytes -> one u32 class id per token, 0 for a sentence delimiter  (Pattern(istream) src/pattern.cpp:483-587,
//                     bytestoint src/classdecoder.cpp:20-43)
//   K1  unigrams      class-indexed histogram (no hashing needed: classes are dense), threshold -> level-1 ids
//   K2  count_ngrams  level n >= 2: a window at position p is valid iff the surviving (n-1)-grams at p and p+1 both exist
//                     (patternmodel.h:1139-1152); its identity IS the pair of their ids, so the table key is 8 bytes for every n.
//                     One thread per position: coalesced id reads, one 32-byte-sector probe, CAS claim or RED increment.
//   K3  prune/compact threshold scan of the table (prune(), patternmodel.h:2107-2128) + survivor compaction, then
//       relabel       per position: id := slot+1 if the slot survived, else 0  -> input of level n+1
//   K5  export        survivors -> varint pattern bytes in the reference's key format (Pattern(PatternPointer) src/pattern.cpp:873-909)
//
// All of it is integer/byte work bound by HBM sector traffic; there is nothing here for tensor cores.
#include "kernels.h"

#include "device_utils.cuh"

namespace colibri {

__host__ __device__ static inline unsigned div_up(uint64_t a, uint64_t b) {
    return (unsigned)((a + b - 1) / b);
}
static inline uint64_t umin64(uint64_t a, uint64_t b) {
    return a < b ? a : b;
}

// =============================================================================================
// K0: tokenise
// The staged body sits 16-byte aligned, preceded by 16 zero bytes (so "previous byte" exists for byte 0 and is < 128)
// and padded to a multiple of kTokTile with 0x80 (a continuation byte: never ends a token).
__device__ __forceinline__ uint32_t token_end_bits(uint32_t w) {
    return ~w & 0x80808080u;  // bit 7 of each byte set where byte < 128
}

__global__ void __launch_bounds__(256) tokenise_count_kernel(const uint8_t* __restrict__ corpus, uint32_t* __restrict__ blk_counts) {
    const uint4* src = reinterpret_cast<const uint4*>(corpus + (uint64_t)blockIdx.x * kTokTile);
    uint4        v   = src[threadIdx.x];
    uint32_t     c   = __popc(token_end_bits(v.x)) + __popc(token_end_bits(v.y)) + __popc(token_end_bits(v.z)) + __popc(token_end_bits(v.w));
    __shared__ uint64_t scratch[8];
    uint64_t            total = block_reduce_sum(c, scratch);
    if (threadIdx.x == 0) blk_counts[blockIdx.x] = (uint32_t)total;
}

// exclusive scan of the per-tile counts, in place, by one block (at most a few hundred thousand entries)
// accumulate: the scan starts at *total (the tiles of one chunk of a corpus that is still being copied; *total carries over to the next chunk)
__global__ void __launch_bounds__(1024) scan_block_counts_kernel(uint32_t* __restrict__ counts, uint32_t n, unsigned long long* __restrict__ total, bool accumulate) {
    __shared__ uint64_t part[1024];
    uint32_t            chunk = (n + blockDim.x - 1) / blockDim.x;
    uint32_t            lo = threadIdx.x * chunk, hi = min(n, lo + chunk);
    uint64_t            s = 0;
    for (uint32_t i = lo; i < hi; ++i) s += counts[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t acc = accumulate ? *total : 0;
        for (uint32_t i = 0; i < blockDim.x; ++i) {
            uint64_t t = part[i];
            part[i]    = acc;
            acc += t;
        }
        *total = acc;
    }
    __syncthreads();
    uint64_t acc = part[threadIdx.x];
    for (uint32_t i = lo; i < hi; ++i) {
        uint32_t t = counts[i];
        counts[i]  = (uint32_t)acc;  // caller guarantees the total fits 32 bits
        acc += t;
    }
}

__global__ void __launch_bounds__(256) tokenise_write_kernel(const uint8_t* __restrict__ corpus, const uint32_t* __restrict__ blk_offsets, uint32_t* __restrict__ tok,
                                                             DeviceStats* __restrict__ st) {
    __shared__ __align__(16) uint8_t tile[16 + kTokTile];
    __shared__ uint32_t staged[kTokTile];  // the tile's tokens in order (a token is at least one byte): written out coalesced below
    __shared__ uint32_t warp_tot[8];
    __shared__ uint64_t scratch[8];
    const uint8_t* base = corpus + (uint64_t)blockIdx.x * kTokTile;
    uint4          v    = reinterpret_cast<const uint4*>(base)[threadIdx.x];
    reinterpret_cast<uint4*>(tile + 16)[threadIdx.x] = v;
    if (threadIdx.x == 0) *reinterpret_cast<uint4*>(tile) = *reinterpret_cast<const uint4*>(base - 16);
    uint32_t w[4]  = {v.x, v.y, v.z, v.w};
    uint32_t ends  = 0;  // bit j set: byte j of this thread's 16 ends a token
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t e = token_end_bits(w[k]);
        ends |= (((e >> 7) & 1u) | ((e >> 14) & 2u) | ((e >> 21) & 4u) | ((e >> 28) & 8u)) << (4 * k);
    }
    uint32_t c    = __popc(ends);
    uint32_t incl = warp_inclusive_scan(c);
    if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (uint32_t i = 0; i < 8; ++i) {
        const uint32_t t = warp_tot[i];
        woff += i < (threadIdx.x >> 5) ? t : 0u;
        total += t;
    }
    uint32_t at = woff + (incl - c);

    uint32_t ntok = 0, maxc = 0, err = 0;
    const uint8_t* mine = tile + 16 + threadIdx.x * 16;
    while (ends) {
        int j = __ffs(ends) - 1;
        ends &= ends - 1;
        const uint8_t* q = mine + j;
        uint64_t       val = q[0];
        int            len = 1;
        while (len < 6 && q[-len] >= 128) {  // continuation bytes precede the final byte; little-endian base 128
            val = (val << 7) | (q[-len] & 0x7Fu);
            ++len;
        }
        if (len == 6 || val > 0xFFFFFFFFull) err |= kErrTokenTooLong;
        if (len > 1 && q[0] == 0) err |= kErrNonCanonical;
        uint32_t cls = (uint32_t)val;
        if (cls == 3 || cls == 4) err |= kErrReservedClass;
        staged[at++] = cls;
        ntok += cls != 0;
        maxc = max(maxc, cls);
    }
    __syncthreads();
    uint32_t* dst = tok + blk_offsets[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < total; i += 256) dst[i] = staged[i];
    uint64_t tot = block_reduce_sum(ntok, scratch);
    maxc         = warp_reduce_max(maxc);
    if (threadIdx.x == 0 && tot) atomicAdd(&st->totaltokens, (unsigned long long)tot);
    if (lane_id() == 0 && maxc) atomicMax(&st->maxclass, maxc);
    if (err) atomicOr(&st->errflags, err);
}

int launch_tokenise_count(cudaStream_t s, const uint8_t* corpus, uint64_t, uint32_t* blk_counts, uint32_t nblocks) {
    tokenise_count_kernel<<<nblocks, 256, 0, s>>>(corpus, blk_counts);
    return 1;
}
int launch_scan_block_counts(cudaStream_t s, uint32_t* blk_counts, uint32_t nblocks, unsigned long long* total, bool accumulate) {
    scan_block_counts_kernel<<<1, 1024, 0, s>>>(blk_counts, nblocks, total, accumulate);
    return 1;
}
int launch_tokenise_write(cudaStream_t s, const uint8_t* corpus, uint64_t, const uint32_t* blk_offsets, uint32_t nblocks, uint32_t* tok, DeviceStats* st) {
    tokenise_write_kernel<<<nblocks, 256, 0, s>>>(corpus, blk_offsets, tok, st);
    return 1;
}

// =============================================================================================
// K1: unigrams.  Classes are dense (the encoder hands out ids by descending frequency, src/classencoder.cpp:213-226),
// so the "hash table" of level 1 is a plain array indexed by class.  The head of a Zipf distribution would serialise
// global atomics on a handful of addresses, so each block first counts the most frequent kHotClasses in shared memory.
constexpr uint32_t kHotClasses = 16384;  // 64 KB of shared memory per block

__global__ void __launch_bounds__(512) unigram_hist_kernel(const uint32_t* __restrict__ tok, uint64_t npos, uint32_t* __restrict__ count1) {
    extern __shared__ uint32_t hot[];
    for (uint32_t i = threadIdx.x; i < kHotClasses; i += blockDim.x) hot[i] = 0;
    __syncthreads();
    const uint64_t nvec   = npos / 4;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        uint4    v    = __ldcs(reinterpret_cast<const uint4*>(tok) + i);
        uint32_t c[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (c[k] == 0) continue;
            if (c[k] < kHotClasses)
                atomicAdd(&hot[c[k]], 1u);
            else
                atomicAdd(&count1[c[k]], 1u);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (npos & 3)) {
        uint32_t c = tok[nvec * 4 + threadIdx.x];
        if (c) atomicAdd(&count1[c], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < kHotClasses; i += blockDim.x)
        if (hot[i]) atomicAdd(&count1[i], hot[i]);
}

// prune(MINTOKENS, 1) (patternmodel.h:2107-2128) over the class array + totaltypes (:1199-1201, counted BEFORE pruning)
// (part_mod, part_rem): in multi-GPU mode every rank sees the same global counts and exports the classes c with c % part_mod == part_rem
__global__ void __launch_bounds__(256) unigram_prune_kernel(const uint32_t* __restrict__ count1, uint32_t nclasses, uint32_t threshold, uint32_t* __restrict__ sv_pos,
                                                            uint32_t* __restrict__ sv_count, uint64_t sv_base, DeviceStats* __restrict__ st, uint32_t part_mod, uint32_t part_rem,
                                                            uint32_t* __restrict__ class_index /* may be NULL: class -> survivor index + 1 */) {
    __shared__ uint64_t scratch[8];
    uint64_t found = 0, kept = 0, occ = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (uint64_t)div_up(nclasses, 32) * 32; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t c    = i < nclasses ? count1[i] : 0;
        bool     keep = c >= threshold && c > 0;
        bool     mine = keep && (uint32_t)(i % part_mod) == part_rem;
        found += c > 0;
        uint64_t idx = warp_aggregated_inc(&st->cursor, mine);
        if (mine) {
            sv_pos[sv_base + idx]   = (uint32_t)i;  // level 1 survivors carry the class id instead of a position
            sv_count[sv_base + idx] = c;
        }
        if (class_index != nullptr && i < nclasses) class_index[i] = mine ? (uint32_t)idx + 1 : 0u;
        if (keep) {
            ++kept;
            occ += c;
        }
    }
    found = block_reduce_sum(found, scratch);
    kept  = block_reduce_sum(kept, scratch);
    occ   = block_reduce_sum(occ, scratch);
    if (threadIdx.x == 0) {
        if (found) atomicAdd(&st->found, (unsigned long long)found);
        if (kept) atomicAdd(&st->kept, (unsigned long long)kept);
        if (occ) atomicAdd(&st->kept_occ, (unsigned long long)occ);
    }
}

// level-1 id of a position: its class if that unigram survived (and meets MINTOKENS_UNIGRAMS, patternmodel.h:1094-1104), else 0
__global__ void __launch_bounds__(256) make_id1_kernel(const uint32_t* __restrict__ tok, uint64_t npos, const uint32_t* __restrict__ count1, uint32_t threshold,
                                                       uint32_t* __restrict__ id1) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npos) return;
    uint32_t c = tok[i];
    id1[i]     = (c != 0 && __ldg(&count1[c]) >= threshold) ? c : 0;
}

int launch_unigram_hist(cudaStream_t s, const uint32_t* tok, uint64_t npos, uint32_t* count1, uint32_t, int sms) {
    // the opt-in is per device (a process may train on several): set it on every call, it costs nothing next to the launch
    cudaFuncSetAttribute(unigram_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHotClasses * 4);
    unigram_hist_kernel<<<sms * 3, 512, kHotClasses * 4, s>>>(tok, npos, count1);
    return 1;
}
int launch_unigram_prune(cudaStream_t s, const uint32_t* count1, uint32_t nclasses, uint32_t threshold, uint32_t* sv_pos, uint32_t* sv_count, uint64_t sv_base, DeviceStats* st,
                         uint32_t part_mod, uint32_t part_rem, uint32_t* class_index) {
    unsigned grid = min(div_up(nclasses, 256), 148u * 8u);
    unigram_prune_kernel<<<grid, 256, 0, s>>>(count1, nclasses, threshold, sv_pos, sv_count, sv_base, st, part_mod ? part_mod : 1u, part_rem, class_index);
    return 1;
}
int launch_make_id1(cudaStream_t s, const uint32_t* tok, uint64_t npos, const uint32_t* count1, uint32_t threshold, uint32_t* id1) {
    make_id1_kernel<<<div_up(npos, 256), 256, 0, s>>>(tok, npos, count1, threshold, id1);
    return 1;
}

// =============================================================================================
// K2: the n-gram upsert kernel (n >= 2).  Replaces, per window: two has() lookups + add() of the reference
// (patternmodel.h:1139-1161), i.e. 3 heap allocations, 3-4 SpookyV2 hashes and 3 chained-bucket walks.
//
// prev[p] = id of the surviving (n-1)-gram starting at position p, 0 if there is none (pruned, or the window
// would cross a sentence delimiter).  Window p of size n is valid iff prev[p] and prev[p+1] are both non-zero,
// and two valid windows are the same n-gram iff they agree on that pair: the pair is the key.
// 16-byte compare-and-swap against the all-zero (empty) slot: atom.global.cas.b128 (sm_90+)
__device__ __forceinline__ void cas128(void* addr, unsigned long long new0, unsigned long long new1, unsigned long long& old0, unsigned long long& old1) {
    asm volatile(
        "{\n\t"
        ".reg .b128 cmp, val, old;\n\t"
        "mov.b128 cmp, {%3, %3};\n\t"
        "mov.b128 val, {%4, %5};\n\t"
        "atom.global.cas.b128 old, [%2], cmp, val;\n\t"
        "mov.b128 {%0, %1}, old;\n\t"
        "}"
        : "=l"(old0), "=l"(old1)
        : "l"(addr), "l"(0ull), "l"(new0), "l"(new1)
        : "memory");
}

constexpr uint64_t kMaxProbe = 8192;  // a probe run this long means the table was sized too small: report, the host retries bigger

__device__ __forceinline__ uint32_t upsert_ngram_at(NgramSlot* __restrict__ table, uint64_t cap, uint64_t slot, unsigned long long key, uint32_t pos, uint32_t& probes, bool& full) {
    const uint64_t limit = cap < kMaxProbe ? cap : kMaxProbe;
    for (uint64_t step = 0; step < limit; ++step) {
        NgramSlot*         s   = table + slot;
        unsigned long long cur = __ldcg(&s->key);  // keys never change once set, so a stale "empty" is the only possible staleness
        ++probes;
        if (cur == 0) {
            // claim the whole slot at once: {key, count = 1, pos} -- one 128-bit CAS instead of CAS + store + RED
            unsigned long long o0, o1;
            cas128(s, key, 1ull | ((unsigned long long)pos << 32), o0, o1);
            if (o0 == 0) return (uint32_t)slot + 1;
            cur = o0;
        }
        if (cur == key) {
            atomicAdd(&s->count, 1u);  // result unused -> RED
            return (uint32_t)slot + 1;
        }
        slot = slot + 1 == cap ? 0 : slot + 1;
    }
    full = true;
    return 0;
}
__device__ __forceinline__ uint32_t upsert_ngram(NgramSlot* __restrict__ table, uint64_t cap, unsigned long long key, uint32_t pos, uint32_t& probes, bool& full) {
    return upsert_ngram_at(table, cap, fast_range(table_hash_u64(key), cap), key, pos, probes, full);
}

// ---- the occurrence filter (MINTOKENS >= 2 only).  Most distinct n-grams of a corpus occur once and are pruned right
// after the pass (86 % of the bigrams of a Zipf corpus).  A first streaming pass counts every valid window into a
// 2-bit saturating counter per hash bucket (bit 0: seen, bit 1: seen twice); the array is small enough to live in L2.
// The counting pass then sends a window to the HBM table only if its bucket was hit at least twice.  A window whose
// bucket was hit once is the ONLY window of its key, so it is a distinct n-gram with count 1: it is counted as "found"
// and as "pruned" without ever touching the table.  Keys that reach the table are counted exactly as before, so the
// surviving patterns, their counts and the found/pruned statistics are unchanged.
__device__ __forceinline__ uint32_t ld_cached(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void filter_locate(uint64_t h, uint64_t nbuckets_mask, uint64_t& word, uint32_t& shift) {
    uint64_t bucket = h & nbuckets_mask;  // low hash bits; the table slot uses the high bits (fast_range)
    word            = bucket >> 4;
    shift           = (uint32_t)(bucket & 15) * 2;
}

// Two ways to enumerate the windows of a level (template parameter kList):
//   dense   item j IS position j; prev[] is streamed (levels whose (n-1)-grams cover most of the corpus)
//   list    item j is position list[j]: the positions whose (n-1)-gram survived, written by the previous level's relabel step.
//           Higher levels of a natural corpus are sparse (Zipf 100 M tokens: 18 % of the positions carry a surviving trigram, 4 % a
//           4-gram), and a dense pass still pays the id stream, the hash of every warp that holds one valid lane, and the id write
//           for ALL positions.  In list mode cur[] is zeroed by a memset and written only where a window exists.
template <bool kList>
__device__ __forceinline__ bool load_window(const uint32_t* __restrict__ prev, const uint32_t* __restrict__ list, uint64_t j, uint64_t& p, uint32_t& a, uint32_t& b) {
    if (kList) {
        p = __ldcs(list + j);
        a = __ldg(prev + p);  // non-zero by construction, but read anyway: it is half of the key
    } else {
        p = j;
        a = __ldcs(prev + p);
    }
    b = __ldg(prev + p + 1);  // prev has npos + 1 readable entries, the last one 0
    return a != 0 && b != 0;
}

template <bool kList>
__global__ void __launch_bounds__(256) ngram_filter_kernel(const uint32_t* __restrict__ prev, const uint32_t* __restrict__ list, uint64_t nitems, uint32_t* __restrict__ filter,
                                                           uint64_t nbuckets_mask, DeviceStats* __restrict__ st, const uint32_t dense) {
    __shared__ uint64_t scratch[8];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t       valid = 0, twice = 0;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nitems; j += stride) {
        uint64_t p;
        uint32_t a, b;
        if (!load_window<kList>(prev, list, j, p, a, b)) continue;
        ++valid;
        if (a < dense && b < dense) continue;  // a pair of frequent classes has its own directly addressed slot (see count_ngrams_kernel)
        uint64_t word;
        uint32_t shift;
        filter_locate(table_hash_u64(((unsigned long long)a << 32) | b), nbuckets_mask, word, shift);
        // Look before touching: filter bits only ever go from 0 to 1 during this launch, so a copy cached in L1 can be stale in one
        // direction only (it may miss bits that are set by now) and then the atomics below fetch the truth.  The frequent keys of a
        // Zipf corpus therefore stop at an L1 hit instead of an L2 round trip.
        uint32_t bits = (ld_cached(filter + word) >> shift) & 3u;
        if (bits == 3u) continue;  // already saturated
        if ((bits & 1u) == 0) {
            uint32_t old = atomicOr(filter + word, 1u << shift);
            if (((old >> shift) & 1u) == 0) continue;  // first hit of this bucket
            bits = (old >> shift) & 3u;
        }
        if ((bits & 2u) == 0) {
            uint32_t old = atomicOr(filter + word, 2u << shift);
            twice += ((old >> shift) & 2u) == 0;  // this thread moved the bucket to "seen twice"
        }
    }
    uint64_t v  = block_reduce_sum(valid, scratch);
    uint64_t tw = block_reduce_sum(twice, scratch);
    if (threadIdx.x == 0) {
        if (v) atomicAdd(&st->valid_windows, (unsigned long long)v);
        if (tw) atomicAdd(&st->found, (unsigned long long)tw);  // buckets hit at least twice
    }
}

// Hot keys: the most frequent n-grams of a corpus put hundreds of thousands to millions of REDs on one L2 address, where
// they serialise.  Each block keeps a small never-evicting cache in shared memory: key -> {table slot, pending count}.  The
// first window of a key that finds its line empty installs it after its normal upsert; later windows of that key in the
// block only bump the shared-memory counter, which is added to the table once when the block retires.  A line never
// changes owner, so counts stay exact; cold keys pay one shared-memory read.
constexpr uint32_t           kHotLines = 1024;
constexpr unsigned long long kHotBusy  = ~0ull;

// Dense pairs (level 2 only, where the ids ARE the class numbers and classes are ranked by frequency, src/classencoder.cpp:213-226):
// a bigram of two classes below `dense` is counted in dense_cnt[a * dense + b], a plain u32 square that follows the occurrence filter
// in one buffer (2048^2 x 4 B = 16 MB; round 1 used full 16-byte slots: 64 MB of random L2 traffic that pushed the filter out).  No
// hash, no filter word, no key to compare, no probing: one RED.  With dense = 2048 that is 46 % of the bigram windows of a Zipf corpus.
// Its id is cap + a * dense + b + 1, as if the square were appended to the hashed table: the survivor bitmap, the relabel step and the
// forward index see nothing special.  A surviving pair needs "a position where it occurs" for the export: prune_dense_kernel writes the
// two class ids into the spare room behind the token array and hands out that position.

template <bool kFilter, bool kDense, bool kList>
__global__ void __launch_bounds__(256, kDense ? 7 : 8) count_ngrams_kernel(const uint32_t* __restrict__ prev, const uint32_t* __restrict__ list, uint64_t nitems,
                                                                           uint32_t* __restrict__ cur, NgramSlot* __restrict__ table, uint64_t cap,
                                                                           const uint32_t* __restrict__ filter, uint64_t nbuckets_mask, DeviceStats* __restrict__ st, const bool hot,
                                                                           const uint32_t dense, uint32_t* __restrict__ dense_cnt, const bool onebit) {
    __shared__ uint64_t scratch[8];
    __shared__ unsigned long long hot_key[kHotLines];
    __shared__ uint32_t hot_slot[kHotLines];
    __shared__ uint32_t hot_pending[kHotLines];
    if (hot) {
        for (uint32_t i = threadIdx.x; i < kHotLines; i += blockDim.x) {
            hot_key[i]     = 0;
            hot_pending[i] = 0;
        }
        __syncthreads();
    }
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t       valid = 0, probes = 0, singles = 0;
    bool           full = false;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nitems; j += stride) {
        uint64_t p;
        uint32_t a, b, id = 0;
        if (load_window<kList>(prev, list, j, p, a, b)) {
            ++valid;
            const unsigned long long key = ((unsigned long long)a << 32) | b;
            if (kDense && a < dense && b < dense) {
                // the pair owns slot cap + a * dense + b (the dense square follows the hashed part of the table)
                const uint32_t     line   = (a * 40503u + b) & (kHotLines - 1);
                unsigned long long cached = ~0ull;
                if (hot) cached = *(volatile unsigned long long*)&hot_key[line];
                if (cached == key) {
                    atomicAdd(&hot_pending[line], 1u);
                    id = *(volatile uint32_t*)&hot_slot[line];
                } else {
                    const uint32_t cell = a * dense + b;
                    atomicAdd(dense_cnt + cell, 1u);  // result unused -> RED
                    id = (uint32_t)cap + cell + 1;
                    if (cached == 0 && atomicCAS(&hot_key[line], 0ull, kHotBusy) == 0ull) {
                        hot_slot[line] = id;
                        __threadfence_block();
                        *(volatile unsigned long long*)&hot_key[line] = key;
                    }
                }
            } else {
                const uint64_t h  = table_hash_u64(key);
                bool           go = true;
                if (kFilter) {
                    uint64_t word;
                    uint32_t shift;
                    if (onebit) {  // the "hit twice" bits alone, one per bucket: half the footprint of the 2-bit counters (launch_filter_to_bitmap)
                        const uint64_t bucket = h & nbuckets_mask;
                        go = ((__ldg(filter + (bucket >> 5)) >> (bucket & 31)) & 1u) != 0;
                    } else {
                        filter_locate(h, nbuckets_mask, word, shift);
                        go = ((__ldg(filter + word) >> shift) & 2u) != 0;
                    }
                    singles += !go;
                }
                if (go) {
                    const uint32_t     line   = (uint32_t)(h >> 20) & (kHotLines - 1);
                    unsigned long long cached = ~0ull;
                    if (hot) cached = *(volatile unsigned long long*)&hot_key[line];
                    if (cached == key) {  // its slot was published before the key (below)
                        atomicAdd(&hot_pending[line], 1u);
                        id = *(volatile uint32_t*)&hot_slot[line];
                    } else {
                        id = upsert_ngram_at(table, cap, fast_range(h, cap), key, (uint32_t)p, probes, full);
                        if (id != 0 && cached == 0 && atomicCAS(&hot_key[line], 0ull, kHotBusy) == 0ull) {
                            hot_slot[line] = id;  // publish the slot ...
                            __threadfence_block();
                            *(volatile unsigned long long*)&hot_key[line] = key;  // ... then the key that makes it visible
                        }
                    }
                }
            }
        }
        if (!kList)
            __stcs(cur + p, id);
        else if (id != 0)
            cur[p] = id;  // cur was zeroed by the host
    }
    if (hot) {
        __syncthreads();
        for (uint32_t l = threadIdx.x; l < kHotLines; l += blockDim.x) {
            uint32_t c = hot_pending[l];
            if (c) {
                const uint64_t slot = hot_slot[l] - 1;
                if (kDense && slot >= cap) atomicAdd(dense_cnt + (slot - cap), c);
                else atomicAdd(&table[slot].count, c);
            }
        }
    }
    uint64_t v  = block_reduce_sum(valid, scratch);
    uint64_t pr = block_reduce_sum(probes, scratch);
    uint64_t sg = block_reduce_sum(singles, scratch);
    if (threadIdx.x == 0) {
        if (v) atomicAdd(&st->valid_windows, (unsigned long long)v);
        if (pr) atomicAdd(&st->probes, (unsigned long long)pr);
        if (sg) atomicAdd(&st->singletons, (unsigned long long)sg);
    }
    if (full) atomicOr(&st->errflags, kErrTableFull);
}

// K3: prune(MINTOKENS, n) as a table scan: statistics, compaction of the survivors, and a 1-bit-per-slot survivor
// bitmap (cap/8 bytes: L2 resident) that the relabel step tests instead of going back to the table in HBM.
// A block handles tiles of 2048 slots: each warp reads 8 x 32 consecutive slots (coalesced 512-byte loads), the block
// reserves its output range with ONE atomicAdd per tile (a per-warp cursor atomic serialises on a single L2 address).
constexpr int kPruneTile = 2048;

template <class Slot>
__device__ __forceinline__ void load_slot(const Slot* table, uint64_t i, uint64_t cap, uint32_t& k_lo, uint32_t& k_hi, uint32_t& count, uint32_t& pos);
template <>
__device__ __forceinline__ void load_slot<NgramSlot>(const NgramSlot* table, uint64_t i, uint64_t cap, uint32_t& k_lo, uint32_t& k_hi, uint32_t& count, uint32_t& pos) {
    uint4 raw = make_uint4(0, 0, 0, 0);
    if (i < cap) raw = __ldcs(reinterpret_cast<const uint4*>(table) + i);
    k_lo = raw.x; k_hi = raw.y; count = raw.z; pos = raw.w;
}
template <>
__device__ __forceinline__ void load_slot<SkipSlot>(const SkipSlot* table, uint64_t i, uint64_t cap, uint32_t& k_lo, uint32_t& k_hi, uint32_t& count, uint32_t& pos) {
    uint4 lo = make_uint4(0, 0, 0, 0), hi = make_uint4(0, 0, 0, 0);
    if (i < cap) {
        lo = __ldcs(reinterpret_cast<const uint4*>(table + i));
        hi = __ldcs(reinterpret_cast<const uint4*>(table + i) + 1);
    }
    k_lo = lo.x; k_hi = lo.y; count = hi.x; pos = hi.y;
}

template <class Slot, bool kSkip>
__global__ void __launch_bounds__(256) prune_table_kernel(const Slot* __restrict__ table, uint64_t cap, uint32_t threshold, uint32_t* __restrict__ sv_pos,
                                                          uint32_t* __restrict__ sv_count, uint32_t* __restrict__ sv_mask, uint32_t* __restrict__ bitmap,
                                                          uint32_t* __restrict__ slot_index /* may be NULL: slot -> survivor index + 1 */, DeviceStats* __restrict__ st,
                                                          const uint32_t* __restrict__ types /* may be NULL */, uint32_t mintypes) {
    __shared__ uint64_t scratch[8];
    __shared__ uint32_t warp_cnt[8];
    __shared__ unsigned long long tile_base;
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint64_t ntiles = (cap + kPruneTile - 1) / kPruneTile;
    uint64_t found = 0, kept = 0, occ = 0;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t base = tile * kPruneTile + (uint64_t)warp * 256;
        uint32_t pos[8], cnt[8], msk[8], keepbits[8];
        uint32_t wtotal = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t k_lo, k_hi;
            load_slot<Slot>(table, base + k * 32 + lane, cap, k_lo, k_hi, cnt[k], pos[k]);
            bool used = (k_lo | k_hi) != 0;
            if (kSkip) used = used && (k_hi & kSkipCombiner) == 0;  // helper entries are ids, not patterns
            bool keep = used && cnt[k] >= threshold;
            if (types != nullptr && keep) keep = __ldg(types + base + k * 32 + lane) >= mintypes;  // skip-type rule of indexed models
            msk[k]    = k_hi & 0x00FFFFFFu;                         // skipgram: high word of k0 = gap mask (+ round bits, dropped)
            keepbits[k] = __ballot_sync(0xffffffffu, keep);
            found += used;
            if (keep) {
                ++kept;
                occ += cnt[k];
            }
            wtotal += __popc(keepbits[k]);
        }
        if (bitmap != nullptr && lane < 8) {
            uint64_t word = (base >> 5) + lane;  // 8 consecutive words per warp: one 32-byte store
            uint32_t bits = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) bits = lane == (uint32_t)k ? keepbits[k] : bits;
            if (word * 32 < (cap + 31) / 32 * 32) bitmap[word] = bits;
        }
        if (lane == 0) warp_cnt[warp] = wtotal;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < 8; ++w) {
                uint32_t c  = warp_cnt[w];
                warp_cnt[w] = tot;
                tot += c;
            }
            tile_base = (tot && sv_pos != nullptr) ? atomicAdd(&st->cursor, (unsigned long long)tot) : 0ull;
        }
        __syncthreads();
        uint64_t out = tile_base + warp_cnt[warp];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const bool mine = (keepbits[k] >> lane) & 1u;
            uint64_t   idx  = out + __popc(keepbits[k] & ((1u << lane) - 1));
            if (sv_pos != nullptr && mine) {
                sv_pos[idx]   = pos[k];
                sv_count[idx] = cnt[k];
                if (kSkip) sv_mask[idx] = msk[k];
            }
            if (slot_index != nullptr && base + k * 32 + lane < cap) slot_index[base + k * 32 + lane] = mine ? (uint32_t)idx + 1 : 0u;
            out += __popc(keepbits[k]);
        }
        __syncthreads();  // warp_cnt / tile_base are reused by the next tile
    }
    found = block_reduce_sum(found, scratch);
    kept  = block_reduce_sum(kept, scratch);
    occ   = block_reduce_sum(occ, scratch);
    if (threadIdx.x == 0) {
        if (found) atomicAdd(&st->found, (unsigned long long)found);
        if (kept) atomicAdd(&st->kept, (unsigned long long)kept);
        if (occ) atomicAdd(&st->kept_occ, (unsigned long long)occ);
    }
}

__device__ __forceinline__ uint64_t block_reserve(uint32_t c, unsigned long long* __restrict__ cursor, uint32_t* warp_tot, unsigned long long* base_smem);
// prune(MINTOKENS, 2) over the dense square: the same outputs as prune_table_kernel (statistics, survivors, survivor bitmap at bit
// cap + cell, slot -> survivor index) for the cells of dense_cnt.  A survivor's two class ids go to tok_ext[2 * cell ..] (the spare room
// behind the token array) and its "position" is where they were put, so the export re-encodes it like any other bigram.
__global__ void __launch_bounds__(256) prune_dense_kernel(const uint32_t* __restrict__ dense_cnt, uint32_t dense, uint64_t cap, uint32_t threshold, uint32_t* __restrict__ sv_pos,
                                                          uint32_t* __restrict__ sv_count, uint32_t* __restrict__ bitmap, uint32_t* __restrict__ slot_index,
                                                          uint32_t* __restrict__ tok_ext, uint32_t ext_pos0, DeviceStats* __restrict__ st) {
    __shared__ uint64_t scratch[8];
    __shared__ uint32_t warp_tot[8];
    __shared__ unsigned long long base_smem;
    const uint64_t cells  = (uint64_t)dense * dense;
    const uint64_t rounds = (cells + (uint64_t)gridDim.x * 256 - 1) / ((uint64_t)gridDim.x * 256);
    uint64_t found = 0, kept = 0, occ = 0;
    for (uint64_t r = 0; r < rounds; ++r) {
        const uint64_t i = (r * gridDim.x + blockIdx.x) * 256 + threadIdx.x;  // cap and the cell count are multiples of 32: a warp covers one bitmap word
        const uint32_t c = i < cells ? __ldcs(dense_cnt + i) : 0u;
        const bool keep  = c != 0 && c >= threshold;
        found += c != 0;
        const uint32_t bits = __ballot_sync(0xffffffffu, keep);
        if (lane_id() == 0 && i < cells) bitmap[(cap + i) >> 5] = bits;
        const uint64_t out = block_reserve(keep ? 1u : 0u, &st->cursor, warp_tot, &base_smem);
        if (keep) {
            ++kept;
            occ += c;
            sv_pos[out]        = ext_pos0 + 2u * (uint32_t)i;
            sv_count[out]      = c;
            tok_ext[2 * i]     = (uint32_t)(i / dense);
            tok_ext[2 * i + 1] = (uint32_t)(i % dense);
        }
        if (slot_index != nullptr && i < cells) slot_index[cap + i] = keep ? (uint32_t)out + 1 : 0u;
    }
    found = block_reduce_sum(found, scratch);
    kept  = block_reduce_sum(kept, scratch);
    occ   = block_reduce_sum(occ, scratch);
    if (threadIdx.x == 0) {
        if (found) atomicAdd(&st->found, (unsigned long long)found);
        if (kept) atomicAdd(&st->kept, (unsigned long long)kept);
        if (occ) atomicAdd(&st->kept_occ, (unsigned long long)occ);
    }
}

// ---- relabel: after pruning a position keeps its id only if its n-gram survived (bit test in the L2-resident survivor bitmap).
// Three shapes: dense -> dense (in place), dense -> dense + the list of surviving positions (when the next level will run in
// list mode), list -> dense + list.  A block compacts its survivors with one atomicAdd on the list cursor (a per-warp cursor atomic
// serialises on a single L2 address); inside a block the positions stay ascending, so the list is sorted in runs of <= 1024.
__device__ __forceinline__ bool id_survives(uint32_t id, const uint32_t* __restrict__ bitmap) {
    return id != 0 && ((__ldg(bitmap + ((id - 1) >> 5)) >> ((id - 1) & 31)) & 1u) != 0;
}
// every thread contributes `c` (0..4) values; returns the thread's first output index in list_out
__device__ __forceinline__ uint64_t block_reserve(uint32_t c, unsigned long long* __restrict__ cursor, uint32_t* warp_tot, unsigned long long* base_smem) {
    uint32_t incl = warp_inclusive_scan(c);
    if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) {
            uint32_t t  = warp_tot[w];
            warp_tot[w] = tot;
            tot += t;
        }
        *base_smem = tot ? atomicAdd(cursor, (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    uint64_t out = *base_smem + warp_tot[threadIdx.x >> 5] + (incl - c);
    __syncthreads();  // warp_tot / base_smem are reused by the caller's next round
    return out;
}

template <bool kEmit>
__global__ void __launch_bounds__(256) relabel_kernel(uint32_t* __restrict__ cur, uint64_t npos, const uint32_t* __restrict__ bitmap, uint32_t* __restrict__ list_out,
                                                      unsigned long long* __restrict__ cursor) {
    __shared__ uint32_t warp_tot[8];
    __shared__ unsigned long long base_smem;
    uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    uint32_t c[4] = {0, 0, 0, 0};
    uint32_t n = 0;
    if (i + 4 <= npos) {
        uint4 v = *reinterpret_cast<uint4*>(cur + i);
        c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
        bool changed = false;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (c[k] != 0) {
                if (id_survives(c[k], bitmap)) ++n;
                else { c[k] = 0; changed = true; }
            }
        if (changed) *reinterpret_cast<uint4*>(cur + i) = make_uint4(c[0], c[1], c[2], c[3]);
    } else if (i < npos) {
        for (int k = 0; k < 4 && i + k < npos; ++k) {
            uint32_t id = cur[i + k];
            if (id_survives(id, bitmap)) { c[k] = id; ++n; }
            else if (id != 0) cur[i + k] = 0;
        }
    }
    if (kEmit) {
        uint64_t out = block_reserve(n, cursor, warp_tot, &base_smem);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (c[k] != 0) list_out[out++] = (uint32_t)(i + k);
    }
}

__global__ void __launch_bounds__(256) relabel_list_kernel(uint32_t* __restrict__ cur, const uint32_t* __restrict__ list_in, uint64_t nitems, const uint32_t* __restrict__ bitmap,
                                                           uint32_t* __restrict__ list_out, unsigned long long* __restrict__ cursor) {
    __shared__ uint32_t warp_tot[8];
    __shared__ unsigned long long base_smem;
    const uint64_t rounds = (nitems + (uint64_t)gridDim.x * blockDim.x - 1) / ((uint64_t)gridDim.x * blockDim.x);
    for (uint64_t r = 0; r < rounds; ++r) {
        // consecutive blocks take consecutive 256-item tiles of a round, so the output keeps the input's rough order
        const uint64_t j = (r * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
        uint32_t p = 0;
        bool     keep = false;
        if (j < nitems) {
            p = __ldcs(list_in + j);
            const uint32_t id = cur[p];
            keep = id_survives(id, bitmap);
            if (id != 0 && !keep) cur[p] = 0;
        }
        if (list_out != nullptr) {  // uniform
            uint64_t out = block_reserve(keep ? 1u : 0u, cursor, warp_tot, &base_smem);
            if (keep) list_out[out] = p;
        }
    }
}

static int blocks_per_sm(const void* fn, int threads, size_t smem) {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, threads, smem);
    return n > 0 ? n : 1;
}

int launch_ngram_filter(cudaStream_t s, const uint32_t* prev, uint64_t npos, uint32_t* filter, uint64_t nbuckets, DeviceStats* st, int sms, uint32_t dense, const uint32_t* list,
                        uint64_t nlist) {
    const uint64_t nitems = list ? nlist : npos;
    if (!nitems) return 0;
    static int bps  = blocks_per_sm((const void*)ngram_filter_kernel<false>, 256, 0);
    unsigned   grid = (unsigned)umin64(div_up(nitems, 256), (uint64_t)sms * bps * 4);
    if (list)
        ngram_filter_kernel<true><<<grid, 256, 0, s>>>(prev, list, nitems, filter, nbuckets - 1, st, dense);
    else
        ngram_filter_kernel<false><<<grid, 256, 0, s>>>(prev, nullptr, nitems, filter, nbuckets - 1, st, dense);
    return 1;
}
template <bool kFilter, bool kDense, bool kList>
static void launch_count_variant(cudaStream_t s, const uint32_t* prev, const uint32_t* list, uint64_t nitems, uint32_t* cur, NgramSlot* table, uint64_t cap, const uint32_t* filter,
                                 uint64_t nbuckets, DeviceStats* st, int sms, bool hot, uint32_t dense, uint32_t* dense_cnt = nullptr, bool onebit = false) {
    static int bps  = blocks_per_sm((const void*)count_ngrams_kernel<kFilter, kDense, kList>, 256, 0);
    unsigned   grid = (unsigned)umin64(div_up(nitems, 256), (uint64_t)sms * bps * 4);
    count_ngrams_kernel<kFilter, kDense, kList><<<grid, 256, 0, s>>>(prev, list, nitems, cur, table, cap, filter, kFilter ? nbuckets - 1 : 0, st, hot, dense, dense_cnt, onebit);
}
// the "hit twice" bit of every 2-bit counter, packed: bitmap[bucket >> 5] bit (bucket & 31); two filter words in, one bitmap word out
__global__ void __launch_bounds__(256) filter_to_bitmap_kernel(const uint32_t* __restrict__ filter, uint64_t nwords_out, uint32_t* __restrict__ bitmap) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwords_out) return;
    const uint2 in = __ldcs(reinterpret_cast<const uint2*>(filter) + w);
    auto odd_bits  = [](uint32_t x) {
        x = (x >> 1) & 0x55555555u;
        x = (x | (x >> 1)) & 0x33333333u;
        x = (x | (x >> 2)) & 0x0F0F0F0Fu;
        x = (x | (x >> 4)) & 0x00FF00FFu;
        x = (x | (x >> 8)) & 0x0000FFFFu;
        return x;
    };
    bitmap[w] = odd_bits(in.x) | (odd_bits(in.y) << 16);
}
int launch_filter_to_bitmap(cudaStream_t s, const uint32_t* filter, uint64_t nbuckets, uint32_t* bitmap) {
    const uint64_t nwords_out = nbuckets / 32;
    if (!nwords_out) return 0;
    filter_to_bitmap_kernel<<<(unsigned)div_up(nwords_out, 256), 256, 0, s>>>(filter, nwords_out, bitmap);
    return 1;
}

int launch_count_ngrams(cudaStream_t s, const uint32_t* prev, uint32_t* cur, uint64_t npos, NgramSlot* table, uint64_t cap, DeviceStats* st, int sms, const uint32_t* filter,
                        uint64_t nbuckets, bool hot, uint32_t dense, const uint32_t* list, uint64_t nlist, uint32_t* dense_cnt, bool onebit) {
    const uint64_t nitems = list ? nlist : npos;
    if (!nitems) return 0;
    const bool f = filter != nullptr;
    if (list) {  // (the dense square belongs to level 2, which never runs from a list: its input is the class ids themselves)
        if (f) launch_count_variant<true, false, true>(s, prev, list, nitems, cur, table, cap, filter, nbuckets, st, sms, hot, 0, nullptr, onebit);
        else launch_count_variant<false, false, true>(s, prev, list, nitems, cur, table, cap, nullptr, 0, st, sms, hot, 0);
    } else if (dense) {
        if (f) launch_count_variant<true, true, false>(s, prev, nullptr, nitems, cur, table, cap, filter, nbuckets, st, sms, hot, dense, dense_cnt, onebit);
        else launch_count_variant<false, true, false>(s, prev, nullptr, nitems, cur, table, cap, nullptr, 0, st, sms, hot, dense, dense_cnt);
    } else {
        if (f) launch_count_variant<true, false, false>(s, prev, nullptr, nitems, cur, table, cap, filter, nbuckets, st, sms, hot, 0, nullptr, onebit);
        else launch_count_variant<false, false, false>(s, prev, nullptr, nitems, cur, table, cap, nullptr, 0, st, sms, hot, 0);
    }
    return 1;
}
int launch_prune_ngrams(cudaStream_t s, const NgramSlot* table, uint64_t cap, uint32_t threshold, uint32_t* sv_pos, uint32_t* sv_count, uint32_t* bitmap, DeviceStats* st, int sms,
                        uint32_t* slot_index) {
    static int bps = blocks_per_sm((const void*)prune_table_kernel<NgramSlot, false>, 256, 0);
    unsigned   grid = (unsigned)umin64(div_up(cap, kPruneTile), (uint64_t)sms * bps);
    prune_table_kernel<NgramSlot, false><<<grid ? grid : 1, 256, 0, s>>>(table, cap, threshold, sv_pos, sv_count, nullptr, bitmap, slot_index, st, nullptr, 0);
    return 1;
}
int launch_prune_dense(cudaStream_t s, const uint32_t* dense_cnt, uint32_t dense, uint64_t cap, uint32_t threshold, uint32_t* sv_pos, uint32_t* sv_count, uint32_t* bitmap,
                       uint32_t* slot_index, uint32_t* tok_ext, uint32_t ext_pos0, DeviceStats* st, int sms) {
    if (!dense) return 0;
    unsigned grid = (unsigned)umin64(div_up((uint64_t)dense * dense, 256), (uint64_t)sms * 8);
    prune_dense_kernel<<<grid, 256, 0, s>>>(dense_cnt, dense, cap, threshold, sv_pos, sv_count, bitmap, slot_index, tok_ext, ext_pos0, st);
    return 1;
}
int launch_relabel(cudaStream_t s, uint32_t* cur, uint64_t npos, const uint32_t* bitmap, const uint32_t* list_in, uint64_t nlist_in, uint32_t* list_out, unsigned long long* cursor,
                   int sms) {
    if (list_in != nullptr) {
        if (!nlist_in) return 0;
        unsigned grid = (unsigned)umin64(div_up(nlist_in, 256), (uint64_t)sms * 8);
        relabel_list_kernel<<<grid, 256, 0, s>>>(cur, list_in, nlist_in, bitmap, list_out, cursor);
    } else if (list_out != nullptr) {
        relabel_kernel<true><<<div_up(div_up(npos, 4), 256), 256, 0, s>>>(cur, npos, bitmap, list_out, cursor);
    } else {
        relabel_kernel<false><<<div_up(div_up(npos, 4), 256), 256, 0, s>>>(cur, npos, bitmap, nullptr, nullptr);
    }
    return 1;
}

// =============================================================================================
// Skipgrams (exhaustive mode, patternmodel.h:1163-1171 -> computeskipgrams :1370-1527).  A window that is valid for
// the n-gram pass is valid for every gap mask.  The skipgram's identity is the mask plus the ids of its
// contiguous non-gap runs, each of which is a surviving k-gram (k < n) whose id sits in ids[k][p + start].
// find-or-claim the slot of a 128-bit key; returns slot index + 1 (0 when the table is full)
__device__ __forceinline__ uint32_t upsert_skipkey(SkipSlot* __restrict__ table, uint64_t cap, unsigned long long k0, unsigned long long k1, bool count, uint32_t pos) {
    uint64_t slot = fast_range(table_hash_u128(k0, k1), cap);
    for (uint64_t step = 0; step < cap; ++step) {
        SkipSlot*          s  = table + slot;
        ulonglong2         kv = __ldcg(reinterpret_cast<const ulonglong2*>(s));
        unsigned long long c0 = kv.x, c1 = kv.y;
        if (c0 == 0 || c1 == 0) {  // empty (or caught mid-claim): the CAS result is authoritative
            cas128(s, k0, k1, c0, c1);
            if (c0 == 0 && c1 == 0) {
                s->pos = pos;
                c0     = k0;
                c1     = k1;
            }
        }
        if (c0 == k0 && c1 == k1) {
            if (count) atomicAdd(&s->count, 1u);
            return (uint32_t)slot + 1;
        }
        slot = slot + 1 == cap ? 0 : slot + 1;
    }
    return 0;
}

// Two ways to enumerate the windows: every position whose two (n-1)-grams survive (exhaustive mode, occ_pos == NULL), or an
// explicit list of positions (occ_pos[j], indexed models: the occurrences of the surviving n-grams, trainskipgrams
// include/patternmodel.h:2969-3010).  item_slot (optional) receives the slot + 1 of every (window, mask) item.
__global__ void __launch_bounds__(256) count_skipgrams_kernel(const uint32_t* const* __restrict__ ids, int n, const SkipMask* __restrict__ masks, int nmasks, uint64_t npos,
                                                              SkipSlot* __restrict__ table, uint64_t cap, DeviceStats* __restrict__ st, const uint32_t* __restrict__ occ_pos,
                                                              uint32_t* __restrict__ item_slot) {
    __shared__ uint64_t scratch[8];
    const uint32_t* prev  = ids[n - 1];
    const uint64_t  total = npos * (uint64_t)nmasks;
    uint32_t        valid = 0;
    bool            full  = false;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t p = t / nmasks;
        int      m = (int)(t - p * nmasks);
        if (occ_pos != nullptr)
            p = __ldg(occ_pos + p);
        else if (__ldg(prev + p) == 0 || __ldg(prev + p + 1) == 0)
            continue;
        const SkipMask* sm     = masks + m;
        const uint32_t  mask   = __ldg(&sm->mask);
        const uint32_t  nparts = __ldg(&sm->nparts);
        uint32_t part[kMaxSkipParts];
#pragma unroll
        for (int k = 0; k < kMaxSkipParts; ++k)
            part[k] = (uint32_t)k < nparts ? __ldg(ids[__ldg(&sm->len[k])] + p + __ldg(&sm->start[k])) : 0;
        // fold the three leading ids into one until at most three remain (only masks with > 3 runs, i.e. n >= 7)
        uint32_t left = nparts, rounds = 0, first = 0;  // part[first..first+left) are the live ids
        while (left > 3) {
            ++rounds;
            unsigned long long h0 = ((unsigned long long)(mask | kSkipCombiner | (rounds << kSkipRoundShift)) << 32) | part[first];
            unsigned long long h1 = ((unsigned long long)part[first + 1] << 32) | part[first + 2];
            uint32_t           u  = upsert_skipkey(table, cap, h0, h1, false, (uint32_t)p);
            if (u == 0) full = true;
            first += 2;
            part[first] = u;
            left -= 2;
        }
        unsigned long long k0 = ((unsigned long long)(mask | (rounds << kSkipRoundShift)) << 32) | part[first];
        unsigned long long k1 = ((unsigned long long)part[first + 1] << 32) | (left > 2 ? part[first + 2] : 0u);
        ++valid;
        const uint32_t slot = upsert_skipkey(table, cap, k0, k1, true, (uint32_t)p);
        if (slot == 0) full = true;
        if (item_slot != nullptr) item_slot[t] = slot;
    }
    uint64_t v = block_reduce_sum(valid, scratch);
    if (threadIdx.x == 0 && v) atomicAdd(&st->valid_windows, (unsigned long long)v);
    if (full) atomicOr(&st->errflags, kErrTableFull);
}

// Skip-type counting of indexed models (IndexedPatternModel::pruneskipgrams :3362-3383, getskipcontent :3029-3059): the
// "content" of an occurrence is the raw token span from the first to the last gap, i.e. the surviving k-gram starting at
// p + head; a skipgram survives only if at least MINSKIPTYPES distinct contents occur.  Distinct (skipgram slot, content id)
// pairs are claimed in a scratch table; the first claim of a pair bumps types[slot].
__global__ void __launch_bounds__(256) skip_types_kernel(const uint32_t* const* __restrict__ ids, int n, const SkipMask* __restrict__ masks, int nmasks, uint64_t nocc,
                                                         const uint32_t* __restrict__ occ_pos, const uint32_t* __restrict__ item_slot, NgramSlot* __restrict__ pairs, uint64_t cap,
                                                         uint32_t* __restrict__ types, DeviceStats* __restrict__ st) {
    const uint64_t total = nocc * (uint64_t)nmasks;
    bool           full  = false;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t slot = item_slot[t];
        if (slot == 0) continue;
        const uint64_t j    = t / nmasks;
        const uint32_t mask = __ldg(&masks[t - j * nmasks].mask);
        const uint32_t p    = __ldg(occ_pos + j);
        const int      head = __ffs(mask) - 1;                       // leading non-gap tokens
        const int      tail = n - (32 - __clz(mask));                // trailing non-gap tokens
        const uint32_t cid  = __ldg(ids[n - head - tail] + p + head);  // id of the content k-gram
        const unsigned long long key = ((unsigned long long)slot << 32) | cid;
        uint64_t       at    = fast_range(table_hash_u64(key), cap);
        const uint64_t limit = cap < kMaxProbe ? cap : kMaxProbe;
        uint64_t       step  = 0;
        for (; step < limit; ++step) {
            unsigned long long cur = __ldcg(&pairs[at].key);
            if (cur == 0) {
                cur = atomicCAS(&pairs[at].key, 0ull, key);
                if (cur == 0) {
                    atomicAdd(&types[slot - 1], 1u);  // first sighting of this content for this skipgram
                    break;
                }
            }
            if (cur == key) break;
            at = at + 1 == cap ? 0 : at + 1;
        }
        if (step == limit) full = true;
    }
    if (full) atomicOr(&st->errflags, kErrTableFull);
}

int launch_count_skipgrams(cudaStream_t s, const uint32_t* const* ids, int n, const SkipMask* masks, int nmasks, uint64_t npos, SkipSlot* table, uint64_t cap, DeviceStats* st,
                           int sms, const uint32_t* occ_pos, uint32_t* item_slot) {
    static int bps  = blocks_per_sm((const void*)count_skipgrams_kernel, 256, 0);
    unsigned   grid = (unsigned)umin64(div_up(npos * nmasks, 256), (uint64_t)sms * bps * 4);
    count_skipgrams_kernel<<<grid ? grid : 1, 256, 0, s>>>(ids, n, masks, nmasks, npos, table, cap, st, occ_pos, item_slot);
    return 1;
}
int launch_skip_types(cudaStream_t s, const uint32_t* const* ids, int n, const SkipMask* masks, int nmasks, uint64_t nocc, const uint32_t* occ_pos, const uint32_t* item_slot,
                      NgramSlot* pairs, uint64_t cap, uint32_t* types, DeviceStats* st, int sms) {
    if (!nocc) return 0;
    unsigned grid = (unsigned)umin64(div_up(nocc * nmasks, 256), (uint64_t)sms * 16);
    skip_types_kernel<<<grid ? grid : 1, 256, 0, s>>>(ids, n, masks, nmasks, nocc, occ_pos, item_slot, pairs, cap, types, st);
    return 1;
}
int launch_prune_skipgrams(cudaStream_t s, const SkipSlot* table, uint64_t cap, uint32_t threshold, uint32_t* sv_pos, uint32_t* sv_count, uint32_t* sv_mask, DeviceStats* st, int sms,
                           uint32_t* slot_index, const uint32_t* types, uint32_t mintypes) {
    static int bps  = blocks_per_sm((const void*)prune_table_kernel<SkipSlot, true>, 256, 0);
    unsigned   grid = (unsigned)umin64(div_up(cap, kPruneTile), (uint64_t)sms * bps);
    prune_table_kernel<SkipSlot, true><<<grid ? grid : 1, 256, 0, s>>>(table, cap, threshold, sv_pos, sv_count, sv_mask, nullptr, slot_index, st, types, mintypes);
    return 1;
}

// =============================================================================================
// K5: export.  A survivor is (position | class, count, n | mask << 8); its key bytes are re-encoded from the token
// array in the reference's pattern format: varint per token, gap tokens collapsed to the single byte 0x03
// (Pattern(const PatternPointer&), src/pattern.cpp:873-909).
// A few words device -> mapped pinned host memory by plain stores over PCIe.  A cudaMemcpyAsync of the same bytes would queue on the D2H copy
// engine behind whatever the export stream is copying (a finished level: tens of MB) and stall the stream that is waiting for a statistics block.
__global__ void copy_words_to_host_kernel(const uint32_t* __restrict__ src, volatile uint32_t* __restrict__ dst, uint32_t nwords) {
    for (uint32_t i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}
int launch_copy_words_to_host(cudaStream_t s, const void* src, void* dst_mapped, uint32_t nbytes) {
    copy_words_to_host_kernel<<<1, 32, 0, s>>>(static_cast<const uint32_t*>(src), static_cast<volatile uint32_t*>(dst_mapped), (nbytes + 3) / 4);
    return 1;
}

__global__ void __launch_bounds__(256) fill_u32_kernel(uint32_t* __restrict__ dst, uint64_t n, uint32_t value) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = value;
}

__global__ void __launch_bounds__(256) pack_nm_kernel(uint32_t* __restrict__ nm, uint64_t count, uint32_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) nm[i] = n | (nm[i] << 8);
}
int launch_pack_nm(cudaStream_t s, uint32_t* nm, uint64_t count, uint32_t n) {
    if (!count) return 0;
    pack_nm_kernel<<<div_up(count, 256), 256, 0, s>>>(nm, count, n);
    return 1;
}

__device__ __forceinline__ uint32_t pattern_bytes(const uint32_t* __restrict__ tok, uint32_t pos, uint32_t nm, uint8_t* out) {
    uint32_t n = nm & 0xFFu, mask = nm >> 8, len = 0;
    if (n == 1) return out ? varint_put(out, pos) : varint_len(pos);
    // the first eight tokens are fetched before any of them is used: one memory latency per pattern instead of one per token
    uint32_t head[8];
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j) head[j] = j < n ? __ldg(tok + pos + j) : 0u;
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j) {
        if (j < n) {
            if ((mask >> j) & 1u) {
                if (out) out[len] = 3;
                len += 1;
            } else {
                len += out ? varint_put(out + len, head[j]) : varint_len(head[j]);
            }
        }
    }
    for (uint32_t j = 8; j < n; ++j) {
        bool gap = j < 24 && ((mask >> j) & 1u);
        if (gap) {
            if (out) out[len] = 3;
            len += 1;
        } else {
            uint32_t c = __ldg(tok + pos + j);
            len += out ? varint_put(out + len, c) : varint_len(c);
        }
    }
    return len;
}

__global__ void __launch_bounds__(256) export_lengths_kernel(const uint32_t* __restrict__ tok, const uint32_t* __restrict__ sv_pos, const uint32_t* __restrict__ sv_nm, uint64_t n,
                                                             uint32_t* __restrict__ lens, uint16_t* __restrict__ lens16) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        uint32_t l = pattern_bytes(tok, sv_pos[i], sv_nm[i], nullptr);
        lens[i]    = l;
        lens16[i]  = (uint16_t)l;  // <= 255 tokens x 5 bytes
    }
}
// the 256 patterns of a block are neighbours in the key blob: their bytes are put together in shared memory and leave as aligned words
// (a block whose patterns are too long for the stage writes them directly)
constexpr uint32_t kKeyStage = 12288;
__global__ void __launch_bounds__(256) export_write_kernel(const uint32_t* __restrict__ tok, const uint32_t* __restrict__ sv_pos, const uint32_t* __restrict__ sv_nm,
                                                           const uint64_t* __restrict__ off, uint64_t n, uint8_t* __restrict__ keys) {
    __shared__ __align__(16) uint8_t stage[kKeyStage];
    const uint64_t first = (uint64_t)blockIdx.x * blockDim.x, i = first + threadIdx.x;
    const uint64_t last  = first + blockDim.x < n ? first + blockDim.x : n;
    const uint64_t base = off[first], span = off[last] - base;
    if (span > kKeyStage) {
        if (i < n) pattern_bytes(tok, sv_pos[i], sv_nm[i], keys + off[i]);
        return;
    }
    if (i < n) pattern_bytes(tok, sv_pos[i], sv_nm[i], stage + (off[i] - base));
    __syncthreads();
    uint8_t*       dst  = keys + base;
    const uint32_t lead = min((uint32_t)span, (uint32_t)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3));
    const uint32_t nw   = ((uint32_t)span - lead) / 4, tail = lead + nw * 4;
    if (threadIdx.x < lead) dst[threadIdx.x] = stage[threadIdx.x];
    for (uint32_t w = threadIdx.x; w < nw; w += blockDim.x) {
        const uint8_t* q = stage + lead + 4 * w;
        reinterpret_cast<uint32_t*>(dst + lead)[w] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
    }
    if (tail + threadIdx.x < span) dst[tail + threadIdx.x] = stage[tail + threadIdx.x];
}
// exclusive scan u32 -> u64 over n items, out has n+1 entries (out[n] = total).  Three small kernels, 2048 items per block.
constexpr int kScanItems = 2048;
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(const uint32_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ sums) {
    __shared__ uint64_t scratch[32];
    uint64_t base = (uint64_t)blockIdx.x * kScanItems;
    uint64_t v    = 0;
    for (int k = 0; k < 2; ++k) {
        uint64_t i = base + threadIdx.x + (uint64_t)k * 1024;
        if (i < n) v += in[i];
    }
    v = block_reduce_sum(v, scratch);
    if (threadIdx.x == 0) sums[blockIdx.x] = v;
}
__global__ void __launch_bounds__(1024) scan_sums_kernel(uint64_t* __restrict__ sums, uint64_t nb) {
    __shared__ uint64_t part[1024];
    uint64_t chunk = (nb + blockDim.x - 1) / blockDim.x;
    uint64_t lo = threadIdx.x * chunk, hi = min(nb, lo + chunk);
    uint64_t s = 0;
    for (uint64_t i = lo; i < hi; ++i) s += sums[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t acc = 0;
        for (uint32_t i = 0; i < blockDim.x; ++i) {
            uint64_t t = part[i];
            part[i]    = acc;
            acc += t;
        }
        sums[nb] = acc;
    }
    __syncthreads();
    uint64_t acc = part[threadIdx.x];
    for (uint64_t i = lo; i < hi; ++i) {
        uint64_t t = sums[i];
        sums[i]    = acc;
        acc += t;
    }
}
__global__ void __launch_bounds__(1024) scan_apply_kernel(const uint32_t* __restrict__ in, uint64_t n, const uint64_t* __restrict__ sums, uint64_t nb, uint64_t* __restrict__ out) {
    __shared__ uint32_t warp_tot[32];
    uint64_t base = (uint64_t)blockIdx.x * kScanItems;
    uint64_t i0 = base + (uint64_t)threadIdx.x * 2, i1 = i0 + 1;  // two consecutive items per thread
    uint32_t a = i0 < n ? in[i0] : 0, b = i1 < n ? in[i1] : 0;
    uint32_t c = a + b;  // a block covers 2048 items of < 2^20 each in practice; per-block sums stay far below 2^32
    uint32_t incl = warp_inclusive_scan(c);
    if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t t = warp_tot[threadIdx.x];
        uint32_t s = warp_inclusive_scan(t);
        warp_tot[threadIdx.x] = s - t;
    }
    __syncthreads();
    uint64_t excl = sums[blockIdx.x] + warp_tot[threadIdx.x >> 5] + (incl - c);
    if (i0 < n) out[i0] = excl;
    if (i1 < n) out[i1] = excl + a;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = sums[nb];
}

__global__ void __launch_bounds__(256) sum_u32_kernel(const uint32_t* __restrict__ v, uint64_t n, unsigned long long* __restrict__ total) {
    __shared__ uint64_t scratch[8];
    uint64_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) acc += v[i];
    acc = block_reduce_sum(acc, scratch);
    if (threadIdx.x == 0 && acc) atomicAdd(total, (unsigned long long)acc);
}
int launch_sum_u32(cudaStream_t s, const uint32_t* v, uint64_t n, unsigned long long* total) {
    if (!n) return 0;
    sum_u32_kernel<<<(unsigned)umin64(div_up(n, 256), 148), 256, 0, s>>>(v, n, total);
    return 1;
}
int launch_fill_u32(cudaStream_t s, uint32_t* dst, uint64_t n, uint32_t value) {
    if (!n) return 0;
    fill_u32_kernel<<<div_up(n, 256), 256, 0, s>>>(dst, n, value);
    return 1;
}
int launch_export_lengths(cudaStream_t s, const uint32_t* tok, const uint32_t* sv_pos, const uint32_t* sv_nm, uint64_t n, uint32_t* lens, uint16_t* lens16) {
    if (!n) return 0;
    export_lengths_kernel<<<div_up(n, 256), 256, 0, s>>>(tok, sv_pos, sv_nm, n, lens, lens16);
    return 1;
}
int launch_exclusive_scan_u32_u64(cudaStream_t s, const uint32_t* in, uint64_t* out, uint64_t n, uint64_t* tmp) {
    uint64_t nb = n ? (n + kScanItems - 1) / kScanItems : 1;
    scan_block_sums_kernel<<<(unsigned)nb, 1024, 0, s>>>(in, n, tmp);
    scan_sums_kernel<<<1, 1024, 0, s>>>(tmp, nb);
    scan_apply_kernel<<<(unsigned)nb, 1024, 0, s>>>(in, n, tmp, nb, out);
    return 3;
}
int launch_export_write(cudaStream_t s, const uint32_t* tok, const uint32_t* sv_pos, const uint32_t* sv_nm, const uint64_t* off, uint64_t n, uint8_t* keys) {
    if (!n) return 0;
    export_write_kernel<<<div_up(n, 256), 256, 0, s>>>(tok, sv_pos, sv_nm, off, n, keys);
    return 1;
}
// =============================================================================================
// Pattern::hash on the device for arbitrary pattern bytes (parity row a5)
__global__ void __launch_bounds__(256) hash64_batch_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ off, uint64_t n, uint64_t* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t b = off[i], e = off[i + 1];
    // src/pattern.cpp:234-236: an empty pattern hashes to 0
    out[i] = (e == b || keys[b] == 0) ? 0 : spooky_hash64(keys + b, (uint32_t)(e - b), 0);
}
int launch_hash64_batch(cudaStream_t s, const uint8_t* keys, const uint64_t* off, uint64_t n, uint64_t* out) {
    if (!n) return 0;
    hash64_batch_kernel<<<div_up(n, 256), 256, 0, s>>>(keys, off, n, out);
    return 1;
}

// =============================================================================================
// Synthetic corpus (measurement input).  Must stay bit-identical to oracle/oracle.c: oracle_synth_*.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t rnd(uint64_t seed, uint64_t stream, uint64_t i) {
    return mix64((seed + stream * 0xD1B54A32D192ED03ULL) ^ mix64(i));
}
struct SynthArgs {
    uint64_t        seed, ntokens, first;
    uint32_t        vocab, mean_sentence, phrase_permille, nphrases;
    const uint64_t* cdf;
};
__device__ __forceinline__ uint32_t zipf_rank(const SynthArgs& a, uint64_t u) {
    uint64_t x  = u % __ldg(a.cdf + a.vocab - 1);
    uint32_t lo = 0, hi = a.vocab - 1;
    while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (__ldg(a.cdf + mid) > x)
            hi = mid;
        else
            lo = mid + 1;
    }
    return lo;
}
__device__ __forceinline__ bool phrase_info(const SynthArgs& a, uint64_t i, uint64_t& id, uint32_t& j, uint32_t& L) {
    if (!a.phrase_permille || !a.nphrases) return false;
    uint64_t b = i >> 3;
    if (rnd(a.seed, 2, b) % 1000 >= a.phrase_permille) return false;
    id         = rnd(a.seed, 4, b) % a.nphrases;
    L          = 3 + (uint32_t)(rnd(a.seed, 3, id) % 4);
    uint32_t k = (uint32_t)(i & 7);
    if (k >= L) return false;
    j = k;
    return true;
}
__device__ __forceinline__ void synth_token(const SynthArgs& a, uint64_t local, uint32_t& cls, bool& brk) {
    const uint64_t i = a.first + local;  // index in the global stream
    uint64_t id;
    uint32_t j, L;
    bool     inphrase = phrase_info(a, i, id, j, L);
    cls               = 6 + zipf_rank(a, inphrase ? rnd(a.seed, 5, id * 8 + j) : rnd(a.seed, 0, i));
    brk               = (inphrase && j + 1 < L) ? false : (rnd(a.seed, 1, i) % a.mean_sentence == 0);
    if (local + 1 == a.ntokens) brk = true;  // a corpus (or shard) always ends with a delimiter
}
__global__ void __launch_bounds__(256) synth_lengths_kernel(SynthArgs a, uint32_t* __restrict__ lens) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.ntokens) return;
    uint32_t cls;
    bool     brk;
    synth_token(a, i, cls, brk);
    lens[i] = varint_len(cls) + (brk ? 1u : 0u);
}
__global__ void __launch_bounds__(256) synth_write_kernel(SynthArgs a, const uint64_t* __restrict__ off, uint8_t* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.ntokens) return;
    uint32_t cls;
    bool     brk;
    synth_token(a, i, cls, brk);
    uint8_t* o = out + off[i];
    uint32_t l = varint_put(o, cls);
    if (brk) o[l] = 0;
}
int launch_synth_lengths(cudaStream_t s, uint64_t seed, uint64_t ntokens, uint64_t first, uint32_t vocab, uint32_t mean_sentence, uint32_t phrase_permille, uint32_t nphrases,
                         const uint64_t* cdf, uint32_t* lens) {
    SynthArgs a{seed, ntokens, first, vocab, mean_sentence, phrase_permille, nphrases, cdf};
    synth_lengths_kernel<<<div_up(ntokens, 256), 256, 0, s>>>(a, lens);
    return 1;
}
int launch_synth_write(cudaStream_t s, uint64_t seed, uint64_t ntokens, uint64_t first, uint32_t vocab, uint32_t mean_sentence, uint32_t phrase_permille, uint32_t nphrases,
                       const uint64_t* cdf, const uint64_t* off, uint8_t* out) {
    SynthArgs a{seed, ntokens, first, vocab, mean_sentence, phrase_permille, nphrases, cdf};
    synth_write_kernel<<<div_up(ntokens, 256), 256, 0, s>>>(a, off, out);
    return 1;
}

}  // namespace colibri
