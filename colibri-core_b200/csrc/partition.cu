// partition.cu -- the partitioned counting path of a level n >= 2 (large levels; the hashed HBM table of kernels.cu stays for small ones).
//
// Replaces, per window, the two has() lookups + add() of the reference (include/patternmodel.h:1139-1161) and, per level, prune()
// (:2107-2128).  Round 1 counted every window with one random 32-byte-sector access into a table in HBM; ncu showed the level-2 launch
// moving 130 B of DRAM per table window at 38 % of the HBM peak (profiles/r02_ncu.md) -- random sectors are what HBM is worst at.  Here
// the windows of a level are radix-partitioned by the high bits of their key hash, twice (b1 + b2 bits), with streaming passes only, until
// a partition holds a few hundred windows; one thread block then counts a partition in a SHARED-MEMORY table, applies the threshold there
// and hands every window of a surviving n-gram its id.  Nothing in HBM is ever probed:
//   A  part_hist      stream prev[] (or the position list): histogram of the b1+b2 partition bits (REDs on an L2-resident array);
//                     level 2: pairs of frequent classes are counted in the dense square instead and never become records
//   -  prune_dense    (kernels.cu) dense square -> survivors, dense_id[cell] = survivor index + 1
//   -  part_scan      exclusive scan of the histogram: partition offsets, cursors, tile starts of pass D
//   B  part_split1    stream prev[] again: record (key, position) of every hashed window -> its b1-partition (block-staged in shared
//                     memory, one cursor reservation per bin per tile); writes cur[] = dense id or 0 for every position on the way
//   D  part_split2    every b1-partition -> its 2^b2 sub-partitions, same scheme
//   E  part_count     one block per final partition: shared-memory open addressing (64-bit CAS), threshold, survivor compaction
//                     (one global cursor reservation per partition), then cur[position] = survivor index + 1 for the windows that stay
// ids are survivor indices + 1 (dense, no table slots), a pruned window keeps id 0: no survivor bitmap, no relabel pass, no table memset,
// no table scan.  Exactness: keys are compared in full, partitions are exhaustive and disjoint (a key has one hash), counts are exact.
// A partition with more distinct keys than the shared-memory table holds raises kErrTableFull and the host reruns the level on the HBM-table
// path (never seen on hashed keys; partitions are sized for <= 512 windows on average, the table holds 2048 keys).
#include <algorithm>

#include "device_utils.cuh"
#include "kernels.h"

namespace colibri {

namespace {

constexpr int      kTile1      = 4096;  // items per block of pass B (16 per thread, four 16-byte loads)
constexpr int      kTile2      = 2048;  // records per block of pass D
constexpr uint32_t kPartSlots  = 2048;  // shared-memory table of pass E
constexpr uint32_t kHotSide    = 64;    // pass A counts the pairs of the 64 most frequent classes in shared memory first

__device__ __forceinline__ uint32_t part_of(uint64_t h, int shift) {
    return (uint32_t)(h >> shift);
}

// ---- pass A ---------------------------------------------------------------------------------------------------------------------
template <bool kList, bool kDense>
__global__ void __launch_bounds__(256) part_hist_kernel(const uint32_t* __restrict__ prev, const uint32_t* __restrict__ list, uint64_t nitems, uint32_t dense,
                                                        uint32_t* __restrict__ dense_cnt, uint32_t* __restrict__ hist, int shift, DeviceStats* __restrict__ st) {
    __shared__ uint64_t scratch[8];
    __shared__ uint32_t hot[kDense ? kHotSide * kHotSide : 1];
    if (kDense) {
        for (uint32_t i = threadIdx.x; i < kHotSide * kHotSide; i += blockDim.x) hot[i] = 0;
        __syncthreads();
    }
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t       valid  = 0;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nitems; j += stride) {
        uint64_t p = j;
        uint32_t a;
        if (kList) {
            p = __ldcs(list + j);
            a = __ldg(prev + p);
        } else {
            a = __ldcs(prev + p);
        }
        const uint32_t b = __ldg(prev + p + 1);
        if (a == 0 || b == 0) continue;
        ++valid;
        if (kDense && a < dense && b < dense) {
            if (a < kHotSide && b < kHotSide) atomicAdd(&hot[a * kHotSide + b], 1u);
            else atomicAdd(dense_cnt + a * dense + b, 1u);
            continue;
        }
        atomicAdd(hist + part_of(table_hash_u64(((unsigned long long)a << 32) | b), shift), 1u);
    }
    if (kDense) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < kHotSide * kHotSide; i += blockDim.x) {
            const uint32_t c = hot[i], a = i / kHotSide, b = i % kHotSide;
            if (c && a < dense && b < dense) atomicAdd(dense_cnt + a * dense + b, c);
        }
    }
    uint64_t v = block_reduce_sum(valid, scratch);
    if (threadIdx.x == 0 && v) atomicAdd(&st->valid_windows, (unsigned long long)v);
}

// ---- scan: off[i] = sum of hist[0..i), cursor2 = off, cursor1[p1] = off[p1 << b2], tstart[p1] = tiles of pass D before partition p1 -----------
__global__ void __launch_bounds__(1024) part_scan_kernel(const uint32_t* __restrict__ hist, uint32_t nparts, int b2, uint32_t* off, uint32_t* __restrict__ cursor2,
                                                         uint32_t* __restrict__ cursor1, uint32_t* __restrict__ tstart) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < nparts; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nparts ? hist[i] : 0u;
        uint32_t incl = warp_inclusive_scan(v);
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_tot[lane];
            uint32_t s = warp_inclusive_scan(w);
            warp_tot[lane] = s - w;
        }
        __syncthreads();
        const uint32_t excl = carry + warp_tot[warp] + incl - v;
        if (i < nparts) {
            off[i]     = excl;
            cursor2[i] = excl;
            if ((i & ((1u << b2) - 1)) == 0) cursor1[i >> b2] = excl;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[nparts] = carry;
    __syncthreads();
    // tiles of pass D per b1-partition: tstart[q] = tiles before partition q (<= 2048 partitions: two rounds of the same block scan)
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const uint32_t p1n = nparts >> b2, total = off[nparts];
    for (uint32_t base = 0; base <= p1n; base += 1024) {
        const uint32_t q = base + threadIdx.x;
        uint32_t       v = 0;
        if (q < p1n) {
            const uint32_t beg = off[q << b2], end = (q + 1 == p1n) ? total : off[(q + 1) << b2];
            v                  = (end - beg + kTile2 - 1) / kTile2;
        }
        uint32_t incl = warp_inclusive_scan(v);
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_tot[lane];
            uint32_t s = warp_inclusive_scan(w);
            warp_tot[lane] = s - w;
        }
        __syncthreads();
        const uint32_t excl = carry + warp_tot[warp] + incl - v;
        if (q <= p1n) tstart[q] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}

// ---- the block-level multi-split both passes share: nbins counters, exclusive prefix, one cursor reservation per non-empty bin -------------
// bins[0..nbins) = per-bin record count of the tile on entry; on exit bins[b] = first staging index of bin b, gdelta[b] = (global index of
// the bin's run) - (staging index), returns the tile's record total
__device__ __forceinline__ uint32_t split_reserve(uint32_t* bins, uint32_t* gdelta, uint32_t nbins, uint32_t* __restrict__ cursor, uint32_t* warp_tot) {
    // nbins <= 2048, 256 threads: thread t owns bins [t * per, (t + 1) * per)
    const uint32_t per = (nbins + 255) / 256;
    uint32_t       local[8];
    uint32_t       sum = 0;
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) {
        const uint32_t b = threadIdx.x * per + k;
        local[k]         = (k < per && b < nbins) ? bins[b] : 0u;
        sum += local[k];
    }
    uint32_t incl = warp_inclusive_scan(sum);
    if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (uint32_t w = 0; w < 8; ++w) {
        const uint32_t c = warp_tot[w];
        if (w < (threadIdx.x >> 5)) before += c;
        total += c;
    }
    uint32_t run = before + incl - sum;
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) {
        const uint32_t b = threadIdx.x * per + k;
        if (k < per && b < nbins) {
            bins[b] = run;
            if (local[k]) gdelta[b] = atomicAdd(cursor + b, local[k]) - run;
            run += local[k];
        }
    }
    __syncthreads();
    return total;
}

// ---- pass B ---------------------------------------------------------------------------------------------------------------------
// dynamic shared memory: stage_key u64[kTile1] | stage_pos u32[kTile1] | stage_bin u16[kTile1] | bins u32[nbins] | gdelta u32[nbins]
template <bool kList, bool kDense>
__global__ void __launch_bounds__(256) part_split1_kernel(const uint32_t* __restrict__ prev, const uint32_t* __restrict__ list, uint64_t nitems, uint32_t dense,
                                                          const uint32_t* __restrict__ dense_id, uint32_t* __restrict__ cur, int shift1, uint32_t nbins,
                                                          uint32_t* __restrict__ cursor1, unsigned long long* __restrict__ rk, uint32_t* __restrict__ rp) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long* stage_key = reinterpret_cast<unsigned long long*>(smem);
    uint32_t*           stage_pos = reinterpret_cast<uint32_t*>(stage_key + kTile1);
    uint16_t*           stage_bin = reinterpret_cast<uint16_t*>(stage_pos + kTile1);
    uint32_t*           bins      = reinterpret_cast<uint32_t*>(stage_bin + kTile1);
    uint32_t*           gdelta    = bins + nbins;
    __shared__ uint32_t warp_tot[8];

    for (uint32_t i = threadIdx.x; i < nbins; i += 256) bins[i] = 0;
    __syncthreads();

    const uint64_t tile = (uint64_t)blockIdx.x * kTile1;
    uint32_t       a[16], b[16], pos[16];
    uint32_t       binrank[16];  // bin << 16 | rank inside (tile, bin); 0xFFFFFFFF = no record
    if (!kList) {
        // four 16-byte loads per thread: positions tile + k * 1024 + 4 * tid .. + 3; the right neighbour of the fourth comes from the next lane
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint64_t p0 = tile + (uint64_t)k * 1024 + 4u * threadIdx.x;
            uint4          v  = make_uint4(0, 0, 0, 0);
            if (p0 < nitems) v = __ldcs(reinterpret_cast<const uint4*>(prev + p0));  // prev has nitems + 8 entries
            uint32_t nxt = __shfl_down_sync(0xffffffffu, v.x, 1);
            if (lane_id() == 31) nxt = p0 + 4 <= nitems ? __ldg(prev + p0 + 4) : 0u;
            a[4 * k] = v.x; a[4 * k + 1] = v.y; a[4 * k + 2] = v.z; a[4 * k + 3] = v.w;
            b[4 * k] = v.y; b[4 * k + 1] = v.z; b[4 * k + 2] = v.w; b[4 * k + 3] = nxt;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                pos[4 * k + e] = (uint32_t)(p0 + e);
                if (p0 + e >= nitems) a[4 * k + e] = 0;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint64_t j = tile + (uint64_t)k * 256 + threadIdx.x;
            a[k] = 0; b[k] = 0; pos[k] = 0;
            if (j < nitems) {
                pos[k] = __ldcs(list + j);
                a[k]   = __ldg(prev + pos[k]);
                b[k]   = __ldg(prev + pos[k] + 1);
            }
        }
    }
    uint32_t dense_out[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        binrank[k]   = 0xFFFFFFFFu;
        dense_out[k] = 0;
        if (a[k] == 0 || b[k] == 0) continue;
        if (kDense && a[k] < dense && b[k] < dense) {
            dense_out[k] = __ldg(dense_id + a[k] * dense + b[k]);
            continue;
        }
        const uint32_t bin = part_of(table_hash_u64(((unsigned long long)a[k] << 32) | b[k]), shift1);
        binrank[k]         = (bin << 16) | atomicAdd(&bins[bin], 1u);  // <= 4096 records per tile: the rank fits 16 bits (4095 at most)
    }
    if (!kList) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint64_t p0 = tile + (uint64_t)k * 1024 + 4u * threadIdx.x;
            if (p0 < nitems) __stcs(reinterpret_cast<uint4*>(cur + p0), make_uint4(dense_out[4 * k], dense_out[4 * k + 1], dense_out[4 * k + 2], dense_out[4 * k + 3]));
        }
    }
    __syncthreads();
    const uint32_t total = split_reserve(bins, gdelta, nbins, cursor1, warp_tot);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (binrank[k] == 0xFFFFFFFFu) continue;
        const uint32_t bin = binrank[k] >> 16, idx = bins[bin] + (binrank[k] & 0xFFFFu);
        stage_key[idx]     = ((unsigned long long)a[k] << 32) | b[k];
        stage_pos[idx]     = pos[k];
        stage_bin[idx]     = (uint16_t)bin;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < total; i += 256) {
        const uint32_t dst = gdelta[stage_bin[i]] + i;
        rk[dst]            = stage_key[i];
        rp[dst]            = stage_pos[i];
    }
}

// ---- pass D ---------------------------------------------------------------------------------------------------------------------
// dynamic shared memory: stage_key u64[kTile2] | stage_pos u32[kTile2] | stage_bin u16[kTile2] | bins u32[nbins] | gdelta u32[nbins]
__global__ void __launch_bounds__(256) part_split2_kernel(const unsigned long long* __restrict__ rk_in, const uint32_t* __restrict__ rp_in, const uint32_t* __restrict__ off,
                                                          const uint32_t* __restrict__ tstart, uint32_t p1n, int b2, int shift2, uint32_t* __restrict__ cursor2,
                                                          unsigned long long* __restrict__ rk_out, uint32_t* __restrict__ rp_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t      nbins     = 1u << b2;
    unsigned long long* stage_key = reinterpret_cast<unsigned long long*>(smem);
    uint32_t*           stage_pos = reinterpret_cast<uint32_t*>(stage_key + kTile2);
    uint16_t*           stage_bin = reinterpret_cast<uint16_t*>(stage_pos + kTile2);
    uint32_t*           bins      = reinterpret_cast<uint32_t*>(stage_bin + kTile2);
    uint32_t*           gdelta    = bins + nbins;
    __shared__ uint32_t warp_tot[8];
    if (blockIdx.x >= __ldg(tstart + p1n)) return;
    // which b1-partition this tile belongs to: last q with tstart[q] <= blockIdx.x
    uint32_t lo = 0, hi = p1n;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(tstart + mid) <= blockIdx.x) lo = mid;
        else hi = mid;
    }
    const uint32_t q    = lo;
    const uint32_t beg  = __ldg(off + ((uint64_t)q << b2)) + (blockIdx.x - __ldg(tstart + q)) * kTile2;
    const uint32_t pend = __ldg(off + ((uint64_t)(q + 1) << b2));  // off has nparts + 1 entries: the last partition ends at the total
    const uint32_t end  = beg + kTile2 < pend ? beg + kTile2 : pend;

    for (uint32_t i = threadIdx.x; i < nbins; i += 256) bins[i] = 0;
    __syncthreads();
    unsigned long long key[8];
    uint32_t           pos[8], binrank[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t i = beg + k * 256 + threadIdx.x;
        binrank[k]       = 0xFFFFFFFFu;
        if (i < end) {
            key[k] = __ldcs(rk_in + i);
            pos[k] = __ldcs(rp_in + i);
            const uint32_t bin = part_of(table_hash_u64(key[k]), shift2) & (nbins - 1);
            binrank[k]         = (bin << 16) | atomicAdd(&bins[bin], 1u);
        }
    }
    __syncthreads();
    const uint32_t total = split_reserve(bins, gdelta, nbins, cursor2 + ((uint64_t)q << b2), warp_tot);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (binrank[k] == 0xFFFFFFFFu) continue;
        const uint32_t bin = binrank[k] >> 16, idx = bins[bin] + (binrank[k] & 0xFFFFu);
        stage_key[idx]     = key[k];
        stage_pos[idx]     = pos[k];
        stage_bin[idx]     = (uint16_t)bin;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < total; i += 256) {
        const uint32_t dst = gdelta[stage_bin[i]] + i;
        rk_out[dst]        = stage_key[i];
        rp_out[dst]        = stage_pos[i];
    }
}

// ---- pass E ---------------------------------------------------------------------------------------------------------------------
// One block per partition (persistent blocks take partitions from a work counter).  out[rp[i]] = (survivor index) * id_mul + id_add for the
// records of surviving keys (single GPU: id_mul = id_add = 1, out = cur; owner of a sharded run: out = the reply array, global ids).
constexpr int kHeld = 4;  // records per thread kept in registers between the two phases (partitions of <= 1024 records never re-read)

__device__ __forceinline__ uint32_t smem_upsert(unsigned long long* tk, uint32_t* tc, uint32_t* tp, uint32_t mask, unsigned long long key, uint32_t pos, bool& full) {
    uint32_t slot = (uint32_t)table_hash_u64(key) & mask;
    for (uint32_t step = 0; step <= mask; ++step) {
        const unsigned long long old = atomicCAS(tk + slot, 0ull, key);
        if (old == 0ull) tp[slot] = pos;  // the claimer's position names the n-gram (any occurrence does)
        if (old == 0ull || old == key) {
            atomicAdd(tc + slot, 1u);
            return slot;
        }
        slot = (slot + 1) & mask;
    }
    full = true;
    return 0xFFFFFFFFu;
}
__device__ __forceinline__ uint32_t smem_find(const unsigned long long* tk, uint32_t mask, unsigned long long key) {
    uint32_t slot = (uint32_t)table_hash_u64(key) & mask;
    for (uint32_t step = 0; step <= mask; ++step) {
        if (tk[slot] == key) return slot;
        slot = (slot + 1) & mask;
    }
    return 0xFFFFFFFFu;
}

__global__ void __launch_bounds__(256) part_count_kernel(const unsigned long long* __restrict__ rk, const uint32_t* __restrict__ rp, const uint32_t* __restrict__ off,
                                                         uint32_t nparts, uint32_t threshold, uint32_t* __restrict__ out, uint32_t id_mul, uint32_t id_add,
                                                         uint32_t* __restrict__ sv_pos, uint32_t* __restrict__ sv_cnt, DeviceStats* __restrict__ st,
                                                         unsigned int* __restrict__ work) {
    __shared__ unsigned long long tk[kPartSlots];
    __shared__ uint32_t           tc[kPartSlots];
    __shared__ uint32_t           tp[kPartSlots];
    __shared__ uint64_t           scratch[8];
    __shared__ uint32_t           warp_tot[8];
    __shared__ uint32_t           s_part;
    __shared__ unsigned long long s_base;
    uint64_t found = 0, kept = 0, occ = 0, singles = 0;
    bool     full = false;
    for (;;) {
        __syncthreads();  // the previous partition's tables are done with
        if (threadIdx.x == 0) s_part = atomicAdd(work, 1u);
        __syncthreads();
        const uint32_t q = s_part;
        if (q >= nparts) break;
        const uint32_t beg = __ldg(off + q), end = __ldg(off + q + 1), n = end - beg;
        if (n == 0) continue;
        uint32_t size = 64;
        while (size < 2 * n && size < kPartSlots) size <<= 1;
        const uint32_t mask = size - 1;
        for (uint32_t i = threadIdx.x; i < size; i += 256) {
            tk[i] = 0ull;
            tc[i] = 0u;
        }
        __syncthreads();
        // phase 1: count
        uint32_t held_slot[kHeld], held_pos[kHeld];
#pragma unroll
        for (int k = 0; k < kHeld; ++k) {
            const uint32_t i = beg + k * 256 + threadIdx.x;
            held_slot[k]     = 0xFFFFFFFFu;
            held_pos[k]      = 0;
            if (i < end) {
                const unsigned long long key = __ldcs(rk + i);
                held_pos[k]                  = __ldcs(rp + i);
                held_slot[k]                 = smem_upsert(tk, tc, tp, mask, key, held_pos[k], full);
            }
        }
        for (uint32_t i = beg + kHeld * 256 + threadIdx.x; i < end; i += 256) smem_upsert(tk, tc, tp, mask, __ldg(rk + i), __ldg(rp + i), full);
        __syncthreads();
        // threshold + compaction: thread t owns slots t, t + 256, ...
        uint32_t nkeep = 0;
        for (uint32_t sl = threadIdx.x; sl < size; sl += 256) {
            if (tk[sl] != 0ull) {
                const uint32_t c = tc[sl];
                ++found;
                singles += c == 1;
                if (c >= threshold) {
                    ++nkeep;
                    ++kept;
                    occ += c;
                }
            }
        }
        uint32_t incl = warp_inclusive_scan(nkeep);
        if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < 8; ++w) {
                const uint32_t c = warp_tot[w];
                warp_tot[w]      = tot;
                tot += c;
            }
            s_base = tot ? atomicAdd(&st->cursor, (unsigned long long)tot) : 0ull;
        }
        __syncthreads();
        uint64_t o = s_base + warp_tot[threadIdx.x >> 5] + incl - nkeep;
        for (uint32_t sl = threadIdx.x; sl < size; sl += 256) {
            uint32_t id = 0;
            if (tk[sl] != 0ull) {
                const uint32_t c = tc[sl];
                if (c >= threshold) {
                    sv_pos[o] = tp[sl];
                    sv_cnt[o] = c;
                    id        = (uint32_t)o * id_mul + id_add;
                    ++o;
                }
            }
            tc[sl] = id;  // the slot now answers "which id", 0 = pruned
        }
        __syncthreads();
        // phase 2: ids to the windows that stay
#pragma unroll
        for (int k = 0; k < kHeld; ++k) {
            if (held_slot[k] == 0xFFFFFFFFu) continue;
            const uint32_t id = tc[held_slot[k]];
            if (id) out[held_pos[k]] = id;
        }
        for (uint32_t i = beg + kHeld * 256 + threadIdx.x; i < end; i += 256) {
            const uint32_t sl = smem_find(tk, mask, __ldg(rk + i));
            if (sl != 0xFFFFFFFFu) {
                const uint32_t id = tc[sl];
                if (id) out[__ldg(rp + i)] = id;
            }
        }
    }
    found   = block_reduce_sum(found, scratch);
    kept    = block_reduce_sum(kept, scratch);
    occ     = block_reduce_sum(occ, scratch);
    singles = block_reduce_sum(singles, scratch);
    if (threadIdx.x == 0) {
        if (found) atomicAdd(&st->found, (unsigned long long)found);
        if (kept) atomicAdd(&st->kept, (unsigned long long)kept);
        if (occ) atomicAdd(&st->kept_occ, (unsigned long long)occ);
        if (singles) atomicAdd(&st->singletons, (unsigned long long)singles);
    }
    if (full) atomicOr(&st->errflags, kErrTableFull);
}

// ---- the positions whose id is non-zero, in corpus order inside 1024-position tiles (the next level's list mode) -------------------------
__global__ void __launch_bounds__(256) compact_nonzero_kernel(const uint32_t* __restrict__ cur, uint64_t npos, uint32_t* __restrict__ list_out, unsigned long long* __restrict__ cursor) {
    __shared__ uint32_t           warp_tot[8];
    __shared__ unsigned long long base_smem;
    const uint64_t i = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    uint32_t       c[4] = {0, 0, 0, 0};
    if (i + 4 <= npos) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4*>(cur + i));
        c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
    } else {
        for (int k = 0; k < 4; ++k)
            if (i + k < npos) c[k] = cur[i + k];
    }
    const uint32_t n = (c[0] != 0) + (c[1] != 0) + (c[2] != 0) + (c[3] != 0);
    uint32_t incl = warp_inclusive_scan(n);
    if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < 8; ++w) {
            const uint32_t t = warp_tot[w];
            warp_tot[w]      = tot;
            tot += t;
        }
        base_smem = tot ? atomicAdd(cursor, (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    uint64_t o = base_smem + warp_tot[threadIdx.x >> 5] + incl - n;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (c[k] != 0) list_out[o++] = (uint32_t)(i + k);
}

__global__ void __launch_bounds__(256) iota_plus1_kernel(uint32_t* __restrict__ out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[i] = (uint32_t)i + 1;
}

size_t split_smem(int tile, uint32_t nbins) {
    return (size_t)tile * (8 + 4 + 2) + (size_t)nbins * 8;
}

template <class K>
void opt_in_smem(K kernel, size_t bytes) {
    // per device (a process may train on several): cheap enough to repeat per launch
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace

PartPlan part_plan(uint64_t bound) {
    PartPlan pl;
    uint64_t want = bound / 512 + 1;
    int      b    = 8;
    while ((1ull << b) < want && b < 22) ++b;
    pl.b1     = (b + 1) / 2;
    pl.b2     = b - pl.b1;
    pl.nparts = 1u << b;
    return pl;
}

int launch_part_hist(cudaStream_t s, const uint32_t* prev, const uint32_t* list, uint64_t nitems, uint32_t dense, uint32_t* dense_cnt, uint32_t* hist, const PartPlan& pl,
                     DeviceStats* st, int sms) {
    if (!nitems) return 0;
    const int      shift = 64 - pl.b1 - pl.b2;
    const unsigned grid  = (unsigned)std::min<uint64_t>((nitems + 255) / 256, (uint64_t)sms * 32);
    if (list) part_hist_kernel<true, false><<<grid, 256, 0, s>>>(prev, list, nitems, 0, nullptr, hist, shift, st);
    else if (dense) part_hist_kernel<false, true><<<grid, 256, 0, s>>>(prev, nullptr, nitems, dense, dense_cnt, hist, shift, st);
    else part_hist_kernel<false, false><<<grid, 256, 0, s>>>(prev, nullptr, nitems, 0, nullptr, hist, shift, st);
    return 1;
}

int launch_part_scan(cudaStream_t s, const uint32_t* hist, const PartPlan& pl, uint32_t* off, uint32_t* cursor2, uint32_t* cursor1, uint32_t* tstart) {
    part_scan_kernel<<<1, 1024, 0, s>>>(hist, pl.nparts, pl.b2, off, cursor2, cursor1, tstart);
    return 1;
}

int launch_part_split1(cudaStream_t s, const uint32_t* prev, const uint32_t* list, uint64_t nitems, uint32_t dense, const uint32_t* dense_id, uint32_t* cur, const PartPlan& pl,
                       uint32_t* cursor1, void* rk, uint32_t* rp) {
    if (!nitems) return 0;
    const uint32_t nbins  = 1u << pl.b1;
    const int      shift1 = 64 - pl.b1;
    const size_t   smem   = split_smem(kTile1, nbins);
    const unsigned grid   = (unsigned)((nitems + kTile1 - 1) / kTile1);
    auto*          keys   = static_cast<unsigned long long*>(rk);
    if (list) {
        opt_in_smem(part_split1_kernel<true, false>, smem);
        part_split1_kernel<true, false><<<grid, 256, smem, s>>>(prev, list, nitems, 0, nullptr, cur, shift1, nbins, cursor1, keys, rp);
    } else if (dense) {
        opt_in_smem(part_split1_kernel<false, true>, smem);
        part_split1_kernel<false, true><<<grid, 256, smem, s>>>(prev, nullptr, nitems, dense, dense_id, cur, shift1, nbins, cursor1, keys, rp);
    } else {
        opt_in_smem(part_split1_kernel<false, false>, smem);
        part_split1_kernel<false, false><<<grid, 256, smem, s>>>(prev, nullptr, nitems, 0, nullptr, cur, shift1, nbins, cursor1, keys, rp);
    }
    return 1;
}

int launch_part_split2(cudaStream_t s, const void* rk_in, const uint32_t* rp_in, const uint32_t* off, const uint32_t* tstart, const PartPlan& pl, uint64_t max_records,
                       uint32_t* cursor2, void* rk_out, uint32_t* rp_out) {
    const uint32_t p1n    = 1u << pl.b1;
    const int      shift2 = 64 - pl.b1 - pl.b2;
    const size_t   smem   = split_smem(kTile2, 1u << pl.b2);
    const unsigned grid   = (unsigned)(max_records / kTile2 + p1n + 1);  // every b1-partition rounds its tile count up
    opt_in_smem(part_split2_kernel, smem);
    part_split2_kernel<<<grid, 256, smem, s>>>(static_cast<const unsigned long long*>(rk_in), rp_in, off, tstart, p1n, pl.b2, shift2, cursor2,
                                                static_cast<unsigned long long*>(rk_out), rp_out);
    return 1;
}

int launch_part_count(cudaStream_t s, const void* rk, const uint32_t* rp, const uint32_t* off, const PartPlan& pl, uint32_t threshold, uint32_t* out, uint32_t id_mul,
                      uint32_t id_add, uint32_t* sv_pos, uint32_t* sv_cnt, DeviceStats* st, unsigned int* work /* zeroed */, int sms) {
    static int bps = 0;
    if (!bps) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, part_count_kernel, 256, 0);
        if (bps < 1) bps = 1;
    }
    const unsigned grid = (unsigned)std::min<uint64_t>(pl.nparts, (uint64_t)sms * bps);
    part_count_kernel<<<grid, 256, 0, s>>>(static_cast<const unsigned long long*>(rk), rp, off, pl.nparts, threshold, out, id_mul, id_add, sv_pos, sv_cnt, st, work);
    return 1;
}

int launch_compact_nonzero(cudaStream_t s, const uint32_t* cur, uint64_t npos, uint32_t* list_out, unsigned long long* cursor) {
    if (!npos) return 0;
    compact_nonzero_kernel<<<(unsigned)((npos + 1023) / 1024), 256, 0, s>>>(cur, npos, list_out, cursor);
    return 1;
}

int launch_iota_plus1(cudaStream_t s, uint32_t* out, uint64_t n) {
    if (!n) return 0;
    iota_plus1_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(out, n);
    return 1;
}

}  // namespace colibri
