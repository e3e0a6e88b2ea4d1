// partition.cu -- the partitioned counting path of a level n >= 2 (large levels; the hashed HBM table of kernels.cu stays for small ones).
//
// Replaces, per window, the two has() lookups + add() of the reference (include/patternmodel.h:1139-1161) and, per level, prune()
// (:2107-2128).  Round 1 counted every window with one random 32-byte-sector access into a table in HBM; ncu showed the level-2 launch
// moving 130 B of DRAM per table window at 38 % of the HBM peak (profiles/r02_ncu.md) -- random sectors are what HBM is worst at.  Here
// the windows of a level are radix-partitioned by the high bits of their key hash, twice (b1 + b2 bits), with streaming passes only, until
// a partition holds a few hundred windows; one WARP then counts a partition in a shared-memory table, applies the threshold there
// and hands every window of a surviving n-gram its id.  No table in HBM is ever probed, and no pass takes a global atomic per window:
//   A  part_hist      stream prev[] (or the position list): histogram of the b1 bits, privatised in shared memory;
//                     level 2: pairs of frequent classes are counted in the dense square instead and never become records
//   -  prune_dense    (kernels.cu) dense square -> survivors, dense_id[cell] = survivor index + 1
//   -  part_bases     exclusive scan of the 2^b1 counts (one small block)
//   B  part_split1    stream prev[] again: record (key, position) of every hashed window -> its b1-partition (block-staged in shared
//                     memory, one cursor reservation per bin per tile); writes cur[] = dense id or 0 for every position on the way
//   -  scan of the b1+b2-bit histogram pass B took on the way (REDs on an L2-resident array): the final partition offsets
//   D  part_split2    one block per b1-partition: its records go to their sub-partition in staged chunks through shared-memory cursors
//   E  part_count     one warp per final partition: shared-memory open addressing (64-bit CAS), threshold; the survivors stay in the
//                     partition's own range for now and name the ids: id = id_base + (offset of the partition) + (rank in it) + 1;
//                     cur[position] = id for the windows that stay
//   -  scan of the per-partition survivor counts, part_gather: survivors -> the level's segment, compacted
// ids are unique non-zero numbers (not table slots), a pruned window keeps id 0: no survivor bitmap, no relabel pass, no table memset,
// no table scan.  Exactness: keys are compared in full, partitions are exhaustive and disjoint (a key has one hash), counts are exact.
// A partition with more distinct keys than the shared-memory table holds raises kErrTableFull and the host reruns the level on the HBM-table
// path (never seen on hashed keys; partitions are sized for <= 384 windows on average (usually half of that), a table holds 512 keys).
#include <algorithm>

#include "device_utils.cuh"
#include "kernels.h"

namespace colibri {

namespace {

constexpr int      kTile1   = 4096;  // items per block of pass B (16 per thread, four 16-byte loads)
constexpr int      kTile2   = 4096;  // records per chunk of pass D (512 threads, 8 per thread)
constexpr uint32_t kHotSide = 64;    // pass A counts the pairs of the 64 most frequent classes in shared memory first

__device__ __forceinline__ uint32_t part_of(uint64_t h, int shift) {
    return (uint32_t)(h >> shift);
}

// ---- pass A ---------------------------------------------------------------------------------------------------------------------
// dynamic shared memory: hist1 u32[nbins] (| hot u32[64 * 64] when kDense)
template <bool kList, bool kDense>
__global__ void __launch_bounds__(256) part_hist_kernel(const uint32_t* __restrict__ prev, const uint32_t* __restrict__ list, uint64_t nitems, uint32_t dense,
                                                        uint32_t* __restrict__ dense_cnt, uint32_t* __restrict__ hist, uint32_t nbins, int shift1, DeviceStats* __restrict__ st) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint64_t scratch[8];
    uint32_t* h1  = reinterpret_cast<uint32_t*>(smem);
    uint32_t* hot = h1 + nbins;
    for (uint32_t i = threadIdx.x; i < nbins + (kDense ? kHotSide * kHotSide : 0); i += blockDim.x) h1[i] = 0;
    __syncthreads();
    uint32_t valid = 0;
    auto     one   = [&](uint32_t a, uint32_t b) {
        if (a == 0 || b == 0) return;
        ++valid;
        if (kDense && a < dense && b < dense) {
            if (a < kHotSide && b < kHotSide) atomicAdd(&hot[a * kHotSide + b], 1u);
            else atomicAdd(dense_cnt + a * dense + b, 1u);
            return;
        }
        atomicAdd(&h1[part_of(table_hash_u64(((unsigned long long)a << 32) | b), shift1)], 1u);
    };
    if (kList) {
        const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
        for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nitems; j += stride) {
            const uint32_t p = __ldcs(list + j);
            one(__ldg(prev + p), __ldg(prev + p + 1));
        }
    } else {
        // four positions per thread and iteration (one 16-byte load); the right neighbour of the fourth comes from the next lane.
        // A warp leaves the loop as a whole (its first lane's position decides), so the shuffle always sees 32 lanes.
        const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
        for (uint64_t p0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;; p0 += stride) {
            if (p0 - (uint64_t)lane_id() * 4 >= nitems) break;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (p0 < nitems) v = __ldcs(reinterpret_cast<const uint4*>(prev + p0));  // prev has nitems + 8 entries
            uint32_t nxt = __shfl_down_sync(0xffffffffu, v.x, 1);
            if (lane_id() == 31) nxt = p0 + 4 <= nitems ? __ldg(prev + p0 + 4) : 0u;
            if (p0 < nitems) one(v.x, v.y);
            if (p0 + 1 < nitems) one(v.y, v.z);
            if (p0 + 2 < nitems) one(v.z, v.w);
            if (p0 + 3 < nitems) one(v.w, nxt);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x)
        if (h1[i]) atomicAdd(hist + i, h1[i]);
    if (kDense) {
        for (uint32_t i = threadIdx.x; i < kHotSide * kHotSide; i += blockDim.x) {
            const uint32_t c = hot[i], a = i / kHotSide, b = i % kHotSide;
            if (c && a < dense && b < dense) atomicAdd(dense_cnt + a * dense + b, c);
        }
    }
    uint64_t v = block_reduce_sum(valid, scratch);
    if (threadIdx.x == 0 && v) atomicAdd(&st->valid_windows, (unsigned long long)v);
}

// ---- level-1 ids and pass A of level 2 in one sweep -----------------------------------------------------------------------------------
// make_id1 (kernels.cu) streams the tokens once to write id1[p] = class of p, or 0 when the class was pruned or p is a delimiter; level 2's pass A would
// stream id1 right afterwards to count the dense pairs and take the first-level histogram.  When level 2 is known to run on the partitioned path
// the two are one kernel: the atomics of pass A hide behind the token stream.
// dynamic shared memory: hist1 u32[nbins] | hot u32[64 * 64]
// keep[c] = 1 if class c stays at level 1 (count1[c] >= threshold), one bit per class: 12.5 KB for 100 000 classes, at home in L1 where the
// 400 KB of counts are not -- the id sweep below waits on its dependent loads, not on bandwidth
__global__ void __launch_bounds__(256) class_keep_bits_kernel(const uint32_t* __restrict__ count1, uint32_t nclasses, uint32_t threshold, uint32_t* __restrict__ keep) {
    const uint32_t i    = blockIdx.x * blockDim.x + threadIdx.x;
    const bool     k    = i != 0 && i < nclasses && count1[i] >= threshold;
    const uint32_t bits = __ballot_sync(0xffffffffu, k);
    if (lane_id() == 0) keep[i >> 5] = bits;
}

__global__ void __launch_bounds__(256) make_id1_hist_kernel(const uint32_t* __restrict__ tok, uint64_t n /* entries of id1 to write: positions + 1 */,
                                                            const uint32_t* __restrict__ keep, uint32_t* __restrict__ id1, uint32_t dense,
                                                            uint32_t* __restrict__ dense_cnt, uint32_t* __restrict__ hist, uint32_t nbins, int shift1, DeviceStats* __restrict__ st) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint64_t scratch[8];
    uint32_t* h1  = reinterpret_cast<uint32_t*>(smem);
    uint32_t* hot = h1 + nbins;
    for (uint32_t i = threadIdx.x; i < nbins + kHotSide * kHotSide; i += blockDim.x) h1[i] = 0;
    __syncthreads();
    uint32_t valid = 0;
    auto     idof  = [&](uint32_t c) { return ((__ldg(keep + (c >> 5)) >> (c & 31u)) & 1u) ? c : 0u; };
    auto     one   = [&](uint32_t a, uint32_t b) {
        if (a == 0 || b == 0) return;
        ++valid;
        if (a < dense && b < dense) {
            if (a < kHotSide && b < kHotSide) atomicAdd(&hot[a * kHotSide + b], 1u);
            else atomicAdd(dense_cnt + a * dense + b, 1u);
            return;
        }
        atomicAdd(&h1[part_of(table_hash_u64(((unsigned long long)a << 32) | b), shift1)], 1u);
    };
    // the tokens of the next round are on their way while this round's are looked up and counted (the token array has positions + 8 entries,
    // zeros behind the last position; lane 31 also needs the first token of the next lane group)
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
    auto fetch = [&](uint64_t p, uint4& t, uint32_t& edge) {
        t    = make_uint4(0, 0, 0, 0);
        edge = 0;
        if (p < n) t = __ldcs(reinterpret_cast<const uint4*>(tok + p));
        if (lane_id() == 31 && p + 4 < n) edge = __ldg(tok + p + 4);
    };
    uint64_t p0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    uint4    t;
    uint32_t edge;
    fetch(p0, t, edge);
    for (;; p0 += stride) {
        if (p0 - (uint64_t)lane_id() * 4 >= n) break;  // whole warps leave together: the shuffle below always sees 32 lanes
        uint4    tn;
        uint32_t edgen;
        fetch(p0 + stride, tn, edgen);
        const uint4 x = make_uint4(idof(t.x), idof(t.y), idof(t.z), idof(t.w));
        uint32_t nxt = __shfl_down_sync(0xffffffffu, x.x, 1);
        if (lane_id() == 31) nxt = idof(edge);
        if (p0 < n) {
            __stcs(reinterpret_cast<uint4*>(id1 + p0), x);  // (id1 has positions + 8 entries; what lies behind n is zero because the tokens there are)
            one(x.x, x.y);
            one(x.y, x.z);
            one(x.z, x.w);
            one(x.w, nxt);
        }
        t    = tn;
        edge = edgen;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x)
        if (h1[i]) atomicAdd(hist + i, h1[i]);
    for (uint32_t i = threadIdx.x; i < kHotSide * kHotSide; i += blockDim.x) {
        const uint32_t c = hot[i], a = i / kHotSide, b = i % kHotSide;
        if (c && a < dense && b < dense) atomicAdd(dense_cnt + a * dense + b, c);
    }
    uint64_t v = block_reduce_sum(valid, scratch);
    if (threadIdx.x == 0 && v) atomicAdd(&st->valid_windows, (unsigned long long)v);
}

// ---- exclusive scan of up to 2048 counts in one block: out[i] = base + sum of in[0..i), out[n] = base + total ------------------------------
// base_ptr (may be NULL): a device-side value added to everything (the survivors the dense square already produced)
__global__ void __launch_bounds__(1024) part_bases_kernel(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ out, uint32_t* __restrict__ out2,
                                                          const unsigned long long* __restrict__ base_ptr) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = base_ptr ? (uint32_t)*base_ptr : 0u;
    __syncthreads();
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t q = base + threadIdx.x;
        const uint32_t v = q < n ? in[q] : 0u;
        uint32_t incl = warp_inclusive_scan(v);
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t a = warp_tot[lane], s1 = warp_inclusive_scan(a);
            warp_tot[lane] = s1 - a;
        }
        __syncthreads();
        const uint32_t excl = carry + warp_tot[warp] + incl - v;
        if (q < n) {
            out[q] = excl;
            if (out2) out2[q] = excl;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

// ---- scan of nparts per-partition counts in three small launches (totals per group of 2^b2, scan of the group totals, scan inside each group)
__global__ void __launch_bounds__(256) part_totals_kernel(const uint32_t* __restrict__ cnt, int b2, uint32_t* __restrict__ totals) {
    __shared__ uint64_t scratch[8];
    const uint32_t n = 1u << b2;
    uint32_t       v = 0;
    for (uint32_t i = threadIdx.x; i < n; i += 256) v += cnt[((uint64_t)blockIdx.x << b2) + i];
    uint64_t t = block_reduce_sum(v, scratch);
    if (threadIdx.x == 0) totals[blockIdx.x] = (uint32_t)t;
}
__global__ void __launch_bounds__(256) part_offsets_kernel(const uint32_t* __restrict__ cnt, int b2, const uint32_t* __restrict__ group_base, uint32_t* __restrict__ off) {
    __shared__ uint32_t warp_tot[8];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = group_base[blockIdx.x];
    __syncthreads();
    const uint32_t n = 1u << b2, lane = lane_id(), warp = threadIdx.x >> 5;
    const uint64_t first = (uint64_t)blockIdx.x << b2;
    for (uint32_t base = 0; base < n; base += 256) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? cnt[first + i] : 0u;
        uint32_t incl = warp_inclusive_scan(v);
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (uint32_t w = 0; w < 8; ++w) {
            const uint32_t c = warp_tot[w];
            if (w < warp) before += c;
            total += c;
        }
        const uint32_t excl = carry + before + incl - v;
        if (i < n) off[first + i] = excl;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (blockIdx.x + 1 == gridDim.x && threadIdx.x == 0) off[first + n] = carry;  // off[nparts] = grand total
}

// ---- block-level multi-split: bins[0..nbins) = per-bin record count of the tile on entry; on exit bins[b] = first staging index of bin b,
// gdelta[b] = (global index of the bin's run) - (staging index); returns the tile's record total.  The run is reserved with one atomicAdd per
// non-empty bin on cursor[] (global memory, pass B) or, kLocal, by advancing the block's own shared-memory cursors (pass D).
template <bool kLocal>
__device__ __forceinline__ uint32_t split_reserve(uint32_t* bins, uint32_t* gdelta, uint32_t nbins, uint32_t* cursor, uint32_t* warp_tot) {
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
            if (local[k]) {
                if (kLocal) {
                    gdelta[b] = cursor[b] - run;
                    cursor[b] += local[k];
                } else {
                    gdelta[b] = atomicAdd(cursor + b, local[k]) - run;
                }
            }
            run += local[k];
        }
    }
    __syncthreads();
    return total;
}

// ---- pass B ---------------------------------------------------------------------------------------------------------------------
// dynamic shared memory: stage_key u64[kTile1] | stage_pos u32[kTile1] | stage_bin u16[kTile1] | bins u32[nbins] | gdelta u32[nbins]
template <bool kList, bool kDense>
__global__ void __launch_bounds__(256, 3) part_split1_kernel(const uint32_t* __restrict__ prev, const uint32_t* __restrict__ list, uint64_t nitems, uint32_t dense,
                                                             const uint32_t* __restrict__ dense_id, uint32_t* __restrict__ cur, int shift1, int b2, uint32_t nbins,
                                                             uint32_t* __restrict__ cursor1, uint32_t* __restrict__ hist2, unsigned long long* __restrict__ rk,
                                                             uint32_t* __restrict__ rp) {
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
    uint32_t       binrank[16];  // bin << 16 | rank inside (tile, bin); 0xFFFFFFFF = no record (<= 4096 records per tile: the rank fits 16 bits)
    if (!kList) {
        // four 16-byte loads per thread: positions tile + k * 1024 + 4 * tid .. + 3; the right neighbour of the fourth comes from the next lane
        uint32_t a[16], nx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint64_t p0 = tile + (uint64_t)k * 1024 + 4u * threadIdx.x;
            uint4          v  = make_uint4(0, 0, 0, 0);
            if (p0 < nitems) v = __ldcs(reinterpret_cast<const uint4*>(prev + p0));  // prev has nitems + 8 entries
            uint32_t nxt = __shfl_down_sync(0xffffffffu, v.x, 1);
            if (lane_id() == 31) nxt = p0 + 4 <= nitems ? __ldg(prev + p0 + 4) : 0u;
            a[4 * k] = v.x; a[4 * k + 1] = v.y; a[4 * k + 2] = v.z; a[4 * k + 3] = v.w;
            nx[k] = nxt;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (p0 + e >= nitems) a[4 * k + e] = 0;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t dense_out[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t x = a[4 * k + e], y = e < 3 ? a[4 * k + e + 1] : nx[k];
                binrank[4 * k + e] = 0xFFFFFFFFu;
                dense_out[e]       = 0;
                if (x == 0 || y == 0) continue;
                if (kDense && x < dense && y < dense) {
                    dense_out[e] = __ldg(dense_id + x * dense + y);
                    continue;
                }
                const uint32_t part = part_of(table_hash_u64(((unsigned long long)x << 32) | y), shift1 - b2), bin = part >> b2;
                atomicAdd(hist2 + part, 1u);  // result unused -> RED: the final partitions' sizes, for pass D
                binrank[4 * k + e] = (bin << 16) | atomicAdd(&bins[bin], 1u);
            }
            const uint64_t p0 = tile + (uint64_t)k * 1024 + 4u * threadIdx.x;
            if (p0 < nitems) __stcs(reinterpret_cast<uint4*>(cur + p0), make_uint4(dense_out[0], dense_out[1], dense_out[2], dense_out[3]));
        }
        __syncthreads();
        const uint32_t total = split_reserve<false>(bins, gdelta, nbins, cursor1, warp_tot);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t br = binrank[4 * k + e];
                if (br == 0xFFFFFFFFu) continue;
                const uint32_t x = a[4 * k + e], y = e < 3 ? a[4 * k + e + 1] : nx[k];
                const uint32_t bin = br >> 16, idx = bins[bin] + (br & 0xFFFFu);
                stage_key[idx]     = ((unsigned long long)x << 32) | y;
                stage_pos[idx]     = (uint32_t)(tile + (uint64_t)k * 1024 + 4u * threadIdx.x + e);
                stage_bin[idx]     = (uint16_t)bin;
            }
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < total; i += 256) {
            const uint32_t dst = gdelta[stage_bin[i]] + i;
            rk[dst]            = stage_key[i];
            rp[dst]            = stage_pos[i];
        }
        return;
    }
    // list mode: item j is position list[j]; cur was zeroed by the host
    uint32_t a[16], b[16], pos[16];
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
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        binrank[k] = 0xFFFFFFFFu;
        if (a[k] == 0 || b[k] == 0) continue;
        const uint32_t part = part_of(table_hash_u64(((unsigned long long)a[k] << 32) | b[k]), shift1 - b2), bin = part >> b2;
        atomicAdd(hist2 + part, 1u);
        binrank[k]         = (bin << 16) | atomicAdd(&bins[bin], 1u);
    }
    __syncthreads();
    const uint32_t total = split_reserve<false>(bins, gdelta, nbins, cursor1, warp_tot);
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
// One block per b1-partition q (records off[q << b2] .. off[(q + 1) << b2]): chunks of kTile2 records are ranked by their next b2 hash bits,
// grouped per bin in shared memory and copied to their sub-partition through the block's own cursors (= the final partition offsets, from the
// scan of the histogram pass B took).  No global atomic.
// dynamic shared memory: stage_key u64[kTile2] | stage_pos u32[kTile2] | stage_bin u16[kTile2] | cursors u32[nbins] | bins u32[nbins] | gdelta u32[nbins]
__global__ void __launch_bounds__(512, 2) part_split2_kernel(const unsigned long long* __restrict__ rk_in, const uint32_t* __restrict__ rp_in, const uint32_t* __restrict__ off,
                                                             int b2, int shift2, unsigned long long* __restrict__ rk_out, uint32_t* __restrict__ rp_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t      nbins     = 1u << b2;
    unsigned long long* stage_key = reinterpret_cast<unsigned long long*>(smem);
    uint32_t*           stage_pos = reinterpret_cast<uint32_t*>(stage_key + kTile2);
    uint16_t*           stage_bin = reinterpret_cast<uint16_t*>(stage_pos + kTile2);
    uint32_t*           cursors   = reinterpret_cast<uint32_t*>(stage_bin + kTile2);
    uint32_t*           bins      = cursors + nbins;
    uint32_t*           gdelta    = bins + nbins;
    __shared__ uint32_t warp_tot[16];
    const uint64_t first = (uint64_t)blockIdx.x << b2;
    const uint32_t beg = __ldg(off + first), end = __ldg(off + first + nbins);
    for (uint32_t i = threadIdx.x; i < nbins; i += 512) cursors[i] = __ldg(off + first + i);
    for (uint32_t c0 = beg; c0 < end; c0 += kTile2) {
        for (uint32_t i = threadIdx.x; i < nbins; i += 512) bins[i] = 0;
        __syncthreads();
        unsigned long long key[8];
        uint32_t           pos[8], binrank[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t i = c0 + k * 512 + threadIdx.x;
            key[k]           = 0;
            pos[k]           = 0;
            if (i < end) {
                key[k] = __ldcs(rk_in + i);
                pos[k] = __ldcs(rp_in + i);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            binrank[k] = 0xFFFFFFFFu;
            if (key[k]) {
                const uint32_t bin = part_of(table_hash_u64(key[k]), shift2) & (nbins - 1);
                binrank[k]         = (bin << 16) | atomicAdd(&bins[bin], 1u);
            }
        }
        __syncthreads();
        // exclusive scan of the chunk's bin counts (512 threads, <= 4 bins each); the runs go where the block's cursors point
        uint32_t total;
        {
            const uint32_t per = (nbins + 511) / 512;
            uint32_t       local[4], sum = 0;
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
                const uint32_t b = threadIdx.x * per + k;
                local[k]         = (k < per && b < nbins) ? bins[b] : 0u;
                sum += local[k];
            }
            uint32_t incl = warp_inclusive_scan(sum);
            if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = incl;
            __syncthreads();
            uint32_t before = 0;
            total           = 0;
#pragma unroll
            for (uint32_t w = 0; w < 16; ++w) {
                const uint32_t c = warp_tot[w];
                if (w < (threadIdx.x >> 5)) before += c;
                total += c;
            }
            uint32_t run = before + incl - sum;
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
                const uint32_t b = threadIdx.x * per + k;
                if (k < per && b < nbins) {
                    bins[b] = run;
                    if (local[k]) {
                        gdelta[b] = cursors[b] - run;
                        cursors[b] += local[k];
                    }
                    run += local[k];
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (binrank[k] == 0xFFFFFFFFu) continue;
            const uint32_t bin = binrank[k] >> 16, idx = bins[bin] + (binrank[k] & 0xFFFFu);
            stage_key[idx]     = key[k];
            stage_pos[idx]     = pos[k];
            stage_bin[idx]     = (uint16_t)bin;
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < total; i += 512) {
            const uint32_t dst = gdelta[stage_bin[i]] + i;
            rk_out[dst]        = stage_key[i];
            rp_out[dst]        = stage_pos[i];
        }
        __syncthreads();
    }
}

// ---- pass E ---------------------------------------------------------------------------------------------------------------------
// One WARP per partition (a partition is a few hundred records: a block would spend its time in barriers and in the latency of one
// partition's loads; 24 independent warps per SM overlap them).  Each warp owns a 512-slot table in shared memory.  The survivors of
// partition q stay inside the partition's own record range for now -- tmp_pos / tmp_cnt [off[q] + r], r = 0 .. kept[q] -- and that place names
// the n-gram: id = (id_base + off[q] + r) * id_mul + id_add.  No global atomic at all (one compaction cursor for all partitions costs ~5 ns per
// partition, serialised: 1.4 of this kernel's 1.6 ms when it had one).  out[rp[i]] = id for the records of surviving keys (single GPU:
// id_mul = id_add = 1, out = cur; owner of a sharded run: out = the reply array, global ids).
constexpr int      kHeld      = 8;    // records per lane kept in registers between the two phases (partitions of <= 256 records never re-read)
constexpr uint32_t kWarpSlots = 512;  // per-warp table: keys u64 | counts u32 = 6 KB

__device__ __forceinline__ uint32_t smem_upsert(unsigned long long* tk, uint32_t* tc, uint32_t mask, unsigned long long key, bool& claimed, bool& full) {
    uint32_t slot = (uint32_t)table_hash_u64(key) & mask;
    claimed       = false;
    for (uint32_t step = 0; step <= mask; ++step) {
        const unsigned long long old = atomicCAS(tk + slot, 0ull, key);
        if (old == 0ull) claimed = true;
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
                                                         uint32_t nparts, uint32_t threshold, uint32_t* __restrict__ out, uint32_t id_base, uint32_t id_mul, uint32_t id_add,
                                                         uint32_t* __restrict__ tmp_pos, uint32_t* __restrict__ tmp_cnt, uint32_t* __restrict__ kept_of, DeviceStats* __restrict__ st) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint64_t scratch[8];
    const uint32_t      warp = threadIdx.x >> 5, lane = lane_id();
    unsigned long long* tk = reinterpret_cast<unsigned long long*>(smem) + (size_t)warp * kWarpSlots;
    uint32_t*           tc = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned long long*>(smem) + 8 * kWarpSlots) + (size_t)warp * kWarpSlots;
    uint32_t found = 0, kept = 0, singles = 0;
    uint64_t occ = 0;
    bool     full = false;
    const uint32_t nwarps = gridDim.x * 8, lt = (1u << lane) - 1;
    uint32_t q = blockIdx.x * 8 + warp;
    uint32_t beg = 0, end = 0;
    if (q < nparts) {
        beg = __ldg(off + q);
        end = __ldg(off + q + 1);
    }
    for (; q < nparts; q += nwarps) {
        const uint32_t n = end - beg, cbeg = beg, cend = end;
        // the next partition's bounds are on their way while this one is counted
        if (q + nwarps < nparts) {
            beg = __ldg(off + q + nwarps);
            end = __ldg(off + q + nwarps + 1);
        }
        if (n == 0) {
            if (lane == 0) kept_of[q] = 0;
            continue;
        }
        uint32_t size = 64;
        while (size < 2 * n && size < kWarpSlots) size <<= 1;
        const uint32_t mask = size - 1;
        if (n <= kHeld * 32) {
            // ---- the usual case: every record stays in a register.  All loads are issued before the first table access (one memory latency per
            // partition, not one per 32 records).  The record whose CAS claimed a slot stands for its key afterwards: the claimers enumerate
            // the distinct keys, so the threshold and the compaction walk the records, not the table.
            unsigned long long key[kHeld];
            uint32_t           held_pos[kHeld], held_slot[kHeld], claims = 0;
#pragma unroll
            for (int k = 0; k < kHeld; ++k) {
                const uint32_t i = cbeg + k * 32 + lane;
                key[k]           = 0ull;
                held_pos[k]      = 0;
                if (i < cend) {
                    key[k]      = __ldcs(rk + i);
                    held_pos[k] = __ldcs(rp + i);
                }
            }
            for (uint32_t i = lane; i < size; i += 32) {
                tk[i] = 0ull;
                tc[i] = 0u;
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < kHeld; ++k) {
                held_slot[k] = 0xFFFFFFFFu;
                if (key[k] != 0ull) {
                    bool claimed;
                    held_slot[k] = smem_upsert(tk, tc, mask, key[k], claimed, full);
                    if (claimed) claims |= 1u << k;
                }
            }
            __syncwarp();
            uint32_t run = cbeg;  // the survivors fit the partition's own range: kept <= distinct keys <= records
#pragma unroll
            for (int k = 0; k < kHeld; ++k) {
                if (k * 32 < (int)n) {
                    const bool     cl   = (claims >> k) & 1u;
                    const uint32_t c    = cl ? tc[held_slot[k]] : 0u;
                    const bool     keep = cl && c >= threshold;
                    const uint32_t m    = __ballot_sync(0xffffffffu, keep);
                    if (cl) {
                        ++found;
                        singles += c == 1;
                        uint32_t id = 0;
                        if (keep) {
                            const uint32_t o = run + __popc(m & lt);
                            tmp_pos[o]       = held_pos[k];
                            tmp_cnt[o]       = c;
                            id               = (id_base + o) * id_mul + id_add;
                            occ += c;
                            ++kept;
                        }
                        tc[held_slot[k]] = id;  // the slot now answers "which id", 0 = pruned
                    }
                    run += __popc(m);
                }
            }
            if (lane == 0) kept_of[q] = run - cbeg;
            __syncwarp();
            if (run != cbeg) {
#pragma unroll
                for (int k = 0; k < kHeld; ++k) {
                    if (held_slot[k] == 0xFFFFFFFFu) continue;
                    const uint32_t id = tc[held_slot[k]];
                    if (id) out[held_pos[k]] = id;
                }
            }
            __syncwarp();  // the table is cleared for the next partition
            continue;
        }
        // ---- a partition with more records than the registers hold (a frequent key landed here): records are read twice, the table is scanned
        for (uint32_t i = lane; i < size; i += 32) {
            tk[i] = 0ull;
            tc[i] = 0u;
        }
        __syncwarp();
        for (uint32_t c0 = cbeg; c0 < cend; c0 += kHeld * 32) {  // loads in batches of eight, as above: one memory latency per 256 records
            unsigned long long key[kHeld];
#pragma unroll
            for (int k = 0; k < kHeld; ++k) {
                const uint32_t i = c0 + k * 32 + lane;
                key[k]           = i < cend ? __ldg(rk + i) : 0ull;
            }
#pragma unroll
            for (int k = 0; k < kHeld; ++k) {
                if (key[k] == 0ull) continue;
                // a frequent key fills whole warps: one lane per distinct key of the warp does the table work for all of them
                const uint32_t peers = __match_any_sync(__activemask(), key[k]);
                if ((int)lane == __ffs(peers) - 1) {
                    bool           claimed;
                    const uint32_t sl = smem_upsert(tk, tc, mask, key[k], claimed, full);
                    if (sl != 0xFFFFFFFFu && __popc(peers) > 1) atomicAdd(tc + sl, (uint32_t)__popc(peers) - 1);
                }
            }
        }
        __syncwarp();
        uint32_t nkeep = 0;
        for (uint32_t sl = lane; sl < size; sl += 32) {
            if (tk[sl] != 0ull) {
                const uint32_t c = tc[sl];
                ++found;
                singles += c == 1;
                if (c >= threshold) {
                    ++nkeep;
                    occ += c;
                }
            }
        }
        kept += nkeep;
        const uint32_t incl  = warp_inclusive_scan(nkeep);
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 0) kept_of[q] = total;
        uint32_t o = cbeg + incl - nkeep;
        for (uint32_t sl = lane; sl < size; sl += 32) {
            uint32_t v = 0;
            if (tk[sl] != 0ull) {
                const uint32_t c = tc[sl];
                if (c >= threshold) {
                    tmp_cnt[o] = c;
                    v          = o + 1;  // where the survivor lives; its position is filled in by any of its records below
                    ++o;
                }
            }
            tc[sl] = v;
        }
        __syncwarp();
        if (total) {
            for (uint32_t c0 = cbeg; c0 < cend; c0 += kHeld * 32) {
                unsigned long long key[kHeld];
                uint32_t           pos[kHeld];
#pragma unroll
                for (int k = 0; k < kHeld; ++k) {
                    const uint32_t i = c0 + k * 32 + lane;
                    key[k]           = i < cend ? __ldg(rk + i) : 0ull;
                    pos[k]           = i < cend ? __ldg(rp + i) : 0u;
                }
#pragma unroll
                for (int k = 0; k < kHeld; ++k) {
                    if (key[k] == 0ull) continue;
                    const uint32_t sl = smem_find(tk, mask, key[k]);
                    if (sl != 0xFFFFFFFFu) {
                        const uint32_t v = tc[sl];
                        if (v) {
                            tmp_pos[v - 1] = pos[k];  // every occurrence names the n-gram equally well: whichever store lands last stays
                            out[pos[k]]    = (id_base + v - 1) * id_mul + id_add;
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
    uint64_t f = block_reduce_sum(found, scratch);
    uint64_t k = block_reduce_sum(kept, scratch);
    uint64_t c = block_reduce_sum(occ, scratch);
    uint64_t g = block_reduce_sum(singles, scratch);
    if (threadIdx.x == 0) {
        if (f) atomicAdd(&st->found, (unsigned long long)f);
        if (k) atomicAdd(&st->kept, (unsigned long long)k);
        if (c) atomicAdd(&st->kept_occ, (unsigned long long)c);
        if (g) atomicAdd(&st->singletons, (unsigned long long)g);
    }
    if (full) atomicOr(&st->errflags, kErrTableFull);
}

// survivors of partition q: tmp[off[q] .. + kept_of[q]) -> sv[dst_off[q] ..); slot_index (indexed models): id - 1 -> survivor index + 1
__global__ void __launch_bounds__(256) part_gather_kernel(const uint32_t* __restrict__ tmp_pos, const uint32_t* __restrict__ tmp_cnt, const uint32_t* __restrict__ off,
                                                          const uint32_t* __restrict__ kept_of, const uint32_t* __restrict__ dst_off, uint32_t nparts, uint32_t* __restrict__ sv_pos,
                                                          uint32_t* __restrict__ sv_cnt, uint32_t* __restrict__ slot_index, uint32_t id_base) {
    const uint32_t lane = lane_id();
    const uint32_t q    = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= nparts) return;
    const uint32_t n = __ldg(kept_of + q), src = __ldg(off + q), dst = __ldg(dst_off + q);
    for (uint32_t r = lane; r < n; r += 32) {
        sv_pos[dst + r] = tmp_pos[src + r];
        sv_cnt[dst + r] = tmp_cnt[src + r];
        if (slot_index) slot_index[id_base + src + r] = dst + r + 1;
    }
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

size_t split_smem(int tile, uint32_t nbins, int bin_arrays) {
    return (size_t)tile * (8 + 4 + 2) + (size_t)nbins * 4 * bin_arrays;
}

template <class K>
void opt_in_smem(K kernel, size_t bytes) {
    // per device (a process may train on several): cheap enough to repeat per launch
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace

PartPlan part_plan(uint64_t bound) {
    PartPlan pl;
    uint64_t want = bound / 384 + 1;
    int      b    = 8;
    while ((1ull << b) < want && b < 22) ++b;
    pl.b1     = (b + 1) / 2;
    pl.b2     = b - pl.b1;
    pl.nparts = 1u << b;
    return pl;
}

int launch_part_hist(cudaStream_t s, const uint32_t* prev, const uint32_t* list, uint64_t nitems, uint32_t dense, uint32_t* dense_cnt, uint32_t* hist1, const PartPlan& pl,
                     DeviceStats* st, int sms) {
    if (!nitems) return 0;
    const uint32_t nbins  = 1u << pl.b1;
    const int      shift1 = 64 - pl.b1;
    const size_t   smem   = (size_t)nbins * 4 + (dense && !list ? kHotSide * kHotSide * 4 : 0);
    const uint64_t per    = list ? 256 : 1024;
    const unsigned grid   = (unsigned)std::min<uint64_t>((nitems + per - 1) / per, (uint64_t)sms * 8);
    if (list) part_hist_kernel<true, false><<<grid, 256, smem, s>>>(prev, list, nitems, 0, nullptr, hist1, nbins, shift1, st);
    else if (dense) part_hist_kernel<false, true><<<grid, 256, smem, s>>>(prev, nullptr, nitems, dense, dense_cnt, hist1, nbins, shift1, st);
    else part_hist_kernel<false, false><<<grid, 256, smem, s>>>(prev, nullptr, nitems, 0, nullptr, hist1, nbins, shift1, st);
    return 1;
}

int launch_make_id1_hist(cudaStream_t s, const uint32_t* tok, uint64_t npos, const uint32_t* count1, uint32_t nclasses, uint32_t threshold, uint32_t* keep_bits, uint32_t* id1,
                         uint32_t dense, uint32_t* dense_cnt, uint32_t* hist1, const PartPlan& pl, DeviceStats* st, int sms) {
    const uint32_t nbins  = 1u << pl.b1;
    const size_t   smem   = (size_t)nbins * 4 + kHotSide * kHotSide * 4;
    const uint64_t n      = npos + 1;
    const unsigned grid   = (unsigned)std::min<uint64_t>((n + 1023) / 1024, (uint64_t)sms * 8);
    class_keep_bits_kernel<<<(nclasses + 255) / 256, 256, 0, s>>>(count1, nclasses, threshold, keep_bits);
    make_id1_hist_kernel<<<grid, 256, smem, s>>>(tok, n, keep_bits, id1, dense, dense_cnt, hist1, nbins, 64 - pl.b1, st);
    return 2;
}

int launch_part_bases(cudaStream_t s, const uint32_t* counts, uint32_t n, uint32_t* out, uint32_t* out2, const unsigned long long* base_ptr) {
    part_bases_kernel<<<1, 1024, 0, s>>>(counts, n, out, out2, base_ptr);
    return 1;
}

int launch_part_split1(cudaStream_t s, const uint32_t* prev, const uint32_t* list, uint64_t nitems, uint32_t dense, const uint32_t* dense_id, uint32_t* cur, const PartPlan& pl,
                       uint32_t* cursor1, uint32_t* hist2, void* rk, uint32_t* rp) {
    if (!nitems) return 0;
    const uint32_t nbins  = 1u << pl.b1;
    const int      shift1 = 64 - pl.b1;
    const size_t   smem   = split_smem(kTile1, nbins, 2);
    const unsigned grid   = (unsigned)((nitems + kTile1 - 1) / kTile1);
    auto*          keys   = static_cast<unsigned long long*>(rk);
    if (list) {
        opt_in_smem(part_split1_kernel<true, false>, smem);
        part_split1_kernel<true, false><<<grid, 256, smem, s>>>(prev, list, nitems, 0, nullptr, cur, shift1, pl.b2, nbins, cursor1, hist2, keys, rp);
    } else if (dense) {
        opt_in_smem(part_split1_kernel<false, true>, smem);
        part_split1_kernel<false, true><<<grid, 256, smem, s>>>(prev, nullptr, nitems, dense, dense_id, cur, shift1, pl.b2, nbins, cursor1, hist2, keys, rp);
    } else {
        opt_in_smem(part_split1_kernel<false, false>, smem);
        part_split1_kernel<false, false><<<grid, 256, smem, s>>>(prev, nullptr, nitems, 0, nullptr, cur, shift1, pl.b2, nbins, cursor1, hist2, keys, rp);
    }
    return 1;
}

int launch_part_scan(cudaStream_t s, const uint32_t* counts, const PartPlan& pl, uint32_t* group_tot, uint32_t* group_base, uint32_t* off, const unsigned long long* base_ptr) {
    const uint32_t p1n = 1u << pl.b1;
    part_totals_kernel<<<p1n, 256, 0, s>>>(counts, pl.b2, group_tot);
    part_bases_kernel<<<1, 1024, 0, s>>>(group_tot, p1n, group_base, nullptr, base_ptr);
    part_offsets_kernel<<<p1n, 256, 0, s>>>(counts, pl.b2, group_base, off);
    return 3;
}

int launch_part_split2(cudaStream_t s, const void* rk_in, const uint32_t* rp_in, const uint32_t* off, const PartPlan& pl, void* rk_out, uint32_t* rp_out) {
    const uint32_t p1n    = 1u << pl.b1;
    const int      shift2 = 64 - pl.b1 - pl.b2;
    const size_t   smem   = split_smem(kTile2, 1u << pl.b2, 3);
    opt_in_smem(part_split2_kernel, smem);
    part_split2_kernel<<<p1n, 512, smem, s>>>(static_cast<const unsigned long long*>(rk_in), rp_in, off, pl.b2, shift2, static_cast<unsigned long long*>(rk_out), rp_out);
    return 1;
}

int launch_part_count(cudaStream_t s, const void* rk, const uint32_t* rp, const uint32_t* off, const PartPlan& pl, uint32_t threshold, uint32_t* out, uint32_t id_base,
                      uint32_t id_mul, uint32_t id_add, uint32_t* tmp_pos, uint32_t* tmp_cnt, uint32_t* kept_of, DeviceStats* st, int sms) {
    const size_t smem = (size_t)8 * kWarpSlots * 12;
    opt_in_smem(part_count_kernel, smem);
    int bps = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, part_count_kernel, 256, smem);
    if (bps < 1) bps = 1;
    const unsigned grid = (unsigned)std::min<uint64_t>((pl.nparts + 7) / 8, (uint64_t)sms * bps);
    part_count_kernel<<<grid, 256, smem, s>>>(static_cast<const unsigned long long*>(rk), rp, off, pl.nparts, threshold, out, id_base, id_mul, id_add, tmp_pos, tmp_cnt, kept_of, st);
    return 1;
}

int launch_part_gather(cudaStream_t s, const uint32_t* tmp_pos, const uint32_t* tmp_cnt, const uint32_t* off, const uint32_t* kept_of, const PartPlan& pl, uint32_t* group_tot,
                       uint32_t* group_base, uint32_t* dst_off, const unsigned long long* first /* device: survivors already in the segment */, uint32_t* sv_pos, uint32_t* sv_cnt,
                       uint32_t* slot_index, uint32_t id_base) {
    launch_part_scan(s, kept_of, pl, group_tot, group_base, dst_off, first);
    part_gather_kernel<<<(pl.nparts + 7) / 8, 256, 0, s>>>(tmp_pos, tmp_cnt, off, kept_of, dst_off, pl.nparts, sv_pos, sv_cnt, slot_index, id_base);
    return 4;
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
