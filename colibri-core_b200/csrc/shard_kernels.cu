// shard_kernels.cu -- device side of the multi-GPU path (nothing in the reference corresponds to this).
//
// "Ship the windows to their owners": a rank never aggregates keys it does not own.  Per level n >= 2
//   sender  split      every valid window's 8-byte key goes into the send buffer of owner(key) = hash(key) mod G, with a STABLE
//                      multi-split (corpus order inside every destination group).  rec_of_pos[p] remembers where window p went.
//   [all-to-all of keys]
//   owner   filter + count   exactly the single-GPU kernels, fed from the received key stream instead of a pair of id arrays:
//                      occurrence filter in L2, HBM table for the rest, rid[i] = slot + 1 of received window i
//           prune      threshold scan (the single-GPU kernel); owner_reply turns rid[] into global ids in place
//   [all-to-all back, 4 bytes per window, same routes]
//   sender  relabel    id[p] = reply[rec_of_pos[p]] -- G ascending read streams, no random HBM access on the sender at all
// Survivors travel back as (index inside the sender's group, global count) to the rank whose window claimed the slot; that
// rank owns the tokens and exports the pattern.  The only random HBM traffic of a level is the owner's table upserts -- the
// same amount as on one GPU.
#include "device_utils.cuh"
#include "kernels.h"

namespace colibri {

static inline unsigned sk_div_up(uint64_t a, uint64_t b) {
    return (unsigned)((a + b - 1) / b);
}
static inline uint64_t sk_min(uint64_t a, uint64_t b) {
    return a < b ? a : b;
}

// Which rank owns a key.  Evaluated three times per position per level on the sender, so it is a cheap multiply-xorshift
// mix (the table slot inside the owner still comes from SpookyV2); only uniformity over G ranks matters here.
__device__ __forceinline__ uint32_t owner_of_key(unsigned long long key, uint32_t world) {
    unsigned long long z = key * 0x9E3779B97F4A7C15ull;
    z ^= z >> 32;
    z *= 0xD6E8FEB86659FD93ull;
    return (uint32_t)__umul64hi(z, (unsigned long long)world);
}

constexpr int      kSplitTile = 4096;  // positions per block: 8 warps x 512, each warp walks its slice in order
constexpr uint32_t kNoRec     = 0xFFFFFFFFu;
constexpr uint32_t kDenseRec  = 0x80000000u;  // rec_of_pos: the window is a pair of frequent classes, low bits = its cell of the dense square
constexpr uint32_t kDenseDest = 255u;
constexpr uint32_t kHotSide   = 64;           // the pairs of the 64 most frequent classes are counted in shared memory first

// Dense pairs (level 2, as on one GPU -- kernels.cu): a window of two classes below `dense` is never shipped.  Every rank counts such pairs
// in its own dense square, the squares are summed by one all-reduce, and every rank derives the same verdict from the global square:
// id = cell + 1 if the global count reaches the threshold, else 0.  The ids of the hashed n-grams start above dense^2.
// item j of a level: position j, or -- list mode, the sparse later levels -- position list[j] (the positions whose (n-1)-gram survived, left behind
// by the previous level's finish).  In list mode the per-position arrays (rec_of_pos) are indexed by ITEM.
__device__ __forceinline__ uint32_t window_dest(const uint32_t* __restrict__ prev, const uint32_t* __restrict__ list, uint64_t j, uint64_t nitems, uint32_t world,
                                                unsigned long long& key, uint32_t dense, uint32_t& p) {
    if (j >= nitems) return 256u;
    p = list ? __ldg(list + j) : (uint32_t)j;
    uint32_t a = prev[p], b = prev[p + 1];
    if (a == 0 || b == 0) return 256u;
    if (a < dense && b < dense) {
        key = (unsigned long long)a * dense + b;
        return kDenseDest;
    }
    key = ((unsigned long long)a << 32) | b;
    return owner_of_key(key, world);
}

// pass 1: per block, windows per destination (destination-major so that one exclusive scan yields every block's bases)
__global__ void __launch_bounds__(256) split_count_kernel(const uint32_t* __restrict__ prev, const uint32_t* __restrict__ list, uint64_t npos /* items */, uint32_t world,
                                                          uint32_t nblocks, uint32_t* __restrict__ hist, uint32_t dense, uint32_t* __restrict__ dense_cnt) {
    __shared__ uint32_t h[64];
    __shared__ uint32_t hot[kHotSide * kHotSide];
    if (threadIdx.x < 64) h[threadIdx.x] = 0;
    if (dense)
        for (uint32_t i = threadIdx.x; i < kHotSide * kHotSide; i += 256) hot[i] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kSplitTile;
#pragma unroll 4
    for (int k = 0; k < kSplitTile / 256; ++k) {
        unsigned long long key;
        uint32_t           p;
        uint32_t           d = window_dest(prev, list, base + (uint64_t)k * 256 + threadIdx.x, npos, world, key, dense, p);
        if (d < 64) atomicAdd(&h[d], 1u);
        else if (d == kDenseDest) {
            const uint32_t cell = (uint32_t)key, a = cell / dense, b = cell % dense;
            if (a < kHotSide && b < kHotSide) atomicAdd(&hot[a * kHotSide + b], 1u);
            else atomicAdd(dense_cnt + cell, 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < world) hist[(uint64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
    if (dense)
        for (uint32_t i = threadIdx.x; i < kHotSide * kHotSide; i += 256) {
            const uint32_t c = hot[i], a = i / kHotSide, b = i % kHotSide;
            if (c && a < dense && b < dense) atomicAdd(dense_cnt + a * dense + b, c);
        }
}

// pass 2: stable scatter.  send_keys is laid out [dest 0 | dest 1 | ...]; positions stay ascending inside a group.
// P2P mode (peer_keys != NULL): the keys are stored straight into the owner's receive slot for this rank over NVLink
// (peer_keys[d] = base of rank d's receive buffer, slot of source r at r * slot_cap) -- the pack and the all-to-all are one kernel.
// The block first gathers its keys per destination in shared memory and then copies each destination's run with consecutive
// threads: 256-byte warp stores instead of 8-byte ones scattered over G streams (NVLink packets like them long).
__global__ void __launch_bounds__(256) split_write_kernel(const uint32_t* __restrict__ prev, const uint32_t* __restrict__ list, uint64_t npos /* items */, uint32_t world,
                                                          uint32_t nblocks, const uint64_t* __restrict__ hist_off,
                                                          unsigned long long* __restrict__ send_keys, uint32_t* __restrict__ pos_of_rec, uint32_t* __restrict__ rec_of_pos,
                                                          unsigned long long* const* __restrict__ peer_keys, uint32_t my_rank, uint64_t slot_cap, uint32_t dense) {
    __shared__ uint32_t cnt[8][65];
    __shared__ uint32_t dpre[66];                       // exclusive prefix of this block's per-destination totals
    __shared__ unsigned long long stage[kSplitTile];    // keys grouped by destination
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    for (int i = threadIdx.x; i < 8 * 65; i += 256) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const uint64_t wbase = (uint64_t)blockIdx.x * kSplitTile + (uint64_t)warp * (kSplitTile / 8);
    for (int it = 0; it < kSplitTile / 8 / 32; ++it) {
        unsigned long long key;
        uint32_t           p;
        uint32_t           d = window_dest(prev, list, wbase + (uint64_t)it * 32 + lane, npos, world, key, dense, p);
        if (d > 64) d = 64;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        if ((int)lane == __ffs(peers) - 1) cnt[warp][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        uint32_t run = 0;
        for (int w = 0; w < 8; ++w) {
            uint32_t c          = cnt[w][threadIdx.x];
            cnt[w][threadIdx.x] = run;
            run += c;
        }
        dpre[threadIdx.x + 1] = run;  // totals for now
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        dpre[0]      = 0;
        for (uint32_t d = 0; d < 64; ++d) {
            uint32_t c = dpre[d + 1];
            dpre[d + 1] = acc + c;
            acc += c;
        }
    }
    __syncthreads();
    for (int it = 0; it < kSplitTile / 8 / 32; ++it) {
        const uint64_t     j = wbase + (uint64_t)it * 32 + lane;  // the item; rec_of_pos is indexed by item
        unsigned long long key = 0;
        uint32_t           p = 0;
        uint32_t           d = window_dest(prev, list, j, npos, world, key, dense, p);
        const bool         act = d < 64, is_dense = d == kDenseDest;
        if (d > 64) d = 64;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t base  = cnt[warp][d];
        __syncwarp();
        if ((int)lane == __ffs(peers) - 1) cnt[warp][d] = base + __popc(peers);
        __syncwarp();
        if (act) {
            const uint32_t in_blk = base + __popc(peers & ((1u << lane) - 1));  // rank inside this block's group for owner d
            const uint64_t dst    = hist_off[(uint64_t)d * nblocks + blockIdx.x] + in_blk;
            stage[dpre[d] + in_blk] = key;
            pos_of_rec[dst]         = p;
            if (peer_keys != nullptr)
                rec_of_pos[j] = (uint32_t)((uint64_t)d * slot_cap + (dst - hist_off[(uint64_t)d * nblocks]));  // where the reply will appear in my own reply slots
            else
                rec_of_pos[j] = (uint32_t)dst;
        } else if (j < npos) {
            rec_of_pos[j] = is_dense ? (kDenseRec | (uint32_t)key) : kNoRec;
        }
    }
    __syncthreads();
    const uint32_t total = dpre[64];
    for (uint32_t i = threadIdx.x; i < total; i += 256) {
        uint32_t d = 0;
        while (i >= dpre[d + 1]) ++d;
        const uint64_t dst = hist_off[(uint64_t)d * nblocks + blockIdx.x] + (i - dpre[d]);
        if (peer_keys != nullptr)
            peer_keys[d][(uint64_t)my_rank * slot_cap + (dst - hist_off[(uint64_t)d * nblocks])] = stage[i];
        else
            send_keys[dst] = stage[i];
    }
}

// ---- owner side: the occurrence filter and the counting kernel over a received key stream -------------------------------
__device__ __forceinline__ void stream_filter_locate(uint64_t h, uint64_t mask, uint64_t& word, uint32_t& shift) {
    uint64_t bucket = h & mask;
    word            = bucket >> 4;
    shift           = (uint32_t)(bucket & 15) * 2;
}

// slotted input (P2P mode, slot_cap != 0): the buffer is G slots of slot_cap keys, slot r holds slot_counts[r] keys of source r.  The owner's
// kernels enumerate the nrecv keys that are there (c = 0 .. nrecv), not the slots' capacity: SlotMap turns c into the buffer index
// src * slot_cap + k through the prefix sums of the counts (G <= 64, built once per block in shared memory).
struct SlotMap {
    unsigned long long pre[65];
    uint32_t           world;
    uint64_t           slot_cap;
    __device__ __forceinline__ void init(uint32_t G, uint64_t cap, const unsigned long long* __restrict__ slot_counts) {  // by the whole block; ends with a barrier
        if (threadIdx.x == 0) {
            world    = cap ? G : 0;
            slot_cap = cap;
            pre[0]   = 0;
            for (uint32_t r = 0; r < world; ++r) pre[r + 1] = pre[r] + slot_counts[r];
        }
        __syncthreads();
    }
    __device__ __forceinline__ uint64_t index_of(uint64_t c, uint32_t& src, uint64_t& k) const {
        if (world == 0) {
            src = 0;
            k   = c;
            return c;
        }
        uint32_t lo = 0, hi = world;  // pre[lo] <= c < pre[hi]
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (pre[mid] <= c) lo = mid;
            else hi = mid;
        }
        src = lo;
        k   = c - pre[lo];
        return (uint64_t)lo * slot_cap + k;
    }
};

__global__ void __launch_bounds__(256) stream_filter_kernel(const unsigned long long* __restrict__ keys, uint64_t n, uint32_t* __restrict__ filter, uint64_t nbuckets_mask,
                                                            DeviceStats* __restrict__ st, uint64_t slot_cap, const unsigned long long* __restrict__ slot_counts, uint32_t world) {
    __shared__ uint64_t scratch[8];
    __shared__ SlotMap  map;
    map.init(world, slot_cap, slot_counts);
    uint32_t twice = 0;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t       src;
        uint64_t       k;
        const uint64_t i = map.index_of(c, src, k);
        uint64_t word;
        uint32_t shift;
        stream_filter_locate(table_hash_u64(__ldcs(keys + i)), nbuckets_mask, word, shift);
        uint32_t bits = (__ldcg(filter + word) >> shift) & 3u;
        if (bits == 3u) continue;
        if ((bits & 1u) == 0) {
            uint32_t old = atomicOr(filter + word, 1u << shift);
            if (((old >> shift) & 1u) == 0) continue;
            bits = (old >> shift) & 3u;
        }
        if ((bits & 2u) == 0) {
            uint32_t old = atomicOr(filter + word, 2u << shift);
            twice += ((old >> shift) & 2u) == 0;
        }
    }
    uint64_t tw = block_reduce_sum(twice, scratch);
    if (threadIdx.x == 0 && tw) atomicAdd(&st->found, (unsigned long long)tw);
}

__device__ __forceinline__ void sk_cas128(void* addr, unsigned long long new0, unsigned long long new1, unsigned long long& old0, unsigned long long& old1) {
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

// rid[i] = slot + 1 of received window i (0: the filter proved it to be the only window of its key).  slot.pos = i of the claimer.
//
// Hot keys: an owner receives every occurrence of its keys from all G ranks, so the most frequent n-grams put millions of
// REDs on one L2 address (they serialise: the owner of "the the" became the straggler of the level).  Each block therefore
// keeps a small never-evicting cache in shared memory: key -> {table slot, pending count}.  The first window of a key that
// finds its cache line empty installs it after its normal upsert; later windows of that key in the block only bump the
// shared-memory counter, which is added to the table once when the block retires.  Lines never change owner, so counts stay exact.
constexpr uint32_t kHotLines    = 1024;
constexpr unsigned long long kHotBusy = ~0ull;

__global__ void __launch_bounds__(256) stream_count_kernel(const unsigned long long* __restrict__ keys, uint64_t n, NgramSlot* __restrict__ table, uint64_t cap,
                                                           const uint32_t* __restrict__ filter, uint64_t nbuckets_mask, uint32_t* __restrict__ rid, DeviceStats* __restrict__ st,
                                                           uint64_t slot_cap, const unsigned long long* __restrict__ slot_counts, uint32_t world, bool onebit) {
    __shared__ uint64_t scratch[8];
    __shared__ SlotMap  map;
    __shared__ unsigned long long hot_key[kHotLines];
    __shared__ uint32_t hot_slot[kHotLines];
    __shared__ uint32_t hot_pending[kHotLines];
    for (uint32_t i = threadIdx.x; i < kHotLines; i += blockDim.x) {
        hot_key[i]     = 0;
        hot_pending[i] = 0;
    }
    __syncthreads();
    uint32_t       singles = 0;
    bool           full    = false;
    const uint64_t limit   = cap < 8192 ? cap : 8192;
    map.init(world, slot_cap, slot_counts);
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t       src;
        uint64_t       kk;
        const uint64_t i = map.index_of(c, src, kk);  // rid[] and the claimer's index stay in buffer coordinates
        const unsigned long long key = __ldcs(keys + i);
        const uint64_t           h   = table_hash_u64(key);
        uint32_t                 out = 0;
        bool                     go  = true;
        if (filter != nullptr) {
            uint64_t word;
            uint32_t shift;
            if (onebit) {  // the packed "hit twice" bits (launch_filter_to_bitmap)
                const uint64_t bucket = h & nbuckets_mask;
                go = ((__ldg(filter + (bucket >> 5)) >> (bucket & 31)) & 1u) != 0;
            } else {
                stream_filter_locate(h, nbuckets_mask, word, shift);
                go = ((__ldg(filter + word) >> shift) & 2u) != 0;
            }
            singles += !go;
        }
        if (go) {
            const uint32_t line = (uint32_t)(h >> 20) & (kHotLines - 1);
            unsigned long long cached = *(volatile unsigned long long*)&hot_key[line];
            if (cached == key) {  // the slot was published before the key (see below)
                atomicAdd(&hot_pending[line], 1u);
                out = *(volatile uint32_t*)&hot_slot[line];
            } else {
                uint64_t slot = fast_range(h, cap);
                uint64_t step = 0;
                for (; step < limit; ++step) {
                    NgramSlot*         s   = table + slot;
                    unsigned long long cur = __ldcg(&s->key);
                    if (cur == 0) {
                        unsigned long long o0, o1;
                        sk_cas128(s, key, 1ull | ((unsigned long long)(uint32_t)i << 32), o0, o1);
                        if (o0 == 0) {
                            out = (uint32_t)slot + 1;
                            break;
                        }
                        cur = o0;
                    }
                    if (cur == key) {
                        atomicAdd(&s->count, 1u);
                        out = (uint32_t)slot + 1;
                        break;
                    }
                    slot = slot + 1 == cap ? 0 : slot + 1;
                }
                if (out == 0) {
                    full = true;
                } else if (cached == 0 && atomicCAS(&hot_key[line], 0ull, kHotBusy) == 0ull) {
                    hot_slot[line] = out;          // publish the slot ...
                    __threadfence_block();
                    *(volatile unsigned long long*)&hot_key[line] = key;  // ... then the key that makes it visible
                }
            }
        }
        __stcs(rid + i, out);
    }
    __syncthreads();
    for (uint32_t l = threadIdx.x; l < kHotLines; l += blockDim.x) {
        uint32_t c = hot_pending[l];
        if (c) atomicAdd(&table[hot_slot[l] - 1].count, c);
    }
    uint64_t sg = block_reduce_sum(singles, scratch);
    if (threadIdx.x == 0 && sg) atomicAdd(&st->singletons, (unsigned long long)sg);
    if (full) atomicOr(&st->errflags, kErrTableFull);
}

// P2P mode (peer_reply != NULL): the global id is stored straight into the sender's reply slot for this owner over NVLink
__global__ void __launch_bounds__(256) owner_reply_kernel(uint32_t* __restrict__ rid, uint64_t n, const uint32_t* __restrict__ bitmap, uint32_t world, uint32_t rank,
                                                          uint32_t* const* __restrict__ peer_reply, uint64_t slot_cap, const unsigned long long* __restrict__ slot_counts,
                                                          uint32_t id_off) {
    __shared__ SlotMap map;
    map.init(world, peer_reply != nullptr ? slot_cap : 0, slot_counts);
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (peer_reply != nullptr) {
        uint32_t src;
        uint64_t k;
        i            = map.index_of(i, src, k);
        uint32_t s   = rid[i];
        uint32_t gid = (s != 0 && ((__ldg(bitmap + ((s - 1) >> 5)) >> ((s - 1) & 31)) & 1u)) ? id_off + (s - 1) * world + rank + 1 : 0u;
        peer_reply[src][(uint64_t)rank * slot_cap + k] = gid;
        return;
    }
    uint32_t s = rid[i];
    if (s == 0) return;
    rid[i] = ((__ldg(bitmap + ((s - 1) >> 5)) >> ((s - 1) & 31)) & 1u) ? id_off + (s - 1) * world + rank + 1 : 0u;
}

// P2P survivors: survivor (receive index i, count) -> record (i % slot_cap, count) in the source's survivor slot for this owner;
// per-source cursors (one atomic per source per 2048-survivor tile); the final cursor values are the counts the host publishes.
__global__ void __launch_bounds__(256) owner_survivors_p2p_kernel(const uint32_t* __restrict__ sv_idx, const uint32_t* __restrict__ sv_count, uint64_t n, uint32_t world, uint32_t rank,
                                                                  uint64_t slot_cap, uint64_t surv_cap, unsigned long long* __restrict__ cursors, uint2* const* __restrict__ peer_surv,
                                                                  DeviceStats* __restrict__ st) {
    __shared__ uint32_t tile_cnt[64];
    __shared__ unsigned long long tile_base[64];
    const uint64_t ntiles = (n + 2047) / 2048;
    bool overflow = false;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (threadIdx.x < 64) tile_cnt[threadIdx.x] = 0;
        __syncthreads();
        uint32_t k_in[8], cnt[8], src[8], rk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint64_t i = tile * 2048 + (uint64_t)k * 256 + threadIdx.x;
            src[k]     = 0xFFFFFFFFu;
            if (i < n) {
                uint32_t v = sv_idx[i];
                src[k]     = (uint32_t)(v / slot_cap);
                k_in[k]    = (uint32_t)(v - (uint64_t)src[k] * slot_cap);
                cnt[k]     = sv_count[i];
                rk[k]      = atomicAdd(&tile_cnt[src[k]], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < world) {
            uint32_t c = tile_cnt[threadIdx.x];
            tile_base[threadIdx.x] = c ? atomicAdd(&cursors[threadIdx.x], (unsigned long long)c) : 0ull;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (src[k] == 0xFFFFFFFFu) continue;
            uint64_t o = tile_base[src[k]] + rk[k];
            if (o < surv_cap)
                peer_surv[src[k]][(uint64_t)rank * surv_cap + o] = make_uint2(k_in[k], cnt[k]);
            else
                overflow = true;
        }
        __syncthreads();
    }
    if (overflow) atomicOr(&st->errflags, kErrTableFull);
}

// publish up to 8 u64 values per peer: dst[d][offset + j] = vals[d * nvals + j]   (headers of the symmetric buffers)
__global__ void p2p_publish_kernel(unsigned long long* const* __restrict__ peer_hdr, uint32_t world, uint64_t offset, const unsigned long long* __restrict__ vals, uint32_t nvals) {
    uint32_t d = threadIdx.x / 8, j = threadIdx.x % 8;
    if (d < world && j < nvals) peer_hdr[d][offset + j] = vals[d * nvals + j];
}

// survivors (claimer's receive index, global count) -> per source rank: (index inside that source's group, count).
// src_base[r] = first receive index of source r (G+1 entries).  One atomic per destination per 2048-survivor tile.
__global__ void __launch_bounds__(256) owner_survivors_kernel(const uint32_t* __restrict__ sv_idx, const uint32_t* __restrict__ sv_count, uint64_t n, uint32_t world,
                                                              const unsigned long long* __restrict__ src_base, const unsigned long long* __restrict__ out_base,
                                                              unsigned long long* __restrict__ cursors, uint2* __restrict__ out) {
    __shared__ uint32_t tile_cnt[64];
    __shared__ unsigned long long tile_base[64];
    __shared__ unsigned long long sbase[65];
    if (threadIdx.x <= world) sbase[threadIdx.x] = src_base[threadIdx.x];
    const uint64_t ntiles = (n + 2047) / 2048;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (threadIdx.x < 64) tile_cnt[threadIdx.x] = 0;
        __syncthreads();
        uint32_t idx[8], cnt[8], src[8], rk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint64_t i = tile * 2048 + (uint64_t)k * 256 + threadIdx.x;
            src[k]     = 0xFFFFFFFFu;
            if (i < n) {
                idx[k] = sv_idx[i];
                cnt[k] = sv_count[i];
                uint32_t r = 0;
                while (r + 1 < world && (unsigned long long)idx[k] >= sbase[r + 1]) ++r;
                src[k] = r;
                rk[k]  = atomicAdd(&tile_cnt[r], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < world) {
            uint32_t c = tile_cnt[threadIdx.x];
            tile_base[threadIdx.x] = c ? out_base[threadIdx.x] + atomicAdd(&cursors[threadIdx.x], (unsigned long long)c) : 0ull;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (src[k] != 0xFFFFFFFFu) out[tile_base[src[k]] + rk[k]] = make_uint2((uint32_t)(idx[k] - sbase[src[k]]), cnt[k]);
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) owner_survivor_counts_kernel(const uint32_t* __restrict__ sv_idx, uint64_t n, uint32_t world, const unsigned long long* __restrict__ src_base,
                                                                    unsigned long long* __restrict__ counts) {
    __shared__ uint32_t h[64];
    __shared__ unsigned long long sbase[65];
    if (threadIdx.x < 64) h[threadIdx.x] = 0;
    if (threadIdx.x <= world) sbase[threadIdx.x] = src_base[threadIdx.x];
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t v = sv_idx[i], r = 0;
        while (r + 1 < world && (unsigned long long)v >= sbase[r + 1]) ++r;
        atomicAdd(&h[r], 1u);
    }
    __syncthreads();
    if (threadIdx.x < world && h[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

// ---- sender side --------------------------------------------------------------------------------------------------------
// id[p] = reply[rec_of_pos[p]]: inside every destination group the records are in corpus order, so this reads G ascending streams
// list != NULL: item i is position list[i] and cur was zeroed by the host; only the positions with a surviving n-gram are written
__global__ void __launch_bounds__(256) sender_relabel_kernel(const uint32_t* __restrict__ rec_of_pos, const uint32_t* __restrict__ reply, uint64_t npos /* items */,
                                                             uint32_t* __restrict__ cur, DeviceStats* __restrict__ st, const uint32_t* __restrict__ dense_global, uint32_t threshold,
                                                             const uint32_t* __restrict__ list) {
    __shared__ uint64_t scratch[8];
    uint32_t valid = 0;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npos; p += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t j  = __ldcs(rec_of_pos + p);
        uint32_t id;
        if (j == kNoRec) id = 0u;
        else if (j & kDenseRec) id = __ldg(dense_global + (j & ~kDenseRec)) >= threshold ? (j & ~kDenseRec) + 1u : 0u;  // the same verdict on every rank
        else id = __ldg(reply + j);
        if (list == nullptr) __stcs(cur + p, id);
        else if (id != 0) cur[__ldg(list + p)] = id;
        valid += id != 0;
    }
    uint64_t v = block_reduce_sum(valid, scratch);
    if (threadIdx.x == 0 && v) atomicAdd(&st->kept_occ, (unsigned long long)v);
}
// prune(MINTOKENS, 2) over this rank's share of the GLOBAL dense square (cells = rank mod world): statistics, and the survivors as
// (position of the class pair in the spare room behind the tokens, global count) -- the rank exports them like any other bigram.
__global__ void __launch_bounds__(256) dense_share_kernel(const uint32_t* __restrict__ dense_global, uint32_t dense, uint32_t world, uint32_t rank, uint32_t threshold,
                                                          uint32_t* __restrict__ sv_pos, uint32_t* __restrict__ sv_cnt, uint32_t* __restrict__ tok_ext, uint32_t ext_pos0,
                                                          DeviceStats* __restrict__ st) {
    __shared__ uint64_t scratch[8];
    const uint64_t cells = (uint64_t)dense * dense;
    uint64_t found = 0, kept = 0, occ = 0;
    const uint64_t mine = (cells + world - 1 - rank) / world;  // cells rank, rank + world, ...
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < mine; base += (uint64_t)gridDim.x * blockDim.x) {  // (whole warps iterate together: the cursor is warp-aggregated)
        const uint64_t i    = base + threadIdx.x;
        const uint64_t cell = i * world + rank;
        const uint32_t c    = i < mine ? __ldg(dense_global + cell) : 0u;
        const bool     keep = c != 0 && c >= threshold;
        found += c != 0;
        const uint64_t o = warp_aggregated_inc(&st->cursor, keep);
        if (keep) {
            ++kept;
            occ += c;
            sv_pos[o]          = ext_pos0 + 2u * (uint32_t)i;
            sv_cnt[o]          = c;
            tok_ext[2 * i]     = (uint32_t)(cell / dense);
            tok_ext[2 * i + 1] = (uint32_t)(cell % dense);
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

// received survivor records of owner group g: (index inside my send group to g, global count) -> (position, count)
__global__ void __launch_bounds__(256) sender_survivors_kernel(const uint2* __restrict__ recs, uint64_t n, const uint32_t* __restrict__ pos_of_rec, uint64_t send_base,
                                                               uint32_t* __restrict__ sv_pos, uint32_t* __restrict__ sv_count) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint2 r     = recs[i];
    sv_pos[i]   = pos_of_rec[send_base + r.x];
    sv_count[i] = r.y;
}

// ---- launchers ------------------------------------------------------------------------------------------------------------
int launch_split_count(cudaStream_t s, const uint32_t* prev, uint64_t npos, uint32_t world, uint32_t* hist /* world x nblocks */, uint32_t dense, uint32_t* dense_cnt,
                       const uint32_t* list) {
    uint32_t nblocks = sk_div_up(npos, kSplitTile);
    if (!nblocks) return 0;
    split_count_kernel<<<nblocks, 256, 0, s>>>(prev, list, npos, world, nblocks, hist, dense, dense_cnt);
    return 1;
}
int launch_split_write(cudaStream_t s, const uint32_t* prev, uint64_t npos, uint32_t world, const uint64_t* hist_off, void* send_keys, uint32_t* pos_of_rec, uint32_t* rec_of_pos,
                       void* const* peer_keys, uint32_t my_rank, uint64_t slot_cap, uint32_t dense, const uint32_t* list) {
    uint32_t nblocks = sk_div_up(npos, kSplitTile);
    if (!nblocks) return 0;
    split_write_kernel<<<nblocks, 256, 0, s>>>(prev, list, npos, world, nblocks, hist_off, (unsigned long long*)send_keys, pos_of_rec, rec_of_pos, (unsigned long long* const*)peer_keys,
                                                my_rank, slot_cap, dense);
    return 1;
}
int launch_stream_filter(cudaStream_t s, const void* keys, uint64_t n, uint32_t* filter, uint64_t nbuckets, DeviceStats* st, int sms, uint64_t slot_cap,
                         const unsigned long long* slot_counts, uint32_t world) {
    if (!n) return 0;
    unsigned grid = (unsigned)sk_min(sk_div_up(n, 256), (uint64_t)sms * 32);
    stream_filter_kernel<<<grid, 256, 0, s>>>((const unsigned long long*)keys, n, filter, nbuckets - 1, st, slot_cap, slot_counts, world);
    return 1;
}
int launch_stream_count(cudaStream_t s, const void* keys, uint64_t n, NgramSlot* table, uint64_t cap, const uint32_t* filter, uint64_t nbuckets, uint32_t* rid, DeviceStats* st, int sms,
                        uint64_t slot_cap, const unsigned long long* slot_counts, uint32_t world, bool onebit) {
    if (!n) return 0;
    unsigned grid = (unsigned)sk_min(sk_div_up(n, 256), (uint64_t)sms * 32);
    stream_count_kernel<<<grid, 256, 0, s>>>((const unsigned long long*)keys, n, table, cap, filter, nbuckets ? nbuckets - 1 : 0, rid, st, slot_cap, slot_counts, world, onebit);
    return 1;
}
int launch_owner_reply(cudaStream_t s, uint32_t* rid, uint64_t n, const uint32_t* bitmap, uint32_t world, uint32_t rank, void* const* peer_reply, uint64_t slot_cap,
                       const unsigned long long* slot_counts, uint32_t id_off) {
    if (!n) return 0;
    owner_reply_kernel<<<sk_div_up(n, 256), 256, 0, s>>>(rid, n, bitmap, world, rank, (uint32_t* const*)peer_reply, slot_cap, slot_counts, id_off);
    return 1;
}
int launch_owner_survivors_p2p(cudaStream_t s, const uint32_t* sv_idx, const uint32_t* sv_count, uint64_t n, uint32_t world, uint32_t rank, uint64_t slot_cap, uint64_t surv_cap,
                               unsigned long long* cursors, void* const* peer_surv, DeviceStats* st, int sms) {
    if (!n) return 0;
    unsigned grid = (unsigned)sk_min(sk_div_up(n, 2048), (uint64_t)sms * 4);
    owner_survivors_p2p_kernel<<<grid, 256, 0, s>>>(sv_idx, sv_count, n, world, rank, slot_cap, surv_cap, cursors, (uint2* const*)peer_surv, st);
    return 1;
}
int launch_p2p_publish(cudaStream_t s, void* const* peer_hdr, uint32_t world, uint64_t offset, const unsigned long long* vals, uint32_t nvals) {
    p2p_publish_kernel<<<1, 512, 0, s>>>((unsigned long long* const*)peer_hdr, world, offset, vals, nvals);
    return 1;
}
int launch_owner_survivor_counts(cudaStream_t s, const uint32_t* sv_idx, uint64_t n, uint32_t world, const unsigned long long* src_base, unsigned long long* counts, int sms) {
    if (!n) return 0;
    unsigned grid = (unsigned)sk_min(sk_div_up(n, 256), (uint64_t)sms * 8);
    owner_survivor_counts_kernel<<<grid, 256, 0, s>>>(sv_idx, n, world, src_base, counts);
    return 1;
}
int launch_owner_survivors(cudaStream_t s, const uint32_t* sv_idx, const uint32_t* sv_count, uint64_t n, uint32_t world, const unsigned long long* src_base,
                           const unsigned long long* out_base, unsigned long long* cursors, void* out, int sms) {
    if (!n) return 0;
    unsigned grid = (unsigned)sk_min(sk_div_up(n, 2048), (uint64_t)sms * 4);
    owner_survivors_kernel<<<grid, 256, 0, s>>>(sv_idx, sv_count, n, world, src_base, out_base, cursors, (uint2*)out);
    return 1;
}
int launch_sender_relabel(cudaStream_t s, const uint32_t* rec_of_pos, const uint32_t* reply, uint64_t npos, uint32_t* cur, DeviceStats* st, int sms, const uint32_t* dense_global,
                          uint32_t threshold, const uint32_t* list) {
    unsigned grid = (unsigned)sk_min(sk_div_up(npos, 256), (uint64_t)sms * 16);
    sender_relabel_kernel<<<grid ? grid : 1, 256, 0, s>>>(rec_of_pos, reply, npos, cur, st, dense_global, threshold, list);
    return 1;
}
int launch_dense_share(cudaStream_t s, const uint32_t* dense_global, uint32_t dense, uint32_t world, uint32_t rank, uint32_t threshold, uint32_t* sv_pos, uint32_t* sv_cnt,
                       uint32_t* tok_ext, uint32_t ext_pos0, DeviceStats* st, int sms) {
    const uint64_t mine = ((uint64_t)dense * dense + world - 1) / world;
    unsigned       grid = (unsigned)sk_min(sk_div_up(mine, 256), (uint64_t)sms * 8);
    dense_share_kernel<<<grid ? grid : 1, 256, 0, s>>>(dense_global, dense, world, rank, threshold, sv_pos, sv_cnt, tok_ext, ext_pos0, st);
    return 1;
}
int launch_sender_survivors(cudaStream_t s, const void* recs, uint64_t n, const uint32_t* pos_of_rec, uint64_t send_base, uint32_t* sv_pos, uint32_t* sv_count) {
    if (!n) return 0;
    sender_survivors_kernel<<<sk_div_up(n, 256), 256, 0, s>>>((const uint2*)recs, n, pos_of_rec, send_base, sv_pos, sv_count);
    return 1;
}

}  // namespace colibri

// =============================================================================================================================
// Skipgrams on the multi-GPU path (exhaustive mode, reference include/patternmodel.h:1163-1171).  A skipgram's key is
// (mask, global ids of its <= 3 contiguous non-gap runs) -- global because every level's ids are global here -- so the same
// "ship it to its owner" scheme applies, without the way back: skipgrams feed no later level, only the survivors return.
namespace colibri {

__device__ __forceinline__ uint32_t owner_of_key128(unsigned long long k0, unsigned long long k1, uint32_t world) {
    unsigned long long z = (k0 ^ ((k1 << 29) | (k1 >> 35))) * 0x9E3779B97F4A7C15ull;
    z ^= z >> 32;
    z *= 0xD6E8FEB86659FD93ull;
    return (uint32_t)__umul64hi(z, (unsigned long long)world);
}

// item t = (position p, mask m); returns false if the window is not valid at level n
__device__ __forceinline__ bool skip_item_key(const uint32_t* const* __restrict__ ids, int n, const SkipMask* __restrict__ masks, int nmasks, uint64_t t, uint64_t& p,
                                              unsigned long long& k0, unsigned long long& k1) {
    p           = t / nmasks;
    const int m = (int)(t - p * nmasks);
    const uint32_t* prev = ids[n - 1];
    if (__ldg(prev + p) == 0 || __ldg(prev + p + 1) == 0) return false;
    const SkipMask* sm = masks + m;
    const uint32_t  np = __ldg(&sm->nparts);
    k0 = ((unsigned long long)__ldg(&sm->mask) << 32) | __ldg(ids[__ldg(&sm->len[0])] + p + __ldg(&sm->start[0]));
    k1 = (unsigned long long)__ldg(ids[__ldg(&sm->len[1])] + p + __ldg(&sm->start[1])) << 32;
    if (np > 2) k1 |= __ldg(ids[__ldg(&sm->len[2])] + p + __ldg(&sm->start[2]));
    return true;
}

__global__ void __launch_bounds__(256) skip_split_count_kernel(const uint32_t* const* __restrict__ ids, int n, const SkipMask* __restrict__ masks, int nmasks, uint64_t npos,
                                                               uint32_t world, unsigned long long* __restrict__ dest_counts) {
    __shared__ uint32_t h[64];
    if (threadIdx.x < 64) h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t total = npos * (uint64_t)nmasks;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t           p;
        unsigned long long k0, k1;
        if (skip_item_key(ids, n, masks, nmasks, t, p, k0, k1)) atomicAdd(&h[owner_of_key128(k0, k1, world)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < world && h[threadIdx.x]) atomicAdd(&dest_counts[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

// 16-byte keys grouped by owner (order inside a group is free: nothing comes back per record), pos_of_rec[j] = window position
__global__ void __launch_bounds__(256) skip_split_write_kernel(const uint32_t* const* __restrict__ ids, int n, const SkipMask* __restrict__ masks, int nmasks, uint64_t npos,
                                                               uint32_t world, const unsigned long long* __restrict__ dest_base, unsigned long long* __restrict__ cursors,
                                                               ulonglong2* __restrict__ send, uint32_t* __restrict__ pos_of_rec) {
    __shared__ uint32_t tile_cnt[64];
    __shared__ unsigned long long tile_base[64];
    const uint64_t total  = npos * (uint64_t)nmasks;
    const uint64_t ntiles = (total + 2047) / 2048;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (threadIdx.x < 64) tile_cnt[threadIdx.x] = 0;
        __syncthreads();
        unsigned long long k0[8], k1[8];
        uint32_t           dest[8], rk[8], pos[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint64_t t = tile * 2048 + (uint64_t)k * 256 + threadIdx.x, p = 0;
            dest[k]    = 0xFFFFFFFFu;
            if (t < total && skip_item_key(ids, n, masks, nmasks, t, p, k0[k], k1[k])) {
                dest[k] = owner_of_key128(k0[k], k1[k], world);
                pos[k]  = (uint32_t)p;
                rk[k]   = atomicAdd(&tile_cnt[dest[k]], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < world) {
            uint32_t c = tile_cnt[threadIdx.x];
            tile_base[threadIdx.x] = c ? dest_base[threadIdx.x] + atomicAdd(&cursors[threadIdx.x], (unsigned long long)c) : 0ull;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (dest[k] == 0xFFFFFFFFu) continue;
            uint64_t j    = tile_base[dest[k]] + rk[k];
            send[j]       = make_ulonglong2(k0[k], k1[k]);
            pos_of_rec[j] = pos[k];
        }
        __syncthreads();
    }
}

// owner: count the received skipgram keys; slot.pos = receive index of the claimer
__global__ void __launch_bounds__(256) skip_stream_count_kernel(const ulonglong2* __restrict__ recv, uint64_t n, SkipSlot* __restrict__ table, uint64_t cap, DeviceStats* __restrict__ st) {
    bool           full  = false;
    const uint64_t limit = cap < 8192 ? cap : 8192;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const ulonglong2 key  = __ldcs(recv + i);
        uint64_t         slot = fast_range(table_hash_u128(key.x, key.y), cap);
        uint64_t         step = 0;
        for (; step < limit; ++step) {
            SkipSlot*          s  = table + slot;
            ulonglong2         kv = __ldcg(reinterpret_cast<const ulonglong2*>(s));
            unsigned long long c0 = kv.x, c1 = kv.y;
            if (c0 == 0 || c1 == 0) {
                sk_cas128(s, key.x, key.y, c0, c1);  // expects the all-zero key; the result is authoritative
                if (c0 == 0 && c1 == 0) {
                    s->pos = (uint32_t)i;
                    c0     = key.x;
                    c1     = key.y;
                }
            }
            if (c0 == key.x && c1 == key.y) {
                atomicAdd(&s->count, 1u);
                break;
            }
            slot = slot + 1 == cap ? 0 : slot + 1;
        }
        if (step == limit) full = true;
    }
    if (full) atomicOr(&st->errflags, kErrTableFull);
}

// survivors (receive index, count, mask) -> per source: 16-byte records {index inside the source's group, count, mask, 0}
__global__ void __launch_bounds__(256) skip_owner_survivors_kernel(const uint32_t* __restrict__ sv_idx, const uint32_t* __restrict__ sv_count, const uint32_t* __restrict__ sv_mask,
                                                                   uint64_t n, uint32_t world, const unsigned long long* __restrict__ src_base,
                                                                   const unsigned long long* __restrict__ out_base, unsigned long long* __restrict__ cursors, uint4* __restrict__ out) {
    __shared__ unsigned long long sbase[65];
    if (threadIdx.x <= world) sbase[threadIdx.x] = src_base[threadIdx.x];
    __syncthreads();
    const uint64_t rounded = (n + 31) / 32 * 32;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t src = 0xFFFFFFFFu, idx = 0;
        if (i < n) {
            idx = sv_idx[i];
            src = 0;
            while (src + 1 < world && (unsigned long long)idx >= sbase[src + 1]) ++src;
        }
        uint32_t peers = __match_any_sync(0xffffffffu, src);
        if (i >= n) continue;
        int      leader = __ffs(peers) - 1;
        uint64_t base   = 0;
        if ((int)lane_id() == leader) base = atomicAdd(&cursors[src], (unsigned long long)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        out[out_base[src] + base + __popc(peers & ((1u << lane_id()) - 1))] = make_uint4((uint32_t)(idx - sbase[src]), sv_count[i], sv_mask[i], 0u);
    }
}
__global__ void __launch_bounds__(256) skip_sender_survivors_kernel(const uint4* __restrict__ recs, uint64_t n, const uint32_t* __restrict__ pos_of_rec, uint64_t send_base,
                                                                    uint32_t* __restrict__ sv_pos, uint32_t* __restrict__ sv_count, uint32_t* __restrict__ sv_mask) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 r     = recs[i];
    sv_pos[i]   = pos_of_rec[send_base + r.x];
    sv_count[i] = r.y;
    sv_mask[i]  = r.z;
}

int launch_skip_split_count(cudaStream_t s, const uint32_t* const* ids, int n, const SkipMask* masks, int nmasks, uint64_t npos, uint32_t world, unsigned long long* dest_counts, int sms) {
    unsigned grid = (unsigned)sk_min(sk_div_up(npos * nmasks, 256), (uint64_t)sms * 16);
    skip_split_count_kernel<<<grid ? grid : 1, 256, 0, s>>>(ids, n, masks, nmasks, npos, world, dest_counts);
    return 1;
}
int launch_skip_split_write(cudaStream_t s, const uint32_t* const* ids, int n, const SkipMask* masks, int nmasks, uint64_t npos, uint32_t world, const unsigned long long* dest_base,
                            unsigned long long* cursors, void* send, uint32_t* pos_of_rec, int sms) {
    unsigned grid = (unsigned)sk_min(sk_div_up(npos * nmasks, 2048), (uint64_t)sms * 4);
    skip_split_write_kernel<<<grid ? grid : 1, 256, 0, s>>>(ids, n, masks, nmasks, npos, world, dest_base, cursors, (ulonglong2*)send, pos_of_rec);
    return 1;
}
int launch_skip_stream_count(cudaStream_t s, const void* recv, uint64_t n, SkipSlot* table, uint64_t cap, DeviceStats* st, int sms) {
    if (!n) return 0;
    unsigned grid = (unsigned)sk_min(sk_div_up(n, 256), (uint64_t)sms * 32);
    skip_stream_count_kernel<<<grid, 256, 0, s>>>((const ulonglong2*)recv, n, table, cap, st);
    return 1;
}
int launch_skip_owner_survivors(cudaStream_t s, const uint32_t* sv_idx, const uint32_t* sv_count, const uint32_t* sv_mask, uint64_t n, uint32_t world, const unsigned long long* src_base,
                                const unsigned long long* out_base, unsigned long long* cursors, void* out, int sms) {
    if (!n) return 0;
    unsigned grid = (unsigned)sk_min(sk_div_up(n, 256), (uint64_t)sms * 8);
    skip_owner_survivors_kernel<<<grid, 256, 0, s>>>(sv_idx, sv_count, sv_mask, n, world, src_base, out_base, cursors, (uint4*)out);
    return 1;
}
int launch_skip_sender_survivors(cudaStream_t s, const void* recs, uint64_t n, const uint32_t* pos_of_rec, uint64_t send_base, uint32_t* sv_pos, uint32_t* sv_count, uint32_t* sv_mask) {
    if (!n) return 0;
    skip_sender_survivors_kernel<<<sk_div_up(n, 256), 256, 0, s>>>((const uint4*)recs, n, pos_of_rec, send_base, sv_pos, sv_count, sv_mask);
    return 1;
}

}  // namespace colibri
