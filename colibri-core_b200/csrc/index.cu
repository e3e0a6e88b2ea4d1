// index.cu -- the forward index of indexed pattern models.
//
// Reference: IndexedPatternModel::add pushes an IndexReference (sentence 1-based, token 0-based; include/datatypes.h:33-89)
// onto the pattern's std::vector for every occurrence (include/patternmodel.h:2789-2800, datatypes.h:283-289) and posttrain()
// sorts every vector (:2699-2705).  On the device the occurrences of a level are known after its prune/relabel step: a
// position p belongs to survivor `idx` iff id[p] maps to it.  The (idx, p) pairs are written in corpus order and sorted by
// idx with a STABLE least-significant-digit radix sort, so inside each pattern the positions stay ascending -- which is the
// reference's sorted order -- without a per-pattern sort and without atomics.  Finally p -> (sentence, token).
#include "device_utils.cuh"
#include "kernels.h"

namespace colibri {

static inline unsigned idx_div_up(uint64_t a, uint64_t b) {
    return (unsigned)((a + b - 1) / b);
}

// ---------------------------------------------------------------------------------------------
// sentence bookkeeping: delim[p] = 1 where tok[p] == 0; sent_before[p] = delimiters in tok[0..p) (exclusive scan, done by the
// caller with launch_exclusive_scan_u32_u64); sent_start[k] = first position of sentence k (0-based)
__global__ void __launch_bounds__(256) delim_flags_kernel(const uint32_t* __restrict__ tok, uint64_t npos, uint32_t* __restrict__ flags) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npos) flags[i] = tok[i] == 0 ? 1u : 0u;
}
__global__ void __launch_bounds__(256) sent_start_kernel(const uint32_t* __restrict__ tok, const uint64_t* __restrict__ sent_before, uint64_t npos, uint32_t* __restrict__ sent_start) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) sent_start[0] = 0;
    if (i < npos && tok[i] == 0) sent_start[sent_before[i] + 1] = (uint32_t)i + 1;
}

// ---------------------------------------------------------------------------------------------
// ordered compaction of the occurrences of one level: for every position whose id maps to a survivor, the pair (idx, p),
// in corpus order.  map[] is indexed by id-1 (table slot) for n >= 2 and by class for n == 1; map value = survivor index + 1.
constexpr int kPairTile = 2048;  // positions per block

__device__ __forceinline__ uint32_t survivor_of(const uint32_t* __restrict__ ids, const uint32_t* __restrict__ map, uint64_t p, bool by_class) {
    uint32_t id = ids[p];
    if (id == 0) return 0;
    return __ldg(map + (by_class ? id : id - 1));
}

__global__ void __launch_bounds__(256) pair_count_kernel(const uint32_t* __restrict__ ids, const uint32_t* __restrict__ map, uint64_t npos, bool by_class, uint32_t* __restrict__ blk_counts) {
    __shared__ uint64_t scratch[8];
    uint64_t base = (uint64_t)blockIdx.x * kPairTile;
    uint32_t c    = 0;
#pragma unroll
    for (int k = 0; k < kPairTile / 256; ++k) {
        uint64_t p = base + (uint64_t)k * 256 + threadIdx.x;
        if (p < npos) c += survivor_of(ids, map, p, by_class) != 0;
    }
    uint64_t tot = block_reduce_sum(c, scratch);
    if (threadIdx.x == 0) blk_counts[blockIdx.x] = (uint32_t)tot;
}

__global__ void __launch_bounds__(256) pair_write_kernel(const uint32_t* __restrict__ ids, const uint32_t* __restrict__ map, uint64_t npos, bool by_class,
                                                         const uint64_t* __restrict__ blk_off, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                         const uint32_t* __restrict__ pos_lookup, uint32_t pos_div) {
    __shared__ uint32_t warp_cnt[8];
    uint64_t base = (uint64_t)blockIdx.x * kPairTile;
    uint64_t out  = blk_off[blockIdx.x];
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    for (int k = 0; k < kPairTile / 256; ++k) {  // 256 consecutive positions per round keep the corpus order
        uint64_t p   = base + (uint64_t)k * 256 + threadIdx.x;
        uint32_t idx = p < npos ? survivor_of(ids, map, p, by_class) : 0;
        uint32_t m   = __ballot_sync(0xffffffffu, idx != 0);
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (uint32_t w = 0; w < 8; ++w) {
            uint32_t c = warp_cnt[w];
            if (w < warp) before += c;
            total += c;
        }
        if (idx != 0) {
            uint64_t dst = out + before + __popc(m & ((1u << lane) - 1));
            keys[dst]    = idx - 1;
            vals[dst]    = pos_lookup != nullptr ? __ldg(pos_lookup + p / pos_div) : (uint32_t)p;
        }
        out += total;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort, 8 bits per pass, keys u32 + values u32.  A block owns a tile of 4096 elements, a warp a contiguous
// 512-element slice of it, processed 32 at a time in order; ranks inside a warp come from __match_any_sync.
constexpr int kSortTile = 4096;

__global__ void __launch_bounds__(256) radix_hist_kernel(const uint32_t* __restrict__ keys, uint64_t n, int shift, uint32_t nblocks, uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    uint64_t base = (uint64_t)blockIdx.x * kSortTile;
#pragma unroll
    for (int k = 0; k < kSortTile / 256; ++k) {
        uint64_t i = base + (uint64_t)k * 256 + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];  // digit-major: one exclusive scan gives every block its bases
}

__global__ void __launch_bounds__(256) radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t n, int shift, uint32_t nblocks,
                                                            const uint64_t* __restrict__ hist_off, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t cnt[8][257];  // per warp, per digit (+1 sentinel column for padding lanes)
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    for (int i = threadIdx.x; i < 8 * 257; i += 256) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const uint64_t wbase = (uint64_t)blockIdx.x * kSortTile + (uint64_t)warp * (kSortTile / 8);
    // sweep 1: per-warp digit counts
    for (int it = 0; it < kSortTile / 8 / 32; ++it) {
        uint64_t i     = wbase + (uint64_t)it * 32 + lane;
        uint32_t d     = i < n ? ((keys_in[i] >> shift) & 255u) : 256u;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        if ((int)lane == __ffs(peers) - 1) cnt[warp][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // digit d: warp w starts after the same digit's elements of warps 0..w-1
    {
        uint32_t d = threadIdx.x, run = 0;
        for (int w = 0; w < 8; ++w) {
            uint32_t c = cnt[w][d];
            cnt[w][d]  = run;
            run += c;
        }
    }
    __syncthreads();
    // sweep 2: same order, stable ranks
    for (int it = 0; it < kSortTile / 8 / 32; ++it) {
        uint64_t i     = wbase + (uint64_t)it * 32 + lane;
        bool     act   = i < n;
        uint32_t key   = act ? keys_in[i] : 0;
        uint32_t d     = act ? ((key >> shift) & 255u) : 256u;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t base  = cnt[warp][d];
        __syncwarp();
        if ((int)lane == __ffs(peers) - 1) cnt[warp][d] = base + __popc(peers);
        __syncwarp();
        if (act) {
            uint64_t dst  = hist_off[(uint64_t)d * nblocks + blockIdx.x] + base + __popc(peers & ((1u << lane) - 1));
            keys_out[dst] = key;
            vals_out[dst] = vals_in[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// p -> IndexReference.  sentence = 1 + delimiters before p (empty sentences are numbered too, reference src/pattern.cpp:1947-1958,
// include/patternmodel.h:1031); token = offset inside the sentence, truncated to 16 bits like IndexReference(sentence, (uint16_t)i)
__global__ void __launch_bounds__(256) refs_from_positions_kernel(const uint32_t* __restrict__ pos, uint64_t n, const uint64_t* __restrict__ sent_before,
                                                                  const uint32_t* __restrict__ sent_start, uint32_t* __restrict__ ref_sentence, uint16_t* __restrict__ ref_token,
                                                                  DeviceStats* __restrict__ st) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t p      = pos[i];
    uint64_t s      = __ldg(sent_before + p);
    uint32_t off    = p - __ldg(sent_start + s);
    ref_sentence[i] = (uint32_t)s + 1;
    ref_token[i]    = (uint16_t)off;
    // the reference would wrap the offset and then sort by the wrapped value; the class encoder never writes such sentences
    // (src/classencoder.cpp:581-588), so this is reported instead of imitated
    if (off > 0xFFFFu) atomicOr(&st->errflags, kErrLongSentence);
}

// ---------------------------------------------------------------------------------------------
int launch_delim_flags(cudaStream_t s, const uint32_t* tok, uint64_t npos, uint32_t* flags) {
    delim_flags_kernel<<<idx_div_up(npos, 256), 256, 0, s>>>(tok, npos, flags);
    return 1;
}
int launch_sent_start(cudaStream_t s, const uint32_t* tok, const uint64_t* sent_before, uint64_t npos, uint32_t* sent_start) {
    sent_start_kernel<<<idx_div_up(npos, 256), 256, 0, s>>>(tok, sent_before, npos, sent_start);
    return 1;
}
int launch_pair_count(cudaStream_t s, const uint32_t* ids, const uint32_t* map, uint64_t npos, bool by_class, uint32_t* blk_counts) {
    pair_count_kernel<<<idx_div_up(npos, kPairTile), 256, 0, s>>>(ids, map, npos, by_class, blk_counts);
    return 1;
}
int launch_pair_write(cudaStream_t s, const uint32_t* ids, const uint32_t* map, uint64_t npos, bool by_class, const uint64_t* blk_off, uint32_t* keys, uint32_t* vals,
                      const uint32_t* pos_lookup, uint32_t pos_div) {
    pair_write_kernel<<<idx_div_up(npos, kPairTile), 256, 0, s>>>(ids, map, npos, by_class, blk_off, keys, vals, pos_lookup, pos_div);
    return 1;
}
int launch_radix_pass(cudaStream_t s, const uint32_t* keys_in, const uint32_t* vals_in, uint64_t n, int shift, uint32_t* hist, uint64_t* hist_off, uint64_t* scan_tmp,
                      uint32_t* keys_out, uint32_t* vals_out) {
    if (!n) return 0;
    uint32_t nblocks = idx_div_up(n, kSortTile);
    radix_hist_kernel<<<nblocks, 256, 0, s>>>(keys_in, n, shift, nblocks, hist);
    int launches = 1 + launch_exclusive_scan_u32_u64(s, hist, hist_off, (uint64_t)256 * nblocks, scan_tmp);
    radix_scatter_kernel<<<nblocks, 256, 0, s>>>(keys_in, vals_in, n, shift, nblocks, hist_off, keys_out, vals_out);
    return launches + 1;
}
int launch_refs_from_positions(cudaStream_t s, const uint32_t* pos, uint64_t n, const uint64_t* sent_before, const uint32_t* sent_start, uint32_t* ref_sentence, uint16_t* ref_token,
                               DeviceStats* st) {
    if (!n) return 0;
    refs_from_positions_kernel<<<idx_div_up(n, 256), 256, 0, s>>>(pos, n, sent_before, sent_start, ref_sentence, ref_token, st);
    return 1;
}

}  // namespace colibri
