// shard.h -- the per-rank state of a multi-GPU run (shard.cu: phases driven through NCCL all-to-alls; shard_p2p.cu: the fused
// NVLink peer-store levels) and the small helpers both files use.
#pragma once
#include "engine_common.h"

using namespace colibri;

struct colibri_b200_shard {
    colibri_b200_corpus* corpus = nullptr;
    colibri_b200_options o;
    int                  dev = 0, sms = 148;
    uint32_t             rank = 0, world = 1;
    cudaStream_t         s = nullptr;
    uint64_t             launches = 0;
    DevBuf<DeviceStats>  d_stats;
    DeviceStats          h_stats;
    uint64_t             npos = 0, local_tokens = 0;
    uint32_t             local_maxclass = 0, nclasses = 0;
    DevBuf<uint32_t>     filter1;  // the occurrence filter's "hit twice" bits, packed (Tuning::filter_1bit)
    DevBuf<uint32_t>     tok, count1, prev, cur, bitmap, filter, pos_of_rec, rec_of_pos, split_hist, sv_idx, sv_cnt;
    DevBuf<uint64_t>     split_off, scan_tmp;
    DevBuf<NgramSlot>    owner_table;
    DevBuf<unsigned long long> d_aux;   // [0..65): source bases of the receive buffer, [65..130): output bases, [130..195): cursors, [195..260): counts
    uint64_t             send_base[65] = {0}, nsent = 0, nrecv = 0, nsurv = 0, prev_valid = 0;
    uint64_t             surv_out_counts[64] = {0};
    // NVLink peer-store mode: symmetric receive buffers of every rank (device pointers valid in this process)
    bool                 p2p = false, own_stream = true;
    uint64_t             slot_cap = 0, surv_cap = 0;
    void*                h_keys_rx[64] = {nullptr};
    void*                h_reply_rx[64] = {nullptr};
    void*                h_surv_rx[64] = {nullptr};
    void*                h_hdr[64] = {nullptr};
    DevBuf<void*>        d_peer;   // [0..64) keys_rx, [64..128) reply_rx, [128..192) surv_rx, [192..256) hdr
    DevBuf<unsigned long long> d_vals;
    DevBuf<uint32_t>     rid;
    // skipgrams: every level's (global) ids are kept, the parts of a skipgram are looked up in them
    std::vector<DevBuf<uint32_t>> ids_keep;
    DevBuf<const uint32_t*> d_idptrs;
    DevBuf<SkipMask>     d_masks;
    DevBuf<SkipSlot>     sktable;
    DevBuf<uint32_t>     sk_pos_of_rec, sk_sv_idx, sk_sv_cnt, sk_sv_mask;
    int                  sk_nmasks = 0;
    uint64_t             sk_send_base[65] = {0}, sk_nsent = 0, sk_nsurv = 0;
    // dense pairs of level 2 (shard_set_dense): the caller's zeroed square, summed over the ranks by the caller between split and owner
    uint32_t             dense = 0;
    uint32_t*            dense_cnt = nullptr;
    uint64_t             tok_ext_cells = 0;  // room behind the tokens for the class pairs of this rank's dense survivors
    DevBuf<uint32_t>     dense_sv_pos, dense_sv_cnt;
    uint64_t             dense_stats[3] = {0, 0, 0}, dense_nsurv = 0;
    // list mode of the sparse later levels: the positions whose newest id is non-zero (see kernels.cu: load_window; shard_kernels.cu: window_dest)
    DevBuf<uint32_t>     list;
    uint64_t             nlist = 0;
    bool                 list_valid = false;
    int                  level = 1;
    uint32_t             t = 2;
    std::vector<Segment> segs;
    uint64_t             global_tokens = 0, global_types = 0;
    cudaEvent_t          ev0 = nullptr, ev1 = nullptr;
    double               device_ms = 0;
    double               phase_ms[8] = {0};  // begin, unigrams, count, pack, merge, finish, export
};

// skipgram mode: remember the id array of level n (sh->prev after the level's finish)
inline int shard_keep_ids(colibri_b200_shard* sh, int n) {
    if (!sh->o.DOSKIPGRAMS_EXHAUSTIVE) return 0;
    if ((int)sh->ids_keep.size() <= n) sh->ids_keep.resize(n + 1);
    TRY(sh->ids_keep[n].alloc(sh->dev, sh->npos + 8));
    CUDA_TRY(cudaMemcpyAsync(sh->ids_keep[n].p, sh->prev.p, (sh->npos + 8) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, sh->s));
    return 0;
}

inline int shard_zero_stats(colibri_b200_shard* sh) {
    CUDA_TRY(cudaMemsetAsync(&sh->d_stats.p->found, 0, offsetof(DeviceStats, maxclass) - offsetof(DeviceStats, found), sh->s));
    return 0;
}
inline int shard_read_stats(colibri_b200_shard* sh) {
    CUDA_TRY(cudaMemcpyAsync(&sh->h_stats, sh->d_stats.p, sizeof(DeviceStats), cudaMemcpyDeviceToHost, sh->s));
    CUDA_TRY(cudaStreamSynchronize(sh->s));
    if (sh->h_stats.errflags & kErrTableFull) return set_err(COLIBRI_E_CAPACITY, "device hash table overflow");
    return 0;
}
// level 2 with a dense square: this rank's share of the (already summed) square -> statistics + survivors, kept until the level's finish
inline int shard_dense_owner(colibri_b200_shard* sh) {
    sh->dense_stats[0] = sh->dense_stats[1] = sh->dense_stats[2] = 0;
    sh->dense_nsurv = 0;
    if (!sh->dense || sh->level != 1) return 0;
    const uint64_t mine = ((uint64_t)sh->dense * sh->dense + sh->world - 1) / sh->world;
    if (sh->dense_sv_pos.n < mine + 1) TRY(sh->dense_sv_pos.alloc(sh->dev, mine + 1));
    if (sh->dense_sv_cnt.n < mine + 1) TRY(sh->dense_sv_cnt.alloc(sh->dev, mine + 1));
    TRY(shard_zero_stats(sh));
    sh->launches += launch_dense_share(sh->s, sh->dense_cnt, sh->dense, sh->world, sh->rank, sh->t, sh->dense_sv_pos.p, sh->dense_sv_cnt.p, sh->tok.p + sh->npos + 8,
                                       (uint32_t)(sh->npos + 8), sh->d_stats.p, sh->sms);
    TRY(shard_read_stats(sh));
    sh->dense_stats[0] = sh->h_stats.found;
    sh->dense_stats[1] = sh->h_stats.kept;
    sh->dense_stats[2] = sh->h_stats.kept_occ;
    sh->dense_nsurv    = sh->h_stats.kept;
    return 0;
}
inline uint64_t shard_items(const colibri_b200_shard* sh) { return sh->list_valid ? sh->nlist : sh->npos; }
inline const uint32_t* shard_list(const colibri_b200_shard* sh) { return sh->list_valid ? sh->list.p : nullptr; }
// after a level's finish (sh->prev = the new ids, sh->prev_valid of them non-zero): the next level runs from a position list when few positions are left
inline int shard_next_list(colibri_b200_shard* sh) {
    const Tuning tune = Tuning::from_env();
    sh->list_valid    = false;
    if (sh->level >= sh->o.MAXLENGTH || tune.sparse_div == 0 || sh->prev_valid == 0 || sh->prev_valid * tune.sparse_div > sh->npos) return 0;
    if (sh->list.n < sh->prev_valid + 8) TRY(sh->list.alloc(sh->dev, sh->prev_valid + 8));
    CUDA_TRY(cudaMemsetAsync(&sh->d_stats.p->cursor, 0, sizeof(unsigned long long), sh->s));
    sh->launches += launch_compact_nonzero(sh->s, sh->prev.p, sh->npos, sh->list.p, &sh->d_stats.p->cursor);
    sh->nlist      = sh->prev_valid;
    sh->list_valid = true;
    return 0;
}
inline uint32_t shard_dense_now(const colibri_b200_shard* sh) {  // the dense side of the level being built (level + 1)
    return sh->level == 1 ? sh->dense : 0u;
}
inline uint32_t shard_id_off(const colibri_b200_shard* sh) {
    return sh->level == 1 ? sh->dense * sh->dense : 0u;
}

struct PhaseClock {  // accumulates device time of one ABI call into shard->device_ms / phase_ms[phase]
    colibri_b200_shard* sh;
    int                 phase;
    explicit PhaseClock(colibri_b200_shard* s, int ph) : sh(s), phase(ph) { cudaEventRecord(sh->ev0, sh->s); }
    ~PhaseClock() {
        cudaEventRecord(sh->ev1, sh->s);
        cudaEventSynchronize(sh->ev1);
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sh->ev0, sh->ev1) == cudaSuccess) {
            sh->device_ms += ms;
            sh->phase_ms[phase] += ms;
        }
    }
};

