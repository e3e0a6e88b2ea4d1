// shard_p2p.cu -- the fused NVLink peer-store levels of a multi-GPU run (see shard.h / shard.cu for the phases they replace).
#include "shard.h"

// ---------------------------------------------------------------------------------------------------------------------
// NVLink peer-store mode.  Every rank owns symmetric receive buffers (allocated and rendezvoused by the caller, e.g. with
// torch.distributed._symmetric_memory): keys_rx = G slots x slot_cap x 8 B, reply_rx = G slots x slot_cap x 4 B,
// surv_rx = G slots x surv_cap x 8 B, hdr = 6*G u64 words:
//   hdr[r]            windows rank r sent me at this level          (written by r's split)
//   hdr[G + r]        survivor records owner r sent me              (written by r's owner phase)
//   hdr[2G + 4r ..]   found, kept, kept occurrences of owner r      (written by r's owner phase)
// The split kernel stores keys straight into the owners' slots and the reply kernel stores ids straight into the senders'
// slots, so a level is: p2p_split, [barrier], p2p_owner, [barrier], p2p_finish -- no all-to-all call, no count exchange.
// The caller provides the barrier (symmetric-memory signal pads) on the stream given to shard_set_stream.
extern "C" int colibri_b200_shard_set_stream(colibri_b200_shard* sh, void* cuda_stream) {
    if (!sh) return set_err(COLIBRI_E_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(sh->dev));
    CUDA_TRY(cudaStreamSynchronize(sh->s));
    if (sh->own_stream) cudaStreamDestroy(sh->s);
    sh->s          = (cudaStream_t)cuda_stream;
    sh->own_stream = false;
    return 0;
}
extern "C" int colibri_b200_shard_set_peers(colibri_b200_shard* sh, const uint64_t* keys_rx, const uint64_t* reply_rx, const uint64_t* surv_rx, const uint64_t* hdr, uint64_t slot_cap,
                                            uint64_t surv_cap) {
    if (!sh || !keys_rx || !reply_rx || !surv_rx || !hdr) return set_err(COLIBRI_E_INVALID, "NULL argument");
    if (slot_cap == 0 || surv_cap == 0 || slot_cap * sh->world >= 0xFFFFFFF0ull) return set_err(COLIBRI_E_INVALID, "slot capacity %llu x %u ranks", (unsigned long long)slot_cap, sh->world);
    CUDA_TRY(cudaSetDevice(sh->dev));
    void* table[256] = {nullptr};
    for (uint32_t r = 0; r < sh->world; ++r) {
        table[r]       = sh->h_keys_rx[r]  = (void*)(uintptr_t)keys_rx[r];
        table[64 + r]  = sh->h_reply_rx[r] = (void*)(uintptr_t)reply_rx[r];
        table[128 + r] = sh->h_surv_rx[r]  = (void*)(uintptr_t)surv_rx[r];
        table[192 + r] = sh->h_hdr[r]      = (void*)(uintptr_t)hdr[r];
    }
    TRY(sh->d_peer.alloc(sh->dev, 256));
    TRY(sh->d_vals.alloc(sh->dev, 64 * 8));
    CUDA_TRY(cudaMemcpyAsync(sh->d_peer.p, table, sizeof table, cudaMemcpyHostToDevice, sh->s));
    CUDA_TRY(cudaStreamSynchronize(sh->s));
    sh->slot_cap = slot_cap;
    sh->surv_cap = surv_cap;
    sh->p2p      = true;
    return 0;
}

extern "C" int colibri_b200_shard_p2p_split(colibri_b200_shard* sh, int n, uint64_t* windows) {
    if (!sh || !windows) return set_err(COLIBRI_E_INVALID, "NULL argument");
    if (!sh->p2p) return set_err(COLIBRI_E_INVALID, "shard_set_peers has not been called");
    uint64_t counts[64];
    TRY(colibri_b200_shard_level_split_count(sh, n, counts, windows));
    for (uint32_t d = 0; d < sh->world; ++d)
        if (counts[d] > sh->slot_cap)
            return set_err(COLIBRI_E_CAPACITY, "level %d: %llu windows for owner %u exceed the receive slot of %llu", n, (unsigned long long)counts[d], d, (unsigned long long)sh->slot_cap);
    PhaseClock clk(sh, 3);
    cudaStream_t s = sh->s;
    if (sh->pos_of_rec.n < sh->nsent + 1) TRY(sh->pos_of_rec.alloc(sh->dev, sh->nsent + 1));
    if (sh->rec_of_pos.n < shard_items(sh) + 8) TRY(sh->rec_of_pos.alloc(sh->dev, shard_items(sh) + 8));
    sh->launches += launch_split_write(s, sh->prev.p, shard_items(sh), sh->world, sh->split_off.p, nullptr, sh->pos_of_rec.p, sh->rec_of_pos.p, sh->d_peer.p, sh->rank, sh->slot_cap, shard_dense_now(sh),
                                       shard_list(sh));
    unsigned long long vals[64];
    for (uint32_t d = 0; d < sh->world; ++d) vals[d] = counts[d];
    CUDA_TRY(cudaMemcpyAsync(sh->d_vals.p, vals, sh->world * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
    sh->launches += launch_p2p_publish(s, sh->d_peer.p + 192, sh->world, sh->rank, sh->d_vals.p, 1);  // hdr[rank] of every owner
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
}

extern "C" int colibri_b200_shard_p2p_owner(colibri_b200_shard* sh, uint64_t stats[3]) {
    if (!sh || !stats) return set_err(COLIBRI_E_INVALID, "NULL argument");
    if (!sh->p2p) return set_err(COLIBRI_E_INVALID, "shard_set_peers has not been called");
    CUDA_TRY(cudaSetDevice(sh->dev));
    PhaseClock clk(sh, 4);
    cudaStream_t   s = sh->s;
    const uint32_t G = sh->world;
    unsigned long long* my_hdr = (unsigned long long*)sh->h_hdr[sh->rank];
    unsigned long long  counts[64];
    CUDA_TRY(cudaMemcpyAsync(counts, my_hdr, G * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    uint64_t nrecv = 0;
    for (uint32_t r = 0; r < G; ++r) {
        if (counts[r] > sh->slot_cap) return set_err(COLIBRI_E_CAPACITY, "corrupt slot header: %llu keys from rank %u", counts[r], r);
        nrecv += counts[r];
    }
    sh->nrecv = nrecv;
    TRY(shard_dense_owner(sh));  // (the caller has summed the dense squares by now)
    const uint64_t phys = (uint64_t)G * sh->slot_cap;  // slots are scanned whole; entries past a slot's count are skipped
    const void*    keys = sh->h_keys_rx[sh->rank];
    if (sh->rid.n < phys) TRY(sh->rid.alloc(sh->dev, phys));

    const uint32_t t          = sh->t;
    const Tuning   tune       = Tuning::from_env();
    const bool     use_filter = tune.use_filter(t, nrecv);
    uint64_t       nbuckets = 0, cap = std::max<uint64_t>(64, nrecv + nrecv / 2 + 16);
    if (use_filter) {
        nbuckets = tune.filter_buckets(nrecv);
        if (sh->filter.n < nbuckets / 16) TRY(sh->filter.alloc(sh->dev, nbuckets / 16));
        CUDA_TRY(cudaMemsetAsync(sh->filter.p, 0, nbuckets / 4, s));
        TRY(shard_zero_stats(sh));
        sh->launches += launch_stream_filter(s, keys, nrecv, sh->filter.p, nbuckets, sh->d_stats.p, sh->sms, sh->slot_cap, my_hdr, G);
        TRY(shard_read_stats(sh));
        if (sh->h_stats.found * 8 <= nbuckets) cap = std::min(cap, std::max<uint64_t>(1024, 3 * sh->h_stats.found + 1024));  // (a saturated filter says nothing about the number of keys)
        if (tune.filter_1bit && nbuckets >= 64) {
            if (sh->filter1.n < nbuckets / 32 + 8) TRY(sh->filter1.alloc(sh->dev, nbuckets / 32 + 8));
            sh->launches += launch_filter_to_bitmap(s, sh->filter.p, nbuckets, sh->filter1.p);
        }
    }
    const bool onebit = use_filter && tune.filter_1bit && nbuckets >= 64;
    const uint64_t cap_max = std::max<uint64_t>(64, nrecv + nrecv / 2 + 16);  // a table this large cannot fill up
    uint64_t singles = 0;
    for (;;) {
        if (cap * G + shard_id_off(sh) >= 0xFFFFFFF0ull) return set_err(COLIBRI_E_CAPACITY, "owner table of %llu slots x %u ranks exceeds the 32-bit id space", (unsigned long long)cap, G);
        if (sh->owner_table.n < cap) TRY(sh->owner_table.alloc(sh->dev, cap));
        CUDA_TRY(cudaMemsetAsync(sh->owner_table.p, 0, cap * sizeof(NgramSlot), s));
        TRY(shard_zero_stats(sh));
        sh->launches += launch_stream_count(s, keys, nrecv, sh->owner_table.p, cap, use_filter ? (onebit ? sh->filter1.p : sh->filter.p) : nullptr, nbuckets, sh->rid.p, sh->d_stats.p, sh->sms, sh->slot_cap, my_hdr, G, onebit);
        CUDA_TRY(cudaMemcpyAsync(&sh->h_stats, sh->d_stats.p, sizeof(DeviceStats), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (sh->h_stats.errflags & kErrTableFull) {
            if (cap >= cap_max) return set_err(COLIBRI_E_CAPACITY, "owner hash table overflow");
            CUDA_TRY(cudaMemsetAsync(&sh->d_stats.p->errflags, 0, sizeof(unsigned int), s));
            cap = std::min(cap_max, cap * 4);
            continue;
        }
        singles = sh->h_stats.singletons;
        break;
    }
    const uint64_t sv_bound = (nrecv - singles) / std::max<uint32_t>(t, 1) + 1;
    if (sh->sv_idx.n < sv_bound) TRY(sh->sv_idx.alloc(sh->dev, sv_bound));
    if (sh->sv_cnt.n < sv_bound) TRY(sh->sv_cnt.alloc(sh->dev, sv_bound));
    if (sh->bitmap.n < cap / 32 + 8) TRY(sh->bitmap.alloc(sh->dev, cap / 32 + 8));
    TRY(shard_zero_stats(sh));
    sh->launches += launch_prune_ngrams(s, sh->owner_table.p, cap, t, sh->sv_idx.p, sh->sv_cnt.p, sh->bitmap.p, sh->d_stats.p, sh->sms);
    sh->launches += launch_owner_reply(s, sh->rid.p, nrecv, sh->bitmap.p, G, sh->rank, sh->d_peer.p + 64, sh->slot_cap, my_hdr, shard_id_off(sh));  // ids -> the senders' reply slots
    TRY(shard_read_stats(sh));
    stats[0]  = sh->h_stats.found + singles + sh->dense_stats[0];
    stats[1]  = sh->h_stats.kept + sh->dense_stats[1];
    stats[2]  = sh->h_stats.kept_occ + sh->dense_stats[2];
    sh->nsurv = sh->h_stats.kept;
    // survivors -> the claimers' survivor slots; the cursors become the counts the sources read from their headers
    if (sh->d_aux.n < 260) TRY(sh->d_aux.alloc(sh->dev, 260));
    CUDA_TRY(cudaMemsetAsync(sh->d_aux.p, 0, 64 * sizeof(unsigned long long), s));
    sh->launches += launch_owner_survivors_p2p(s, sh->sv_idx.p, sh->sv_cnt.p, sh->nsurv, G, sh->rank, sh->slot_cap, sh->surv_cap, sh->d_aux.p, sh->d_peer.p + 128, sh->d_stats.p, sh->sms);
    sh->launches += launch_p2p_publish(s, sh->d_peer.p + 192, G, G + sh->rank, sh->d_aux.p, 1);  // hdr[G + rank] of every source
    unsigned long long sv[64 * 3];
    for (uint32_t d = 0; d < G; ++d) {
        sv[3 * d]     = stats[0];
        sv[3 * d + 1] = stats[1];
        sv[3 * d + 2] = stats[2];
    }
    CUDA_TRY(cudaMemcpyAsync(sh->d_vals.p, sv, G * 3 * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
    sh->launches += launch_p2p_publish(s, sh->d_peer.p + 192, G, 2 * G + 4 * sh->rank, sh->d_vals.p, 3);  // hdr[2G + 4*rank ..] of every rank
    TRY(shard_read_stats(sh));
    if (sh->h_stats.errflags & kErrTableFull) return set_err(COLIBRI_E_CAPACITY, "survivor slot overflow (surv_cap %llu)", (unsigned long long)sh->surv_cap);
    return 0;
}

extern "C" int colibri_b200_shard_p2p_finish(colibri_b200_shard* sh, uint64_t global_stats[3], uint64_t* local_valid) {
    if (!sh || !global_stats) return set_err(COLIBRI_E_INVALID, "NULL argument");
    if (!sh->p2p) return set_err(COLIBRI_E_INVALID, "shard_set_peers has not been called");
    CUDA_TRY(cudaSetDevice(sh->dev));
    cudaStream_t   s = sh->s;
    const uint32_t G = sh->world;
    unsigned long long hdr[6 * 64];
    CUDA_TRY(cudaMemcpyAsync(hdr, sh->h_hdr[sh->rank], 6 * G * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    global_stats[0] = global_stats[1] = global_stats[2] = 0;
    uint64_t surv_counts[64];
    for (uint32_t r = 0; r < G; ++r) {
        surv_counts[r] = hdr[G + r];
        for (int k = 0; k < 3; ++k) global_stats[k] += hdr[2 * G + 4 * r + k];
        if (surv_counts[r] > sh->surv_cap) return set_err(COLIBRI_E_CAPACITY, "survivor slot overflow reported by owner %u", r);
    }
    PhaseClock clk(sh, 5);
    const int n = sh->level + 1;
    TRY(shard_zero_stats(sh));
    CUDA_TRY(cudaMemsetAsync(sh->cur.p + sh->npos, 0, 8 * sizeof(uint32_t), s));
    if (sh->list_valid) CUDA_TRY(cudaMemsetAsync(sh->cur.p, 0, sh->npos * sizeof(uint32_t), s));  // list mode writes only the positions that keep an id
    sh->launches += launch_sender_relabel(s, sh->rec_of_pos.p, (const uint32_t*)sh->h_reply_rx[sh->rank], shard_items(sh), sh->cur.p, sh->d_stats.p, sh->sms,
                                          shard_dense_now(sh) ? sh->dense_cnt : nullptr, sh->t, shard_list(sh));
    uint64_t total = 0;
    for (uint32_t g = 0; g < G; ++g) total += surv_counts[g];
    Segment sg;
    sg.n = n;
    const uint64_t nd = shard_dense_now(sh) ? sh->dense_nsurv : 0;  // this rank's share of the dense square's survivors goes into the same segment
    if (total + nd) {
        TRY(sg.pos.alloc(sh->dev, total + nd));
        TRY(sg.cnt.alloc(sh->dev, total + nd));
        uint64_t off = 0;
        for (uint32_t g = 0; g < G; ++g) {
            const uint8_t* recs = (const uint8_t*)sh->h_surv_rx[sh->rank] + (uint64_t)g * sh->surv_cap * 8;
            sh->launches += launch_sender_survivors(s, recs, surv_counts[g], sh->pos_of_rec.p, sh->send_base[g], sg.pos.p + off, sg.cnt.p + off);
            off += surv_counts[g];
        }
        if (nd) {
            CUDA_TRY(cudaMemcpyAsync(sg.pos.p + total, sh->dense_sv_pos.p, nd * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(sg.cnt.p + total, sh->dense_sv_cnt.p, nd * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        }
    }
    TRY(shard_read_stats(sh));
    sg.count = total + nd;
    if (total + nd) sh->segs.push_back(std::move(sg));
    sh->prev_valid = sh->h_stats.kept_occ;
    if (local_valid) *local_valid = sh->prev_valid;
    std::swap(sh->prev, sh->cur);
    sh->level = n;
    TRY(shard_keep_ids(sh, n));
    TRY(shard_next_list(sh));
    return 0;
}

