// shard.cu -- multi-GPU training as a sequence of per-rank phases (C ABI section "multi-GPU" of include/colibri_b200.h).
//
// The reference has no notion of several devices.  The corpus shards at sentence boundaries (windows
// never cross the 0x00 delimiter, reference include/patternmodel.h:1063), one shard per GPU, and the model is partitioned
// by hash: owner(key) = hash(key) mod G.  One process per GPU drives these phases and moves the buffers between ranks
// with torch.distributed (NCCL all-to-all over NVLink); this file contains no communication and no torch types.
//
// Per level n >= 2 and per rank ("ship the windows to their owners", see shard_kernels.cu):
//   level_split_count / level_split_write   valid windows -> 8-byte keys grouped by owner rank, corpus order inside a group
//   [all-to-all of keys]
//   level_owner            occurrence filter + table upserts over the received keys (the single-GPU kernels on a key stream),
//                          threshold (prune(), patternmodel.h:2107-2128), reply = global id per received window (4 bytes)
//   level_owner_survivors  (index in the sender's group, global count) for the window that claimed each surviving slot
//   [all-to-all back of the replies, all-to-all of the survivor records]
//   level_finish           id[p] = reply[rec_of_pos[p]] (streaming), survivors -> (position, count) this rank exports
// Level 1 needs no table: the class histograms are summed with an all-reduce.
#include "shard.h"

using namespace colibri;

extern "C" void colibri_b200_shard_free(colibri_b200_shard* sh) {
    if (!sh) return;
    cudaSetDevice(sh->dev);
    if (sh->s) {
        cudaStreamSynchronize(sh->s);
        if (sh->own_stream) cudaStreamDestroy(sh->s);
    }
    if (sh->ev0) cudaEventDestroy(sh->ev0);
    if (sh->ev1) cudaEventDestroy(sh->ev1);
    delete sh;
}

extern "C" int colibri_b200_shard_begin(colibri_b200_corpus* corpus, const colibri_b200_options* opt, int rank, int world, colibri_b200_shard** out) {
    if (!out) return set_err(COLIBRI_E_INVALID, "out is NULL");
    *out = nullptr;
    if (!corpus || !opt) return set_err(COLIBRI_E_INVALID, "NULL argument");
    if (world < 1 || world > 64 || rank < 0 || rank >= world) return set_err(COLIBRI_E_INVALID, "rank %d / world %d", rank, world);
    colibri_b200_options o = *opt;
    TRY(check_options(o));
    if (o.model_type == COLIBRI_INDEXEDPATTERNMODEL) return set_err(COLIBRI_E_UNSUPPORTED, "indexed models are not on the multi-GPU path yet");
    if (corpus->nbytes == 0) return set_err(COLIBRI_E_FORMAT, "Attempting to read pattern from file, but file is empty?");
    CUDA_TRY(cudaSetDevice(corpus->device));
    auto* sh   = new colibri_b200_shard();
    sh->corpus = corpus;
    sh->o      = o;
    sh->dev    = corpus->device;
    sh->rank   = (uint32_t)rank;
    sh->world  = (uint32_t)world;
    sh->t      = (uint32_t)o.MINTOKENS;
    int rc     = [&]() -> int {
        CUDA_TRY(cudaDeviceGetAttribute(&sh->sms, cudaDevAttrMultiProcessorCount, sh->dev));
        CUDA_TRY(cudaStreamCreateWithFlags(&sh->s, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreate(&sh->ev0));
        CUDA_TRY(cudaEventCreate(&sh->ev1));
        PhaseClock clk(sh, 0);
        cudaStream_t s = sh->s;
        // sentence-source quirk exactly as in the single-GPU driver (a shard is a whole file of its own here)
        size_t staged = corpus->nbytes;
        uint8_t tail[16];
        memset(tail, 0x80, sizeof tail);
        size_t k = 0;
        if (!corpus->ends_with_delim) {
            if (o.streamed) tail[k++] = corpus->last_byte;
            tail[k++] = 0;
        }
        staged += k;
        CUDA_TRY(cudaMemcpyAsync(corpus->body() + corpus->nbytes, tail, sizeof tail, cudaMemcpyHostToDevice, s));
        const uint32_t nblocks = (uint32_t)(corpus->padded(staged) / kTokTile);
        TRY(sh->d_stats.alloc(sh->dev, 1));
        CUDA_TRY(cudaMemsetAsync(sh->d_stats.p, 0, sizeof(DeviceStats), s));
        DevBuf<uint32_t> blk;
        TRY(blk.alloc(sh->dev, nblocks + 1));
        sh->launches += launch_tokenise_count(s, corpus->body(), staged, blk.p, nblocks);
        sh->launches += launch_scan_block_counts(s, blk.p, nblocks, &sh->d_stats.p->cursor);
        TRY(shard_read_stats(sh));
        const uint64_t npos_real = sh->h_stats.cursor;
        if (npos_real >= 0xFFFFFFF0ull) return set_err(COLIBRI_E_CAPACITY, "shard has %llu positions; the device index is 32 bit", (unsigned long long)npos_real);
        sh->npos = npos_real + 1;
        {   // spare room behind the tokens for the class pairs of dense survivors (shard_set_dense; 32 MB at the default side)
            const Tuning tune = Tuning::from_env();
            sh->tok_ext_cells = (uint64_t)tune.dense_dim * tune.dense_dim;
            if (sh->npos + 8 + 2 * sh->tok_ext_cells >= 0xFFFFFFF0ull) sh->tok_ext_cells = 0;
        }
        TRY(sh->tok.alloc(sh->dev, sh->npos + 8 + 2 * sh->tok_ext_cells));
        CUDA_TRY(cudaMemsetAsync(sh->tok.p + npos_real, 0, 8 * sizeof(uint32_t), s));
        sh->launches += launch_tokenise_write(s, corpus->body(), staged, blk.p, nblocks, sh->tok.p, sh->d_stats.p);
        TRY(shard_read_stats(sh));
        if (sh->h_stats.errflags & (kErrTokenTooLong | kErrNonCanonical)) return set_err(COLIBRI_E_FORMAT, "malformed class encoding in corpus");
        if (sh->h_stats.errflags & kErrReservedClass) return set_err(COLIBRI_E_UNSUPPORTED, "corpus contains the reserved skip/flex classes (3, 4) as running text");
        sh->local_tokens   = sh->h_stats.totaltokens;
        sh->local_maxclass = sh->h_stats.maxclass;
        return 0;
    }();
    if (rc) {
        colibri_b200_shard_free(sh);
        return rc;
    }
    *out = sh;
    return 0;
}

extern "C" int colibri_b200_shard_info(const colibri_b200_shard* sh, uint64_t out[4]) {
    if (!sh || !out) return set_err(COLIBRI_E_INVALID, "NULL argument");
    out[0] = sh->local_tokens;
    out[1] = sh->local_maxclass;
    out[2] = sh->npos;
    out[3] = sh->launches;
    return 0;
}
extern "C" double colibri_b200_shard_device_ms(const colibri_b200_shard* sh) {
    return sh ? sh->device_ms : 0.0;
}
extern "C" int colibri_b200_shard_phase_ms(const colibri_b200_shard* sh, double out[8]) {
    if (!sh || !out) return set_err(COLIBRI_E_INVALID, "NULL argument");
    memcpy(out, sh->phase_ms, sizeof sh->phase_ms);
    return 0;
}

// local class histogram into the caller's device buffer (u32[nclasses], zeroed here); the caller all-reduces it
extern "C" int colibri_b200_shard_unigram_counts(colibri_b200_shard* sh, uint32_t nclasses, void* dev_counts) {
    if (!sh || !dev_counts) return set_err(COLIBRI_E_INVALID, "NULL argument");
    if (nclasses <= sh->local_maxclass) return set_err(COLIBRI_E_INVALID, "nclasses %u <= local maximum class %u", nclasses, sh->local_maxclass);
    CUDA_TRY(cudaSetDevice(sh->dev));
    PhaseClock clk(sh, 1);
    sh->nclasses = nclasses;
    CUDA_TRY(cudaMemsetAsync(dev_counts, 0, (size_t)nclasses * 4, sh->s));
    sh->launches += launch_unigram_hist(sh->s, sh->tok.p, sh->npos, (uint32_t*)dev_counts, nclasses, sh->sms);
    CUDA_TRY(cudaStreamSynchronize(sh->s));
    return 0;
}

// global counts in: prune (identical on every rank), export this rank's share of the surviving unigrams, level-1 ids
extern "C" int colibri_b200_shard_unigram_finish(colibri_b200_shard* sh, const void* dev_global_counts, uint64_t global_tokens, uint64_t stats[3]) {
    if (!sh || !dev_global_counts || !stats) return set_err(COLIBRI_E_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(sh->dev));
    PhaseClock clk(sh, 1);
    cudaStream_t s = sh->s;
    TRY(sh->count1.alloc(sh->dev, sh->nclasses));
    CUDA_TRY(cudaMemcpyAsync(sh->count1.p, dev_global_counts, (size_t)sh->nclasses * 4, cudaMemcpyDeviceToDevice, s));
    sh->global_tokens = global_tokens;
    TRY(shard_zero_stats(sh));
    Segment sg;
    sg.n = 1;
    uint64_t bound = (uint64_t)sh->nclasses / sh->world + 2;
    TRY(sg.pos.alloc(sh->dev, bound));
    TRY(sg.cnt.alloc(sh->dev, bound));
    sh->launches += launch_unigram_prune(s, sh->count1.p, sh->nclasses, sh->t, sg.pos.p, sg.cnt.p, 0, sh->d_stats.p, sh->world, sh->rank);
    TRY(shard_read_stats(sh));
    sg.count = sh->h_stats.cursor;
    sh->segs.push_back(std::move(sg));
    stats[0] = sh->h_stats.found;
    stats[1] = sh->h_stats.kept;
    stats[2] = sh->h_stats.kept_occ;
    sh->global_types = sh->h_stats.found;
    const uint32_t t1 = (uint32_t)std::max(sh->o.MINTOKENS, sh->o.MINTOKENS_UNIGRAMS);
    TRY(sh->prev.alloc(sh->dev, sh->npos + 8));
    TRY(sh->cur.alloc(sh->dev, sh->npos + 8));
    sh->launches += launch_make_id1(s, sh->tok.p, sh->npos + 1, sh->count1.p, t1, sh->prev.p);
    CUDA_TRY(cudaMemsetAsync(sh->prev.p + sh->npos, 0, 8 * sizeof(uint32_t), s));
    sh->level = 1;
    TRY(shard_keep_ids(sh, 1));
    CUDA_TRY(cudaStreamSynchronize(s));
    sh->prev_valid = sh->local_tokens;
    return 0;
}

// Dense pairs of level 2.  dev_square: the caller's zeroed u32[dim * dim]; level_split_count / p2p_split count this rank's dense windows
// into it, the CALLER sums the squares of all ranks (one all-reduce) before level_owner / p2p_owner, and every rank then reads the same verdicts
// from it.  All ranks must make the same call (same dim) or none: the ids of level 2 depend on it.
extern "C" int colibri_b200_shard_set_dense(colibri_b200_shard* sh, void* dev_square, uint32_t dim) {
    if (!sh) return set_err(COLIBRI_E_INVALID, "NULL argument");
    if (sh->level != 1) return set_err(COLIBRI_E_INVALID, "shard_set_dense comes before level 2");
    if (dim == 0 || !dev_square) {
        sh->dense     = 0;
        sh->dense_cnt = nullptr;
        return 0;
    }
    if ((uint64_t)dim * dim > sh->tok_ext_cells) return set_err(COLIBRI_E_INVALID, "dense side %u: the shard keeps room for %llu cells (COLIBRI_B200_DENSE at shard_begin)", dim, (unsigned long long)sh->tok_ext_cells);
    if (sh->p2p && sh->slot_cap * sh->world >= 0x80000000ull) return set_err(COLIBRI_E_CAPACITY, "receive slots of %llu x %u: the record index has 31 bits beside the dense square", (unsigned long long)sh->slot_cap, sh->world);
    sh->dense     = dim;
    sh->dense_cnt = static_cast<uint32_t*>(dev_square);
    return 0;
}

// count the valid windows of level n per owner rank.  send_counts[world]; *windows = their sum
extern "C" int colibri_b200_shard_level_split_count(colibri_b200_shard* sh, int n, uint64_t* send_counts, uint64_t* windows) {
    if (!sh || !send_counts || !windows) return set_err(COLIBRI_E_INVALID, "NULL argument");
    if (n != sh->level + 1) return set_err(COLIBRI_E_INVALID, "level %d requested after level %d", n, sh->level);
    CUDA_TRY(cudaSetDevice(sh->dev));
    PhaseClock clk(sh, 2);
    cudaStream_t   s       = sh->s;
    const uint64_t nblocks = (shard_items(sh) + 4095) / 4096;
    const uint64_t nh      = (uint64_t)sh->world * nblocks;
    if (sh->split_hist.n < nh) TRY(sh->split_hist.alloc(sh->dev, nh));
    if (sh->split_off.n < nh + 1) TRY(sh->split_off.alloc(sh->dev, nh + 1));
    if (sh->scan_tmp.n < nh / 2048 + 4) TRY(sh->scan_tmp.alloc(sh->dev, nh / 2048 + 4));
    sh->launches += launch_split_count(s, sh->prev.p, shard_items(sh), sh->world, sh->split_hist.p, shard_dense_now(sh), sh->dense_cnt, shard_list(sh));
    sh->launches += launch_exclusive_scan_u32_u64(s, sh->split_hist.p, sh->split_off.p, nh, sh->scan_tmp.p);
    for (uint32_t d = 0; d <= sh->world; ++d)
        CUDA_TRY(cudaMemcpyAsync(&sh->send_base[d], sh->split_off.p + (uint64_t)d * nblocks, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (uint32_t d = 0; d < sh->world; ++d) send_counts[d] = sh->send_base[d + 1] - sh->send_base[d];
    sh->nsent = sh->send_base[sh->world];
    *windows  = sh->nsent;
    return 0;
}

// write the nsent 8-byte keys, grouped by owner rank in rank order, into dev_send_keys
extern "C" int colibri_b200_shard_level_split_write(colibri_b200_shard* sh, void* dev_send_keys) {
    if (!sh || (!dev_send_keys && sh->nsent)) return set_err(COLIBRI_E_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(sh->dev));
    PhaseClock clk(sh, 3);
    if (sh->pos_of_rec.n < sh->nsent + 1) TRY(sh->pos_of_rec.alloc(sh->dev, sh->nsent + 1));
    if (sh->rec_of_pos.n < shard_items(sh) + 8) TRY(sh->rec_of_pos.alloc(sh->dev, shard_items(sh) + 8));
    if (shard_dense_now(sh) && sh->nsent >= 0x80000000ull) return set_err(COLIBRI_E_CAPACITY, "%llu windows to ship; the record index has 31 bits beside the dense square", (unsigned long long)sh->nsent);
    sh->launches += launch_split_write(sh->s, sh->prev.p, shard_items(sh), sh->world, sh->split_off.p, dev_send_keys, sh->pos_of_rec.p, sh->rec_of_pos.p, nullptr, 0, 0, shard_dense_now(sh), shard_list(sh));
    CUDA_TRY(cudaStreamSynchronize(sh->s));
    return 0;
}

// owner side.  dev_recv_keys: nrecv = sum(recv_counts) keys grouped by source rank; dev_reply: nrecv u32 (global id | 0 per window).
// stats = {distinct keys owned, kept, kept occurrences}; surv_counts[world] = survivor records to return to each source.
extern "C" int colibri_b200_shard_level_owner(colibri_b200_shard* sh, const void* dev_recv_keys, const uint64_t* recv_counts, void* dev_reply, uint64_t stats[3],
                                              uint64_t* surv_counts) {
    if (!sh || !recv_counts || !stats || !surv_counts) return set_err(COLIBRI_E_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(sh->dev));
    PhaseClock clk(sh, 4);
    cudaStream_t s = sh->s;
    unsigned long long src_base[65];
    uint64_t nrecv = 0;
    for (uint32_t r = 0; r < sh->world; ++r) {
        src_base[r] = nrecv;
        nrecv += recv_counts[r];
    }
    src_base[sh->world] = nrecv;
    sh->nrecv = nrecv;
    if (nrecv && (!dev_recv_keys || !dev_reply)) return set_err(COLIBRI_E_INVALID, "NULL buffer");
    if (nrecv >= 0xFFFFFFF0ull) return set_err(COLIBRI_E_CAPACITY, "owner received %llu windows; the receive index is 32 bit", (unsigned long long)nrecv);
    TRY(shard_dense_owner(sh));  // (the caller has summed the dense squares by now)
    if (sh->d_aux.n < 260) TRY(sh->d_aux.alloc(sh->dev, 260));
    CUDA_TRY(cudaMemsetAsync(sh->d_aux.p, 0, 260 * sizeof(unsigned long long), s));
    CUDA_TRY(cudaMemcpyAsync(sh->d_aux.p, src_base, (sh->world + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));

    const uint32_t t          = sh->t;
    const Tuning   tune       = Tuning::from_env();
    const bool     use_filter = tune.use_filter(t, nrecv);
    uint64_t       nbuckets = 0, cap = std::max<uint64_t>(64, nrecv + nrecv / 2 + 16);
    if (use_filter) {
        nbuckets = tune.filter_buckets(nrecv);
        if (sh->filter.n < nbuckets / 16) TRY(sh->filter.alloc(sh->dev, nbuckets / 16));
        CUDA_TRY(cudaMemsetAsync(sh->filter.p, 0, nbuckets / 4, s));
        TRY(shard_zero_stats(sh));
        sh->launches += launch_stream_filter(s, dev_recv_keys, nrecv, sh->filter.p, nbuckets, sh->d_stats.p, sh->sms);
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
        if (cap * sh->world + shard_id_off(sh) >= 0xFFFFFFF0ull) return set_err(COLIBRI_E_CAPACITY, "owner table of %llu slots x %u ranks exceeds the 32-bit id space", (unsigned long long)cap, sh->world);
        if (sh->owner_table.n < cap) TRY(sh->owner_table.alloc(sh->dev, cap));
        CUDA_TRY(cudaMemsetAsync(sh->owner_table.p, 0, cap * sizeof(NgramSlot), s));
        TRY(shard_zero_stats(sh));
        sh->launches += launch_stream_count(s, dev_recv_keys, nrecv, sh->owner_table.p, cap, use_filter ? (onebit ? sh->filter1.p : sh->filter.p) : nullptr, nbuckets, (uint32_t*)dev_reply, sh->d_stats.p, sh->sms, 0, nullptr, 0, onebit);
        CUDA_TRY(cudaMemcpyAsync(&sh->h_stats, sh->d_stats.p, sizeof(DeviceStats), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (sh->h_stats.errflags & kErrTableFull) {  // the estimate was too small: go again with twice the slots
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
    sh->launches += launch_owner_reply(s, (uint32_t*)dev_reply, nrecv, sh->bitmap.p, sh->world, sh->rank, nullptr, 0, nullptr, shard_id_off(sh));
    TRY(shard_read_stats(sh));
    stats[0]  = sh->h_stats.found + singles + sh->dense_stats[0];  // a filtered window is a distinct n-gram with global count 1: found, and pruned
    stats[1]  = sh->h_stats.kept + sh->dense_stats[1];
    stats[2]  = sh->h_stats.kept_occ + sh->dense_stats[2];
    sh->nsurv = sh->h_stats.kept;
    // how many survivor records go back to each source
    sh->launches += launch_owner_survivor_counts(s, sh->sv_idx.p, sh->nsurv, sh->world, sh->d_aux.p, sh->d_aux.p + 195, sh->sms);
    unsigned long long cnt[64];
    CUDA_TRY(cudaMemcpyAsync(cnt, sh->d_aux.p + 195, sh->world * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    unsigned long long obase[65];
    uint64_t acc = 0;
    for (uint32_t r = 0; r < sh->world; ++r) {
        surv_counts[r]          = cnt[r];
        sh->surv_out_counts[r]  = cnt[r];
        obase[r]                = acc;
        acc += cnt[r];
    }
    CUDA_TRY(cudaMemcpyAsync(sh->d_aux.p + 65, obase, sh->world * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
}

// the survivor records, grouped by source rank: 8 bytes each {index inside the source's group, global count}
extern "C" int colibri_b200_shard_level_owner_survivors(colibri_b200_shard* sh, void* dev_out) {
    if (!sh || (!dev_out && sh->nsurv)) return set_err(COLIBRI_E_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(sh->dev));
    PhaseClock clk(sh, 4);
    sh->launches += launch_owner_survivors(sh->s, sh->sv_idx.p, sh->sv_cnt.p, sh->nsurv, sh->world, sh->d_aux.p, sh->d_aux.p + 65, sh->d_aux.p + 130, dev_out, sh->sms);
    CUDA_TRY(cudaStreamSynchronize(sh->s));
    return 0;
}

// sender side: dev_reply_back = nsent u32 in send order; dev_surv = survivor records grouped by owner rank (surv_counts[world]).
extern "C" int colibri_b200_shard_level_finish(colibri_b200_shard* sh, const void* dev_reply_back, const void* dev_surv, const uint64_t* surv_counts, uint64_t* local_valid) {
    if (!sh || !surv_counts || (sh->nsent && !dev_reply_back)) return set_err(COLIBRI_E_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(sh->dev));
    PhaseClock clk(sh, 5);
    cudaStream_t s = sh->s;
    const int n = sh->level + 1;
    TRY(shard_zero_stats(sh));
    CUDA_TRY(cudaMemsetAsync(sh->cur.p + sh->npos, 0, 8 * sizeof(uint32_t), s));
    if (sh->list_valid) CUDA_TRY(cudaMemsetAsync(sh->cur.p, 0, sh->npos * sizeof(uint32_t), s));  // list mode writes only the positions that keep an id
    sh->launches += launch_sender_relabel(s, sh->rec_of_pos.p, (const uint32_t*)dev_reply_back, shard_items(sh), sh->cur.p, sh->d_stats.p, sh->sms,
                                          shard_dense_now(sh) ? sh->dense_cnt : nullptr, sh->t, shard_list(sh));
    uint64_t total = 0;
    for (uint32_t g = 0; g < sh->world; ++g) total += surv_counts[g];
    Segment sg;
    sg.n = n;
    const uint64_t nd = shard_dense_now(sh) ? sh->dense_nsurv : 0;  // this rank's share of the dense square's survivors goes into the same segment
    if (total + nd) {
        if (total && !dev_surv) return set_err(COLIBRI_E_INVALID, "NULL survivor buffer");
        TRY(sg.pos.alloc(sh->dev, total + nd));
        TRY(sg.cnt.alloc(sh->dev, total + nd));
        uint64_t off = 0;
        for (uint32_t g = 0; g < sh->world; ++g) {
            sh->launches += launch_sender_survivors(s, (const uint8_t*)dev_surv + off * 8, surv_counts[g], sh->pos_of_rec.p, sh->send_base[g], sg.pos.p + off, sg.cnt.p + off);
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

// ---------------------------------------------------------------------------------------------------------------------
// Skipgrams of the level just finished (n = current level >= 3; exhaustive mode).  Every valid window of level n sends one
// 16-byte key per gap mask to the key's owner; owners count, apply the skipgram threshold and return the survivors to the
// rank whose record claimed the slot.  Nothing else comes back: skipgrams feed no later level.
extern "C" int colibri_b200_shard_skip_split_count(colibri_b200_shard* sh, uint64_t* send_counts, uint64_t* nrecords) {
    if (!sh || !send_counts || !nrecords) return set_err(COLIBRI_E_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(sh->dev));
    PhaseClock clk(sh, 2);
    cudaStream_t s = sh->s;
    const int    n = sh->level;
    *nrecords      = 0;
    for (uint32_t d = 0; d < sh->world; ++d) send_counts[d] = 0;
    sh->sk_nmasks = 0;
    sh->sk_nsent  = 0;
    if (!sh->o.DOSKIPGRAMS_EXHAUSTIVE || n < 3) return 0;
    std::vector<SkipMask> masks;
    TRY(skip_masks(n, sh->o.MAXSKIPS, masks));
    for (auto& m : masks)
        if (m.nparts > 3) return set_err(COLIBRI_E_UNSUPPORTED, "skipgrams with more than 3 non-gap runs (n >= 7) are not on the multi-GPU path");
    if (masks.empty()) return 0;
    sh->sk_nmasks = (int)masks.size();
    std::vector<const uint32_t*> ptrs(n, nullptr);
    for (int k = 1; k < n; ++k) ptrs[k] = sh->ids_keep[k].p;
    TRY(sh->d_idptrs.alloc(sh->dev, n));
    TRY(sh->d_masks.alloc(sh->dev, masks.size()));
    if (sh->d_aux.n < 260) TRY(sh->d_aux.alloc(sh->dev, 260));
    CUDA_TRY(cudaMemcpyAsync(sh->d_idptrs.p, ptrs.data(), n * sizeof(uint32_t*), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(sh->d_masks.p, masks.data(), masks.size() * sizeof(SkipMask), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemsetAsync(sh->d_aux.p, 0, 260 * sizeof(unsigned long long), s));
    sh->launches += launch_skip_split_count(s, sh->d_idptrs.p, n, sh->d_masks.p, sh->sk_nmasks, sh->npos, sh->world, sh->d_aux.p + 195, sh->sms);
    unsigned long long cnt[64];
    CUDA_TRY(cudaMemcpyAsync(cnt, sh->d_aux.p + 195, sh->world * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));  // also keeps ptrs/masks alive until the copies are done
    unsigned long long bases[65];
    uint64_t acc = 0;
    for (uint32_t d = 0; d < sh->world; ++d) {
        send_counts[d]      = cnt[d];
        bases[d]            = acc;
        sh->sk_send_base[d] = acc;
        acc += cnt[d];
    }
    sh->sk_send_base[sh->world] = acc;
    sh->sk_nsent                = acc;
    *nrecords                   = acc;
    CUDA_TRY(cudaMemcpyAsync(sh->d_aux.p + 65, bases, sh->world * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
}
extern "C" int colibri_b200_shard_skip_split_write(colibri_b200_shard* sh, void* dev_send) {
    if (!sh || (!dev_send && sh->sk_nsent)) return set_err(COLIBRI_E_INVALID, "NULL argument");
    if (!sh->sk_nsent) return 0;
    CUDA_TRY(cudaSetDevice(sh->dev));
    PhaseClock clk(sh, 3);
    TRY(sh->sk_pos_of_rec.alloc(sh->dev, sh->sk_nsent));
    sh->launches += launch_skip_split_write(sh->s, sh->d_idptrs.p, sh->level, sh->d_masks.p, sh->sk_nmasks, sh->npos, sh->world, sh->d_aux.p + 65, sh->d_aux.p + 130, dev_send,
                                            sh->sk_pos_of_rec.p, sh->sms);
    CUDA_TRY(cudaStreamSynchronize(sh->s));
    return 0;
}
// stats = {distinct skipgrams owned, kept}; surv_counts[world]
extern "C" int colibri_b200_shard_skip_owner(colibri_b200_shard* sh, const void* dev_recv, const uint64_t* recv_counts, uint64_t stats[2], uint64_t* surv_counts) {
    if (!sh || !recv_counts || !stats || !surv_counts) return set_err(COLIBRI_E_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(sh->dev));
    PhaseClock clk(sh, 4);
    cudaStream_t s = sh->s;
    unsigned long long src_base[65];
    uint64_t nrecv = 0;
    for (uint32_t r = 0; r < sh->world; ++r) {
        src_base[r] = nrecv;
        nrecv += recv_counts[r];
        surv_counts[r] = 0;
    }
    src_base[sh->world] = nrecv;
    stats[0] = stats[1] = 0;
    sh->sk_nsurv = 0;
    if (nrecv == 0) return 0;
    if (!dev_recv) return set_err(COLIBRI_E_INVALID, "NULL buffer");
    if (nrecv >= 0xFFFFFFF0ull) return set_err(COLIBRI_E_CAPACITY, "owner received %llu skipgram records", (unsigned long long)nrecv);
    if (sh->d_aux.n < 260) TRY(sh->d_aux.alloc(sh->dev, 260));
    CUDA_TRY(cudaMemsetAsync(sh->d_aux.p, 0, 260 * sizeof(unsigned long long), s));
    CUDA_TRY(cudaMemcpyAsync(sh->d_aux.p, src_base, (sh->world + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
    // PatternModel::pruneskipgrams returns early when minskiptypes <= 1 (reference :2170-2171): then only MINTOKENS applies
    const uint32_t ts = sh->o.MINSKIPTYPES > 1 ? (uint32_t)sh->o.MINTOKENS_SKIPGRAMS : sh->t;
    uint64_t cap = std::max<uint64_t>(64, nrecv + nrecv / 2 + 16);
    if (sh->sktable.n < cap) TRY(sh->sktable.alloc(sh->dev, cap));
    CUDA_TRY(cudaMemsetAsync(sh->sktable.p, 0, cap * sizeof(SkipSlot), s));
    TRY(shard_zero_stats(sh));
    sh->launches += launch_skip_stream_count(s, dev_recv, nrecv, sh->sktable.p, cap, sh->d_stats.p, sh->sms);
    const uint64_t bound = nrecv / std::max<uint32_t>(ts, 1) + 1;
    if (sh->sk_sv_idx.n < bound) TRY(sh->sk_sv_idx.alloc(sh->dev, bound));
    if (sh->sk_sv_cnt.n < bound) TRY(sh->sk_sv_cnt.alloc(sh->dev, bound));
    if (sh->sk_sv_mask.n < bound) TRY(sh->sk_sv_mask.alloc(sh->dev, bound));
    sh->launches += launch_prune_skipgrams(s, sh->sktable.p, cap, ts, sh->sk_sv_idx.p, sh->sk_sv_cnt.p, sh->sk_sv_mask.p, sh->d_stats.p, sh->sms);
    TRY(shard_read_stats(sh));
    stats[0]     = sh->h_stats.found;
    stats[1]     = sh->h_stats.kept;
    sh->sk_nsurv = sh->h_stats.kept;
    sh->launches += launch_owner_survivor_counts(s, sh->sk_sv_idx.p, sh->sk_nsurv, sh->world, sh->d_aux.p, sh->d_aux.p + 195, sh->sms);
    unsigned long long cnt[64];
    CUDA_TRY(cudaMemcpyAsync(cnt, sh->d_aux.p + 195, sh->world * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    unsigned long long obase[65];
    uint64_t acc = 0;
    for (uint32_t r = 0; r < sh->world; ++r) {
        surv_counts[r] = cnt[r];
        obase[r]       = acc;
        acc += cnt[r];
    }
    CUDA_TRY(cudaMemcpyAsync(sh->d_aux.p + 65, obase, sh->world * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
}
extern "C" int colibri_b200_shard_skip_owner_survivors(colibri_b200_shard* sh, void* dev_out) {
    if (!sh || (!dev_out && sh->sk_nsurv)) return set_err(COLIBRI_E_INVALID, "NULL argument");
    if (!sh->sk_nsurv) return 0;
    CUDA_TRY(cudaSetDevice(sh->dev));
    PhaseClock clk(sh, 4);
    sh->launches += launch_skip_owner_survivors(sh->s, sh->sk_sv_idx.p, sh->sk_sv_cnt.p, sh->sk_sv_mask.p, sh->sk_nsurv, sh->world, sh->d_aux.p, sh->d_aux.p + 65, sh->d_aux.p + 130, dev_out,
                                                sh->sms);
    CUDA_TRY(cudaStreamSynchronize(sh->s));
    return 0;
}
extern "C" int colibri_b200_shard_skip_finish(colibri_b200_shard* sh, const void* dev_surv, const uint64_t* surv_counts) {
    if (!sh || !surv_counts) return set_err(COLIBRI_E_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(sh->dev));
    PhaseClock clk(sh, 5);
    uint64_t total = 0;
    for (uint32_t g = 0; g < sh->world; ++g) total += surv_counts[g];
    if (!total) return 0;
    if (!dev_surv) return set_err(COLIBRI_E_INVALID, "NULL survivor buffer");
    Segment sg;
    sg.n    = sh->level;
    sg.skip = true;
    TRY(sg.pos.alloc(sh->dev, total));
    TRY(sg.cnt.alloc(sh->dev, total));
    TRY(sg.mask.alloc(sh->dev, total));
    uint64_t off = 0;
    for (uint32_t g = 0; g < sh->world; ++g) {
        sh->launches += launch_skip_sender_survivors(sh->s, (const uint8_t*)dev_surv + off * 16, surv_counts[g], sh->sk_pos_of_rec.p, sh->sk_send_base[g], sg.pos.p + off, sg.cnt.p + off,
                                                     sg.mask.p + off);
        off += surv_counts[g];
    }
    CUDA_TRY(cudaStreamSynchronize(sh->s));
    sg.count = total;
    sh->segs.push_back(std::move(sg));
    return 0;
}

// export this rank's share of the model.  passes = npasses x {n, found, foundskip, pruned} (global numbers, reduced by the caller)
extern "C" int colibri_b200_shard_finish(colibri_b200_shard* sh, const uint64_t* passes, int npasses, uint64_t global_types, int maxn, int minn, colibri_b200_model** out) {
    if (!sh || !out) return set_err(COLIBRI_E_INVALID, "NULL argument");
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(sh->dev));
    auto* m       = new colibri_b200_model();
    m->device     = sh->dev;
    m->model_type = sh->o.model_type;
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete m;
        return set_err(COLIBRI_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    m->totaltokens = sh->global_tokens;
    m->totaltypes  = global_types;
    m->maxn        = maxn;
    m->minn        = minn;
    for (int p = 0; p < npasses; ++p) {
        m->passes.push_back({passes[4 * p], passes[4 * p + 1], passes[4 * p + 2], passes[4 * p + 3]});
        if (passes[4 * p + 2]) m->hasskipgrams = 1;
    }
    {   // which levels end up in the model: the rules of the single-GPU driver (engine.cu, reference include/patternmodel.h:1221-1229, :1278-1280, :1337-1341)
        const colibri_b200_options& o = sh->o;
        const int last_pass = maxn;
        std::vector<Segment> keep;
        for (auto& sg : sh->segs) {
            bool drop = false;
            if (o.MINTOKENS > 1) {
                if (!o.DOSKIPGRAMS_EXHAUSTIVE && !o.DOSKIPGRAMS) {
                    const int k = sg.n;
                    if (k < o.MINLENGTH && last_pass >= k + 1 && k != o.MAXBACKOFFLENGTH && !(k == 1 && o.MINTOKENS_UNIGRAMS > o.MINTOKENS)) drop = true;
                    if (k == o.MAXBACKOFFLENGTH && o.MAXBACKOFFLENGTH < o.MINLENGTH) drop = true;
                } else if (o.MINLENGTH > 1 && sg.n <= o.MINLENGTH - 1) {
                    drop = true;
                }
            }
            if (!drop && sg.count > 0) keep.push_back(std::move(sg));
        }
        sh->segs = std::move(keep);
    }
    int rc;
    {
        PhaseClock clk(sh, 6);
        rc = export_segments(sh->dev, sh->s, sh->segs, sh->tok.p, m, sh->launches);
        cudaStreamSynchronize(sh->s);
    }
    if (rc) {
        colibri_b200_model_free(m);
        return rc;
    }
    m->counters[0]          = sh->npos - 1;
    m->counters[1]          = sh->corpus->nbytes;
    m->counters[2]          = sh->launches;
    m->ms[COLIBRI_T_TOTAL]  = sh->device_ms;
    *out = m;
    return 0;
}
