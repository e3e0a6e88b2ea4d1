// kernels.h -- launch wrappers of the sm_100a kernels (definitions in kernels.cu).
// Every wrapper enqueues on the given stream and returns the number of kernel launches it made.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace colibri {

// One slot of the n-gram table: 16 bytes, two per 32-byte HBM sector.
//   key   = (id of the surviving (n-1)-gram at p) << 32 | (id of the surviving (n-1)-gram at p+1); 0 = empty.
//   count = occurrences, pos = a position where the n-gram occurs (any one; used to print its bytes).
struct alignas(16) NgramSlot {
    unsigned long long key;
    uint32_t           count;
    uint32_t           pos;
};
// One slot of the skipgram table: 32 bytes = one sector.  The 128-bit key is claimed with one atom.cas.b128.
//   k0 = maskfield << 32 | id(part 1),  k1 = id(part 2) << 32 | id(part 3)   (ids of the contiguous non-gap runs).
//   Masks with more than three runs first fold their leading three ids into one through a helper entry of the
//   same table (flag kSkipCombiner; its slot index is the folded id), repeatedly, until three ids remain.
struct alignas(32) SkipSlot {
    unsigned long long k0, k1;
    uint32_t           count;
    uint32_t           pos;
    uint32_t           pad[2];
};

// per-train device-side statistics block (zeroed by the host before each phase that uses it)
struct DeviceStats {
    unsigned long long totaltokens;    // non-delimiter positions
    unsigned long long found;          // occupied slots of the table just scanned
    unsigned long long kept;           // ... with count >= threshold
    unsigned long long kept_occ;       // sum of their counts
    unsigned long long cursor;         // compaction cursor into the survivor arrays
    unsigned long long valid_windows;  // upserts issued by the last count kernel
    unsigned long long probes;         // probe steps of the last count kernel (diagnostic)
    unsigned long long singletons;     // windows the occurrence filter proved to be the only one of their n-gram
    unsigned int       maxclass;
    unsigned int       errflags;       // kErr* bits
};
constexpr unsigned kErrTokenTooLong  = 1u;  // varint longer than 5 bytes / class >= 2^32
constexpr unsigned kErrNonCanonical  = 2u;  // multi-byte token whose last byte is 0
constexpr unsigned kErrReservedClass = 4u;  // class 3 (skip) or 4 (flex) in running text
constexpr unsigned kErrTableFull     = 8u;  // a probe sequence wrapped the whole table
constexpr unsigned kErrLongSentence  = 16u; // indexed model: a token offset does not fit IndexReference's uint16_t

// gap configuration of one skipgram mask, precomputed on the host (compute_skip_configurations, src/algorithms.cpp:79-94)
constexpr int kMaxSkipParts = 12;
struct SkipMask {
    uint32_t mask;
    uint32_t nparts;                 // number of contiguous non-gap runs (2 .. kMaxSkipParts)
    uint8_t  start[kMaxSkipParts];   // first token of each run, relative to the window
    uint8_t  len[kMaxSkipParts];     // tokens in each run
};
constexpr int      kMaxSkipMasks   = 4096;
constexpr uint32_t kSkipCombiner   = 0x80000000u;  // mask-field flag: an id-combining helper entry, not a pattern
constexpr uint32_t kSkipRoundShift = 24;           // mask-field bits 24..30: number of combining rounds behind the key

// ---- K0: corpus bytes -> class ids (0 = sentence delimiter)
constexpr int kTokTile = 4096;  // bytes per block
int launch_tokenise_count(cudaStream_t s, const uint8_t* corpus, uint64_t nbytes, uint32_t* blk_counts, uint32_t nblocks);
int launch_scan_block_counts(cudaStream_t s, uint32_t* blk_counts, uint32_t nblocks, unsigned long long* total, bool accumulate = false /* start at *total */);
int launch_tokenise_write(cudaStream_t s, const uint8_t* corpus, uint64_t nbytes, const uint32_t* blk_offsets, uint32_t nblocks, uint32_t* tok, DeviceStats* st);

// ---- K1: unigram histogram, unigram prune, level-1 ids
int launch_unigram_hist(cudaStream_t s, const uint32_t* tok, uint64_t npos, uint32_t* count1, uint32_t nclasses, int sms);
int launch_unigram_prune(cudaStream_t s, const uint32_t* count1, uint32_t nclasses, uint32_t threshold, uint32_t* sv_pos, uint32_t* sv_count, uint64_t sv_base, DeviceStats* st,
                         uint32_t part_mod = 1, uint32_t part_rem = 0, uint32_t* class_index = nullptr);
int launch_make_id1(cudaStream_t s, const uint32_t* tok, uint64_t npos, const uint32_t* count1, uint32_t threshold, uint32_t* id1);

// ---- K2: n-gram upsert (the dominant kernel), K3: prune/compact, relabel
// occurrence filter: nbuckets (power of two) 2-bit saturating counters, nbuckets/4 bytes, zeroed by the caller; st->found := buckets hit twice
// dense (level 2 only): windows whose two class ids are both below it bypass the filter (they own directly addressed slots, see launch_count_ngrams)
// list != NULL: the windows are the nlist positions list[j] (positions whose (n-1)-gram survived) instead of every position 0..npos
int launch_ngram_filter(cudaStream_t s, const uint32_t* prev, uint64_t npos, uint32_t* filter, uint64_t nbuckets, DeviceStats* st, int sms, uint32_t dense = 0,
                        const uint32_t* list = nullptr, uint64_t nlist = 0);
// filter == NULL: every valid window goes to the table.  dense > 0: the table has cap + dense * dense slots; a window (a, b) with a, b < dense is
// counted in slot cap + a * dense + b (no hash, no filter, no probing), every other window in the hashed part [0, cap)
int launch_count_ngrams(cudaStream_t s, const uint32_t* prev, uint32_t* cur, uint64_t npos, NgramSlot* table, uint64_t cap, DeviceStats* st, int sms,
                        const uint32_t* filter = nullptr, uint64_t nbuckets = 0, bool hot = false /* per-block shared-memory cache for frequent keys */, uint32_t dense = 0,
                        const uint32_t* list = nullptr /* list mode: cur must have been zeroed by the caller */, uint64_t nlist = 0,
                        uint32_t* dense_cnt = nullptr /* dense > 0: the zeroed u32 square dense x dense */,
                        bool onebit = false /* filter = the packed "hit twice" bits of launch_filter_to_bitmap instead of the 2-bit counters */);
// bitmap[nbuckets / 32 words] = the "hit twice" bit of every bucket of the occurrence filter (nbuckets >= 32, a power of two)
int launch_filter_to_bitmap(cudaStream_t s, const uint32_t* filter, uint64_t nbuckets, uint32_t* bitmap);
// prune() over the dense square (cells = ids cap + 1 .. cap + dense^2); cap must be a multiple of 32.  Survivors get tok_ext[2 * cell ..] = their two class
// ids and the position ext_pos0 + 2 * cell (tok_ext = token array + ext_pos0)
int launch_prune_dense(cudaStream_t s, const uint32_t* dense_cnt, uint32_t dense, uint64_t cap, uint32_t threshold, uint32_t* sv_pos, uint32_t* sv_count, uint32_t* bitmap,
                       uint32_t* slot_index, uint32_t* tok_ext, uint32_t ext_pos0, DeviceStats* st, int sms);
// bitmap: (cap+31)/32 words, bit = slot survived (may be NULL)
int launch_prune_ngrams(cudaStream_t s, const NgramSlot* table, uint64_t cap, uint32_t threshold, uint32_t* sv_pos, uint32_t* sv_count, uint32_t* bitmap, DeviceStats* st, int sms,
                        uint32_t* slot_index = nullptr /* slot -> survivor index + 1, for the forward index */);
// cur[p] := 0 where the n-gram of p was pruned.  list_in != NULL: only the nlist_in positions list_in[j] are visited (everything else in cur
// is zero already).  list_out != NULL: the surviving positions are appended to it through *cursor (zeroed by the caller; rough corpus order).
int launch_relabel(cudaStream_t s, uint32_t* cur, uint64_t npos, const uint32_t* bitmap, const uint32_t* list_in = nullptr, uint64_t nlist_in = 0, uint32_t* list_out = nullptr,
                   unsigned long long* cursor = nullptr, int sms = 148);

// ---- partitioned counting path of a level (partition.cu): radix partition of the windows by key hash, counting in shared memory
struct PartPlan {
    int      b1 = 4, b2 = 4;  // bits of the first / second split (<= 11 each)
    uint32_t nparts = 256;    // 1 << (b1 + b2)
};
PartPlan part_plan(uint64_t window_bound);  // <= 256 windows per partition on average
// pass A: hist1[b1 bits of the key hash] += 1 per hashed window (hist1 zeroed by the caller, 1 << b1 entries); level 2 (dense > 0): windows of two classes
// below `dense` are counted in dense_cnt[a * dense + b] (zeroed) instead; st->valid_windows += valid windows
int launch_part_hist(cudaStream_t s, const uint32_t* prev, const uint32_t* list, uint64_t nitems, uint32_t dense, uint32_t* dense_cnt, uint32_t* hist1, const PartPlan& pl,
                     DeviceStats* st, int sms);
// make_id1 + pass A of a dense level 2 in one sweep over the tokens: id1[0 .. npos] as launch_make_id1 writes them, and hist1 / dense_cnt / st->valid_windows
// as launch_part_hist would leave them for level 2 (hist1, dense_cnt zeroed by the caller; dense > 0).  keep_bits: scratch of (nclasses + 255) / 256 * 8 words
// (one bit per class: does it stay at level 1); every token is a class below nclasses
int launch_make_id1_hist(cudaStream_t s, const uint32_t* tok, uint64_t npos, const uint32_t* count1, uint32_t nclasses, uint32_t threshold, uint32_t* keep_bits, uint32_t* id1,
                         uint32_t dense, uint32_t* dense_cnt, uint32_t* hist1, const PartPlan& pl, DeviceStats* st, int sms);
// out[0..n] = exclusive scan of counts[0..n) (+ *base_ptr if given), out2 (may be NULL) = a second copy of out[0..n); n <= 2048 is cheap, larger n works
int launch_part_bases(cudaStream_t s, const uint32_t* counts, uint32_t n, uint32_t* out, uint32_t* out2, const unsigned long long* base_ptr);
// pass B: records (key, position) of the hashed windows grouped by their b1 bits (cursor1 = a copy of the partition offsets, advanced by the kernel);
// dense mode (list == NULL) also writes cur[p] = dense_id[a * dense + b] or 0 for every position (list mode: cur zeroed by the caller)
int launch_part_split1(cudaStream_t s, const uint32_t* prev, const uint32_t* list, uint64_t nitems, uint32_t dense, const uint32_t* dense_id, uint32_t* cur, const PartPlan& pl,
                       uint32_t* cursor1, uint32_t* hist2 /* zeroed, nparts: += records per final partition */, void* rk /* u64 */, uint32_t* rp);
// off[0 .. nparts] = exclusive scan of counts[0 .. nparts) (+ *base_ptr if given); scratch: group_tot 1 << b1, group_base (1 << b1) + 1
int launch_part_scan(cudaStream_t s, const uint32_t* counts, const PartPlan& pl, uint32_t* group_tot, uint32_t* group_base, uint32_t* off, const unsigned long long* base_ptr);
// pass D: every b1-partition split by the next b2 bits; off[nparts + 1] = the final partition offsets
int launch_part_split2(cudaStream_t s, const void* rk_in, const uint32_t* rp_in, const uint32_t* off, const PartPlan& pl, void* rk_out, uint32_t* rp_out);
// pass E: per partition count in shared memory and threshold; survivors of partition q -> tmp_pos / tmp_cnt [off[q] ..], kept_of[q] of them;
// out[rp[i]] = (id_base + index in tmp) * id_mul + id_add for the records of surviving keys; st->found / kept / kept_occ / singletons (keys counted
// once) are added to; kErrTableFull if a partition does not fit its table
int launch_part_count(cudaStream_t s, const void* rk, const uint32_t* rp, const uint32_t* off, const PartPlan& pl, uint32_t threshold, uint32_t* out, uint32_t id_base,
                      uint32_t id_mul, uint32_t id_add, uint32_t* tmp_pos, uint32_t* tmp_cnt, uint32_t* kept_of, DeviceStats* st, int sms);
// survivors out of the partitions' ranges -> sv_pos / sv_cnt [*first ..], compacted (scratch: group_tot 1 << b1, group_base (1 << b1) + 1, dst_off nparts + 1);
// slot_index (may be NULL): slot_index[id_base + index in tmp] = survivor index + 1
int launch_part_gather(cudaStream_t s, const uint32_t* tmp_pos, const uint32_t* tmp_cnt, const uint32_t* off, const uint32_t* kept_of, const PartPlan& pl, uint32_t* group_tot,
                       uint32_t* group_base, uint32_t* dst_off, const unsigned long long* first, uint32_t* sv_pos, uint32_t* sv_cnt, uint32_t* slot_index, uint32_t id_base);
int launch_compact_nonzero(cudaStream_t s, const uint32_t* cur, uint64_t npos, uint32_t* list_out, unsigned long long* cursor /* zeroed */);
int launch_iota_plus1(cudaStream_t s, uint32_t* out, uint64_t n);

// ---- skipgrams (config 3)
// occ_pos != NULL: npos counts entries of occ_pos (explicit window positions) instead of corpus positions; item_slot (optional): slot + 1 per (window, mask)
int launch_count_skipgrams(cudaStream_t s, const uint32_t* const* ids /*device array, index = level*/, int n, const SkipMask* masks /*device*/, int nmasks, uint64_t npos,
                           SkipSlot* table, uint64_t cap, DeviceStats* st, int sms, const uint32_t* occ_pos = nullptr, uint32_t* item_slot = nullptr);
int launch_skip_types(cudaStream_t s, const uint32_t* const* ids, int n, const SkipMask* masks, int nmasks, uint64_t nocc, const uint32_t* occ_pos, const uint32_t* item_slot,
                      NgramSlot* pairs /* zeroed scratch */, uint64_t cap, uint32_t* types /* zeroed, one per skip slot */, DeviceStats* st, int sms);
int launch_prune_skipgrams(cudaStream_t s, const SkipSlot* table, uint64_t cap, uint32_t threshold, uint32_t* sv_pos, uint32_t* sv_count, uint32_t* sv_mask, DeviceStats* st,
                           int sms, uint32_t* slot_index = nullptr, const uint32_t* types = nullptr, uint32_t mintypes = 0);

// ---- export: survivors -> pattern bytes
// sv_nm[i] = n | mask << 8 ; for n == 1 sv_pos holds the class id itself
int launch_fill_u32(cudaStream_t s, uint32_t* dst, uint64_t n, uint32_t value);
// nbytes (multiple of 4, 4-byte aligned) from device memory to MAPPED pinned host memory with a one-warp kernel instead of the copy engine
int launch_copy_words_to_host(cudaStream_t s, const void* src, void* dst_mapped, uint32_t nbytes);
int launch_sum_u32(cudaStream_t s, const uint32_t* v, uint64_t n, unsigned long long* total /* += sum */);
int launch_pack_nm(cudaStream_t s, uint32_t* nm /*in: gap masks, out: n | mask << 8*/, uint64_t count, uint32_t n);
int launch_export_lengths(cudaStream_t s, const uint32_t* tok, const uint32_t* sv_pos, const uint32_t* sv_nm, uint64_t n, uint32_t* lens, uint16_t* lens16);
int launch_exclusive_scan_u32_u64(cudaStream_t s, const uint32_t* in, uint64_t* out /*n+1*/, uint64_t n, uint64_t* tmp /*>= n/2048+2*/);
int launch_export_write(cudaStream_t s, const uint32_t* tok, const uint32_t* sv_pos, const uint32_t* sv_nm, const uint64_t* off, uint64_t n, uint8_t* keys);

// ---- forward index of indexed models (index.cu)
int launch_delim_flags(cudaStream_t s, const uint32_t* tok, uint64_t npos, uint32_t* flags);
int launch_sent_start(cudaStream_t s, const uint32_t* tok, const uint64_t* sent_before, uint64_t npos, uint32_t* sent_start);
// pos_lookup != NULL: item i stands for position pos_lookup[i / pos_div] (skipgram items of indexed models)
int launch_pair_count(cudaStream_t s, const uint32_t* ids, const uint32_t* map, uint64_t npos, bool by_class, uint32_t* blk_counts /* one per 2048 positions */);
int launch_pair_write(cudaStream_t s, const uint32_t* ids, const uint32_t* map, uint64_t npos, bool by_class, const uint64_t* blk_off, uint32_t* keys, uint32_t* vals,
                      const uint32_t* pos_lookup = nullptr, uint32_t pos_div = 1);
int launch_radix_pass(cudaStream_t s, const uint32_t* keys_in, const uint32_t* vals_in, uint64_t n, int shift, uint32_t* hist /* 256*ceil(n/4096) */, uint64_t* hist_off,
                      uint64_t* scan_tmp, uint32_t* keys_out, uint32_t* vals_out);
int launch_refs_from_positions(cudaStream_t s, const uint32_t* pos, uint64_t n, const uint64_t* sent_before, const uint32_t* sent_start, uint32_t* ref_sentence, uint16_t* ref_token,
                               DeviceStats* st);

// ---- multi-GPU phases (hash-partitioned model): shard_kernels.cu, driven by shard.cu
// dense > 0 (level 2): windows of two classes below `dense` are counted in dense_cnt[a * dense + b] and get no destination
int launch_split_count(cudaStream_t s, const uint32_t* prev, uint64_t npos, uint32_t world, uint32_t* hist /* world x ceil(npos/4096), destination-major */, uint32_t dense = 0,
                       uint32_t* dense_cnt = nullptr, const uint32_t* list = nullptr /* list mode: npos items, item j = position list[j] */);
// peer_* != NULL selects the NVLink peer-store variants (symmetric-memory receive slots of slot_cap entries per source rank)
int launch_split_write(cudaStream_t s, const uint32_t* prev, uint64_t npos, uint32_t world, const uint64_t* hist_off, void* send_keys, uint32_t* pos_of_rec, uint32_t* rec_of_pos,
                       void* const* peer_keys = nullptr, uint32_t my_rank = 0, uint64_t slot_cap = 0, uint32_t dense = 0 /* rec_of_pos = 0x80000000 | cell for dense windows */,
                       const uint32_t* list = nullptr /* list mode: rec_of_pos is indexed by item */);
// n = keys received; slot_cap != 0: they sit in `world` slots of slot_cap entries, slot_counts[r] of them in slot r (rid[] is indexed like the buffer)
int launch_stream_filter(cudaStream_t s, const void* keys, uint64_t n, uint32_t* filter, uint64_t nbuckets, DeviceStats* st, int sms, uint64_t slot_cap = 0,
                         const unsigned long long* slot_counts = nullptr, uint32_t world = 0);
int launch_stream_count(cudaStream_t s, const void* keys, uint64_t n, NgramSlot* table, uint64_t cap, const uint32_t* filter, uint64_t nbuckets, uint32_t* rid, DeviceStats* st, int sms,
                        uint64_t slot_cap = 0, const unsigned long long* slot_counts = nullptr, uint32_t world = 0, bool onebit = false /* filter = packed "hit twice" bits */);
int launch_owner_reply(cudaStream_t s, uint32_t* rid, uint64_t n, const uint32_t* bitmap, uint32_t world, uint32_t rank, void* const* peer_reply = nullptr, uint64_t slot_cap = 0,
                       const unsigned long long* slot_counts = nullptr, uint32_t id_off = 0 /* global ids start above the dense square's */);
int launch_owner_survivors_p2p(cudaStream_t s, const uint32_t* sv_idx, const uint32_t* sv_count, uint64_t n, uint32_t world, uint32_t rank, uint64_t slot_cap, uint64_t surv_cap,
                               unsigned long long* cursors, void* const* peer_surv, DeviceStats* st, int sms);
int launch_p2p_publish(cudaStream_t s, void* const* peer_hdr, uint32_t world, uint64_t offset, const unsigned long long* vals /* world x nvals */, uint32_t nvals);
int launch_owner_survivor_counts(cudaStream_t s, const uint32_t* sv_idx, uint64_t n, uint32_t world, const unsigned long long* src_base, unsigned long long* counts, int sms);
int launch_owner_survivors(cudaStream_t s, const uint32_t* sv_idx, const uint32_t* sv_count, uint64_t n, uint32_t world, const unsigned long long* src_base,
                           const unsigned long long* out_base, unsigned long long* cursors, void* out, int sms);
int launch_sender_relabel(cudaStream_t s, const uint32_t* rec_of_pos, const uint32_t* reply, uint64_t npos, uint32_t* cur, DeviceStats* st, int sms,
                          const uint32_t* dense_global = nullptr, uint32_t threshold = 0, const uint32_t* list = nullptr /* list mode: cur zeroed by the caller */);
// this rank's share (cells = rank mod world) of the globally summed dense square: found / kept / kept_occ into st, survivors appended through st->cursor
int launch_dense_share(cudaStream_t s, const uint32_t* dense_global, uint32_t dense, uint32_t world, uint32_t rank, uint32_t threshold, uint32_t* sv_pos, uint32_t* sv_cnt,
                       uint32_t* tok_ext, uint32_t ext_pos0, DeviceStats* st, int sms);
int launch_sender_survivors(cudaStream_t s, const void* recs, uint64_t n, const uint32_t* pos_of_rec, uint64_t send_base, uint32_t* sv_pos, uint32_t* sv_count);

// skipgrams on the multi-GPU path
int launch_skip_split_count(cudaStream_t s, const uint32_t* const* ids, int n, const SkipMask* masks, int nmasks, uint64_t npos, uint32_t world, unsigned long long* dest_counts, int sms);
int launch_skip_split_write(cudaStream_t s, const uint32_t* const* ids, int n, const SkipMask* masks, int nmasks, uint64_t npos, uint32_t world, const unsigned long long* dest_base,
                            unsigned long long* cursors, void* send /* 16 B keys */, uint32_t* pos_of_rec, int sms);
int launch_skip_stream_count(cudaStream_t s, const void* recv, uint64_t n, SkipSlot* table, uint64_t cap, DeviceStats* st, int sms);
int launch_skip_owner_survivors(cudaStream_t s, const uint32_t* sv_idx, const uint32_t* sv_count, const uint32_t* sv_mask, uint64_t n, uint32_t world, const unsigned long long* src_base,
                                const unsigned long long* out_base, unsigned long long* cursors, void* out /* 16 B records */, int sms);
int launch_skip_sender_survivors(cudaStream_t s, const void* recs, uint64_t n, const uint32_t* pos_of_rec, uint64_t send_base, uint32_t* sv_pos, uint32_t* sv_count, uint32_t* sv_mask);

// ---- pattern sets addressed by Pattern::hash of their bytes (pattern_index.cu): constrained training, load filters, queries
constexpr uint32_t kMaxIndexedKeyBytes = 191;  // SpookyV2 "Short" range; longer patterns are refused
struct PatternMetaStats {  // zeroed by the host except minn / kept_minn = 0xFFFFFFFF
    unsigned int       maxn, minn, hasskip, hasflex, maxclass, malformed, duplicates;
    unsigned int       kept_maxn, kept_minn, kept_hasskip, kept_hasflex;
    unsigned int       unigram_ngrams;   // patterns of one token that are plain n-grams (totalwordtypesingroup(NGRAM, 1))
    unsigned long long nhist[256];       // patterns per length (lengths >= 255 share the last bin)
    unsigned long long kept_n[256];      // survivors per length
    unsigned long long kept_occ_n[256];  // their occurrences
};
int launch_pattern_meta(cudaStream_t s, const uint8_t* keys, const uint64_t* off, uint64_t np, uint16_t* pn, uint8_t* pcat, PatternMetaStats* st);
// One slot of the pattern index: 32 bytes = one HBM sector.  idx1 = pattern index + 1 (0 = empty); count = occurrences counted through
// this slot by constrained training (collected into counts[] and zeroed afterwards); k0, k1 = key bytes 0..15; k2 = key bytes 16..22 with
// the key length in the top byte, or 0xFF << 56 for keys longer than 23 bytes (their tail is compared in the blob).
struct alignas(32) PatSlot {
    uint32_t           idx1;
    uint32_t           count;
    unsigned long long k0, k1, k2;
};
// slots: cap_pow2 zeroed entries; presence: presence_bits_pow2 zeroed bits, one per hash bucket
// rep == NULL: equal keys are an error (st->duplicates); rep != NULL: group-by mode, rep[i] = index of the copy of key i that claimed the slot
int launch_index_build(cudaStream_t s, const uint8_t* keys, const uint64_t* off, uint64_t np, PatSlot* slots, uint64_t cap_pow2, uint32_t* presence,
                       uint64_t presence_bits_pow2, PatternMetaStats* st, uint32_t* rep = nullptr);
int launch_index_lookup(cudaStream_t s, const uint8_t* qkeys, const uint64_t* qoff, uint64_t nq, const uint8_t* keys, const uint64_t* off, const PatSlot* slots,
                        uint64_t cap_pow2, const uint32_t* presence, uint64_t presence_bits_pow2, uint32_t* out_idx1 /* pattern index + 1, or 0 */);
int launch_gather_counts(cudaStream_t s, const uint32_t* idx1, uint64_t nq, const uint32_t* counts, uint32_t* out);
// windows of n tokens that are in the set: counts[pattern] += 1, match[p] = pattern index + 1 or 0 (match may be NULL).
// prev (may be NULL) = match[] of length n-1 (npos + 1 readable entries); use_prefix / use_suffix: skip windows whose prefix / suffix did not match
int launch_constrained_match(cudaStream_t s, const uint32_t* tok, uint64_t npos, int n, const uint8_t* keys, const uint64_t* off, PatSlot* slots, uint64_t cap_pow2,
                             const uint32_t* presence, uint64_t presence_bits_pow2, uint32_t* counts, uint32_t* match, const uint32_t* prev, bool use_prefix, bool use_suffix,
                             DeviceStats* st, int sms);
// uni[class] = index + 1 of the unigram pattern of that class (uni zeroed by the caller, nclasses entries)
int launch_unigram_table(cudaStream_t s, const uint8_t* keys, const uint64_t* off, const uint16_t* pn, uint64_t np, uint32_t* uni, uint32_t nclasses);
// level 1 of a constrained run: counts[uni[c] - 1] += hist[c] for c < nclasses (hist = launch_unigram_hist's class histogram); match[p] = uni[tok[p]]
int launch_unigram_apply(cudaStream_t s, const uint32_t* hist, const uint32_t* uni, uint32_t nclasses, uint32_t* counts);
int launch_unigram_match(cudaStream_t s, const uint32_t* tok, uint64_t npos, const uint32_t* uni, uint32_t nclasses, uint32_t* match);
// per pattern length n (bins of 256, zeroed by the caller): patterns whose (n-1)-token prefix / suffix is not in the set
// counts[pattern] += the slot counters of a constrained run, which are reset
int launch_collect_slot_counts(cudaStream_t s, PatSlot* slots, uint64_t cap_pow2, uint32_t* counts);
int launch_closure_check(cudaStream_t s, const uint8_t* keys, const uint64_t* off, const uint16_t* pn, uint64_t np, const PatSlot* slots, uint64_t cap_pow2,
                         const uint32_t* presence, uint64_t presence_bits_pow2, unsigned long long* prefix_open, unsigned long long* suffix_open);
// per_length: also fill st->kept_n / kept_occ_n (indexed models build their occurrence lists per pattern length)
int launch_constrained_stats(cudaStream_t s, const uint32_t* counts, const uint16_t* pn, uint64_t np, uint32_t threshold, uint32_t* flags, PatternMetaStats* st, DeviceStats* ds,
                             bool per_length);
int launch_load_filter(cudaStream_t s, const uint16_t* pn, const uint8_t* pcat, const uint32_t* counts, const uint32_t* constrain_idx1, uint64_t np, uint32_t mintokens,
                       uint32_t minlength, uint32_t maxlength, int dongrams, int doskipgrams, int doflexgrams, uint32_t* flags, PatternMetaStats* st);
int launch_select_scatter(cudaStream_t s, const uint32_t* flags, const uint64_t* newpos, const uint16_t* pn, uint64_t np, uint32_t* sel_idx, uint32_t* sel_n);
int launch_gather_meta(cudaStream_t s, const uint32_t* sel_idx, uint64_t k, const uint64_t* off, const uint32_t* counts, uint32_t* kmap, uint32_t* lens, uint16_t* len16,
                       uint32_t* counts_out);
int launch_gather_keys(cudaStream_t s, const uint32_t* sel_idx, uint64_t k, const uint8_t* keys, const uint64_t* off, const uint64_t* new_off, uint8_t* out);
int launch_gather_refs(cudaStream_t s, const uint32_t* sel_idx, uint64_t k, const uint64_t* ref_off, const uint32_t* rs, const uint16_t* rt, const uint64_t* new_ref_off,
                       uint32_t* rs_out, uint16_t* rt_out);
int launch_iota(cudaStream_t s, uint32_t* out, uint64_t n);
int launch_token_bitmap(cudaStream_t s, const uint8_t* keys, const uint64_t* off, uint64_t np, uint32_t* bitmap /* zeroed, (maxclass >> 5) + 1 words */);
int launch_popcount(cudaStream_t s, const uint32_t* words, uint64_t n, unsigned long long* total /* zeroed */);

// order-independent checksum: out[0] += sum, out[1] ^= xor of the per-pattern values, out[2] += occurrences, out[4] += per-reference values (out zeroed by the caller)
int launch_model_checksum(cudaStream_t s, const uint8_t* keys, const uint64_t* off, const uint32_t* counts, uint64_t np, uint64_t* keyhash /* may be NULL */, unsigned long long* out);
int launch_refs_checksum(cudaStream_t s, const uint64_t* keyhash, const uint64_t* ref_off, uint64_t np, const uint32_t* rs, const uint16_t* rt, uint64_t nrefs, unsigned long long* out);

// ---- parity helpers / measurement input
int launch_hash64_batch(cudaStream_t s, const uint8_t* keys, const uint64_t* off, uint64_t n, uint64_t* out);
int launch_synth_lengths(cudaStream_t s, uint64_t seed, uint64_t ntokens, uint64_t first, uint32_t vocab, uint32_t mean_sentence, uint32_t phrase_permille, uint32_t nphrases,
                         const uint64_t* cdf, uint32_t* lens);
int launch_synth_write(cudaStream_t s, uint64_t seed, uint64_t ntokens, uint64_t first, uint32_t vocab, uint32_t mean_sentence, uint32_t phrase_permille, uint32_t nphrases,
                       const uint64_t* cdf, const uint64_t* off, uint8_t* out);

}  // namespace colibri
