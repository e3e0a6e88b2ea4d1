/* colibri_b200.h -- C ABI of the B200-native PatternModel::train path.
 *
 * The reference (proycon/colibri-core) has no FFI seam of its own: the boundary a
 * caller sees is the C++ class API
 *     PatternModel<uint32_t>::train(std::istream*, const PatternModelOptions&, ...)   include/patternmodel.h:880
 *     PatternModel<uint32_t>::train(const std::string&, ...)                          include/patternmodel.h:1353
 *     IndexedPatternModel<>::train(...)                                               include/patternmodel.h:2821-2844
 * called from src/patternmodeller.cpp:319, src/benchmarks.cpp:224, src/test.cpp:1212/1226/1266 and the Cython
 * wrapper colibricore_patternmodel.pxi:290/294.  This header is what an implementation of that method binds to:
 * plain pointers and sizes, no C++ or torch types.  The host-side mirror of the class API that calls it lives in
 * colibri-core_b200/host/ (C++), INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success and a COLIBRI_E_* code otherwise; the message is in
 *     colibri_b200_last_error() (thread local).  The C++ wrapper prints it on std::cerr and throws InternalError,
 *     which is the reference's error convention (include/common.h:41-44, patternmodel.h:958-963).
 *   - "corpus" always means the BODY of a .colibri.dat v2 file: the bytes after the 0xA2 0x02 header
 *     (src/classencoder.cpp:551-556; train() itself seeks to offset 2, patternmodel.h:997-1004).
 *   - device memory is owned by the library; host buffers are owned by the caller.
 *   - one host thread per handle (the reference is single threaded and not re-entrant either).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with COLIBRI_E_CUDA.
 */
#ifndef COLIBRI_B200_H
#define COLIBRI_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define COLIBRI_OK 0
#define COLIBRI_E_INVALID 1     /* bad argument */
#define COLIBRI_E_UNSUPPORTED 2 /* option combination outside the accelerated subset */
#define COLIBRI_E_CUDA 3        /* CUDA runtime failure, or no device */
#define COLIBRI_E_FORMAT 4      /* corpus bytes are not a well-formed class-encoded v2 stream */
#define COLIBRI_E_CAPACITY 5    /* a device table or index would exceed its addressable size */

#define COLIBRI_UNINDEXEDPATTERNMODEL 10 /* include/patternmodel.h:68-75 */
#define COLIBRI_INDEXEDPATTERNMODEL 20

/* POD mirror of the PatternModelOptions fields train() reads (include/patternmodel.h:103-180; same defaults via
 * colibri_b200_options_default) plus the two things the C++ caller decides outside the options struct. */
typedef struct colibri_b200_options {
    int32_t MINTOKENS;              /* -1 -> 2, 0 -> 1 (patternmodel.h:883-886) */
    int32_t MINTOKENS_SKIPGRAMS;    /* raised to MINTOKENS (:887-888) */
    int32_t MINTOKENS_UNIGRAMS;
    int32_t MINLENGTH;
    int32_t MAXLENGTH;
    int32_t MAXBACKOFFLENGTH;
    int32_t MINSKIPTYPES;
    int32_t MAXSKIPS;
    int32_t DOSKIPGRAMS;            /* non-exhaustive (indexed post-pass, :2969-3010) */
    int32_t DOSKIPGRAMS_EXHAUSTIVE; /* counted inside train(), :1163-1171 */
    int32_t DOPATTERNPERLINE;
    int32_t PRUNENONSUBSUMED;
    int32_t PRUNESUBSUMED;
    int32_t QUIET;
    int32_t DEBUG;
    int32_t model_type;             /* COLIBRI_UNINDEXEDPATTERNMODEL or COLIBRI_INDEXEDPATTERNMODEL */
    int32_t streamed;               /* 1: the caller's sentences come from Pattern(std::istream&) (src/pattern.cpp:483-587),
                                       0: from a preloaded IndexedCorpus (src/pattern.cpp:1916-1967).  Only matters when the
                                       last sentence lacks its 0x00 (the stream reader then repeats the last byte). */
    int32_t device;                 /* CUDA device ordinal */
    /* load-time switches (PatternModel::load, include/patternmodel.h:827-858 -> PatternMapStore::read, include/patternstore.h:555-619) */
    int32_t DOREMOVEINDEX;          /* accepted for signature parity; choosing model_type = unindexed is what drops the index */
    int32_t DOREMOVENGRAMS;
    int32_t DOREMOVESKIPGRAMS;
    int32_t DOREMOVEFLEXGRAMS;
    int32_t DORESET;                /* values start empty (counts 0, no references) */
} colibri_b200_options;

/* Thread safety: calls on DIFFERENT handles may run concurrently (the device pool, the stream / event caches and the error text are
 * thread-safe or thread-local).  One handle must not be used by two calls at a time -- that includes a corpus or a constraint model shared by two
 * concurrent trainings: training writes the sentence-source tail behind the staged body and counts inside the constraint model's index. */
typedef struct colibri_b200_corpus colibri_b200_corpus; /* a corpus body resident in HBM, tokenised lazily */
typedef struct colibri_b200_model  colibri_b200_model;  /* a trained model: device-resident survivors + host-side stats */

const char* colibri_b200_last_error(void);
const char* colibri_b200_version(void);
/* number of usable CUDA devices (0 when there is none or the driver is missing) */
int         colibri_b200_device_count(void);
void        colibri_b200_options_default(colibri_b200_options* opt);

/* ---- corpus staging (replaces: std::ifstream reads of patternmodel.h:1353-1364 / IndexedCorpus::load src/pattern.cpp:1916-1967) */
/* copy nbytes of corpus body from host memory to the device (one cudaMemcpyAsync; pinned or pageable source) */
int  colibri_b200_corpus_stage(const uint8_t* host_body, size_t nbytes, int device, colibri_b200_corpus** out);
/* adopt (copy, device-to-device) a body that already lives in device memory, e.g. produced by colibri_b200_synth_corpus */
int  colibri_b200_corpus_from_device(const void* dev_body, size_t nbytes, int device, colibri_b200_corpus** out);
size_t colibri_b200_corpus_bytes(const colibri_b200_corpus* c);
void colibri_b200_corpus_free(colibri_b200_corpus* c);

/* ---- training (replaces PatternModel::train, include/patternmodel.h:880-1345, for the subset documented in DESIGN.md) */
/* end to end from host memory: stage + train + leave the result ready for export */
int  colibri_b200_train(const uint8_t* host_body, size_t nbytes, const colibri_b200_options* opt, colibri_b200_model** out);
/* end to end into caller-owned host buffers, with the PCIe legs overlapped: the survivors of level n are turned into pattern bytes and copied
 * to the host (second stream, asynchronous when the buffers are pinned) while level n+1 counts.  Output = the compact flat form of
 * colibri_b200_model_export_compact (keys blob, key_len[], counts[]; unindexed models).  No device-resident model is left behind; what a
 * caller reads after train() comes back in the summary.  If a buffer is too small the call returns COLIBRI_E_CAPACITY and the summary
 * holds the sizes needed.  Replaces train() + iteration for callers that materialise the map on the host (host/patternmodel.h: adopt()). */
typedef struct colibri_b200_train_summary {
    uint64_t npatterns, keybytes, totaltokens, totaltypes;
    int32_t  maxn, minn, hasskipgrams, npasses;
    uint64_t passes[32][4];          /* n, found n-grams, found skipgrams, pruned (first 32 passes) */
    double   ms[16];                 /* COLIBRI_T_* phase times of the device work */
    uint64_t counters[8];            /* as colibri_b200_model_counters */
} colibri_b200_train_summary;
int  colibri_b200_train_export(const uint8_t* host_body, size_t nbytes, const colibri_b200_options* opt, uint8_t* keys, uint64_t keys_cap, uint16_t* key_len,
                               uint32_t* counts, uint64_t patterns_cap, colibri_b200_train_summary* summary);
/* from a staged corpus (device resident input); the corpus can be trained repeatedly with different options */
int  colibri_b200_train_corpus(colibri_b200_corpus* corpus, const colibri_b200_options* opt, colibri_b200_model** out);
void colibri_b200_model_free(colibri_b200_model* m);

/* ---- what callers read after train(): size() :744, tokens() :1709, types() :1700, maxlength()/minlength() :1640-1648 */
uint64_t colibri_b200_model_size(const colibri_b200_model* m);
uint64_t colibri_b200_model_tokens(const colibri_b200_model* m);
uint64_t colibri_b200_model_types(const colibri_b200_model* m);
int      colibri_b200_model_maxn(const colibri_b200_model* m);
int      colibri_b200_model_minn(const colibri_b200_model* m);
int      colibri_b200_model_hasskipgrams(const colibri_b200_model* m);
int      colibri_b200_model_type(const colibri_b200_model* m);
/* the numbers of the progress lines " Found X ngrams...pruned Y...total kept: Z" (patternmodel.h:1195-1245):
 * out[0]=n, out[1]=found n-grams (distinct, new), out[2]=found skipgrams (distinct, new), out[3]=pruned */
int      colibri_b200_model_passes(const colibri_b200_model* m);
int      colibri_b200_model_pass_stats(const colibri_b200_model* m, int pass, uint64_t out[4]);

/* ---- export (replaces iteration over the PatternMap / PatternMapStore::write, include/patternstore.h:534-542)
 * Flat form: keys = concatenated pattern bytes (no terminators), key_off[npatterns+1], counts[npatterns];
 * indexed models add ref_off[npatterns+1] and (sentence 1-based, token 0-based) pairs sorted ascending
 * (IndexReference, include/datatypes.h:33-89).  Pattern order is unspecified (so is the reference's). */
int colibri_b200_model_export_sizes(colibri_b200_model* m, uint64_t* npatterns, uint64_t* keybytes, uint64_t* nrefs);
int colibri_b200_model_export(colibri_b200_model* m, uint8_t* keys, uint64_t* key_off, uint32_t* counts, uint32_t* ref_sentence, uint16_t* ref_token, uint64_t* ref_off);
/* the same without offsets: key_len[npatterns] = bytes of each key, in pattern order (2 instead of 8 bytes per pattern to copy) */
int colibri_b200_model_export_compact(colibri_b200_model* m, uint8_t* keys, uint16_t* key_len, uint32_t* counts);
/* the .colibri.patternmodel byte stream (include/patternmodel.h:1609-1624): returns needed size in *nbytes when buf is NULL */
int colibri_b200_model_write(colibri_b200_model* m, uint8_t* buf, size_t cap, size_t* nbytes);
/* occurrencecount(pattern) (include/patternmodel.h:1653-1669): key = pattern bytes without terminator; *count = 0 if absent.
 * Runs on the device: SpookyV2 of the key bytes (Pattern::hash, src/pattern.cpp:234-238) into the model's HBM-resident index. */
int colibri_b200_model_lookup(colibri_b200_model* m, const uint8_t* key, uint32_t len, uint32_t* count);
/* has() / occurrencecount() for n patterns at once (include/patternmodel.h:751-756, :1653-1669): keys = concatenated pattern bytes,
 * key_off[n+1]; counts[i] = 0 and index[i] = -1 when pattern i is absent, else its position in the export order.  Either output may be NULL. */
int colibri_b200_model_lookup_batch(colibri_b200_model* m, const uint8_t* keys, const uint64_t* key_off, uint64_t n, uint32_t* counts, int64_t* index);

/* ---- several GPUs of one node from ONE process (the C++ host / CLI; torchrun-launched runs drive the shard phases above themselves) */
/* PatternModel::train (include/patternmodel.h:880-1345), unindexed n-gram models: the corpus is cut at sentence boundaries into ndev shards, one host
 * thread per device runs the shard phases, keys and replies travel through peer-mapped buffers over NVLink (cudaDeviceEnablePeerAccess).  out[r]
 * receives device r's SHARE of the model -- the patterns it exports, with their global counts; tokens, types and passes are the global numbers in
 * every share -- so the model is the union of the shares (disjoint).  Skipgrams, indexed models and MINLENGTH > 1 are refused here. */
int  colibri_b200_train_multi(const uint8_t* host_body, size_t nbytes, const colibri_b200_options* opt, const int* devices, int ndev, colibri_b200_model** out /* ndev handles */);

/* ---- the reverse index of a model over a corpus, and the relations computed from it */
typedef struct colibri_b200_rindex colibri_b200_rindex;
/* Replaces the ReverseIndex / IndexedCorpus argument of IndexedPatternModel (include/patternmodel.h:2681-2720) for the queries below: every window of
 * every length the model holds is matched against the model once (the kernels of constrained training), so that getreverseindex
 * (:1746-1824) becomes a gather.  n-grams only: a model with skipgrams or flexgrams answers for its n-grams.  `streamed` as in
 * the options struct (the sentence source).  The model and the corpus must stay alive while the index is used. */
int  colibri_b200_rindex_build(colibri_b200_model* model, colibri_b200_corpus* corpus, int streamed, colibri_b200_rindex** out);
void colibri_b200_rindex_free(colibri_b200_rindex* r);
/* out[0] = sentences (IndexedCorpus::sentences()), out[1] = positions (tokens + delimiters), out[2], out[3] = shortest / longest n-gram length indexed */
int  colibri_b200_rindex_info(const colibri_b200_rindex* r, uint64_t out[4]);
/* the n-gram lengths the index holds, ascending: column k of a query answer is the pattern of lengths[k] tokens */
int  colibri_b200_rindex_lengths(const colibri_b200_rindex* r, uint32_t* lengths, uint32_t cap, uint32_t* n);
/* out[k] = position of the first token of sentence k + 1 (k = 0 .. sentences); sentence k + 1 has out[k + 1] - out[k] - 1 tokens (IndexedCorpus::sentencelength) */
int  colibri_b200_rindex_sentence_starts(const colibri_b200_rindex* r, uint32_t* out, uint64_t cap);
/* getreverseindex(IndexReference(sentence, token)) for nq references at once (sentences count from 1, tokens from 0): out[q * nlengths + k] = index + 1
 * (export order of the model) of the model's n-gram of lengths[k] tokens that starts there, 0 if none -- also for a reference outside its sentence. */
int  colibri_b200_rindex_query(colibri_b200_rindex* r, const uint32_t* sentence, const uint16_t* token, uint64_t nq, uint32_t* out);
/* getrightcooc (direction 0, :3460-3493) / getleftcooc (direction 1, :3502-3531) of EVERY pattern, with the reference's arithmetic: its
 * getreverseindex_right / _left (:1867-1878, :1885-1892) look each neighbouring position up at the original reference, so a relation (P, Q) joins two
 * model patterns that start at the same position (s, t), which adds max(0, sl - 1 - (t + |P|)) (right) or max(0, t - |Q|) (left) to joint(P, Q).
 * *nrel = relations found; when the three buffers are given (cap entries each) they receive (index of P, index of Q, joint) in no particular order. */
int  colibri_b200_rindex_cooc(colibri_b200_rindex* r, int direction, uint32_t* idx_p, uint32_t* idx_q, uint64_t* joint, uint64_t cap, uint64_t* nrel);
/* getcooc (:3543-3576) of the pattern with export index `pattern`, both directions: count(Q) = pairs (occurrence of P at token t, start t2 of the model
 * n-gram Q in the same sentence) with t2 + |Q| < t or t2 > t + |P| (no overlap, not adjacent).  *nrel = patterns Q with a count; when both buffers are
 * given (cap entries each) they receive (index of Q, count) in no particular order.  The reference's occurrencethreshold / category / size /
 * ordersignificant arguments are filters on this list and live on the host side (host/patternmodel.h, pybinding.py).  Asking twice for the same
 * pattern (first for *nrel, then with buffers) counts once. */
int  colibri_b200_rindex_cooc_of(colibri_b200_rindex* r, uint64_t pattern, uint32_t* idx_q, uint64_t* count, uint64_t cap, uint64_t* nrel);

/* ---- models that do not come out of train() */
/* a pattern set given as flat host arrays (the export form) becomes a device-resident model: replaces building a PatternModel /
 * PatternSetModel by insert() (include/patternstore.h:520, include/patternmodel.h:296-470).  counts and the three ref arrays may be NULL. */
int colibri_b200_model_from_flat(const uint8_t* keys, const uint64_t* key_off, const uint32_t* counts, uint64_t npatterns, const uint32_t* ref_sentence,
                                 const uint16_t* ref_token, const uint64_t* ref_off, uint64_t totaltokens, uint64_t totaltypes, int model_type, int device,
                                 colibri_b200_model** out);
/* PatternModel::load (include/patternmodel.h:781-861): the bytes of a .colibri.patternmodel file (type 10 or 20, version 2) read AS
 * opt->model_type, with the options acting as filters exactly like PatternMapStore::read (include/patternstore.h:555-619): count >= MINTOKENS
 * (-1 -> 0), MINLENGTH <= n <= MAXLENGTH, DOREMOVE{NGRAMS,SKIPGRAMS,FLEXGRAMS}, membership in `constrain` (may be NULL), DORESET.  The record
 * stream is scanned on the host (its boundaries are sequential), the filters run on the device. */
int colibri_b200_model_load(const uint8_t* file, size_t nbytes, const colibri_b200_options* opt, colibri_b200_model* constrain, colibri_b200_model** out);
/* PatternModel::train with constrainbymodel != NULL (include/patternmodel.h:880-1345: one scan :1064-1072, membership :1088-1089, prune :1211-1218):
 * only patterns of `constrain` are counted.  inplace = 0: constrain is another model (the CLI's -j; totals start from its totals, :892-895);
 * inplace = 1: constrainbymodel == this, i.e. `constrain` is the model being rebuilt, loaded with DORESET (the CLI's -I and stage 2 of -2;
 * :889-891, :970-971, :1199-1201).  `constrain` is not modified; the result is a new model. */
int colibri_b200_train_constrained(colibri_b200_corpus* corpus, const colibri_b200_options* opt, colibri_b200_model* constrain, int inplace, colibri_b200_model** out);

/* The same over a corpus sharded at sentence boundaries, one shard per rank (SURVEY.md 8e applied to 8f-2; unindexed models): the constraint set is
 * replicated on every rank; _count adds this shard's occurrences of every pattern to dev_counts (device u32[size of constrain], zeroed by the
 * caller) -- no communication; the caller sums dev_counts and the token counts over the ranks (one all-reduce each); _finish thresholds the sums
 * and returns the model (identical on every rank).  Every rank must hold the SAME pattern numbering, i.e. load the same model file / upload the
 * same flat arrays (the export order of a freshly trained model is not deterministic).  colibri_b200_train_constrained == _count + _finish on one rank. */
int colibri_b200_constrained_count(colibri_b200_corpus* shard, const colibri_b200_options* opt, colibri_b200_model* constrain, void* dev_counts, uint64_t* shard_tokens,
                                   uint64_t* kernel_launches);
int colibri_b200_constrained_finish(const colibri_b200_options* opt, colibri_b200_model* constrain, void* dev_counts, uint64_t corpus_tokens, int inplace, colibri_b200_model** out);

/* ---- flexgrams (first piece)
 * IndexedPatternModel::computeflexgrams_fromskipgrams (include/patternmodel.h:3724-3744; CLI `-F S`, src/patternmodeller.cpp:330-337): every
 * skipgram of an indexed model is abstracted to its flexgram (Pattern::toflexgram, src/pattern.cpp:145-180) and hands it all its occurrences.
 * *found = the new flexgrams (the reference's return value); *out = a new model = m's patterns + the flexgrams, each flexgram's occurrence
 * list ascending.  m is not modified. */
int colibri_b200_model_flexgrams_fromskipgrams(colibri_b200_model* m, uint64_t* found, colibri_b200_model** out);
int colibri_b200_model_hasflexgrams(const colibri_b200_model* m);

/* ---- measurement hooks (bench.py): device time by phase, from CUDA events on the library's stream */
#define COLIBRI_T_TOTAL 0     /* whole train_corpus call on the device (tokenise .. survivors ready) */
#define COLIBRI_T_TOKENISE 1  /* K0 */
#define COLIBRI_T_UNIGRAMS 2  /* K1 histogram + unigram prune */
#define COLIBRI_T_COUNT 3     /* sum over n >= 2 of the n-gram upsert kernel (the dominant kernel) */
#define COLIBRI_T_SKIPGRAMS 4 /* sum of the skipgram upsert kernel */
#define COLIBRI_T_PRUNE 5     /* table scan + compaction + relabel, all levels */
#define COLIBRI_T_EXPORT 6    /* device side of the flat export (lengths, scan, byte writer) */
#define COLIBRI_T_H2D 7       /* corpus staging copy (colibri_b200_train only) */
#define COLIBRI_T_INDEX 8     /* forward index fill + sort (indexed models) */
#define COLIBRI_T_NPHASES 9
int colibri_b200_model_timings(const colibri_b200_model* m, double ms[COLIBRI_T_NPHASES]);
/* work counters of the last train: out[0]=positions (tokens+delimiters), out[1]=corpus bytes, out[2]=kernel launches,
 * out[3]=n-gram upserts (valid windows, all n>=2), out[4]=skipgram upserts, out[5]=table slots initialised (sum),
 * out[6]=unigram increments, out[7]=bytes of device memory at the peak */
int colibri_b200_model_counters(const colibri_b200_model* m, uint64_t out[8]);
/* per level n>=2: out[0]=valid windows, out[1]=table capacity in slots, out[2]=ms of the level's filter + count kernels,
 * out[3]=windows the occurrence filter proved to be singletons (they never reach the table: upserts = out[0] - out[3]) */
/* order-independent checksum of the model's (pattern bytes, count[, occurrence list]) content: out[0] = sum and out[1] = xor over the patterns of
 * fmix64(FNV1a64(key bytes) ^ count * 0x9E3779B97F4A7C15), out[2] = total occurrences, out[3] = patterns, out[4] = sum over the stored references
 * of fmix64(FNV1a64(key) ^ (sentence << 16 | token) * 0xD6E8FEB86659FD93) (indexed models), out[5] = references.  The shares of a sharded model
 * combine (add / xor) to the checksum of the whole; the reference's model file yields the same numbers (tests/checksum.py).  Measurement aid: lets
 * bench.py compare 10^8-pattern models across GPU counts and against committed fixtures without shipping and sorting them. */
int colibri_b200_model_checksum(colibri_b200_model* m, uint64_t out[6]);
int colibri_b200_model_level_counters(const colibri_b200_model* m, int n, double out[4]);
/* same plus out[4]=items the level's kernels enumerated: every position (dense mode), or the length of the position list the previous
 * level left behind (list mode: only positions whose (n-1)-gram survived are visited); out[5]=1 if the level ran on the partitioned path
 * (shared-memory counting, out[1] = partitions, out[3] = n-grams that occur once) else 0 (HBM table, out[3] = windows the occurrence filter held
 * back); out[6]=1 if the occurrence filter ran; out[7]=1 if out[2] includes the sweep that writes the level-1 ids
 * (level 2 on the partitioned path: one kernel writes the ids and takes level 2's first pass) */
int colibri_b200_model_level_info(const colibri_b200_model* m, int n, double out[8]);

/* ---- multi-GPU: one process per GPU drives these per-rank phases and moves the buffers between ranks itself
 * (torch.distributed / NCCL all-to-all); the library does no communication.  Model = hash-partitioned across ranks,
 * corpus = sharded at sentence boundaries (nothing in the reference corresponds to this).
 * Order per rank: shard_begin, shard_unigram_counts, [all-reduce SUM of the u32 counts], shard_unigram_finish, then for
 * n = 2.. : shard_level_split_count, shard_level_split_write, [all-to-all of 8-byte keys], shard_level_owner,
 * shard_level_owner_survivors, [all-to-all back of 4-byte replies, all-to-all of 8-byte survivor records],
 * shard_level_finish; finally shard_finish. */
typedef struct colibri_b200_shard colibri_b200_shard;
int    colibri_b200_shard_begin(colibri_b200_corpus* corpus, const colibri_b200_options* opt, int rank, int world, colibri_b200_shard** out);
/* out[0]=local tokens, out[1]=local maximum class, out[2]=positions, out[3]=kernel launches so far */
int    colibri_b200_shard_info(const colibri_b200_shard* sh, uint64_t out[4]);
double colibri_b200_shard_device_ms(const colibri_b200_shard* sh);
/* device ms by phase: [0] tokenise, [1] unigrams, [2] level_count, [3] level_pack, [4] level_merge, [5] level_finish, [6] export */
int    colibri_b200_shard_phase_ms(const colibri_b200_shard* sh, double out[8]);
int    colibri_b200_shard_unigram_counts(colibri_b200_shard* sh, uint32_t nclasses, void* dev_counts /* u32[nclasses] */);
/* stats[0]=distinct unigrams (global), [1]=kept, [2]=occurrences kept */
int    colibri_b200_shard_unigram_finish(colibri_b200_shard* sh, const void* dev_global_counts, uint64_t global_tokens, uint64_t stats[3]);
/* Dense pairs of level 2 (optional, before level 2; all ranks alike): dev_square = the caller's zeroed u32[dim * dim] on this rank's device.  The split
 * of level 2 counts this rank's windows of two classes below dim into it instead of shipping them; the CALLER sums the squares of all ranks
 * (one all-reduce) before shard_level_owner / shard_p2p_owner.  dim = 0 switches it off. */
int    colibri_b200_shard_set_dense(colibri_b200_shard* sh, void* dev_square, uint32_t dim);
/* send_counts[world]: valid windows of level n whose key is owned by each rank; *windows = their sum */
int    colibri_b200_shard_level_split_count(colibri_b200_shard* sh, int n, uint64_t* send_counts, uint64_t* windows);
int    colibri_b200_shard_level_split_write(colibri_b200_shard* sh, void* dev_send_keys /* 8 B per window, grouped by owner */);
/* recv_counts[world] per source; dev_reply: one u32 per received window; stats[0]=distinct keys owned, [1]=kept, [2]=occurrences kept;
 * surv_counts[world]: survivor records to return to each source */
int    colibri_b200_shard_level_owner(colibri_b200_shard* sh, const void* dev_recv_keys, const uint64_t* recv_counts, void* dev_reply, uint64_t stats[3], uint64_t* surv_counts);
int    colibri_b200_shard_level_owner_survivors(colibri_b200_shard* sh, void* dev_out /* 8 B per record, grouped by source */);
int    colibri_b200_shard_level_finish(colibri_b200_shard* sh, const void* dev_reply_back, const void* dev_surv, const uint64_t* surv_counts /* per owner */, uint64_t* local_valid);
/* exhaustive skipgrams of the level just finished (n >= 3): shard_skip_split_count, shard_skip_split_write,
 * [all-to-all of 16-byte keys], shard_skip_owner, shard_skip_owner_survivors, [all-to-all of 16-byte survivor records],
 * shard_skip_finish.  stats[0]=distinct skipgrams owned, [1]=kept. */
int    colibri_b200_shard_skip_split_count(colibri_b200_shard* sh, uint64_t* send_counts, uint64_t* nrecords);
int    colibri_b200_shard_skip_split_write(colibri_b200_shard* sh, void* dev_send);
int    colibri_b200_shard_skip_owner(colibri_b200_shard* sh, const void* dev_recv, const uint64_t* recv_counts, uint64_t stats[2], uint64_t* surv_counts);
int    colibri_b200_shard_skip_owner_survivors(colibri_b200_shard* sh, void* dev_out);
int    colibri_b200_shard_skip_finish(colibri_b200_shard* sh, const void* dev_surv, const uint64_t* surv_counts);
/* NVLink peer-store mode: the caller allocates symmetric receive buffers on every rank (e.g. torch.distributed._symmetric_memory),
 * passes every rank's device pointers (keys_rx: G slots x slot_cap x 8 B; reply_rx: G x slot_cap x 4 B; surv_rx: G x surv_cap x 8 B;
 * hdr: 6*G u64 words) and provides a device-side barrier on the stream given to shard_set_stream.  A level is then
 * shard_p2p_split, [barrier], shard_p2p_owner, [barrier], shard_p2p_finish: the split kernel stores keys straight into the
 * owners' slots and the reply kernel stores ids straight into the senders' slots over NVLink; no all-to-all call is made. */
int    colibri_b200_shard_set_stream(colibri_b200_shard* sh, void* cuda_stream);
int    colibri_b200_shard_set_peers(colibri_b200_shard* sh, const uint64_t* keys_rx, const uint64_t* reply_rx, const uint64_t* surv_rx, const uint64_t* hdr, uint64_t slot_cap,
                                    uint64_t surv_cap);
int    colibri_b200_shard_p2p_split(colibri_b200_shard* sh, int n, uint64_t* windows);
int    colibri_b200_shard_p2p_owner(colibri_b200_shard* sh, uint64_t stats[3]);
int    colibri_b200_shard_p2p_finish(colibri_b200_shard* sh, uint64_t global_stats[3], uint64_t* local_valid);
/* passes: npasses x {n, found, foundskip, pruned} global numbers; the returned model holds THIS RANK'S share of the patterns */
int    colibri_b200_shard_finish(colibri_b200_shard* sh, const uint64_t* passes, int npasses, uint64_t global_types, int maxn, int minn, colibri_b200_model** out);
void   colibri_b200_shard_free(colibri_b200_shard* sh);

/* ---- L1 pieces exposed for parity tests of SURVEY.md 8(a) rows a1/a5/a10 */
/* SpookyHash::Hash64(key, len, 0) computed ON THE DEVICE for n variable-length messages (Pattern::hash, src/pattern.cpp:234-238) */
int colibri_b200_hash64_batch(const uint8_t* keys, const uint64_t* key_off, uint64_t n, uint64_t* out, int device);
/* decode the corpus on the device and return class ids (0 = sentence delimiter) -- bytestoint, src/classdecoder.cpp:20-43 */
int colibri_b200_corpus_tokens(colibri_b200_corpus* c, uint32_t* out, uint64_t cap, uint64_t* ntokens_and_delims);

/* ---- synthetic corpus generator (measurement input, not part of the reference): counter based and integer only,
 * bit-identical to oracle_synth_corpus.  Writes the body into a new staged corpus. */
typedef struct colibri_b200_synth_params {
    uint64_t seed;
    uint64_t ntokens;
    uint32_t vocab;
    uint32_t mean_sentence;
    uint32_t phrase_permille;
    uint32_t nphrases;
    uint64_t first_token;   /* index of this corpus's first token in the global stream (multi-GPU shards); 0 for a whole corpus */
} colibri_b200_synth_params;
int colibri_b200_synth_corpus(const colibri_b200_synth_params* p, int device, colibri_b200_corpus** out);
/* copy a staged corpus body back to the host */
int colibri_b200_corpus_download(const colibri_b200_corpus* c, uint8_t* host, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
