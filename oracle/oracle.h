/* oracle/oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference algorithm on the hot path
 * (PatternModel::train and what it calls).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library, and
 * only as the checker.  The shipped path (colibri-core_b200/) never links it.
 *
 * Parity status: PINNED.  The restatement is checked against
 *   - the reference's own known-answer numbers (src/test.cpp:1211-1283, :1321-1337),
 *   - model files written by the unmodified reference compiled into oracle/_ref/
 *     (tests/golden/, produced by tests/golden/make_golden.py),
 *   - SpookyV2 known-answer hashes produced by the reference library.
 * Every function cites the reference file:line it follows.
 */
#ifndef COLIBRI_ORACLE_H
#define COLIBRI_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Mirror of the PatternModelOptions fields train() reads (include/patternmodel.h:103-180). */
typedef struct oracle_options {
    int32_t mintokens;              /* MINTOKENS            (-1 -> 2, 0 -> 1; :883-886) */
    int32_t mintokens_skipgrams;    /* MINTOKENS_SKIPGRAMS  (raised to MINTOKENS; :887-888) */
    int32_t mintokens_unigrams;     /* MINTOKENS_UNIGRAMS   (default 1) */
    int32_t minlength;              /* MINLENGTH            (default 1) */
    int32_t maxlength;              /* MAXLENGTH            (default 100) */
    int32_t maxbackofflength;       /* MAXBACKOFFLENGTH     (default 100) */
    int32_t minskiptypes;           /* MINSKIPTYPES         (default 2) */
    int32_t maxskips;               /* MAXSKIPS             (default 3) */
    int32_t doskipgrams;            /* DOSKIPGRAMS          (indexed models: trainskipgrams post-pass, patternmodel.h:2969-3010) */
    int32_t doskipgrams_exhaustive; /* DOSKIPGRAMS_EXHAUSTIVE */
    int32_t indexed;                /* 0: PatternModel<uint32_t>, 1: IndexedPatternModel<> */
    int32_t streamed;               /* 1: sentences come from Pattern(std::istream&) (src/pattern.cpp:483-587),
                                       0: from a preloaded IndexedCorpus (src/pattern.cpp:1916-1967, :2135-2158).
                                       They differ when the final 0x00 is missing. */
} oracle_options;

void oracle_options_default(oracle_options* o);

typedef struct oracle_model oracle_model;

/* Returns 0 on success; non-zero with a message in oracle_last_error() otherwise.
 * corpus = the bytes of a .colibri.dat v2 file AFTER its 2-byte header (0xA2 0x02). */
int         oracle_train(const uint8_t* corpus, size_t nbytes, const oracle_options* opt, oracle_model** out);
void        oracle_model_free(oracle_model* m);
const char* oracle_last_error(void);

uint64_t    oracle_model_size(const oracle_model* m);
uint64_t    oracle_model_tokens(const oracle_model* m);
uint64_t    oracle_model_types(const oracle_model* m);
int         oracle_model_maxn(const oracle_model* m);
int         oracle_model_minn(const oracle_model* m);
int         oracle_model_hasskipgrams(const oracle_model* m);
/* Per executed pass: out[0]=n, out[1]=found n-grams, out[2]=found skipgrams (new distinct), out[3]=pruned (incl. extra skipgram pruning).
 * These are the numbers of the reference's " Found X ngrams...pruned Y...total kept: Z" lines (patternmodel.h:1195-1245). */
int         oracle_model_passes(const oracle_model* m);
int         oracle_model_pass_stats(const oracle_model* m, int pass, uint64_t out[4]);
/* occurrence count of one pattern (key bytes without terminator); 0 if absent. */
uint32_t    oracle_model_count(const oracle_model* m, const uint8_t* key, uint32_t len);

/* Canonical export: patterns sorted by key bytes (memcmp order, shorter first on ties).
 * key_off has npatterns+1 entries; ref_off has npatterns+1 entries (indexed only, may be NULL). */
void        oracle_model_export_sizes(const oracle_model* m, uint64_t* npatterns, uint64_t* keybytes, uint64_t* nrefs);
void        oracle_model_export(const oracle_model* m, uint8_t* keys, uint64_t* key_off, uint32_t* counts, uint32_t* ref_sentence, uint16_t* ref_token, uint64_t* ref_off);

/* Model file (.colibri.patternmodel) layout: include/patternmodel.h:1609-1624, include/patternstore.h:534-542,
 * src/pattern.cpp:268-277, include/datatypes.h:55-58, :263-270.
 *   0x00, type (10 unindexed / 20 indexed), version 2, u64 totaltokens, u64 totaltypes, u64 npatterns,
 *   then per pattern: key bytes, 0x00, u32 count [, count x (u32 sentence, u16 token)].
 * oracle_model_write emits patterns in canonical (sorted) order; returns bytes needed/written. */
size_t      oracle_model_write(const oracle_model* m, uint8_t* buf, size_t cap);
/* Parse any model file of type 10 or 20 (file order preserved).  scan fills hdr[0..6] =
 * type, version, totaltokens, totaltypes, npatterns, keybytes, nrefs; returns 0 on success. */
int         oracle_modelfile_scan(const uint8_t* data, size_t n, uint64_t hdr[7]);
int         oracle_modelfile_parse(const uint8_t* data, size_t n, uint8_t* keys, uint64_t* key_off, uint32_t* counts, uint32_t* ref_sentence, uint16_t* ref_token,
                                   uint64_t* ref_off);

/* Model load with the options acting as filters: PatternModel::load (include/patternmodel.h:781-861) over
 * PatternMapStore::read (include/patternstore.h:555-619). */
typedef struct oracle_load_options {
    int32_t mintokens;    /* MINTOKENS (-1 -> 0 here, patternstore.h:565-566): keep count >= mintokens */
    int32_t minlength;    /* keep MINLENGTH <= n <= MAXLENGTH */
    int32_t maxlength;
    int32_t dongrams;     /* !DOREMOVENGRAMS */
    int32_t doskipgrams;  /* !DOREMOVESKIPGRAMS */
    int32_t doflexgrams;  /* !DOREMOVEFLEXGRAMS */
    int32_t doreset;      /* DORESET: values start empty */
    int32_t load_indexed; /* 0: read as PatternModel<uint32_t>, 1: as IndexedPatternModel<> */
} oracle_load_options;
int         oracle_model_load(const uint8_t* file, size_t nbytes, const oracle_load_options* lo, const oracle_model* constrain, oracle_model** out);
int         oracle_model_from_keys(const uint8_t* keys, const uint64_t* key_off, uint64_t npatterns, uint64_t totaltokens, uint64_t totaltypes, int indexed,
                                   oracle_model** out);
/* train() under a constraint model (include/patternmodel.h:880-1345, constrainbymodel != NULL); inplace = (constrainbymodel == this) */
int         oracle_train_constrained(const uint8_t* corpus, size_t nbytes, const oracle_options* opt, const oracle_model* constrain, int inplace, oracle_model** out);

/* IndexedPatternModel::computeflexgrams_fromskipgrams (include/patternmodel.h:3724-3744) on an indexed model that holds skipgrams:
 * returns the number of new flexgrams (-1 on error); every flexgram's occurrence list comes out ascending. */
int64_t     oracle_computeflexgrams_fromskipgrams(oracle_model* m);
int         oracle_model_hasflexgrams(const oracle_model* m);

/* Codec + hash + masks (the L1 layer). */
unsigned    oracle_inttobytes(uint8_t* buf, uint32_t cls);               /* src/classencoder.cpp:22-42 */
uint32_t    oracle_bytestoint(const uint8_t* a, unsigned* length);        /* src/classdecoder.cpp:20-43 */
uint64_t    oracle_spooky_hash64(const void* msg, size_t len, uint64_t seed); /* src/SpookyV2.cpp:21-120 (Short path, len < 192) */
uint64_t    oracle_pattern_hash(const uint8_t* key, size_t len);         /* src/pattern.cpp:234-238 */
int         oracle_skip_configurations(int n, int maxskips, uint32_t* out, int cap); /* src/algorithms.cpp:79-94 */
/* Pattern(const PatternPointer&) for a skipgram: src/pattern.cpp:886-908.  Returns the collapsed byte length. */
size_t      oracle_skipgram_collapse(const uint8_t* ngram, size_t len, uint32_t mask, uint8_t* out);

/* Counter-based, integer-only synthetic corpus.  Writes the body of a .colibri.dat v2 file
 * (no 2-byte header).  Returns the number of bytes written (or needed, when out==NULL / cap too small). */
typedef struct oracle_synth_params {
    uint64_t seed;
    uint64_t ntokens;
    uint32_t vocab;         /* number of word types V; classes are 6 .. V+5 */
    uint32_t mean_sentence; /* a sentence ends after token i iff mix(seed^K, i) % mean_sentence == 0 */
    uint32_t phrase_permille; /* 0..1000: probability (per mille) that a position starts an injected phrase */
    uint32_t nphrases;      /* number of distinct fixed phrases (each 3..6 tokens) */
    uint64_t first_token;   /* index of the first token in the global stream (multi-GPU shards); 0 for a whole corpus */
} oracle_synth_params;
uint64_t    oracle_synth_token(const oracle_synth_params* p, const uint64_t* cdf, uint64_t i); /* class id of token i */
size_t      oracle_synth_corpus(const oracle_synth_params* p, uint8_t* out, size_t cap);
/* integer Zipf table: cdf[r] = sum_{k<=r} floor(2^40/(k+1)), r = 0..vocab-1 */
void        oracle_synth_cdf(uint32_t vocab, uint64_t* cdf);

#ifdef __cplusplus
}
#endif
#endif
