/* oracle/oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See oracle.h.
 *
 * A sequential, single-threaded restatement (in our own words) of what the
 * reference does on the PatternModel::train path.  It deliberately keeps the
 * reference's structure -- one corpus scan per n, a byte-string keyed hash map
 * hashed with SpookyV2, lookback of the two (n-1)-grams, prune after each pass
 * -- so that it is an independent check on the very differently organised CUDA
 * path.  Parity status: PINNED (tests/test_oracle_golden.py).
 */
#include "oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* errors                                                                     */
static char g_err[512];
const char* oracle_last_error(void) {
    return g_err;
}
static int fail(const char* msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return 1;
}

/* ------------------------------------------------------------------------- */
/* class codec: src/classencoder.cpp:22-42, src/classdecoder.cpp:20-43        */
unsigned oracle_inttobytes(uint8_t* buf, uint32_t cls) {
    /* little-endian base 128; bit 7 set on every byte but the last */
    unsigned len = 0;
    do {
        uint8_t digit = (uint8_t)(cls & 0x7F);
        cls >>= 7;
        if (cls)
            digit |= 0x80;
        if (buf)
            buf[len] = digit;
        ++len;
    } while (cls);
    return len;
}

uint32_t oracle_bytestoint(const uint8_t* a, unsigned* length) {
    uint32_t v = 0;
    unsigned i = 0;
    for (;; ++i) {
        uint8_t b = a[i];
        v += (uint32_t)(b & 0x7F) << (7 * i);
        if (b < 0x80)
            break;
    }
    if (length)
        *length = i + 1;
    return v;
}

/* ------------------------------------------------------------------------- */
/* SpookyHash V2 (Bob Jenkins, public domain), the "Short" code path that
 * Hash64 -> Hash128 takes for messages under 192 bytes:
 * include/SpookyV2.h:59-66, :277-362, :383-392; src/SpookyV2.cpp:21-120.    */
#define SPOOKY_CONST 0xdeadbeefdeadbeefULL
static inline uint64_t rotl64(uint64_t x, int k) {
    return (x << k) | (x >> (64 - k));
}
static inline uint64_t load_le64(const uint8_t* p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return v; /* host is little-endian, as the reference assumes */
}
static void spooky_short_mix(uint64_t* h) {
    static const int rot[12] = {50, 52, 30, 41, 54, 48, 38, 37, 62, 34, 5, 36};
    /* h[(i+2)&3] = rot(h[(i+2)&3], r); h[(i+2)&3] += h[(i+3)&3]; h[i&3] ^= h[(i+2)&3]  for i = 0..11 */
    for (int i = 0; i < 12; ++i) {
        uint64_t* x = &h[(i + 2) & 3];
        *x          = rotl64(*x, rot[i]);
        *x += h[(i + 3) & 3];
        h[i & 3] ^= *x;
    }
}
static void spooky_short_end(uint64_t* h) {
    static const int rot[11] = {15, 52, 26, 51, 28, 9, 47, 54, 32, 25, 63};
    /* h[(i+3)&3] ^= h[(i+2)&3]; h[(i+2)&3] = rot(h[(i+2)&3], r); h[(i+3)&3] += h[(i+2)&3]  for i = 0..10 */
    for (int i = 0; i < 11; ++i) {
        uint64_t* x = &h[(i + 3) & 3];
        uint64_t* y = &h[(i + 2) & 3];
        *x ^= *y;
        *y = rotl64(*y, rot[i]);
        *x += *y;
    }
}
uint64_t oracle_spooky_hash64(const void* msg, size_t len, uint64_t seed) {
    const uint8_t* p = (const uint8_t*)msg;
    uint64_t       h[4]; /* a, b, c, d */
    size_t         rem = len & 31;
    h[0]               = seed;
    h[1]               = seed;
    h[2]               = SPOOKY_CONST;
    h[3]               = SPOOKY_CONST;
    if (len > 15) {
        const uint8_t* end = p + (len / 32) * 32;
        for (; p < end; p += 32) {
            h[2] += load_le64(p);
            h[3] += load_le64(p + 8);
            spooky_short_mix(h);
            h[0] += load_le64(p + 16);
            h[1] += load_le64(p + 24);
        }
        if (rem >= 16) {
            h[2] += load_le64(p);
            h[3] += load_le64(p + 8);
            spooky_short_mix(h);
            p += 16;
            rem -= 16;
        }
    }
    h[3] += (uint64_t)len << 56;
    /* last 0..15 bytes, zero padded into (c,d); an empty tail adds the constant instead */
    if (rem == 0) {
        h[2] += SPOOKY_CONST;
        h[3] += SPOOKY_CONST;
    } else {
        uint8_t tail[16];
        memset(tail, 0, sizeof tail);
        memcpy(tail, p, rem);
        h[2] += load_le64(tail);
        h[3] += load_le64(tail + 8);
    }
    spooky_short_end(h);
    return h[0];
}
uint64_t oracle_pattern_hash(const uint8_t* key, size_t len) {
    /* src/pattern.cpp:234-238: empty pattern hashes to 0, otherwise Hash64(data, bytesize, seed 0) */
    if (len == 0 || key[0] == 0)
        return 0;
    return oracle_spooky_hash64(key, len, 0);
}

/* ------------------------------------------------------------------------- */
/* gap masks: src/algorithms.cpp:33-53 (mask2vector) and :79-94               */
static int count_gap_runs(uint32_t mask, int n) {
    int runs = 0, in = 0;
    for (int i = 0; i < n; ++i) {
        int bit = (i < 32) ? (int)((mask >> i) & 1u) : 0;
        if (bit && !in)
            ++runs;
        in = bit;
    }
    return runs;
}
int oracle_skip_configurations(int n, int maxskips, uint32_t* out, int cap) {
    if (n < 3)
        return 0;
    if (n > 30)
        return -1; /* 2^(n-2) configurations: the reference would not finish either */
    int      k     = 0;
    uint32_t order = 1u << (n - 2);
    for (uint32_t i = 1; i < order; ++i) {
        uint32_t mask = i << 1; /* bits 1..n-2: first and last token are never gaps */
        if (n - 2 >= maxskips && count_gap_runs(mask, n) > maxskips)
            continue;
        if (out && k < cap)
            out[k] = mask;
        ++k;
    }
    return k;
}

size_t oracle_skipgram_collapse(const uint8_t* ngram, size_t len, uint32_t mask, uint8_t* out) {
    /* src/pattern.cpp:886-908: every gap token becomes the single byte 0x03 (its continuation bytes are dropped) */
    size_t cur = 0;
    int    tok = 0;
    for (size_t i = 0; i < len; ++i) {
        int     gap = (tok < 31) && ((mask >> tok) & 1u); /* PatternPointer::isgap, src/pattern.cpp:139-143 */
        uint8_t c   = ngram[i];
        if (c < 128) {
            out[cur++] = gap ? 3 : c;
            ++tok;
        } else if (!gap) {
            out[cur++] = c;
        }
    }
    return cur;
}

/* ------------------------------------------------------------------------- */
/* The pattern store: stands in for PatternMap<uint32_t> / PatternMap<IndexedData>
 * (include/patternstore.h:937-1011): byte-string keys, hashed with Pattern::hash. */
typedef struct entry {
    uint64_t  hash;
    uint64_t  keyoff;
    uint32_t  len;
    uint32_t  count;
    uint16_t  n;       /* tokens */
    uint8_t   skipgram; /* category: 0 NGRAM, 1 SKIPGRAM */
    uint8_t   used;
    uint32_t  refcap;
    uint64_t* refs; /* indexed: (sentence << 16) | token, in insertion order */
} entry;

typedef struct store {
    entry*   tab;
    uint64_t cap; /* power of two */
    uint64_t size;
    uint8_t* arena;
    uint64_t arena_len, arena_cap;
} store;

struct oracle_model {
    store    st;
    int      indexed;
    uint64_t totaltokens, totaltypes;
    int      maxn, minn, hasskipgrams;
    int      hasflexgrams;
    int      npasses;
    uint64_t pass[128][4];
};

static void store_init(store* s, uint64_t cap) {
    s->cap       = cap;
    s->tab       = (entry*)calloc(cap, sizeof(entry));
    s->size      = 0;
    s->arena_cap = 1 << 16;
    s->arena     = (uint8_t*)malloc(s->arena_cap);
    s->arena_len = 0;
}
static void store_free(store* s) {
    if (s->tab) {
        for (uint64_t i = 0; i < s->cap; ++i)
            free(s->tab[i].refs);
    }
    free(s->tab);
    free(s->arena);
    memset(s, 0, sizeof *s);
}
static entry* store_find(const store* s, const uint8_t* key, uint32_t len, uint64_t h) {
    uint64_t i = h & (s->cap - 1);
    for (;;) {
        entry* e = &s->tab[i];
        if (!e->used)
            return NULL;
        if (e->hash == h && e->len == len && memcmp(s->arena + e->keyoff, key, len) == 0)
            return e;
        i = (i + 1) & (s->cap - 1);
    }
}
static void store_place(store* s, const entry* src) {
    uint64_t i = src->hash & (s->cap - 1);
    while (s->tab[i].used)
        i = (i + 1) & (s->cap - 1);
    s->tab[i] = *src;
}
static void store_grow(store* s) {
    entry*   old    = s->tab;
    uint64_t oldcap = s->cap;
    s->cap *= 2;
    s->tab = (entry*)calloc(s->cap, sizeof(entry));
    for (uint64_t i = 0; i < oldcap; ++i)
        if (old[i].used)
            store_place(s, &old[i]);
    free(old);
}
static entry* store_insert(store* s, const uint8_t* key, uint32_t len, uint64_t h, uint16_t n, uint8_t skipgram) {
    if ((s->size + 1) * 10 > s->cap * 7)
        store_grow(s);
    if (s->arena_len + len > s->arena_cap) {
        while (s->arena_len + len > s->arena_cap)
            s->arena_cap *= 2;
        s->arena = (uint8_t*)realloc(s->arena, s->arena_cap);
    }
    entry e;
    memset(&e, 0, sizeof e);
    e.hash     = h;
    e.keyoff   = s->arena_len;
    e.len      = len;
    e.n        = n;
    e.skipgram = skipgram;
    e.used     = 1;
    memcpy(s->arena + s->arena_len, key, len);
    s->arena_len += len;
    uint64_t i = h & (s->cap - 1);
    while (s->tab[i].used)
        i = (i + 1) & (s->cap - 1);
    s->tab[i] = e;
    ++s->size;
    return &s->tab[i];
}
/* erase every entry for which drop(e) is true; returns the number erased.
 * Stands in for the erase loops of prune() / pruneskipgrams() / prunebylength()
 * (include/patternmodel.h:2107-2128, :2167-2186, :2137-2158). */
typedef struct prune_rule {
    int      n;         /* 0: any size */
    int      maxn;      /* >0: size <= maxn (prunebylength) */
    int      category;  /* 0 any, 1 ngram only, 2 skipgram only */
    int64_t  threshold; /* -1: everything */
} prune_rule;
static int rule_hits(const prune_rule* r, const entry* e) {
    if (r->n && e->n != r->n)
        return 0;
    if (r->maxn && e->n > r->maxn)
        return 0;
    if (r->category == 1 && e->skipgram)
        return 0;
    if (r->category == 2 && !e->skipgram)
        return 0;
    return r->threshold < 0 || (int64_t)e->count < r->threshold;
}
static uint64_t store_prune(store* s, const prune_rule* r) {
    uint64_t hits = 0;
    for (uint64_t i = 0; i < s->cap; ++i)
        if (s->tab[i].used && rule_hits(r, &s->tab[i]))
            ++hits;
    if (!hits)
        return 0;
    store ns;
    ns.cap = s->cap;
    while (ns.cap > 1024 && (s->size - hits) * 4 < ns.cap)
        ns.cap /= 2;
    ns.tab       = (entry*)calloc(ns.cap, sizeof(entry));
    ns.size      = 0;
    ns.arena_cap = s->arena_cap;
    ns.arena     = (uint8_t*)malloc(ns.arena_cap);
    ns.arena_len = 0;
    for (uint64_t i = 0; i < s->cap; ++i) {
        entry* e = &s->tab[i];
        if (!e->used)
            continue;
        if (rule_hits(r, e)) {
            free(e->refs);
            e->refs = NULL;
            continue;
        }
        entry c  = *e;
        c.keyoff = ns.arena_len;
        memcpy(ns.arena + ns.arena_len, s->arena + e->keyoff, e->len);
        ns.arena_len += e->len;
        store_place(&ns, &c);
        ++ns.size;
        e->refs = NULL;
    }
    free(s->tab);
    free(s->arena);
    *s = ns;
    return hits;
}

/* valuehandler.add: BaseValueHandler (+1) include/datatypes.h:228-230; IndexedDataHandler (push_back) :283-289 */
static void entry_add(entry* e, int indexed, uint32_t sentence, uint16_t token) {
    if (indexed) {
        if (e->count == e->refcap) {
            e->refcap = e->refcap ? e->refcap * 2 : 2;
            e->refs   = (uint64_t*)realloc(e->refs, (size_t)e->refcap * sizeof(uint64_t));
        }
        e->refs[e->count] = ((uint64_t)sentence << 16) | token;
    }
    e->count += 1;
}

/* ------------------------------------------------------------------------- */
void oracle_options_default(oracle_options* o) {
    /* include/patternmodel.h:153-180 */
    memset(o, 0, sizeof *o);
    o->mintokens           = -1;
    o->mintokens_skipgrams = -1;
    o->mintokens_unigrams  = 1;
    o->minlength           = 1;
    o->maxlength           = 100;
    o->maxbackofflength    = 100;
    o->minskiptypes        = 2;
    o->maxskips            = 3;
    o->streamed            = 1;
}

/* One sentence as token spans over the corpus bytes. */
typedef struct span {
    uint64_t off;
    uint32_t len;
} span;

static int cmp_u64(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : (x > y);
}

/* add(patternpointer, ref): include/patternmodel.h:2059-2073 */
static entry* model_add(oracle_model* m, const uint8_t* key, uint32_t len, uint16_t n, uint8_t skipgram, uint32_t sentence, uint16_t token) {
    uint64_t h = oracle_pattern_hash(key, len);
    entry*   e = store_find(&m->st, key, len, h);
    if (!e)
        e = store_insert(&m->st, key, len, h, n, skipgram);
    entry_add(e, m->indexed, sentence, token);
    return e;
}
static int model_has(const oracle_model* m, const uint8_t* key, uint32_t len) {
    return store_find(&m->st, key, len, oracle_pattern_hash(key, len)) != NULL;
}


/* ------------------------------------------------------------------------- */
/* IndexedPatternModel::trainskipgrams (include/patternmodel.h:2969-3010) with computeskipgrams(pattern, ..., multiplerefs)
 * (:1370-1527) and the indexed pruneskipgrams (:3362-3383) / getskipcontent (:3029-3059), restated with CLEAN iteration
 * semantics: every surviving n-gram of size n is visited exactly once.  (The reference inserts into the unordered_map it is
 * iterating over; where a rehash happens mid-loop its own visiting order is undefined.  tests/ pin what agrees.) */
typedef struct tokindex {
    span*     toks;      /* every token of the corpus */
    uint64_t  ntoks;
    uint64_t* sentfirst; /* sentfirst[s] = index of the first token of sentence s (1-based); nsent+2 entries */
    uint32_t  nsent;
} tokindex;

static void tokindex_build(tokindex* ti, const uint8_t* corpus, size_t nbytes) {
    size_t cap = 1024, scap = 256;
    ti->toks      = (span*)malloc(cap * sizeof(span));
    ti->sentfirst = (uint64_t*)malloc(scap * sizeof(uint64_t));
    ti->ntoks     = 0;
    ti->nsent     = 0;
    size_t pos    = 0;
    while (pos < nbytes) {
        ++ti->nsent;
        if (ti->nsent + 2 >= scap) {
            scap *= 2;
            ti->sentfirst = (uint64_t*)realloc(ti->sentfirst, scap * sizeof(uint64_t));
        }
        ti->sentfirst[ti->nsent] = ti->ntoks;
        size_t start = pos;
        int    prevhigh = 0;
        for (;; ++pos) {
            uint8_t c = corpus[pos];
            if (!prevhigh && c == 0) {
                ++pos;
                break;
            }
            if (c < 128) {
                if (ti->ntoks == cap) {
                    cap *= 2;
                    ti->toks = (span*)realloc(ti->toks, cap * sizeof(span));
                }
                ti->toks[ti->ntoks].off = start;
                ti->toks[ti->ntoks].len = (uint32_t)(pos + 1 - start);
                ++ti->ntoks;
                start    = pos + 1;
                prevhigh = 0;
            } else {
                prevhigh = 1;
            }
        }
    }
    ti->sentfirst[ti->nsent + 1] = ti->ntoks;
}

static const uint8_t* g_cmp_corpus;
static int cmp_span(const void* a, const void* b) {
    const span* x = (const span*)a;
    const span* y = (const span*)b;
    uint32_t    l = x->len < y->len ? x->len : y->len;
    int         c = memcmp(g_cmp_corpus + x->off, g_cmp_corpus + y->off, l);
    if (c)
        return c;
    return x->len < y->len ? -1 : (x->len > y->len);
}

/* number of distinct raw token spans "first gap .. last gap" over the occurrences of one skipgram (getskipcontent().size()) */
static uint64_t skip_types(const tokindex* ti, const uint8_t* corpus, const entry* e, uint32_t mask) {
    int n = e->n, head = 0, tail = 0;
    while (head < n && !((mask >> head) & 1u))
        ++head; /* maskheadskip(reversemask(mask)) : leading non-gap tokens */
    while (tail < n && !((mask >> (n - 1 - tail)) & 1u))
        ++tail;
    span* v = (span*)malloc((size_t)(e->count ? e->count : 1) * sizeof(span));
    for (uint32_t k = 0; k < e->count; ++k) {
        uint32_t sentence = (uint32_t)(e->refs[k] >> 16), token = (uint32_t)(e->refs[k] & 0xFFFF);
        uint64_t first = ti->sentfirst[sentence] + token + (uint64_t)head, last = ti->sentfirst[sentence] + token + (uint64_t)(n - tail) - 1;
        v[k].off = ti->toks[first].off;
        v[k].len = (uint32_t)(ti->toks[last].off + ti->toks[last].len - ti->toks[first].off);
    }
    g_cmp_corpus = corpus;
    qsort(v, e->count, sizeof(span), cmp_span);
    uint64_t types = e->count ? 1 : 0;
    for (uint32_t k = 1; k < e->count; ++k)
        if (cmp_span(&v[k - 1], &v[k]) != 0)
            ++types;
    free(v);
    return types;
}

/* gap mask of a stored skipgram key (bytes with 0x03 gap markers): Pattern::getmask, src/pattern.cpp:189-210 */
static uint32_t key_mask(const uint8_t* key, uint32_t len) {
    uint32_t mask = 0;
    int      tok = 0, prevhigh = 0;
    for (uint32_t i = 0; i < len; ++i) {
        uint8_t c = key[i];
        if (c < 128) {
            if (!prevhigh && c == 3 && tok < 31)
                mask |= 1u << tok;
            ++tok;
            prevhigh = 0;
        } else {
            prevhigh = 1;
        }
    }
    return mask;
}

static int oracle_trainskipgrams(oracle_model* m, const oracle_options* o, const uint8_t* corpus, size_t nbytes) {
    tokindex ti;
    tokindex_build(&ti, corpus, nbytes);
    uint32_t masks[4096];
    uint8_t  skipkey[1024];
    int      rc = 0;
    for (int n = 3; n <= o->maxlength; ++n) {
        int nmasks = oracle_skip_configurations(n, o->maxskips, masks, 4096);
        if (nmasks < 0 || nmasks > 4096) {
            rc = fail("oracle: too many skip configurations");
            break;
        }
        /* snapshot of the surviving n-grams of this size (key bytes + occurrence lists) */
        uint64_t cnt = 0;
        for (uint64_t i = 0; i < m->st.cap; ++i)
            if (m->st.tab[i].used && m->st.tab[i].n == n && !m->st.tab[i].skipgram)
                ++cnt;
        uint64_t* idx = (uint64_t*)malloc((cnt + 1) * sizeof(uint64_t));
        uint64_t  k   = 0;
        for (uint64_t i = 0; i < m->st.cap; ++i)
            if (m->st.tab[i].used && m->st.tab[i].n == n && !m->st.tab[i].skipgram)
                idx[k++] = i;
        /* copy what is needed: inserting may grow/move the table */
        typedef struct snap { uint8_t* key; uint32_t len; uint32_t count; uint64_t* refs; } snap;
        snap* sn = (snap*)malloc((cnt + 1) * sizeof(snap));
        for (uint64_t q = 0; q < cnt; ++q) {
            entry* e    = &m->st.tab[idx[q]];
            sn[q].len   = e->len;
            sn[q].count = e->count;
            sn[q].key   = (uint8_t*)malloc(e->len);
            memcpy(sn[q].key, m->st.arena + e->keyoff, e->len);
            sn[q].refs = (uint64_t*)malloc((size_t)(e->count ? e->count : 1) * sizeof(uint64_t));
            memcpy(sn[q].refs, e->refs, (size_t)e->count * sizeof(uint64_t));
        }
        free(idx);
        uint64_t foundskipgrams = 0;
        for (uint64_t q = 0; q < cnt; ++q) {
            for (int mi = 0; mi < nmasks; ++mi) {
                size_t sl = oracle_skipgram_collapse(sn[q].key, sn[q].len, masks[mi], skipkey);
                if (!model_has(m, skipkey, (uint32_t)sl))
                    ++foundskipgrams; /* :1504-1505 */
                for (uint32_t r = 0; r < sn[q].count; ++r)
                    model_add(m, skipkey, (uint32_t)sl, (uint16_t)n, 1, (uint32_t)(sn[q].refs[r] >> 16), (uint16_t)(sn[q].refs[r] & 0xFFFF)); /* :1508-1512 */
            }
        }
        for (uint64_t q = 0; q < cnt; ++q) {
            free(sn[q].key);
            free(sn[q].refs);
        }
        free(sn);
        if (!foundskipgrams)
            break; /* " None found" :2991-2993 */
        m->hasskipgrams = 1;
        prune_rule r      = {n, 0, 0, o->mintokens};
        uint64_t   pruned = store_prune(&m->st, &r); /* :2999 */
        uint64_t   extra  = 0;
        if (o->minskiptypes > 1) { /* :3362-3383 */
            /* mark, then erase: a skipgram whose gaps were filled by fewer than MINSKIPTYPES distinct contents goes */
            for (uint64_t i = 0; i < m->st.cap; ++i) {
                entry* e = &m->st.tab[i];
                if (e->used && e->skipgram && e->n == n) {
                    uint32_t mask = key_mask(m->st.arena + e->keyoff, e->len);
                    if (skip_types(&ti, corpus, e, mask) < (uint64_t)o->minskiptypes) {
                        e->count = 0; /* below any threshold: picked up by the prune below */
                        ++extra;
                    }
                }
            }
            prune_rule z = {n, 0, 2, 1};
            store_prune(&m->st, &z);
        }
        if (m->npasses < 128) { /* recorded as extra "passes": n, 0 n-grams, found skipgrams, pruned */
            m->pass[m->npasses][0] = (uint64_t)n;
            m->pass[m->npasses][1] = 0;
            m->pass[m->npasses][2] = foundskipgrams;
            m->pass[m->npasses][3] = pruned + extra;
            ++m->npasses;
        }
    }
    free(ti.toks);
    free(ti.sentfirst);
    return rc;
}

int oracle_train(const uint8_t* corpus_in, size_t nbytes_in, const oracle_options* opt_in, oracle_model** out) {
    oracle_options o = *opt_in;
    *out             = NULL;
    /* include/patternmodel.h:883-888 */
    if (o.mintokens == -1)
        o.mintokens = 2;
    if (o.mintokens == 0)
        o.mintokens = 1;
    if (o.mintokens_skipgrams < o.mintokens)
        o.mintokens_skipgrams = o.mintokens;
    if (o.doskipgrams && o.doskipgrams_exhaustive)
        return fail("Both DOSKIPGRAMS as well as DOSKIPGRAMS_EXHAUSTIVE are set"); /* :958-963 */
    if (o.doskipgrams && !o.indexed)
        return fail("Can not compute skipgrams on unindexed model (except exhaustively during train() )"); /* :1554-1561 */
    if (o.maxlength > 127)
        return fail("oracle: MAXLENGTH > 127 not supported");
    if ((o.minlength > 1 || o.mintokens == 1) && o.mintokens_unigrams > o.mintokens)
        return fail("oracle: the iter_unigramsonly pre-pass (patternmodel.h:918-920) is not restated");
    if (nbytes_in == 0)
        return fail("Attempting to read pattern from file, but file is empty?"); /* src/pattern.cpp:520-523 (and :545-549) */

    /* Sentence source.  Streamed (Pattern(istream), src/pattern.cpp:483-587): when the last sentence has no 0x00,
     * the stage-1 length count includes the failed read, so stage 2 stores the last byte twice before the added
     * end marker (checked against the reference binary: tests/golden/quirk_noeos*).  Preloaded (IndexedCorpus,
     * src/pattern.cpp:2135-2154) simply stops at the end of the buffer. */
    size_t   nbytes = nbytes_in;
    uint8_t* corpus = (uint8_t*)malloc(nbytes_in + 2);
    memcpy(corpus, corpus_in, nbytes_in);
    {
        int ends_with_delim = 0;
        if (nbytes_in >= 1 && corpus_in[nbytes_in - 1] == 0)
            ends_with_delim = (nbytes_in == 1) || (corpus_in[nbytes_in - 2] < 128);
        if (!ends_with_delim) {
            if (o.streamed)
                corpus[nbytes++] = corpus_in[nbytes_in - 1];
            corpus[nbytes++] = 0;
        }
    }

    oracle_model* m = (oracle_model*)calloc(1, sizeof *m);
    store_init(&m->st, 1024);
    m->indexed = o.indexed;
    m->maxn    = 0;
    m->minn    = 999; /* include/patternmodel.h:647-648 */

    span*    toks    = NULL;
    size_t   tokscap = 0;
    uint32_t masks[4096];
    uint8_t  skipkey[1024];
    uint64_t prevsize   = 0;
    const int singlepass = (o.mintokens == 1); /* :1062-1072: MINTOKENS==1 extracts every length in one scan */
    int      rc         = 0;

    for (int n = 1; n <= o.maxlength; ++n) {
        int nmasks = 0;
        uint64_t foundskipgrams = 0;
        uint32_t sentence       = 0;
        size_t   pos            = 0;
        while (pos < nbytes) { /* :1030 */
            ++sentence;
            /* tokenise one sentence: delimiter = 0x00 whose predecessor is not a continuation byte */
            size_t ntok = 0, start = pos;
            int    prevhigh = 0;
            for (;; ++pos) {
                uint8_t c = corpus[pos];
                if (!prevhigh && c == 0) {
                    ++pos;
                    break;
                }
                if (c < 128) {
                    if (ntok == tokscap) {
                        tokscap = tokscap ? tokscap * 2 : 256;
                        toks    = (span*)realloc(toks, tokscap * sizeof(span));
                    }
                    toks[ntok].off = start;
                    toks[ntok].len = (uint32_t)(pos + 1 - start);
                    ++ntok;
                    start    = pos + 1;
                    prevhigh = 0;
                } else {
                    prevhigh = 1;
                }
            }
            if (ntok == 0)
                continue; /* :1042-1045 empty lines are numbered but skipped */
            if (n == 1)
                m->totaltokens += ntok; /* :1047-1048 */

            int lo = n, hi = n;
            if (singlepass) { /* line.subngrams(ngrams, MINLENGTH, MAXLENGTH): src/pattern.cpp:1363-1374 */
                lo = o.minlength;
                hi = o.maxlength < (int)ntok ? o.maxlength : (int)ntok;
                if (lo > (int)ntok)
                    continue;
            }
            for (int len = lo; len <= hi; ++len) {
                if ((size_t)len > ntok)
                    break;
                if (o.doskipgrams_exhaustive && (len >= 3)) {
                    nmasks = oracle_skip_configurations(len, o.maxskips, masks, 4096); /* :1021-1022, :1388-1389 */
                    if (nmasks < 0 || nmasks > 4096) {
                        rc = fail("oracle: too many skip configurations");
                        goto done;
                    }
                } else {
                    nmasks = 0;
                }
                for (size_t i = 0; i + len <= ntok; ++i) { /* :1078 */
                    const uint8_t* key    = corpus + toks[i].off;
                    uint32_t       keylen = (uint32_t)(toks[i + len - 1].off + toks[i + len - 1].len - toks[i].off);
                    int            found  = 1;
                    /* :1094-1104 secondary unigram threshold */
                    if (o.mintokens_unigrams > o.mintokens && (len > 1 || singlepass)) {
                        for (int j = 0; j < len && found; ++j)
                            if ((int64_t)oracle_model_count(m, corpus + toks[i + j].off, toks[i + j].len) < o.mintokens_unigrams)
                                found = 0;
                    }
                    /* :1139-1152 lookback: all sub-n-grams of size min(n-1, MAXBACKOFFLENGTH) must still be in the model */
                    int subsok = 1;
                    if (len > 1 && o.mintokens > 1) {
                        int b = len - 1;
                        for (int j = 0; j + b <= len; ++j) {
                            uint32_t sl = (uint32_t)(toks[i + j + b - 1].off + toks[i + j + b - 1].len - toks[i + j].off);
                            if (!model_has(m, corpus + toks[i + j].off, sl)) {
                                subsok = 0;
                                break;
                            }
                        }
                    }
                    if (found && len > 1 && o.mintokens > 1) {
                        int b = len - 1;
                        if (b > o.maxbackofflength)
                            b = o.maxbackofflength;
                        if (b == len - 1) {
                            found = subsok;
                        } else {
                            for (int j = 0; j + b <= len; ++j) {
                                uint32_t sl = (uint32_t)(toks[i + j + b - 1].off + toks[i + j + b - 1].len - toks[i + j].off);
                                if (!model_has(m, corpus + toks[i + j].off, sl)) {
                                    found = 0;
                                    break;
                                }
                            }
                        }
                    }
                    if (found)
                        model_add(m, key, keylen, (uint16_t)len, 0, sentence, (uint16_t)i); /* :1155-1161 */
                    /* :1163-1171 exhaustive skipgrams are attempted for EVERY window, found or not; computeskipgrams
                     * (:1370-1527) then validates each mask against the two RAW (n-1)-grams, because the sub-slices of a
                     * masked PatternPointer recompute their mask from the bytes (src/pattern.cpp:855) -> never a gap. */
                    if (nmasks > 0 && (len >= 3 || o.mintokens == 1)) {
                        int valid = (o.mintokens_skipgrams == 1) ? 1 : subsok;
                        if (o.mintokens_skipgrams != 1 && o.mintokens == 1) {
                            /* single-pass with a skipgram threshold > 1: the lookup happens against the model as built so far */
                            int b = len - 1;
                            valid = 1;
                            for (int j = 0; j + b <= len; ++j) {
                                uint32_t sl = (uint32_t)(toks[i + j + b - 1].off + toks[i + j + b - 1].len - toks[i + j].off);
                                if (!model_has(m, corpus + toks[i + j].off, sl)) {
                                    valid = 0;
                                    break;
                                }
                            }
                        }
                        if (valid) {
                            if (keylen > sizeof skipkey) {
                                rc = fail("oracle: pattern too long");
                                goto done;
                            }
                            for (int k = 0; k < nmasks; ++k) {
                                size_t sl = oracle_skipgram_collapse(key, keylen, masks[k], skipkey);
                                if (!model_has(m, skipkey, (uint32_t)sl))
                                    ++foundskipgrams; /* :1504-1505 */
                                model_add(m, skipkey, (uint32_t)sl, (uint16_t)len, 1, sentence, (uint16_t)i);
                            }
                        }
                    }
                }
            }
        }

        /* :1181-1194 */
        uint64_t foundngrams = m->st.size - foundskipgrams - prevsize;
        if (foundskipgrams)
            m->hasskipgrams = 1; /* :1168-1169 */
        if (foundngrams || foundskipgrams) {
            if (n > m->maxn)
                m->maxn = n;
            if (n < m->minn)
                m->minn = n;
        } else {
            break; /* "None found" */
        }
        /* :1199-1209 */
        if (o.mintokens > 1 && n == 1) {
            m->totaltypes = m->st.size;
        } else if (o.mintokens == 1 && o.minlength == 1) {
            uint64_t types = 0; /* totalwordtypesingroup(NGRAM, 1): distinct unigram n-grams */
            for (uint64_t i = 0; i < m->st.cap; ++i)
                if (m->st.tab[i].used && m->st.tab[i].n == 1 && !m->st.tab[i].skipgram)
                    ++types;
            m->totaltypes = types;
        }
        /* :1210-1230 */
        uint64_t   pruned;
        prune_rule r = {0, 0, 0, 0};
        if (singlepass) {
            r.threshold = o.mintokens;
            pruned      = store_prune(&m->st, &r);
        } else {
            r.n         = n;
            r.threshold = o.mintokens;
            pruned      = store_prune(&m->st, &r);
            if (!o.doskipgrams && !o.doskipgrams_exhaustive && n - 1 >= 1 && n - 1 < o.minlength && n - 1 != o.maxbackofflength &&
                !(n - 1 == 1 && o.mintokens_unigrams > o.mintokens)) {
                prune_rule all = {n - 1, 0, 0, -1};
                store_prune(&m->st, &all);
            }
        }
        /* :1233-1243.  train() is PatternModel code, so the call binds to PatternModel::pruneskipgrams(unsigned, int, int)
         * (:2167-2186) for indexed models too: IndexedPatternModel::pruneskipgrams(int, int, int) (:3362) has a different
         * signature and does not override it.  That version returns early when minskiptypes <= 1 and otherwise applies
         * only the occurrence threshold. */
        if (foundskipgrams && o.minskiptypes > 1) {
            prune_rule sk = {singlepass ? 0 : n, 0, 2, o.mintokens_skipgrams};
            pruned += store_prune(&m->st, &sk);
        }
        if (m->npasses < 128) {
            m->pass[m->npasses][0] = (uint64_t)n;
            m->pass[m->npasses][1] = foundngrams;
            m->pass[m->npasses][2] = foundskipgrams;
            m->pass[m->npasses][3] = pruned;
            ++m->npasses;
        }
        if (o.mintokens == 1)
            break; /* :1246-1247 */
        prevsize = m->st.size; /* :1269 */
    }

    if (o.doskipgrams && !o.doskipgrams_exhaustive) { /* :1271-1273 */
        rc = oracle_trainskipgrams(m, &o, corpus, nbytes);
        if (rc)
            goto done;
    }
    if (o.mintokens == 1) { /* :1274-1277 postread: maxn/minn/hasskipgrams from the stored patterns */
        for (uint64_t i = 0; i < m->st.cap; ++i) {
            entry* e = &m->st.tab[i];
            if (!e->used)
                continue;
            if (e->n > m->maxn)
                m->maxn = e->n;
            if (e->n < m->minn)
                m->minn = e->n;
            if (e->skipgram)
                m->hasskipgrams = 1;
        }
    }
    if (o.maxbackofflength < o.minlength) { /* :1278-1280 */
        prune_rule r = {o.maxbackofflength, 0, 0, -1};
        store_prune(&m->st, &r);
    }
    if (o.minlength > 1 && o.mintokens_unigrams > o.mintokens) { /* :1281-1284 */
        prune_rule r = {1, 0, 0, -1};
        store_prune(&m->st, &r);
    }
    if (o.minlength > 1 && (o.doskipgrams || o.doskipgrams_exhaustive)) { /* :1337-1341 prunebylength */
        prune_rule r = {0, o.minlength - 1, 0, -1};
        store_prune(&m->st, &r);
    }
    if (o.indexed) { /* posttrain: :2699-2705, IndexedData::sort include/datatypes.h:170-172 */
        for (uint64_t i = 0; i < m->st.cap; ++i)
            if (m->st.tab[i].used && m->st.tab[i].count > 1)
                qsort(m->st.tab[i].refs, m->st.tab[i].count, sizeof(uint64_t), cmp_u64);
    }
done:
    free(toks);
    free(corpus);
    if (rc) {
        oracle_model_free(m);
        return rc;
    }
    *out = m;
    return 0;
}

/* ------------------------------------------------------------------------- */
static int modelfile_walk(const uint8_t* d, size_t n, uint64_t hdr[7], uint8_t* keys, uint64_t* key_off, uint32_t* counts, uint32_t* ref_sentence, uint16_t* ref_token,
                          uint64_t* ref_off);
/* Shape of a stored key: tokens (Pattern::n -> datasize, src/pattern.cpp:74-97: one per byte < 128) and category
 * (datacategory, src/pattern.cpp:23-43: the FIRST skip (3) / flex (4) token decides). */
static void key_shape(const uint8_t* key, uint32_t len, uint16_t* n_out, uint8_t* cat_out) {
    uint16_t n = 0;
    uint8_t  cat = 0;
    int      tokstart = 1;
    for (uint32_t i = 0; i < len; ++i) {
        uint8_t c = key[i];
        if (c < 128) {
            if (tokstart && cat == 0 && c == 3)
                cat = 1;
            if (tokstart && cat == 0 && c == 4)
                cat = 2;
            ++n;
            tokstart = 1;
        } else {
            tokstart = 0;
        }
    }
    *n_out   = n;
    *cat_out = cat;
}

/* postread (include/patternmodel.h:572-588): maxn / minn / hasskipgrams from the stored patterns */
static void model_postread(oracle_model* m) {
    for (uint64_t i = 0; i < m->st.cap; ++i) {
        entry* e = &m->st.tab[i];
        if (!e->used)
            continue;
        if (e->n > m->maxn)
            m->maxn = e->n;
        if (e->n < m->minn)
            m->minn = e->n;
        if (e->skipgram == 1)
            m->hasskipgrams = 1;
    }
}

/* PatternModel::load (include/patternmodel.h:781-861) + PatternMapStore::read (include/patternstore.h:555-619):
 * read a model file of type 10/20 AS an unindexed (load_indexed=0) or indexed (1) model, applying the options as filters. */
int oracle_model_load(const uint8_t* file, size_t nbytes, const oracle_load_options* lo, const oracle_model* constrain, oracle_model** out) {
    *out = NULL;
    uint64_t hdr[7];
    if (modelfile_walk(file, nbytes, hdr, NULL, NULL, NULL, NULL, NULL, NULL))
        return 1;
    const int filetype = (int)hdr[0];
    uint64_t  np = hdr[4], kb = hdr[5], nr = hdr[6];
    uint8_t*  keys = (uint8_t*)malloc(kb + 1);
    uint64_t* ko   = (uint64_t*)malloc((np + 1) * sizeof(uint64_t));
    uint32_t* cnt  = (uint32_t*)malloc((np + 1) * sizeof(uint32_t));
    uint32_t* rs   = (uint32_t*)malloc((nr + 1) * sizeof(uint32_t));
    uint16_t* rt   = (uint16_t*)malloc((nr + 1) * sizeof(uint16_t));
    uint64_t* ro   = (uint64_t*)malloc((np + 1) * sizeof(uint64_t));
    modelfile_walk(file, nbytes, NULL, keys, ko, cnt, rs, rt, ro);
    oracle_model* m = (oracle_model*)calloc(1, sizeof *m);
    store_init(&m->st, 1024);
    m->indexed     = lo->load_indexed;
    m->maxn        = 0;
    m->minn        = 999;
    m->totaltokens = hdr[2]; /* :815-816 */
    m->totaltypes  = hdr[3];
    int64_t mintokens = lo->mintokens == -1 ? 0 : lo->mintokens; /* patternstore.h:565-566 */
    for (uint64_t i = 0; i < np; ++i) {
        const uint8_t* k = keys + ko[i];
        uint32_t       l = (uint32_t)(ko[i + 1] - ko[i]);
        uint16_t       n;
        uint8_t        cat;
        key_shape(k, l, &n, &cat);
        if ((!lo->dongrams && cat == 0) || (!lo->doskipgrams && cat == 1) || (!lo->doflexgrams && cat == 2))
            continue; /* patternstore.h:574-578 */
        if ((int)n < lo->minlength || (int)n > lo->maxlength)
            continue; /* :585 */
        if ((int64_t)cnt[i] < mintokens)
            continue; /* :586 */
        if (constrain && !model_has(constrain, k, l))
            continue;
        uint64_t h = oracle_pattern_hash(k, l);
        if (store_find(&m->st, k, l, h))
            continue;
        entry* e = store_insert(&m->st, k, l, h, n, cat);
        if (lo->doreset)
            continue; /* :588-589: a fresh value */
        if (m->indexed) {
            /* 20 -> 20 keeps the occurrence list; 10 -> 20 "will load the patterns but lose all the counts" (:833-837) */
            if (filetype == 20)
                for (uint64_t j = ro[i]; j < ro[i + 1]; ++j)
                    entry_add(e, 1, rs[j], rt[j]);
        } else {
            e->count = cnt[i]; /* 20 -> 10: IndexedDataHandler::convertto -> count (:827-832) */
        }
    }
    model_postread(m); /* :860 */
    free(keys);
    free(ko);
    free(cnt);
    free(rs);
    free(rt);
    free(ro);
    *out = m;
    return 0;
}

/* a model holding the given patterns with zero counts and the given totals: the constraint side of
 * oracle_train_constrained when the caller has no model file at hand */
int oracle_model_from_keys(const uint8_t* keys, const uint64_t* key_off, uint64_t npatterns, uint64_t totaltokens, uint64_t totaltypes, int indexed, oracle_model** out) {
    oracle_model* m = (oracle_model*)calloc(1, sizeof *m);
    store_init(&m->st, 1024);
    m->indexed     = indexed;
    m->maxn        = 0;
    m->minn        = 999;
    m->totaltokens = totaltokens;
    m->totaltypes  = totaltypes;
    for (uint64_t i = 0; i < npatterns; ++i) {
        const uint8_t* k = keys + key_off[i];
        uint32_t       l = (uint32_t)(key_off[i + 1] - key_off[i]);
        uint16_t       n;
        uint8_t        cat;
        key_shape(k, l, &n, &cat);
        uint64_t h = oracle_pattern_hash(k, l);
        if (!store_find(&m->st, k, l, h))
            store_insert(&m->st, k, l, h, n, cat);
    }
    model_postread(m);
    *out = m;
    return 0;
}

/* IndexedPatternModel::computeflexgrams_fromskipgrams (include/patternmodel.h:3724-3744): every skipgram of the model is abstracted to its
 * flexgram (Pattern::toflexgram, src/pattern.cpp:145-180: each run of gap tokens 0x03 becomes ONE dynamic gap 0x04) and hands all its
 * occurrences to it; returns the number of flexgrams that were not in the model before.  Restated with CLEAN iteration semantics: the
 * skipgrams are visited once each (the reference inserts into the unordered_map it iterates over).  The reference appends the
 * occurrences in iteration order and never sorts them; here every flexgram's list ends up ascending -- compare them as multisets. */
static uint32_t key_toflexgram(const uint8_t* key, uint32_t len, uint8_t* out) {
    uint32_t j = 0;
    int      skipgap = 0, prevhigh = 0;
    for (uint32_t i = 0; i < len; ++i) {
        uint8_t c = key[i];
        if (!prevhigh && c == 3) {
            if (!skipgap) {
                out[j++] = 4;
                skipgap  = 1;
            }
        } else {
            out[j++] = c;
            skipgap  = 0;
        }
        prevhigh = c >= 128;
    }
    return j;
}
int64_t oracle_computeflexgrams_fromskipgrams(oracle_model* m) {
    if (!m->indexed) {
        fail("computeflexgrams_fromskipgrams needs an indexed model");
        return -1;
    }
    /* snapshot of the skipgrams: (key copy, refs copy), since inserting may move the table */
    uint64_t nsk = 0;
    for (uint64_t i = 0; i < m->st.cap; ++i)
        if (m->st.tab[i].used && m->st.tab[i].skipgram == 1)
            ++nsk;
    uint8_t** keys  = (uint8_t**)malloc((nsk + 1) * sizeof(uint8_t*));
    uint32_t* lens  = (uint32_t*)malloc((nsk + 1) * sizeof(uint32_t));
    uint64_t** refs = (uint64_t**)malloc((nsk + 1) * sizeof(uint64_t*));
    uint32_t* cnts  = (uint32_t*)malloc((nsk + 1) * sizeof(uint32_t));
    uint16_t* ns    = (uint16_t*)malloc((nsk + 1) * sizeof(uint16_t));
    uint64_t  k     = 0;
    for (uint64_t i = 0; i < m->st.cap; ++i) {
        entry* e = &m->st.tab[i];
        if (!e->used || e->skipgram != 1)
            continue;
        keys[k] = (uint8_t*)malloc(e->len + 1);
        memcpy(keys[k], m->st.arena + e->keyoff, e->len);
        lens[k] = e->len;
        cnts[k] = e->count;
        ns[k]   = e->n;
        refs[k] = (uint64_t*)malloc(((size_t)e->count + 1) * sizeof(uint64_t));
        memcpy(refs[k], e->refs, (size_t)e->count * sizeof(uint64_t));
        ++k;
    }
    int64_t found = 0;
    uint8_t flex[1024];
    for (uint64_t i = 0; i < nsk; ++i) {
        if (lens[i] > sizeof flex)
            continue;
        uint32_t fl = key_toflexgram(keys[i], lens[i], flex);
        uint64_t h  = oracle_pattern_hash(flex, fl);
        entry*   e  = store_find(&m->st, flex, fl, h);
        if (!e) {
            uint16_t n;
            uint8_t  cat;
            key_shape(flex, fl, &n, &cat);
            e = store_insert(&m->st, flex, fl, h, n, 2);
            ++found;
        }
        for (uint32_t j = 0; j < cnts[i]; ++j)
            entry_add(e, 1, (uint32_t)(refs[i][j] >> 16), (uint16_t)(refs[i][j] & 0xFFFF));
    }
    for (uint64_t i = 0; i < m->st.cap; ++i)
        if (m->st.tab[i].used && m->st.tab[i].skipgram == 2 && m->st.tab[i].count > 1)
            qsort(m->st.tab[i].refs, m->st.tab[i].count, sizeof(uint64_t), cmp_u64);
    if (found)
        m->hasflexgrams = 1;
    for (uint64_t i = 0; i < nsk; ++i) {
        free(keys[i]);
        free(refs[i]);
    }
    free(keys);
    free(lens);
    free(refs);
    free(cnts);
    free(ns);
    return found;
}
int oracle_model_hasflexgrams(const oracle_model* m) {
    return m->hasflexgrams;
}

/* PatternModel::train with constrainbymodel != NULL (include/patternmodel.h:880-1345): ONE scan of the corpus
 * (:1064-1072 subngrams(MINLENGTH, MAXLENGTH)), a window is counted iff the constraint model has it (:1088-1089),
 * then prune(MINTOKENS, 0) (:1211-1218) and stop (:1246-1247).
 *   inplace == 0: `constrain` is another model (the CLI's -j: a PatternSetModel); the result starts empty;
 *                 totaltypes/totaltokens start from the constraint model's (:892-895) and the corpus tokens are ADDED (:1047-1048).
 *   inplace == 1: constrainbymodel == this (the CLI's -I / stage 2 of -2): `constrain` is the model itself, loaded with
 *                 DORESET; totals restart at 0 (:889-891), found = the whole loaded model (prevsize = 0, :970-971),
 *                 totaltypes = size() for MINTOKENS > 1 (:1199-1201).
 * The constraint model is not modified; the result is a new model. */
int oracle_train_constrained(const uint8_t* corpus_in, size_t nbytes_in, const oracle_options* opt_in, const oracle_model* constrain, int inplace, oracle_model** out) {
    oracle_options o = *opt_in;
    *out             = NULL;
    if (o.mintokens == -1)
        o.mintokens = 2;
    if (o.mintokens == 0)
        o.mintokens = 1;
    if (o.doskipgrams || o.doskipgrams_exhaustive)
        return fail("oracle: skipgrams under a constraint model are not restated");
    if (o.mintokens_unigrams > o.mintokens)
        return fail("oracle: the secondary unigram threshold under a constraint model is not restated");
    if (o.maxlength > 127)
        return fail("oracle: MAXLENGTH > 127 not supported");
    if (nbytes_in == 0)
        return fail("Attempting to read pattern from file, but file is empty?");
    size_t   nbytes = nbytes_in;
    uint8_t* corpus = (uint8_t*)malloc(nbytes_in + 2);
    memcpy(corpus, corpus_in, nbytes_in);
    {
        int ends_with_delim = 0;
        if (nbytes_in >= 1 && corpus_in[nbytes_in - 1] == 0)
            ends_with_delim = (nbytes_in == 1) || (corpus_in[nbytes_in - 2] < 128);
        if (!ends_with_delim) {
            if (o.streamed)
                corpus[nbytes++] = corpus_in[nbytes_in - 1];
            corpus[nbytes++] = 0;
        }
    }
    oracle_model* m = (oracle_model*)calloc(1, sizeof *m);
    store_init(&m->st, 1024);
    m->indexed = o.indexed;
    m->maxn    = 0;
    m->minn    = 999;
    if (inplace) {
        /* the loaded model itself: every pattern present with a reset value; maxn/minn as postread left them */
        for (uint64_t i = 0; i < constrain->st.cap; ++i) {
            const entry* e = &constrain->st.tab[i];
            if (e->used)
                store_insert(&m->st, constrain->st.arena + e->keyoff, e->len, e->hash, e->n, e->skipgram);
        }
        m->maxn         = constrain->maxn;
        m->minn         = constrain->minn;
        m->hasskipgrams = constrain->hasskipgrams;
    } else {
        m->totaltypes  = constrain->totaltypes; /* :892-895 */
        m->totaltokens = constrain->totaltokens;
    }
    span*    toks    = NULL;
    size_t   tokscap = 0;
    uint32_t sentence = 0;
    size_t   pos      = 0;
    while (pos < nbytes) {
        ++sentence;
        size_t ntok = 0, start = pos;
        int    prevhigh = 0;
        for (;; ++pos) {
            uint8_t c = corpus[pos];
            if (!prevhigh && c == 0) {
                ++pos;
                break;
            }
            if (c < 128) {
                if (ntok == tokscap) {
                    tokscap = tokscap ? tokscap * 2 : 256;
                    toks    = (span*)realloc(toks, tokscap * sizeof(span));
                }
                toks[ntok].off = start;
                toks[ntok].len = (uint32_t)(pos + 1 - start);
                ++ntok;
                start    = pos + 1;
                prevhigh = 0;
            } else {
                prevhigh = 1;
            }
        }
        if (ntok == 0)
            continue;
        m->totaltokens += ntok; /* n == 1 && !continued, :1047-1048 */
        int lo = o.minlength, hi = o.maxlength < (int)ntok ? o.maxlength : (int)ntok;
        if (lo > (int)ntok)
            continue;
        for (int len = lo; len <= hi; ++len)
            for (size_t i = 0; i + len <= ntok; ++i) {
                const uint8_t* key    = corpus + toks[i].off;
                uint32_t       keylen = (uint32_t)(toks[i + len - 1].off + toks[i + len - 1].len - toks[i].off);
                if (!model_has(constrain, key, keylen))
                    continue; /* :1088-1089 */
                model_add(m, key, keylen, (uint16_t)len, 0, sentence, (uint16_t)i);
            }
    }
    uint64_t foundngrams = m->st.size; /* :1182 with prevsize = 0 (fresh model, or :970-971) */
    if (foundngrams) {                 /* :1184-1188 with n == 1 */
        if (1 > m->maxn)
            m->maxn = 1;
        if (1 < m->minn)
            m->minn = 1;
    }
    if (inplace) { /* :1199-1209 */
        if (o.mintokens > 1) {
            m->totaltypes = m->st.size;
        } else if (o.minlength == 1) {
            uint64_t types = 0;
            for (uint64_t i = 0; i < m->st.cap; ++i)
                if (m->st.tab[i].used && m->st.tab[i].n == 1 && !m->st.tab[i].skipgram)
                    ++types;
            m->totaltypes = types;
        }
    }
    uint64_t pruned = 0;
    if (foundngrams) { /* "None found" breaks before the prune (:1189-1194) */
        prune_rule r = {0, 0, 0, o.mintokens};
        pruned       = store_prune(&m->st, &r); /* :1217 */
        m->pass[0][0] = 1;
        m->pass[0][1] = foundngrams;
        m->pass[0][2] = 0;
        m->pass[0][3] = pruned;
        m->npasses    = 1;
    }
    if (o.mintokens == 1)
        model_postread(m); /* :1274-1277 */
    if (o.maxbackofflength < o.minlength) { /* :1278-1280 */
        prune_rule r = {o.maxbackofflength, 0, 0, -1};
        store_prune(&m->st, &r);
    }
    if (o.indexed) {
        for (uint64_t i = 0; i < m->st.cap; ++i)
            if (m->st.tab[i].used && m->st.tab[i].count > 1)
                qsort(m->st.tab[i].refs, m->st.tab[i].count, sizeof(uint64_t), cmp_u64);
    }
    if (m->totaltypes == 0 && m->st.size > 0 && !(inplace && o.mintokens == 1 && o.minlength == 1)) {
        /* types() (:1700-1704) falls back to totalwordtypesingroup(0, 0) when totaltypes was never set (in-place rebuild with
         * MINTOKENS == 1 and MINLENGTH > 1): the distinct word types covered by the patterns left (:1953-1975); write() stores it.
         * Not after the MINTOKENS == 1 / MINLENGTH == 1 branch above: that call filled the coverage cache, so a zero stays zero. */
        oracle_model* seen = (oracle_model*)calloc(1, sizeof *seen);
        store_init(&seen->st, 1024);
        for (uint64_t i = 0; i < m->st.cap; ++i) {
            const entry* e = &m->st.tab[i];
            if (!e->used)
                continue;
            const uint8_t* k = m->st.arena + e->keyoff;
            uint32_t       a = 0;
            for (uint32_t b = 0; b < e->len; ++b)
                if (k[b] < 128) {
                    uint64_t h = oracle_pattern_hash(k + a, b + 1 - a);
                    if (!store_find(&seen->st, k + a, b + 1 - a, h))
                        store_insert(&seen->st, k + a, b + 1 - a, h, 1, 0);
                    a = b + 1;
                }
        }
        m->totaltypes = seen->st.size;
        oracle_model_free(seen);
    }
    free(toks);
    free(corpus);
    *out = m;
    return 0;
}

void oracle_model_free(oracle_model* m) {
    if (!m)
        return;
    store_free(&m->st);
    free(m);
}
uint64_t oracle_model_size(const oracle_model* m) {
    return m->st.size;
}
uint64_t oracle_model_tokens(const oracle_model* m) {
    return m->totaltokens;
}
uint64_t oracle_model_types(const oracle_model* m) {
    return m->totaltypes;
}
int oracle_model_maxn(const oracle_model* m) {
    return m->maxn;
}
int oracle_model_minn(const oracle_model* m) {
    return m->minn;
}
int oracle_model_hasskipgrams(const oracle_model* m) {
    return m->hasskipgrams;
}
int oracle_model_passes(const oracle_model* m) {
    return m->npasses;
}
int oracle_model_pass_stats(const oracle_model* m, int pass, uint64_t out[4]) {
    if (pass < 0 || pass >= m->npasses)
        return 1;
    memcpy(out, m->pass[pass], 4 * sizeof(uint64_t));
    return 0;
}
uint32_t oracle_model_count(const oracle_model* m, const uint8_t* key, uint32_t len) {
    entry* e = store_find(&m->st, key, len, oracle_pattern_hash(key, len));
    return e ? e->count : 0;
}

/* canonical order: bytewise, shorter key first when one is a prefix of the other */
static const store* g_sort_store;
static int          cmp_entry(const void* a, const void* b) {
    const entry* x = *(const entry* const*)a;
    const entry* y = *(const entry* const*)b;
    uint32_t     l = x->len < y->len ? x->len : y->len;
    int          c = memcmp(g_sort_store->arena + x->keyoff, g_sort_store->arena + y->keyoff, l);
    if (c)
        return c;
    return x->len < y->len ? -1 : (x->len > y->len);
}
void oracle_model_export_sizes(const oracle_model* m, uint64_t* npatterns, uint64_t* keybytes, uint64_t* nrefs) {
    uint64_t kb = 0, nr = 0;
    for (uint64_t i = 0; i < m->st.cap; ++i)
        if (m->st.tab[i].used) {
            kb += m->st.tab[i].len;
            nr += m->st.tab[i].count;
        }
    *npatterns = m->st.size;
    *keybytes  = kb;
    *nrefs     = m->indexed ? nr : 0;
}
void oracle_model_export(const oracle_model* m, uint8_t* keys, uint64_t* key_off, uint32_t* counts, uint32_t* ref_sentence, uint16_t* ref_token, uint64_t* ref_off) {
    const entry** v = (const entry**)malloc((m->st.size + 1) * sizeof(entry*));
    uint64_t      k = 0;
    for (uint64_t i = 0; i < m->st.cap; ++i)
        if (m->st.tab[i].used)
            v[k++] = &m->st.tab[i];
    g_sort_store = &m->st;
    qsort(v, k, sizeof(entry*), cmp_entry);
    uint64_t ko = 0, ro = 0;
    for (uint64_t i = 0; i < k; ++i) {
        key_off[i] = ko;
        memcpy(keys + ko, m->st.arena + v[i]->keyoff, v[i]->len);
        ko += v[i]->len;
        counts[i] = v[i]->count;
        if (m->indexed && ref_off) {
            ref_off[i] = ro;
            for (uint32_t j = 0; j < v[i]->count; ++j) {
                ref_sentence[ro] = (uint32_t)(v[i]->refs[j] >> 16);
                ref_token[ro]    = (uint16_t)(v[i]->refs[j] & 0xFFFF);
                ++ro;
            }
        }
    }
    key_off[k] = ko;
    if (m->indexed && ref_off)
        ref_off[k] = ro;
    free(v);
}

/* ------------------------------------------------------------------------- */
/* model file: see oracle.h for the layout and its citations */
size_t oracle_model_write(const oracle_model* m, uint8_t* buf, size_t cap) {
    uint64_t np, kb, nr;
    oracle_model_export_sizes(m, &np, &kb, &nr);
    size_t need = 3 + 24 + kb + np * 5 + nr * 6;
    if (!buf || cap < need)
        return need;
    uint8_t*  keys = (uint8_t*)malloc(kb + 1);
    uint64_t* ko   = (uint64_t*)malloc((np + 1) * sizeof(uint64_t));
    uint32_t* cnt  = (uint32_t*)malloc((np + 1) * sizeof(uint32_t));
    uint32_t* rs   = (uint32_t*)malloc((nr + 1) * sizeof(uint32_t));
    uint16_t* rt   = (uint16_t*)malloc((nr + 1) * sizeof(uint16_t));
    uint64_t* ro   = (uint64_t*)malloc((np + 1) * sizeof(uint64_t));
    oracle_model_export(m, keys, ko, cnt, rs, rt, ro);
    size_t w = 0;
    buf[w++] = 0;
    buf[w++] = m->indexed ? 20 : 10; /* include/patternmodel.h:68-75 */
    buf[w++] = 2;
    memcpy(buf + w, &m->totaltokens, 8);
    w += 8;
    memcpy(buf + w, &m->totaltypes, 8);
    w += 8;
    memcpy(buf + w, &np, 8);
    w += 8;
    for (uint64_t i = 0; i < np; ++i) {
        size_t l = (size_t)(ko[i + 1] - ko[i]);
        memcpy(buf + w, keys + ko[i], l);
        w += l;
        buf[w++] = 0;
        memcpy(buf + w, &cnt[i], 4);
        w += 4;
        if (m->indexed) {
            for (uint64_t j = ro[i]; j < ro[i + 1]; ++j) {
                memcpy(buf + w, &rs[j], 4);
                w += 4;
                memcpy(buf + w, &rt[j], 2);
                w += 2;
            }
        }
    }
    free(keys);
    free(ko);
    free(cnt);
    free(rs);
    free(rt);
    free(ro);
    return w;
}

static int modelfile_walk(const uint8_t* d, size_t n, uint64_t hdr[7], uint8_t* keys, uint64_t* key_off, uint32_t* counts, uint32_t* ref_sentence, uint16_t* ref_token,
                          uint64_t* ref_off) {
    if (n < 27 || d[0] != 0)
        return fail("model file: bad header");
    uint64_t type = d[1], version = d[2], tokens, types, np;
    if (type != 10 && type != 20)
        return fail("model file: only types 10 (unindexed) and 20 (indexed) are parsed");
    memcpy(&tokens, d + 3, 8);
    memcpy(&types, d + 11, 8);
    memcpy(&np, d + 19, 8);
    size_t   pos = 27;
    uint64_t kb = 0, nr = 0;
    for (uint64_t i = 0; i < np; ++i) {
        size_t start    = pos;
        int    prevhigh = 0;
        for (;; ++pos) { /* key runs to the first 0x00 that does not follow a continuation byte */
            if (pos >= n)
                return fail("model file: truncated key");
            if (!prevhigh && d[pos] == 0)
                break;
            prevhigh = d[pos] >= 128;
        }
        size_t l = pos - start;
        ++pos;
        if (pos + 4 > n)
            return fail("model file: truncated value");
        uint32_t c;
        memcpy(&c, d + pos, 4);
        pos += 4;
        if (keys) {
            key_off[i] = kb;
            memcpy(keys + kb, d + start, l);
            counts[i] = c;
        }
        kb += l;
        if (type == 20) {
            if (pos + (size_t)c * 6 > n)
                return fail("model file: truncated index");
            if (keys && ref_off) {
                ref_off[i] = nr;
                for (uint32_t j = 0; j < c; ++j) {
                    memcpy(&ref_sentence[nr + j], d + pos + (size_t)j * 6, 4);
                    memcpy(&ref_token[nr + j], d + pos + (size_t)j * 6 + 4, 2);
                }
            }
            pos += (size_t)c * 6;
            nr += c;
        }
    }
    if (keys) {
        key_off[np] = kb;
        if (type == 20 && ref_off)
            ref_off[np] = nr;
    }
    if (hdr) {
        hdr[0] = type;
        hdr[1] = version;
        hdr[2] = tokens;
        hdr[3] = types;
        hdr[4] = np;
        hdr[5] = kb;
        hdr[6] = nr;
    }
    return 0;
}
int oracle_modelfile_scan(const uint8_t* data, size_t n, uint64_t hdr[7]) {
    return modelfile_walk(data, n, hdr, NULL, NULL, NULL, NULL, NULL, NULL);
}
int oracle_modelfile_parse(const uint8_t* data, size_t n, uint8_t* keys, uint64_t* key_off, uint32_t* counts, uint32_t* ref_sentence, uint16_t* ref_token,
                           uint64_t* ref_off) {
    return modelfile_walk(data, n, NULL, keys, key_off, counts, ref_sentence, ref_token, ref_off);
}

/* ------------------------------------------------------------------------- */
/* Synthetic corpus: counter-based, integer only, so that the
 * CUDA generator in the product library's test helpers and this one emit the
 * same bytes.  Not part of the reference. */
static inline uint64_t mix64(uint64_t z) { /* splitmix64 finaliser */
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint64_t rnd(uint64_t seed, uint64_t stream, uint64_t i) {
    return mix64((seed + stream * 0xD1B54A32D192ED03ULL) ^ mix64(i));
}
void oracle_synth_cdf(uint32_t vocab, uint64_t* cdf) {
    uint64_t acc = 0;
    for (uint32_t r = 0; r < vocab; ++r) {
        acc += (1ULL << 40) / (uint64_t)(r + 1);
        cdf[r] = acc;
    }
}
static uint32_t zipf_rank(const uint64_t* cdf, uint32_t vocab, uint64_t u) {
    uint64_t x  = u % cdf[vocab - 1];
    uint32_t lo = 0, hi = vocab - 1; /* first r with cdf[r] > x */
    while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (cdf[mid] > x)
            hi = mid;
        else
            lo = mid + 1;
    }
    return lo;
}
/* phrase layout: token index space is cut into blocks of 8; a block flagged as a phrase block starts with a
 * fixed phrase of 3..6 tokens (its id and content depend only on the seed), the rest of the block is ordinary text */
static int phrase_info(const oracle_synth_params* p, uint64_t i, uint64_t* phrase_id, uint32_t* j, uint32_t* plen) {
    if (!p->phrase_permille || !p->nphrases)
        return 0;
    uint64_t b = i >> 3;
    if (rnd(p->seed, 2, b) % 1000 >= p->phrase_permille)
        return 0;
    uint64_t id = rnd(p->seed, 4, b) % p->nphrases;
    uint32_t L  = 3 + (uint32_t)(rnd(p->seed, 3, id) % 4);
    uint32_t k  = (uint32_t)(i & 7);
    if (k >= L)
        return 0;
    *phrase_id = id;
    *j         = k;
    *plen      = L;
    return 1;
}
uint64_t oracle_synth_token(const oracle_synth_params* p, const uint64_t* cdf, uint64_t i) {
    uint64_t id;
    uint32_t j, L;
    if (phrase_info(p, i, &id, &j, &L))
        return 6 + zipf_rank(cdf, p->vocab, rnd(p->seed, 5, id * 8 + j));
    return 6 + zipf_rank(cdf, p->vocab, rnd(p->seed, 0, i));
}
static int synth_break_after(const oracle_synth_params* p, uint64_t i) {
    uint64_t id;
    uint32_t j, L;
    if (phrase_info(p, i, &id, &j, &L) && j + 1 < L)
        return 0; /* never break inside a phrase */
    return rnd(p->seed, 1, i) % p->mean_sentence == 0;
}
size_t oracle_synth_corpus(const oracle_synth_params* p, uint8_t* out, size_t cap) {
    uint64_t* cdf = (uint64_t*)malloc((size_t)p->vocab * sizeof(uint64_t));
    oracle_synth_cdf(p->vocab, cdf);
    size_t  n = 0;
    uint8_t buf[8];
    for (uint64_t k = 0; k < p->ntokens; ++k) {
        uint64_t i = p->first_token + k; /* index in the global stream */
        unsigned l = oracle_inttobytes(buf, (uint32_t)oracle_synth_token(p, cdf, i));
        if (out && n + l <= cap)
            memcpy(out + n, buf, l);
        n += l;
        if (synth_break_after(p, i) || k + 1 == p->ntokens) {
            if (out && n < cap)
                out[n] = 0;
            ++n;
        }
    }
    free(cdf);
    return n;
}
