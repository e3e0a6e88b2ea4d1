// engine_common.h -- host-side plumbing shared by engine.cu (single-GPU driver + C ABI) and shard.cu (multi-GPU phases):
// error reporting, the device memory pool, phase timers and the handle structs behind the opaque C types.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/colibri_b200.h"
#include "kernels.h"

namespace colibri {

// ------------------------------------------------------------------------------------------------ errors
extern thread_local char g_err[1024];
// An error return unwinds through scopes that own device blocks (DevBuf): work that was enqueued before the error may still be writing to them.
// set_err raises this flag; the first DevBuf released afterwards drains the device before its block goes back to the pool, where another
// stream or thread could be handed it.  (The flag costs one thread-local load on the normal path.)
extern thread_local bool g_unwinding;
inline int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    g_unwinding = true;
    return code;
}
#define CUDA_TRY(expr)                                                                                                        \
    do {                                                                                                                      \
        cudaError_t e__ = (expr);                                                                                             \
        if (e__ != cudaSuccess) return set_err(COLIBRI_E_CUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, cudaGetErrorString(e__)); \
    } while (0)
#define TRY(expr)                 \
    do {                          \
        int rc__ = (expr);        \
        if (rc__ != 0) return rc__; \
    } while (0)

// ------------------------------------------------------------------------------------------------ device memory pool
// cudaMalloc/cudaFree of multi-gigabyte buffers costs milliseconds and synchronises the device; training the same
// corpus repeatedly (the benchmark loop, or a CLI run with several models) reuses the blocks instead.
struct Pool {
    std::mutex                        mu;
    std::multimap<size_t, void*>      free_blocks;  // size -> ptr
    std::map<void*, size_t>           live;
    size_t                            in_use = 0, peak = 0, cached = 0, misses = 0;
    // Requests are rounded up to size classes (powers of two below 1 MB, quarter-octave steps above: <= 25 % slack) and a
    // free block is reused only by a request of its own class.  The same sequence of requests (every training step makes
    // one) then always finds its blocks; with "any block up to 25 % larger" a small request could take the block a later,
    // larger one needed, and the resulting cudaMalloc is millisecond-expensive once peer access is enabled (multi-GPU).
    static size_t size_class(size_t bytes) {
        if (bytes <= 256) return 256;
        if (bytes <= (1u << 20)) {
            size_t c = 256;
            while (c < bytes) c <<= 1;
            return c;
        }
        int k = 63 - __builtin_clzll((unsigned long long)bytes);
        size_t step = (size_t)1 << (k - 2);
        return (bytes + step - 1) / step * step;
    }
    int alloc(void** out, size_t bytes) {
        bytes = size_class(bytes);
        std::lock_guard<std::mutex> g(mu);
        auto it = free_blocks.find(bytes);
        if (it != free_blocks.end()) {
            *out = it->second;
            live[*out] = it->first;
            in_use += it->first;
            cached -= it->first;
            free_blocks.erase(it);
        } else {
            cudaError_t e = cudaMalloc(out, bytes);
            if (e != cudaSuccess) {  // drop the cache and retry once
                cudaGetLastError();
                for (auto& kv : free_blocks) cudaFree(kv.second);
                free_blocks.clear();
                cached = 0;
                e = cudaMalloc(out, bytes);
            }
            if (e != cudaSuccess) {
                cudaGetLastError();
                return set_err(COLIBRI_E_CUDA, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
            }
            live[*out] = bytes;
            in_use += bytes;
            ++misses;
        }
        peak = std::max(peak, in_use);
        return 0;
    }
    void release(void* p) {
        if (!p) return;
        std::lock_guard<std::mutex> g(mu);
        auto it = live.find(p);
        if (it == live.end()) return;
        in_use -= it->second;
        cached += it->second;
        free_blocks.emplace(it->second, p);
        live.erase(it);
    }
};
extern Pool g_pool[16];

template <class T>
struct DevBuf {
    T*     p   = nullptr;
    size_t n   = 0;
    int    dev = 0;
    DevBuf() {}
    DevBuf(const DevBuf&)            = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), dev(o.dev) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            reset();
            p = o.p; n = o.n; dev = o.dev;
            o.p = nullptr; o.n = 0;
        }
        return *this;
    }
    ~DevBuf() { reset(); }
    int alloc(int device, size_t count) {
        reset();
        dev = device;
        void* q = nullptr;
        const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
        TRY(g_pool[device & 15].alloc(&q, bytes));
        p = (T*)q;
        n = Pool::size_class(bytes) / sizeof(T);  // the whole block is usable: grow-only buffers re-allocate less often
        return 0;
    }
    void reset() {
        if (p && g_unwinding) {
            int cur = 0;
            cudaGetDevice(&cur);
            cudaSetDevice(dev);
            cudaDeviceSynchronize();
            cudaGetLastError();
            cudaSetDevice(cur);
            g_unwinding = false;
        }
        if (p) g_pool[dev & 15].release(p);
        p = nullptr;
        n = 0;
    }
};

// Timing events are reused across calls: cudaEventCreate / Destroy cost microseconds each and a train call takes dozens of spans.
struct EventCache {
    std::mutex               mu;
    std::vector<cudaEvent_t> free_events[16];
    cudaEvent_t get(int dev) {
        {
            std::lock_guard<std::mutex> g(mu);
            auto& v = free_events[dev & 15];
            if (!v.empty()) {
                cudaEvent_t e = v.back();
                v.pop_back();
                return e;
            }
        }
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        return e;
    }
    void put(int dev, cudaEvent_t e) {
        if (!e) return;
        std::lock_guard<std::mutex> g(mu);
        free_events[dev & 15].push_back(e);
    }
};
extern EventCache g_events;

// Host-side trace of one ABI call (COLIBRI_B200_TRACE=1): wall-clock marks, printed to stderr when the call returns.  The aux "tracing"
// subsystem of this path: it answers where the time between the device phases goes (allocation, synchronisation, copies).
struct HostTrace {
    bool                                        on = false;
    std::vector<std::pair<const char*, double>> marks;
    static double now_ms() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
    }
    void begin() {
        on = getenv("COLIBRI_B200_TRACE") != nullptr;
        marks.clear();
        mark("begin");
    }
    void mark(const char* what) {
        if (on) marks.emplace_back(what, now_ms());
    }
    void dump(const char* call) {
        if (!on || marks.empty()) return;
        fprintf(stderr, "[colibri_b200 trace] %s:", call);
        for (size_t i = 1; i < marks.size(); ++i) fprintf(stderr, " %s +%.3f", marks[i].first, marks[i].second - marks[i - 1].second);
        fprintf(stderr, " | total %.3f ms\n", marks.back().second - marks.front().second);
        marks.clear();
    }
};
extern thread_local HostTrace g_trace;

struct PhaseTimer {  // CUDA events on the library's stream, resolved after the final synchronise
    struct Span { int phase; cudaEvent_t a, b; int level; };
    std::vector<Span> spans;
    cudaStream_t      s = nullptr;
    int               dev = 0;
    int begin(int phase, int level = 0) {
        Span sp{phase, g_events.get(dev), g_events.get(dev), level};
        if (!sp.a || !sp.b) return -1;
        cudaEventRecord(sp.a, s);
        spans.push_back(sp);
        return (int)spans.size() - 1;
    }
    void end(int h) { if (h >= 0) cudaEventRecord(spans[h].b, s); }
    ~PhaseTimer() {
        for (auto& sp : spans) {
            g_events.put(dev, sp.a);
            g_events.put(dev, sp.b);
        }
    }
    void resolve(double ms[COLIBRI_T_NPHASES], std::map<int, double>* level_ms) {
        for (auto& sp : spans) {
            float t = 0;
            if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) {
                ms[sp.phase] += t;
                if (level_ms && sp.phase == COLIBRI_T_COUNT) (*level_ms)[sp.level] += t;
            }
            g_events.put(dev, sp.a);
            g_events.put(dev, sp.b);
        }
        spans.clear();
    }
};

struct Segment {  // survivors of one (level, category)
    int              n = 0;
    bool             skip = false;
    uint64_t         count = 0;
    DevBuf<uint32_t> pos, cnt, mask;
    // indexed models: the occurrences of the segment's patterns, grouped by pattern (segment order), ascending inside a pattern
    uint64_t         nrefs = 0;
    DevBuf<uint32_t> ref_sentence;
    DevBuf<uint16_t> ref_token;
    DevBuf<uint32_t> occ_pos;  // indexed + DOSKIPGRAMS: the nrefs occurrence positions of this level in corpus order
};

// ------------------------------------------------------------------------------------------------ tuning knobs
// When the large-corpus machinery switches on.  Read from the environment at every train call (not cached), so that the parity
// suite can force every path -- occurrence filter with a crowded bucket array, hot-key cache, dense pair slots, list mode --
// on corpora small enough for the oracle, and so that A/B measurements need no rebuild.
struct Tuning {
    uint64_t filter_min      = 1ull << 20;  // COLIBRI_B200_FILTER_MIN: smallest level (upper bound of its windows) that gets the occurrence filter
    int      filter_log2_min = 20;          // COLIBRI_B200_FILTER_LOG2_MIN / _LOG2: the filter has 2^min .. 2^max buckets (>= 2 per window where that fits)
    int      filter_log2_max = 30;          // (2^28 until round 2: at 1 B tokens the 64 MB filter saturated and level 3 took a 15 GB table; 2^30 buckets = 256 MB, table 3 GB)
    bool     no_filter       = false;       // COLIBRI_B200_NO_FILTER
    bool     filter_1bit     = true;        // COLIBRI_B200_FILTER_1BIT=0: the count launch reads the 2-bit counters instead of their packed "hit twice" bits
                                            // (half the footprint in L2: level 3 of the 100 M-token corpus 2.49 -> 2.30 ms)
    int      hot_mode        = 1;           // COLIBRI_B200_HOT: per-block hot-key cache 0 never, 1 levels >= hot_min, 2 always
    uint64_t hot_min         = 1ull << 25;  // COLIBRI_B200_HOT_MIN
    uint32_t dense_dim       = 3072;        // COLIBRI_B200_DENSE: side of the directly addressed square of level 2 (0 = off).  Measured with the partitioned level 2,
                                            // 100 M tokens: 1024 -> 9.81 ms per step, 2048 -> 9.48, 3072 -> 9.25, 4096 -> 9.45 (the 64 MB square no longer lives in L2)
    uint64_t dense_min       = 1ull << 25;  // COLIBRI_B200_DENSE_MIN
    uint64_t part_min        = 1ull << 26;  // COLIBRI_B200_PART_MIN: smallest level (upper bound of its windows) counted on the partitioned path (partition.cu)
    bool     part_all        = false;       // COLIBRI_B200_PART_ALL: every level >= part_min, not only level 2 with its dense square (measured: the later
                                            // levels are faster on the HBM table -- their frequent keys make partitions only one warp works through)
    uint32_t sparse_div      = 4;           // COLIBRI_B200_SPARSE_DIV: level n+1 runs from a position list when occurrences(n) * div <= positions (0 = never)
    static uint64_t env_u64(const char* name, uint64_t dflt) {
        const char* e = getenv(name);
        return e && *e ? strtoull(e, nullptr, 10) : dflt;
    }
    static Tuning from_env() {
        Tuning t;
        t.filter_min      = env_u64("COLIBRI_B200_FILTER_MIN", t.filter_min);
        t.filter_log2_min = (int)env_u64("COLIBRI_B200_FILTER_LOG2_MIN", t.filter_log2_min);
        t.filter_log2_max = (int)env_u64("COLIBRI_B200_FILTER_LOG2", t.filter_log2_max);
        t.filter_log2_min = std::max(6, std::min(t.filter_log2_min, 32));
        t.filter_log2_max = std::max(t.filter_log2_min, std::min(t.filter_log2_max, 32));
        t.no_filter       = getenv("COLIBRI_B200_NO_FILTER") != nullptr;
        t.filter_1bit     = env_u64("COLIBRI_B200_FILTER_1BIT", t.filter_1bit ? 1 : 0) != 0;
        t.hot_mode        = (int)env_u64("COLIBRI_B200_HOT", t.hot_mode);
        t.hot_min         = env_u64("COLIBRI_B200_HOT_MIN", t.hot_min);
        t.dense_dim       = (uint32_t)std::min<uint64_t>(env_u64("COLIBRI_B200_DENSE", t.dense_dim), 16384);
        t.dense_min       = env_u64("COLIBRI_B200_DENSE_MIN", t.dense_min);
        t.sparse_div      = (uint32_t)env_u64("COLIBRI_B200_SPARSE_DIV", t.sparse_div);
        t.part_min        = env_u64("COLIBRI_B200_PART_MIN", t.part_min);
        t.part_all        = env_u64("COLIBRI_B200_PART_ALL", 0) != 0;
        return t;
    }
    bool     use_filter(uint32_t mintokens, uint64_t bound) const { return mintokens >= 2 && !no_filter && bound >= filter_min; }
    bool     use_partition(uint64_t windows, bool dense_level) const { return windows >= part_min && (dense_level || part_all); }
    bool     use_hot(uint64_t bound) const { return hot_mode == 2 || (hot_mode == 1 && bound >= hot_min); }
    uint64_t filter_buckets(uint64_t bound) const {
        uint64_t nb = 1ull << filter_log2_min;
        while (nb < 2 * bound && nb < (1ull << filter_log2_max)) nb <<= 1;
        return nb;
    }
};

int check_options(colibri_b200_options& o);
// compute_skip_configurations (reference src/algorithms.cpp:79-94) with the run decomposition the kernels need
int skip_masks(int n, int maxskips, std::vector<SkipMask>& out);

}  // namespace colibri

using colibri::DevBuf;
using colibri::kTokTile;

// ------------------------------------------------------------------------------------------------ corpus
constexpr size_t kHalo = 16;  // zero bytes in front of the body: byte -1 must read as "< 128"
struct colibri_b200_corpus {
    int              device = 0;
    DevBuf<uint8_t>  buf;            // [kHalo zeros][body][2 spare][0x80 padding to a tile multiple + one tile]
    size_t           nbytes = 0;     // body bytes as given
    bool             ends_with_delim = true;
    uint8_t          last_byte = 0;
    cudaStream_t     stream = nullptr;
    double           h2d_ms = 0;
    cudaEvent_t      ev_h2d0 = nullptr, ev_h2d1 = nullptr;  // around the staging copy; ev_h2d1 doubles as "the body is in HBM" for other streams
    bool             h2d_pending = false;                   // staged asynchronously: h2d_ms is resolved after the caller's final synchronise
    size_t           chunk_bytes = 0;                       // > 0: the body was copied in chunks of this many bytes (a multiple of the tokeniser's tile),
    std::vector<cudaEvent_t> chunk_ev;                      //      chunk_ev[k] fires when chunk k is in HBM: the tokeniser follows the copy chunk by chunk
    uint8_t*         body() const { return buf.p + kHalo; }
    size_t           padded(size_t staged) const { return (staged + kTokTile - 1) / kTokTile * kTokTile; }
};


struct PassStat { uint64_t n, found, foundskip, pruned; };
struct LevelInfo { uint64_t windows = 0, cap = 0, singles = 0, items = 0; double ms = 0; int path = 0 /* 0 HBM table, 1 partitioned */; int filtered = 0; int fused_id1 = 0 /* the level's time includes the sweep that writes the level-1 ids */; };
struct colibri_b200_model {
    int      device = 0;
    int      model_type = COLIBRI_UNINDEXEDPATTERNMODEL;
    uint64_t npatterns = 0, keybytes = 0, nrefs = 0;
    uint64_t totaltokens = 0, totaltypes = 0;
    int      maxn = 0, minn = 999, hasskipgrams = 0, hasflexgrams = 0;
    std::vector<PassStat> passes;
    // device-resident flat export
    DevBuf<uint8_t>  d_keys;
    DevBuf<uint64_t> d_off;
    DevBuf<uint32_t> d_counts;
    DevBuf<uint16_t> d_len16;         // key length in bytes per pattern (compact export)
    DevBuf<uint32_t> d_ref_sentence;  // indexed models only
    DevBuf<uint16_t> d_ref_token;
    DevBuf<uint64_t> d_ref_off;
    // host copies (filled on first export / lookup)
    bool                  host_ready = false;
    std::vector<uint8_t>  h_keys;
    std::vector<uint64_t> h_off;
    std::vector<uint32_t> h_counts;
    std::vector<uint32_t> h_sorted;  // lookup index
    std::vector<uint32_t> h_ref_sentence;
    std::vector<uint16_t> h_ref_token;
    std::vector<uint64_t> h_ref_off;
    // shape of every pattern (tokens, category) and the hash index over the pattern bytes (pattern_index.cu); built on first
    // use by constrained training, load filters and lookups
    DevBuf<uint16_t>           d_pn;
    DevBuf<uint8_t>            d_pcat;
    bool                       meta_ready = false;
    colibri::PatternMetaStats  meta;
    DevBuf<colibri::PatSlot>   d_index;
    bool                       index_counts_dirty = false;  // a constrained run died between counting in the slots and collecting them
    uint64_t                   index_cap = 0;
    DevBuf<uint32_t>           d_presence;  // one bit per hash bucket (16 per pattern): negative lookups end in L2
    uint64_t                   presence_bits = 0;
    // constrained training only: class -> unigram pattern, and per length how many patterns lack their (n-1)-token prefix / suffix
    DevBuf<uint32_t>           d_uni;
    uint32_t                   uni_classes = 0;
    bool                       closure_ready = false;
    unsigned long long         prefix_open[256] = {0}, suffix_open[256] = {0};
    bool                       index_ready = false;
    double   ms[COLIBRI_T_NPHASES] = {0};
    uint64_t counters[8] = {0};
    std::map<int, LevelInfo> levels;
    cudaStream_t stream = nullptr;
};

// ------------------------------------------------------------------------------------------------ reverse index of a model over a corpus
// What IndexedPatternModel::getreverseindex (include/patternmodel.h:1746-1824) answers one position at a time, for the whole corpus at once:
// match[k][p] = (index + 1) of the model's n-gram of lengths[k] tokens that starts at position p, 0 if there is none.
struct colibri_b200_rindex {
    int                            device = 0;
    colibri_b200_model*            model = nullptr;  // not owned; must outlive the index
    cudaStream_t                   stream = nullptr;
    uint64_t                       npos = 0, nsentences = 0;
    int                            minn = 0, maxn = 0;
    std::vector<int>               lengths;
    std::vector<DevBuf<uint32_t>>  match;
    DevBuf<uint32_t>               tok;
    DevBuf<uint64_t>               sent_before;  // delimiters in tok[0..p)
    DevBuf<uint32_t>               sent_start;   // first position of sentence k (0-based); sent_start[nsentences] = one past the last delimiter
    DevBuf<const uint32_t*>        d_match_ptrs;
    DevBuf<uint32_t>               d_lengths;
};

namespace colibri {
// survivors of all levels -> the flat device-resident export of the model (engine.cu)
int export_segments(int dev, cudaStream_t s, std::vector<Segment>& segs, const uint32_t* tok, colibri_b200_model* m, uint64_t& launches);
// model_io.cu: a fresh handle with its own stream on `device`
int new_model(int device, int model_type, colibri_b200_model** out);
// model_io.cu: per-pattern shape (d_pn, d_pcat, meta) / the hash index over the pattern bytes, built once per model
int ensure_meta(colibri_b200_model* m, uint64_t* launches);
int ensure_index(colibri_b200_model* m, uint64_t* launches);
// model_io.cu: what constrained training wants to know about its constraint set (unigram table, closure per length), built once per model
int ensure_closure(colibri_b200_model* m, uint64_t* launches);
// model_io.cu: the patterns of `src` whose flag is set become the flat arrays of `dst` (work enqueued on dst->stream, synchronised on return).
//   d_counts: their counts (NULL: zeros); order_by_length: stable order by token count, so that occurrence lists built per length concatenate;
//   copy_refs: carry src's occurrence lists over; d_kmap (optional, src->npatterns entries): old index -> new index + 1, 0 = dropped.
int compact_patterns(const colibri_b200_model* src, const uint32_t* d_flags, const uint32_t* d_counts, bool order_by_length, bool copy_refs, uint32_t* d_kmap, colibri_b200_model* dst,
                     uint64_t* launches);
}  // namespace colibri
