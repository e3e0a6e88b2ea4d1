// engine.cu -- host-side driver of the device passes and the C ABI (include/colibri_b200.h).
//
// Restates the control flow of PatternModel::train (reference include/patternmodel.h:880-1345) around the kernels in
// kernels.cu: option normalisation (:883-888), one pass per n (:981), "None found" early stop (:1189-1194),
// totaltypes before the unigram prune (:1199-1201), prune after each pass (:1220), optional skipgram threshold
// (:1233-1243), MINLENGTH clean-up (:1221-1229, :1337-1341).  No CPU implementation of the counting exists here:
// without a CUDA device every entry point fails with COLIBRI_E_CUDA.
#include "engine_common.h"
#include "relations.h"
#include "spooky.h"

using namespace colibri;

thread_local char colibri::g_err[1024] = "";
thread_local bool colibri::g_unwinding = false;
colibri::Pool colibri::g_pool[16];
colibri::EventCache colibri::g_events;
thread_local colibri::HostTrace colibri::g_trace;

namespace {
// Per-call CUDA objects that are expensive to make are kept: streams (a create/destroy pair costs tens of microseconds and the destroy
// waits for the stream), and one small pinned block per device for the scalars the export stream reads back (cudaHostAlloc / cudaFreeHost
// cost up to a millisecond each and the free synchronises the whole device).
struct StreamCache {
    std::mutex                mu;
    std::vector<cudaStream_t> free_streams[16];
    int get(int dev, cudaStream_t* out) {
        {
            std::lock_guard<std::mutex> g(mu);
            auto& v = free_streams[dev & 15];
            if (!v.empty()) {
                *out = v.back();
                v.pop_back();
                return 0;
            }
        }
        CUDA_TRY(cudaStreamCreateWithFlags(out, cudaStreamNonBlocking));
        return 0;
    }
    void put(int dev, cudaStream_t s) {  // the stream must be idle
        if (!s) return;
        std::lock_guard<std::mutex> g(mu);
        free_streams[dev & 15].push_back(s);
    }
} g_streams;
struct PinnedScalars {
    std::mutex                       mu;
    std::vector<unsigned long long*> free_blocks;
    unsigned long long* get() {
        {
            std::lock_guard<std::mutex> g(mu);
            if (!free_blocks.empty()) {
                auto* p = free_blocks.back();
                free_blocks.pop_back();
                return p;
            }
        }
        unsigned long long* p = nullptr;
        if (cudaHostAlloc((void**)&p, 1024 * sizeof(unsigned long long), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return p;
    }
    void put(unsigned long long* p) {
        if (!p) return;
        std::lock_guard<std::mutex> g(mu);
        free_blocks.push_back(p);
    }
} g_pinned;
}  // namespace


extern "C" const char* colibri_b200_last_error(void) {
    return g_err;
}
extern "C" const char* colibri_b200_version(void) {
    return "colibri-core_b200 0.1 (sm_100a)";
}
extern "C" int colibri_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
extern "C" void colibri_b200_options_default(colibri_b200_options* o) {
    memset(o, 0, sizeof *o);
    o->MINTOKENS           = -1;  // reference include/patternmodel.h:153-180
    o->MINTOKENS_SKIPGRAMS = -1;
    o->MINTOKENS_UNIGRAMS  = 1;
    o->MINLENGTH           = 1;
    o->MAXLENGTH           = 100;
    o->MAXBACKOFFLENGTH    = 100;
    o->MINSKIPTYPES        = 2;
    o->MAXSKIPS            = 3;
    o->model_type          = COLIBRI_UNINDEXEDPATTERNMODEL;
    o->streamed            = 1;
    o->device              = 0;
}

static int corpus_alloc(colibri_b200_corpus* c, int device, size_t nbytes) {
    int ndev = colibri_b200_device_count();
    if (ndev <= 0) return set_err(COLIBRI_E_CUDA, "no CUDA device available: the B200 path has no CPU fallback");
    if (device < 0 || device >= ndev) return set_err(COLIBRI_E_INVALID, "device %d out of range (have %d)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    // the tables are probed one 32-byte sector at a time at random addresses: ask L2 not to fetch the neighbouring sector too (once per device)
    static bool limit_set[16] = {false};
    if (!limit_set[device & 15]) {
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, getenv("COLIBRI_B200_L2_FETCH") ? (size_t)atoi(getenv("COLIBRI_B200_L2_FETCH")) : 32);
        cudaGetLastError();
        limit_set[device & 15] = true;
    }
    c->device = device;
    c->nbytes = nbytes;
    size_t total = kHalo + c->padded(nbytes + 2) + kTokTile;
    TRY(c->buf.alloc(device, total));
    TRY(g_streams.get(device, &c->stream));
    CUDA_TRY(cudaMemsetAsync(c->buf.p, 0, kHalo, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->buf.p + kHalo + nbytes, 0x80, total - kHalo - nbytes, c->stream));
    return 0;
}
static int corpus_finish(colibri_b200_corpus* c) {
    // the last two body bytes decide whether the final sentence is terminated (delimiter = 0x00 not preceded by a continuation byte)
    uint8_t tail[2] = {0, 0};
    size_t  k       = std::min<size_t>(2, c->nbytes);
    if (k) CUDA_TRY(cudaMemcpyAsync(tail + (2 - k), c->body() + c->nbytes - k, k, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->last_byte       = tail[1];
    c->ends_with_delim = c->nbytes >= 1 && tail[1] == 0 && (c->nbytes == 1 || tail[0] < 128);
    return 0;
}

// Staging without a host synchronise: whether the last sentence is terminated is read from the caller's bytes, the copy is left running on
// the corpus stream and ev_h2d1 tells other streams when the body is in HBM.  host_body must stay valid until that event has passed
// (colibri_b200_train / _train_export synchronise before they return; the public colibri_b200_corpus_stage waits right away).
static int corpus_stage_async(const uint8_t* host_body, size_t nbytes, int device, colibri_b200_corpus** out) {
    *out = nullptr;
    if (!host_body && nbytes) return set_err(COLIBRI_E_INVALID, "host_body is NULL");
    auto* c = new colibri_b200_corpus();
    int   rc = corpus_alloc(c, device, nbytes);
    if (rc == 0) {
        c->ev_h2d0 = g_events.get(device);
        c->ev_h2d1 = g_events.get(device);
        if (!c->ev_h2d0 || !c->ev_h2d1) rc = set_err(COLIBRI_E_CUDA, "cudaEventCreate failed");
    }
    if (rc == 0) {
        cudaEventRecord(c->ev_h2d0, c->stream);
        // a large body goes in chunks with an event behind each, so that the tokeniser can start on chunk k while chunk k + 1 is on the bus
        size_t chunk = Tuning::env_u64("COLIBRI_B200_H2D_CHUNK", 16u << 20) / kTokTile * kTokTile;
        if (chunk == 0 || nbytes < 4 * chunk) chunk = 0;
        if (chunk) {
            c->chunk_bytes = chunk;
            for (size_t o0 = 0; o0 < nbytes && rc == 0; o0 += chunk) {
                cudaError_t e = cudaMemcpyAsync(c->body() + o0, host_body + o0, std::min(chunk, nbytes - o0), cudaMemcpyHostToDevice, c->stream);
                if (e != cudaSuccess) rc = set_err(COLIBRI_E_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
                cudaEvent_t ev = g_events.get(device);
                if (!ev) rc = set_err(COLIBRI_E_CUDA, "cudaEventCreate failed");
                else {
                    cudaEventRecord(ev, c->stream);
                    c->chunk_ev.push_back(ev);
                }
            }
        } else if (nbytes) {
            cudaError_t e = cudaMemcpyAsync(c->body(), host_body, nbytes, cudaMemcpyHostToDevice, c->stream);
            if (e != cudaSuccess) rc = set_err(COLIBRI_E_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
        }
        cudaEventRecord(c->ev_h2d1, c->stream);
        c->h2d_pending = true;
        // the last two body bytes decide whether the final sentence is terminated (delimiter = 0x00 not preceded by a continuation byte)
        c->last_byte       = nbytes ? host_body[nbytes - 1] : 0;
        c->ends_with_delim = nbytes >= 1 && host_body[nbytes - 1] == 0 && (nbytes == 1 || host_body[nbytes - 2] < 128);
    }
    if (rc) {
        colibri_b200_corpus_free(c);
        return rc;
    }
    *out = c;
    return 0;
}
static void corpus_resolve_h2d(colibri_b200_corpus* c) {  // after a synchronise that covers ev_h2d1
    if (!c->h2d_pending) return;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev_h2d0, c->ev_h2d1) == cudaSuccess) c->h2d_ms = ms;
    else cudaGetLastError();
    c->h2d_pending = false;
}
extern "C" int colibri_b200_corpus_stage(const uint8_t* host_body, size_t nbytes, int device, colibri_b200_corpus** out) {
    if (!out) return set_err(COLIBRI_E_INVALID, "out is NULL");
    TRY(corpus_stage_async(host_body, nbytes, device, out));
    cudaError_t e = cudaStreamSynchronize((*out)->stream);
    if (e != cudaSuccess) {
        colibri_b200_corpus_free(*out);
        *out = nullptr;
        return set_err(COLIBRI_E_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
    }
    corpus_resolve_h2d(*out);
    return 0;
}
extern "C" int colibri_b200_corpus_from_device(const void* dev_body, size_t nbytes, int device, colibri_b200_corpus** out) {
    if (!out) return set_err(COLIBRI_E_INVALID, "out is NULL");
    *out = nullptr;
    auto* c = new colibri_b200_corpus();
    int   rc = corpus_alloc(c, device, nbytes);
    if (rc == 0 && nbytes) {
        cudaError_t e = cudaMemcpyAsync(c->body(), dev_body, nbytes, cudaMemcpyDeviceToDevice, c->stream);
        if (e != cudaSuccess) rc = set_err(COLIBRI_E_CUDA, "D2D copy failed: %s", cudaGetErrorString(e));
    }
    if (rc == 0) rc = corpus_finish(c);
    if (rc) {
        colibri_b200_corpus_free(c);
        return rc;
    }
    *out = c;
    return 0;
}
extern "C" size_t colibri_b200_corpus_bytes(const colibri_b200_corpus* c) {
    return c ? c->nbytes : 0;
}
extern "C" void colibri_b200_corpus_free(colibri_b200_corpus* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) {
        cudaStreamSynchronize(c->stream);
        g_streams.put(c->device, c->stream);
    }
    g_events.put(c->device, c->ev_h2d0);
    g_events.put(c->device, c->ev_h2d1);
    for (auto ev : c->chunk_ev) g_events.put(c->device, ev);
    delete c;
}
extern "C" int colibri_b200_corpus_download(const colibri_b200_corpus* c, uint8_t* host, size_t cap) {
    if (!c || !host) return set_err(COLIBRI_E_INVALID, "NULL argument");
    if (cap < c->nbytes) return set_err(COLIBRI_E_INVALID, "buffer too small");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(host, c->body(), c->nbytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" void colibri_b200_model_free(colibri_b200_model* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) {
        cudaStreamSynchronize(m->stream);
        cudaStreamDestroy(m->stream);
    }
    delete m;
}
extern "C" uint64_t colibri_b200_model_size(const colibri_b200_model* m) { return m->npatterns; }
extern "C" uint64_t colibri_b200_model_tokens(const colibri_b200_model* m) { return m->totaltokens; }
extern "C" uint64_t colibri_b200_model_types(const colibri_b200_model* m) { return m->totaltypes; }
extern "C" int      colibri_b200_model_maxn(const colibri_b200_model* m) { return m->maxn; }
extern "C" int      colibri_b200_model_minn(const colibri_b200_model* m) { return m->minn; }
extern "C" int      colibri_b200_model_hasskipgrams(const colibri_b200_model* m) { return m->hasskipgrams; }
extern "C" int      colibri_b200_model_type(const colibri_b200_model* m) { return m->model_type; }
extern "C" int      colibri_b200_model_passes(const colibri_b200_model* m) { return (int)m->passes.size(); }
extern "C" int colibri_b200_model_pass_stats(const colibri_b200_model* m, int pass, uint64_t out[4]) {
    if (pass < 0 || pass >= (int)m->passes.size()) return set_err(COLIBRI_E_INVALID, "pass %d out of range", pass);
    out[0] = m->passes[pass].n;
    out[1] = m->passes[pass].found;
    out[2] = m->passes[pass].foundskip;
    out[3] = m->passes[pass].pruned;
    return 0;
}
extern "C" int colibri_b200_model_timings(const colibri_b200_model* m, double ms[COLIBRI_T_NPHASES]) {
    memcpy(ms, m->ms, sizeof m->ms);
    return 0;
}
extern "C" int colibri_b200_model_counters(const colibri_b200_model* m, uint64_t out[8]) {
    memcpy(out, m->counters, sizeof m->counters);
    return 0;
}
extern "C" int colibri_b200_model_level_counters(const colibri_b200_model* m, int n, double out[4]) {
    auto it = m->levels.find(n);
    if (it == m->levels.end()) return set_err(COLIBRI_E_INVALID, "level %d was not run", n);
    out[0] = (double)it->second.windows;
    out[1] = (double)it->second.cap;
    out[2] = it->second.ms;
    out[3] = (double)it->second.singles;
    return 0;
}
extern "C" int colibri_b200_model_level_info(const colibri_b200_model* m, int n, double out[8]) {
    auto it = m->levels.find(n);
    if (it == m->levels.end()) return set_err(COLIBRI_E_INVALID, "level %d was not run", n);
    memset(out, 0, 8 * sizeof(double));
    out[0] = (double)it->second.windows;
    out[1] = (double)it->second.cap;
    out[2] = it->second.ms;
    out[3] = (double)it->second.singles;
    out[4] = (double)it->second.items;
    out[5] = (double)it->second.path;
    out[6] = (double)it->second.filtered;
    out[7] = (double)it->second.fused_id1;
    return 0;
}

// ------------------------------------------------------------------------------------------------ training
namespace {

// compute_skip_configurations (reference src/algorithms.cpp:79-94): every non-empty subset of the inner positions
// 1..n-2, dropped when it has more than maxskips separate gaps (and n-2 >= maxskips)
}  // namespace
int colibri::skip_masks(int n, int maxskips, std::vector<SkipMask>& out) {
    out.clear();
    if (n < 3) return 0;
    if (n > 24) return set_err(COLIBRI_E_UNSUPPORTED, "skipgrams of %d tokens: the device key holds masks up to n=24", n);
    for (uint32_t i = 1; i < (1u << (n - 2)); ++i) {
        uint32_t mask = i << 1;
        int      runs = 0;
        for (int j = 0, in = 0; j < n; ++j) {
            int bit = (mask >> j) & 1;
            if (bit && !in) ++runs;
            in = bit;
        }
        if (n - 2 >= maxskips && runs > maxskips) continue;
        SkipMask sm;
        memset(&sm, 0, sizeof sm);
        sm.mask = mask;
        int parts = 0;
        for (int j = 0; j < n;) {
            if ((mask >> j) & 1) { ++j; continue; }
            int k = j;
            while (k < n && !((mask >> k) & 1)) ++k;
            if (parts < kMaxSkipParts) {
                sm.start[parts] = (uint8_t)j;
                sm.len[parts]   = (uint8_t)(k - j);
            }
            ++parts;
            j = k;
        }
        if (parts > kMaxSkipParts)
            return set_err(COLIBRI_E_UNSUPPORTED, "skipgram mask 0x%x of size %d has %d non-gap runs; the device path folds at most %d", mask, n, parts, kMaxSkipParts);
        sm.nparts = (uint32_t)parts;
        out.push_back(sm);
        if ((int)out.size() > kMaxSkipMasks) return set_err(COLIBRI_E_UNSUPPORTED, "more than %d skip configurations for n=%d (the reference enumerates 2^(n-2) too)", kMaxSkipMasks, n);
    }
    return 0;
}

int colibri::check_options(colibri_b200_options& o) {
    // include/patternmodel.h:883-888
    if (o.MINTOKENS == -1) o.MINTOKENS = 2;
    if (o.MINTOKENS == 0) o.MINTOKENS = 1;
    if (o.MINTOKENS_SKIPGRAMS < o.MINTOKENS) o.MINTOKENS_SKIPGRAMS = o.MINTOKENS;
    if (o.MINTOKENS < 1) return set_err(COLIBRI_E_INVALID, "MINTOKENS=%d", o.MINTOKENS);
    if (o.DOSKIPGRAMS && o.DOSKIPGRAMS_EXHAUSTIVE)
        return set_err(COLIBRI_E_INVALID, "Both DOSKIPGRAMS as well as DOSKIPGRAMS_EXHAUSTIVE are set, this shouldn't happen, choose one.");  // :958-963
    if (o.model_type != COLIBRI_UNINDEXEDPATTERNMODEL && o.model_type != COLIBRI_INDEXEDPATTERNMODEL)
        return set_err(COLIBRI_E_UNSUPPORTED, "model type %d (only 10 = unindexed and 20 = indexed run on the device)", o.model_type);
    if (o.model_type == COLIBRI_INDEXEDPATTERNMODEL && o.DOSKIPGRAMS_EXHAUSTIVE)
        return set_err(COLIBRI_E_UNSUPPORTED, "exhaustive skipgrams on an indexed model are not on the device path yet");
    if (o.DOSKIPGRAMS && o.model_type != COLIBRI_INDEXEDPATTERNMODEL)
        return set_err(COLIBRI_E_INVALID, "Can not compute skipgrams on unindexed model (except exhaustively during train() )");  // :1554-1561
    if (o.DOSKIPGRAMS && o.MINTOKENS == 1) return set_err(COLIBRI_E_UNSUPPORTED, "indexed skipgrams with MINTOKENS=1 are not on the device path");
    if (o.DOPATTERNPERLINE) return set_err(COLIBRI_E_UNSUPPORTED, "DOPATTERNPERLINE is not on the device path");
    if (o.PRUNENONSUBSUMED || o.PRUNESUBSUMED) return set_err(COLIBRI_E_UNSUPPORTED, "PRUNE(NON)SUBSUMED is not on the device path");
    if (o.MAXLENGTH < 1 || o.MAXLENGTH > 255) return set_err(COLIBRI_E_UNSUPPORTED, "MAXLENGTH=%d (device path supports 1..255)", o.MAXLENGTH);
    if (o.MINLENGTH < 1) o.MINLENGTH = 1;
    if (o.MINLENGTH > o.MAXLENGTH) return set_err(COLIBRI_E_INVALID, "MINLENGTH > MAXLENGTH");
    if (o.MINTOKENS > 1 && o.MAXBACKOFFLENGTH < o.MAXLENGTH - 1)
        return set_err(COLIBRI_E_UNSUPPORTED, "MAXBACKOFFLENGTH=%d < MAXLENGTH-1: truncated back-off is not on the device path", o.MAXBACKOFFLENGTH);
    if ((o.MINLENGTH > 1 || o.MINTOKENS == 1) && o.MINTOKENS_UNIGRAMS > o.MINTOKENS)
        return set_err(COLIBRI_E_UNSUPPORTED, "the separate unigram pre-pass (MINTOKENS_UNIGRAMS with MINLENGTH>1 or MINTOKENS=1, patternmodel.h:918-920) is not on the device path");
    if (o.MINTOKENS == 1 && o.MINLENGTH > 1) return set_err(COLIBRI_E_UNSUPPORTED, "MINTOKENS=1 with MINLENGTH>1 is not on the device path");
    return 0;
}

namespace {
// colibri_b200_train_export: the caller's host buffers and the second stream the finished levels leave on
struct ExportSink {
    uint8_t*     keys = nullptr;
    uint16_t*    len16 = nullptr;
    uint32_t*    counts = nullptr;
    uint64_t     keys_cap = 0, pat_cap = 0;
    uint64_t     npat = 0, nbytes = 0;  // what has been handed to the copy stream so far (also: what is needed, on overflow)
    bool         overflow = false;
    cudaStream_t xs = nullptr;
    struct Pending {
        uint64_t           count = 0;
        DevBuf<uint32_t>   nm, lens, cnt;
        DevBuf<uint64_t>   off, tmp;
        DevBuf<uint16_t>   len16;
        DevBuf<uint8_t>    keys;
        unsigned long long* h_kb = nullptr;  // pinned: total key bytes of the segment, written by an async copy on the main stream
        cudaEvent_t        ready = nullptr;
    };
    std::vector<Pending> pending, inflight;
    unsigned long long*  h_scalars = nullptr;  // pinned, one per segment
    int                  nscalars = 0;
    int                  dev = 0;
    ~ExportSink() {
        if (xs) {
            cudaStreamSynchronize(xs);
            g_streams.put(dev, xs);
        }
        for (auto& p : pending) g_events.put(dev, p.ready);
        for (auto& p : inflight) g_events.put(dev, p.ready);
        g_pinned.put(h_scalars);
    }
};

struct Trainer {
    colibri_b200_corpus*       c;
    colibri_b200_options       o;
    colibri_b200_model*        m;
    cudaStream_t               s;
    int                        dev, sms = 148;
    PhaseTimer                 timer;
    uint64_t                   launches = 0;
    DevBuf<DeviceStats>        d_stats;
    DeviceStats                h_stats;
    unsigned long long*        h_pinned = nullptr;  // mapped pinned block: the statistics arrive here by kernel stores, not through the copy engine
    ~Trainer() { g_pinned.put(h_pinned); }
    int fetch_stats() {  // d_stats -> h_stats, synchronising the stream
        if (!h_pinned) {
            h_pinned = g_pinned.get();
            if (!h_pinned) return set_err(COLIBRI_E_CUDA, "no pinned memory for the statistics block");
        }
        launches += launch_copy_words_to_host(s, d_stats.p, h_pinned, (uint32_t)sizeof(DeviceStats));
        CUDA_TRY(cudaStreamSynchronize(s));
        memcpy(&h_stats, h_pinned, sizeof(DeviceStats));
        return 0;
    }
    std::vector<Segment>       segs;
    uint64_t                   slots_init = 0, ngram_upserts = 0, skip_upserts = 0, filtered_windows = 0;
    Tuning                     tune = Tuning::from_env();
    ExportSink*                sink = nullptr;
    uint64_t                   tok_ext_cells = 0;  // the token array has room for this many (class, class) pairs behind position npos + 8
    size_t                     l2_persist_max = 0, l2_window_max = 0;
    // partitioned counting path (partition.cu): record buffers of the two splits and the small per-partition arrays, reused across levels
    DevBuf<unsigned long long> part_rk1, part_rk2;
    DevBuf<uint32_t>           part_rp1, part_rp2, part_small, part_dense_id, part_dense_bits;
    int  level_partitioned(int n, const uint32_t* prev, uint32_t* cur, uint64_t npos, const uint32_t* list, uint64_t nlist, uint64_t wbound, uint32_t dense, uint32_t* dense_cnt,
                           uint32_t t, uint32_t* tok_ext, Segment& sg, bool& overflow, double hashed_share, DevBuf<uint32_t>* slot_index, bool pre_hist);
    int  part_small_layout(const PartPlan& pl);
    // windows the partitions are sized for.  COLIBRI_B200_PART_SCALE (a test knob) scales it: a value far below 1 makes partitions that cannot fit their
    // shared-memory tables, which exercises the overflow -> HBM-table fallback
    static uint64_t plan_bound(uint64_t wbound, double hashed_share) {
        const char*  e     = getenv("COLIBRI_B200_PART_SCALE");
        const double scale = e && *e ? atof(e) : 1.0;
        return (uint64_t)((double)wbound * std::min(1.0, hashed_share) * scale) + 1024;
    }
    int  l2_pin(const void* base, size_t bytes);
    void l2_unpin();
    const uint32_t*            tok_for_sink = nullptr;
    uint32_t                   sink_maxclass = 0;
    int emit_segment(Segment& sg);
    int flush_sink(bool final);

    int zero_stats(bool keep_global = true) {
        // found/kept/kept_occ/cursor/valid_windows/probes are per-phase; totaltokens/maxclass/errflags live for the whole train
        (void)keep_global;
        CUDA_TRY(cudaMemsetAsync(&d_stats.p->found, 0, offsetof(DeviceStats, maxclass) - offsetof(DeviceStats, found), s));
        return 0;
    }
    int read_stats() {
        CUDA_TRY(cudaGetLastError());  // a kernel of this phase that failed to launch (bad configuration, missing opt-in) must not read as "nothing found"
        TRY(fetch_stats());
        if (h_stats.errflags & kErrTableFull) return set_err(COLIBRI_E_CAPACITY, "device hash table overflow");
        if (sink && !sink->pending.empty()) {  // every host sync of the level loop is a chance to hand finished segments to the copy stream
            g_trace.mark("(");
            TRY(flush_sink(false));
            g_trace.mark("flush)");
        }
        return 0;
    }
    // forward index (indexed models)
    bool               indexed = false;
    DevBuf<uint64_t>   sent_before;  // delimiters in tok[0..p)
    DevBuf<uint32_t>   sent_start;   // first position of sentence k (0-based)
    int prepare_index(const uint32_t* tok, uint64_t npos);
    int build_refs(Segment& sg, const uint32_t* ids, const uint32_t* map, bool by_class, uint64_t npos, uint64_t expect, bool keep_positions = false,
                   const uint32_t* pos_lookup = nullptr, uint32_t pos_div = 1);
    int indexed_skipgrams(int n, Segment& ng, const std::vector<DevBuf<uint32_t>>& ids, uint64_t& foundskip, uint64_t& keptskip, Segment& out);
    int tokenise(DevBuf<uint32_t>& tok, uint64_t& npos, uint32_t& nclasses);
    int run();
    // phase 0: the whole call; 1: count this corpus into ext_counts only (a shard of a multi-GPU run); 2: threshold + compaction of
    // ext_counts (summed over the shards by the caller), corpus_tokens = ext_tokens.  Phases 1 and 2 are for unindexed models.
    int run_constrained(colibri_b200_model* cm, bool inplace, int phase = 0, uint32_t* ext_counts = nullptr, uint64_t ext_tokens = 0);
    int build_rindex(colibri_b200_model* cm, colibri_b200_rindex* r);
};

int Trainer::prepare_index(const uint32_t* tok, uint64_t npos) {
    DevBuf<uint32_t> flags;
    DevBuf<uint64_t> tmp;
    TRY(flags.alloc(dev, npos));
    TRY(tmp.alloc(dev, npos / 2048 + 4));
    TRY(sent_before.alloc(dev, npos + 1));
    const uint64_t ndelims = npos - m->totaltokens;
    TRY(sent_start.alloc(dev, ndelims + 2));
    launches += launch_delim_flags(s, tok, npos, flags.p);
    launches += launch_exclusive_scan_u32_u64(s, flags.p, sent_before.p, npos, tmp.p);
    launches += launch_sent_start(s, tok, sent_before.p, npos, sent_start.p);
    CUDA_TRY(cudaStreamSynchronize(s));  // the temporaries go back to the pool
    return 0;
}

// occurrences of one level -> (sentence, token) lists grouped by survivor, ascending inside each (see index.cu)
int Trainer::build_refs(Segment& sg, const uint32_t* ids, const uint32_t* map, bool by_class, uint64_t npos, uint64_t expect, bool keep_positions, const uint32_t* pos_lookup,
                        uint32_t pos_div) {
    if (sg.count == 0) return 0;
    const uint64_t   nblk = (npos + 2047) / 2048;
    DevBuf<uint32_t> blk;
    DevBuf<uint64_t> blk_off, tmp;
    TRY(blk.alloc(dev, nblk + 1));
    TRY(blk_off.alloc(dev, nblk + 2));
    TRY(tmp.alloc(dev, nblk / 2048 + 4));
    launches += launch_pair_count(s, ids, map, npos, by_class, blk.p);
    launches += launch_exclusive_scan_u32_u64(s, blk.p, blk_off.p, nblk, tmp.p);
    uint64_t R = 0;
    CUDA_TRY(cudaMemcpyAsync(&R, blk_off.p + nblk, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (R != expect) return set_err(COLIBRI_E_CUDA, "forward index of level %d: %llu occurrences found, %llu expected", sg.n, (unsigned long long)R, (unsigned long long)expect);
    if (R >= 0xFFFFFFF0ull) return set_err(COLIBRI_E_CAPACITY, "forward index of level %d has %llu entries", sg.n, (unsigned long long)R);
    DevBuf<uint32_t> ka, va, kb, vb, hist;
    DevBuf<uint64_t> hist_off, stmp;
    const uint64_t   nsort = (R + 4095) / 4096;
    TRY(ka.alloc(dev, R));
    TRY(va.alloc(dev, R));
    TRY(kb.alloc(dev, R));
    TRY(vb.alloc(dev, R));
    TRY(hist.alloc(dev, 256 * nsort));
    TRY(hist_off.alloc(dev, 256 * nsort + 1));
    TRY(stmp.alloc(dev, 256 * nsort / 2048 + 4));
    launches += launch_pair_write(s, ids, map, npos, by_class, blk_off.p, ka.p, va.p, pos_lookup, pos_div);
    if (keep_positions) {  // corpus-order occurrence positions of this level: the windows of its skipgrams (trainskipgrams)
        TRY(sg.occ_pos.alloc(dev, R));
        CUDA_TRY(cudaMemcpyAsync(sg.occ_pos.p, va.p, R * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    }
    uint32_t *kin = ka.p, *vin = va.p, *kout = kb.p, *vout = vb.p;
    for (int shift = 0; shift < 32 && ((sg.count - 1) >> shift) != 0; shift += 8) {
        launches += launch_radix_pass(s, kin, vin, R, shift, hist.p, hist_off.p, stmp.p, kout, vout);
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
    TRY(sg.ref_sentence.alloc(dev, R));
    TRY(sg.ref_token.alloc(dev, R));
    launches += launch_refs_from_positions(s, vin, R, sent_before.p, sent_start.p, sg.ref_sentence.p, sg.ref_token.p, d_stats.p);
    sg.nrefs = R;
    TRY(read_stats());
    if (h_stats.errflags & kErrLongSentence)
        return set_err(COLIBRI_E_UNSUPPORTED, "indexed model: a sentence has more than 65536 tokens (IndexReference.token is 16 bit; the class encoder splits such lines)");
    return 0;
}

// IndexedPatternModel::trainskipgrams for one n (reference include/patternmodel.h:2969-3010): every occurrence of every surviving
// n-gram contributes to each of its gap configurations (computeskipgrams with multiplerefs, :1370-1527), then prune(MINTOKENS, n)
// and the skip-type rule of the indexed pruneskipgrams (:3362-3383).  The skipgrams' occurrence lists come out of the same
// ordered-pairs + stable radix sort as the n-grams'.
int Trainer::indexed_skipgrams(int n, Segment& ng, const std::vector<DevBuf<uint32_t>>& ids, uint64_t& foundskip, uint64_t& keptskip, Segment& out) {
    foundskip = keptskip = 0;
    std::vector<SkipMask> masks;
    TRY(skip_masks(n, o.MAXSKIPS, masks));
    const uint64_t nocc = ng.nrefs;
    if (masks.empty() || nocc == 0) return 0;
    const uint64_t items = nocc * masks.size();
    if (items >= 0xFFFFFFF0ull) return set_err(COLIBRI_E_CAPACITY, "level %d has %llu skipgram occurrences", n, (unsigned long long)items);
    uint64_t sbound = 0;
    for (auto& sm : masks) sbound += nocc * (1 + (sm.nparts > 3 ? (sm.nparts - 2) / 2 : 0));
    const uint64_t scap = std::max<uint64_t>(64, sbound + sbound / 2 + 16);
    DevBuf<SkipSlot>        table;
    DevBuf<const uint32_t*> d_ptrs;
    DevBuf<SkipMask>        d_m;
    DevBuf<uint32_t>        item_slot, types, slot_index;
    DevBuf<NgramSlot>       pairs;
    TRY(table.alloc(dev, scap));
    TRY(item_slot.alloc(dev, items));
    TRY(slot_index.alloc(dev, scap));
    TRY(d_ptrs.alloc(dev, n + 1));
    TRY(d_m.alloc(dev, masks.size()));
    std::vector<const uint32_t*> ptrs(n + 1, nullptr);
    for (int k = 1; k <= n; ++k) ptrs[k] = ids[k].p;
    CUDA_TRY(cudaMemcpyAsync(d_ptrs.p, ptrs.data(), (n + 1) * sizeof(uint32_t*), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_m.p, masks.data(), masks.size() * sizeof(SkipMask), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemsetAsync(table.p, 0, scap * sizeof(SkipSlot), s));
    CUDA_TRY(cudaMemsetAsync(item_slot.p, 0, items * sizeof(uint32_t), s));
    TRY(zero_stats());
    int hs = timer.begin(COLIBRI_T_SKIPGRAMS, n);
    launches += launch_count_skipgrams(s, d_ptrs.p, n, d_m.p, (int)masks.size(), nocc, table.p, scap, d_stats.p, sms, ng.occ_pos.p, item_slot.p);
    const bool by_types = o.MINSKIPTYPES > 1;
    if (by_types) {
        const uint64_t pcap = std::max<uint64_t>(64, items + items / 2 + 16);
        TRY(types.alloc(dev, scap));
        TRY(pairs.alloc(dev, pcap));
        CUDA_TRY(cudaMemsetAsync(types.p, 0, scap * sizeof(uint32_t), s));
        CUDA_TRY(cudaMemsetAsync(pairs.p, 0, pcap * sizeof(NgramSlot), s));
        launches += launch_skip_types(s, d_ptrs.p, n, d_m.p, (int)masks.size(), nocc, ng.occ_pos.p, item_slot.p, pairs.p, pcap, types.p, d_stats.p, sms);
    }
    timer.end(hs);
    TRY(read_stats());  // also keeps ptrs/masks alive until the copies are done
    skip_upserts += h_stats.valid_windows;
    out.n    = n;
    out.skip = true;
    const uint64_t bound = items / std::max<uint32_t>((uint32_t)o.MINTOKENS, 1) + 1;
    TRY(out.pos.alloc(dev, bound));
    TRY(out.cnt.alloc(dev, bound));
    TRY(out.mask.alloc(dev, bound));
    TRY(zero_stats());
    int hp = timer.begin(COLIBRI_T_PRUNE);
    launches += launch_prune_skipgrams(s, table.p, scap, (uint32_t)o.MINTOKENS, out.pos.p, out.cnt.p, out.mask.p, d_stats.p, sms, slot_index.p, by_types ? types.p : nullptr,
                                       (uint32_t)std::max(o.MINSKIPTYPES, 0));
    timer.end(hp);
    TRY(read_stats());
    foundskip = h_stats.found;
    keptskip  = h_stats.kept;
    out.count = keptskip;
    if (keptskip) {
        int hi = timer.begin(COLIBRI_T_INDEX);
        TRY(build_refs(out, item_slot.p, slot_index.p, false, items, h_stats.kept_occ, false, ng.occ_pos.p, (uint32_t)masks.size()));
        timer.end(hi);
    }
    return 0;
}

// ---- L2 residency for the small random-access structures of a level (occurrence filter + dense square, <= ~50 MB): without it the
// table's HBM traffic keeps evicting them and the count launch re-fetches filter lines from DRAM (profiles/r02_ncu.md: 4.4 GB of the
// 6.3 GB the level-2 count launch read).  A stream access-policy window marks the buffer persisting; the set-aside is released after the level.
int Trainer::l2_pin(const void* base, size_t bytes) {
    // MEASURED (B200, 100 M tokens, profiles/r02_ncu.md): with the window on, the whole step went 10.9 -> 15.2 ms (every kernel slower, the
    // tokeniser included: the set-aside shrinks the L2 everything else lives in).  Off unless asked for; kept as the evidence.
    if (!getenv("COLIBRI_B200_L2_PIN") || bytes == 0) return 0;
    if (l2_persist_max == 0) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, dev);
        l2_persist_max = (size_t)std::max(v, 0);
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        l2_window_max = (size_t)std::max(v, 0);
        if (l2_persist_max == 0 || l2_window_max == 0) {
            l2_persist_max = l2_window_max = 1;  // not supported: stay quiet
            return 0;
        }
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, l2_persist_max);
    }
    if (l2_persist_max <= 1) return 0;
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof av);
    av.accessPolicyWindow.base_ptr  = const_cast<void*>(base);
    av.accessPolicyWindow.num_bytes = std::min(bytes, l2_window_max);
    av.accessPolicyWindow.hitRatio  = (float)std::min(1.0, (double)l2_persist_max / (double)av.accessPolicyWindow.num_bytes);
    av.accessPolicyWindow.hitProp   = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp  = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
    return 0;
}
void Trainer::l2_unpin() {
    if (l2_persist_max <= 1) return;
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof av);
    av.accessPolicyWindow.num_bytes = 0;
    av.accessPolicyWindow.hitProp   = cudaAccessPropertyNormal;
    av.accessPolicyWindow.missProp  = cudaAccessPropertyNormal;
    if (cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
    if (cudaCtxResetPersistingL2Cache() != cudaSuccess) cudaGetLastError();
}

// ---- streamed export (colibri_b200_train_export).  A level's survivors are final once its prune scan ran, so their pattern bytes are
// produced right away (the same kernels export_segments runs at the end, on the main stream: small, and their inputs are hot) into a
// staging block sized by an upper bound; the total lands in pinned memory with the next statistics read, and from then on the host knows
// how many bytes to copy: keys, lengths and counts leave on the second stream while the next level counts.
int Trainer::emit_segment(Segment& sg) {
    if (sg.count == 0) return 0;
    if (sink->nscalars >= 1024) return set_err(COLIBRI_E_CAPACITY, "streamed export: more than 1024 segments");
    ExportSink::Pending p;
    p.count = sg.count;
    const uint32_t vmax = varint_len(sink_maxclass);
    uint64_t       bound = 0;  // bytes of the segment's keys at most
    if (sg.n == 1) bound = sg.count * vmax;
    else bound = sg.count * (uint64_t)sg.n * vmax;
    TRY(p.nm.alloc(dev, sg.count));
    TRY(p.lens.alloc(dev, sg.count));
    TRY(p.off.alloc(dev, sg.count + 1));
    TRY(p.tmp.alloc(dev, sg.count / 2048 + 4));
    TRY(p.len16.alloc(dev, sg.count));
    TRY(p.keys.alloc(dev, bound + 16));
    if (sg.skip) {
        CUDA_TRY(cudaMemcpyAsync(p.nm.p, sg.mask.p, sg.count * 4, cudaMemcpyDeviceToDevice, s));
        launches += launch_pack_nm(s, p.nm.p, sg.count, (uint32_t)sg.n);
    } else {
        launches += launch_fill_u32(s, p.nm.p, sg.count, (uint32_t)sg.n);
    }
    launches += launch_export_lengths(s, tok_for_sink, sg.pos.p, p.nm.p, sg.count, p.lens.p, p.len16.p);
    launches += launch_exclusive_scan_u32_u64(s, p.lens.p, p.off.p, sg.count, p.tmp.p);
    launches += launch_export_write(s, tok_for_sink, sg.pos.p, p.nm.p, p.off.p, sg.count, p.keys.p);
    p.h_kb = sink->h_scalars + sink->nscalars++;
    launches += launch_copy_words_to_host(s, p.off.p + sg.count, p.h_kb, (uint32_t)sizeof(unsigned long long));  // (not the copy engine: it is busy with the previous level)
    // the counts travel from the segment's own array; it stays alive (segs) until the end of the call
    p.cnt = std::move(sg.cnt);
    p.ready = g_events.get(dev);
    if (!p.ready) return set_err(COLIBRI_E_CUDA, "cudaEventCreate failed");
    CUDA_TRY(cudaEventRecord(p.ready, s));
    sink->pending.push_back(std::move(p));
    return 0;
}

// called right after a synchronise of the main stream: every pending segment's byte total is known now
int Trainer::flush_sink(bool final) {
    for (auto& p : sink->pending) {
        const uint64_t kb = *p.h_kb;
        if (sink->npat + p.count > sink->pat_cap || sink->nbytes + kb > sink->keys_cap) sink->overflow = true;
        if (!sink->overflow) {
            CUDA_TRY(cudaStreamWaitEvent(sink->xs, p.ready, 0));
            if (kb) CUDA_TRY(cudaMemcpyAsync(sink->keys + sink->nbytes, p.keys.p, kb, cudaMemcpyDeviceToHost, sink->xs));
            CUDA_TRY(cudaMemcpyAsync(sink->len16 + sink->npat, p.len16.p, p.count * sizeof(uint16_t), cudaMemcpyDeviceToHost, sink->xs));
            CUDA_TRY(cudaMemcpyAsync(sink->counts + sink->npat, p.cnt.p, p.count * sizeof(uint32_t), cudaMemcpyDeviceToHost, sink->xs));
        }
        sink->npat += p.count;
        sink->nbytes += kb;
        sink->inflight.push_back(std::move(p));  // its blocks go back to the pool when the sink dies, after the copy stream drained
    }
    sink->pending.clear();
    if (final) CUDA_TRY(cudaStreamSynchronize(sink->xs));
    return 0;
}

int Trainer::part_small_layout(const PartPlan& pl) {
    const uint64_t small_words = 5ull * (1u << pl.b1) + 3ull * pl.nparts + 16;
    if (part_small.n < small_words) TRY(part_small.alloc(dev, small_words));
    return 0;
}

// One level on the partitioned path (partition.cu).  On return h_stats holds the level's device statistics (valid_windows, found, kept,
// kept_occ, singletons = keys that occur once); sg.pos / sg.cnt hold the survivors, cur[] the final ids (survivor index + 1, 0 = pruned or no
// window).  overflow: a partition did not fit its shared-memory table -- nothing of the level is usable, the caller reruns it on the HBM table.
int Trainer::level_partitioned(int n, const uint32_t* prev, uint32_t* cur, uint64_t npos, const uint32_t* list, uint64_t nlist, uint64_t wbound, uint32_t dense,
                               uint32_t* dense_cnt, uint32_t t, uint32_t* tok_ext, Segment& sg, bool& overflow, double hashed_share, DevBuf<uint32_t>* slot_index, bool pre_hist) {
    overflow = false;
    // partitions are sized for the windows expected to become records (an estimate that is too low shows up as an overflow, not as a wrong count)
    const PartPlan pl  = part_plan(plan_bound(wbound, hashed_share));
    const uint32_t p1n = 1u << pl.b1;
    TRY(part_small_layout(pl));  // hist1 | off1 (+1) | cursor1 | group_tot | group_base (+1) | off (+1) | kept_of (first: hist2) | dst_off (+1)
    uint32_t* hist1      = part_small.p;
    uint32_t* off1       = hist1 + p1n;
    uint32_t* cursor1    = off1 + p1n + 1;
    uint32_t* group_tot  = cursor1 + p1n;
    uint32_t* group_base = group_tot + p1n;
    uint32_t* off        = group_base + p1n + 1;
    uint32_t* kept_of    = off + pl.nparts + 1;
    uint32_t* dst_off    = kept_of + pl.nparts;
    if (part_rk1.n < wbound + 1) TRY(part_rk1.alloc(dev, wbound + 1));
    if (part_rp1.n < wbound + 1) TRY(part_rp1.alloc(dev, wbound + 1));
    if (part_rk2.n < wbound + 1) TRY(part_rk2.alloc(dev, wbound + 1));
    if (part_rp2.n < wbound + 1) TRY(part_rp2.alloc(dev, wbound + 1));
    const uint64_t dense_cells = (uint64_t)dense * dense;
    if (dense_cells + wbound + 2 >= 0xFFFFFFF0ull) return set_err(COLIBRI_E_CAPACITY, "level %d: %llu windows; ids are 32 bit", n, (unsigned long long)wbound);
    if (dense) {
        if (part_dense_id.n < dense_cells) TRY(part_dense_id.alloc(dev, dense_cells));
        if (part_dense_bits.n < dense_cells / 32 + 8) TRY(part_dense_bits.alloc(dev, dense_cells / 32 + 8));
    }
    const uint64_t sv_bound = wbound / std::max<uint32_t>(t, 1) + 1;
    TRY(sg.pos.alloc(dev, sv_bound));
    TRY(sg.cnt.alloc(dev, sv_bound));
    if (slot_index) {  // indexed models: id - 1 -> survivor index + 1 (the dense survivors' ids are their indices already)
        if (slot_index->n < dense_cells + wbound + 8) TRY(slot_index->alloc(dev, dense_cells + wbound + 8));
        launches += launch_iota_plus1(s, slot_index->p, dense_cells);
    }
    const uint64_t nitems = list ? nlist : npos;
    // the survivors wait inside their partition's record range until part_gather compacts them: the first split's buffers are free by then
    uint32_t* tmp_pos = reinterpret_cast<uint32_t*>(part_rk1.p);
    uint32_t* tmp_cnt = part_rp1.p;

    int hc = timer.begin(COLIBRI_T_COUNT, n);
    uint32_t* hist2 = kept_of;  // the final partitions' sizes (pass B) are consumed by the scan before pass E writes the survivor counts there
    CUDA_TRY(cudaMemsetAsync(hist2, 0, (size_t)pl.nparts * sizeof(uint32_t), s));
    if (!pre_hist) {  // (pre_hist: pass A ran inside the level-1 id sweep, launch_make_id1_hist: hist1, the dense square and valid_windows are there)
        CUDA_TRY(cudaMemsetAsync(hist1, 0, (size_t)p1n * sizeof(uint32_t), s));
        if (dense) CUDA_TRY(cudaMemsetAsync(dense_cnt, 0, dense_cells * sizeof(uint32_t), s));
        TRY(zero_stats());
        launches += launch_part_hist(s, prev, list, nitems, dense, dense_cnt, hist1, pl, d_stats.p, sms);
    }
    if (dense)  // prune(MINTOKENS, 2) of the dense square first: its survivors open the segment, dense_id[cell] = survivor index + 1
        launches += launch_prune_dense(s, dense_cnt, dense, 0, t, sg.pos.p, sg.cnt.p, part_dense_bits.p, part_dense_id.p, tok_ext, (uint32_t)(npos + 8), d_stats.p, sms);
    launches += launch_part_bases(s, hist1, p1n, off1, cursor1, nullptr);
    launches += launch_part_split1(s, prev, list, nitems, dense, part_dense_id.p, cur, pl, cursor1, hist2, part_rk1.p, part_rp1.p);
    launches += launch_part_scan(s, hist2, pl, group_tot, group_base, off, nullptr);
    launches += launch_part_split2(s, part_rk1.p, part_rp1.p, off, pl, part_rk2.p, part_rp2.p);
    launches += launch_part_count(s, part_rk2.p, part_rp2.p, off, pl, t, cur, (uint32_t)dense_cells, 1, 1, tmp_pos, tmp_cnt, kept_of, d_stats.p, sms);
    launches += launch_part_gather(s, tmp_pos, tmp_cnt, off, kept_of, pl, group_tot, group_base, dst_off, &d_stats.p->cursor, sg.pos.p, sg.cnt.p,
                                   slot_index ? slot_index->p : nullptr, (uint32_t)dense_cells);
    timer.end(hc);
    CUDA_TRY(cudaGetLastError());
    TRY(fetch_stats());
    if (h_stats.errflags & kErrTableFull) {
        CUDA_TRY(cudaMemsetAsync(&d_stats.p->errflags, 0, sizeof(unsigned int), s));
        overflow = true;
        return 0;
    }
    if (sink) TRY(flush_sink(false));
    m->levels[n].cap  = pl.nparts;
    m->levels[n].path = 1;
    return 0;
}

// K0: stage the sentence-source tail, tokenise, check the encoding.  npos includes one virtual delimiter closing the last sentence.
int Trainer::tokenise(DevBuf<uint32_t>& tok, uint64_t& npos, uint32_t& nclasses) {
    // ---- the sentence source quirk (see include/colibri_b200.h: streamed)
    if (c->nbytes == 0) return set_err(COLIBRI_E_FORMAT, "Attempting to read pattern from file, but file is empty?");  // src/pattern.cpp:520-523
    // a corpus staged asynchronously is still on its way: chunk by chunk (below) or as a whole
    const bool piped = c->h2d_pending && c->chunk_bytes && !c->chunk_ev.empty();
    if (piped) CUDA_TRY(cudaStreamWaitEvent(s, c->chunk_ev[0], 0));  // (also orders the tail below after the padding memset of the corpus stream)
    else if (c->ev_h2d1) CUDA_TRY(cudaStreamWaitEvent(s, c->ev_h2d1, 0));
    size_t staged = c->nbytes;
    {
        uint8_t tail[16];
        memset(tail, 0x80, sizeof tail);
        size_t k = 0;
        if (!c->ends_with_delim) {
            if (o.streamed) tail[k++] = c->last_byte;  // Pattern(istream) stores the last byte twice when the final 0x00 is missing
            tail[k++] = 0;
        }
        staged += k;
        CUDA_TRY(cudaMemcpyAsync(c->body() + c->nbytes, tail, sizeof tail, cudaMemcpyHostToDevice, s));
    }
    const uint32_t nblocks = (uint32_t)(c->padded(staged) / kTokTile);

    TRY(d_stats.alloc(dev, 1));
    CUDA_TRY(cudaMemsetAsync(d_stats.p, 0, sizeof(DeviceStats), s));

    // ---- K0 tokenise
    int h = timer.begin(COLIBRI_T_TOKENISE);
    DevBuf<uint32_t> blk;
    TRY(blk.alloc(dev, nblocks + 1));
    uint64_t npos_real = 0;
    if (piped) {
        // The tokeniser follows the copy: a token belongs to the tile in which it ENDS and is decoded backwards, so the tiles of chunk k need
        // nothing of chunk k + 1.  The token array is sized by its upper bound (a token has at least one byte) because the count is not known yet.
        const uint64_t pos_bound = staged;
        tok_ext_cells = (tune.dense_dim && pos_bound >= tune.dense_min) ? (uint64_t)tune.dense_dim * tune.dense_dim : 0;
        if (pos_bound + 9 + 2 * tok_ext_cells >= 0xFFFFFFF0ull) {
            tok_ext_cells = 0;
            if (pos_bound + 9 >= 0xFFFFFFF0ull) return set_err(COLIBRI_E_CAPACITY, "corpus of %llu bytes; the device index is 32 bit", (unsigned long long)pos_bound);
        }
        TRY(tok.alloc(dev, pos_bound + 9 + 2 * tok_ext_cells));
        const uint32_t tiles_per_chunk = (uint32_t)(c->chunk_bytes / kTokTile);
        for (size_t k = 0; k < c->chunk_ev.size(); ++k) {
            const uint32_t t0 = (uint32_t)k * tiles_per_chunk;
            const uint32_t t1 = k + 1 == c->chunk_ev.size() ? nblocks : std::min<uint32_t>(nblocks, t0 + tiles_per_chunk);  // the last chunk takes the tail and the padding along
            if (t0 >= t1) break;
            if (k) CUDA_TRY(cudaStreamWaitEvent(s, c->chunk_ev[k], 0));
            const uint8_t* base = c->body() + (uint64_t)t0 * kTokTile;
            launches += launch_tokenise_count(s, base, 0, blk.p + t0, t1 - t0);
            launches += launch_scan_block_counts(s, blk.p + t0, t1 - t0, &d_stats.p->cursor, true);
            launches += launch_tokenise_write(s, base, 0, blk.p + t0, t1 - t0, tok.p, d_stats.p);
        }
        timer.end(h);
        TRY(read_stats());
        npos_real = h_stats.cursor;  // tokens + delimiters
        npos      = npos_real + 1;   // one virtual delimiter closes the last sentence
        if (!(tune.dense_dim && npos >= tune.dense_min)) tok_ext_cells = 0;
        CUDA_TRY(cudaMemsetAsync(tok.p + npos_real, 0, 8 * sizeof(uint32_t), s));
    } else {
        launches += launch_tokenise_count(s, c->body(), staged, blk.p, nblocks);
        launches += launch_scan_block_counts(s, blk.p, nblocks, &d_stats.p->cursor);
        TRY(read_stats());
        npos_real = h_stats.cursor;  // tokens + delimiters
        if (npos_real >= 0xFFFFFFF0ull) return set_err(COLIBRI_E_CAPACITY, "corpus has %llu positions; the device index is 32 bit", (unsigned long long)npos_real);
        npos = npos_real + 1;  // one virtual delimiter closes the last sentence
        // spare room behind the tokens: the class pairs of the surviving dense bigrams (kernels.cu: prune_dense_kernel)
        tok_ext_cells = (tune.dense_dim && npos >= tune.dense_min) ? (uint64_t)tune.dense_dim * tune.dense_dim : 0;
        if (npos + 8 + 2 * tok_ext_cells >= 0xFFFFFFF0ull) tok_ext_cells = 0;
        TRY(tok.alloc(dev, npos + 8 + 2 * tok_ext_cells));
        CUDA_TRY(cudaMemsetAsync(tok.p + npos_real, 0, 8 * sizeof(uint32_t), s));
        launches += launch_tokenise_write(s, c->body(), staged, blk.p, nblocks, tok.p, d_stats.p);
        timer.end(h);
        TRY(read_stats());
    }
    if (h_stats.errflags & kErrTokenTooLong) return set_err(COLIBRI_E_FORMAT, "corpus contains a class wider than 32 bits / 5 bytes");
    if (h_stats.errflags & kErrNonCanonical) return set_err(COLIBRI_E_FORMAT, "corpus contains a non-canonical class encoding (multi-byte token ending in 0x00)");
    if (h_stats.errflags & kErrReservedClass)
        return set_err(COLIBRI_E_UNSUPPORTED, "corpus contains the reserved skip/flex classes (3, 4) as running text; not on the device path");
    m->totaltokens = h_stats.totaltokens;
    nclasses       = h_stats.maxclass + 1;
    m->counters[0] = npos_real;
    m->counters[1] = c->nbytes;
    return 0;
}

int Trainer::run() {
    CUDA_TRY(cudaSetDevice(dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));  // (cudaGetDeviceProperties costs milliseconds per call)
    timer.s   = s;
    timer.dev = dev;
    int h_total = timer.begin(COLIBRI_T_TOTAL);
    DevBuf<uint32_t> tok;
    uint64_t         npos = 0;
    uint32_t         nclasses = 0;
    int              h = -1;
    TRY(tokenise(tok, npos, nclasses));
    g_trace.mark("tokenise");
    tok_for_sink  = tok.p;
    sink_maxclass = nclasses ? nclasses - 1 : 0;
    // a level can be handed to the caller as soon as it is pruned unless a later rule may still drop it (MINLENGTH clean-up, :1221-1229, :1337-1341)
    const bool stream_levels = sink != nullptr && o.MINLENGTH <= 1 && o.MINTOKENS > 1;

    indexed = o.model_type == COLIBRI_INDEXEDPATTERNMODEL;
    if (indexed) {
        int hi = timer.begin(COLIBRI_T_INDEX);
        TRY(prepare_index(tok.p, npos));
        timer.end(hi);
    }
    const uint32_t t  = (uint32_t)o.MINTOKENS;
    const uint32_t t1 = (uint32_t)std::max(o.MINTOKENS, o.MINTOKENS_UNIGRAMS);  // what higher orders require of their unigrams (:1094-1104)
    const uint32_t ts = o.MINSKIPTYPES > 1 ? (uint32_t)o.MINTOKENS_SKIPGRAMS : t;  // PatternModel::pruneskipgrams returns early when minskiptypes <= 1 (:2170-2171)
    const bool     skipgrams = o.DOSKIPGRAMS_EXHAUSTIVE != 0;
    const bool     indexed_skip = indexed && o.DOSKIPGRAMS != 0;  // trainskipgrams after the n-gram levels
    const bool     keep_all_ids = skipgrams || indexed_skip;

    // ---- K1 unigrams
    h = timer.begin(COLIBRI_T_UNIGRAMS);
    DevBuf<uint32_t> count1;
    TRY(count1.alloc(dev, nclasses));
    CUDA_TRY(cudaMemsetAsync(count1.p, 0, (size_t)nclasses * 4, s));
    launches += launch_unigram_hist(s, tok.p, npos, count1.p, nclasses, sms);
    TRY(zero_stats());
    {
        Segment sg;
        sg.n = 1;
        uint64_t bound = std::min<uint64_t>(nclasses, m->totaltokens / std::max<uint32_t>(t, 1) + 1);
        TRY(sg.pos.alloc(dev, bound));
        TRY(sg.cnt.alloc(dev, bound));
        DevBuf<uint32_t> class_index;
        if (indexed) TRY(class_index.alloc(dev, nclasses));
        launches += launch_unigram_prune(s, count1.p, nclasses, t, sg.pos.p, sg.cnt.p, 0, d_stats.p, 1, 0, indexed ? class_index.p : nullptr);
        TRY(read_stats());
        sg.count = h_stats.kept;
        if (indexed) {
            int hi = timer.begin(COLIBRI_T_INDEX);
            TRY(build_refs(sg, tok.p, class_index.p, true, npos, h_stats.kept_occ));
            timer.end(hi);
        }
        if (h_stats.found == 0) sg.count = 0;
        segs.push_back(std::move(sg));
        if (stream_levels) {
            int he = timer.begin(COLIBRI_T_EXPORT);
            TRY(emit_segment(segs.back()));
            timer.end(he);
        }
    }
    timer.end(h);
    m->counters[6] = m->totaltokens;
    g_trace.mark("unigrams");
    const DeviceStats uni = h_stats;  // found / kept / kept occurrences of the unigram level
    // tokens of the classes the dense square covers (sizes the occurrence filter of level 2)
    uint64_t dense_tokens = 0;
    if (tok_ext_cells && nclasses) {
        TRY(zero_stats());
        launches += launch_sum_u32(s, count1.p, std::min<uint32_t>(tune.dense_dim, nclasses), &d_stats.p->found);
        TRY(read_stats());
        dense_tokens = h_stats.found;
    }
    std::vector<PassStat> passes;
    uint64_t prev_kept = uni.kept, prev_occ = uni.kept_occ;
    if (uni.found == 0) {  // nothing at all ("None found", :1189-1194): an empty model
        segs.clear();
    } else {
        passes.push_back({1, uni.found, 0, uni.found - uni.kept});
        m->maxn = 1;
        m->minn = 1;
        m->totaltypes = uni.found;  // :1199-1201 (t > 1) and totalwordtypesingroup(NGRAM,1) (t == 1, :1202-1208)
    }
    int last_pass = uni.found ? 1 : 0;

    // ---- levels n >= 2
    std::vector<DevBuf<uint32_t>> ids(2);  // ids[k] = id array of level k (all kept when skipgrams need their parts, else ping-pong)
    DevBuf<NgramSlot>       table;
    DevBuf<uint32_t>        bitmap;  // survivor bit per table slot of the level just pruned
    DevBuf<uint32_t>        filter;  // 2-bit occurrence filter of the level being counted
    DevBuf<uint32_t>        filter1; // its "hit twice" bits, packed (Tuning::filter_1bit)
    DevBuf<uint32_t>        class_bits;  // one bit per class: kept at level 1 (launch_make_id1_hist)
    DevBuf<uint32_t>        slot_index;  // indexed models: table slot -> survivor index + 1
    DevBuf<uint32_t>        list_cur, list_next;  // list mode: positions whose newest id is non-zero (see kernels.cu: load_window)
    uint64_t                nlist = 0;
    bool                    list_valid = false;
    DevBuf<SkipSlot>        sktable;
    DevBuf<const uint32_t*> d_idptrs;
    DevBuf<SkipMask>        d_masks;
    bool pre_hist2 = false;  // level 2's pass A (dense square, first-level histogram) was taken by the sweep that writes the level-1 ids
    if (last_pass == 1 && o.MAXLENGTH >= 2) {
        TRY(ids[1].alloc(dev, npos + 8));
        // what level 2 will decide below, decided here already: a dense, partitioned level 2 lets one kernel do both sweeps
        uint64_t bound2 = prev_occ;
        if (prev_kept < (1ull << 31)) bound2 = std::min(bound2, prev_kept * prev_kept);
        const uint64_t wbound2 = std::min<uint64_t>(prev_occ, npos);
        const uint32_t dense2  = (bound2 >= tune.dense_min && tok_ext_cells) ? std::min<uint32_t>(tune.dense_dim, nclasses) : 0;
        if (dense2 && bound2 > 0 && tune.use_partition(wbound2, true) && wbound2 < 0xFFFFFFF0ull && m->totaltokens && !getenv("COLIBRI_B200_NO_FUSE_ID1")) {
            const double   f  = (double)dense_tokens / (double)m->totaltokens;
            const PartPlan pl = part_plan(plan_bound(wbound2, 1.1 * (1.0 - f * f)));
            const uint64_t dense_cells2 = (uint64_t)dense2 * dense2;
            TRY(part_small_layout(pl));
            if (filter.n < dense_cells2 + 8) TRY(filter.alloc(dev, dense_cells2 + 8));
            int hc = timer.begin(COLIBRI_T_COUNT, 2);
            CUDA_TRY(cudaMemsetAsync(part_small.p, 0, ((size_t)1 << pl.b1) * sizeof(uint32_t), s));
            CUDA_TRY(cudaMemsetAsync(filter.p, 0, dense_cells2 * sizeof(uint32_t), s));
            TRY(zero_stats());
            TRY(class_bits.alloc(dev, (uint64_t)(nclasses + 255) / 256 * 8 + 8));
            launches += launch_make_id1_hist(s, tok.p, npos, count1.p, nclasses, t1, class_bits.p, ids[1].p, dense2, filter.p, part_small.p, pl, d_stats.p, sms);
            timer.end(hc);
            pre_hist2 = true;
            m->levels[2].fused_id1 = 1;
        } else {
            launches += launch_make_id1(s, tok.p, npos + 1, count1.p, t1, ids[1].p);
        }
    }
    for (int n = 2; n <= o.MAXLENGTH && last_pass == n - 1; ++n) {
        // every valid window starts at a position whose (n-1)-gram survived, and is a pair of surviving (n-1)-grams
        uint64_t bound = prev_occ;
        if (prev_kept < (1ull << 31)) bound = std::min(bound, prev_kept * prev_kept);
        if (bound == 0) break;  // nothing can be found
        if ((int)ids.size() <= n) ids.resize(n + 1);
        DevBuf<uint32_t>& prev = ids[n - 1];
        DevBuf<uint32_t>& cur  = ids[n];
        if (!cur.p) TRY(cur.alloc(dev, npos + 8));
        CUDA_TRY(cudaMemsetAsync(cur.p + npos, 0, 8 * sizeof(uint32_t), s));

        // ---- list mode: the previous level's relabel step left the positions whose (n-1)-gram survived
        const bool      use_list = list_valid;
        const uint32_t* list     = use_list ? list_cur.p : nullptr;
        if (use_list) CUDA_TRY(cudaMemsetAsync(cur.p, 0, npos * sizeof(uint32_t), s));  // only the windows that exist are written

        // ---- dense pairs (level 2 of a large corpus): the ids of level 1 are the class numbers, frequent classes are the small ones
        uint32_t dense = 0;
        if (n == 2 && !use_list && bound >= tune.dense_min && tok_ext_cells) dense = std::min<uint32_t>(tune.dense_dim, nclasses);
        const uint64_t dense_cells = (uint64_t)dense * dense;

        // ---- large levels: radix-partitioned counting in shared memory (partition.cu); the HBM-table path below serves the small ones
        const uint64_t wbound = use_list ? nlist : std::min<uint64_t>(prev_occ, npos);  // valid windows start where the (n-1)-gram survived
        bool     parted = false;
        uint64_t windows = 0, singles = 0;
        Segment  sg;
        sg.n = n;
        if (tune.use_partition(wbound, dense != 0) && wbound < 0xFFFFFFF0ull) {
            if (filter.n < dense_cells + 8) TRY(filter.alloc(dev, dense_cells + 8));
            bool overflow = false;
            double share = 1.0;  // windows of two dense classes never become records: their share follows from the class histogram
            if (dense && m->totaltokens) {
                const double f = (double)dense_tokens / (double)m->totaltokens;
                share = 1.1 * (1.0 - f * f);
            }
            TRY(level_partitioned(n, prev.p, cur.p, npos, list, nlist, wbound, dense, filter.p, t, tok.p + npos + 8, sg, overflow, share, indexed ? &slot_index : nullptr, n == 2 && pre_hist2));
            parted = !overflow;
            if (parted) {
                windows = h_stats.valid_windows;
                singles = h_stats.singletons;
                ngram_upserts += windows;
                m->levels[n].windows = windows;
                m->levels[n].singles = singles;
                m->levels[n].items   = use_list ? nlist : npos;
            } else if (use_list) {
                CUDA_TRY(cudaMemsetAsync(cur.p, 0, npos * sizeof(uint32_t), s));
            }
        }
        uint64_t found = 0, kept = 0, occ = 0;
        if (parted) {
            found = h_stats.found; kept = h_stats.kept; occ = h_stats.kept_occ;
        } else {
        // ---- occurrence filter (t >= 2): the 2-bit counters and the dense square share one buffer, pinned in L2 for the level
        const bool use_filter = tune.use_filter(t, bound);
        uint64_t   nbuckets = 0, cap = 0;
        if (use_filter) {
            // windows of two dense classes never look at the filter: size it for the rest (share estimated from the class histogram)
            uint64_t fbound = bound;
            if (dense && m->totaltokens && getenv("COLIBRI_B200_FILTER_SHRINK")) {
                const double f = (double)dense_tokens / (double)m->totaltokens;
                fbound = (uint64_t)((double)bound * std::min(1.0, 1.05 * (1.0 - f * f))) + 1024;
            }
            nbuckets = tune.filter_buckets(fbound);
        }
        const uint64_t filter_words = nbuckets / 16;
        if (filter.n < filter_words + dense_cells + 8) TRY(filter.alloc(dev, filter_words + dense_cells + 8));
        uint32_t* dense_cnt = filter.p + filter_words;
        if (use_filter || dense) TRY(l2_pin(filter.p, (filter_words + dense_cells) * sizeof(uint32_t)));
        if (use_filter) {
            int hf = timer.begin(COLIBRI_T_COUNT, n);
            CUDA_TRY(cudaMemsetAsync(filter.p, 0, filter_words * sizeof(uint32_t), s));
            TRY(zero_stats());
            launches += launch_ngram_filter(s, prev.p, npos, filter.p, nbuckets, d_stats.p, sms, dense, list, nlist);
            timer.end(hf);
            TRY(read_stats());
            // keys that reach the table live in buckets hit at least twice; there are at most ~2 such keys per bucket
            // when the filter is crowded with singletons, ~1 otherwise.  Overflow is detected and retried.
            cap = std::max<uint64_t>(1024, 3 * h_stats.found + 1024);
            if (h_stats.found * 8 > nbuckets) cap = bound + bound / 2 + 16;  // a saturated filter says nothing about the number of keys
            cap = std::min(cap, std::max<uint64_t>(64, bound + bound / 2 + 16));
            if (tune.filter_1bit && nbuckets >= 64) {
                if (filter1.n < nbuckets / 32 + 8) TRY(filter1.alloc(dev, nbuckets / 32 + 8));
                int hf1 = timer.begin(COLIBRI_T_COUNT, n);
                launches += launch_filter_to_bitmap(s, filter.p, nbuckets, filter1.p);
                timer.end(hf1);
            }
        } else {
            cap = std::max<uint64_t>(64, bound + bound / 2 + 16);  // load factor <= 2/3
        }
        const bool onebit = use_filter && tune.filter_1bit && nbuckets >= 64;
        const uint64_t cap_max = std::max<uint64_t>(64, bound + bound / 2 + 16);  // windows <= bound: a table this large cannot fill up
        for (;;) {
            cap = (cap + 31) / 32 * 32;  // the dense cells' survivor bits start on a bitmap word
            if (cap + dense_cells >= 0xFFFFFFF0ull)
                return set_err(COLIBRI_E_CAPACITY, "level %d needs %llu table slots; slot ids are 32 bit", n, (unsigned long long)(cap + dense_cells));
            if (table.n < cap) TRY(table.alloc(dev, cap));
            int hp0 = timer.begin(COLIBRI_T_PRUNE);
            CUDA_TRY(cudaMemsetAsync(table.p, 0, cap * sizeof(NgramSlot), s));
            if (dense) CUDA_TRY(cudaMemsetAsync(dense_cnt, 0, dense_cells * sizeof(uint32_t), s));
            timer.end(hp0);
            slots_init += cap + dense_cells / 4;
            TRY(zero_stats());
            int hc = timer.begin(COLIBRI_T_COUNT, n);
            const bool hot = tune.use_hot(bound);
            launches += launch_count_ngrams(s, prev.p, cur.p, npos, table.p, cap, d_stats.p, sms, use_filter ? (onebit ? filter1.p : filter.p) : nullptr, nbuckets, hot, dense, list, nlist, dense_cnt,
                                            onebit);
            timer.end(hc);
            CUDA_TRY(cudaGetLastError());
            TRY(fetch_stats());
            if (h_stats.errflags & kErrTableFull) {  // the estimate was too small: clear the flag and go again with more slots
                if (cap >= cap_max) return set_err(COLIBRI_E_CAPACITY, "device hash table overflow at level %d", n);
                CUDA_TRY(cudaMemsetAsync(&d_stats.p->errflags, 0, sizeof(unsigned int), s));
                if (use_list) CUDA_TRY(cudaMemsetAsync(cur.p, 0, npos * sizeof(uint32_t), s));
                cap = std::min(cap_max, cap * 4);
                continue;
            }
            windows = h_stats.valid_windows;
            singles = h_stats.singletons;
            break;
        }
        ngram_upserts += windows - singles;
        filtered_windows += singles;
        m->levels[n].windows = windows;
        m->levels[n].cap     = cap;
        m->levels[n].path    = 0;
        m->levels[n].filtered = use_filter ? 1 : 0;
        m->levels[n].singles = singles;
        m->levels[n].items   = use_list ? nlist : npos;

        int hp = timer.begin(COLIBRI_T_PRUNE);
        uint64_t sv_bound = windows / std::max<uint32_t>(t, 1) + 1;
        TRY(sg.pos.alloc(dev, sv_bound));
        TRY(sg.cnt.alloc(dev, sv_bound));
        const uint64_t slots_total = cap + dense_cells;  // ids cap + 1 .. cap + dense^2 name the cells of the dense square
        if (bitmap.n < slots_total / 32 + 8) TRY(bitmap.alloc(dev, slots_total / 32 + 8));
        if (indexed && slot_index.n < slots_total) TRY(slot_index.alloc(dev, slots_total));
        launches += launch_prune_ngrams(s, table.p, cap, t, sg.pos.p, sg.cnt.p, bitmap.p, d_stats.p, sms, indexed ? slot_index.p : nullptr);
        if (dense)
            launches += launch_prune_dense(s, dense_cnt, dense, cap, t, sg.pos.p, sg.cnt.p, bitmap.p, indexed ? slot_index.p : nullptr, tok.p + npos + 8, (uint32_t)(npos + 8), d_stats.p, sms);
        timer.end(hp);
        TRY(read_stats());
        if (use_filter || dense) l2_unpin();
        // a window the filter held back is a distinct n-gram with exactly one occurrence: found, and pruned (t >= 2)
        found = h_stats.found + singles; kept = h_stats.kept; occ = h_stats.kept_occ;
        }  // HBM-table path
        sg.count = kept;
        int hp = -1;
        if (indexed && kept > 0) {  // IndexedPatternModel::add (:2789-2800) + posttrain sort (:2699-2705)
            int hi = timer.begin(COLIBRI_T_INDEX);
            TRY(build_refs(sg, cur.p, slot_index.p, false, npos, occ, indexed_skip && n >= 3));
            timer.end(hi);
        }

        // ---- exhaustive skipgrams of this level (:1163-1171)
        uint64_t foundskip = 0, keptskip = 0;
        Segment  sk;
        if (skipgrams && n >= 3 && windows > 0) {
            std::vector<SkipMask> masks;
            TRY(skip_masks(n, o.MAXSKIPS, masks));
            if (!masks.empty()) {
                uint64_t sbound = 0;  // distinct keys <= one per (window, mask) plus one helper per folding round
                for (auto& sm : masks) sbound += windows * (1 + (sm.nparts > 3 ? (sm.nparts - 2) / 2 : 0));
                uint64_t scap   = std::max<uint64_t>(64, sbound + sbound / 2 + 16);
                size_t   freeb = 0, totalb = 0;
                cudaMemGetInfo(&freeb, &totalb);
                if (sktable.n < scap) {
                    sktable.reset();
                    if (scap * sizeof(SkipSlot) > freeb + g_pool[dev & 15].cached)
                        return set_err(COLIBRI_E_CAPACITY, "skipgram table of level %d needs %.1f GB", n, scap * 32.0 / 1e9);
                    TRY(sktable.alloc(dev, scap));
                }
                std::vector<const uint32_t*> ptrs(n, nullptr);
                for (int k = 1; k < n; ++k) ptrs[k] = ids[k].p;
                TRY(d_idptrs.alloc(dev, n));
                TRY(d_masks.alloc(dev, masks.size()));
                CUDA_TRY(cudaMemcpyAsync(d_idptrs.p, ptrs.data(), n * sizeof(uint32_t*), cudaMemcpyHostToDevice, s));
                CUDA_TRY(cudaMemcpyAsync(d_masks.p, masks.data(), masks.size() * sizeof(SkipMask), cudaMemcpyHostToDevice, s));
                CUDA_TRY(cudaMemsetAsync(sktable.p, 0, scap * sizeof(SkipSlot), s));
                slots_init += scap * 2;
                TRY(zero_stats());
                int hs = timer.begin(COLIBRI_T_SKIPGRAMS, n);
                launches += launch_count_skipgrams(s, d_idptrs.p, n, d_masks.p, (int)masks.size(), npos, sktable.p, scap, d_stats.p, sms);
                timer.end(hs);
                TRY(read_stats());  // also keeps ptrs/masks alive until the copies are done
                skip_upserts += h_stats.valid_windows;
                sk.n    = n;
                sk.skip = true;
                uint64_t kb = h_stats.valid_windows / std::max<uint32_t>(ts, 1) + 1;
                TRY(sk.pos.alloc(dev, kb));
                TRY(sk.cnt.alloc(dev, kb));
                TRY(sk.mask.alloc(dev, kb));
                TRY(zero_stats());
                hp = timer.begin(COLIBRI_T_PRUNE);
                launches += launch_prune_skipgrams(s, sktable.p, scap, ts, sk.pos.p, sk.cnt.p, sk.mask.p, d_stats.p, sms);
                timer.end(hp);
                TRY(read_stats());
                foundskip = h_stats.found;
                keptskip  = h_stats.kept;
                sk.count  = keptskip;
            }
        }
        if (found == 0 && foundskip == 0) break;  // "None found" (:1189-1194): maxn/minn untouched, no prune
        last_pass = n;
        m->maxn   = std::max(m->maxn, n);
        m->minn   = std::min(m->minn, n);
        if (foundskip) m->hasskipgrams = 1;
        passes.push_back({(uint64_t)n, found, foundskip, (found - kept) + (foundskip - keptskip)});
        segs.push_back(std::move(sg));
        if (stream_levels) {
            g_trace.mark("(");
            int he = timer.begin(COLIBRI_T_EXPORT);
            TRY(emit_segment(segs.back()));
            timer.end(he);
            g_trace.mark("emit)");
        }
        if (sk.skip) {
            segs.push_back(std::move(sk));
            if (stream_levels) {
                int he = timer.begin(COLIBRI_T_EXPORT);
                TRY(emit_segment(segs.back()));
                timer.end(he);
            }
        }

        bool next_list = false;
        if ((n < o.MAXLENGTH || indexed_skip) && kept > 0) {
            if (t > 1) {
                // the next level runs from a position list when this level's survivors cover a small part of the corpus
                next_list = n < o.MAXLENGTH && tune.sparse_div > 0 && occ * tune.sparse_div <= npos && occ < 0xFFFFFFF0ull;
                hp = timer.begin(COLIBRI_T_PRUNE);
                if (next_list) {
                    if (list_next.n < occ + 8) TRY(list_next.alloc(dev, occ + 8));
                    CUDA_TRY(cudaMemsetAsync(&d_stats.p->cursor, 0, sizeof(unsigned long long), s));
                }
                if (!parted) launches += launch_relabel(s, cur.p, npos, bitmap.p, list, nlist, next_list ? list_next.p : nullptr, &d_stats.p->cursor, sms);
                else if (next_list) launches += launch_compact_nonzero(s, cur.p, npos, list_next.p, &d_stats.p->cursor);  // the ids are final already
                timer.end(hp);
                if (next_list) {
                    TRY(read_stats());
                    if (h_stats.cursor != occ)
                        return set_err(COLIBRI_E_CUDA, "level %d: %llu surviving positions listed, %llu occurrences kept", n, (unsigned long long)h_stats.cursor, (unsigned long long)occ);
                    std::swap(list_cur, list_next);
                    nlist = occ;
                }
            }
        }
        list_valid = next_list;
        {
            static const char* const kLevelNames[] = {"L0", "L1", "L2", "L3", "L4", "L5", "L6", "L7", "L8", "L9+"};
            g_trace.mark(kLevelNames[std::min(n, 9)]);
        }
        if (!keep_all_ids) ids[n - 1].reset();  // ping-pong: only the newest level is needed
        prev_kept = kept;
        prev_occ  = occ;
        if (kept == 0) {  // the next pass cannot find anything: it would print "None found" and stop
            break;
        }
    }

    // ---- indexed models: skipgrams from the surviving n-grams (train() tail, :1271-1273 -> trainskipgrams :2969-3010)
    if (indexed_skip) {
        std::vector<Segment> extra;
        for (int n = 3; n <= o.MAXLENGTH; ++n) {
            Segment* ng = nullptr;
            for (auto& sg : segs)
                if (sg.n == n && !sg.skip) ng = &sg;
            uint64_t foundskip = 0, keptskip = 0;
            Segment  sk;
            if (ng != nullptr && ng->count > 0) TRY(indexed_skipgrams(n, *ng, ids, foundskip, keptskip, sk));
            if (foundskip == 0) break;  // " None found"
            m->hasskipgrams = 1;
            passes.push_back({(uint64_t)n, 0, foundskip, foundskip - keptskip});
            if (keptskip) extra.push_back(std::move(sk));
        }
        for (auto& sk : extra) segs.push_back(std::move(sk));
    }

    // ---- which levels end up in the model
    std::vector<Segment> keep;
    for (auto& sg : segs) {
        bool drop = false;
        if (o.MINTOKENS > 1) {
            if (!skipgrams && !o.DOSKIPGRAMS) {
                // :1221-1229: after pass n, level n-1 is emptied when it is below MINLENGTH
                int k = sg.n;
                if (k < o.MINLENGTH && last_pass >= k + 1 && k != o.MAXBACKOFFLENGTH && !(k == 1 && o.MINTOKENS_UNIGRAMS > o.MINTOKENS)) drop = true;
                // :1278-1280: the level the back-off rule needed until the end goes last, `prune(-1, MAXBACKOFFLENGTH)` when it is below MINLENGTH
                if (k == o.MAXBACKOFFLENGTH && o.MAXBACKOFFLENGTH < o.MINLENGTH) drop = true;
            } else if (o.MINLENGTH > 1 && sg.n <= o.MINLENGTH - 1) {
                drop = true;  // :1337-1341 prunebylength
            }
        }
        if (!drop && sg.count > 0) keep.push_back(std::move(sg));
    }
    segs = std::move(keep);

    // reference reports a single pass when MINTOKENS == 1 (all lengths extracted in one scan, :1062-1072, :1246-1247)
    if (o.MINTOKENS == 1 && !passes.empty()) {
        PassStat one{1, 0, 0, 0};
        for (auto& p : passes) {
            one.found += p.found;
            one.foundskip += p.foundskip;
            one.pruned += p.pruned;
        }
        passes.assign(1, one);
        // postread (:1274-1277): maxn/minn come from the stored patterns
        m->maxn = 0;
        m->minn = 999;
        for (auto& sg : segs) {
            m->maxn = std::max(m->maxn, sg.n);
            m->minn = std::min(m->minn, sg.n);
        }
    }
    m->passes = passes;

    h = timer.begin(COLIBRI_T_EXPORT);
    if (sink != nullptr) {
        if (!stream_levels)
            for (auto& sg : segs) TRY(emit_segment(sg));
        timer.end(h);
        timer.end(h_total);
        g_trace.mark("emit-last");
        CUDA_TRY(cudaStreamSynchronize(s));
        g_trace.mark("sync-main");
        TRY(flush_sink(true));
        g_trace.mark("sync-copy");
        m->npatterns = sink->npat;
        m->keybytes  = sink->nbytes;
    } else {
        TRY(colibri::export_segments(dev, s, segs, tok.p, m, launches));
        timer.end(h);
        timer.end(h_total);
        CUDA_TRY(cudaStreamSynchronize(s));
    }
    {
        std::map<int, double> lvl;
        timer.resolve(m->ms, &lvl);
        for (auto& kv : lvl) m->levels[kv.first].ms = kv.second;
    }
    m->counters[2] = launches;
    m->counters[3] = ngram_upserts;
    m->counters[4] = skip_upserts;
    m->counters[5] = slots_init;
    m->counters[7] = g_pool[dev & 15].peak;
    return 0;
}


// The reverse index of a model over a corpus (colibri_b200_rindex_build): the matching half of constrained training -- every window of every
// length the model holds is looked up in the model's pattern index -- with all match arrays kept, plus the sentence tables of indexed models.
int Trainer::build_rindex(colibri_b200_model* cm, colibri_b200_rindex* r) {
    CUDA_TRY(cudaSetDevice(dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    timer.s   = s;
    timer.dev = dev;
    TRY(ensure_closure(cm, &launches));
    uint64_t npos = 0;
    uint32_t nclasses = 0;
    TRY(tokenise(r->tok, npos, nclasses));
    TRY(prepare_index(r->tok.p, npos));
    r->npos        = npos;
    r->nsentences  = npos - m->totaltokens;  // delimiters, the virtual one closing the last sentence included
    r->sent_before = std::move(sent_before);
    r->sent_start  = std::move(sent_start);
    r->minn        = cm->minn;
    r->maxn        = cm->maxn;
    const uint64_t np = cm->npatterns;
    for (int n = 1; n <= cm->maxn && n <= 255; ++n)
        if (np && cm->meta.nhist[n]) r->lengths.push_back(n);
    DevBuf<uint32_t> counts;  // the matching kernels count as they go; the counts are not wanted here
    TRY(counts.alloc(dev, std::max<uint64_t>(np, 1)));
    CUDA_TRY(cudaMemsetAsync(counts.p, 0, std::max<uint64_t>(np, 1) * sizeof(uint32_t), s));
    r->match.resize(r->lengths.size());
    for (auto& mb : r->match) {
        TRY(mb.alloc(dev, npos + 8));
        CUDA_TRY(cudaMemsetAsync(mb.p + npos, 0, 8 * sizeof(uint32_t), s));
    }
    TRY(zero_stats());
    cm->index_counts_dirty = true;
    const bool chain = !getenv("COLIBRI_B200_NO_CHAIN");
    DevBuf<uint32_t> hist;
    for (size_t k = 0; k < r->lengths.size(); ++k) {
        const int       n    = r->lengths[k];
        uint32_t*       cur  = r->match[k].p;
        const uint32_t* prev = (k > 0 && r->lengths[k - 1] == n - 1 && chain) ? r->match[k - 1].p : nullptr;
        const bool use_prefix = prev && cm->prefix_open[n] == 0, use_suffix = prev && cm->suffix_open[n] == 0;
        if (n == 1 && cm->uni_classes) {
            launches += launch_unigram_match(s, r->tok.p, npos, cm->d_uni.p, cm->uni_classes, cur);
        } else {
            launches += launch_constrained_match(s, r->tok.p, npos, n, cm->d_keys.p, cm->d_off.p, cm->d_index.p, cm->index_cap, cm->d_presence.p, cm->presence_bits, counts.p, cur, prev,
                                                 use_prefix, use_suffix, d_stats.p, sms);
        }
    }
    launches += launch_collect_slot_counts(s, cm->d_index.p, cm->index_cap, counts.p);
    TRY(read_stats());
    cm->index_counts_dirty = false;
    std::vector<const uint32_t*> ptrs;
    std::vector<uint32_t>        lens;
    for (size_t k = 0; k < r->lengths.size(); ++k) {
        ptrs.push_back(r->match[k].p);
        lens.push_back((uint32_t)r->lengths[k]);
    }
    TRY(r->d_match_ptrs.alloc(dev, std::max<size_t>(ptrs.size(), 1)));
    TRY(r->d_lengths.alloc(dev, std::max<size_t>(lens.size(), 1)));
    if (!ptrs.empty()) {
        CUDA_TRY(cudaMemcpyAsync(r->d_match_ptrs.p, ptrs.data(), ptrs.size() * sizeof(uint32_t*), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(r->d_lengths.p, lens.data(), lens.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
}

// PatternModel::train with constrainbymodel != NULL (reference include/patternmodel.h:880-1345): ONE scan of the corpus that extracts every
// window of MINLENGTH..MAXLENGTH tokens (:1064-1072) and counts it iff the constraint model has it (:1088-1089); then prune(MINTOKENS, 0)
// (:1211-1218) and stop (:1246-1247).  Here: one launch per pattern length that the constraint set actually holds; the window's varint bytes
// are rebuilt in registers, hashed with SpookyV2 (Pattern::hash) and probed in the set's HBM index (pattern_index.cu); survivors are
// compacted out of the set's own blob.  Indexed models remember the match of every position and build the occurrence lists with the
// ordered-pairs + stable radix sort of index.cu, one length at a time (survivors are ordered by length so the lists concatenate).
int Trainer::run_constrained(colibri_b200_model* cm, bool inplace, int phase, uint32_t* ext_counts, uint64_t ext_tokens) {
    CUDA_TRY(cudaSetDevice(dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    timer.s   = s;
    timer.dev = dev;
    int h_total = timer.begin(COLIBRI_T_TOTAL);
    indexed = o.model_type == COLIBRI_INDEXEDPATTERNMODEL;
    if (phase != 0 && (indexed || ext_counts == nullptr)) return set_err(COLIBRI_E_UNSUPPORTED, "sharded constrained training is for unindexed models");
    TRY(ensure_closure(cm, &launches));
    DevBuf<uint32_t> tok;
    uint64_t         npos = 0;
    uint32_t         nclasses = 0;
    if (phase != 2) {
        TRY(tokenise(tok, npos, nclasses));
    } else {
        TRY(d_stats.alloc(dev, 1));
        CUDA_TRY(cudaMemsetAsync(d_stats.p, 0, sizeof(DeviceStats), s));
        m->totaltokens = ext_tokens;
    }
    const uint64_t corpus_tokens = m->totaltokens;
    if (indexed) {
        int hi = timer.begin(COLIBRI_T_INDEX);
        TRY(prepare_index(tok.p, npos));
        timer.end(hi);
    }
    const uint64_t np = cm->npatterns;
    const uint32_t t  = (uint32_t)o.MINTOKENS;
    std::vector<int> lengths;  // the window lengths worth a scan
    for (int n = o.MINLENGTH; n <= o.MAXLENGTH && n <= 255; ++n)
        if (np && cm->meta.nhist[n]) lengths.push_back(n);

    DevBuf<uint32_t> counts_own, flags, kmap;
    if (ext_counts == nullptr) {
        TRY(counts_own.alloc(dev, std::max<uint64_t>(np, 1)));
        CUDA_TRY(cudaMemsetAsync(counts_own.p, 0, std::max<uint64_t>(np, 1) * sizeof(uint32_t), s));
        ext_counts = counts_own.p;
    }
    struct { uint32_t* p; } counts{ext_counts};  // this call's counter array: its own, or the caller's (sharded run)
    TRY(flags.alloc(dev, np + 1));
    // match[k][p] = pattern (index + 1) of the window of lengths[k] tokens at p.  Indexed models keep every level (the occurrence lists are
    // built from them); otherwise two buffers alternate: level n only looks at level n-1, to skip windows whose prefix / suffix did not match
    const bool chain = !getenv("COLIBRI_B200_NO_CHAIN");
    DevBuf<uint32_t> count1_scratch;
    uint64_t         unigram_windows = 0;
    std::vector<DevBuf<uint32_t>> match(phase == 2 ? 0 : (indexed ? lengths.size() : std::min<size_t>(lengths.size(), 2)));
    for (auto& mb : match) {
        TRY(mb.alloc(dev, npos + 8));
        CUDA_TRY(cudaMemsetAsync(mb.p + npos, 0, 8 * sizeof(uint32_t), s));
    }
    TRY(zero_stats());
    if (phase != 2) cm->index_counts_dirty = true;  // until the slot counters have been collected (an error in between forces a rebuild of the index)
    for (size_t k = 0; k < lengths.size() && phase != 2; ++k) {
        const int n   = lengths[k];
        uint32_t* cur = match[indexed ? k : k % 2].p;
        const uint32_t* prev = (k > 0 && lengths[k - 1] == n - 1 && chain) ? match[indexed ? k - 1 : (k - 1) % 2].p : nullptr;
        const bool use_prefix = prev && cm->prefix_open[n] == 0, use_suffix = prev && cm->suffix_open[n] == 0;
        if (!indexed && !(k + 1 < lengths.size() && lengths[k + 1] == n + 1 && chain)) cur = nullptr;  // nobody will read it
        int hc = timer.begin(COLIBRI_T_COUNT, n);
        if (n == 1 && cm->uni_classes) {
            // the class histogram of unconstrained training, handed to the unigram patterns
            DevBuf<uint32_t>& hist = count1_scratch;
            TRY(hist.alloc(dev, nclasses));
            CUDA_TRY(cudaMemsetAsync(hist.p, 0, (size_t)nclasses * sizeof(uint32_t), s));
            launches += launch_unigram_hist(s, tok.p, npos, hist.p, nclasses, sms);
            launches += launch_unigram_apply(s, hist.p, cm->d_uni.p, std::min<uint32_t>(nclasses, cm->uni_classes), counts.p);
            if (cur) launches += launch_unigram_match(s, tok.p, npos, cm->d_uni.p, cm->uni_classes, cur);
            unigram_windows = corpus_tokens;
        } else
            launches += launch_constrained_match(s, tok.p, npos, n, cm->d_keys.p, cm->d_off.p, cm->d_index.p, cm->index_cap, cm->d_presence.p, cm->presence_bits, counts.p, cur,
                                                 prev, use_prefix, use_suffix, d_stats.p, sms);
        timer.end(hc);
    }
    if (phase != 2) {
        int hc = timer.begin(COLIBRI_T_PRUNE);
        launches += launch_collect_slot_counts(s, cm->d_index.p, cm->index_cap, counts.p);
        timer.end(hc);
    }
    TRY(read_stats());
    if (phase != 2) cm->index_counts_dirty = false;
    ngram_upserts = h_stats.valid_windows + unigram_windows;
    if (phase == 1) {  // a shard: the caller sums the counters over the ranks and runs phase 2
        timer.end(h_total);
        CUDA_TRY(cudaStreamSynchronize(s));
        timer.resolve(m->ms, nullptr);
        m->counters[2] = launches;
        m->counters[3] = ngram_upserts;
        return 0;
    }

    // ---- threshold (prune(MINTOKENS, 0)) and the numbers of the progress line
    DevBuf<PatternMetaStats> d_pst;
    PatternMetaStats         pst;
    memset(&pst, 0, sizeof pst);
    pst.minn = pst.kept_minn = 0xFFFFFFFFu;
    TRY(d_pst.alloc(dev, 1));
    CUDA_TRY(cudaMemcpyAsync(d_pst.p, &pst, sizeof pst, cudaMemcpyHostToDevice, s));
    TRY(zero_stats());
    int hp = timer.begin(COLIBRI_T_PRUNE);
    launches += launch_constrained_stats(s, counts.p, cm->d_pn.p, np, t, flags.p, d_pst.p, d_stats.p, indexed);
    timer.end(hp);
    CUDA_TRY(cudaMemcpyAsync(&pst, d_pst.p, sizeof pst, cudaMemcpyDeviceToHost, s));
    TRY(read_stats());
    const uint64_t found = inplace ? np : h_stats.found;  // :1182 with prevsize = 0 (:970-971): an in-place rebuild "finds" every loaded pattern
    const uint64_t kept = h_stats.kept, kept_occ = h_stats.kept_occ;

    // ---- survivors -> the flat model (ordered by length for indexed models)
    int he = timer.begin(COLIBRI_T_EXPORT);
    if (indexed) TRY(kmap.alloc(dev, std::max<uint64_t>(np, 1)));
    TRY(compact_patterns(cm, flags.p, counts.p, indexed, false, indexed ? kmap.p : nullptr, m, &launches));
    timer.end(he);
    if (m->npatterns != kept) return set_err(COLIBRI_E_CUDA, "constrained training: %llu survivors compacted, %llu counted", (unsigned long long)m->npatterns, (unsigned long long)kept);
    if (indexed && kept) {
        int hi = timer.begin(COLIBRI_T_INDEX);
        DevBuf<uint64_t> tmp;
        TRY(tmp.alloc(dev, kept / 2048 + 4));
        launches += launch_exclusive_scan_u32_u64(s, m->d_counts.p, m->d_ref_off.p, kept, tmp.p);
        m->nrefs = kept_occ;
        TRY(m->d_ref_sentence.alloc(dev, std::max<uint64_t>(kept_occ, 1)));
        TRY(m->d_ref_token.alloc(dev, std::max<uint64_t>(kept_occ, 1)));
        uint64_t base = 0, rbase = 0;
        for (size_t k = 0; k < lengths.size(); ++k) {
            const int      n  = lengths[k];
            const uint64_t kn = pst.kept_n[n], on = pst.kept_occ_n[n];
            if (kn == 0) continue;
            Segment sg;
            sg.n     = n;
            sg.count = base + kn;  // keys are model-wide survivor indices: this bounds the radix passes
            TRY(build_refs(sg, match[k].p, kmap.p, false, npos, on));
            CUDA_TRY(cudaMemcpyAsync(m->d_ref_sentence.p + rbase, sg.ref_sentence.p, on * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(m->d_ref_token.p + rbase, sg.ref_token.p, on * sizeof(uint16_t), cudaMemcpyDeviceToDevice, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            base += kn;
            rbase += on;
        }
        timer.end(hi);
        if (base != kept || rbase != kept_occ) return set_err(COLIBRI_E_CUDA, "constrained training: occurrence lists cover %llu of %llu patterns", (unsigned long long)base, (unsigned long long)kept);
    }

    // ---- header numbers (see oracle_train_constrained for the line-by-line account)
    std::vector<PassStat> passes;
    int maxn = 0, minn = 999;
    if (inplace) {
        maxn            = cm->maxn;
        minn            = cm->minn;
        m->hasskipgrams = cm->hasskipgrams;
        m->totaltokens  = corpus_tokens;  // :889-891, :1047-1048
        m->totaltypes   = 0;
        if (t > 1)
            m->totaltypes = np;  // :1199-1201: size() of the model being rebuilt
        else if (o.MINLENGTH == 1)
            m->totaltypes = cm->meta.unigram_ngrams;  // totalwordtypesingroup(NGRAM, 1), :1202-1208
    } else {
        m->totaltypes  = cm->totaltypes;  // :892-895
        m->totaltokens = cm->totaltokens + corpus_tokens;
    }
    if (found) {  // :1184-1188 with n == 1
        maxn = std::max(maxn, 1);
        minn = std::min(minn, 1);
        passes.push_back({1, found, 0, found - kept});
    }
    if (t == 1 && kept) {  // postread, :1274-1277
        maxn = std::max(maxn, (int)pst.kept_maxn);
        minn = std::min(minn, (int)pst.kept_minn);
    }
    if (m->totaltypes == 0 && kept && !(inplace && t == 1 && o.MINLENGTH == 1)) {
        // types() (:1700-1704) computes totalwordtypesingroup(0, 0) on demand when totaltypes was never set, and write() stores that:
        // the distinct classes occurring in the patterns that are left
        const uint64_t   words = ((uint64_t)cm->meta.maxclass >> 5) + 1;
        DevBuf<uint32_t> bitmap;
        TRY(bitmap.alloc(dev, words));
        CUDA_TRY(cudaMemsetAsync(bitmap.p, 0, words * sizeof(uint32_t), s));
        TRY(zero_stats());
        launches += launch_token_bitmap(s, m->d_keys.p, m->d_off.p, kept, bitmap.p);
        launches += launch_popcount(s, bitmap.p, words, &d_stats.p->found);
        TRY(read_stats());
        m->totaltypes = h_stats.found;
    }
    m->maxn   = maxn;
    m->minn   = minn;
    m->passes = passes;
    timer.end(h_total);
    CUDA_TRY(cudaStreamSynchronize(s));
    std::map<int, double> level_ms;
    timer.resolve(m->ms, &level_ms);
    for (auto& kv : level_ms) {
        m->levels[kv.first].ms      = kv.second;
        m->levels[kv.first].windows = 0;
    }
    m->counters[2] = launches;
    m->counters[3] = ngram_upserts;
    m->counters[6] = corpus_tokens;
    m->counters[7] = g_pool[dev & 15].peak;
    return 0;
}

}  // namespace

// survivors of all levels -> the flat device-resident export (keys blob, offsets, counts) of the model
int colibri::export_segments(int dev, cudaStream_t s, std::vector<Segment>& segs, const uint32_t* tok, colibri_b200_model* m, uint64_t& launches) {
    uint64_t total = 0;
    for (auto& sg : segs) total += sg.count;
    m->npatterns = total;
    TRY(m->d_off.alloc(dev, total + 1));
    TRY(m->d_counts.alloc(dev, std::max<uint64_t>(total, 1)));
    if (total == 0) {
        CUDA_TRY(cudaMemsetAsync(m->d_off.p, 0, sizeof(uint64_t), s));
        m->keybytes = 0;
        TRY(m->d_keys.alloc(dev, 1));
        return 0;
    }
    DevBuf<uint32_t> pos, nm, lens;
    DevBuf<uint64_t> tmp;
    TRY(pos.alloc(dev, total));
    TRY(nm.alloc(dev, total));
    TRY(lens.alloc(dev, total));
    TRY(tmp.alloc(dev, total / 2048 + 4));
    uint64_t base = 0;
    for (auto& sg : segs) {
        CUDA_TRY(cudaMemcpyAsync(pos.p + base, sg.pos.p, sg.count * 4, cudaMemcpyDeviceToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(m->d_counts.p + base, sg.cnt.p, sg.count * 4, cudaMemcpyDeviceToDevice, s));
        if (sg.skip) {
            // raw gap masks first; packed into n | mask << 8 below
            CUDA_TRY(cudaMemcpyAsync(nm.p + base, sg.mask.p, sg.count * 4, cudaMemcpyDeviceToDevice, s));
        } else {
            launches += launch_fill_u32(s, nm.p + base, sg.count, (uint32_t)sg.n);
        }
        base += sg.count;
    }
    // skip segments hold raw masks in nm: turn them into n | mask << 8
    base = 0;
    for (auto& sg : segs) {
        if (sg.skip) {
            launches += launch_pack_nm(s, nm.p + base, sg.count, (uint32_t)sg.n);
        }
        base += sg.count;
    }
    TRY(m->d_len16.alloc(dev, total));
    launches += launch_export_lengths(s, tok, pos.p, nm.p, total, lens.p, m->d_len16.p);
    launches += launch_exclusive_scan_u32_u64(s, lens.p, m->d_off.p, total, tmp.p);
    uint64_t kb = 0;
    CUDA_TRY(cudaMemcpyAsync(&kb, m->d_off.p + total, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    m->keybytes = kb;
    TRY(m->d_keys.alloc(dev, std::max<uint64_t>(kb, 1)));
    launches += launch_export_write(s, tok, pos.p, nm.p, m->d_off.p, total, m->d_keys.p);
    if (m->model_type == COLIBRI_INDEXEDPATTERNMODEL) {
        uint64_t nrefs = 0;
        for (auto& sg : segs) nrefs += sg.nrefs;
        m->nrefs = nrefs;
        TRY(m->d_ref_sentence.alloc(dev, std::max<uint64_t>(nrefs, 1)));
        TRY(m->d_ref_token.alloc(dev, std::max<uint64_t>(nrefs, 1)));
        TRY(m->d_ref_off.alloc(dev, total + 1));
        uint64_t rbase = 0;
        for (auto& sg : segs) {
            if (sg.nrefs) {
                CUDA_TRY(cudaMemcpyAsync(m->d_ref_sentence.p + rbase, sg.ref_sentence.p, sg.nrefs * 4, cudaMemcpyDeviceToDevice, s));
                CUDA_TRY(cudaMemcpyAsync(m->d_ref_token.p + rbase, sg.ref_token.p, sg.nrefs * 2, cudaMemcpyDeviceToDevice, s));
            }
            rbase += sg.nrefs;
        }
        // a pattern's occurrence list has exactly `count` entries, and lists follow the pattern order
        launches += launch_exclusive_scan_u32_u64(s, m->d_counts.p, m->d_ref_off.p, total, tmp.p);
    }
    CUDA_TRY(cudaStreamSynchronize(s));  // the temporaries above go back to the pool when this returns
    segs.clear();
    return 0;
}

extern "C" int colibri_b200_train_corpus(colibri_b200_corpus* corpus, const colibri_b200_options* opt, colibri_b200_model** out) {
    if (!out) return set_err(COLIBRI_E_INVALID, "out is NULL");
    *out = nullptr;
    if (!corpus || !opt) return set_err(COLIBRI_E_INVALID, "NULL argument");
    colibri_b200_options o = *opt;
    TRY(check_options(o));
    if (o.device != corpus->device) return set_err(COLIBRI_E_INVALID, "options.device=%d but the corpus is staged on device %d", o.device, corpus->device);
    CUDA_TRY(cudaSetDevice(corpus->device));
    auto* m       = new colibri_b200_model();
    m->device     = corpus->device;
    m->model_type = o.model_type;
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete m;
        return set_err(COLIBRI_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    int rc;
    {
        Trainer tr;
        tr.c   = corpus;
        tr.o   = o;
        tr.m   = m;
        tr.s   = m->stream;
        tr.dev = corpus->device;
        rc     = tr.run();
        if (rc) cudaStreamSynchronize(m->stream);
    }
    if (rc) {
        colibri_b200_model_free(m);
        return rc;
    }
    *out = m;
    return 0;
}

// options of a constrained run: the same normalisation as train() (:883-888); what the single constrained scan never consults is not checked
static int check_constrained_options(colibri_b200_options& o) {
    if (o.MINTOKENS == -1) o.MINTOKENS = 2;
    if (o.MINTOKENS == 0) o.MINTOKENS = 1;
    if (o.MINTOKENS < 1) return set_err(COLIBRI_E_INVALID, "MINTOKENS=%d", o.MINTOKENS);
    if (o.model_type != COLIBRI_UNINDEXEDPATTERNMODEL && o.model_type != COLIBRI_INDEXEDPATTERNMODEL)
        return set_err(COLIBRI_E_UNSUPPORTED, "model type %d (only 10 = unindexed and 20 = indexed run on the device)", o.model_type);
    if (o.DOSKIPGRAMS || o.DOSKIPGRAMS_EXHAUSTIVE) return set_err(COLIBRI_E_UNSUPPORTED, "skipgrams under a constraint model are not on the device path yet");
    if (o.MINTOKENS_UNIGRAMS > o.MINTOKENS) return set_err(COLIBRI_E_UNSUPPORTED, "MINTOKENS_UNIGRAMS > MINTOKENS under a constraint model is not on the device path");
    if (o.DOPATTERNPERLINE) return set_err(COLIBRI_E_UNSUPPORTED, "DOPATTERNPERLINE is not on the device path");
    if (o.PRUNENONSUBSUMED || o.PRUNESUBSUMED) return set_err(COLIBRI_E_UNSUPPORTED, "PRUNE(NON)SUBSUMED is not on the device path");
    if (o.MAXLENGTH < 1 || o.MAXLENGTH > 255) return set_err(COLIBRI_E_UNSUPPORTED, "MAXLENGTH=%d (device path supports 1..255)", o.MAXLENGTH);
    if (o.MINLENGTH < 1) o.MINLENGTH = 1;
    return 0;
}

extern "C" int colibri_b200_train_constrained(colibri_b200_corpus* corpus, const colibri_b200_options* opt, colibri_b200_model* constrain, int inplace, colibri_b200_model** out) {
    if (!out) return set_err(COLIBRI_E_INVALID, "out is NULL");
    *out = nullptr;
    if (!corpus || !opt || !constrain) return set_err(COLIBRI_E_INVALID, "NULL argument");
    colibri_b200_options o = *opt;
    TRY(check_constrained_options(o));
    if (o.device != corpus->device) return set_err(COLIBRI_E_INVALID, "options.device=%d but the corpus is staged on device %d", o.device, corpus->device);
    if (constrain->device != corpus->device) return set_err(COLIBRI_E_INVALID, "the constraint model lives on device %d, the corpus on device %d", constrain->device, corpus->device);
    colibri_b200_model* m = nullptr;
    TRY(new_model(corpus->device, o.model_type, &m));
    int rc;
    {
        Trainer tr;
        tr.c   = corpus;
        tr.o   = o;
        tr.m   = m;
        tr.s   = m->stream;
        tr.dev = corpus->device;
        rc     = tr.run_constrained(constrain, inplace != 0);
        if (rc) cudaStreamSynchronize(m->stream);
    }
    if (rc) {
        colibri_b200_model_free(m);
        return rc;
    }
    *out = m;
    return 0;
}

// ---- constrained training over a sharded corpus: the constraint set is replicated, every rank counts its shard (no communication), the
// caller sums the counter arrays (one all-reduce), then every rank thresholds the same sums
extern "C" int colibri_b200_constrained_count(colibri_b200_corpus* shard, const colibri_b200_options* opt, colibri_b200_model* constrain, void* dev_counts, uint64_t* shard_tokens,
                                              uint64_t* kernel_launches) {
    if (!shard || !opt || !constrain || !dev_counts) return set_err(COLIBRI_E_INVALID, "NULL argument");
    colibri_b200_options o = *opt;
    TRY(check_constrained_options(o));
    if (o.device != shard->device || constrain->device != shard->device) return set_err(COLIBRI_E_INVALID, "corpus, constraint model and options must name the same device");
    colibri_b200_model* scratch = nullptr;
    TRY(new_model(shard->device, o.model_type, &scratch));
    int rc;
    {
        Trainer tr;
        tr.c   = shard;
        tr.o   = o;
        tr.m   = scratch;
        tr.s   = scratch->stream;
        tr.dev = shard->device;
        rc     = tr.run_constrained(constrain, false, 1, (uint32_t*)dev_counts, 0);
        if (rc) cudaStreamSynchronize(scratch->stream);
    }
    if (rc == 0) {
        if (shard_tokens) *shard_tokens = scratch->totaltokens;
        if (kernel_launches) *kernel_launches = scratch->counters[2];
    }
    colibri_b200_model_free(scratch);
    return rc;
}

extern "C" int colibri_b200_constrained_finish(const colibri_b200_options* opt, colibri_b200_model* constrain, void* dev_counts, uint64_t corpus_tokens, int inplace,
                                               colibri_b200_model** out) {
    if (!out) return set_err(COLIBRI_E_INVALID, "out is NULL");
    *out = nullptr;
    if (!opt || !constrain || !dev_counts) return set_err(COLIBRI_E_INVALID, "NULL argument");
    colibri_b200_options o = *opt;
    TRY(check_constrained_options(o));
    if (o.device != constrain->device) return set_err(COLIBRI_E_INVALID, "options.device=%d but the constraint model lives on device %d", o.device, constrain->device);
    colibri_b200_model* m = nullptr;
    TRY(new_model(constrain->device, o.model_type, &m));
    int rc;
    {
        Trainer tr;
        tr.c   = nullptr;
        tr.o   = o;
        tr.m   = m;
        tr.s   = m->stream;
        tr.dev = constrain->device;
        rc     = tr.run_constrained(constrain, inplace != 0, 2, (uint32_t*)dev_counts, corpus_tokens);
        if (rc) cudaStreamSynchronize(m->stream);
    }
    if (rc) {
        colibri_b200_model_free(m);
        return rc;
    }
    *out = m;
    return 0;
}

extern "C" int colibri_b200_train(const uint8_t* host_body, size_t nbytes, const colibri_b200_options* opt, colibri_b200_model** out) {
    if (!out) return set_err(COLIBRI_E_INVALID, "out is NULL");
    *out = nullptr;
    if (!opt) return set_err(COLIBRI_E_INVALID, "options is NULL");
    colibri_b200_options o = *opt;
    TRY(check_options(o));
    colibri_b200_corpus* c = nullptr;
    TRY(corpus_stage_async(host_body, nbytes, o.device, &c));  // the copy runs while the host sets the training up
    int rc = colibri_b200_train_corpus(c, opt, out);
    cudaStreamSynchronize(c->stream);  // host_body must not be in use after this call returns, whatever happened
    corpus_resolve_h2d(c);
    if (rc == 0) (*out)->ms[COLIBRI_T_H2D] = c->h2d_ms;
    colibri_b200_corpus_free(c);
    return rc;
}

extern "C" int colibri_b200_train_export(const uint8_t* host_body, size_t nbytes, const colibri_b200_options* opt, uint8_t* keys, uint64_t keys_cap, uint16_t* key_len,
                                         uint32_t* counts, uint64_t patterns_cap, colibri_b200_train_summary* summary) {
    if (!opt || !summary) return set_err(COLIBRI_E_INVALID, "NULL argument");
    memset(summary, 0, sizeof *summary);
    if ((keys_cap && !keys) || (patterns_cap && (!key_len || !counts))) return set_err(COLIBRI_E_INVALID, "NULL output buffer");
    colibri_b200_options o = *opt;
    TRY(check_options(o));
    if (o.model_type != COLIBRI_UNINDEXEDPATTERNMODEL) return set_err(COLIBRI_E_UNSUPPORTED, "colibri_b200_train_export streams unindexed models; use colibri_b200_train + colibri_b200_model_export for indexed ones");
    g_trace.begin();
    colibri_b200_corpus* c = nullptr;
    TRY(corpus_stage_async(host_body, nbytes, o.device, &c));
    g_trace.mark("stage");
    colibri_b200_model* m = nullptr;
    int rc = new_model(o.device, o.model_type, &m);
    if (rc == 0) {
        ExportSink sink;
        sink.dev = c->device;
        sink.keys = keys; sink.len16 = key_len; sink.counts = counts; sink.keys_cap = keys_cap; sink.pat_cap = patterns_cap;
        rc = g_streams.get(c->device, &sink.xs);
        if (rc == 0) {
            sink.h_scalars = g_pinned.get();
            if (!sink.h_scalars) rc = set_err(COLIBRI_E_CUDA, "streamed export set-up: no pinned memory");
        }
        g_trace.mark("setup");
        if (rc == 0) {
            Trainer tr;
            tr.c = c; tr.o = o; tr.m = m; tr.s = m->stream; tr.dev = c->device;
            tr.sink = &sink;
            rc = tr.run();
            if (rc) cudaStreamSynchronize(m->stream);
            g_trace.mark("run");
        }
        cudaStreamSynchronize(c->stream);  // host_body is the caller's again
        corpus_resolve_h2d(c);
        if (rc == 0) {
            summary->npatterns = sink.npat; summary->keybytes = sink.nbytes;
            summary->totaltokens = m->totaltokens; summary->totaltypes = m->totaltypes;
            summary->maxn = m->maxn; summary->minn = m->minn; summary->hasskipgrams = m->hasskipgrams;
            summary->npasses = (int32_t)m->passes.size();
            for (size_t i = 0; i < m->passes.size() && i < 32; ++i) {
                summary->passes[i][0] = m->passes[i].n; summary->passes[i][1] = m->passes[i].found;
                summary->passes[i][2] = m->passes[i].foundskip; summary->passes[i][3] = m->passes[i].pruned;
            }
            m->ms[COLIBRI_T_H2D] = c->h2d_ms;
            for (int i = 0; i < COLIBRI_T_NPHASES && i < 16; ++i) summary->ms[i] = m->ms[i];
            memcpy(summary->counters, m->counters, sizeof m->counters);
            if (sink.overflow)
                rc = set_err(COLIBRI_E_CAPACITY, "output buffers too small: %llu patterns / %llu key bytes needed", (unsigned long long)sink.npat, (unsigned long long)sink.nbytes);
        }
    }
    g_trace.mark("teardown-sink");
    if (m) colibri_b200_model_free(m);
    colibri_b200_corpus_free(c);
    g_trace.mark("free");
    g_trace.dump("train_export");
    return rc;
}

// ------------------------------------------------------------------------------------------------ reverse index (queries: relations.cu)
extern "C" int colibri_b200_rindex_build(colibri_b200_model* model, colibri_b200_corpus* corpus, int streamed, colibri_b200_rindex** out) {
    if (!out) return set_err(COLIBRI_E_INVALID, "out is NULL");
    *out = nullptr;
    if (!model || !corpus) return set_err(COLIBRI_E_INVALID, "NULL argument");
    if (model->device != corpus->device) return set_err(COLIBRI_E_INVALID, "model on device %d, corpus on device %d", model->device, corpus->device);
    colibri_b200_options o;
    colibri_b200_options_default(&o);
    o.MINTOKENS = 1;
    o.MAXLENGTH = std::max(model->maxn, 1);
    o.streamed  = streamed;
    o.device    = model->device;
    o.model_type = COLIBRI_INDEXEDPATTERNMODEL;
    colibri_b200_model* scratch = nullptr;  // the tokeniser reports into a model handle
    TRY(new_model(model->device, COLIBRI_INDEXEDPATTERNMODEL, &scratch));
    auto* r   = new colibri_b200_rindex();
    r->device = model->device;
    r->model  = model;
    r->stream = model->stream;
    int rc;
    {
        Trainer tr;
        tr.c = corpus; tr.o = o; tr.m = scratch; tr.s = model->stream; tr.dev = model->device;
        rc = tr.build_rindex(model, r);
        if (rc) cudaStreamSynchronize(model->stream);
    }
    colibri_b200_model_free(scratch);
    if (rc) {
        delete r;
        return rc;
    }
    *out = r;
    return 0;
}
extern "C" void colibri_b200_rindex_free(colibri_b200_rindex* r) {
    if (!r) return;
    cudaSetDevice(r->device);
    cudaStreamSynchronize(r->stream);
    colibri::rindex_cooc_forget(r);
    delete r;
}
extern "C" int colibri_b200_rindex_info(const colibri_b200_rindex* r, uint64_t out[4]) {
    if (!r || !out) return set_err(COLIBRI_E_INVALID, "NULL argument");
    out[0] = r->nsentences;
    out[1] = r->npos;
    out[2] = r->lengths.empty() ? 0 : (uint64_t)r->lengths.front();
    out[3] = r->lengths.empty() ? 0 : (uint64_t)r->lengths.back();
    return 0;
}

// ------------------------------------------------------------------------------------------------ export / lookup
extern "C" int colibri_b200_model_export_sizes(colibri_b200_model* m, uint64_t* npatterns, uint64_t* keybytes, uint64_t* nrefs) {
    if (!m) return set_err(COLIBRI_E_INVALID, "model is NULL");
    if (npatterns) *npatterns = m->npatterns;
    if (keybytes) *keybytes = m->keybytes;
    if (nrefs) *nrefs = m->nrefs;
    return 0;
}
extern "C" int colibri_b200_model_export(colibri_b200_model* m, uint8_t* keys, uint64_t* key_off, uint32_t* counts, uint32_t* ref_sentence, uint16_t* ref_token, uint64_t* ref_off) {
    if (!m || !key_off) return set_err(COLIBRI_E_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(m->device));
    if (m->keybytes && keys) CUDA_TRY(cudaMemcpyAsync(keys, m->d_keys.p, m->keybytes, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaMemcpyAsync(key_off, m->d_off.p, (m->npatterns + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, m->stream));
    if (m->npatterns && counts) CUDA_TRY(cudaMemcpyAsync(counts, m->d_counts.p, m->npatterns * sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream));
    if (m->model_type == COLIBRI_INDEXEDPATTERNMODEL && m->d_ref_off.p) {
        if (ref_off) CUDA_TRY(cudaMemcpyAsync(ref_off, m->d_ref_off.p, (m->npatterns + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, m->stream));
        if (ref_sentence && m->nrefs) CUDA_TRY(cudaMemcpyAsync(ref_sentence, m->d_ref_sentence.p, m->nrefs * sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream));
        if (ref_token && m->nrefs) CUDA_TRY(cudaMemcpyAsync(ref_token, m->d_ref_token.p, m->nrefs * sizeof(uint16_t), cudaMemcpyDeviceToHost, m->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return 0;
}
// compact flat form: key_len[npatterns] (bytes per key, u16) instead of the u64 offsets -- 2 instead of 8 bytes per pattern over PCIe
extern "C" int colibri_b200_model_export_compact(colibri_b200_model* m, uint8_t* keys, uint16_t* key_len, uint32_t* counts) {
    if (!m || !key_len) return set_err(COLIBRI_E_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(m->device));
    if (m->keybytes && keys) CUDA_TRY(cudaMemcpyAsync(keys, m->d_keys.p, m->keybytes, cudaMemcpyDeviceToHost, m->stream));
    if (m->npatterns) {
        CUDA_TRY(cudaMemcpyAsync(key_len, m->d_len16.p, m->npatterns * sizeof(uint16_t), cudaMemcpyDeviceToHost, m->stream));
        if (counts) CUDA_TRY(cudaMemcpyAsync(counts, m->d_counts.p, m->npatterns * sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return 0;
}
static int ensure_host(colibri_b200_model* m) {
    if (m->host_ready) return 0;
    m->h_keys.resize(m->keybytes + 1);
    m->h_off.resize(m->npatterns + 1);
    m->h_counts.resize(m->npatterns + 1);
    const bool indexed = m->model_type == COLIBRI_INDEXEDPATTERNMODEL;
    if (indexed) {
        m->h_ref_sentence.resize(m->nrefs + 1);
        m->h_ref_token.resize(m->nrefs + 1);
        m->h_ref_off.resize(m->npatterns + 1);
    }
    TRY(colibri_b200_model_export(m, m->h_keys.data(), m->h_off.data(), m->h_counts.data(), indexed ? m->h_ref_sentence.data() : nullptr, indexed ? m->h_ref_token.data() : nullptr,
                                  indexed ? m->h_ref_off.data() : nullptr));
    m->host_ready = true;
    return 0;
}
extern "C" int colibri_b200_model_write(colibri_b200_model* m, uint8_t* buf, size_t cap, size_t* nbytes) {
    // layout: include/patternmodel.h:1609-1624 + include/patternstore.h:534-542 + src/pattern.cpp:268-277 + include/datatypes.h:216-221
    if (!m || !nbytes) return set_err(COLIBRI_E_INVALID, "NULL argument");
    const bool indexed = m->model_type == COLIBRI_INDEXEDPATTERNMODEL;  // value = u32 count + count x (u32 sentence, u16 token), include/datatypes.h:263-270, :55-58
    size_t need = 3 + 24 + m->keybytes + m->npatterns * 5 + (indexed ? m->nrefs * 6 : 0);
    *nbytes     = need;
    if (!buf) return 0;
    if (cap < need) return set_err(COLIBRI_E_INVALID, "buffer too small: need %zu bytes", need);
    TRY(ensure_host(m));
    size_t w = 0;
    buf[w++] = 0;
    buf[w++] = (uint8_t)m->model_type;
    buf[w++] = 2;
    memcpy(buf + w, &m->totaltokens, 8); w += 8;
    memcpy(buf + w, &m->totaltypes, 8);  w += 8;
    memcpy(buf + w, &m->npatterns, 8);   w += 8;
    for (uint64_t i = 0; i < m->npatterns; ++i) {
        size_t l = (size_t)(m->h_off[i + 1] - m->h_off[i]);
        memcpy(buf + w, m->h_keys.data() + m->h_off[i], l);
        w += l;
        buf[w++] = 0;
        memcpy(buf + w, &m->h_counts[i], 4);
        w += 4;
        if (indexed) {
            for (uint64_t j = m->h_ref_off[i]; j < m->h_ref_off[i + 1]; ++j) {
                memcpy(buf + w, &m->h_ref_sentence[j], 4);
                memcpy(buf + w + 4, &m->h_ref_token[j], 2);
                w += 6;
            }
        }
    }
    return 0;
}
// ------------------------------------------------------------------------------------------------ parity helpers
extern "C" int colibri_b200_hash64_batch(const uint8_t* keys, const uint64_t* key_off, uint64_t n, uint64_t* out, int device) {
    if (colibri_b200_device_count() <= 0) return set_err(COLIBRI_E_CUDA, "no CUDA device available: the B200 path has no CPU fallback");
    if (!key_off || !out) return set_err(COLIBRI_E_INVALID, "NULL argument");
    if (n == 0) return 0;
    CUDA_TRY(cudaSetDevice(device));
    uint64_t kb = key_off[n];
    DevBuf<uint8_t>  dk;
    DevBuf<uint64_t> doff, dout;
    TRY(dk.alloc(device, kb + 1));
    TRY(doff.alloc(device, n + 1));
    TRY(dout.alloc(device, n));
    if (kb) CUDA_TRY(cudaMemcpy(dk.p, keys, kb, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(doff.p, key_off, (n + 1) * 8, cudaMemcpyHostToDevice));
    launch_hash64_batch(nullptr, dk.p, doff.p, n, dout.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, dout.p, n * 8, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int colibri_b200_corpus_tokens(colibri_b200_corpus* c, uint32_t* out, uint64_t cap, uint64_t* npositions) {
    if (!c || !npositions) return set_err(COLIBRI_E_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    uint8_t tail[16];
    memset(tail, 0x80, sizeof tail);
    CUDA_TRY(cudaMemcpyAsync(c->body() + c->nbytes, tail, sizeof tail, cudaMemcpyHostToDevice, s));
    const uint32_t nblocks = (uint32_t)(c->padded(c->nbytes) / kTokTile);
    DevBuf<uint32_t>    blk;
    DevBuf<DeviceStats> st;
    TRY(blk.alloc(c->device, nblocks + 1));
    TRY(st.alloc(c->device, 1));
    CUDA_TRY(cudaMemsetAsync(st.p, 0, sizeof(DeviceStats), s));
    if (nblocks) {
        launch_tokenise_count(s, c->body(), c->nbytes, blk.p, nblocks);
        launch_scan_block_counts(s, blk.p, nblocks, &st.p->cursor);
    }
    DeviceStats h;
    CUDA_TRY(cudaMemcpyAsync(&h, st.p, sizeof h, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    *npositions = h.cursor;
    if (!out) return 0;
    if (cap < h.cursor) return set_err(COLIBRI_E_INVALID, "buffer too small: %llu positions", (unsigned long long)h.cursor);
    DevBuf<uint32_t> tok;
    TRY(tok.alloc(c->device, h.cursor + 8));
    if (nblocks) launch_tokenise_write(s, c->body(), c->nbytes, blk.p, nblocks, tok.p, st.p);
    CUDA_TRY(cudaMemcpyAsync(out, tok.p, h.cursor * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(&h, st.p, sizeof h, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (h.errflags & (kErrTokenTooLong | kErrNonCanonical)) return set_err(COLIBRI_E_FORMAT, "malformed class encoding in corpus (flags %u)", h.errflags);
    return 0;
}

// ------------------------------------------------------------------------------------------------ synthetic corpus
extern "C" int colibri_b200_synth_corpus(const colibri_b200_synth_params* p, int device, colibri_b200_corpus** out) {
    if (!p || !out) return set_err(COLIBRI_E_INVALID, "NULL argument");
    *out = nullptr;
    if (colibri_b200_device_count() <= 0) return set_err(COLIBRI_E_CUDA, "no CUDA device available: the B200 path has no CPU fallback");
    if (p->vocab == 0 || p->mean_sentence == 0 || p->ntokens == 0) return set_err(COLIBRI_E_INVALID, "vocab, mean_sentence and ntokens must be positive");
    CUDA_TRY(cudaSetDevice(device));
    std::vector<uint64_t> cdf(p->vocab);
    uint64_t acc = 0;
    for (uint32_t r = 0; r < p->vocab; ++r) {  // integer Zipf table, identical to oracle_synth_cdf
        acc += (1ULL << 40) / (uint64_t)(r + 1);
        cdf[r] = acc;
    }
    DevBuf<uint64_t> dcdf, off, tmp;
    DevBuf<uint32_t> lens;
    TRY(dcdf.alloc(device, p->vocab));
    TRY(lens.alloc(device, p->ntokens));
    TRY(off.alloc(device, p->ntokens + 1));
    TRY(tmp.alloc(device, p->ntokens / 2048 + 4));
    CUDA_TRY(cudaMemcpy(dcdf.p, cdf.data(), cdf.size() * 8, cudaMemcpyHostToDevice));
    launch_synth_lengths(nullptr, p->seed, p->ntokens, p->first_token, p->vocab, p->mean_sentence, p->phrase_permille, p->nphrases, dcdf.p, lens.p);
    launch_exclusive_scan_u32_u64(nullptr, lens.p, off.p, p->ntokens, tmp.p);
    uint64_t nbytes = 0;
    CUDA_TRY(cudaMemcpy(&nbytes, off.p + p->ntokens, 8, cudaMemcpyDeviceToHost));
    auto* c = new colibri_b200_corpus();
    int   rc = corpus_alloc(c, device, nbytes);
    if (rc == 0) {
        cudaStreamSynchronize(c->stream);
        launch_synth_write(nullptr, p->seed, p->ntokens, p->first_token, p->vocab, p->mean_sentence, p->phrase_permille, p->nphrases, dcdf.p, off.p, c->body());
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = set_err(COLIBRI_E_CUDA, "synthetic corpus kernel failed: %s", cudaGetErrorString(e));
    }
    if (rc == 0) rc = corpus_finish(c);
    if (rc) {
        colibri_b200_corpus_free(c);
        return rc;
    }
    *out = c;
    return 0;
}
