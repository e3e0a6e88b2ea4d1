// model_io.cu -- models that do not come out of train(): uploaded pattern sets, model files read with the options as
// filters, and queries against a device-resident model.
//
// Reference: PatternModel(filename, options, constrainmodel, corpus) / load() include/patternmodel.h:700-726, :781-861,
// PatternMapStore::read include/patternstore.h:555-619 (filters), postread :572-588, has() :751-756,
// occurrencecount() :1653-1669.  The file format is a sequence of variable-length records (key bytes, 0x00, u32 count
// [, count x 6-byte references]) whose boundaries only a sequential scan can find, so the scan runs on the host; what it
// yields (flat blob + offsets + counts + references) goes to HBM, and everything after that -- pattern shapes, the
// filters, the membership test against a constraint model, compaction, the hash index, lookups -- runs on the device.
#include "engine_common.h"

using namespace colibri;

int colibri::new_model(int device, int model_type, colibri_b200_model** out) {
    *out = nullptr;
    if (colibri_b200_device_count() <= 0) return set_err(COLIBRI_E_CUDA, "no CUDA device available: the B200 path has no CPU fallback");
    if (model_type != COLIBRI_UNINDEXEDPATTERNMODEL && model_type != COLIBRI_INDEXEDPATTERNMODEL)
        return set_err(COLIBRI_E_UNSUPPORTED, "model type %d (only 10 = unindexed and 20 = indexed live on the device)", model_type);
    CUDA_TRY(cudaSetDevice(device));
    auto* m       = new colibri_b200_model();
    m->device     = device;
    m->model_type = model_type;
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete m;
        return set_err(COLIBRI_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    *out = m;
    return 0;
}

int colibri::ensure_meta(colibri_b200_model* m, uint64_t* launches) {
    if (m->meta_ready) return 0;
    CUDA_TRY(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    memset(&m->meta, 0, sizeof m->meta);
    m->meta.minn = m->meta.kept_minn = 0xFFFFFFFFu;
    uint64_t l = 0;
    if (m->npatterns) {
        DevBuf<PatternMetaStats> st;
        TRY(st.alloc(m->device, 1));
        TRY(m->d_pn.alloc(m->device, m->npatterns));
        TRY(m->d_pcat.alloc(m->device, m->npatterns));
        CUDA_TRY(cudaMemcpyAsync(st.p, &m->meta, sizeof m->meta, cudaMemcpyHostToDevice, s));
        l += launch_pattern_meta(s, m->d_keys.p, m->d_off.p, m->npatterns, m->d_pn.p, m->d_pcat.p, st.p);
        CUDA_TRY(cudaMemcpyAsync(&m->meta, st.p, sizeof m->meta, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
    }
    if (launches) *launches += l;
    m->meta_ready = true;
    return 0;
}

int colibri::ensure_index(colibri_b200_model* m, uint64_t* launches) {
    if (m->index_ready && !m->index_counts_dirty) return 0;
    m->index_ready        = false;
    m->closure_ready      = false;
    m->index_counts_dirty = false;
    TRY(ensure_meta(m, launches));
    if (m->meta.malformed)
        return set_err(COLIBRI_E_UNSUPPORTED, "%u pattern(s) are empty, longer than %u bytes or not a well-formed class sequence: not indexable on the device", m->meta.malformed,
                       kMaxIndexedKeyBytes);
    if (m->npatterns >= 0xFFFFFFF0ull) return set_err(COLIBRI_E_CAPACITY, "%llu patterns; the device index is 32 bit", (unsigned long long)m->npatterns);
    cudaStream_t s = m->stream;
    uint64_t cap = 1024;
    while (cap < 2 * m->npatterns) cap <<= 1;
    uint64_t pbits = 1ull << 15;
    while (pbits < 16 * m->npatterns) pbits <<= 1;
    TRY(m->d_index.alloc(m->device, cap));
    TRY(m->d_presence.alloc(m->device, pbits / 32));
    CUDA_TRY(cudaMemsetAsync(m->d_index.p, 0, cap * sizeof(PatSlot), s));
    CUDA_TRY(cudaMemsetAsync(m->d_presence.p, 0, pbits / 8, s));
    uint64_t l = 0;
    if (m->npatterns) {
        DevBuf<PatternMetaStats> st;
        TRY(st.alloc(m->device, 1));
        CUDA_TRY(cudaMemsetAsync(st.p, 0, sizeof(PatternMetaStats), s));
        l += launch_index_build(s, m->d_keys.p, m->d_off.p, m->npatterns, m->d_index.p, cap, m->d_presence.p, pbits, st.p);
        PatternMetaStats h;
        CUDA_TRY(cudaMemcpyAsync(&h, st.p, sizeof h, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (h.duplicates) return set_err(COLIBRI_E_INVALID, "the pattern set holds %u duplicate pattern(s)", h.duplicates);
        if (h.malformed) return set_err(COLIBRI_E_CAPACITY, "pattern index overflow");
    } else {
        CUDA_TRY(cudaStreamSynchronize(s));
    }
    if (launches) *launches += l;
    m->index_cap     = cap;
    m->presence_bits = pbits;
    m->index_ready   = true;
    return 0;
}

int colibri::ensure_closure(colibri_b200_model* m, uint64_t* launches) {
    if (m->closure_ready) return 0;
    TRY(ensure_index(m, launches));
    cudaStream_t s = m->stream;
    uint64_t     l = 0;
    memset(m->prefix_open, 0, sizeof m->prefix_open);
    memset(m->suffix_open, 0, sizeof m->suffix_open);
    m->uni_classes = 0;
    if (m->npatterns) {
        DevBuf<unsigned long long> open;
        TRY(open.alloc(m->device, 512));
        CUDA_TRY(cudaMemsetAsync(open.p, 0, 512 * sizeof(unsigned long long), s));
        l += launch_closure_check(s, m->d_keys.p, m->d_off.p, m->d_pn.p, m->npatterns, m->d_index.p, m->index_cap, m->d_presence.p, m->presence_bits, open.p, open.p + 256);
        CUDA_TRY(cudaMemcpyAsync(m->prefix_open, open.p, 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(m->suffix_open, open.p + 256, 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        if (m->meta.nhist[1] && m->meta.maxclass < (1u << 28)) {
            m->uni_classes = m->meta.maxclass + 1;
            TRY(m->d_uni.alloc(m->device, m->uni_classes));
            CUDA_TRY(cudaMemsetAsync(m->d_uni.p, 0, (size_t)m->uni_classes * sizeof(uint32_t), s));
            l += launch_unigram_table(s, m->d_keys.p, m->d_off.p, m->d_pn.p, m->npatterns, m->d_uni.p, m->uni_classes);
        }
        CUDA_TRY(cudaStreamSynchronize(s));
    }
    if (launches) *launches += l;
    m->closure_ready = true;
    return 0;
}

int colibri::compact_patterns(const colibri_b200_model* src, const uint32_t* d_flags, const uint32_t* d_counts, bool order_by_length, bool copy_refs, uint32_t* d_kmap,
                              colibri_b200_model* dst, uint64_t* launches) {
    const int      dev = dst->device;
    cudaStream_t   s   = dst->stream;
    const uint64_t np  = src->npatterns;
    uint64_t       l   = 0;
    DevBuf<uint64_t> newpos, tmp;
    TRY(newpos.alloc(dev, np + 1));
    TRY(tmp.alloc(dev, np / 2048 + 4));
    uint64_t K = 0;
    if (np) {
        l += launch_exclusive_scan_u32_u64(s, d_flags, newpos.p, np, tmp.p);
        CUDA_TRY(cudaMemcpyAsync(&K, newpos.p + np, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
    }
    if (d_kmap && np) CUDA_TRY(cudaMemsetAsync(d_kmap, 0, np * sizeof(uint32_t), s));
    dst->npatterns = K;
    dst->keybytes  = 0;
    dst->nrefs     = 0;
    TRY(dst->d_off.alloc(dev, K + 1));
    TRY(dst->d_counts.alloc(dev, std::max<uint64_t>(K, 1)));
    TRY(dst->d_len16.alloc(dev, std::max<uint64_t>(K, 1)));
    const bool indexed = dst->model_type == COLIBRI_INDEXEDPATTERNMODEL;
    if (K == 0) {
        CUDA_TRY(cudaMemsetAsync(dst->d_off.p, 0, sizeof(uint64_t), s));
        TRY(dst->d_keys.alloc(dev, 1));
        if (indexed) {
            TRY(dst->d_ref_off.alloc(dev, 1));
            TRY(dst->d_ref_sentence.alloc(dev, 1));
            TRY(dst->d_ref_token.alloc(dev, 1));
            CUDA_TRY(cudaMemsetAsync(dst->d_ref_off.p, 0, sizeof(uint64_t), s));
        }
        CUDA_TRY(cudaStreamSynchronize(s));
        if (launches) *launches += l;
        return 0;
    }
    DevBuf<uint32_t> sel_idx, sel_n, sel_idx2, sel_n2, lens;
    TRY(sel_idx.alloc(dev, K));
    TRY(sel_n.alloc(dev, K));
    TRY(lens.alloc(dev, K));
    l += launch_select_scatter(s, d_flags, newpos.p, src->d_pn.p, np, sel_idx.p, sel_n.p);
    uint32_t* order = sel_idx.p;
    DevBuf<uint32_t> hist;
    DevBuf<uint64_t> hist_off, stmp;
    if (order_by_length) {  // stable LSD radix sort on the token count (index.cu): survivors stay in index order inside a length
        const uint64_t nsort = (K + 4095) / 4096;
        TRY(sel_idx2.alloc(dev, K));
        TRY(sel_n2.alloc(dev, K));
        TRY(hist.alloc(dev, 256 * nsort));
        TRY(hist_off.alloc(dev, 256 * nsort + 1));
        TRY(stmp.alloc(dev, 256 * nsort / 2048 + 4));
        uint32_t *kin = sel_n.p, *vin = sel_idx.p, *kout = sel_n2.p, *vout = sel_idx2.p;
        for (int shift = 0; shift < 16 && (src->meta.maxn >> shift) != 0; shift += 8) {
            l += launch_radix_pass(s, kin, vin, K, shift, hist.p, hist_off.p, stmp.p, kout, vout);
            std::swap(kin, kout);
            std::swap(vin, vout);
        }
        order = vin;
    }
    l += launch_gather_meta(s, order, K, src->d_off.p, d_counts, d_kmap, lens.p, dst->d_len16.p, dst->d_counts.p);
    l += launch_exclusive_scan_u32_u64(s, lens.p, dst->d_off.p, K, tmp.p);
    uint64_t kb = 0;
    CUDA_TRY(cudaMemcpyAsync(&kb, dst->d_off.p + K, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    dst->keybytes = kb;
    TRY(dst->d_keys.alloc(dev, std::max<uint64_t>(kb, 1)));
    l += launch_gather_keys(s, order, K, src->d_keys.p, src->d_off.p, dst->d_off.p, dst->d_keys.p);
    if (indexed) {
        TRY(dst->d_ref_off.alloc(dev, K + 1));
        if (copy_refs && src->d_ref_off.p && d_counts) {
            // a pattern's occurrence list has exactly `count` entries
            l += launch_exclusive_scan_u32_u64(s, dst->d_counts.p, dst->d_ref_off.p, K, tmp.p);
            uint64_t nr = 0;
            CUDA_TRY(cudaMemcpyAsync(&nr, dst->d_ref_off.p + K, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            dst->nrefs = nr;
            TRY(dst->d_ref_sentence.alloc(dev, std::max<uint64_t>(nr, 1)));
            TRY(dst->d_ref_token.alloc(dev, std::max<uint64_t>(nr, 1)));
            l += launch_gather_refs(s, order, K, src->d_ref_off.p, src->d_ref_sentence.p, src->d_ref_token.p, dst->d_ref_off.p, dst->d_ref_sentence.p, dst->d_ref_token.p);
        } else {
            CUDA_TRY(cudaMemsetAsync(dst->d_ref_off.p, 0, (K + 1) * sizeof(uint64_t), s));
            TRY(dst->d_ref_sentence.alloc(dev, 1));
            TRY(dst->d_ref_token.alloc(dev, 1));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(s));  // the temporaries go back to the pool when this returns
    if (launches) *launches += l;
    return 0;
}

// ------------------------------------------------------------------------------------------------ upload
extern "C" int colibri_b200_model_from_flat(const uint8_t* keys, const uint64_t* key_off, const uint32_t* counts, uint64_t npatterns, const uint32_t* ref_sentence,
                                            const uint16_t* ref_token, const uint64_t* ref_off, uint64_t totaltokens, uint64_t totaltypes, int model_type, int device,
                                            colibri_b200_model** out) {
    if (!out) return set_err(COLIBRI_E_INVALID, "out is NULL");
    *out = nullptr;
    if (!key_off || (npatterns && !keys && key_off[npatterns])) return set_err(COLIBRI_E_INVALID, "NULL argument");
    for (uint64_t i = 0; i < npatterns; ++i)
        if (key_off[i + 1] < key_off[i]) return set_err(COLIBRI_E_INVALID, "key_off is not ascending at %llu", (unsigned long long)i);
    colibri_b200_model* m = nullptr;
    TRY(new_model(device, model_type, &m));
    auto bail = [&](int rc) {
        cudaStreamSynchronize(m->stream);
        colibri_b200_model_free(m);
        return rc;
    };
#define MTRY(expr)                       \
    do {                                 \
        int rc2__ = (expr);              \
        if (rc2__ != 0) return bail(rc2__); \
    } while (0)
#define MCUDA(expr)                                                                                                                        \
    do {                                                                                                                                   \
        cudaError_t e2__ = (expr);                                                                                                         \
        if (e2__ != cudaSuccess) return bail(set_err(COLIBRI_E_CUDA, "CUDA error %s at %s:%d", cudaGetErrorName(e2__), __FILE__, __LINE__)); \
    } while (0)
    cudaStream_t   s  = m->stream;
    const uint64_t kb = npatterns ? key_off[npatterns] : 0;
    m->npatterns   = npatterns;
    m->keybytes    = kb;
    m->totaltokens = totaltokens;
    m->totaltypes  = totaltypes;
    MTRY(m->d_keys.alloc(device, std::max<uint64_t>(kb, 1)));
    MTRY(m->d_off.alloc(device, npatterns + 1));
    MTRY(m->d_counts.alloc(device, std::max<uint64_t>(npatterns, 1)));
    MTRY(m->d_len16.alloc(device, std::max<uint64_t>(npatterns, 1)));
    if (kb) MCUDA(cudaMemcpyAsync(m->d_keys.p, keys, kb, cudaMemcpyHostToDevice, s));
    if (npatterns) {
        MCUDA(cudaMemcpyAsync(m->d_off.p, key_off, (npatterns + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
        if (counts)
            MCUDA(cudaMemcpyAsync(m->d_counts.p, counts, npatterns * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        else
            MCUDA(cudaMemsetAsync(m->d_counts.p, 0, npatterns * sizeof(uint32_t), s));
        std::vector<uint16_t> l16(npatterns);
        for (uint64_t i = 0; i < npatterns; ++i) l16[i] = (uint16_t)std::min<uint64_t>(key_off[i + 1] - key_off[i], 65535);
        MCUDA(cudaMemcpyAsync(m->d_len16.p, l16.data(), npatterns * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
        MCUDA(cudaStreamSynchronize(s));
    } else {
        MCUDA(cudaMemsetAsync(m->d_off.p, 0, sizeof(uint64_t), s));
    }
    if (model_type == COLIBRI_INDEXEDPATTERNMODEL) {
        const bool     have = ref_off && ref_sentence && ref_token && counts;
        const uint64_t nr   = (have && npatterns) ? ref_off[npatterns] : 0;
        m->nrefs = nr;
        MTRY(m->d_ref_off.alloc(device, npatterns + 1));
        MTRY(m->d_ref_sentence.alloc(device, std::max<uint64_t>(nr, 1)));
        MTRY(m->d_ref_token.alloc(device, std::max<uint64_t>(nr, 1)));
        if (have && npatterns) {
            for (uint64_t i = 0; i < npatterns; ++i)
                if (ref_off[i + 1] - ref_off[i] != counts[i]) return bail(set_err(COLIBRI_E_INVALID, "pattern %llu: %llu references for a count of %u", (unsigned long long)i,
                                                                                   (unsigned long long)(ref_off[i + 1] - ref_off[i]), counts[i]));
            MCUDA(cudaMemcpyAsync(m->d_ref_off.p, ref_off, (npatterns + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
            if (nr) {
                MCUDA(cudaMemcpyAsync(m->d_ref_sentence.p, ref_sentence, nr * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
                MCUDA(cudaMemcpyAsync(m->d_ref_token.p, ref_token, nr * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
            }
        } else {
            // an indexed model without occurrence lists has no counts either (IndexedData::count() is the list length)
            MCUDA(cudaMemsetAsync(m->d_ref_off.p, 0, (npatterns + 1) * sizeof(uint64_t), s));
            if (npatterns) MCUDA(cudaMemsetAsync(m->d_counts.p, 0, npatterns * sizeof(uint32_t), s));
        }
    }
    MCUDA(cudaStreamSynchronize(s));
    // postread (include/patternmodel.h:572-588)
    uint64_t launches = 0;
    MTRY(ensure_meta(m, &launches));
    if (npatterns) {
        m->maxn         = (int)m->meta.maxn;
        m->minn         = (int)m->meta.minn;
        m->hasskipgrams = m->meta.hasskip ? 1 : 0;
    }
    m->counters[2] = launches;
    *out = m;
    return 0;
}

// ------------------------------------------------------------------------------------------------ load
namespace {
struct ParsedFile {
    int                   type = 0;
    uint64_t              tokens = 0, types = 0;
    std::vector<uint8_t>  keys;
    std::vector<uint64_t> off;
    std::vector<uint32_t> counts;
    std::vector<uint32_t> rs;
    std::vector<uint16_t> rt;
    std::vector<uint64_t> roff;
};
// sequential scan of the record stream (include/patternmodel.h:781-816, include/patternstore.h:559-584, src/pattern.cpp:483-587,
// include/datatypes.h:209-221, :263-270, :55-58)
int parse_modelfile(const uint8_t* d, size_t n, bool want_refs, ParsedFile& f) {
    if (n < 27 || d[0] != 0) return set_err(COLIBRI_E_FORMAT, "File is not a colibri model file (or a very old one)");
    f.type = d[1];
    if (f.type != COLIBRI_UNINDEXEDPATTERNMODEL && f.type != COLIBRI_INDEXEDPATTERNMODEL)
        return set_err(COLIBRI_E_UNSUPPORTED, "model file of type %d: only unindexed (10) and indexed (20) pattern models are read on the B200 path", f.type);
    if (d[2] != 2) return set_err(COLIBRI_E_UNSUPPORTED, "model file version %d: only version 2 (class encoding v2) is read", (int)d[2]);
    uint64_t np;
    memcpy(&f.tokens, d + 3, 8);
    memcpy(&f.types, d + 11, 8);
    memcpy(&np, d + 19, 8);
    if (np > n) return set_err(COLIBRI_E_FORMAT, "model file: pattern count %llu exceeds the file size", (unsigned long long)np);
    f.off.reserve(np + 1);
    f.counts.reserve(np);
    f.keys.reserve(n / 2);
    const bool indexed = f.type == COLIBRI_INDEXEDPATTERNMODEL;
    size_t pos = 27;
    f.off.push_back(0);
    if (indexed && want_refs) f.roff.push_back(0);
    for (uint64_t i = 0; i < np; ++i) {
        const size_t start = pos;
        bool prevhigh = false;
        for (;; ++pos) {  // a key runs to the first 0x00 that does not follow a continuation byte
            if (pos >= n) return set_err(COLIBRI_E_FORMAT, "model file: truncated pattern %llu", (unsigned long long)i);
            if (!prevhigh && d[pos] == 0) break;
            prevhigh = d[pos] >= 128;
        }
        f.keys.insert(f.keys.end(), d + start, d + pos);
        f.off.push_back(f.keys.size());
        ++pos;
        if (pos + 4 > n) return set_err(COLIBRI_E_FORMAT, "model file: truncated value of pattern %llu", (unsigned long long)i);
        uint32_t c;
        memcpy(&c, d + pos, 4);
        pos += 4;
        f.counts.push_back(c);
        if (indexed) {
            if (pos + (size_t)c * 6 > n) return set_err(COLIBRI_E_FORMAT, "model file: truncated index of pattern %llu", (unsigned long long)i);
            if (want_refs) {
                for (uint32_t j = 0; j < c; ++j) {
                    uint32_t sent;
                    uint16_t tokn;
                    memcpy(&sent, d + pos + (size_t)j * 6, 4);
                    memcpy(&tokn, d + pos + (size_t)j * 6 + 4, 2);
                    f.rs.push_back(sent);
                    f.rt.push_back(tokn);
                }
                f.roff.push_back(f.rs.size());
            }
            pos += (size_t)c * 6;
        }
    }
    return 0;
}
}  // namespace

extern "C" int colibri_b200_model_load(const uint8_t* file, size_t nbytes, const colibri_b200_options* opt, colibri_b200_model* constrain, colibri_b200_model** out) {
    if (!out) return set_err(COLIBRI_E_INVALID, "out is NULL");
    *out = nullptr;
    if (!file || !opt) return set_err(COLIBRI_E_INVALID, "NULL argument");
    if (colibri_b200_device_count() <= 0) return set_err(COLIBRI_E_CUDA, "no CUDA device available: the B200 path has no CPU fallback");
    if (constrain && constrain->device != opt->device) return set_err(COLIBRI_E_INVALID, "the constraint model lives on device %d, options.device=%d", constrain->device, opt->device);
    const int  as_type  = opt->model_type;
    const bool as_index = as_type == COLIBRI_INDEXEDPATTERNMODEL;
    ParsedFile f;
    const bool keep_refs = as_index && !opt->DORESET;
    TRY(parse_modelfile(file, nbytes, keep_refs, f));
    const uint64_t np = f.counts.size();
    // everything read, as one device-resident set (type of the file: counts as stored, references if they will be kept)
    colibri_b200_model* all = nullptr;
    const bool file_refs = keep_refs && f.type == COLIBRI_INDEXEDPATTERNMODEL;
    TRY(colibri_b200_model_from_flat(f.keys.data(), f.off.data(), f.counts.data(), np, file_refs ? f.rs.data() : nullptr, file_refs ? f.rt.data() : nullptr,
                                     file_refs ? f.roff.data() : nullptr, f.tokens, f.types, file_refs ? COLIBRI_INDEXEDPATTERNMODEL : COLIBRI_UNINDEXEDPATTERNMODEL, opt->device, &all));
    colibri_b200_model* m = nullptr;
    int rc = new_model(opt->device, as_type, &m);
    uint64_t launches = all->counters[2];
    if (rc == 0) {
        const int      dev = opt->device;
        cudaStream_t   s   = m->stream;
        DevBuf<uint32_t>         flags, cidx;
        DevBuf<PatternMetaStats> st;
        PatternMetaStats         h;
        memset(&h, 0, sizeof h);
        h.minn = h.kept_minn = 0xFFFFFFFFu;
        auto body = [&]() -> int {
            TRY(flags.alloc(dev, np + 1));
            TRY(st.alloc(dev, 1));
            CUDA_TRY(cudaMemcpyAsync(st.p, &h, sizeof h, cudaMemcpyHostToDevice, s));
            if (constrain && np) {  // constrainstore->has(p), include/patternstore.h:586
                TRY(ensure_index(constrain, &launches));
                TRY(cidx.alloc(dev, np));
                launches += launch_index_lookup(s, all->d_keys.p, all->d_off.p, np, constrain->d_keys.p, constrain->d_off.p, constrain->d_index.p, constrain->index_cap, constrain->d_presence.p,
                                                constrain->presence_bits, cidx.p);
            }
            const int64_t mintokens = opt->MINTOKENS == -1 ? 0 : opt->MINTOKENS;  // include/patternstore.h:565-566
            launches += launch_load_filter(s, all->d_pn.p, all->d_pcat.p, all->d_counts.p, (constrain && np) ? cidx.p : nullptr, np, (uint32_t)std::max<int64_t>(mintokens, 0),
                                           (uint32_t)std::max(opt->MINLENGTH, 0), (uint32_t)std::max(opt->MAXLENGTH, 0), !opt->DOREMOVENGRAMS, !opt->DOREMOVESKIPGRAMS,
                                           !opt->DOREMOVEFLEXGRAMS, flags.p, st.p);
            CUDA_TRY(cudaMemcpyAsync(&h, st.p, sizeof h, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            // values: DORESET -> empty; unindexed file read as indexed -> the patterns without their counts (include/patternmodel.h:833-837)
            const bool keep_counts = !opt->DORESET && !(as_index && f.type != COLIBRI_INDEXEDPATTERNMODEL);
            TRY(compact_patterns(all, flags.p, keep_counts ? all->d_counts.p : nullptr, false, keep_counts && file_refs, nullptr, m, &launches));
            return 0;
        };
        rc = body();
        if (rc == 0) {
            m->totaltokens = f.tokens;  // include/patternmodel.h:815-816
            m->totaltypes  = f.types;
            if (m->npatterns) {  // postread, :572-588
                m->maxn         = (int)h.kept_maxn;
                m->minn         = (int)h.kept_minn;
                m->hasskipgrams = h.kept_hasskip ? 1 : 0;
            }
            m->counters[2] = launches;
        } else {
            cudaStreamSynchronize(m->stream);
        }
    }
    colibri_b200_model_free(all);
    if (rc) {
        if (m) colibri_b200_model_free(m);
        return rc;
    }
    *out = m;
    return 0;
}

// ------------------------------------------------------------------------------------------------ checksum
// out[0] = sum and out[1] = xor over the patterns of fmix64(FNV1a64(key bytes) ^ count * K), out[2] = total occurrences, out[3] = patterns,
// out[4] = sum over the stored references of fmix64(FNV1a64(key) ^ (sentence << 16 | token) * K') (indexed models, else 0), out[5] = references.
// Order-independent: shares of a sharded model add up (out[1]: xor) to the checksum of the whole.
extern "C" int colibri_b200_model_checksum(colibri_b200_model* m, uint64_t out[6]) {
    if (!m || !out) return set_err(COLIBRI_E_INVALID, "NULL argument");
    memset(out, 0, 6 * sizeof(uint64_t));
    out[3] = m->npatterns;
    if (m->npatterns == 0) return 0;
    CUDA_TRY(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    const bool   refs = m->model_type == COLIBRI_INDEXEDPATTERNMODEL && m->d_ref_off.p && m->nrefs;
    DevBuf<unsigned long long> d;
    DevBuf<uint64_t>           kh;
    TRY(d.alloc(m->device, 8));
    if (refs) TRY(kh.alloc(m->device, m->npatterns));
    CUDA_TRY(cudaMemsetAsync(d.p, 0, 8 * sizeof(unsigned long long), s));
    launch_model_checksum(s, m->d_keys.p, m->d_off.p, m->d_counts.p, m->npatterns, refs ? kh.p : nullptr, d.p);
    if (refs) launch_refs_checksum(s, kh.p, m->d_ref_off.p, m->npatterns, m->d_ref_sentence.p, m->d_ref_token.p, m->nrefs, d.p);
    unsigned long long h[8];
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h, d.p, sizeof h, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    out[0] = h[0]; out[1] = h[1]; out[2] = h[2]; out[4] = h[4];
    out[5] = refs ? m->nrefs : 0;
    return 0;
}

// ------------------------------------------------------------------------------------------------ queries
extern "C" int colibri_b200_model_lookup_batch(colibri_b200_model* m, const uint8_t* keys, const uint64_t* key_off, uint64_t n, uint32_t* counts, int64_t* index) {
    if (!m || !key_off || (!counts && !index)) return set_err(COLIBRI_E_INVALID, "NULL argument");
    if (n == 0) return 0;
    const uint64_t kb = key_off[n];
    if (kb && !keys) return set_err(COLIBRI_E_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(m->device));
    uint64_t launches = 0;
    TRY(ensure_index(m, &launches));
    cudaStream_t     s = m->stream;
    DevBuf<uint8_t>  dk;
    DevBuf<uint64_t> doff;
    DevBuf<uint32_t> didx, dcnt;
    TRY(dk.alloc(m->device, kb + 1));
    TRY(doff.alloc(m->device, n + 1));
    TRY(didx.alloc(m->device, n));
    TRY(dcnt.alloc(m->device, n));
    if (kb) CUDA_TRY(cudaMemcpyAsync(dk.p, keys, kb, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(doff.p, key_off, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    launches += launch_index_lookup(s, dk.p, doff.p, n, m->d_keys.p, m->d_off.p, m->d_index.p, m->index_cap, m->d_presence.p, m->presence_bits, didx.p);
    if (counts) {
        launches += launch_gather_counts(s, didx.p, n, m->d_counts.p, dcnt.p);
        CUDA_TRY(cudaMemcpyAsync(counts, dcnt.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    std::vector<uint32_t> hidx;
    if (index) {
        hidx.resize(n);
        CUDA_TRY(cudaMemcpyAsync(hidx.data(), didx.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    if (index)
        for (uint64_t i = 0; i < n; ++i) index[i] = (int64_t)hidx[i] - 1;
    m->counters[2] += launches;
    return 0;
}

extern "C" int colibri_b200_model_lookup(colibri_b200_model* m, const uint8_t* key, uint32_t len, uint32_t* count) {
    if (!m || !count || (!key && len)) return set_err(COLIBRI_E_INVALID, "NULL argument");
    const uint64_t off[2] = {0, len};
    *count = 0;
    if (len == 0) return 0;
    return colibri_b200_model_lookup_batch(m, key, off, 1, count, nullptr);
}
