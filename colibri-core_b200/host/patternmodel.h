// patternmodel.h -- the reference's C++ training API, re-implemented on top of the B200 C ABI (include/colibri_b200.h).
//
// Same class names, method names, argument order and error behaviour as the reference for the training path:
//   PatternModelOptions                       reference include/patternmodel.h:103-213
//   IndexedCorpus                             reference include/patternstore.h:43-401, src/pattern.cpp:1900-2162 (load / sentences only)
//   PatternModel<uint32_t>::train(...)        reference include/patternmodel.h:880 (istream) and :1353 (filename)
//   IndexedPatternModel<>::train(...)         reference include/patternmodel.h:2821-2844
//   size/has/occurrencecount/types/tokens/maxlength/minlength/begin/end/write   reference :744, :751, :1653, :1700, :1709, :1640-1648, :1609-1629
// so that a caller such as src/patternmodeller.cpp:316-319 compiles against this header unchanged.  train() stages the
// corpus bytes in HBM, runs every counting pass on the GPU and keeps the resulting patterns as flat host arrays; the
// std::unordered_map view the reference exposes is materialised lazily, only if a caller iterates or looks patterns up.
// Progress lines on std::cerr are the reference's own (patternmodel.h:922-931, :1005-1019, :1190-1245) unless options.QUIET.
// Errors print a message on std::cerr and throw InternalError (reference include/common.h:41-44).  There is no CPU
// implementation behind this header: option combinations outside the accelerated subset throw, they do not fall back.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <chrono>
#include <fstream>
#include <iostream>
#include <istream>
#include <iterator>
#include <map>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/colibri_b200.h"
#include "pattern.h"

enum ModelType {  // reference include/patternmodel.h:68-75
    UNINDEXEDPATTERNMODEL = 10,
    UNINDEXEDPATTERNPOINTERMODEL = 11,
    INDEXEDPATTERNMODEL = 20,
    INDEXEDPATTERNPOINTERMODEL = 21,
    PATTERNSETMODEL = 30,
    PATTERNALIGNMENTMODEL = 40,
};

/// Same public fields and defaults as the reference (include/patternmodel.h:105-180).
class PatternModelOptions {
  public:
    int  MINTOKENS = -1;
    int  MINTOKENS_SKIPGRAMS = -1;
    int  MINTOKENS_UNIGRAMS = 1;
    int  MINLENGTH = 1;
    int  MAXLENGTH = 100;
    int  MAXBACKOFFLENGTH = 100;
    bool DOSKIPGRAMS = false;
    bool DOSKIPGRAMS_EXHAUSTIVE = false;
    int  MINSKIPTYPES = 2;
    int  MAXSKIPS = 3;
    bool DOREVERSEINDEX = true;
    bool DOPATTERNPERLINE = false;
    int  PRUNENONSUBSUMED = 0;
    int  PRUNESUBSUMED = 0;
    bool DOREMOVEINDEX = false;
    bool DOREMOVENGRAMS = false;
    bool DOREMOVESKIPGRAMS = false;
    bool DOREMOVEFLEXGRAMS = false;
    bool DORESET = false;
    bool QUIET = false;
    bool DEBUG = false;
};

namespace colibri_b200_detail {
// where the host side of a call spends its time (seconds, summed over the process): reading the corpus file, the device call (staging +
// training), taking the flat result over (device -> host), building the unordered_map view, writing the model file.  The CLI prints it.
struct HostTimes {
    double read = 0, device = 0, adopt = 0, materialise = 0, write = 0;
};
inline HostTimes& host_times() {
    static HostTimes t;
    return t;
}
struct Stopwatch {
    double&                               acc;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit Stopwatch(double& a) : acc(a) {}
    ~Stopwatch() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};
// the rest of a stream into memory: one read() when the stream can tell how much is left (a file), byte iterators otherwise
inline std::vector<unsigned char> read_all(std::istream& in) {
    std::vector<unsigned char> all;
    std::streamoff             left = -1;
    const std::streampos       cur  = in.tellg();
    if (in.good() && cur != std::streampos(-1)) {
        in.seekg(0, std::ios::end);
        const std::streampos end = in.tellg();
        if (in.good() && end != std::streampos(-1)) left = end - cur;
        in.clear();
        in.seekg(cur);
    } else {
        in.clear();
    }
    if (left > 0) {
        all.resize((size_t)left);
        in.read(reinterpret_cast<char*>(all.data()), left);
        all.resize((size_t)std::max<std::streamsize>(in.gcount(), 0));
        in.clear();
    } else if (left != 0) {
        all.assign((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    }
    return all;
}
}  // namespace colibri_b200_detail

/// A corpus held in host memory (the reference's "reverse index").  Only what the training path uses: load(), sentences().
class IndexedCorpus {
    std::vector<unsigned char> body_;  // bytes after the 0xA2 0x02 header
    uint32_t                   sentences_ = 0;

  public:
    IndexedCorpus() {}
    explicit IndexedCorpus(std::istream& in, bool debug = false) { load(in, debug); }
    explicit IndexedCorpus(const std::string& filename, bool debug = false) { load(filename, debug); }
    void load(std::istream& in, bool = false) {
        if (!in.good()) {
            std::cerr << "ERROR: Supplied data file can not be opened. Check whether it exists and whether you have proper permissions..." << std::endl;
            throw InternalError();
        }
        std::vector<unsigned char> all;
        {
            colibri_b200_detail::Stopwatch sw(colibri_b200_detail::host_times().read);
            all = colibri_b200_detail::read_all(in);
        }
        if (all.size() < 2 || all[0] != 0xA2 || all[1] != 2) {
            std::cerr << "ERROR: the B200 build reads class-encoded corpora of data version 2 only (0xA2 0x02 header)" << std::endl;
            throw InternalError();
        }
        body_.assign(all.begin() + 2, all.end());
        // sentence index as in reference src/pattern.cpp:1944-1958: a sentence starts at byte 0 and after every delimiter
        sentences_         = 0;
        bool prevdelimiter = true, prevhigh = false;
        for (unsigned char c : body_) {
            if (prevdelimiter) {
                ++sentences_;
                prevdelimiter = false;
            }
            if (!prevhigh && c == 0) prevdelimiter = true;
            prevhigh = c >= 128;
        }
    }
    void load(const std::string& filename, bool debug = false) {
        std::ifstream in(filename, std::ios::in | std::ios::binary);
        if (!in.good()) {
            std::cerr << "ERROR: Unable to load file " << filename << std::endl;
            throw InternalError();
        }
        load(in, debug);
    }
    unsigned int         sentences() const { return sentences_; }
    size_t               bytesize() const { return body_.size(); }
    const unsigned char* beginpointer() const { return body_.data(); }
};

namespace colibri_b200_detail {
struct FlatView {  // a model's patterns as the C ABI takes them (colibri_b200_model_from_flat)
    const uint8_t*  keys = nullptr;
    const uint64_t* off = nullptr;
    const uint32_t* counts = nullptr;
    uint64_t        npatterns = 0;
};
}  // namespace colibri_b200_detail

class PatternModelInterface {  // reference include/patternmodel.h:234-287
  public:
    virtual ~PatternModelInterface() {}
    virtual int    getmodeltype() const = 0;
    virtual int    getmodelversion() const = 0;
    virtual size_t occurrencecount(const Pattern& pattern) = 0;
    virtual int    maxlength() const = 0;
    virtual int    minlength() const = 0;
    virtual size_t types() = 0;
    virtual size_t tokens() const = 0;
    /// B200 build: what stands in for getstoreinterface() (reference :286) -- the patterns as flat arrays, to be uploaded as a constraint set
    virtual colibri_b200_detail::FlatView flatview() const = 0;
};

/// Placeholder so that the train() signatures match the reference's (filter argument); filters are not on the device path.
template <class T = uint32_t>
class PatternSet {
  public:
    size_t size() const { return 0; }
};

namespace colibri_b200_detail {
inline int& default_device() {  // CUDA device the next train() call uses (the reference has no such notion)
    static int device = 0;
    return device;
}
inline std::vector<int>& device_list() {  // more than one entry: train() shards the corpus over these GPUs (colibri_b200_train_multi; CLI -d 0-7)
    static std::vector<int> devices;
    return devices;
}
inline void fail(const std::string& msg) {
    std::cerr << "ERROR: " << msg << std::endl;
    throw InternalError();
}
struct ModelHandle {  // frees the C-ABI handle
    colibri_b200_model* h = nullptr;
    ~ModelHandle() { colibri_b200_model_free(h); }
};
}  // namespace colibri_b200_detail

template <class ValueType, int kModelType>
class DevicePatternModel : public PatternModelInterface {
  protected:
    std::vector<uint8_t>  keys_;
    std::vector<uint64_t> off_;
    std::vector<uint32_t> counts_;
    std::vector<uint32_t> ref_sentence_;
    std::vector<uint16_t> ref_token_;
    std::vector<uint64_t> ref_off_;
    uint64_t              totaltokens = 0, totaltypes = 0;
    int                   maxn = 0, minn = 999;
    IndexedCorpus*        reverseindex = nullptr;
    std::unordered_map<Pattern, ValueType> map_;  // lazily built view
    bool                                   map_ready_ = false;

    void materialise();

    static colibri_b200_options c_options(const PatternModelOptions& options, bool streamed) {
        colibri_b200_options o;
        colibri_b200_options_default(&o);
        o.MINTOKENS              = options.MINTOKENS;
        o.MINTOKENS_SKIPGRAMS    = options.MINTOKENS_SKIPGRAMS;
        o.MINTOKENS_UNIGRAMS     = options.MINTOKENS_UNIGRAMS;
        o.MINLENGTH              = options.MINLENGTH;
        o.MAXLENGTH              = options.MAXLENGTH;
        o.MAXBACKOFFLENGTH       = options.MAXBACKOFFLENGTH;
        o.MINSKIPTYPES           = options.MINSKIPTYPES;
        o.MAXSKIPS               = options.MAXSKIPS;
        o.DOSKIPGRAMS            = options.DOSKIPGRAMS;
        o.DOSKIPGRAMS_EXHAUSTIVE = options.DOSKIPGRAMS_EXHAUSTIVE;
        o.DOPATTERNPERLINE       = options.DOPATTERNPERLINE;
        o.PRUNENONSUBSUMED       = options.PRUNENONSUBSUMED;
        o.PRUNESUBSUMED          = options.PRUNESUBSUMED;
        o.DOREMOVEINDEX          = options.DOREMOVEINDEX;
        o.DOREMOVENGRAMS         = options.DOREMOVENGRAMS;
        o.DOREMOVESKIPGRAMS      = options.DOREMOVESKIPGRAMS;
        o.DOREMOVEFLEXGRAMS      = options.DOREMOVEFLEXGRAMS;
        o.DORESET                = options.DORESET;
        o.QUIET                  = options.QUIET;
        o.DEBUG                  = options.DEBUG;
        o.model_type             = kModelType;
        o.streamed               = streamed ? 1 : 0;
        o.device                 = colibri_b200_detail::default_device();
        return o;
    }

    /// take over the result of a device call: header numbers + the flat export
    void adopt(colibri_b200_model* h) {
        using colibri_b200_detail::fail;
        colibri_b200_detail::Stopwatch sw(colibri_b200_detail::host_times().adopt);
        totaltokens  = colibri_b200_model_tokens(h);
        totaltypes   = colibri_b200_model_types(h);
        maxn         = colibri_b200_model_maxn(h);
        minn         = colibri_b200_model_minn(h);
        hasskipgrams = colibri_b200_model_hasskipgrams(h) != 0;
        uint64_t np = 0, kb = 0, nr = 0;
        if (colibri_b200_model_export_sizes(h, &np, &kb, &nr) != COLIBRI_OK) fail(colibri_b200_last_error());
        keys_.assign(kb + 1, 0);
        off_.assign(np + 1, 0);
        counts_.assign(np + 1, 0);
        ref_sentence_.clear();
        ref_token_.clear();
        ref_off_.clear();
        if (kModelType == INDEXEDPATTERNMODEL) {
            ref_sentence_.assign(nr + 1, 0);
            ref_token_.assign(nr + 1, 0);
            ref_off_.assign(np + 1, 0);
        }
        if (colibri_b200_model_export(h, keys_.data(), off_.data(), counts_.data(), kModelType == INDEXEDPATTERNMODEL ? ref_sentence_.data() : nullptr,
                                      kModelType == INDEXEDPATTERNMODEL ? ref_token_.data() : nullptr, kModelType == INDEXEDPATTERNMODEL ? ref_off_.data() : nullptr) != COLIBRI_OK)
            fail(colibri_b200_last_error());
        counts_.resize(np);
        map_.clear();
        map_ready_ = false;
    }

    /// train() under a constraint model (reference include/patternmodel.h:880-1345 with constrainbymodel != NULL): one scan, only the
    /// constraint model's patterns are counted.  constrainbymodel == this is the in-place rebuild of the CLI's -I / -2.
    void train_constrained_body(const unsigned char* body, size_t nbytes, bool streamed, const PatternModelOptions& options, PatternModelInterface* constrainbymodel) {
        using colibri_b200_detail::fail;
        const bool           inplace = constrainbymodel == static_cast<PatternModelInterface*>(this);
        colibri_b200_options o       = c_options(options, streamed);
        const int mintokens          = options.MINTOKENS == -1 ? 2 : (options.MINTOKENS == 0 ? 1 : options.MINTOKENS);
        if (!options.QUIET) std::cerr << "Training patternmodel, constrained by another model, occurrence threshold: " << mintokens << std::endl;  // reference :922-931
        if (options.DOSKIPGRAMS || options.DOSKIPGRAMS_EXHAUSTIVE) fail("skipgrams under a constraint model are not available in the B200 build");
        const colibri_b200_detail::FlatView v = constrainbymodel->flatview();
        colibri_b200_detail::ModelHandle    cm, mh;
        const uint64_t                      zero = 0;
        // an in-place rebuild starts from reset values (the caller loaded with DORESET): only the patterns matter
        if (colibri_b200_model_from_flat(v.keys, v.npatterns ? v.off : &zero, nullptr, v.npatterns, nullptr, nullptr, nullptr, constrainbymodel->tokens(), constrainbymodel->types(),
                                         UNINDEXEDPATTERNMODEL, o.device, &cm.h) != COLIBRI_OK)
            fail(colibri_b200_last_error());
        colibri_b200_corpus* corpus = nullptr;
        if (colibri_b200_corpus_stage(body, nbytes, o.device, &corpus) != COLIBRI_OK) fail(colibri_b200_last_error());
        if (!options.QUIET) std::cerr << "Counting n-grams that occur in constraint model" << std::endl;  // reference :1010-1011
        const int rc = colibri_b200_train_constrained(corpus, &o, cm.h, inplace ? 1 : 0, &mh.h);
        colibri_b200_corpus_free(corpus);
        if (rc != COLIBRI_OK) fail(colibri_b200_last_error());
        if (!options.QUIET) {
            uint64_t st[4];
            if (colibri_b200_model_passes(mh.h) > 0 && colibri_b200_model_pass_stats(mh.h, 0, st) == COLIBRI_OK)
                std::cerr << " Found " << st[1] << " ngrams...pruned " << st[3] << "...total kept: " << st[1] - st[3] << std::endl;  // reference :1195-1245
            else
                std::cerr << "None found" << std::endl;
        }
        const int keep_maxn = maxn, keep_minn = minn;
        adopt(mh.h);
        if (inplace) {  // maxn / minn only ever widen across load() and train() (reference :578-581, :1184-1188)
            maxn = std::max(maxn, keep_maxn);
            minn = std::min(minn, keep_minn);
        }
    }

    void train_body(const unsigned char* body, size_t nbytes, bool streamed, const PatternModelOptions& options, PatternModelInterface* constrainbymodel, PatternSet<>* filter,
                    bool continued, uint32_t firstsentence) {
        using colibri_b200_detail::fail;
        if (filter != nullptr && filter->size() > 0) fail("training with a pattern filter is not available in the B200 build");
        if (continued) fail("continued training is not available in the B200 build");
        if (firstsentence != 1) fail("firstsentence != 1 is not available in the B200 build");
        if (constrainbymodel != nullptr) {
            train_constrained_body(body, nbytes, streamed, options, constrainbymodel);
            return;
        }
        colibri_b200_options o;
        colibri_b200_options_default(&o);
        o.MINTOKENS              = options.MINTOKENS;
        o.MINTOKENS_SKIPGRAMS    = options.MINTOKENS_SKIPGRAMS;
        o.MINTOKENS_UNIGRAMS     = options.MINTOKENS_UNIGRAMS;
        o.MINLENGTH              = options.MINLENGTH;
        o.MAXLENGTH              = options.MAXLENGTH;
        o.MAXBACKOFFLENGTH       = options.MAXBACKOFFLENGTH;
        o.MINSKIPTYPES           = options.MINSKIPTYPES;
        o.MAXSKIPS               = options.MAXSKIPS;
        o.DOSKIPGRAMS            = options.DOSKIPGRAMS;
        o.DOSKIPGRAMS_EXHAUSTIVE = options.DOSKIPGRAMS_EXHAUSTIVE;
        o.DOPATTERNPERLINE       = options.DOPATTERNPERLINE;
        o.PRUNENONSUBSUMED       = options.PRUNENONSUBSUMED;
        o.PRUNESUBSUMED          = options.PRUNESUBSUMED;
        o.QUIET                  = options.QUIET;
        o.DEBUG                  = options.DEBUG;
        o.model_type             = kModelType;
        o.streamed               = streamed ? 1 : 0;
        o.device                 = colibri_b200_detail::default_device();
        const int mintokens      = options.MINTOKENS == -1 ? 2 : (options.MINTOKENS == 0 ? 1 : options.MINTOKENS);
        if (!options.QUIET) std::cerr << "Training patternmodel, occurrence threshold: " << mintokens << std::endl;  // reference :922-931
        colibri_b200_detail::ModelHandle mh;
        const std::vector<int>&          devs = colibri_b200_detail::device_list();
        std::vector<colibri_b200_model*> shares;
        {
            colibri_b200_detail::Stopwatch sw(colibri_b200_detail::host_times().device);
            if (devs.size() > 1) {
                // several GPUs: every device returns its share of the model (same header numbers in each); the shares are concatenated below
                shares.assign(devs.size(), nullptr);
                if (colibri_b200_train_multi(body, nbytes, &o, devs.data(), (int)devs.size(), shares.data()) != COLIBRI_OK) fail(colibri_b200_last_error());
                mh.h      = shares[0];
                shares[0] = nullptr;
            } else if (colibri_b200_train(body, nbytes, &o, &mh.h) != COLIBRI_OK)
                fail(colibri_b200_last_error());
        }
        totaltokens  = colibri_b200_model_tokens(mh.h);
        totaltypes   = colibri_b200_model_types(mh.h);
        maxn         = colibri_b200_model_maxn(mh.h);
        minn         = colibri_b200_model_minn(mh.h);
        hasskipgrams = colibri_b200_model_hasskipgrams(mh.h) != 0;
        if (!options.QUIET) {  // the reference's per-pass progress lines (:1005-1019, :1190-1245)
            const int np = colibri_b200_model_passes(mh.h);
            int ngrampasses = 0;
            for (int p = 0; p < np; ++p) {
                uint64_t st[4];
                colibri_b200_model_pass_stats(mh.h, p, st);
                if (options.DOSKIPGRAMS && st[1] == 0) continue;  // trainskipgrams passes: printed below, after the n-gram passes
                ++ngrampasses;
                if (mintokens > 1)
                    std::cerr << "Counting " << st[0] << "-grams" << std::endl;
                else
                    std::cerr << "Counting *all* n-grams (occurrence threshold=1)" << std::endl;
                std::cerr << " Found " << st[1] << " ngrams...";
                if (options.DOSKIPGRAMS_EXHAUSTIVE) std::cerr << st[2] << " skipgram occurrences...";
                std::cerr << "pruned " << st[3] << "...total kept: " << (st[1] + st[2]) - st[3] << std::endl;
            }
            if (mintokens > 1 && ngrampasses > 0 && ngrampasses < options.MAXLENGTH) {
                uint64_t st[4];
                colibri_b200_model_pass_stats(mh.h, ngrampasses - 1, st);
                std::cerr << "Counting " << st[0] + 1 << "-grams" << std::endl << "None found" << std::endl;
            }
            if (options.DOSKIPGRAMS) {  // IndexedPatternModel::trainskipgrams, reference :2980-3008
                int last = 2;
                for (int p = ngrampasses; p < np; ++p) {
                    uint64_t st[4];
                    colibri_b200_model_pass_stats(mh.h, p, st);
                    last = (int)st[0];
                    std::cerr << "Counting " << st[0] << "-skipgrams" << std::endl;
                    std::cerr << " Found " << st[2] << " skipgrams...pruned " << st[3] << "...total kept: " << st[2] - st[3] << std::endl;
                }
                if (last < options.MAXLENGTH) std::cerr << "Counting " << last + 1 << "-skipgrams" << std::endl << " None found" << std::endl;
            }
        }
        adopt(mh.h);
        for (size_t r = 1; r < shares.size(); ++r) {  // the other devices' shares of a multi-GPU run
            colibri_b200_detail::ModelHandle sh;
            sh.h = shares[r];
            append_share(sh.h);
        }
    }

    /// append the flat export of another share of the same model (unindexed shares of colibri_b200_train_multi)
    void append_share(colibri_b200_model* h) {
        using colibri_b200_detail::fail;
        uint64_t np = 0, kb = 0, nr = 0;
        if (colibri_b200_model_export_sizes(h, &np, &kb, &nr) != COLIBRI_OK) fail(colibri_b200_last_error());
        if (np == 0) return;
        std::vector<uint8_t>  k(kb + 1);
        std::vector<uint64_t> of(np + 1);
        std::vector<uint32_t> c(np + 1);
        if (colibri_b200_model_export(h, k.data(), of.data(), c.data(), nullptr, nullptr, nullptr) != COLIBRI_OK) fail(colibri_b200_last_error());
        const size_t   np0  = counts_.size();
        const uint64_t base = off_[np0];
        keys_.resize(base);
        keys_.insert(keys_.end(), k.begin(), k.begin() + kb);
        keys_.push_back(0);
        off_.resize(np0 + 1);
        for (uint64_t i = 1; i <= np; ++i) off_.push_back(base + of[i]);
        counts_.insert(counts_.end(), c.begin(), c.begin() + np);
        map_.clear();
        map_ready_ = false;
    }

  public:
    bool hasskipgrams = false;
    bool hasflexgrams = false;

    DevicePatternModel(IndexedCorpus* corpus = nullptr) : reverseindex(corpus) {}
    /// Read a pattern model from file (reference include/patternmodel.h:700-726); the options act as filters
    DevicePatternModel(const std::string& filename, const PatternModelOptions& options, PatternModelInterface* constrainmodel = nullptr, IndexedCorpus* corpus = nullptr)
        : reverseindex(corpus) {
        if (!options.QUIET) std::cerr << "Loading " << filename << std::endl;
        std::ifstream in(filename, std::ios::in | std::ios::binary);
        if (!in.good()) {
            std::cerr << "ERROR: Unable to load file " << filename << std::endl;
            throw InternalError();
        }
        this->load(in, options, constrainmodel);
    }
    DevicePatternModel(std::istream& f, const PatternModelOptions& options, PatternModelInterface* constrainmodel = nullptr, IndexedCorpus* corpus = nullptr) : reverseindex(corpus) {
        this->load(f, options, constrainmodel);
    }
    virtual ~DevicePatternModel() {}

    /// Read a pattern model from a stream (reference include/patternmodel.h:781-861 over PatternMapStore::read, include/patternstore.h:555-619):
    /// count >= MINTOKENS, MINLENGTH <= n <= MAXLENGTH, DOREMOVE{NGRAMS,SKIPGRAMS,FLEXGRAMS}, membership in constrainmodel, DORESET.
    /// The record stream is scanned on the host; shapes, filters, the constraint test and the compaction run on the device.
    virtual void load(std::istream& f, const PatternModelOptions& options, PatternModelInterface* constrainmodel = nullptr) {
        using colibri_b200_detail::fail;
        std::vector<unsigned char> all;
        {
            colibri_b200_detail::Stopwatch sw(colibri_b200_detail::host_times().read);
            all = colibri_b200_detail::read_all(f);
        }
        if (all.size() < 3 || all[0] != 0 || (all[1] != UNINDEXEDPATTERNMODEL && all[1] != INDEXEDPATTERNMODEL)) {
            if (all.size() >= 3 && all[0] == 0 && (all[1] == UNINDEXEDPATTERNPOINTERMODEL || all[1] == INDEXEDPATTERNPOINTERMODEL || all[1] == PATTERNALIGNMENTMODEL))
                fail("pointer models and alignment models are not read by the B200 build");
            std::cerr << "File is not a colibri model file (or a very old one)" << std::endl;  // reference :788-793
            throw InternalError();
        }
        colibri_b200_options             o = c_options(options, true);
        colibri_b200_detail::ModelHandle cm, mh;
        if (constrainmodel != nullptr) {
            const colibri_b200_detail::FlatView v    = constrainmodel->flatview();
            const uint64_t                      zero = 0;
            if (colibri_b200_model_from_flat(v.keys, v.npatterns ? v.off : &zero, nullptr, v.npatterns, nullptr, nullptr, nullptr, 0, 0, UNINDEXEDPATTERNMODEL, o.device, &cm.h) != COLIBRI_OK)
                fail(colibri_b200_last_error());
        }
        if (colibri_b200_model_load(all.data(), all.size(), &o, cm.h, &mh.h) != COLIBRI_OK) fail(colibri_b200_last_error());
        adopt(mh.h);
    }
    virtual void load(std::string& filename, const PatternModelOptions& options, PatternModelInterface* constrainmodel = nullptr) {
        if (!options.QUIET) std::cerr << "Loading " << filename << std::endl;
        std::ifstream in(filename, std::ios::in | std::ios::binary);
        if (!in.good()) {
            std::cerr << "ERROR: Unable to load file " << filename << std::endl;
            throw InternalError();
        }
        this->load(in, options, constrainmodel);
    }

    /// Compute flexgrams by abstracting from the skipgrams in the model (reference include/patternmodel.h:3724-3744, IndexedPatternModel only):
    /// returns the number of flexgrams found.  The model goes to the device as it is, comes back with the flexgrams appended.
    int computeflexgrams_fromskipgrams() {
        using colibri_b200_detail::fail;
        if (kModelType != INDEXEDPATTERNMODEL) fail("computeflexgrams_fromskipgrams is defined for indexed models only");
        colibri_b200_detail::ModelHandle cur, res;
        const uint64_t                   zero = 0;
        const bool                       any  = !counts_.empty();
        if (colibri_b200_model_from_flat(keys_.data(), any ? off_.data() : &zero, counts_.data(), counts_.size(), ref_sentence_.data(), ref_token_.data(), any ? ref_off_.data() : &zero,
                                         totaltokens, totaltypes, INDEXEDPATTERNMODEL, colibri_b200_detail::default_device(), &cur.h) != COLIBRI_OK)
            fail(colibri_b200_last_error());
        uint64_t found = 0;
        if (colibri_b200_model_flexgrams_fromskipgrams(cur.h, &found, &res.h) != COLIBRI_OK) fail(colibri_b200_last_error());
        const int keep_maxn = maxn, keep_minn = minn;
        const bool keep_skip = hasskipgrams;
        adopt(res.h);
        maxn         = keep_maxn;  // the reference leaves maxn / minn alone here
        minn         = keep_minn;
        hasskipgrams = keep_skip;
        if (found) hasflexgrams = true;
        return (int)found;
    }

    /// reference include/patternmodel.h:866-868
    PatternModelInterface* getinterface() { return static_cast<PatternModelInterface*>(this); }
    colibri_b200_detail::FlatView flatview() const override {
        colibri_b200_detail::FlatView v;
        v.keys      = keys_.data();
        v.off       = off_.data();
        v.counts    = counts_.data();
        v.npatterns = counts_.size();
        return v;
    }

    int getmodeltype() const override { return kModelType; }
    int getmodelversion() const override { return 2; }
    unsigned char type() const { return (unsigned char)kModelType; }
    unsigned char version() const { return 2; }

    /// Train on corpus data read from a stream (a *.colibri.dat); `in` may be NULL if a preloaded corpus was given to the constructor.
    virtual void train(std::istream* in, const PatternModelOptions& options, PatternModelInterface* constrainbymodel = nullptr, PatternSet<>* filter = nullptr, bool continued = false,
                       uint32_t firstsentence = 1, bool ignoreerrors = false) {
        (void)ignoreerrors;
        if (reverseindex != nullptr) {  // reference :1030-1037: sentences come from the preloaded corpus when there is one
            train_body(reverseindex->beginpointer(), reverseindex->bytesize(), false, options, constrainbymodel, filter, continued, firstsentence);
            return;
        }
        if (in == nullptr) colibri_b200_detail::fail("train() needs an input stream or a preloaded IndexedCorpus");
        if (!in->good()) {
            std::cerr << "ERROR: Supplied data file can not be opened. Check whether it exists and whether you have proper permissions..." << std::endl;  // classdecoder.cpp:263-266
            throw InternalError();
        }
        std::vector<unsigned char> all;
        {
            colibri_b200_detail::Stopwatch sw(colibri_b200_detail::host_times().read);
            in->clear();
            in->seekg(0);
            all = colibri_b200_detail::read_all(*in);
        }
        if (all.size() < 2 || all[0] != 0xA2 || all[1] != 2)
            colibri_b200_detail::fail("the B200 build reads class-encoded corpora of data version 2 only (0xA2 0x02 header)");
        train_body(all.data() + 2, all.size() - 2, true, options, constrainbymodel, filter, continued, firstsentence);
    }
    /// Train on a corpus file (*.colibri.dat).
    virtual void train(const std::string& filename, const PatternModelOptions& options, PatternModelInterface* constrainbymodel = nullptr, PatternSet<>* filter = nullptr,
                       bool continued = false, uint32_t firstsentence = 1, bool ignoreerrors = false) {
        if (filename.size() > 3 && filename.substr(filename.size() - 3) == ".bz2") colibri_b200_detail::fail("bz2-compressed corpora are not read by the B200 build");
        std::ifstream in(filename, std::ios::in | std::ios::binary);
        this->train(&in, options, constrainbymodel, filter, continued, firstsentence, ignoreerrors);
    }

    size_t size() const { return counts_.size(); }
    size_t tokens() const override { return totaltokens; }
    size_t types() override { return totaltypes; }
    int    maxlength() const override { return maxn; }
    int    minlength() const override { return minn; }

    bool has(const Pattern& pattern) {
        materialise();
        return map_.find(pattern) != map_.end();
    }
    size_t occurrencecount(const Pattern& pattern) override;
    ValueType* getdata(const Pattern& pattern, bool = false) {
        materialise();
        auto it = map_.find(pattern);
        return it == map_.end() ? nullptr : &it->second;
    }

    typedef typename std::unordered_map<Pattern, ValueType>::iterator       iterator;
    typedef typename std::unordered_map<Pattern, ValueType>::const_iterator const_iterator;
    iterator begin() { materialise(); return map_.begin(); }
    iterator end() { materialise(); return map_.end(); }

    /// The i-th pattern of the flat result (no map involved): fast path for bulk consumers.
    Pattern  flat_pattern(size_t i) const { return Pattern(keys_.data() + off_[i], (int)(off_[i + 1] - off_[i])); }
    uint32_t flat_count(size_t i) const { return counts_[i]; }

    /// Group statistics (reference computestats :1903-1933, computecoveragestats :1946-1984): category 0 = all, NGRAM, SKIPGRAM, FLEXGRAM;
    /// n = 0: all lengths (flexgrams have no per-length entry).  Computed from the flat result, nothing cached -- the reference's cache
    /// answers 0 for a group asked after another one was computed alone; that is not reproduced.
    unsigned int totaloccurrencesingroup(int category, int n) const { return (unsigned int)group_total(category, n, true); }
    unsigned int totalpatternsingroup(int category, int n) const { return (unsigned int)group_total(category, n, false); }
    /// distinct tokens of the group's patterns (a gap counts as a token); asked for length 1, the unigram patterns themselves
    unsigned int totalwordtypesingroup(int category, int n) const {
        std::unordered_map<std::string, char> types;
        for (size_t i = 0; i < counts_.size(); ++i) {
            const Pattern p = flat_pattern(i);
            if (category != 0 && (int)p.category() != category) continue;
            const int pn = (int)p.n();
            if (pn == 1 && n <= 1) {
                types[std::string(reinterpret_cast<const char*>(p.data()), p.bytesize())] = 1;
            } else if (n == 0 || pn == n) {
                const unsigned char* d = p.data();
                size_t               b = 0;
                for (size_t e = 0; e < p.bytesize(); ++e)
                    if (d[e] < 128) {  // a token ends at its first byte below 128
                        types[std::string(reinterpret_cast<const char*>(d + b), e + 1 - b)] = 1;
                        b = e + 1;
                    }
            }
        }
        return (unsigned int)types.size();
    }

  protected:
    uint64_t group_total(int category, int n, bool occurrences) const {
        uint64_t t = 0;
        for (size_t i = 0; i < counts_.size(); ++i) {
            const Pattern p = flat_pattern(i);
            const int     c = (int)p.category();
            if (category != 0 && c != category) continue;
            if (n != 0 && (c == FLEXGRAM || (int)p.n() != n)) continue;
            t += occurrences ? counts_[i] : 1;
        }
        return t;
    }

  public:
    /// Write the model in the reference's binary format (reference :1609-1624, patternstore.h:534-542, datatypes.h:216-221, :263-270).
    void write(std::ostream& out) {
        colibri_b200_detail::Stopwatch sw(colibri_b200_detail::host_times().write);
        const char    null = 0;
        unsigned char t = (unsigned char)kModelType, v = 2;
        out.write(&null, 1);
        out.write(reinterpret_cast<char*>(&t), 1);
        out.write(reinterpret_cast<char*>(&v), 1);
        out.write(reinterpret_cast<char*>(&totaltokens), sizeof(uint64_t));
        const uint64_t tp = totaltypes;
        out.write(reinterpret_cast<const char*>(&tp), sizeof(uint64_t));
        const uint64_t s = counts_.size();
        out.write(reinterpret_cast<const char*>(&s), sizeof(uint64_t));
        // the records are put together in 4 MB pieces: one ostream call per piece instead of three per pattern
        std::string buf;
        buf.reserve((4u << 20) + 4096);
        auto put = [&](const void* p, size_t n) { buf.append(reinterpret_cast<const char*>(p), n); };
        for (size_t i = 0; i < counts_.size(); ++i) {
            put(keys_.data() + off_[i], (size_t)(off_[i + 1] - off_[i]));
            put(&null, 1);
            put(&counts_[i], sizeof(uint32_t));
            if (kModelType == INDEXEDPATTERNMODEL) {
                for (uint64_t j = ref_off_[i]; j < ref_off_[i + 1]; ++j) {
                    put(&ref_sentence_[j], 4);
                    put(&ref_token_[j], 2);
                    if (buf.size() >= (4u << 20)) {
                        out.write(buf.data(), (std::streamsize)buf.size());
                        buf.clear();
                    }
                }
            }
            if (buf.size() >= (4u << 20)) {
                out.write(buf.data(), (std::streamsize)buf.size());
                buf.clear();
            }
        }
        out.write(buf.data(), (std::streamsize)buf.size());
    }
    void write(const std::string& filename) {
        std::ofstream out(filename, std::ios::out | std::ios::binary);
        this->write(out);
    }
};

// ---- unindexed: value = occurrence count
template <class ValueType, int kModelType>
inline void DevicePatternModel<ValueType, kModelType>::materialise() {
    if (map_ready_) return;
    colibri_b200_detail::Stopwatch sw(colibri_b200_detail::host_times().materialise);
    map_.reserve(counts_.size());
    for (size_t i = 0; i < counts_.size(); ++i) {
        if constexpr (kModelType == INDEXEDPATTERNMODEL) {
            ValueType& v = map_[flat_pattern(i)];
            for (uint64_t j = ref_off_[i]; j < ref_off_[i + 1]; ++j) v.data.push_back(IndexReference(ref_sentence_[j], ref_token_[j]));
        } else {
            map_[flat_pattern(i)] = (ValueType)counts_[i];
        }
    }
    map_ready_ = true;
}
template <class ValueType, int kModelType>
inline size_t DevicePatternModel<ValueType, kModelType>::occurrencecount(const Pattern& pattern) {
    materialise();
    auto it = map_.find(pattern);
    if (it == map_.end()) return 0;
    if constexpr (kModelType == INDEXEDPATTERNMODEL)
        return it->second.count();
    else
        return (size_t)it->second;
}

/// PatternModel<uint32_t>: the unindexed model (reference include/patternmodel.h:545)
template <class ValueType = uint32_t>
class PatternModel : public DevicePatternModel<ValueType, UNINDEXEDPATTERNMODEL> {
  public:
    PatternModel(IndexedCorpus* corpus = nullptr) : DevicePatternModel<ValueType, UNINDEXEDPATTERNMODEL>(corpus) {}
    PatternModel(const std::string& filename, const PatternModelOptions& options, PatternModelInterface* constrainmodel = nullptr, IndexedCorpus* corpus = nullptr)
        : DevicePatternModel<ValueType, UNINDEXEDPATTERNMODEL>(filename, options, constrainmodel, corpus) {}
    PatternModel(std::istream& f, const PatternModelOptions& options, PatternModelInterface* constrainmodel = nullptr, IndexedCorpus* corpus = nullptr)
        : DevicePatternModel<ValueType, UNINDEXEDPATTERNMODEL>(f, options, constrainmodel, corpus) {}
};

/// PatternSetModel: patterns without values, what the CLI loads as the constraint model of -j (reference include/patternmodel.h:296-470).
/// Reading an (un)indexed model file as a set applies the same filters (readmap, :421-431); the counts are simply not looked at.
class PatternSetModel : public DevicePatternModel<uint32_t, UNINDEXEDPATTERNMODEL> {
  public:
    PatternSetModel() {}
    PatternSetModel(const std::string& filename, const PatternModelOptions& options, PatternModelInterface* constrainmodel = nullptr)
        : DevicePatternModel<uint32_t, UNINDEXEDPATTERNMODEL>(filename, options, constrainmodel, nullptr) {}
    PatternSetModel(std::istream& f, const PatternModelOptions& options, PatternModelInterface* constrainmodel = nullptr)
        : DevicePatternModel<uint32_t, UNINDEXEDPATTERNMODEL>(f, options, constrainmodel, nullptr) {}
    int getmodeltype() const override { return PATTERNSETMODEL; }
};

/// IndexedPatternModel<>: value = sorted list of positions (reference include/patternmodel.h:2681)
template <class MapType = void>
class IndexedPatternModel : public DevicePatternModel<IndexedData, INDEXEDPATTERNMODEL> {
  public:
    IndexedPatternModel(IndexedCorpus* corpus = nullptr) : DevicePatternModel<IndexedData, INDEXEDPATTERNMODEL>(corpus) {}
    IndexedPatternModel(const std::string& filename, const PatternModelOptions& options, PatternModelInterface* constrainmodel = nullptr, IndexedCorpus* corpus = nullptr)
        : DevicePatternModel<IndexedData, INDEXEDPATTERNMODEL>(filename, options, constrainmodel, corpus) {}
    IndexedPatternModel(std::istream& f, const PatternModelOptions& options, PatternModelInterface* constrainmodel = nullptr, IndexedCorpus* corpus = nullptr)
        : DevicePatternModel<IndexedData, INDEXEDPATTERNMODEL>(f, options, constrainmodel, corpus) {}
    ~IndexedPatternModel() { drop_rindex(); }

    typedef std::map<Pattern, uint64_t> t_relationmap;  // reference include/patternmodel.h:2650 (unordered there; ordered here for stable output)

    /// The model's patterns that begin at a corpus position (reference include/patternmodel.h:1746-1824; n-grams).  The reverse index is built
    /// on the device on first use: every window of the corpus is matched against the model once (colibri_b200_rindex_build).
    std::vector<Pattern> getreverseindex(const IndexReference ref, int occurrencecount = 0, int category = 0, unsigned int size = 0) {
        ensure_rindex();
        std::vector<uint32_t> row(rlengths_.size(), 0);
        if (!row.empty() && colibri_b200_rindex_query(rindex_, &ref.sentence, &ref.token, 1, row.data()) != COLIBRI_OK) colibri_b200_detail::fail(colibri_b200_last_error());
        std::vector<Pattern> out;
        for (size_t k = 0; k < row.size(); ++k) {
            if (!row[k]) continue;
            if (size && rlengths_[k] != size) continue;
            if (category && category != NGRAM) continue;
            if (occurrencecount && counts_[row[k] - 1] < (uint32_t)occurrencecount) continue;
            out.push_back(pattern_at(row[k] - 1));
        }
        return out;
    }
    /// getrightcooc / getleftcooc (reference :3460-3493, :3502-3531) with the reference's arithmetic -- see include/colibri_b200.h, colibri_b200_rindex_cooc.
    t_relationmap getrightcooc(const Pattern& pattern, unsigned int occurrencethreshold = 0, int category = 0, unsigned int size = 0) { return cooc_of(pattern, 0, occurrencethreshold, category, size); }
    t_relationmap getleftcooc(const Pattern& pattern, unsigned int occurrencethreshold = 0, int category = 0, unsigned int size = 0) { return cooc_of(pattern, 1, occurrencethreshold, category, size); }
    /// getcooc (reference :3543-3576): both directions, neither overlapping nor adjacent; counted on the device (colibri_b200_rindex_cooc_of), the
    /// reference's filters applied here (threshold on the neighbour's occurrence count, category, size, ordersignificant, prunerelations)
    t_relationmap getcooc(const Pattern& pattern, unsigned int occurrencethreshold = 0, int category = 0, unsigned int size = 0, bool ordersignificant = false) {
        ensure_rindex();
        if (!this->has(pattern)) throw NoSuchPattern();
        const uint64_t idx = index_of(pattern);
        uint64_t n = 0;
        if (colibri_b200_rindex_cooc_of(rindex_, idx, nullptr, nullptr, 0, &n) != COLIBRI_OK) colibri_b200_detail::fail(colibri_b200_last_error());
        std::vector<uint32_t> q(n);
        std::vector<uint64_t> c(n);
        if (n && colibri_b200_rindex_cooc_of(rindex_, idx, q.data(), c.data(), n, &n) != COLIBRI_OK) colibri_b200_detail::fail(colibri_b200_last_error());
        t_relationmap out;
        for (uint64_t k = 0; k < n; ++k) {
            const Pattern nb = pattern_at(q[k]);
            if (ordersignificant && nb < pattern) continue;
            if (occurrencethreshold && (counts_[q[k]] < occurrencethreshold || c[k] < occurrencethreshold)) continue;
            if (category && (int)nb.category() != category) continue;
            if (size && nb.n() != size) continue;
            out[nb] = c[k];
        }
        return out;
    }
    /// reference :3582-3585, the same expression (the product of the two counts is an unsigned 32-bit product there too)
    double npmi(const Pattern& key1, const Pattern& key2, int jointcount) {
        return log((double)jointcount / ((unsigned int)this->occurrencecount(key1) * (unsigned int)this->occurrencecount(key2))) / -log((double)jointcount / (double)totaloccurrences());
    }
    uint64_t totaloccurrences() const {  // totaloccurrencesingroup(0, 0)
        uint64_t t = 0;
        for (uint32_t c : counts_) t += c;
        return t;
    }
    /// computeflexgrams_fromcooc (reference :3751-3774) over the patterns the model holds when it is called: P {**} Q for every right co-occurrence
    /// whose npmi passes; every match of getrightcooc(P) adds its reference to each of P's flexgrams.  Returns the number of flexgrams found.
    int computeflexgrams_fromcooc(double threshold) {
        ensure_rindex();
        load_cooc(0);
        const uint64_t total = totaloccurrences();
        std::vector<std::string>                 newkeys;
        std::vector<std::vector<IndexReference>> newrefs;
        size_t i = 0;
        while (i < cooc_[0].size()) {
            size_t j = i;
            while (j < cooc_[0].size() && cooc_[0][j].p == cooc_[0][i].p) ++j;
            const uint32_t p = cooc_[0][i].p;
            std::vector<uint32_t> passing;
            for (size_t k = i; k < j; ++k) {
                const double v = log((double)cooc_[0][k].joint / (counts_[p] * counts_[cooc_[0][k].q])) / -log((double)cooc_[0][k].joint / (double)total);
                if (v >= threshold) passing.push_back(cooc_[0][k].q);
            }
            if (!passing.empty()) {
                // the matches of P: per occurrence, one per (position right of the pattern) x (pattern that starts at the occurrence)
                const uint64_t nocc = ref_off_[p + 1] - ref_off_[p];
                std::vector<uint32_t> rows(nocc * rlengths_.size(), 0);
                if (nocc && colibri_b200_rindex_query(rindex_, ref_sentence_.data() + ref_off_[p], ref_token_.data() + ref_off_[p], nocc, rows.data()) != COLIBRI_OK)
                    colibri_b200_detail::fail(colibri_b200_last_error());
                const unsigned int n = pattern_at(p).n();
                std::vector<IndexReference> plist;
                for (uint64_t o = 0; o < nocc; ++o) {
                    const uint32_t s = ref_sentence_[ref_off_[p] + o];
                    const uint16_t t = ref_token_[ref_off_[p] + o];
                    const int64_t  sl = (int64_t)sent_start_[s] - 1 - (int64_t)sent_start_[s - 1];
                    const int64_t  w  = sl - 1 - ((int64_t)t + n);
                    uint64_t here = 0;
                    for (size_t k = 0; k < rlengths_.size(); ++k) here += rows[o * rlengths_.size() + k] != 0;
                    for (int64_t r = 0; w > 0 && r < w * (int64_t)here; ++r) plist.push_back(IndexReference(s, t));
                }
                for (uint32_t q : passing) {
                    std::string k(reinterpret_cast<const char*>(keys_.data() + off_[p]), off_[p + 1] - off_[p]);
                    k.push_back((char)Pattern::flexclass);
                    k.append(reinterpret_cast<const char*>(keys_.data() + off_[q]), off_[q + 1] - off_[q]);
                    newkeys.push_back(k);
                    newrefs.push_back(plist);
                }
            }
            i = j;
        }
        if (newkeys.empty()) return 0;
        const size_t np0 = counts_.size();
        keys_.resize(off_[np0]);
        ref_sentence_.resize(ref_off_[np0]);
        ref_token_.resize(ref_off_[np0]);
        for (size_t f = 0; f < newkeys.size(); ++f) {
            keys_.insert(keys_.end(), newkeys[f].begin(), newkeys[f].end());
            off_.push_back(keys_.size());
            counts_.push_back((uint32_t)newrefs[f].size());
            for (const auto& r : newrefs[f]) {
                ref_sentence_.push_back(r.sentence);
                ref_token_.push_back(r.token);
            }
            ref_off_.push_back(ref_sentence_.size());
        }
        keys_.push_back(0);
        this->map_.clear();
        this->map_ready_ = false;
        this->hasflexgrams = true;
        drop_rindex();
        return (int)newkeys.size();
    }

  private:
    struct Rel { uint32_t p, q; uint64_t joint; };
    colibri_b200_rindex* rindex_ = nullptr;
    colibri_b200_model*  rmodel_ = nullptr;
    colibri_b200_corpus* rcorpus_ = nullptr;
    std::vector<uint32_t> rlengths_, sent_start_;
    std::vector<Rel>      cooc_[2];
    bool                  cooc_ready_[2] = {false, false};

    Pattern pattern_at(uint64_t i) const { return Pattern(keys_.data() + off_[i], (int)(off_[i + 1] - off_[i])); }
    void drop_rindex() {
        colibri_b200_rindex_free(rindex_);
        colibri_b200_model_free(rmodel_);
        colibri_b200_corpus_free(rcorpus_);
        rindex_ = nullptr; rmodel_ = nullptr; rcorpus_ = nullptr;
        cooc_ready_[0] = cooc_ready_[1] = false;
    }
    void ensure_rindex() {
        using colibri_b200_detail::fail;
        if (rindex_ != nullptr) return;
        if (this->reverseindex == nullptr) {
            std::cerr << "ERROR: No reverse index present" << std::endl;
            throw InternalError();
        }
        const int      dev  = colibri_b200_detail::default_device();
        const uint64_t zero = 0;
        const bool     any  = !counts_.empty();
        if (colibri_b200_model_from_flat(keys_.data(), any ? off_.data() : &zero, counts_.data(), counts_.size(), ref_sentence_.data(), ref_token_.data(), any ? ref_off_.data() : &zero,
                                         totaltokens, totaltypes, INDEXEDPATTERNMODEL, dev, &rmodel_) != COLIBRI_OK)
            fail(colibri_b200_last_error());
        if (colibri_b200_corpus_stage(this->reverseindex->beginpointer(), this->reverseindex->bytesize(), dev, &rcorpus_) != COLIBRI_OK) fail(colibri_b200_last_error());
        if (colibri_b200_rindex_build(rmodel_, rcorpus_, 0, &rindex_) != COLIBRI_OK) fail(colibri_b200_last_error());
        uint32_t n = 0, buf[256];
        if (colibri_b200_rindex_lengths(rindex_, buf, 256, &n) != COLIBRI_OK) fail(colibri_b200_last_error());
        rlengths_.assign(buf, buf + n);
        uint64_t info[4];
        if (colibri_b200_rindex_info(rindex_, info) != COLIBRI_OK) fail(colibri_b200_last_error());
        sent_start_.assign(info[0] + 1, 0);
        if (colibri_b200_rindex_sentence_starts(rindex_, sent_start_.data(), sent_start_.size()) != COLIBRI_OK) fail(colibri_b200_last_error());
    }
    void load_cooc(int direction) {
        using colibri_b200_detail::fail;
        ensure_rindex();
        if (cooc_ready_[direction]) return;
        uint64_t n = 0;
        if (colibri_b200_rindex_cooc(rindex_, direction, nullptr, nullptr, nullptr, 0, &n) != COLIBRI_OK) fail(colibri_b200_last_error());
        std::vector<uint32_t> p(n + 1), q(n + 1);
        std::vector<uint64_t> j(n + 1);
        if (n && colibri_b200_rindex_cooc(rindex_, direction, p.data(), q.data(), j.data(), n, &n) != COLIBRI_OK) fail(colibri_b200_last_error());
        cooc_[direction].resize(n);
        for (uint64_t i = 0; i < n; ++i) cooc_[direction][i] = Rel{p[i], q[i], j[i]};
        std::sort(cooc_[direction].begin(), cooc_[direction].end(), [](const Rel& a, const Rel& b) { return a.p < b.p || (a.p == b.p && a.q < b.q); });
        cooc_ready_[direction] = true;
    }
    // the pattern's index in the flat arrays (counts_.size() if it is not there)
    uint64_t index_of(const Pattern& pattern) const {
        for (uint64_t i = 0; i < counts_.size(); ++i)
            if (off_[i + 1] - off_[i] == pattern.bytesize() && memcmp(keys_.data() + off_[i], pattern.data(), pattern.bytesize()) == 0) return i;
        return counts_.size();
    }
    t_relationmap cooc_of(const Pattern& pattern, int direction, unsigned int occurrencethreshold, int category, unsigned int size) {
        load_cooc(direction);
        if (!this->has(pattern)) throw NoSuchPattern();
        const uint64_t idx = index_of(pattern);
        t_relationmap out;
        auto lo = std::lower_bound(cooc_[direction].begin(), cooc_[direction].end(), (uint32_t)idx, [](const Rel& a, uint32_t v) { return a.p < v; });
        for (; lo != cooc_[direction].end() && lo->p == idx; ++lo) {
            const Pattern q = pattern_at(lo->q);
            if (occurrencethreshold && (counts_[lo->q] < occurrencethreshold || lo->joint < occurrencethreshold)) continue;
            if (category && (int)q.category() != category) continue;
            if (size && q.n() != size) continue;
            out[q] = lo->joint;
        }
        return out;
    }
};
