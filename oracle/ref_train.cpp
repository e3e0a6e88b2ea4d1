// oracle/ref_train.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A tiny driver around the UNMODIFIED reference (compiled from /root/reference by
// oracle/Makefile into oracle/_ref/).  It calls the reference's own
// PatternModel<uint32_t>::train / IndexedPatternModel<>::train
// (include/patternmodel.h:880, :1353, :2837) with the option set the
// colibri-patternmodeller CLI would build (src/patternmodeller.cpp:504-618,
// :721-754), times ONLY the train() call with steady_clock, writes the model
// file with the reference's own write() and prints one JSON line on stdout.
// The reference's per-pass progress lines go to stderr unchanged.
//
// usage: ref_train --spooky HEX...   |   ref_train --masks N MAXSKIPS   |
//        ref_train -f corpus.colibri.dat [-o model] [-u] [-s] [-t N] [-l N] [-m N]
//                  [-b N] [-y N] [-T N] [-W N] [-q]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>

#include "SpookyV2.h"
#include "algorithms.h"
#include "patternmodel.h"

template <class ModelType>
static int run(const std::string& corpusfile, const std::string& outfile, PatternModelOptions& options, bool loadcorpus, bool indexed) {
    IndexedCorpus* corpus = NULL;
    double         load_s = 0.0;
    if (loadcorpus) {
        auto          t0 = std::chrono::steady_clock::now();
        std::ifstream f(corpusfile, std::ifstream::in | std::ifstream::binary);
        if (!f.good()) {
            std::cerr << "Can't open corpus data: " << corpusfile << std::endl;
            return 2;
        }
        corpus = new IndexedCorpus(f, options.DEBUG);
        load_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    ModelType* model = new ModelType(corpus);
    auto       t0    = std::chrono::steady_clock::now();
    model->train(corpusfile, options, NULL, NULL, false, 1, false);
    if (indexed && options.DOSKIPGRAMS && !model->hasskipgrams) {
        // src/patternmodeller.cpp:326-337: the CLI runs trainskipgrams() after train() for indexed models
        model->trainskipgrams(options);
    }
    double train_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double write_s = 0.0;
    if (!outfile.empty()) {
        auto t1 = std::chrono::steady_clock::now();
        model->write(outfile);
        write_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    }
    printf("{\"train_seconds\": %.6f, \"corpus_load_seconds\": %.6f, \"write_seconds\": %.6f, \"tokens\": %llu, \"types\": %llu, "
           "\"patterns\": %llu, \"maxn\": %d, \"minn\": %d, \"hasskipgrams\": %d}\n",
           train_s, load_s, write_s, (unsigned long long)model->tokens(), (unsigned long long)model->types(), (unsigned long long)model->size(), model->maxlength(),
           model->minlength(), (int)model->hasskipgrams);
    fflush(stdout);
    delete model;
    if (corpus != NULL)
        delete corpus;
    return 0;
}

int main(int argc, char** argv) {
    std::string         corpusfile, outfile;
    bool                unindexed = false;
    PatternModelOptions options;
    if (argc >= 3 && std::string(argv[1]) == "--spooky") {
        // known-answer mode: SpookyHash::Hash64 (include/SpookyV2.h:59-66) of each hex-encoded message, one per line
        for (int i = 2; i < argc; ++i) {
            std::string   hex = argv[i];
            unsigned char buf[512];
            size_t        n = 0;
            for (size_t k = 0; k + 1 < hex.size() && n < sizeof buf; k += 2)
                buf[n++] = (unsigned char)strtoul(hex.substr(k, 2).c_str(), NULL, 16);
            printf("%llu\n", (unsigned long long)SpookyHash::Hash64((const void*)buf, n, 0));
        }
        return 0;
    }
    if (argc >= 4 && std::string(argv[1]) == "--masks") {
        // known-answer mode: compute_skip_configurations(n, maxskips) (src/algorithms.cpp:79-94)
        std::vector<uint32_t> masks = compute_skip_configurations(atoi(argv[2]), atoi(argv[3]));
        for (auto m : masks)
            printf("%u\n", m);
        return 0;
    }
    for (int i = 1; i < argc; ++i) {
        std::string a    = argv[i];
        auto        next = [&]() -> const char* {
            if (i + 1 >= argc) {
                std::cerr << "missing value for " << a << std::endl;
                exit(2);
            }
            return argv[++i];
        };
        if (a == "-f")
            corpusfile = next();
        else if (a == "-o")
            outfile = next();
        else if (a == "-u")
            unindexed = true;
        else if (a == "-s")
            options.DOSKIPGRAMS = true;
        else if (a == "-q")
            options.QUIET = true;
        else if (a == "-t")
            options.MINTOKENS = atoi(next());
        else if (a == "-l")
            options.MAXLENGTH = atoi(next());
        else if (a == "-m")
            options.MINLENGTH = atoi(next());
        else if (a == "-b")
            options.MAXBACKOFFLENGTH = atoi(next());
        else if (a == "-y")
            options.MINTOKENS_SKIPGRAMS = atoi(next());
        else if (a == "-T")
            options.MINSKIPTYPES = atoi(next());
        else if (a == "-W")
            options.MINTOKENS_UNIGRAMS = atoi(next());
        else {
            std::cerr << "unknown argument " << a << std::endl;
            return 2;
        }
    }
    if (corpusfile.empty()) {
        std::cerr << "usage: ref_train -f corpus.colibri.dat [-o model] [-u] [-s] [-t N] [-l N] ..." << std::endl;
        return 2;
    }
    bool loadcorpus = true;
    if (unindexed) {
        // src/patternmodeller.cpp:721-737
        if (options.DOSKIPGRAMS) {
            options.DOSKIPGRAMS_EXHAUSTIVE = true;
            options.DOSKIPGRAMS            = false;
        } else {
            loadcorpus = false;
        }
        return run<PatternModel<uint32_t>>(corpusfile, outfile, options, loadcorpus, false);
    }
    return run<IndexedPatternModel<>>(corpusfile, outfile, options, loadcorpus, true);
}
