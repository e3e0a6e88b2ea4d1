// colibri-patternmodeller (B200 build) -- the training front end of the reference's CLI (reference src/patternmodeller.cpp)
// for the options that lead into PatternModel::train: -f -o -u -t -l -m -b -s -y -T -W (same letters, same meaning,
// reference src/patternmodeller.cpp:504-618; Appendix C of SURVEY.md).  Model views and queries (-P -R -H -Q ...) and the
// constrained / continued / pointer-model modes belong to the reference's CPU code and are refused here with exit code 2.
#include <getopt.h>

#include <chrono>
#include <cstdlib>
#include <iostream>
#include <string>

#include "patternmodel.h"

static void usage() {
    std::cerr << "Usage: colibri-patternmodeller -f corpus.colibri.dat -o model.colibri.patternmodel [options]\n"
                 " Builds a pattern model on a B200 GPU.  Training options (as in Colibri Core):\n"
                 "  -f FILE   class-encoded corpus (*.colibri.dat, data version 2)\n"
                 "  -o FILE   output model file\n"
                 "  -u        build an unindexed model (default: indexed)\n"
                 "  -t N      occurrence threshold (default 2)\n"
                 "  -l N      maximum pattern length (default 100)\n"
                 "  -m N      minimum pattern length (default 1)\n"
                 "  -b N      maximum back-off length (default 100)\n"
                 "  -s        compute skipgrams (exhaustively for unindexed models)\n"
                 "  -y N      occurrence threshold for skipgrams (default: same as -t)\n"
                 "  -T N      skip type threshold (default 2)\n"
                 "  -W N      word occurrence threshold\n"
                 "  -q        quiet\n"
                 "  -d N      CUDA device ordinal (default 0)\n";
}

template <class ModelType>
static int run(const std::string& corpusfile, const std::string& outputmodelfile, IndexedCorpus* corpus, const PatternModelOptions& options, const std::string& qualifier) {
    ModelType model(corpus);
    std::cerr << "Training" << qualifier << " model on  " << corpusfile << std::endl;  // reference src/patternmodeller.cpp:318
    auto t0 = std::chrono::steady_clock::now();
    model.train(corpusfile, options, nullptr, nullptr, false, 1, false);
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!options.QUIET)
        std::cerr << "Trained in " << sec << " s: " << model.size() << " patterns, " << model.types() << " types, " << model.tokens() << " tokens (" << model.tokens() / sec / 1e6
                  << " M tokens/s)" << std::endl;
    if (!outputmodelfile.empty()) {
        std::cerr << "Writing model to " << outputmodelfile << std::endl;  // reference :389
        model.write(outputmodelfile);
    }
    return 0;
}

int main(int argc, char** argv) {
    std::string         corpusfile, outputmodelfile;
    PatternModelOptions options;
    bool                unindexed = false;
    int                 device    = 0;
    int                 c;
    while ((c = getopt(argc, argv, "hf:o:ut:l:m:b:sy:T:W:qd:c:i:j:PRHQDrgGF:S:xXNIVC:Y:L2Zvp:Ee:0M")) != -1) {
        switch (c) {
            case 'f': corpusfile = optarg; break;
            case 'o': outputmodelfile = optarg; break;
            case 'u': unindexed = true; break;
            case 't': options.MINTOKENS = atoi(optarg); break;
            case 'l': options.MAXLENGTH = atoi(optarg); break;
            case 'm': options.MINLENGTH = atoi(optarg); break;
            case 'b': options.MAXBACKOFFLENGTH = atoi(optarg); break;
            case 's': options.DOSKIPGRAMS = true; break;
            case 'y': options.MINTOKENS_SKIPGRAMS = atoi(optarg); break;
            case 'T': options.MINSKIPTYPES = atoi(optarg); break;
            case 'W': options.MINTOKENS_UNIGRAMS = atoi(optarg); break;
            case 'q': options.QUIET = true; break;
            case 'd': device = atoi(optarg); break;
            case 'h': usage(); return 0;
            default:
                std::cerr << "ERROR: option -" << (char)(c == '?' ? optopt : c)
                          << " is not part of the B200 training front end (model views, queries, constrained/continued training and pointer models "
                             "stay with the reference's CPU build)"
                          << std::endl;
                return 2;
        }
    }
    if (corpusfile.empty()) {
        usage();
        return 2;
    }
    if (outputmodelfile.empty()) {
        // reference src/patternmodeller.cpp:296-301
        std::cerr << "Ooops... You didn't really give me anything to do...that can't be right.. Did you perhaps forget --outputmodel?" << std::endl;
        return 2;
    }
    {
        std::ifstream probe(corpusfile, std::ios::in | std::ios::binary);
        if (!probe.good()) {
            std::cerr << "Can't open corpus data: " << corpusfile << std::endl;  // reference :749-751, exit 2
            return 2;
        }
    }
    colibri_b200_detail::default_device() = device;
    try {
        if (unindexed) {
            // reference :721-737: an unindexed model streams the corpus, unless skipgrams are wanted -- then they are computed
            // exhaustively from a preloaded corpus
            if (options.DOSKIPGRAMS) {
                std::cerr << "NOTE: Skipgram generation on unindexed pattern models can only be done exhaustively!" << std::endl;
                options.DOSKIPGRAMS_EXHAUSTIVE = true;
                options.DOSKIPGRAMS            = false;
                std::cerr << "Loading corpus data..." << std::endl;
                IndexedCorpus corpus(corpusfile);
                return run<PatternModel<uint32_t>>(corpusfile, outputmodelfile, &corpus, options, " unindexed");
            }
            return run<PatternModel<uint32_t>>(corpusfile, outputmodelfile, nullptr, options, " unindexed");
        }
        std::cerr << "Loading corpus data..." << std::endl;
        IndexedCorpus corpus(corpusfile);
        return run<IndexedPatternModel<>>(corpusfile, outputmodelfile, &corpus, options, "");
    } catch (const std::exception& e) {
        std::cerr << "FATAL: " << e.what() << std::endl;
        return 1;
    }
}
