// colibri-patternmodeller (B200 build) -- the training front end of the reference's CLI (reference src/patternmodeller.cpp)
// for the options that lead into PatternModel::train: -f -o -u -t -l -m -b -s -y -T -W, and the model-driven modes around it:
// -i (input model), -j (constraint model), -I (constrained in-place rebuild), -2 (two-stage building) -- same letters, same
// meaning, reference src/patternmodeller.cpp:504-618, :620-660, :700-721, :777-852; Appendix C of SURVEY.md.  Model views and
// queries (-P -R -H -Q ...), continued / expanded training and pointer models belong to the reference's CPU code and are
// refused here with exit code 2.
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
                 "  -F S      compute flexgrams by abstracting from the skipgrams (implies -s; indexed models)\n"
                 "  -i FILE   input model (with -I: the model to rebuild on the corpus; alone with -o: load, filter by the options, write)\n"
                 "  -j FILE   constraint model: only patterns that occur in it are counted\n"
                 "  -I        constrained in-place rebuild of the input model (-i) on the corpus (-f)\n"
                 "  -2        two-stage building: an unindexed model first, then an indexed model constrained by it\n"
                 "  -q        quiet\n"
                 "  -d N      CUDA device ordinal (default 0); a list or range (-d 0-7, -d 0,2,5) shards the corpus over several GPUs\n"
                 "            (unindexed n-gram models; the model is hash-partitioned over the GPUs while it is built)\n";
}

static bool g_flexfromskip = false;  // -F S / -S S: abstract flexgrams from the skipgrams after training (reference src/patternmodeller.cpp:330-337)

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
    if (g_flexfromskip) {
        if (model.getmodeltype() != INDEXEDPATTERNMODEL) {
            std::cerr << "WARNING: Can't compute flexgrams from skipgrams on unindexed model" << std::endl;  // reference :327-329
        } else {
            std::cerr << "Computing flexgrams from skipgrams" << corpusfile << std::endl;  // reference :333-336 (sic)
            int found = model.computeflexgrams_fromskipgrams();
            std::cerr << found << " flexgrams found" << corpusfile << std::endl;
        }
    }
    if (!outputmodelfile.empty()) {
        std::cerr << "Writing model to " << outputmodelfile << std::endl;  // reference :389
        model.write(outputmodelfile);
    }
    return 0;
}

// -I (reference src/patternmodeller.cpp:777-852): load the input model AS the output type with DORESET, widen the length window to the
// model's, then train on the corpus constrained by the model itself
template <class ModelType>
static int rebuild_inplace(const std::string& inputmodelfile, const std::string& corpusfile, const std::string& outputmodelfile, PatternSetModel* constrainbymodel,
                           IndexedCorpus* corpus, PatternModelOptions options, const char* kind) {
    std::cerr << "Loading model " << inputmodelfile << " as " << kind << " pattern model..." << std::endl;
    PatternModelOptions optionscopy = options;
    optionscopy.DORESET             = true;
    ModelType model(inputmodelfile, optionscopy, constrainbymodel ? constrainbymodel->getinterface() : nullptr, corpus);
    std::cerr << "(" << model.size() << " patterns" << ")" << std::endl;
    if (model.maxlength() > options.MAXLENGTH) options.MAXLENGTH = model.maxlength();
    if (model.minlength() < options.MINLENGTH) options.MINLENGTH = model.minlength();
    std::cerr << "Building new " << kind << " model from  " << corpusfile << std::endl;
    model.train(corpusfile, options, model.getinterface(), nullptr, false, 1, false);
    if (!outputmodelfile.empty()) model.write(outputmodelfile);
    return 0;
}

// -i without -I (reference processmodel, :364-368, :389-392): load with the options as filters, write
template <class ModelType>
static int convert(const std::string& inputmodelfile, const std::string& outputmodelfile, PatternSetModel* constrainbymodel, const PatternModelOptions& options,
                   const std::string& qualifier) {
    std::cerr << "Loading pattern model " << inputmodelfile << " as" << qualifier << " model..." << std::endl;
    ModelType model(inputmodelfile, options, constrainbymodel ? constrainbymodel->getinterface() : nullptr, nullptr);
    std::cerr << "Writing model to " << outputmodelfile << std::endl;
    model.write(outputmodelfile);
    return 0;
}

template <class ModelType>
static int run_constrained(const std::string& corpusfile, const std::string& outputmodelfile, IndexedCorpus* corpus, const PatternModelOptions& options, PatternSetModel* constrainbymodel,
                           const std::string& qualifier) {
    ModelType model(corpus);
    std::cerr << "Training" << qualifier << " model on  " << corpusfile << std::endl;
    model.train(corpusfile, options, constrainbymodel->getinterface(), nullptr, false, 1, false);
    std::cerr << "Unloading constraint model" << std::endl;
    if (!outputmodelfile.empty()) {
        std::cerr << "Writing model to " << outputmodelfile << std::endl;
        model.write(outputmodelfile);
    }
    return 0;
}

static bool file_exists(const std::string& filename) {
    std::ifstream testf(filename);
    return testf.good();
}

int main(int argc, char** argv) {
    std::string         corpusfile, outputmodelfile, inputmodelfile, inputmodelfile2;
    PatternModelOptions options;
    bool                unindexed = false, inplace = false, twostage = false;
    int                 device    = 0;
    std::string         devicespec;
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
            case 'd': devicespec = optarg; break;
            case 'c':  // the class file only matters to the views (decoding patterns for print/report); training never opens it (reference :506, :857-865)
                if (!options.QUIET) std::cerr << "Note: class file " << optarg << " is not needed to build a model; ignored" << std::endl;
                break;
            case 'D': options.DEBUG = true; break;  // reference :511
            case 'i': inputmodelfile = optarg; break;
            case 'j': inputmodelfile2 = optarg; break;
            case 'S':
            case 'F':
                if (std::string(optarg) == "S") {  // reference src/patternmodeller.cpp:550-554
                    g_flexfromskip      = true;
                    options.DOSKIPGRAMS = true;
                } else {
                    std::cerr << "ERROR: flexgrams from co-occurrence (-F <threshold>) are not part of the B200 training front end" << std::endl;
                    return 2;
                }
                break;
            case 'I': inplace = true; break;
            case '2': twostage = true; break;
            case 'h': usage(); return 0;
            default:
                std::cerr << "ERROR: option -" << (char)(c == '?' ? optopt : c)
                          << " is not part of the B200 training front end (model views, queries, continued/expanded training and pointer models "
                             "stay with the reference's CPU build)"
                          << std::endl;
                return 2;
        }
    }
    if (!devicespec.empty()) {  // N | A-B | A,B,C
        std::vector<int> devs;
        size_t           i = 0;
        while (i < devicespec.size()) {
            size_t j = devicespec.find(',', i);
            if (j == std::string::npos) j = devicespec.size();
            const std::string part = devicespec.substr(i, j - i);
            const size_t      dash = part.find('-');
            if (dash != std::string::npos && dash > 0) {
                for (int d = atoi(part.substr(0, dash).c_str()); d <= atoi(part.substr(dash + 1).c_str()); ++d) devs.push_back(d);
            } else if (!part.empty()) {
                devs.push_back(atoi(part.c_str()));
            }
            i = j + 1;
        }
        if (devs.empty()) {
            std::cerr << "ERROR: -d expects a device ordinal, a list (0,1,2) or a range (0-7)" << std::endl;
            return 2;
        }
        device = devs[0];
        if (devs.size() > 1) colibri_b200_detail::device_list() = devs;
    }
    colibri_b200_detail::default_device() = device;
    int stages = 1;
    const std::string cached_outputmodelfile = outputmodelfile;
    const bool        cached_DOSKIPGRAMS     = options.DOSKIPGRAMS;
    if (twostage) {  // reference :630-646
        if (options.MINTOKENS == 1) {
            std::cerr << "Two stage building was requested but has no value with --threshold 1 , disabling..." << std::endl;
            twostage = false;
        } else {
            stages = 2;
            if (outputmodelfile.empty()) {
                std::cerr << "ERROR: An output model file (--outputmodel) is mandatory for two-stage building!" << std::endl;
                return 2;
            }
        }
    }
    try {
        for (int stage = 1; stage <= stages; ++stage) {
            if (twostage) {  // reference :650-671
                if (stage == 1) {
                    std::cerr << "********* STARTING STAGE 1/2: Building intermediary unindexed patternmodel ******" << std::endl;
                    inplace             = false;
                    outputmodelfile     = cached_outputmodelfile + ".stage1";
                    unindexed           = true;
                    options.DOSKIPGRAMS = false;
                } else {
                    std::cerr << "********* STARTING STAGE 2/2: Building indexed patternmodel ******" << std::endl;
                    inplace             = true;
                    outputmodelfile     = cached_outputmodelfile;
                    unindexed           = false;
                    options.DOSKIPGRAMS = cached_DOSKIPGRAMS;
                    inputmodelfile      = outputmodelfile + ".stage1";
                    inputmodelfile2     = "";
                }
            }
            if (inputmodelfile.empty() && corpusfile.empty()) {
                if (argc <= 1) {
                    usage();
                    return 2;
                }
                std::cerr << "ERROR: No input model (--inputmodel) or corpus data file specified (--datafile|-f), specify at least one." << std::endl;  // reference :675-685
                return 2;
            }
            if (outputmodelfile.empty()) {
                // reference src/patternmodeller.cpp:296-301
                std::cerr << "Ooops... You didn't really give me anything to do...that can't be right.. Did you perhaps forget --outputmodel?" << std::endl;
                return 2;
            }
            if (!inputmodelfile.empty() && !file_exists(inputmodelfile)) {
                std::cerr << "No such file: " << inputmodelfile << std::endl;  // reference assert_file_exists, :233-239
                return 2;
            }
            if (!corpusfile.empty() && !file_exists(corpusfile)) {
                std::cerr << "Can't open corpus data: " << corpusfile << std::endl;  // reference :749-751, exit 2
                return 2;
            }
            if ((inplace || !inputmodelfile2.empty()) && options.DOSKIPGRAMS) {
                std::cerr << "ERROR: skipgrams under a constraint model (-s with -j / -I / -2) are not part of the B200 training front end" << std::endl;
                return 2;
            }
            PatternSetModel* constrainbymodel = nullptr;
            if (!inputmodelfile2.empty()) {  // reference :713-721
                if (!file_exists(inputmodelfile2)) {
                    std::cerr << "No such file: " << inputmodelfile2 << std::endl;
                    return 2;
                }
                std::cerr << "Loading constraint model (aka training/intersection model)" << std::endl;
                constrainbymodel = new PatternSetModel(inputmodelfile2, options);
                std::cerr << " (Contains " << constrainbymodel->size() << " patterns)" << std::endl;
            }
            struct Drop {
                PatternSetModel*& p;
                ~Drop() { delete p; p = nullptr; }
            } drop{constrainbymodel};
            const std::string qualifier = unindexed ? " unindexed" : "";
            int rc = 0;
            if (inplace) {
                std::cerr << "Constrained in-place rebuild (--constrained|-I) enabled, on " << corpusfile << std::endl;
                if (corpusfile.empty() || inputmodelfile.empty()) {
                    std::cerr << "ERROR: Corpus data file (--datafile|-f) must be specified when --constrained|-I is set!." << std::endl;  // reference :781-786
                    return 2;
                }
                std::cerr << "Loading corpus data..." << std::endl;  // -I keeps the corpus preloaded (reference :728-737, :745-754)
                IndexedCorpus corpus(corpusfile);
                rc = unindexed ? rebuild_inplace<PatternModel<uint32_t>>(inputmodelfile, corpusfile, outputmodelfile, constrainbymodel, &corpus, options, "unindexed")
                               : rebuild_inplace<IndexedPatternModel<>>(inputmodelfile, corpusfile, outputmodelfile, constrainbymodel, &corpus, options, "indexed");
            } else if (!inputmodelfile.empty()) {
                if (!corpusfile.empty()) {
                    std::cerr << "ERROR: expanding / continuing an input model on a corpus (-i with -f, without -I) is not part of the B200 training front end" << std::endl;
                    return 2;
                }
                rc = unindexed ? convert<PatternModel<uint32_t>>(inputmodelfile, outputmodelfile, constrainbymodel, options, qualifier)
                               : convert<IndexedPatternModel<>>(inputmodelfile, outputmodelfile, constrainbymodel, options, qualifier);
            } else if (unindexed) {
                // reference :721-737: an unindexed model streams the corpus, unless skipgrams are wanted -- then they are computed
                // exhaustively from a preloaded corpus
                if (options.DOSKIPGRAMS) {
                    std::cerr << "NOTE: Skipgram generation on unindexed pattern models can only be done exhaustively!" << std::endl;
                    options.DOSKIPGRAMS_EXHAUSTIVE = true;
                    options.DOSKIPGRAMS            = false;
                    std::cerr << "Loading corpus data..." << std::endl;
                    IndexedCorpus corpus(corpusfile);
                    rc = run<PatternModel<uint32_t>>(corpusfile, outputmodelfile, &corpus, options, " unindexed");
                } else if (constrainbymodel) {
                    rc = run_constrained<PatternModel<uint32_t>>(corpusfile, outputmodelfile, nullptr, options, constrainbymodel, " unindexed");
                } else {
                    rc = run<PatternModel<uint32_t>>(corpusfile, outputmodelfile, nullptr, options, " unindexed");
                }
            } else {
                std::cerr << "Loading corpus data..." << std::endl;
                IndexedCorpus corpus(corpusfile);
                rc = constrainbymodel ? run_constrained<IndexedPatternModel<>>(corpusfile, outputmodelfile, &corpus, options, constrainbymodel, "")
                                      : run<IndexedPatternModel<>>(corpusfile, outputmodelfile, &corpus, options, "");
            }
            if (rc) return rc;
            if (!options.QUIET) {  // (B200 build only: where the host side spent its time)
                const colibri_b200_detail::HostTimes& ht = colibri_b200_detail::host_times();
                std::cerr << "Host timing: read files " << ht.read << " s, device calls (staging + training) " << ht.device << " s, result to host " << ht.adopt
                          << " s, map view " << ht.materialise << " s, write " << ht.write << " s" << std::endl;
            }
        }
    } catch (const std::exception& e) {
        std::cerr << "FATAL: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
