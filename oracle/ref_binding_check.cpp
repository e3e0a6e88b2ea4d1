// oracle/ref_binding_check.cpp -- TEST INFRASTRUCTURE, not product code.
//
// One process, the UNMODIFIED reference (headers and library from /root/reference, compiled by oracle/Makefile into oracle/_ref/) twice:
//   A  PatternModel<uint32_t>::train              the reference's own CPU loop (include/patternmodel.h:880-1345)
//   B  B200PatternModel::train                    the same class with the one override of examples/reference_binding/b200_patternmodel.h,
//                                                 i.e. the C ABI of this repository on the GPU, filling the reference's own PatternMap
// and a comparison of the two models with the reference's own accessors: size(), tokens(), types(), maxlength(), minlength(), and
// occurrencecount() of every pattern of A in B and of B in A.  Prints "IDENTICAL ..." and exits 0, or the first differences and exits 1.
//
//   -i: the same with IndexedPatternModel<> / B200IndexedPatternModel, comparing the (sentence, token) list of every pattern
//
// usage: ref_binding_check -f corpus.colibri.dat [-t N] [-l N] [-s (exhaustive skipgrams)] [-p (preloaded corpus instead of the stream)]
//                          [-i (indexed models) [-S (trainskipgrams)]]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "b200_patternmodel.h"

static int indexed_check(const std::string& corpusfile, const PatternModelOptions& options) {
    std::ifstream  f(corpusfile, std::ifstream::in | std::ifstream::binary);
    IndexedCorpus* corpus = new IndexedCorpus(f, false);
    IndexedPatternModel<> a(corpus);
    a.train(corpusfile, options);
    B200IndexedPatternModel b(corpus);
    try {
        b.train(corpusfile, options);
    } catch (const InternalError&) {
        printf("B200 train() failed\n");
        return 3;
    }
    int bad = 0;
    if (a.size() != b.size() || a.tokens() != b.tokens() || a.types() != b.types() || a.maxlength() != b.maxlength() || a.minlength() != b.minlength()) {
        printf("DIFFERENT header: reference %llu patterns / %llu tokens / %llu types / %d..%d, B200 binding %llu / %llu / %llu / %d..%d\n", (unsigned long long)a.size(),
               (unsigned long long)a.tokens(), (unsigned long long)a.types(), a.minlength(), a.maxlength(), (unsigned long long)b.size(), (unsigned long long)b.tokens(),
               (unsigned long long)b.types(), b.minlength(), b.maxlength());
        ++bad;
    }
    unsigned long long refs = 0;
    for (IndexedPatternModel<>::iterator it = a.begin(); it != a.end() && bad < 10; ++it) {
        const Pattern p = it->first;
        IndexedData*  x = a.getdata(p);
        IndexedData*  y = b.getdata(p);
        if (y == NULL || x->data.size() != y->data.size()) {
            printf("DIFFERENT pattern of the reference model: %u occurrences there, %u in the binding's\n", (unsigned)x->data.size(), (unsigned)(y ? y->data.size() : 0));
            ++bad;
            continue;
        }
        for (size_t j = 0; j < x->data.size(); ++j)
            if (x->data[j].sentence != y->data[j].sentence || x->data[j].token != y->data[j].token) {
                printf("DIFFERENT reference %zu of a pattern: (%u, %u) there, (%u, %u) in the binding's\n", j, x->data[j].sentence, (unsigned)x->data[j].token, y->data[j].sentence,
                       (unsigned)y->data[j].token);
                ++bad;
                break;
            }
        refs += x->data.size();
    }
    if (bad) return 1;
    printf("IDENTICAL indexed patterns=%llu references=%llu tokens=%llu types=%llu\n", (unsigned long long)a.size(), refs, (unsigned long long)a.tokens(), (unsigned long long)a.types());
    return 0;
}

int main(int argc, char** argv) {
    std::string         corpusfile;
    PatternModelOptions options;
    options.MINTOKENS = 2;
    options.MAXLENGTH = 5;
    options.QUIET     = true;
    bool preloaded    = false;
    bool indexed      = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-f" && i + 1 < argc) corpusfile = argv[++i];
        else if (a == "-t" && i + 1 < argc) options.MINTOKENS = atoi(argv[++i]);
        else if (a == "-l" && i + 1 < argc) options.MAXLENGTH = atoi(argv[++i]);
        else if (a == "-s") options.DOSKIPGRAMS_EXHAUSTIVE = true;
        else if (a == "-p") preloaded = true;
        else if (a == "-i") indexed = true;
        else if (a == "-S") options.DOSKIPGRAMS = true;
        else {
            std::cerr << "unknown argument " << a << std::endl;
            return 2;
        }
    }
    if (indexed) return indexed_check(corpusfile, options);
    IndexedCorpus* corpus = NULL;
    if (preloaded || options.DOSKIPGRAMS_EXHAUSTIVE) {  // (the CLI preloads the corpus for exhaustive skipgrams, src/patternmodeller.cpp:721-737)
        std::ifstream f(corpusfile, std::ifstream::in | std::ifstream::binary);
        corpus = new IndexedCorpus(f, false);
    }
    PatternModel<uint32_t> a(corpus);
    a.train(corpusfile, options);
    B200PatternModel b(corpus);
    try {
        b.train(corpusfile, options);
    } catch (const InternalError&) {
        printf("B200 train() failed\n");
        return 3;
    }
    int bad = 0;
    auto differ = [&](const char* what, unsigned long long x, unsigned long long y) {
        if (x != y) {
            printf("DIFFERENT %s: reference %llu, B200 binding %llu\n", what, x, y);
            ++bad;
        }
    };
    differ("size()", a.size(), b.size());
    differ("tokens()", a.tokens(), b.tokens());
    differ("types()", a.types(), b.types());
    differ("maxlength()", a.maxlength(), b.maxlength());
    differ("minlength()", a.minlength(), b.minlength());
    differ("hasskipgrams", a.hasskipgrams, b.hasskipgrams);
    for (PatternModel<uint32_t>::iterator it = a.begin(); it != a.end() && bad < 10; ++it) {
        const Pattern p = it->first;
        if (!b.has(p) || b.occurrencecount(p) != a.occurrencecount(p)) {
            printf("DIFFERENT pattern of the reference model: count %u there, %u in the binding's\n", (unsigned)a.occurrencecount(p), (unsigned)(b.has(p) ? b.occurrencecount(p) : 0));
            ++bad;
        }
    }
    for (PatternModel<uint32_t>::iterator it = b.begin(); it != b.end() && bad < 10; ++it) {
        const Pattern p = it->first;
        if (!a.has(p)) {
            printf("DIFFERENT pattern only in the binding's model (count %u)\n", (unsigned)b.occurrencecount(p));
            ++bad;
        }
    }
    if (bad) return 1;
    printf("IDENTICAL patterns=%llu tokens=%llu types=%llu maxlength=%d\n", (unsigned long long)a.size(), (unsigned long long)a.tokens(), (unsigned long long)a.types(), a.maxlength());
    return 0;
}
