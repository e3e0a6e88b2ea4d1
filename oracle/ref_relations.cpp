// oracle/ref_relations.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A driver around the UNMODIFIED reference (compiled from /root/reference by oracle/Makefile into oracle/_ref/): trains an
// IndexedPatternModel with a reverse index on a corpus and dumps what the reference itself answers for
//   G  getreverseindex(ref)            include/patternmodel.h:1746-1824   every position of the corpus
//   R  getrightcooc(pattern)           :3460-3493                          every pattern
//   L  getleftcooc(pattern)            :3502-3531
//   C  getcooc(pattern)                :3543-3576 (both directions, no overlap), defaults; O: the same with occurrencethreshold 2 and ordersignificant
//   N  npmi() :3582-3585 of every getrightcooc relation that passes the threshold (what computenpmi(map, th, true, false) :3671-3691 collects)
//   F/X computeflexgrams_fromcooc(th)  :3751-3774                          found, then every flexgram with its occurrence count
// as text lines (patterns as hex of their bytes), sorted, so that tests/golden/make_golden_relations.py can pin the oracle to them.
//
// usage: ref_relations -f corpus.colibri.dat [-t N] [-l N] [-Y npmi-threshold] [-x (also run computeflexgrams_fromcooc)] [-s (skipgrams)] [-S (group statistics only)]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "patternmodel.h"

static std::string hexof(const PatternPointer& p) {
    static const char* d = "0123456789abcdef";
    std::string        s;
    const Pattern      q(p);
    for (size_t i = 0; i < q.bytesize(); ++i) {
        s.push_back(d[q.data[i] >> 4]);
        s.push_back(d[q.data[i] & 15]);
    }
    return s;
}

int main(int argc, char** argv) {
    std::string         corpusfile;
    PatternModelOptions options;
    options.MINTOKENS = 2;
    options.MAXLENGTH = 5;
    options.QUIET     = true;
    double threshold  = 0.0;
    bool   doflex     = false;
    bool   statsonly  = false;  // -S: print the group statistics (computestats :1903-1933, computecoveragestats :1946-1984) and stop
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-f" && i + 1 < argc) corpusfile = argv[++i];
        else if (a == "-t" && i + 1 < argc) options.MINTOKENS = atoi(argv[++i]);
        else if (a == "-l" && i + 1 < argc) options.MAXLENGTH = atoi(argv[++i]);
        else if (a == "-Y" && i + 1 < argc) threshold = atof(argv[++i]);
        else if (a == "-x") doflex = true;
        else if (a == "-s") options.DOSKIPGRAMS = true;
        else if (a == "-S") statsonly = true;
        else {
            std::cerr << "unknown argument " << a << std::endl;
            return 2;
        }
    }
    std::ifstream f(corpusfile, std::ifstream::in | std::ifstream::binary);
    if (!f.good()) {
        std::cerr << "Can't open corpus data: " << corpusfile << std::endl;
        return 2;
    }
    IndexedCorpus* corpus = new IndexedCorpus(f, false);
    IndexedPatternModel<>* model = new IndexedPatternModel<>(corpus);
    model->train(corpusfile, options, NULL, NULL, false, 1, false);
    printf("H patterns=%llu tokens=%llu types=%llu sentences=%d total=%llu\n", (unsigned long long)model->size(), (unsigned long long)model->tokens(),
           (unsigned long long)model->types(), corpus->sentences(), (unsigned long long)model->totaloccurrencesingroup(0, 0));
    if (statsonly) {
        for (int c = 0; c <= 3; ++c)
            for (int n = 0; n <= model->maxlength(); ++n)
                printf("S %d %d %u %u %u\n", c, n, model->totaloccurrencesingroup(c, n), model->totalpatternsingroup(c, n), model->totalwordtypesingroup(c, n));
        return 0;
    }
    // snapshot of the patterns (the model must not be iterated while it changes)
    std::vector<Pattern> patterns;
    for (IndexedPatternModel<>::iterator it = model->begin(); it != model->end(); ++it) patterns.push_back(it->first);
    std::sort(patterns.begin(), patterns.end());

    std::vector<std::string> lines;
    for (int s = 1; s <= corpus->sentences(); ++s) {
        const int sl = corpus->sentencelength(s);
        for (int t = 0; t < sl; ++t) {
            std::unordered_set<PatternPointer> r = model->getreverseindex(IndexReference(s, t));
            std::vector<std::string> hs;
            for (const auto& p : r) hs.push_back(hexof(p));
            std::sort(hs.begin(), hs.end());
            std::string line = "G " + std::to_string(s) + " " + std::to_string(t);
            for (auto& h : hs) line += " " + h;
            lines.push_back(line);
        }
    }
    for (auto& l : lines) puts(l.c_str());
    lines.clear();
    for (const Pattern& p : patterns) {
        t_relationmap r = model->getrightcooc(p);
        for (const auto& kv : r) lines.push_back("R " + hexof(p) + " " + hexof(kv.first) + " " + std::to_string(kv.second));
        t_relationmap l = model->getleftcooc(p);
        for (const auto& kv : l) lines.push_back("L " + hexof(p) + " " + hexof(kv.first) + " " + std::to_string(kv.second));
        t_relationmap c = model->getcooc(p);
        for (const auto& kv : c) lines.push_back("C " + hexof(p) + " " + hexof(kv.first) + " " + std::to_string(kv.second));
        t_relationmap o = model->getcooc(p, 2, 0, 0, true);
        for (const auto& kv : o) lines.push_back("O " + hexof(p) + " " + hexof(kv.first) + " " + std::to_string(kv.second));
    }
    // computenpmi(map, threshold, right = true, left = false) keys its result map with PatternPointers into a loop-local Pattern (:3674, :3686: dangling
    // once the iteration moves on), so the map cannot be read back; its content is restated here with the reference's own getrightcooc() and npmi()
    for (const Pattern& p : patterns) {
        t_relationmap r = model->getrightcooc(p);
        for (const auto& kv : r) {
            const double value = model->npmi(p, kv.first, kv.second);
            if (value >= threshold) {
                char buf[64];
                snprintf(buf, sizeof buf, "%.17g", value);
                lines.push_back("N " + hexof(p) + " " + hexof(kv.first) + " " + buf);
            }
        }
    }
    std::sort(lines.begin(), lines.end());
    for (auto& l : lines) puts(l.c_str());
    if (doflex) {
        const size_t before = model->size();
        int          found  = model->computeflexgrams_fromcooc(threshold);
        printf("F found=%d size_before=%llu size_after=%llu\n", found, (unsigned long long)before, (unsigned long long)model->size());
        lines.clear();
        for (IndexedPatternModel<>::iterator it = model->begin(); it != model->end(); ++it) {
            const Pattern p = it->first;
            if (p.category() == FLEXGRAM) lines.push_back("X " + hexof(p) + " " + std::to_string(model->occurrencecount(p)));
        }
        std::sort(lines.begin(), lines.end());
        for (auto& l : lines) puts(l.c_str());
    }
    return 0;
}
