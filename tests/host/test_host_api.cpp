// CPU-side checks of the C++ API mirror (colibri-core_b200/host): the value types behave like the reference's
// (reference include/pattern.h, src/pattern.cpp) and train() refuses loudly without a GPU.  Prints "ok"/"FAILED" lines in
// the spirit of the reference's own src/test.cpp and exits 2 on the first failure.
#include <cstdlib>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "patternmodel.h"

static int testnr = 0;
template <class A, class B>
static void test(const char* what, const A& value, const B& ref) {
    ++testnr;
    if (value == (A)ref) {
        std::cerr << testnr << " " << what << " ... ok" << std::endl;
    } else {
        std::cerr << testnr << " " << what << " ... " << value << " ... FAILED, expected " << ref << std::endl;
        exit(2);
    }
}

int main(int argc, char** argv) {
    // Pattern basics; hash KATs produced by the reference library
    Pattern unigram = Pattern::fromclasses({6});
    Pattern bigram  = Pattern::fromclasses({6, 7});
    test("unigram n()", unigram.n(), 1u);
    test("bigram n()", bigram.n(), 2u);
    test("bigram bytesize()", bigram.bytesize(), 2u);
    test("Pattern::hash [06]", (unsigned long long)unigram.hash(), 6716366417670780154ull);
    test("Pattern::hash [06 07]", (unsigned long long)bigram.hash(), 7201352277971678815ull);
    const unsigned char multibyte[] = {0x86, 0x01, 0x07, 0x80, 0x80, 0x01};
    Pattern mb(multibyte, 6);
    test("multibyte n()", mb.n(), 3u);
    test("Pattern::hash multibyte", (unsigned long long)mb.hash(), 6910059722266967716ull);
    test("multibyte tovector[0]", mb.tovector()[0], 134u);
    test("multibyte tovector[2]", mb.tovector()[2], 16384u);
    test("fromclasses roundtrip", Pattern::fromclasses(mb.tovector()) == mb, true);
    test("empty hash", (unsigned long long)Pattern().hash(), 0ull);
    const unsigned char skip[] = {0x0A, 0x03, 0x03, 0x0D, 0x0E};
    Pattern sg(skip, 5);
    test("skipgram category", (int)sg.category(), (int)SKIPGRAM);
    test("skipgram isgap(1)", sg.isgap(1), true);
    test("skipgram isgap(0)", sg.isgap(0), false);
    test("skipgram n()", sg.n(), 5u);
    test("Pattern::hash skipgram", (unsigned long long)sg.hash(), 17915007264098384048ull);
    test("ngram category", (int)bigram.category(), (int)NGRAM);
    test("std::hash<Pattern>", (unsigned long long)std::hash<Pattern>()(bigram), 7201352277971678815ull);
    std::ostringstream os;
    bigram.write(os);
    test("write() appends the end marker", os.str().size(), 3u);

    // options defaults (reference include/patternmodel.h:153-180)
    PatternModelOptions o;
    test("MINTOKENS default", o.MINTOKENS, -1);
    test("MAXLENGTH default", o.MAXLENGTH, 100);
    test("MINSKIPTYPES default", o.MINSKIPTYPES, 2);
    test("MAXSKIPS default", o.MAXSKIPS, 3);
    test("DOSKIPGRAMS default", o.DOSKIPGRAMS, false);

    {  // the file readers take what is left of a stream, with one read() when the stream knows its size
        const std::string tmp = "/tmp/colibri_b200_host_test.bin";
        {
            std::ofstream out(tmp, std::ios::binary);
            for (int i = 0; i < 100000; ++i) out.put((char)(i * 7));
        }
        std::ifstream f(tmp, std::ios::binary);
        std::vector<unsigned char> all = colibri_b200_detail::read_all(f);
        test("read_all(file) size", all.size(), (size_t)100000);
        test("read_all(file) content", (int)all[3], 21);
        f.clear();
        f.seekg(10);
        all = colibri_b200_detail::read_all(f);
        test("read_all from a position", all.size(), (size_t)99990);
        test("read_all from a position: first byte", (int)all[0], 70);
        std::istringstream text(std::string("hello"));
        test("read_all(stringstream)", colibri_b200_detail::read_all(text).size(), (size_t)5);
        std::ifstream missing("/nonexistent/file", std::ios::binary);
        test("read_all(unopened stream)", colibri_b200_detail::read_all(missing).size(), (size_t)0);
        remove(tmp.c_str());
        PatternModel<uint32_t> empty;  // an empty model is its 27-byte header (reference include/patternmodel.h:1609-1624)
        std::ostringstream     blob;
        empty.write(blob);
        test("empty model file size", blob.str().size(), (size_t)27);
        test("empty model file type byte", (int)(unsigned char)blob.str()[1], (int)UNINDEXEDPATTERNMODEL);
    }
    if (argc > 1) {
        IndexedCorpus corpus{std::string(argv[1])};
        test("IndexedCorpus sentences (hamlet)", corpus.sentences(), 40u);  // reference src/test.cpp:1549
        if (colibri_b200_device_count() > 0) {
            // the training part of the reference's own test program (src/test.cpp:1207-1340) with its expected numbers; patterns are given by
            // their classes in tests/golden/hamlet.colibri.cls ("or not to" = 14 12 7, "give us fortune" = 156 27 135)
            {
                const std::string   infilename = argv[1];
                PatternModelOptions options;
                options.QUIET = true;
                PatternModel<uint32_t> unindexedmodelNSR(&corpus);  // from the preloaded corpus, no skipgrams (:1210-1219)
                unindexedmodelNSR.train(infilename, options);
                test("unindexed, preloaded: patterns", unindexedmodelNSR.size(), (size_t)111);
                test("unindexed, preloaded: types", unindexedmodelNSR.types(), (size_t)186);
                test("unindexed, preloaded: tokens", unindexedmodelNSR.tokens(), (size_t)354);
                PatternModel<uint32_t> unindexedmodelNS;  // streamed from the file (:1224-1233)
                unindexedmodelNS.train(infilename, options);
                test("unindexed, streamed: patterns", unindexedmodelNS.size(), (size_t)111);
                test("unindexed, streamed: types", unindexedmodelNS.types(), (size_t)186);
                test("unindexed, streamed: tokens", unindexedmodelNS.tokens(), (size_t)354);
                const Pattern ngram    = Pattern::fromclasses({14, 12, 7});
                const Pattern ngram_ne = Pattern::fromclasses({156, 27, 135});
                test("unindexedmodel.has(or not to)", unindexedmodelNS.has(ngram), true);             // :1241-1243
                test("!unindexedmodel.has(give us fortune)", unindexedmodelNS.has(ngram_ne), false);  // :1244-1246
                test("unindexedmodel.occurrencecount(or not to)", unindexedmodelNS.occurrencecount(ngram), (size_t)6);  // :1247-1248
                options.DOSKIPGRAMS_EXHAUSTIVE = true;  // :1262-1283
                options.DOSKIPGRAMS            = false;
                PatternModel<uint32_t> unindexedmodelR(&corpus);
                unindexedmodelR.train(infilename, options);
                test("unindexed + skipgrams, preloaded: patterns", unindexedmodelR.size(), (size_t)385);
                PatternModel<uint32_t> unindexedmodel;
                unindexedmodel.train(infilename, options);
                test("unindexed + skipgrams, streamed: patterns", unindexedmodel.size(), (size_t)385);
                test("unindexed + skipgrams: types", unindexedmodel.types(), (size_t)186);
                test("unindexed + skipgrams: tokens", unindexedmodel.tokens(), (size_t)354);
                test("unindexed + skipgrams: has", unindexedmodel.has(ngram), true);
                test("unindexed + skipgrams: occurrencecount", unindexedmodel.occurrencecount(ngram), (size_t)6);
                const std::string outputfilename = "/tmp/colibri_b200_host_test.colibri.patternmodel";  // write, read back (:1303-1322)
                unindexedmodel.write(outputfilename);
                PatternModel<uint32_t> unindexedmodel2(outputfilename, options);
                test("read back: equal tokens", unindexedmodel.tokens() == unindexedmodel2.tokens(), true);
                test("read back: equal types", unindexedmodel.types() == unindexedmodel2.types(), true);
                test("read back: equal size", unindexedmodel.size() == unindexedmodel2.size(), true);
                test("read back: has", unindexedmodel2.has(ngram), true);
                test("read back: occurrencecount", unindexedmodel2.occurrencecount(ngram), (size_t)6);
                remove(outputfilename.c_str());
                options.DOSKIPGRAMS_EXHAUSTIVE = false;  // indexed model with skipgrams (:1324-1340)
                options.DOSKIPGRAMS            = true;
                IndexedPatternModel<> indexedmodel(&corpus);
                indexedmodel.train(infilename, options);
                test("indexed + skipgrams: patterns", indexedmodel.size(), (size_t)133);
                test("indexed: equal tokens", unindexedmodel.tokens() == indexedmodel.tokens(), true);
                test("indexed: equal types", unindexedmodel.types() == indexedmodel.types(), true);
                test("indexed: has", indexedmodel.has(ngram), true);
                test("indexed: occurrencecount", indexedmodel.occurrencecount(ngram), (size_t)6);
                test("indexed: size = n-grams + skipgrams", indexedmodel.size(), unindexedmodelNS.size() + indexedmodel.totalpatternsingroup(SKIPGRAM, 0));  // :1329-1330
                test("indexed: unigram types", indexedmodel.totalwordtypesingroup(0, 1), 45u);              // :1344-1345
                test("unindexed: unigram types", unindexedmodelNSR.totalwordtypesingroup(0, 1), 45u);      // :1221-1222
            }
            // reverse index + co-occurrence relations of the indexed model (answers of the unmodified reference, oracle/_ref/ref_relations -t 2 -l 3)
            PatternModelOptions ro;
            ro.QUIET     = true;
            ro.MINTOKENS = 2;
            ro.MAXLENGTH = 3;
            IndexedPatternModel<> im(&corpus);
            im.train((std::istream*)nullptr, ro);
            test("indexed model patterns (hamlet -t 2 -l 3)", im.size(), (size_t)81);
            std::vector<Pattern> at = im.getreverseindex(IndexReference(1, 2));
            test("getreverseindex((1, 2)) patterns", at.size(), (size_t)3);  // 0e, 0e0c, 0e0c07
            test("getreverseindex outside the sentence", im.getreverseindex(IndexReference(1, 9999)).size(), (size_t)0);
            Pattern six = Pattern::fromclasses({6});
            auto    lc  = im.getleftcooc(six);
            test("getleftcooc([06]) relations", lc.size(), (size_t)4);
            test("getleftcooc([06])[06]", (unsigned long long)lc[six], 165ull);
            test("getleftcooc([06])[06 07]", (unsigned long long)lc[Pattern::fromclasses({6, 7})], 12ull);
            auto bc = im.getcooc(six);  // both directions, no overlap (reference :3543-3576; numbers of oracle/_ref/ref_relations)
            test("getcooc([06]) relations", bc.size(), (size_t)56);
            test("getcooc([06])[06]", (unsigned long long)bc[six], 14ull);
            test("getcooc([06], 2, 0, 0, ordersignificant) relations", im.getcooc(six, 2, 0, 0, true).size(), (size_t)24);
            const size_t before = im.size();
            int          found  = im.computeflexgrams_fromcooc(0.1);
            test("computeflexgrams_fromcooc(0.1) found", found, 43);  // the clean iteration; the reference's own run adds one flexgram of a flexgram
            test("model grew by the flexgrams", im.size(), before + 43);
        }
        // without a GPU train() must throw InternalError, never compute on the CPU
        if (colibri_b200_device_count() == 0) {
            PatternModel<uint32_t> model;
            bool threw = false;
            o.QUIET = true;
            try {
                model.train(std::string(argv[1]), o);
            } catch (const InternalError&) {
                threw = true;
            }
            test("train() without a GPU throws InternalError", threw, true);
            test("model stays empty", model.size(), 0u);
        }
        // refusals of the training front end
        PatternModel<uint32_t> model2;
        bool threw = false;
        try {
            model2.train(std::string(argv[1]), o, nullptr, nullptr, /*continued=*/true);
        } catch (const InternalError&) {
            threw = true;
        }
        test("continued training is refused", threw, true);
        // loading and constrained training keep the reference's signatures (include/patternmodel.h:700-726, :781-861, :880)
        threw = false;
        try {
            PatternModel<uint32_t> missing(std::string("/nonexistent.colibri.patternmodel"), o);
        } catch (const InternalError&) {
            threw = true;
        }
        test("loading a missing model file throws InternalError", threw, true);
        threw = false;
        try {
            IndexedPatternModel<> notamodel(std::string(argv[1]), o);  // a corpus file is not a model file (reference :788-793)
        } catch (const InternalError&) {
            threw = true;
        }
        test("loading a file that is not a model throws InternalError", threw, true);
        if (colibri_b200_device_count() == 0) {
            PatternSetModel        constraint;  // an empty set still has to go through the device
            PatternModel<uint32_t> model3;
            threw = false;
            try {
                model3.train(std::string(argv[1]), o, constraint.getinterface());
            } catch (const InternalError&) {
                threw = true;
            }
            test("constrained train() without a GPU throws InternalError", threw, true);
        }
    }
    std::cerr << "all " << testnr << " host API tests ok" << std::endl;
    return 0;
}
