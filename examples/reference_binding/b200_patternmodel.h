// b200_patternmodel.h -- the binding a Colibri Core maintainer adds, as code that compiles against the UNMODIFIED
// reference headers (-I <colibri-core>/include) and the C ABI of this repository (-I include, -lcolibri_b200).
//
// B200PatternModel IS the reference's PatternModel<uint32_t>, B200IndexedPatternModel its IndexedPatternModel<> -- their own PatternMap, write(),
// has(), occurrencecount(), iteration -- with ONE override: train() hands the corpus bytes to the library and fills the model's map from the flat
// result (counts, or the sorted (sentence, token) lists of an indexed model).  Outside the accelerated
// subset (constraint model, filter, continued training) it calls the reference's own train().  Nothing of the reference is copied here.
//
// oracle/ref_binding_check.cpp trains the same corpus through PatternModel<uint32_t>::train (reference, CPU) and through this override (GPU)
// in one process and compares the two models entry by entry (tests/test_reference_binding_gpu.py).
#pragma once

#include <iostream>
#include <iterator>
#include <vector>

#include "patternmodel.h"  // the reference's include/patternmodel.h

extern "C" {
#include "colibri_b200.h"
}

template <class Base, bool kIndexed>
class B200TrainOverride : public Base {
  public:
    explicit B200TrainOverride(IndexedCorpus* corpus = NULL) : Base(corpus) {}

    using Base::train;  // the filename overload (:1353-1364) opens the file and calls the stream version below

    void train(std::istream* in, const PatternModelOptions& options, PatternModelInterface* constrainbymodel = NULL, PatternSet<>* filter = NULL, bool continued = false,
               uint32_t firstsentence = 1, bool ignoreerrors = false) override {
        if (constrainbymodel != NULL || (filter != NULL && filter->size() > 0) || continued || firstsentence != 1 || options.DOPATTERNPERLINE) {
            Base::train(in, options, constrainbymodel, filter, continued, firstsentence, ignoreerrors);  // the reference's own loop
            return;
        }
        // the corpus bytes after the 2-byte header: from the preloaded corpus if the model has one (:1030-1037), else from the stream
        std::vector<unsigned char> streamed;
        const unsigned char*       body   = NULL;
        size_t                     nbytes = 0;
        if (this->reverseindex != NULL) {
            body   = this->reverseindex->beginpointer();
            nbytes = this->reverseindex->bytesize();
        } else {
            if (in == NULL || !in->good()) {
                std::cerr << "ERROR: Supplied data file can not be opened. Check whether it exists and whether you have proper permissions..." << std::endl;
                throw InternalError();
            }
            in->clear();
            in->seekg(0);
            streamed.assign((std::istreambuf_iterator<char>(*in)), std::istreambuf_iterator<char>());
            if (streamed.size() < 2 || streamed[0] != 0xA2 || streamed[1] != 2) {
                Base::train(in, options, constrainbymodel, filter, continued, firstsentence, ignoreerrors);  // old corpus format: not on the device
                return;
            }
            body   = streamed.data() + 2;
            nbytes = streamed.size() - 2;
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
        o.PRUNENONSUBSUMED       = options.PRUNENONSUBSUMED;
        o.PRUNESUBSUMED          = options.PRUNESUBSUMED;
        o.QUIET                  = true;
        o.model_type             = this->getmodeltype();          // 10 unindexed, 20 indexed
        o.streamed               = this->reverseindex == NULL;    // Pattern(istream) vs IndexedCorpus as the sentence source
        colibri_b200_model* m = NULL;
        if (colibri_b200_train(body, nbytes, &o, &m) != COLIBRI_OK) {
            std::cerr << "ERROR: " << colibri_b200_last_error() << std::endl;  // the reference's error convention (include/common.h:41-44)
            throw InternalError();
        }
        this->totaltokens  = colibri_b200_model_tokens(m);
        this->totaltypes   = colibri_b200_model_types(m);
        this->maxn         = colibri_b200_model_maxn(m);
        this->minn         = colibri_b200_model_minn(m);
        this->hasskipgrams = colibri_b200_model_hasskipgrams(m) != 0;
        uint64_t np = 0, kb = 0, nr = 0;
        colibri_b200_model_export_sizes(m, &np, &kb, &nr);
        std::vector<uint8_t>  keys(kb + 1);
        std::vector<uint64_t> off(np + 1), ref_off(kIndexed ? np + 1 : 1);
        std::vector<uint32_t> counts(np + 1), ref_sentence(kIndexed ? nr + 1 : 1);
        std::vector<uint16_t> ref_token(kIndexed ? nr + 1 : 1);
        if (colibri_b200_model_export(m, keys.data(), off.data(), counts.data(), kIndexed ? ref_sentence.data() : NULL, kIndexed ? ref_token.data() : NULL,
                                      kIndexed ? ref_off.data() : NULL) != COLIBRI_OK) {
            std::cerr << "ERROR: " << colibri_b200_last_error() << std::endl;
            colibri_b200_model_free(m);
            throw InternalError();
        }
        colibri_b200_model_free(m);
        this->reserve(np);
        for (uint64_t i = 0; i < np; ++i) {
            const Pattern pattern(keys.data() + off[i], (int)(off[i + 1] - off[i]));
            fill((*this)[pattern], i, counts, ref_sentence, ref_token, ref_off);  // PatternMap::operator[] (patternstore.h:968-973)
        }
    }

  private:
    static void fill(uint32_t& value, uint64_t i, const std::vector<uint32_t>& counts, const std::vector<uint32_t>&, const std::vector<uint16_t>&, const std::vector<uint64_t>&) {
        value = counts[i];
    }
    static void fill(IndexedData& value, uint64_t i, const std::vector<uint32_t>&, const std::vector<uint32_t>& ref_sentence, const std::vector<uint16_t>& ref_token,
                     const std::vector<uint64_t>& ref_off) {
        value.data.reserve(ref_off[i + 1] - ref_off[i]);
        for (uint64_t j = ref_off[i]; j < ref_off[i + 1]; ++j) value.data.push_back(IndexReference(ref_sentence[j], ref_token[j]));  // sorted, as after posttrain (:2699-2705)
    }
};

typedef B200TrainOverride<PatternModel<uint32_t>, false> B200PatternModel;
typedef B200TrainOverride<IndexedPatternModel<>, true>   B200IndexedPatternModel;
