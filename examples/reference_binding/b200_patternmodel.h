// b200_patternmodel.h -- the binding a Colibri Core maintainer adds, as code that compiles against the UNMODIFIED
// reference headers (-I <colibri-core>/include) and the C ABI of this repository (-I include, -lcolibri_b200).
//
// B200PatternModel IS the reference's PatternModel<uint32_t> -- its own PatternMap, its own write(), has(), occurrencecount(), iteration --
// with ONE override: train() hands the corpus bytes to the library and fills the model's map from the flat result.  Outside the accelerated
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

class B200PatternModel : public PatternModel<uint32_t> {
  public:
    explicit B200PatternModel(IndexedCorpus* corpus = NULL) : PatternModel<uint32_t>(corpus) {}

    using PatternModel<uint32_t>::train;  // the filename overload (:1353-1364) opens the file and calls the stream version below

    void train(std::istream* in, const PatternModelOptions& options, PatternModelInterface* constrainbymodel = NULL, PatternSet<>* filter = NULL, bool continued = false,
               uint32_t firstsentence = 1, bool ignoreerrors = false) override {
        if (constrainbymodel != NULL || (filter != NULL && filter->size() > 0) || continued || firstsentence != 1 || options.DOPATTERNPERLINE) {
            PatternModel<uint32_t>::train(in, options, constrainbymodel, filter, continued, firstsentence, ignoreerrors);  // the reference's own loop
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
                PatternModel<uint32_t>::train(in, options, constrainbymodel, filter, continued, firstsentence, ignoreerrors);  // old corpus format: not on the device
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
        o.model_type             = this->getmodeltype();          // 10: unindexed
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
        std::vector<uint64_t> off(np + 1);
        std::vector<uint32_t> counts(np + 1);
        if (colibri_b200_model_export(m, keys.data(), off.data(), counts.data(), NULL, NULL, NULL) != COLIBRI_OK) {
            std::cerr << "ERROR: " << colibri_b200_last_error() << std::endl;
            colibri_b200_model_free(m);
            throw InternalError();
        }
        colibri_b200_model_free(m);
        this->reserve(np);
        for (uint64_t i = 0; i < np; ++i) (*this)[Pattern(keys.data() + off[i], (int)(off[i + 1] - off[i]))] = counts[i];  // PatternMap::operator[] (patternstore.h:968-973)
    }
};
