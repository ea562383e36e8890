// ORACLE (test infrastructure): C entry points around the reference's OWN vocabulary code -- DBoW2's
// TemplatedVocabulary<FORB::TDescriptor, FORB> (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h, FORB.cpp, ScoringObject.cpp,
// BowVector.cpp, FeatureVector.cpp) compiled VERBATIM from where it lies against oracle/cvshim_m.  Recipe: oracle/Makefile
// -> oracle/_ref/libfbe_refvoc.so.  The vocabulary is read with the reference's own loadFromTextFile from a text file the
// test writes (the shipped ORBvoc.txt is not part of the reference checkout).
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include "opencv2/core/core.hpp"
#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> Voc;

extern "C" {

void* refv_load_text(const char* path) {
    Voc* v = new Voc();
    if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
    return v;
}
void refv_free(void* h) { delete static_cast<Voc*>(h); }
int refv_size(void* h) { return (int)static_cast<Voc*>(h)->size(); }

// TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup) (TemplatedVocabulary.h:1127-1205), the call
// Frame::ComputeBoW / KeyFrame::ComputeBoW make.  Outputs: the bag-of-words vector as (word id, value) in map order and the
// feature vector as CSR over ascending node ids.  Returns the number of words; *n_nodes = number of nodes.
int refv_transform(void* h, const uint8_t* desc, int n, int levelsup, int32_t* bow_ids, double* bow_vals, int32_t* fv_ids,
                   int32_t* fv_start, int32_t* fv_items, int32_t* n_nodes) {
    std::vector<cv::Mat> feats(n);
    for (int i = 0; i < n; ++i) { feats[i].create(1, 32, CV_8U); std::memcpy(feats[i].ptr(0), desc + (size_t)i * 32, 32); }
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    static_cast<Voc*>(h)->transform(feats, bv, fv, levelsup);
    int k = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++k) { bow_ids[k] = (int32_t)it->first; bow_vals[k] = it->second; }
    int a = 0, p = 0;
    fv_start[0] = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++a) {
        fv_ids[a] = (int32_t)it->first;
        for (size_t j = 0; j < it->second.size(); ++j) fv_items[p++] = (int32_t)it->second[j];
        fv_start[a + 1] = p;
    }
    *n_nodes = a;
    return k;
}

}  // extern "C"
