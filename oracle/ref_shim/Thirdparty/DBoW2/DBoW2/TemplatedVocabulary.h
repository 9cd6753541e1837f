// Test-only stand-in for DBoW2's TemplatedVocabulary.h. The vocabulary tree itself is out of scope (SURVEY.md
// section 2: feature vectors are an INPUT of the BoW matchers); transform() and score() abort if reached.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <utility>
#include "BowVector.h"
#include "FeatureVector.h"
// The DBoW2 copy ORB-SLAM2 ships opens namespace std in this header, and the reference's own headers rely on it
// (include/Frame.h:70 names a bare `string`, include/Map.h:62 a bare `vector`).
using namespace std;
namespace DBoW2 {
template <class TDescriptor, class F> class TemplatedVocabulary {
public:
    void transform(const std::vector<TDescriptor>&, BowVector&, FeatureVector&, int) const {
        std::fprintf(stderr, "ref_shim: DBoW2 vocabulary transform is outside the hot path\n");
        std::abort();
    }
    double score(const BowVector&, const BowVector&) const {
        std::fprintf(stderr, "ref_shim: DBoW2 vocabulary score is outside the hot path\n");
        std::abort();
    }
};
}  // namespace DBoW2
