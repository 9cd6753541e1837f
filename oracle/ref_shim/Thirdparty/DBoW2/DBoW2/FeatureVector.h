// Test-only stand-in for DBoW2's FeatureVector.h: std::map<NodeId, std::vector<unsigned int>>, features appended in
// ascending index order (what TemplatedVocabulary::transform produces). The SearchByBoW / SearchForTriangulation loops of
// src/ORBmatcher.cc:159-288, 522-655, 657-824 only iterate it.
#pragma once
#include <map>
#include <vector>
#include "BowVector.h"
namespace DBoW2 {
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {
public:
    void addFeature(NodeId id, unsigned int i_feature) { (*this)[id].push_back(i_feature); }
};
}  // namespace DBoW2
