// Test-only stand-in for DBoW2's BowVector.h (the reference's Thirdparty/DBoW2 copy is not in /root/reference).
// DBoW2 publishes BowVector as a std::map<WordId, WordValue>; the hot path only carries it around.
#pragma once
#include <map>
#include <vector>
namespace DBoW2 {
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
class BowVector : public std::map<WordId, WordValue> {
public:
    void addWeight(WordId id, WordValue v) { (*this)[id] += v; }
    void addIfNotExist(WordId id, WordValue v) { insert(std::make_pair(id, v)); }
};
}  // namespace DBoW2
