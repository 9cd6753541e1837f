// Test-only stand-in for DBoW2's FORB.h: only the descriptor typedef include/ORBVocabulary.h needs.
#pragma once
#include <opencv2/core/core.hpp>
namespace DBoW2 {
class FORB {
public:
    typedef cv::Mat TDescriptor;
    typedef const TDescriptor* pDescriptor;
    static const int L = 32;
};
}  // namespace DBoW2
