// Handle types shared by the two C wrappers of the oracle/_ref build (ref_c.cpp, ref_match_c.cpp). Test infrastructure only.
#pragma once
namespace ORB_SLAM2 { class ORBextractor; }
struct ref_extractor {
    ORB_SLAM2::ORBextractor* ex;
    int nlevels;
    int w0 = 0, h0 = 0;
};
