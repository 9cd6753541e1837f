// coeb_oracle.cpp -- extractor driver of the CPU oracle. TEST INFRASTRUCTURE ONLY (see the header).
#include "coeb_oracle.hpp"

#include <chrono>

namespace orc {

static inline double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ORBextractor::operator() (src/ORBextractor.cc:1088-1342) with ComputeKeyPointsOctTree (:771-904)
// inlined. Debug drawing / imshow / waitKey (:1214-1288) is not part of the algorithm and is dropped.
int Extractor::extract(const uint8_t* gray, int w, int h, int stride, const float* boxes, int nbox,
                       const float* tm, int ntm, const int* blur_flag, int nblur,
                       std::vector<coeb_keypoint>& kps, std::vector<uint8_t>& desc) {
    kps.clear();
    desc.clear();
    if (!gray || w <= 0 || h <= 0) return COEB_OK;  // `if (_image.empty()) return;` (:1096)
    const double t0 = now_s();
    if (!classify_boxes(w, h, boxes, nbox, tm, ntm, blur_flag, nblur)) return COEB_ERR_BAD_BOX;
    const bool area_flag = dyn.area_flag != 0;

    compute_pyramid(gray, w, h, stride);
    const double t1 = now_s();
    times.pyramid += t1 - t0;

    // threshold override (:775-784)
    if (area_flag) { iniThFAST = 30; minThFAST = 10; }
    else { iniThFAST = 20; minThFAST = 7; }

    candidates.assign(nlevels, {});
    level_keys.assign(nlevels, {});
    for (int level = 0; level < nlevels; level++) {
        const double ta = now_s();
        const Image& im = pyramid[level];
        const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
        const int maxBorderX = im.w - EDGE_THRESHOLD + 3, maxBorderY = im.h - EDGE_THRESHOLD + 3;
        std::vector<Cand>& cand = candidates[level];
        detect_level(level, iniThFAST, minThFAST, cand);
        if (area_flag) {  // CheckMovingKeyPoints before the octree (:854-858, :1410-1450)
            std::vector<Cand> keep;
            keep.reserve(cand.size());
            for (const Cand& c : cand)
                if (!is_moving(c.x, c.y, level, w, h)) keep.push_back(c);
            cand.swap(keep);
        }
        const double tb = now_s();
        times.fast += tb - ta;

        int N = mnFeaturesPerLevel[level];
        if (area_flag) N = (int)((int)(mnFeaturesPerLevel[level]) * 0.7);  // :869
        std::vector<int> sel = octree_distribute(cand, minBorderX, maxBorderX, minBorderY, maxBorderY, N);
        const int scaledPatchSize = (int)(PATCH_SIZE * mvScaleFactor[level]);  // :877
        std::vector<coeb_keypoint>& keys = level_keys[level];
        keys.reserve(sel.size());
        for (int idx : sel) {
            coeb_keypoint k;
            k.x = cand[idx].x + minBorderX;
            k.y = cand[idx].y + minBorderY;
            k.size = (float)scaledPatchSize;
            k.angle = -1;
            k.response = cand[idx].response;
            k.octave = level;
            k.class_id = -1;
            keys.push_back(k);
        }
        times.octree += now_s() - tb;
    }
    const double t2 = now_s();
    for (int level = 0; level < nlevels; level++)  // computeOrientation (:902-903)
        for (coeb_keypoint& k : level_keys[level]) k.angle = ic_angle(pyramid[level], k.x, k.y, umax);
    if (!area_flag) {  // CheckMovingKeyPoints_finall (:1204-1207, :1371-1408)
        for (int level = 0; level < nlevels && level < 8; level++) {
            std::vector<coeb_keypoint> keep;
            for (const coeb_keypoint& k : level_keys[level])
                if (!is_moving(k.x, k.y, level, w, h)) keep.push_back(k);
            level_keys[level].swap(keep);
        }
    }
    const double t3 = now_s();
    times.angle += t3 - t2;

    size_t total = 0;
    for (int level = 0; level < nlevels; level++) total += level_keys[level].size();
    kps.reserve(total);
    desc.resize(total * 32);
    blurred.resize(nlevels);
    size_t offset = 0;
    for (int level = 0; level < nlevels; level++) {  // :1308-1337
        std::vector<coeb_keypoint>& keys = level_keys[level];
        if (keys.empty()) { blurred[level].alloc(0, 0); continue; }
        const double ta = now_s();
        const Image& im = pyramid[level];
        blurred[level].alloc(im.w, im.h);
        gaussian7x7_8u(im.px.data(), im.w, im.h, im.w, blurred[level].px.data(), im.w);
        const double tb = now_s();
        times.blur += tb - ta;
        for (size_t i = 0; i < keys.size(); i++)
            orb_descriptor(blurred[level], keys[i].x, keys[i].y, keys[i].angle, &desc[(offset + i) * 32]);
        offset += keys.size();
        const float scale = mvScaleFactor[level];
        for (const coeb_keypoint& k0 : keys) {
            coeb_keypoint k = k0;
            if (level != 0) { k.x *= scale; k.y *= scale; }
            kps.push_back(k);
        }
        times.desc += now_s() - tb;
    }
    times.total += now_s() - t0;
    times.frames++;
    return COEB_OK;
}

}  // namespace orc
