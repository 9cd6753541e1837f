// coeb_oracle_c.cpp -- C entry points of the CPU oracle for ctypes (tests/, smoke(), bench.py's CPU
// arm). TEST INFRASTRUCTURE ONLY: the product library never links or loads this.
#include <atomic>
#include <chrono>
#include <memory>
#include <thread>

#include "coeb_oracle.hpp"
#include "coeb_oracle_match.hpp"
#include "coeb_oracle_frame.hpp"

using namespace orc;

extern "C" {

// ---- OpenCV primitive models ---------------------------------------------------------------
void orc_resize_linear(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh,
                       int dstride) {
    resize_linear_8u(src, sw, sh, sstride, dst, dw, dh, dstride);
}
void orc_gaussian7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    gaussian7x7_8u(src, w, h, sstride, dst, dstride);
}
float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }
void orc_rgb_to_gray(const uint8_t* src, int w, int h, int sstride, int channels, int bgr, uint8_t* dst, int dstride) {
    rgb_to_gray_8u(src, w, h, sstride, channels, bgr != 0, dst, dstride);
}
// blur_flag producer: means[b] = Laplacian mean of box b (-1 if invalid), flags[b] = mean in [0, 4.2)
void orc_blur_flags(const uint8_t* gray, int w, int h, int stride, const float* boxes, int nbox, int* flags, double* means) {
    for (int b = 0; b < nbox; b++) {
        const double m = laplacian_box_mean(gray, w, h, stride, boxes + 4 * b);
        if (means) means[b] = m;
        flags[b] = (m >= 0 && m < 4.2) ? 1 : 0;
    }
}
int orc_cv_round_f(float v) { return cv_round(v); }

// FAST-9/16 + NMS on a ROI; out is n x 3 int32 (x, y, score). Returns the count (may exceed cap).
int orc_fast_roi(const uint8_t* img, int w, int h, int stride, int threshold, int* out, int cap) {
    std::vector<FastPt> pts;
    fast9_nms(img, w, h, stride, threshold, pts);
    for (size_t i = 0; i < pts.size() && (int)i < cap; i++) {
        out[3 * i] = pts[i].x; out[3 * i + 1] = pts[i].y; out[3 * i + 2] = pts[i].score;
    }
    return (int)pts.size();
}

// DistributeOctTree on explicit candidates (x, y, response as float triples).
int orc_octree(const float* cand_xyr, int n, int minX, int maxX, int minY, int maxY, int N, int* out_idx,
               int cap) {
    std::vector<Cand> c(n);
    for (int i = 0; i < n; i++) c[i] = {cand_xyr[3 * i], cand_xyr[3 * i + 1], cand_xyr[3 * i + 2]};
    std::vector<int> sel = octree_distribute(c, minX, maxX, minY, maxY, N);
    for (size_t i = 0; i < sel.size() && (int)i < cap; i++) out_idx[i] = sel[i];
    return (int)sel.size();
}

// out4 = {octree calls, careful-phase rounds, rounds with an order tie, rounds with a cut tie} of this thread since the last reset
void orc_octree_tie_stats(long* out4, int reset) {
    if (out4) { out4[0] = g_tie_stats.calls; out4[1] = g_tie_stats.careful_rounds; out4[2] = g_tie_stats.rounds_with_order_tie; out4[3] = g_tie_stats.rounds_with_cut_tie; }
    if (reset) g_tie_stats = OctreeTieStats();
}

// ---- extractor -----------------------------------------------------------------------------
struct orc_extractor {
    Extractor ex;
    std::vector<coeb_keypoint> kps;
    std::vector<uint8_t> desc;
    int w0 = 0, h0 = 0;
    orc_extractor(const coeb_orb_params& p)
        : ex(p.nfeatures, p.scale_factor, p.nlevels, p.ini_th_fast, p.min_th_fast) {}
};

orc_extractor* orc_extractor_create(const coeb_orb_params* p) { return new orc_extractor(*p); }
void orc_extractor_destroy(orc_extractor* e) { delete e; }

void orc_extractor_tables(orc_extractor* e, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                          int* per_level, int* umax16) {
    for (int i = 0; i < e->ex.nlevels; i++) {
        if (scale) scale[i] = e->ex.mvScaleFactor[i];
        if (inv_scale) inv_scale[i] = e->ex.mvInvScaleFactor[i];
        if (sigma2) sigma2[i] = e->ex.mvLevelSigma2[i];
        if (inv_sigma2) inv_sigma2[i] = e->ex.mvInvLevelSigma2[i];
        if (per_level) per_level[i] = e->ex.mnFeaturesPerLevel[i];
    }
    if (umax16)
        for (int i = 0; i < 16; i++) umax16[i] = e->ex.umax[i];
}

int orc_extract(orc_extractor* e, const uint8_t* gray, int w, int h, int stride, const float* boxes, int nbox,
                const float* tm, int ntm, const int* blur_flag, int nblur, coeb_keypoint* kps_out,
                uint8_t* desc_out, int cap, int* n_out) {
    e->w0 = w; e->h0 = h;
    int st = e->ex.extract(gray, w, h, stride, boxes, nbox, tm, ntm, blur_flag, nblur, e->kps, e->desc);
    if (st != COEB_OK) { if (n_out) *n_out = 0; return st; }
    const int n = (int)e->kps.size();
    if (n_out) *n_out = n;
    if (n > cap) return COEB_ERR_CAPACITY;
    if (kps_out) std::memcpy(kps_out, e->kps.data(), sizeof(coeb_keypoint) * n);
    if (desc_out) std::memcpy(desc_out, e->desc.data(), (size_t)32 * n);
    return COEB_OK;
}

void orc_dyn_info(orc_extractor* e, coeb_dyn_info* out) { *out = e->ex.dyn; }

int orc_level_size(orc_extractor* e, int level, int* w, int* h) {
    if (level < 0 || level >= (int)e->ex.pyramid.size()) return -1;
    *w = e->ex.pyramid[level].w; *h = e->ex.pyramid[level].h;
    return 0;
}
// which: 0 = pyramid level, 1 = blurred level (empty if the level had no keypoints)
int orc_level_image(orc_extractor* e, int level, int which, uint8_t* dst) {
    const Image& im = which ? e->ex.blurred[level] : e->ex.pyramid[level];
    if (im.px.empty()) return 0;
    std::memcpy(dst, im.px.data(), im.px.size());
    return (int)im.px.size();
}
// FAST candidates handed to the octree, minBorder-relative (x, y, response) float triples.
int orc_level_candidates(orc_extractor* e, int level, float* out, int cap) {
    const auto& c = e->ex.candidates[level];
    for (size_t i = 0; i < c.size() && (int)i < cap; i++) {
        out[3 * i] = c[i].x; out[3 * i + 1] = c[i].y; out[3 * i + 2] = c[i].response;
    }
    return (int)c.size();
}
// Keypoints of one level after octree + angle (+ post cull), level coordinates.
int orc_level_keypoints(orc_extractor* e, int level, coeb_keypoint* out, int cap) {
    const auto& k = e->ex.level_keys[level];
    for (size_t i = 0; i < k.size() && (int)i < cap; i++) out[i] = k[i];
    return (int)k.size();
}

void orc_stage_times(orc_extractor* e, double* out7, long* frames) {
    const StageTimes& t = e->ex.times;
    out7[0] = t.pyramid; out7[1] = t.fast; out7[2] = t.octree; out7[3] = t.angle; out7[4] = t.blur;
    out7[5] = t.desc; out7[6] = t.total;
    *frames = t.frames;
}

// Multi-threaded batch extraction for the CPU baseline: one Extractor per thread (the reference
// class is not re-entrant), frames handed out dynamically. Inputs are packed per frame:
// gray [B][h][w]; boxes [B][max_box][4] with nbox[B]; tm [B][max_tm][2] with ntm[B];
// blur [B][max_box]. Outputs: counts[B], optional kps [B][cap], desc [B][cap][32].
// Returns wall seconds for the whole batch.
double orc_extract_batch_mt(const coeb_orb_params* p, int B, const uint8_t* gray, int w, int h, const float* boxes,
                            const int* nbox, int max_box, const float* tm, const int* ntm, int max_tm,
                            const int* blur, int nthreads, int* counts, coeb_keypoint* kps_out,
                            uint8_t* desc_out, int cap) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next(0);
    auto t0 = std::chrono::steady_clock::now();
    auto worker = [&]() {
        Extractor ex(p->nfeatures, p->scale_factor, p->nlevels, p->ini_th_fast, p->min_th_fast);
        std::vector<coeb_keypoint> kps;
        std::vector<uint8_t> desc;
        for (;;) {
            int f = next.fetch_add(1);
            if (f >= B) break;
            ex.extract(gray + (size_t)f * w * h, w, h, w, boxes ? boxes + (size_t)f * max_box * 4 : nullptr,
                       nbox ? nbox[f] : 0, tm ? tm + (size_t)f * max_tm * 2 : nullptr, ntm ? ntm[f] : 0,
                       blur ? blur + (size_t)f * max_box : nullptr, nbox ? nbox[f] : 0, kps, desc);
            int n = (int)kps.size();
            if (counts) counts[f] = n;
            int m = std::min(n, cap);
            if (kps_out) std::memcpy(kps_out + (size_t)f * cap, kps.data(), sizeof(coeb_keypoint) * m);
            if (desc_out) std::memcpy(desc_out + (size_t)f * cap * 32, desc.data(), (size_t)32 * m);
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; i++) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

// ---- grid + matchers -----------------------------------------------------------------------
int orc_hamming256(const uint8_t* a, const uint8_t* b) { return hamming256(a, b); }

struct orc_frame {
    FrameView v;
    std::vector<coeb_keypoint> kps;
    std::vector<uint8_t> desc;
    std::vector<float> uright, scale;
};

orc_frame* orc_frame_create(const coeb_keypoint* kps, const uint8_t* desc, int n, const float* uright,
                            const coeb_camera* cam, const float* scale, int nlevels) {
    orc_frame* f = new orc_frame();
    f->kps.assign(kps, kps + n);
    f->desc.assign(desc, desc + (size_t)n * 32);
    if (uright) f->uright.assign(uright, uright + n);
    f->scale.assign(scale, scale + nlevels);
    f->v.n = n;
    f->v.kps = f->kps.data();
    f->v.desc = f->desc.data();
    f->v.uright = uright ? f->uright.data() : nullptr;
    f->v.cam = *cam;
    f->v.scale = f->scale.data();
    f->v.nlevels = nlevels;
    f->v.build_grid();
    return f;
}
void orc_frame_destroy(orc_frame* f) { delete f; }

int orc_features_in_area(orc_frame* f, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap) {
    std::vector<int> v;
    f->v.features_in_area(x, y, r, minLevel, maxLevel, v);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = v[i];
    return (int)v.size();
}

int orc_grid_cell(orc_frame* f, int ix, int iy, int* out, int cap) {
    const auto& c = f->v.grid[ix][iy];
    for (size_t i = 0; i < c.size() && (int)i < cap; i++) out[i] = c[i];
    return (int)c.size();
}

int orc_match_projection(orc_frame* f, int n, const uint8_t* track_in_view, const uint8_t* bad,
                         const uint8_t* has_obs, const float* proj_x, const float* proj_y, const float* proj_xr,
                         const int* level, const float* view_cos, const uint8_t* desc, float th, float nnratio,
                         int* kp_match) {
    MapPointsSoA mp;
    mp.n = n; mp.track_in_view = track_in_view; mp.bad = bad; mp.has_obs = has_obs;
    mp.proj_x = proj_x; mp.proj_y = proj_y; mp.proj_xr = proj_xr; mp.level = level; mp.view_cos = view_cos;
    mp.desc = desc;
    return search_by_projection_map(f->v, mp, th, nnratio, kp_match);
}

int orc_match_lastframe(orc_frame* cur, int n, const uint8_t* valid, const uint8_t* has_obs, const float* xyz,
                        const int* octave, const float* angle, const uint8_t* desc, const float* Tcw_cur,
                        const float* Tcw_last, float th, int mono, int check_ori, int* kp_match) {
    LastFrameSoA L;
    L.n = n; L.valid = valid; L.has_obs = has_obs; L.xyz = xyz; L.octave = octave; L.angle = angle; L.desc = desc;
    return search_by_projection_last(cur->v, L, Tcw_cur, Tcw_last, th, mono != 0, check_ori != 0, kp_match);
}

int orc_match_init(orc_frame* f1, orc_frame* f2, float* prev_matched, int* matches12, int window, float nnratio,
                   int check_ori) {
    return search_for_initialization(f1->v, f2->v, prev_matched, matches12, window, nnratio, check_ori != 0);
}

// Stereo: the two pyramids are taken from two oracle extractors that have just run.
int orc_stereo_match(orc_extractor* exL, orc_extractor* exR, int N, const coeb_keypoint* keysL,
                     const uint8_t* descL, int Nr, const coeb_keypoint* keysR, const uint8_t* descR, float mbf,
                     float mb, float* uright, float* depth) {
    return compute_stereo_matches(N, keysL, descL, Nr, keysR, descR, exL->ex.pyramid, exR->ex.pyramid,
                                  exL->ex.mvScaleFactor.data(), exL->ex.mvInvScaleFactor.data(), mbf, mb, uright,
                                  depth);
}

int orc_knn2(int nq, const uint8_t* q, int nt, const uint8_t* t, float nnratio, int* best_idx, int* d1, int* d2) {
    return knn2_ratio(nq, q, nt, t, nnratio, best_idx, d1, d2);
}

// Multi-threaded kNN for the CPU baseline (queries split across threads).
double orc_knn2_mt(int nq, const uint8_t* q, int nt, const uint8_t* t, float nnratio, int* best_idx, int* d1,
                   int* d2, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    auto work = [&](int lo, int hi) {
        if (hi > lo) knn2_ratio(hi - lo, q + (size_t)lo * 32, nt, t, nnratio, best_idx + lo, d1 + lo, d2 + lo);
    };
    int per = (nq + nthreads - 1) / nthreads;
    for (int i = 1; i < nthreads; i++) th.emplace_back(work, std::min(nq, i * per), std::min(nq, (i + 1) * per));
    work(0, std::min(nq, per));
    for (auto& x : th) x.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- Frame constructor tail + SearchLocalPoints (coeb_oracle_frame.hpp) -----------------------------------
void orc_undistort_keypoints(const coeb_keypoint* keys, int n, const coeb_camera* cam, const float* dist5,
                             coeb_keypoint* keys_un) {
    undistort_keypoints(keys, n, *cam, dist5, keys_un);
}

void orc_stereo_from_rgbd(const coeb_keypoint* keys, const coeb_keypoint* keys_un, int n, const void* depth, int kind,
                          int stride_bytes, float factor, float mbf, float* uright, float* depth_out) {
    DepthView D;
    D.data = depth; D.kind = depth ? kind : 0; D.stride_bytes = stride_bytes; D.factor = factor;
    stereo_from_rgbd(keys, keys_un, n, D, mbf, uright, depth_out);
}

int orc_match_bow(orc_frame* f1, orc_frame* f2, const uint8_t* valid1, const uint8_t* valid2, int nn1, const int* node1,
                  const int* start1, const int* items1, int nn2, const int* node2, const int* start2, const int* items2,
                  float nnratio, int check_ori, int strict_low, int* match12) {
    FeatVecCSR V1, V2;
    V1.nn = nn1; V1.node = node1; V1.start = start1; V1.items = items1;
    V2.nn = nn2; V2.node = node2; V2.start = start2; V2.items = items2;
    return search_by_bow(f1->v, f2->v, valid1, valid2, V1, V2, nnratio, check_ori != 0, strict_low != 0, match12);
}

int orc_match_triangulation(orc_frame* f1, orc_frame* f2, const uint8_t* free1, const uint8_t* free2, int nn1, const int* node1,
                            const int* start1, const int* items1, int nn2, const int* node2, const int* start2,
                            const int* items2, const float* F12, float ex, float ey, int only_stereo, int check_ori,
                            int* match12) {
    FeatVecCSR V1, V2;
    V1.nn = nn1; V1.node = node1; V1.start = start1; V1.items = items1;
    V2.nn = nn2; V2.node = node2; V2.start = start2; V2.items = items2;
    return search_for_triangulation(f1->v, f2->v, free1, free2, V1, V2, F12, ex, ey, only_stereo != 0, check_ori != 0, match12);
}

int orc_match_reloc(orc_frame* cur, int n, const uint8_t* valid, const float* xyz, const float* min_dist, const float* max_dist,
                    const float* angle, const uint8_t* desc, const float* Tcw, const float* Ow, float th, int orb_dist, int check_ori,
                    int* kp_match) {
    return search_by_projection_reloc(cur->v, n, valid, xyz, min_dist, max_dist, angle, desc, Tcw, Ow, th, orb_dist, check_ori != 0, kp_match);
}

int orc_fuse_search(orc_frame* f, int n, const float* xyz, const float* normal, const float* min_dist, const float* max_dist,
                    const uint8_t* desc, const uint8_t* valid, const float* Tcw, const float* Ow, float th, int chi2_tests, int* best_idx) {
    LocalMapSoA M;
    M.n = n; M.xyz = xyz; M.normal = normal; M.min_dist = min_dist; M.max_dist = max_dist; M.desc = desc;
    return fuse_search(f->v, M, valid, Tcw, Ow, th, chi2_tests != 0, best_idx);
}

// Number of floats in [lo, hi) (bit patterns, stepped by `step`) whose restated logf differs from the C library's.
long orc_logf_mismatches(uint32_t lo, uint32_t hi, uint32_t step) {
    long bad = 0;
    for (uint64_t b = lo; b < hi; b += step) {
        const uint32_t bits = (uint32_t)b;
        float x;
        std::memcpy(&x, &bits, 4);
        const float a = std::log(x), m = glibc_logf(x);
        bad += std::memcmp(&a, &m, 4) != 0;
    }
    return bad;
}

int orc_search_local_points(orc_frame* f, int n, const float* xyz, const float* normal, const float* min_dist,
                            const float* max_dist, const uint8_t* desc, const uint8_t* skip, const uint8_t* has_obs,
                            const float* Tcw, const float* Ow, float cos_limit, float th, float nnratio, int* kp_match,
                            uint8_t* in_view, float* proj) {
    LocalMapSoA M;
    M.n = n; M.xyz = xyz; M.normal = normal; M.min_dist = min_dist; M.max_dist = max_dist; M.desc = desc;
    return search_local_points(f->v, M, skip, has_obs, Tcw, Ow, cos_limit, th, nnratio, kp_match, in_view, proj);
}

}  // extern "C"
