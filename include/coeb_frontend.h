/*
 * coeb_frontend.h -- C ABI of the B200-native COEB-SLAM front end (libcoeb_frontend.so).
 *
 * Drop-in boundary for the reference's data-parallel hot path. Every entry point names the reference
 * interface it replaces (paths relative to the COEB-SLAM tree). Plain pointers and sizes only; no
 * C++, OpenCV or torch types. All functions return COEB_OK (0) or a negative coeb_status;
 * coeb_last_error() gives the message for the calling thread. There is no CPU fallback: without an
 * sm_100 device creation fails with COEB_ERR_NO_DEVICE.
 *
 * Threading: a coeb_extractor / coeb_matcher is owned by one thread at a time, like the reference's
 * ORBextractor (non re-entrant: it mutates mvImagePyramid, include/ORBextractor.h:99). Use one
 * handle per stream / per GPU. Handles are cheap to re-create (COEB's frame-loss logic does
 * `new ORBextractor(nFeatures+500, ...)`, src/Tracking.cc:434-465): device arenas are pooled per device.
 */
#ifndef COEB_FRONTEND_H
#define COEB_FRONTEND_H

#include <stddef.h>
#include <stdint.h>

#include "coeb_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct coeb_extractor coeb_extractor;
typedef struct coeb_frame coeb_frame;
typedef struct coeb_matcher coeb_matcher;

const char* coeb_last_error(void);
const char* coeb_version(void);
/* Number of usable sm_100 devices (0 if none). */
int coeb_device_count(void);

/* Pinned host memory for the host-buffer entry points (cudaHostAlloc / cudaFreeHost). */
int coeb_host_alloc(void** ptr, size_t bytes);
int coeb_host_free(void* ptr);

/* ---------------------------------------------------------------------------------------------
 * ORBextractor
 * ------------------------------------------------------------------------------------------- */

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * (include/ORBextractor.h:50-51, src/ORBextractor.cc:418-477). `device` is the CUDA ordinal. */
int coeb_extractor_create(const coeb_orb_params* params, int device, coeb_extractor** out);
void coeb_extractor_destroy(coeb_extractor* ex);

/* Run on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL restores the handle's own
 * stream. The *_device entry point only enqueues work on it. */
int coeb_extractor_set_stream(coeb_extractor* ex, void* cuda_stream);

/* Pre-size the device arenas for `max_batch` frames of width x height so that later calls do not
 * allocate (replaces the per-call pyramid reallocation of src/ORBextractor.cc:1351). */
int coeb_extractor_reserve(coeb_extractor* ex, int width, int height, int max_batch);

/* GetLevels / GetScaleFactor(s) / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares (include/ORBextractor.h:77-97) and mnFeaturesPerLevel.
 * Any pointer may be NULL; arrays hold nlevels entries. */
int coeb_extractor_tables(const coeb_extractor* ex, int* nlevels, float* scale, float* inv_scale, float* sigma2,
                          float* inv_sigma2, int* features_per_level);

/* ORBextractor::operator()(image, mask, img, imD, keypoints, descriptors, box, T_M, mask_result, blur_flag)
 * (include/ORBextractor.h:73-75, src/ORBextractor.cc:1088-1342) for one 8-bit gray frame in HOST memory.
 * `mask`, `img`, `imD`, `mask_result` of the reference are never read and have no counterpart here.
 *   boxes_xyxy : nbox x 4 floats (xmin, ymin, xmax, ymax), person boxes from YOLO
 *   tm_xy      : ntm x 2 floats, the moving points T_M (Frame::ProcessMovingObject)
 *   blur_flag  : nblur ints (missing entries count as 0)
 *   kps_out / desc_out : caller arrays of `cap` keypoints / cap x 32 bytes; *n_out = keypoints produced.
 * nbox == 0 gives the classic 4-argument ORB-SLAM2 operator()(image, mask, keypoints, descriptors).
 * Blocking. COEB_ERR_CAPACITY if cap is too small (*n_out still holds the required size). */
int coeb_extract(coeb_extractor* ex, const uint8_t* gray, int width, int height, int stride, const float* boxes_xyxy,
                 int nbox, const float* tm_xy, int ntm, const int* blur_flag, int nblur, coeb_keypoint* kps_out,
                 uint8_t* desc_out, int cap, int* n_out);

/* Batched form of the same call for B independent frames. Packed layouts:
 *   gray [B] frames `frame_stride` bytes apart, rows `stride` bytes apart
 *   boxes [B][max_box][4], nbox [B], tm [B][max_tm][2], ntm [B], blur_flag [B][max_box]
 *   kps_out [B][cap], desc_out [B][cap][32], counts_out [B], status_out [B] (per-frame coeb_status)
 * boxes/nbox/tm/ntm/blur_flag may all be NULL (no dynamic objects).
 * _host: all pointers are host memory; copies in, runs, copies out, blocks until done.
 * _device: all pointers are device memory on the extractor's device; enqueues on the stream and returns. */
int coeb_extract_batch_host(coeb_extractor* ex, int B, const uint8_t* gray, int width, int height, int stride,
                            size_t frame_stride, const float* boxes, const int* nbox, int max_box, const float* tm,
                            const int* ntm, int max_tm, const int* blur_flag, coeb_keypoint* kps_out,
                            uint8_t* desc_out, int* counts_out, int* status_out, int cap);
int coeb_extract_batch_device(coeb_extractor* ex, int B, const uint8_t* gray, int width, int height, int stride,
                              size_t frame_stride, const float* boxes, const int* nbox, int max_box, const float* tm,
                              const int* ntm, int max_tm, const int* blur_flag, coeb_keypoint* kps_out,
                              uint8_t* desc_out, int* counts_out, int* status_out, int cap);

/* Device-resident results of the last coeb_extract* call: keypoints [cap], descriptors [cap][32] and the count of
 * frame `frame` (the handle's own output arrays for the host entry points, the caller's arrays for
 * coeb_extract_batch_device). Valid until the next call on the handle. Consumers on another stream must order
 * themselves after the extractor's stream (coeb_frame_from_extractor does). */
int coeb_extractor_device_outputs(coeb_extractor* ex, int frame, const coeb_keypoint** d_kps, const uint8_t** d_desc,
                                  const int** d_count, int* cap);

/* Host-side copy of the same keypoints when the last call was a single-frame host call (coeb_extract, or coeb_extract_batch_host
 * with one frame): the descriptor stage writes its results into pinned host memory as well, and consumers that only need a few
 * values per keypoint from a host image (coeb_frame_from_extractor gathers the depth of ~1000 pixels instead of uploading the
 * depth map) read them there. *h_kps == NULL when the last call left no host copy. Valid until the next call on the handle. */
int coeb_extractor_host_outputs(coeb_extractor* ex, int frame, const coeb_keypoint** h_kps, int* count);

/* Upper bound of the keypoints one width x height frame can yield, for sizing `cap`: per level max(N_l + 3, 4 * nIni_l).
 * DistributeOctTree stops once it holds N_l nodes but its last expansion may overshoot by three, and its first round splits
 * every one of the nIni_l = round(w/h) root nodes whatever N_l is (src/ORBextractor.cc:546-676), so nfeatures alone is not a
 * bound (small nfeatures, wide images). */
int coeb_extractor_max_keypoints(const coeb_extractor* ex, int width, int height, int* max_out);

/* Kernel launches the last coeb_extract_batch_* call enqueued (for benchmark accounting; large batches run as several
 * sub-batches, each with its own launches). */
int coeb_extractor_launches_per_call(const coeb_extractor* ex);

/* Benchmark accounting: with profiling on, every coeb_extract_batch_* call brackets its six stages
 * (classify, pyramid, blur, FAST, octree-select, describe) with CUDA events on the launch stream;
 * coeb_extractor_stage_ms returns the device time of each stage of the last call (blocks on its end). */
int coeb_extractor_set_profiling(coeb_extractor* ex, int on);
int coeb_extractor_stage_ms(coeb_extractor* ex, float* ms6);

/* Dynamic-object decision of frame `frame` of the last call (area_flag, zero-filled mask rectangles;
 * src/ORBextractor.cc:1101-1195). Blocking. */
int coeb_extractor_dyn_info(coeb_extractor* ex, int frame, coeb_dyn_info* out);

/* ORBextractor::mvImagePyramid[level] (include/ORBextractor.h:99) of frame `frame` of the last call:
 * device pointer + geometry, or a tight host copy. blurred != 0 selects the 7x7 sigma=2 blurred level
 * used for descriptors. The 19-px border of the reference's buffers is never read on this path and
 * is not materialised. */
int coeb_pyramid_level(coeb_extractor* ex, int frame, int level, int blurred, const uint8_t** dev_ptr, int* width,
                       int* height, int* pitch);
int coeb_pyramid_level_copy(coeb_extractor* ex, int frame, int level, int blurred, uint8_t* host_dst);

/* Stage outputs of the last call for parity tests (blocking host copies).
 * candidates: FAST keypoints handed to the octree (vToDistributeKeys, src/ORBextractor.cc:846) as
 *   packed uint32 x | y<<12 | response<<24, minBorder-relative, in unspecified order.
 * level keys: per-level keypoints after octree, orientation and culling, as float4 {x, y, response, angle}
 *   in level coordinates, reference list order. */
int coeb_debug_candidates(coeb_extractor* ex, int frame, int level, uint32_t* host_out, int cap, int* n_out);
int coeb_debug_level_keys(coeb_extractor* ex, int frame, int level, float* host_out, int cap, int* n_out);

/* ---------------------------------------------------------------------------------------------
 * Producers directly in front of the extractor (run on the extractor's device and stream)
 * ------------------------------------------------------------------------------------------- */

/* cv::cvtColor(im, gray, CV_RGB2GRAY | CV_BGR2GRAY | CV_RGBA2GRAY | CV_BGRA2GRAY) of Tracking::GrabImageRGBD
 * (src/Tracking.cc:212-224): 8-bit, channels = 3 or 4, bgr != 0 for the BGR orders. OpenCV's fixed-point weights,
 * bit-exact. _batch_device: device pointers, enqueued on the stream; the host form is blocking. */
int coeb_rgb_to_gray_batch_device(coeb_extractor* ex, int B, const uint8_t* d_rgb, int width, int height, int stride,
                                  size_t frame_stride, int channels, int bgr, uint8_t* d_gray, int gray_stride,
                                  size_t gray_frame_stride);
int coeb_rgb_to_gray(coeb_extractor* ex, const uint8_t* rgb, int width, int height, int stride, int channels, int bgr,
                     uint8_t* gray_out, int gray_stride);

/* blur_flag producer of the RGB-D Frame constructor (src/Frame.cc:171-202 with Frame::detect_laplacian, :905-913):
 * per YOLO box, mean of abs(cv::Laplacian(box.clone(), CV_16U)); flag = 1 if the mean is < 4.2. Boxes are
 * [B][max_box][4] (xmin, ymin, xmax, ymax), nbox [B] (NULL: max_box boxes in every frame); flags [B][max_box] are
 * exactly the blur_flag array coeb_extract_batch_* takes; means (optional) [B][max_box] doubles, -1 for an unused or
 * out-of-image box. The reference's first-frame rule (no previous image: blur_flag = {0, 0}, :205-209) is the caller's. */
int coeb_blur_flags_batch_device(coeb_extractor* ex, int B, const uint8_t* d_gray, int width, int height, int stride,
                                 size_t frame_stride, const float* d_boxes, const int* d_nbox, int max_box, int* d_flags,
                                 double* d_means);
int coeb_blur_flags(coeb_extractor* ex, const uint8_t* gray, int width, int height, int stride, const float* boxes_xyxy,
                    int nbox, int* flags_out, double* means_out);

/* ---------------------------------------------------------------------------------------------
 * Frame keypoint grid + ORBmatcher (declared here, implemented in match.cu)
 * ------------------------------------------------------------------------------------------- */

/* Matcher context: device scratch + stream for one GPU. */
int coeb_matcher_create(int device, coeb_matcher** out);
void coeb_matcher_destroy(coeb_matcher* m);
int coeb_matcher_set_stream(coeb_matcher* m, void* cuda_stream);

/* The part of ORB_SLAM2::Frame the matchers read, resident on the device: mvKeysUn, mDescriptors,
 * mvuRight (NULL: all -1), image bounds and mvScaleFactors, plus the 64x48 grid of
 * Frame::AssignFeaturesToGrid / PosInGrid (src/Frame.cc:396-411, 558-568). Host pointers. */
int coeb_frame_create(coeb_matcher* m, const coeb_keypoint* kps_un, const uint8_t* desc, int n, const float* uright,
                      const coeb_camera* cam, const float* scale_factors, int nlevels, coeb_frame** out);
void coeb_frame_destroy(coeb_frame* f);

/* Tail of the RGB-D Frame constructor (src/Frame.cc:213-240) straight from the extractor's device-resident output, with
 * no host round trip of keypoints or descriptors: UndistortKeyPoints (:579-609, cv::undistortPoints in double, 5
 * iterations; skipped when dist_coef5 is NULL or k1 == 0 like the reference), ComputeStereoFromRGBD (:820-842, depth
 * read at the truncated distorted coordinates, uRight = xUn - bf / d) and AssignFeaturesToGrid (:396-411).
 *   frame_index : frame of the extractor's last call;  n : its keypoint count (< 0: read it from the device)
 *   dist_coef5  : k1, k2, p1, p2, k3 (mDistCoef)
 *   depth       : NULL or kind 0 = none (mvuRight = mvDepth = -1); kind 1 = float32 metres as Tracking::GrabImageRGBD
 *                 hands it over; kind 2 = raw uint16 with `factor` = mDepthMapFactor, i.e. the reference's
 *                 imDepth.convertTo(CV_32F, factor) (src/Tracking.cc:226-229) folded into the gather. Host memory
 *                 (uploaded by the call; pinned memory from coeb_host_alloc makes that copy asynchronous) unless
 *                 on_device != 0.
 *   keys_un_out / uright_out / depth_out : optional host arrays of n entries (mvKeysUn, mvuRight, mvDepth).
 * The frame is usable by every matcher entry point below. */
typedef struct coeb_depth_image {
    const void* data;
    int32_t kind;          /* 0 none, 1 float32, 2 uint16 */
    int32_t stride_bytes;
    float factor;          /* kind 2 only */
    int32_t on_device;
    int32_t width, height;
} coeb_depth_image;
int coeb_frame_from_extractor(coeb_matcher* m, coeb_extractor* ex, int frame_index, int n, const coeb_camera* cam,
                              const float* dist_coef5, const coeb_depth_image* depth, coeb_keypoint* keys_un_out,
                              float* uright_out, float* depth_out, int* n_out, coeb_frame** out);

/* Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel) (src/Frame.cc:503-556): indices in the
 * reference's traversal order (ix outer, iy inner, insertion order). */
int coeb_frame_features_in_area(coeb_frame* f, float x, float y, float r, int min_level, int max_level, int* idx_out,
                                int cap, int* n_out);

/* ORBmatcher::DescriptorDistance for ONE pair (src/ORBmatcher.cc:1648-1664), host pointers to two 32-byte rows. The reference
 * calls it for single pairs from host loops (src/Frame.cc:719, src/MapPoint.cc:281): eight popcounts on the calling thread, no
 * device work, no handle. Batches go through coeb_hamming256_batch. */
int coeb_hamming256(const void* a, const void* b);

/* ORBmatcher::DescriptorDistance for n pairs (src/ORBmatcher.cc:1648-1664). Host pointers, n x 32 bytes each. */
int coeb_hamming256_batch(coeb_matcher* m, const uint8_t* a, const uint8_t* b, int n, int* dist_out);

/* ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:45-129), with the
 * MapPoint fields flattened (set by Frame::isInFrustum, src/Frame.cc:492-498):
 *   track_in_view, bad, has_obs (Observations()>0): n bytes each; proj_x, proj_y, proj_xr, view_cos: n floats;
 *   level: n ints (mnTrackScaleLevel); desc: n x 32 bytes (GetDescriptor()).
 * kp_match (in/out, F.n ints) encodes F.mvpMapPoints: -1 empty, -2 holds a MapPoint with observations,
 * -3 holds one without; on return >= 0 is the index of the map point assigned by this call.
 * nnratio is the matcher's mfNNratio. *nmatches_out = return value of the reference function. */
int coeb_match_projection(coeb_matcher* m, coeb_frame* F, int n, const uint8_t* track_in_view, const uint8_t* bad,
                          const uint8_t* has_obs, const float* proj_x, const float* proj_y, const float* proj_xr,
                          const int* level, const float* view_cos, const uint8_t* desc, float th, float nnratio,
                          int* kp_match, int* nmatches_out);

/* Tracking::mvpLocalMapPoints resident on the device (src/Tracking.cc:1284-1311 rebuilds it per keyframe change, not per
 * frame): GetWorldPos() n x 3, GetNormal() n x 3, mfMinDistance / mfMaxDistance (the 0.8 / 1.2 invariance factors of
 * src/MapPoint.cc:373-383 are applied on the device), GetDescriptor() n x 32. Host pointers. */
typedef struct coeb_local_map coeb_local_map;
int coeb_local_map_create(coeb_matcher* m, int n, const float* xyz, const float* normal, const float* min_dist,
                          const float* max_dist, const uint8_t* desc, coeb_local_map** out);
void coeb_local_map_destroy(coeb_local_map* lm);

/* Tracking::SearchLocalPoints (src/Tracking.cc:1222-1272), second loop + matcher call, as one device pass:
 * Frame::isInFrustum(pMP, viewing_cos_limit) (src/Frame.cc:445-501) with MapPoint::PredictScale (src/MapPoint.cc:402-417)
 * for every map point with skip[i] == 0, then ORBmatcher::SearchByProjection(F, mvpLocalMapPoints, th) (:45-129).
 *   skip    : n bytes, 1 = mnLastFrameSeen == current frame id or isBad() (the caller's first loop)
 *   has_obs : n bytes, Observations() > 0
 *   Tcw     : 3x4 row-major [mRcw | mtcw];  Ow : mOw
 *   kp_match: as coeb_match_projection
 *   in_view_out : n bytes, isInFrustum's return value (the points IncreaseVisible() applies to; their sum is nToMatch)
 *   proj_out    : optional n x 5 floats {mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos, mnTrackScaleLevel}
 * The only per-frame upload is skip/has_obs/kp_match; the projected fields never leave the device. */
int coeb_search_local_points(coeb_matcher* m, coeb_frame* F, coeb_local_map* lm, const uint8_t* skip,
                             const uint8_t* has_obs, const float* Tcw, const float* Ow, float viewing_cos_limit, float th,
                             float nnratio, int* kp_match, uint8_t* in_view_out, float* proj_out, int* nmatches_out);

/* Search half of ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, th) (src/ORBmatcher.cc:826-961;
 * LocalMapping::SearchInNeighbors): for every map point of `lm` the keypoint of the keyframe `kf` it would be fused into
 * (projection, IsInImage, distance range, 60-degree viewing cone, PredictScale, window of radius th * scale, level and
 * chi-square filters (5.99 mono / 7.8 stereo), best Hamming distance <= TH_LOW), or -1. The MapPoint side effects of the
 * reference loop (Replace / AddObservation / AddMapPoint, :938-957) depend only on these indices and stay with the caller,
 * applied in list order.  valid: n bytes, pMP && !isBad() && !IsInKeyFrame(pKF).  Tcw: [GetRotation() | GetTranslation()],
 * Ow: GetCameraCenter().  best_idx: n ints out; *nfused_out = nFused.
 * chi2_tests = 0 gives the search of the loop-closing overload Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (:963-1093): the same
 * projection and filters without the reprojection-error test, with [Rcw | tcw] and Ow from the decomposed Sim3 (:973-977, the
 * caller's cv::Mat expressions) and valid = !isBad() && !spAlreadyFound.count(pMP). */
int coeb_fuse_search(coeb_matcher* m, coeb_frame* kf, coeb_local_map* lm, const uint8_t* valid, const float* Tcw, const float* Ow,
                     float th, int chi2_tests, int* best_idx, int* nfused_out);

/* ORBmatcher::SearchByProjection(Frame& cur, const Frame& last, th, bMono) (src/ORBmatcher.cc:1329-1471).
 * Per last-frame keypoint i: valid (mvpMapPoints[i] && !mvbOutlier[i]), has_obs, xyz (GetWorldPos),
 * octave (LastFrame.mvKeys[i].octave), angle (LastFrame.mvKeysUn[i].angle), desc (pMP->GetDescriptor()).
 * Tcw_cur / Tcw_last: 3x4 row-major [R|t]. check_ori is the matcher's mbCheckOrientation. */
int coeb_match_lastframe(coeb_matcher* m, coeb_frame* cur, int n, const uint8_t* valid, const uint8_t* has_obs,
                         const float* xyz, const int* octave, const float* angle, const uint8_t* desc,
                         const float* Tcw_cur, const float* Tcw_last, float th, int mono, int check_ori,
                         int* kp_match, int* nmatches_out);

/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist)
 * (src/ORBmatcher.cc:1473-1600), the guided search of Tracking::Relocalization. Per keyframe map point i:
 *   valid (pMP && !isBad() && !sAlreadyFound.count(pMP)), xyz (GetWorldPos), min_dist / max_dist (mfMinDistance / mfMaxDistance),
 *   angle (pKF->mvKeysUn[i].angle), desc (GetDescriptor()).
 * Tcw: CurrentFrame.mTcw, 3x4 row-major; Ow: -Rcw^T tcw (:1479, equal to CurrentFrame.mOw).
 * kp_match (in/out, cur.n ints): -1 = CurrentFrame.mvpMapPoints[k] is NULL, any other negative value = it holds a MapPoint
 * (every non-null entry blocks, :1546-1547; such entries come back unchanged); on return >= 0 is the keyframe map point
 * assigned by this call. orb_dist is the ORBdist acceptance threshold. */
int coeb_match_reloc(coeb_matcher* m, coeb_frame* cur, int n, const uint8_t* valid, const float* xyz, const float* min_dist,
                     const float* max_dist, const float* angle, const uint8_t* desc, const float* Tcw, const float* Ow, float th,
                     int orb_dist, int check_ori, int* kp_match, int* nmatches_out);

/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)
 * (src/ORBmatcher.cc:405-520). prev_matched: F1.n x 2 floats in/out; matches12: F1.n ints out. */
int coeb_match_init(coeb_matcher* m, coeb_frame* f1, coeb_frame* f2, float* prev_matched, int* matches12,
                    int window_size, float nnratio, int check_ori, int* nmatches_out);

/* ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vpMapPointMatches) (src/ORBmatcher.cc:158-288; strict_low = 0,
 * valid2 = NULL) and ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12) (:522-655; strict_low = 1).
 * The DBoW2::FeatureVector of each side (mFeatVec, produced by the reference's vocabulary, which stays the reference's)
 * is passed flattened: nn nodes with ascending ids `node`, node k owning items[start[k] .. start[k+1]) = feature
 * indices in DBoW2's push order. f1 holds the query side (the keyframe), f2 the searched side.
 *   valid1 : f1.n bytes, MapPoint exists and !isBad() (:190-196, :558-562)
 *   valid2 : f2.n bytes or NULL (:575-581; a Frame has no such filter)
 *   match12: f1.n ints out, matched feature of f2 or -1. For the Frame overload the caller writes
 *            vpMapPointMatches[match12[i]] = vpMapPointsKF[i].
 * Returns the match count after the rotation-histogram check (check_ori = mbCheckOrientation). */
int coeb_match_bow(coeb_matcher* m, coeb_frame* f1, coeb_frame* f2, const uint8_t* valid1, const uint8_t* valid2, int nn1,
                   const int* node1, const int* start1, const int* items1, int nn2, const int* node2, const int* start2,
                   const int* items2, float nnratio, int check_ori, int strict_low, int* match12, int* nmatches_out);

/* ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) (src/ORBmatcher.cc:657-824), with
 * CheckDistEpipolarLine (:140-156). Feature vectors as in coeb_match_bow; the frames must carry mvuRight when the
 * keyframes are stereo / RGB-D (a frame without it counts as monocular).
 *   free1 / free2 : n bytes each, the feature has NO MapPoint yet (:697-701, :722-726)
 *   F12           : 3x3 row-major fundamental matrix;  epipole_xy : (ex, ey) of :663-670, which the caller computes with
 *                   the reference's own cv::Mat expressions (R2w * Cw + t2w, projected)
 *   match12       : f1.n ints out, vMatches12 (:685); the caller builds vMatchedPairs from the entries >= 0 (:812-819). */
int coeb_match_triangulation(coeb_matcher* m, coeb_frame* f1, coeb_frame* f2, const uint8_t* free1, const uint8_t* free2,
                             int nn1, const int* node1, const int* start1, const int* items1, int nn2, const int* node2,
                             const int* start2, const int* items2, const float* F12, const float* epipole_xy,
                             int only_stereo, int check_ori, int* match12, int* nmatches_out);

/* Frame::ComputeStereoMatches (src/Frame.cc:644-818). Keypoints/descriptors are host arrays; the two
 * pyramids are taken from the extractors that produced them (their last call, frame 0).
 * uright_out / depth_out: N floats (mvuRight, mvDepth; -1 = none). */
int coeb_stereo_match(coeb_matcher* m, coeb_extractor* left, coeb_extractor* right, int N,
                      const coeb_keypoint* keys_left, const uint8_t* desc_left, int Nr,
                      const coeb_keypoint* keys_right, const uint8_t* desc_right, float bf, float b,
                      float* uright_out, float* depth_out, int* nmatched_out);

/* The same with the two pyramids named by (extractor, frame of its last call). A stereo rig whose two ORBextractors have the same parameters
 * (ORB-SLAM's do: src/Tracking.cc constructs mpORBextractorLeft / Right from the same settings) can extract both images as ONE two-frame
 * call of one extractor (coeb_extract_batch_host with B = 2: 254 us instead of 431 us for a 1241x376 pair) and match frames 0 and 1 of it. */
int coeb_stereo_match_frames(coeb_matcher* m, coeb_extractor* left, int frame_left, coeb_extractor* right, int frame_right, int N,
                             const coeb_keypoint* keys_left, const uint8_t* desc_left, int Nr, const coeb_keypoint* keys_right,
                             const uint8_t* desc_right, float bf, float b, float* uright_out, float* depth_out, int* nmatched_out);

/* Brute-force k=2 Hamming search with ratio test over a whole train set (BASELINE.json config 5;
 * semantics of the SearchByBoW inner loop, src/ORBmatcher.cc:201-231: strict '<', first index wins
 * ties, accept if best <= TH_LOW and best < nnratio * second). Host pointers.
 * best_idx: nq ints (-1 if rejected); d1/d2: best and second-best distances (may be NULL). */
int coeb_knn2(coeb_matcher* m, const uint8_t* query, int nq, const uint8_t* train, int nt, float nnratio,
              int* best_idx, int* d1, int* d2, int* naccepted_out);
/* Same with device-resident descriptors/outputs, enqueued on the matcher's stream. */
int coeb_knn2_device(coeb_matcher* m, const uint8_t* d_query, int nq, const uint8_t* d_train, int nt, float nnratio,
                     int* d_best_idx, int* d_d1, int* d_d2);

/* ---- Frame::ProcessMovingObject (src/Frame.cc:311-393): the producer of T_M ---------------------------------------------------
 * The reference body is goodFeaturesToTrack -> cornerSubPix -> calcOpticalFlowPyrLK -> its own border / SAD tests ->
 * findFundamentalMat(RANSAC) -> its own epipolar-distance test. The arithmetic of those OpenCV calls is not the reference's;
 * parity for these entry points is by the tolerances stated in DESIGN.md section 2 and tests/test_motion_gpu.py, not bit-exact. */
typedef struct coeb_motion coeb_motion;
int coeb_motion_create(int device, coeb_motion** out);
void coeb_motion_destroy(coeb_motion* m);

#define COEB_MOTION_TRACE_POINTS 1000
/* Intermediates of one coeb_process_moving_object call (for tests and diagnostics): prepoint / nextpoint / state of
 * src/Frame.cc:32-35 after the SAD test, and the fundamental matrix of :370. */
typedef struct coeb_motion_trace {
    int32_t n_points, n_tracked, n_inliers, have_F;
    float pre_xy[2 * COEB_MOTION_TRACE_POINTS], next_xy[2 * COEB_MOTION_TRACE_POINTS];
    uint8_t state[COEB_MOTION_TRACE_POINTS];
    double F[9];
} coeb_motion_trace;

/* The whole function: prev_gray is imGrayPre, cur_gray the frame being constructed (8-bit, same size). tm_xy_out receives
 * T_M (nextpoint of every tracked point whose epipolar distance exceeds 1, in point order), at most cap pairs. */
int coeb_process_moving_object(coeb_motion* m, const uint8_t* prev_gray, const uint8_t* cur_gray, int width, int height, int stride,
                               float* tm_xy_out, int cap, int* n_tm_out, coeb_motion_trace* trace /* may be NULL */);

/* The same for a sequence, as the reference runs it (imGrayPre is the frame of the previous call, src/Frame.cc:164-209): the previous frame is
 * the current frame of the last coeb_process_moving_object / coeb_process_moving_object_next call on this handle, still on the device with
 * its pyramid, so only cur_gray travels. The first call of a sequence (or the first after a change of size) keeps the frame and returns no
 * T_M, like the reference's first frame. Results are identical to coeb_process_moving_object(previous frame, cur_gray). */
int coeb_process_moving_object_next(coeb_motion* m, const uint8_t* cur_gray, int width, int height, int stride, float* tm_xy_out, int cap,
                                    int* n_tm_out, coeb_motion_trace* trace /* may be NULL */);

/* The stages, callable on their own (host arrays in and out):
 * cv::goodFeaturesToTrack(gray, corners, max_corners, quality, min_distance, noArray(), 3, true, harris_k)       (:333) */
int coeb_motion_good_features(coeb_motion* m, const uint8_t* gray, int width, int height, int stride, int max_corners, double quality,
                              double min_distance, double harris_k, float* xy_out, int cap, int* n_out);
/* cv::cornerSubPix(gray, corners, Size(half_win, half_win), Size(-1,-1), TermCriteria(ITER|EPS, max_iters, eps))    (:334) */
int coeb_motion_corner_subpix(coeb_motion* m, const uint8_t* gray, int width, int height, int stride, float* xy_inout, int n, int half_win,
                              int max_iters, double eps);
/* cv::calcOpticalFlowPyrLK(prev, cur, prev_xy, next_xy, status, err, Size(win, win), max_level,
 *                          TermCriteria(ITER|EPS, max_iters, eps), 0, min_eig_threshold)                             (:335) */
int coeb_motion_lk(coeb_motion* m, const uint8_t* prev_gray, const uint8_t* cur_gray, int width, int height, int stride, const float* prev_xy,
                   int n, int win, int max_level, int max_iters, double eps, double min_eig_threshold, float* next_xy, uint8_t* status);
/* cv::findFundamentalMat(p1, p2, FM_RANSAC, threshold, confidence) -- host code: normalised 8-point inside RANSAC with its own
 * generator (seed), final fit on the consensus set, F scaled so that F[8] = 1.                                        (:370) */
int coeb_fundamental_ransac(const float* p1_xy, const float* p2_xy, int n, double threshold, double confidence, int max_iters,
                            unsigned seed, double F_out[9], uint8_t* inlier_mask /* may be NULL */, int* n_inliers /* may be NULL */);
/* The epipolar test of :372-385 for every point with status != 0: moving_out[i] = distance > limit. */
int coeb_epipolar_outliers(coeb_motion* m, const float* pre_xy, const float* next_xy, const uint8_t* status, int n, const double F[9],
                           double limit, uint8_t* moving_out, double* dist_out /* may be NULL */);

#ifdef __cplusplus
}
#endif
#endif /* COEB_FRONTEND_H */
