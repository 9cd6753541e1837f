/*
 * coeb_types.h -- plain-old-data types shared by the C ABI (coeb_frontend.h), the C++ drop-in
 * adapters and the CPU oracle. No functions here.
 *
 * Layouts follow the reference's containers so an adapter is a memcpy:
 *   coeb_keypoint  == cv::KeyPoint field order {pt.x, pt.y, size, angle, response, octave, class_id}
 *                     (28 bytes; reference: src/ORBextractor.cc:1336, include/Frame.h mvKeys)
 *   descriptors    == row-major n x 32 uint8 (cv::Mat CV_8U, src/ORBextractor.cc:1300)
 */
#ifndef COEB_TYPES_H
#define COEB_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COEB_DESC_BYTES 32
#define COEB_MAX_LEVELS 16
#define COEB_MAX_BOXES 32        /* person boxes per frame the device path keeps resident */
#define COEB_GRID_COLS 64        /* FRAME_GRID_COLS, reference include/Frame.h:38 */
#define COEB_GRID_ROWS 48        /* FRAME_GRID_ROWS, reference include/Frame.h:37 */
#define COEB_TH_HIGH 100         /* ORBmatcher::TH_HIGH, src/ORBmatcher.cc:37 */
#define COEB_TH_LOW 50           /* ORBmatcher::TH_LOW,  src/ORBmatcher.cc:38 */
#define COEB_HISTO_LENGTH 30     /* ORBmatcher::HISTO_LENGTH, src/ORBmatcher.cc:39 */

typedef struct coeb_keypoint {
    float x, y;      /* pt, level-0 pixel coordinates after the final per-level rescale */
    float size;      /* (int)(31 * scale[octave]) stored as float */
    float angle;     /* degrees in [0,360), OpenCV fastAtan2 */
    float response;  /* FAST-9/16 corner score */
    int32_t octave;  /* pyramid level */
    int32_t class_id;/* always -1 */
} coeb_keypoint;

/* ORBextractor constructor arguments (reference include/ORBextractor.h:50-51). */
typedef struct coeb_orb_params {
    int32_t nfeatures;
    float scale_factor;
    int32_t nlevels;
    int32_t ini_th_fast; /* kept for API parity; the reference overrides both thresholds per frame */
    int32_t min_th_fast; /* (src/ORBextractor.cc:775-784): 20/7, or 30/10 when area_flag is set    */
} coeb_orb_params;

/* Per-frame summary of the COEB dynamic-object decision (src/ORBextractor.cc:1101-1195). */
typedef struct coeb_dyn_info {
    int32_t area_flag;                 /* summed dynamic box area > 200000 px^2 */
    int32_t n_dynamic;                 /* boxes judged dynamic */
    int32_t rect[COEB_MAX_BOXES][4];   /* zero-filled mask rectangles: x0,y0,x1,y1 (half-open) */
    float area;                        /* summed area of dynamic boxes, accumulated in box order */
} coeb_dyn_info;

/* Camera / frame constants the matchers read from ORB_SLAM2::Frame. */
typedef struct coeb_camera {
    float fx, fy, cx, cy;
    float bf;   /* mbf */
    float b;    /* mb  */
    float min_x, max_x, min_y, max_y; /* mnMinX.. image bounds (Frame.cc:611-642) */
} coeb_camera;

typedef enum coeb_status {
    COEB_OK = 0,
    COEB_ERR_INVALID_ARG = -1,
    COEB_ERR_NO_DEVICE = -2,     /* no sm_100 device: there is no CPU fallback */
    COEB_ERR_CUDA = -3,
    COEB_ERR_CAPACITY = -4,      /* caller buffer or preallocated arena too small */
    COEB_ERR_BAD_BOX = -5,       /* box not inside the image (the reference throws cv::Exception) */
    COEB_ERR_UNSUPPORTED = -6
} coeb_status;

#ifdef __cplusplus
}
#endif
#endif /* COEB_TYPES_H */
