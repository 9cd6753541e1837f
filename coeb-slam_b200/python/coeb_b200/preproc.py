"""ctypes binding of the producers in front of the extractor (cvtColor, blur flags). See include/coeb_frontend.h."""
import ctypes as C

import numpy as np

from . import _check, _p, lib


def rgb_to_gray(ex, img, bgr=False):
    """cv::cvtColor(img, CV_RGB2GRAY / CV_BGR2GRAY / CV_RGBA2GRAY / CV_BGRA2GRAY) on the extractor's device. img [h,w,3|4] u8."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, c = img.shape
    out = np.empty((h, w), np.uint8)
    _check(lib().coeb_rgb_to_gray(ex.h, _p(img), w, h, img.strides[0], c, int(bool(bgr)), _p(out), w))
    return out


def blur_flags(ex, gray, boxes):
    """blur_flag per box (Laplacian mean < 4.2), plus the means. gray [h,w] u8, boxes [n,4] xyxy."""
    gray = np.ascontiguousarray(gray, dtype=np.uint8)
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    flags = np.zeros(len(boxes), np.int32)
    means = np.zeros(len(boxes), np.float64)
    _check(lib().coeb_blur_flags(ex.h, _p(gray), gray.shape[1], gray.shape[0], gray.strides[0], _p(boxes), len(boxes), _p(flags), _p(means)))
    return flags, means


def rgb_to_gray_batch_device(ex, B, d_rgb, w, h, stride, frame_stride, channels, bgr, d_gray, gray_stride, gray_frame_stride):
    _check(lib().coeb_rgb_to_gray_batch_device(ex.h, int(B), _p(d_rgb), int(w), int(h), int(stride), C.c_size_t(frame_stride), int(channels),
                                               int(bool(bgr)), _p(d_gray), int(gray_stride), C.c_size_t(gray_frame_stride)))


def blur_flags_batch_device(ex, B, d_gray, w, h, stride, frame_stride, d_boxes, d_nbox, max_box, d_flags, d_means=None):
    _check(lib().coeb_blur_flags_batch_device(ex.h, int(B), _p(d_gray), int(w), int(h), int(stride), C.c_size_t(frame_stride), _p(d_boxes),
                                              _p(d_nbox), int(max_box), _p(d_flags), _p(d_means)))
