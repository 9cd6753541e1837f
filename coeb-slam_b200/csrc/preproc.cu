// preproc.cu -- the two cheap producers that sit directly in front of the extractor ("next" rows 2 and 3 of
// SURVEY.md section 8f), so that a frame can enter the device path as RGB + boxes and nothing else:
//   * cv::cvtColor(RGB/BGR/RGBA/BGRA -> GRAY), reference src/Tracking.cc:212-224
//   * Frame::detect_laplacian per YOLO box -> blur_flag, reference src/Frame.cc:171-202, 905-913
// Both are exact integer restatements of OpenCV 4.13 (pinned in tests/test_oracle_vs_cv2.py).
#include <cstring>
#include <vector>

#include "../../include/coeb_frontend.h"
#include "coeb_device.cuh"
#include "coeb_host.hpp"

namespace coeb {

// RGB2Gray<uchar>: (R*9798 + G*19235 + B*3735 + 2^14) >> 15. One thread per 4 output pixels (one 32-bit store).
__global__ void __launch_bounds__(256) rgb_to_gray_kernel(const uint8_t* __restrict__ src, int w, int h, int sstride, size_t sframe, int channels,
                                                          int bgr, uint8_t* __restrict__ dst, int dstride, size_t dframe) {
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4, y = blockIdx.y * 8 + threadIdx.y, f = blockIdx.z;
    if (x0 >= w || y >= h) return;
    const uint8_t* s = src + f * sframe + (size_t)y * sstride + (size_t)x0 * channels;
    uint8_t* d = dst + f * dframe + (size_t)y * dstride + x0;
    uint32_t packed = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (x0 + i < w) {
            const int c0 = s[i * channels], c1 = s[i * channels + 1], c2 = s[i * channels + 2];
            const int r = bgr ? c2 : c0, b = bgr ? c0 : c2;
            packed |= (uint32_t)((r * 9798 + c1 * 19235 + b * 3735 + (1 << 14)) >> 15) << (8 * i);
        }
    }
    if (x0 + 3 < w && (dstride & 3) == 0 && ((size_t)dst & 3) == 0 && (dframe & 3) == 0) {
        *reinterpret_cast<uint32_t*>(d) = packed;
    } else {
        for (int i = 0; i < 4 && x0 + i < w; i++) d[i] = (uint8_t)(packed >> (8 * i));
    }
}

__device__ __forceinline__ int reflect101_box(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// One CTA per (box, frame): sum over the box of max(0, 4-neighbour Laplacian) with BORDER_REFLECT_101 on the box
// itself (the reference clones the ROI first), mean in double, flag = mean < 4.2.
__global__ void __launch_bounds__(256) blur_flag_kernel(const uint8_t* __restrict__ gray, int w, int h, int stride, size_t frame_stride,
                                                        const float* __restrict__ boxes, const int* __restrict__ nbox, int max_box,
                                                        int* __restrict__ flags, double* __restrict__ means) {
    __shared__ unsigned long long s_sum[8];
    const int b = blockIdx.x, f = blockIdx.y;
    const int tid = threadIdx.x;
    const int nb = nbox ? nbox[f] : max_box;
    const size_t o = (size_t)f * max_box + b;
    if (b >= nb) {
        if (tid == 0) { flags[o] = 0; if (means) means[o] = -1.0; }
        return;
    }
    const float* bx = boxes + o * 4;
    const int x0 = (int)bx[0], y0 = (int)bx[1], bw = (int)(bx[2] - bx[0]), bh = (int)(bx[3] - bx[1]);
    if (x0 < 0 || y0 < 0 || bw <= 0 || bh <= 0 || x0 + bw > w || y0 + bh > h) {   // the reference's cv::Mat ROI would throw
        if (tid == 0) { flags[o] = 0; if (means) means[o] = -1.0; }
        return;
    }
    const uint8_t* img = gray + f * frame_stride + (size_t)y0 * stride + x0;
    unsigned long long sum = 0;
    const int lane = tid & 31, warp = tid >> 5;
    for (int y = warp; y < bh; y += 8) {
        const uint8_t* rc = img + (size_t)y * stride;
        const uint8_t* ru = img + (size_t)reflect101_box(y - 1, bh) * stride;
        const uint8_t* rd = img + (size_t)reflect101_box(y + 1, bh) * stride;
        for (int x = lane; x < bw; x += 32) {
            const int lap = (int)ru[x] + (int)rd[x] + (int)rc[reflect101_box(x - 1, bw)] + (int)rc[reflect101_box(x + 1, bw)] - 4 * (int)rc[x];
            sum += (unsigned)max(lap, 0);
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    if (lane == 0) s_sum[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; i++) t += s_sum[i];
        const double m = (double)t / ((double)bw * (double)bh);
        flags[o] = m < 4.2 ? 1 : 0;
        if (means) means[o] = m;
    }
}

}  // namespace coeb

using namespace coeb;

extern "C" int coeb_extractor_device_stream(coeb_extractor* ex, int* device, void** stream);

extern "C" {

int coeb_rgb_to_gray_batch_device(coeb_extractor* ex, int B, const uint8_t* d_rgb, int width, int height, int stride, size_t frame_stride,
                                  int channels, int bgr, uint8_t* d_gray, int gray_stride, size_t gray_frame_stride) {
    if (!ex || !d_rgb || !d_gray || B < 1 || width < 1 || height < 1) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (channels != 3 && channels != 4) return fail(COEB_ERR_INVALID_ARG, "channels must be 3 or 4");
    if (stride < width * channels || gray_stride < width) return fail(COEB_ERR_INVALID_ARG, "row stride smaller than the row");
    int dev = 0;
    void* s = nullptr;
    int st = coeb_extractor_device_stream(ex, &dev, &s);
    if (st != COEB_OK) return st;
    CUDA_TRY(cudaSetDevice(dev));
    rgb_to_gray_kernel<<<dim3((width + 127) / 128, (height + 7) / 8, B), dim3(32, 8), 0, (cudaStream_t)s>>>(
        d_rgb, width, height, stride, frame_stride, channels, bgr, d_gray, gray_stride, gray_frame_stride);
    CUDA_TRY(cudaGetLastError());
    return COEB_OK;
}

int coeb_blur_flags_batch_device(coeb_extractor* ex, int B, const uint8_t* d_gray, int width, int height, int stride, size_t frame_stride,
                                 const float* d_boxes, const int* d_nbox, int max_box, int* d_flags, double* d_means) {
    if (!ex || !d_gray || !d_boxes || !d_flags || B < 1 || max_box < 1) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    int dev = 0;
    void* s = nullptr;
    int st = coeb_extractor_device_stream(ex, &dev, &s);
    if (st != COEB_OK) return st;
    CUDA_TRY(cudaSetDevice(dev));
    blur_flag_kernel<<<dim3(max_box, B), 256, 0, (cudaStream_t)s>>>(d_gray, width, height, stride, frame_stride, d_boxes, d_nbox, max_box, d_flags,
                                                                   d_means);
    CUDA_TRY(cudaGetLastError());
    return COEB_OK;
}

// Host-buffer conveniences (blocking): one frame.
int coeb_rgb_to_gray(coeb_extractor* ex, const uint8_t* rgb, int width, int height, int stride, int channels, int bgr, uint8_t* gray_out,
                     int gray_stride) {
    if (!ex || !rgb || !gray_out || width < 1 || height < 1) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (channels != 3 && channels != 4) return fail(COEB_ERR_INVALID_ARG, "channels must be 3 or 4");
    int dev = 0;
    void* sv = nullptr;
    int st = coeb_extractor_device_stream(ex, &dev, &sv);
    if (st != COEB_OK) return st;
    cudaStream_t s = (cudaStream_t)sv;
    CUDA_TRY(cudaSetDevice(dev));
    uint8_t *d_in = nullptr, *d_out = nullptr;
    const size_t in_pitch = (size_t)width * channels, out_pitch = (width + 3) & ~3;
    CUDA_TRY(cudaMalloc(&d_in, in_pitch * height));
    if (cudaMalloc(&d_out, out_pitch * height) != cudaSuccess) { cudaFree(d_in); return fail(COEB_ERR_CUDA, "cudaMalloc failed"); }
    cudaError_t e = cudaMemcpy2DAsync(d_in, in_pitch, rgb, stride, in_pitch, height, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        st = coeb_rgb_to_gray_batch_device(ex, 1, d_in, width, height, (int)in_pitch, in_pitch * height, channels, bgr, d_out, (int)out_pitch,
                                           out_pitch * height);
        if (st == COEB_OK) e = cudaMemcpy2DAsync(gray_out, gray_stride, d_out, out_pitch, width, height, cudaMemcpyDeviceToHost, s);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_in);
    cudaFree(d_out);
    if (st != COEB_OK) return st;
    if (e != cudaSuccess) return fail(COEB_ERR_CUDA, "rgb_to_gray failed: %s", cudaGetErrorString(e));
    return COEB_OK;
}

int coeb_blur_flags(coeb_extractor* ex, const uint8_t* gray, int width, int height, int stride, const float* boxes_xyxy, int nbox,
                    int* flags_out, double* means_out) {
    if (!ex || !gray || !flags_out || nbox < 0 || (nbox > 0 && !boxes_xyxy)) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (nbox == 0) return COEB_OK;
    int dev = 0;
    void* sv = nullptr;
    int st = coeb_extractor_device_stream(ex, &dev, &sv);
    if (st != COEB_OK) return st;
    cudaStream_t s = (cudaStream_t)sv;
    CUDA_TRY(cudaSetDevice(dev));
    uint8_t* d_gray = nullptr;
    char* d_misc = nullptr;
    const size_t misc = (size_t)nbox * (16 + 4 + 8) + 64;
    CUDA_TRY(cudaMalloc(&d_gray, (size_t)width * height));
    if (cudaMalloc(&d_misc, misc) != cudaSuccess) { cudaFree(d_gray); return fail(COEB_ERR_CUDA, "cudaMalloc failed"); }
    double* d_means = (double*)d_misc;
    float* d_boxes = (float*)(d_misc + (size_t)nbox * 8);
    int* d_flags = (int*)(d_misc + (size_t)nbox * 24);
    cudaError_t e = cudaMemcpy2DAsync(d_gray, width, gray, stride, width, height, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_boxes, boxes_xyxy, (size_t)nbox * 16, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        st = coeb_blur_flags_batch_device(ex, 1, d_gray, width, height, width, (size_t)width * height, d_boxes, nullptr, nbox, d_flags, d_means);
        if (st == COEB_OK) e = cudaMemcpyAsync(flags_out, d_flags, (size_t)nbox * 4, cudaMemcpyDeviceToHost, s);
        if (st == COEB_OK && e == cudaSuccess && means_out) e = cudaMemcpyAsync(means_out, d_means, (size_t)nbox * 8, cudaMemcpyDeviceToHost, s);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_gray);
    cudaFree(d_misc);
    if (st != COEB_OK) return st;
    if (e != cudaSuccess) return fail(COEB_ERR_CUDA, "blur_flags failed: %s", cudaGetErrorString(e));
    return COEB_OK;
}

}  // extern "C"
