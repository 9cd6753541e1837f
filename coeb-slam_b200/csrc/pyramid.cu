// pyramid.cu -- image pyramid (E2), per-level Gaussian blur (E10) and the dynamic-box decision (E3).
//
// Replaces, bit for bit, what the reference obtains from OpenCV:
//   cv::resize(prev, level, sz, 0, 0, INTER_LINEAR)             src/ORBextractor.cc:1356
//   cv::GaussianBlur(level, level, Size(7,7), 2, 2, REFLECT_101) src/ORBextractor.cc:1318
// and the box -> mask logic of src/ORBextractor.cc:1101-1195.
#include "coeb_device.cuh"

namespace coeb {

// ------------------------------------------------------------------------------------------------
// Resize: OpenCV's 8-bit INTER_LINEAR is an 11-bit fixed-point separable filter:
//   h(dx)  = S[sx]*a0 + S[sx+1]*a1                       (a0,a1 = rint((1-fx)*2048), rint(fx*2048))
//   out    = (((b0*(h0>>4))>>16) + ((b1*(h1>>4))>>16) + 2) >> 2
// The per-column / per-row (offset, a0|a1<<16) tables are built on the host (coeb_api.cu) with the
// exact double/float arithmetic of cv::resize and kept resident.
// Each thread produces 4 adjacent output pixels and stores them as one 32-bit word.
// ------------------------------------------------------------------------------------------------
// kResizeRows = destination rows per thread (rows dy, dy+8, ... of a 128 x 8*kResizeRows tile): 4 for batches, 1 when the
// grid would otherwise be too small to fill the GPU (single-frame latency path).
template <int kResizeRows>
__global__ void __launch_bounds__(256) resize_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v,
                                                     int level) {
    const LevelGeom& D = g.lv[level];
    const LevelGeom& S = g.lv[level - 1];
    const int frame = blockIdx.z;
    const int dx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    if (dx0 >= D.w) return;
    const uint8_t* __restrict__ src = level_ptr(g, v, level - 1, frame);
    const int spitch = level_pitch(g, v, level - 1);
    uint8_t* dst = v.pyr + D.img_base + (unsigned long long)frame * D.img_stride;
    const int2* __restrict__ xt = v.tabs + D.tab_base;
    const int2* __restrict__ yt = xt + D.w;
    // the column entries (source offsets and weights) are the same for every row: fetch them once, then walk
    // kResizeRows destination rows
    int sx[4], sx1[4], a0[4], a1[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int2 xe = __ldg(&xt[min(dx0 + i, D.w - 1)]);   // the tail of the last word lands in row padding
        sx[i] = xe.x;
        sx1[i] = min(xe.x + 1, S.w - 1);
        a0[i] = xe.y & 0xFFFF;
        a1[i] = xe.y >> 16;
    }
#pragma unroll
    for (int j = 0; j < kResizeRows; j++) {
        const int dy = blockIdx.y * (8 * kResizeRows) + threadIdx.y + 8 * j;
        if (dy >= D.h) break;
        const int2 ye = __ldg(&yt[dy]);
        const int sy0 = min(max(ye.x, 0), S.h - 1), sy1 = min(max(ye.x + 1, 0), S.h - 1);
        const int b0 = ye.y & 0xFFFF, b1 = ye.y >> 16;
        const uint8_t* r0 = src + (size_t)sy0 * spitch;
        const uint8_t* r1 = src + (size_t)sy1 * spitch;
        uint32_t packed = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int h0 = __ldg(r0 + sx[i]) * a0[i] + __ldg(r0 + sx1[i]) * a1[i];
            const int h1 = __ldg(r1 + sx[i]) * a0[i] + __ldg(r1 + sx1[i]) * a1[i];
            int o = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            o = min(max(o, 0), 255);
            packed |= (uint32_t)o << (8 * i);
        }
        *reinterpret_cast<uint32_t*>(dst + (size_t)dy * D.pitch + dx0) = packed;  // pitch is a multiple of 64: padding absorbs the tail
    }
}

void launch_pyramid(const Geometry& g, const BatchView& v, cudaStream_t stream) {
    for (int l = 1; l < g.nlevels; l++) {
        dim3 block(32, 8);
        const int tiles_x = (g.lv[l].w + 127) / 128;
        if ((long long)tiles_x * ((g.lv[l].h + 31) / 32) * v.B >= 2 * 148) {
            resize_kernel<4><<<dim3(tiles_x, (g.lv[l].h + 31) / 32, v.B), block, 0, stream>>>(g, v, l);
        } else {
            resize_kernel<1><<<dim3(tiles_x, (g.lv[l].h + 7) / 8, v.B), block, 0, stream>>>(g, v, l);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Gaussian 7x7 sigma=2, OpenCV bit-exact fixed point: Q8 kernel {18,34,48,56,48,34,18};
// horizontal pass -> Q8.8 (uint16), vertical pass -> Q16.16, (v + 32768) >> 16. BORDER_REFLECT_101
// on the level itself (the reference blurs a clone of the ROI, so the pyramid border is not seen).
// One CTA = 128 x 32 output tile of one level of one frame; all levels in one launch.
//   stage : aligned 32-bit words into shared memory (rows reflected by index, the <= 3 reflected
//           columns at the image's left/right edge patched in place)
//   H pass: 4 px per thread; the 10 source bytes become nine s16x2 pairs (funnel shift + PRMT) and
//           the taps are packed 16-bit multiply-adds (no lane can overflow: 255*256 < 65536)
//   V pass: 4 px x 4 rows per thread from the uint16 intermediate, 32-bit accumulation, one 32-bit store per row
// ------------------------------------------------------------------------------------------------
constexpr int kBlurTW = 128, kBlurTH = 32, kBlurThreads = 256;
constexpr int kBlurInWords = kBlurTW / 4 + 8;       // 40 words = ten 16-byte vectors per row, x from tx0-16
constexpr int kBlurRows = kBlurTH + 6;

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

__global__ void __launch_bounds__(kBlurThreads, 6) blur_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v,
                                                            const __grid_constant__ TileMap tm) {
    __shared__ __align__(16) uint32_t s_in[kBlurRows * kBlurInWords];
    __shared__ uint2 s_h[kBlurRows * (kBlurTW / 4 + 1)];
    const int frame = blockIdx.y;
    int level = 0;
    while (level + 1 < g.nlevels && (int)blockIdx.x >= tm.tile_base[level + 1]) level++;
    const LevelGeom& L = g.lv[level];
    const int t = blockIdx.x - tm.tile_base[level];
    const int tx0 = (t % tm.tiles_x[level]) * kBlurTW, ty0 = (t / tm.tiles_x[level]) * kBlurTH;
    const uint8_t* __restrict__ src = level_ptr(g, v, level, frame);
    const int spitch = level_pitch(g, v, level);
    const int tid = threadIdx.x;

    // stage rows ty0-3 .. ty0+34 (reflected by index), 160 bytes from x = tx0-16 as ten 16-byte loads per row
    for (int i = tid; i < kBlurRows * 10; i += kBlurThreads) {
        const int ry = i / 10, q = i - ry * 10;
        const int gy = reflect101(min(ty0 + ry - 3, L.h + 2), L.h);
        const int gx = tx0 - 16 + 16 * q;
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        if (gx >= 0 && gx + 16 <= spitch) w = __ldg(reinterpret_cast<const uint4*>(src + (size_t)gy * spitch + gx));
        *reinterpret_cast<uint4*>(&s_in[ry * kBlurInWords + 4 * q]) = w;
    }
    __syncthreads();
    uint8_t* s_b = reinterpret_cast<uint8_t*>(s_in);   // byte view: column c <-> x = tx0 - 16 + c
    constexpr int kRowBytes = kBlurInWords * 4;
    if (tx0 == 0) {   // x = -1,-2,-3  <-  x = 1,2,3
        for (int i = tid; i < kBlurRows * 3; i += kBlurThreads) {
            const int ry = i / 3, k = i - ry * 3 + 1;
            s_b[ry * kRowBytes + 16 - k] = s_b[ry * kRowBytes + 16 + k];
        }
    }
    if (tx0 + kBlurTW + 3 > L.w - 1) {   // x = w, w+1, w+2  <-  x = w-2, w-3, w-4
        for (int i = tid; i < kBlurRows * 3; i += kBlurThreads) {
            const int ry = i / 3, k = i - ry * 3;
            const int c = L.w + k - tx0 + 16, cs = L.w - 2 - k - tx0 + 16;
            if (c < kRowBytes && cs >= 0) s_b[ry * kRowBytes + c] = s_b[ry * kRowBytes + cs];
        }
    }
    __syncthreads();

    for (int i = tid; i < kBlurRows * (kBlurTW / 4); i += kBlurThreads) {
        const int ry = i >> 5, gx = i & 31;
        const uint32_t* w = &s_in[ry * kBlurInWords + gx + 3]; // w[0]: x-4.., w[1]: the 4 output pixels, w[2]: x+4..
        const uint32_t wm = w[0], w0 = w[1], wp = w[2];
        // B[i] = byte i of (wm, w0, wp); pair P_i = (B[i], B[i+1]) widened to 16 bits
        const uint32_t Sa = __funnelshift_r(wm, w0, 8), Sb = __funnelshift_r(wm, w0, 16), Sc = __funnelshift_r(wm, w0, 24);
        const uint32_t Sd = __funnelshift_r(w0, wp, 8), Se = __funnelshift_r(w0, wp, 16), Sf = __funnelshift_r(w0, wp, 24);
        const uint32_t P1 = __byte_perm(Sa, 0u, 0x4140), P2 = __byte_perm(Sb, 0u, 0x4140), P3 = __byte_perm(Sa, 0u, 0x4342),
                       P4 = __byte_perm(Sb, 0u, 0x4342), P5 = __byte_perm(Sc, 0u, 0x4342), P6 = __byte_perm(w0, 0u, 0x4342),
                       P7 = __byte_perm(Sd, 0u, 0x4342), P8 = __byte_perm(Se, 0u, 0x4342), P9 = __byte_perm(Sf, 0u, 0x4342);
        uint2 o;
        o.x = 18u * (P1 + P7) + 34u * (P2 + P6) + 48u * (P3 + P5) + 56u * P4;   // outputs x, x+1 (Q8.8 each)
        o.y = 18u * (P3 + P9) + 34u * (P4 + P8) + 48u * (P5 + P7) + 56u * P6;   // outputs x+2, x+3
        s_h[ry * (kBlurTW / 4 + 1) + gx] = o;
    }
    __syncthreads();

    uint8_t* dst = blur_ptr(g, v, level, frame);
    {
        const int gx = tid & 31, strip = tid >> 5;           // 32 column groups x 8 strips of 4 rows
        const int x = tx0 + 4 * gx;
        if (x < L.w) {
            uint32_t hv[10][4];
#pragma unroll
            for (int r = 0; r < 10; r++) {
                const uint2 q = s_h[(strip * 4 + r) * (kBlurTW / 4 + 1) + gx];
                hv[r][0] = q.x & 0xFFFFu; hv[r][1] = q.x >> 16; hv[r][2] = q.y & 0xFFFFu; hv[r][3] = q.y >> 16;
            }
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int y = ty0 + strip * 4 + r;
                if (y < L.h) {
                    uint32_t packed = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t acc = 18u * (hv[r][k] + hv[r + 6][k]) + 34u * (hv[r + 1][k] + hv[r + 5][k]) +
                                             48u * (hv[r + 2][k] + hv[r + 4][k]) + 56u * hv[r + 3][k];
                        packed |= ((acc + 32768u) >> 16) << (8 * k);
                    }
                    *reinterpret_cast<uint32_t*>(dst + (size_t)y * L.pitch + x) = packed;   // tail lands in row padding
                }
            }
        }
    }
}

void launch_blur(const Geometry& g, const BatchView& v, cudaStream_t stream) {
    TileMap tm;
    int total = 0;
    for (int l = 0; l < g.nlevels; l++) {
        tm.tile_base[l] = total;
        tm.tiles_x[l] = (g.lv[l].w + kBlurTW - 1) / kBlurTW;
        total += tm.tiles_x[l] * ((g.lv[l].h + kBlurTH - 1) / kBlurTH);
    }
    tm.tile_base[g.nlevels] = total;
    blur_kernel<<<dim3(total, v.B), kBlurThreads, 0, stream>>>(g, v, tm);
}

// ------------------------------------------------------------------------------------------------
// Dynamic-object decision (src/ORBextractor.cc:1101-1195). One warp per frame.
//   count  = #T_M points whose truncated coordinates fall in Rect(int(xmin), int(ymin), int(xmax-xmin), int(ymax-ymin))
//   layer1 = T_M not empty && count*10000 > area_box   (the reference tests after every point with an
//            early break; the count is monotone, so this equals the test on the final count)
//   layer2 = blur_flag[b]==1 && count>0
// Dynamic boxes are recorded as the rectangle [(int)xmin,(int)xmax) x [(int)ymin,(int)ymax) that the
// reference zero-fills in its 480x640 mask; area accumulates in box order (fp32).
// ------------------------------------------------------------------------------------------------
__global__ void classify_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v) {
    const int frame = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (frame >= v.B) return;
    DynState* out = &v.dyn[frame];
    const int nbox = v.nbox ? v.nbox[frame] : 0;
    const int ntm = v.ntm ? v.ntm[frame] : 0;
    const float* boxes = v.boxes + (size_t)frame * v.max_box * 4;
    const float* tm = v.tm + (size_t)frame * v.max_tm * 2;
    const int* blur = v.blur_flag + (size_t)frame * v.max_box;
    float area = 0.f;
    int ndyn = 0, bad = 0;
    if (nbox > v.max_box || nbox > COEB_MAX_BOXES || ntm > v.max_tm) bad = 1;
    for (int b = 0; b < nbox && !bad; b++) {
        const float xmin = boxes[4 * b], ymin = boxes[4 * b + 1], xmax = boxes[4 * b + 2], ymax = boxes[4 * b + 3];
        const int rx = (int)xmin, ry = (int)ymin, rw = (int)(xmax - xmin), rh = (int)(ymax - ymin);
        if (rx < 0 || ry < 0 || rw < 0 || rh < 0 || rx + rw > g.w0 || ry + rh > g.h0 || (int)xmax > g.w0 || (int)ymax > g.h0) {
            bad = 1;
            break;
        }
        const float area_box = __fmul_rn(xmax - xmin, ymax - ymin);
        int count = 0;
        for (int t0 = 0; t0 < ntm; t0 += 32) {
            const int t = t0 + lane;
            bool in = false;
            if (t < ntm) {
                const int tx = (int)tm[2 * t], ty = (int)tm[2 * t + 1];
                in = tx >= rx && tx < rx + rw && ty >= ry && ty < ry + rh;
            }
            count += __popc(__ballot_sync(0xffffffffu, in));
        }
        bool mark = ntm > 0 && (float)((unsigned long long)count * 10000ull) > area_box;
        if (!mark && blur[b] == 1 && count > 0) mark = true;
        if (mark) {
            area = __fadd_rn(area, area_box);
            if (lane == 0) {
                out->rect[ndyn][0] = (int)xmin; out->rect[ndyn][1] = (int)ymin;
                out->rect[ndyn][2] = (int)xmax; out->rect[ndyn][3] = (int)ymax;
            }
            ndyn++;
        }
    }
    if (lane == 0) {
        out->n_dynamic = bad ? 0 : ndyn;
        out->area = area;
        out->area_flag = (!bad && area > 200000.f) ? 1 : 0;
        out->bad_box = bad;
        v.status[frame] = bad ? COEB_ERR_BAD_BOX : COEB_OK;
    }
}

void launch_classify(const Geometry& g, const BatchView& v, cudaStream_t stream) {
    const int warps = 4;
    classify_kernel<<<(v.B + warps - 1) / warps, warps * 32, 0, stream>>>(g, v);
}

}  // namespace coeb
