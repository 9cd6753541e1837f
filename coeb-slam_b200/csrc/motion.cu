// motion.cu -- Frame::ProcessMovingObject (reference src/Frame.cc:311-393), SURVEY.md section 8(f) row 1: the producer of T_M,
// the "moving" points the extractor's box classification consumes (src/ORBextractor.cc:1101-1195).
//
// The reference body is six OpenCV calls and two loops of its own:
//   goodFeaturesToTrack(prev, 1000, 0.01, 8, noArray, 3, useHarris, 0.04)        :333   -> harris_response / harris_candidates / sort_candidates kernels; the
//                                                                                           minimum-distance pass in the sort kernel's CTA (whole call) or on the host
//   cornerSubPix(prev, pts, (10,10), (-1,-1), (ITER|EPS, 20, 0.03))              :334   -> corner_subpix_cta_kernel (one CTA of four warps per corner)
//   calcOpticalFlowPyrLK(prev, cur, pts, next, state, err, (22,22), 5, (20,.01)) :335   -> pyr_down / scharr kernels (one captured graph on a side stream),
//                                                                                           lk_cta_kernel (one CTA per point, all pyramid levels in one launch)
//   5-px border test, 3x3 SAD test (limit 2120)                                  :336-364 -> tail of the LK kernel
//   findFundamentalMat(F_pre, F_next, mask, FM_RANSAC, 0.1, 0.99)                :370   -> host (normalised 8-point inside RANSAC, own generator)
//   epipolar distance > 1 -> T_M                                                 :372-385 -> on the host tracks inside coeb_process_moving_object, epipolar_kernel
//                                                                                           behind coeb_epipolar_outliers (double precision, the reference's expression)
//
// The arithmetic of the OpenCV calls is OpenCV's (third party, pinned to 4.13.0). It is restated here operation by operation
// (integer pyramid, integer Scharr derivatives and Q14 bilinear weights of the LK tracker are exact; the float parts follow the
// scalar code paths without FMA), but OpenCV's own float paths are SIMD builds with fused multiply-adds and another summation
// order, and its RANSAC draws from its own generator: parity for this row is by tolerance (tests/test_motion_gpu.py).
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/coeb_frontend.h"
#include "coeb_device.cuh"
#include "coeb_host.hpp"

namespace coeb {

constexpr int kMoMaxLevels = 8;
constexpr int kMoMaxPts = 4096;
constexpr int kMoMaxCand = 1 << 17;

__device__ __forceinline__ int refl101(int i, int n) {   // one reflection: valid for -n < i < 2n - 1 (n >= 2)
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

// ---------------------------------------------------------------------------------------------------------------------
// cv::cornerHarris(src, dst, blockSize 3, ksize 3, k) on CV_8U: Sobel derivatives scaled by 1 / (4 * 3 * 255), products,
// unnormalised 3x3 box sums (accumulated in double by OpenCV's box filter for float input), a*c - b*b - k*(a+c)^2 in double.
// Borders: BORDER_REFLECT_101 for the Sobel pass and for the box sums.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kHtW = 32, kHtH = 8;
__global__ void __launch_bounds__(kHtW * kHtH) harris_response_kernel(const uint8_t* __restrict__ img, int w, int h, int pitch, double k, float s1, float s2,
                                                                      float* __restrict__ resp, unsigned* __restrict__ max_bits) {
    __shared__ float s_dx[(kHtH + 2) * (kHtW + 2)], s_dy[(kHtH + 2) * (kHtW + 2)];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * kHtW + tx;
    const int x0 = blockIdx.x * kHtW, y0 = blockIdx.y * kHtH;
    // The derivative at a position outside the image is the derivative at its mirror image (the box filter reflects the
    // derivative planes), so the 1-px frame of the tile is filled through the mirrored CENTRE coordinates.
    // derivative planes for the tile with a 1-px frame: each entry is evaluated at its own (mirrored) centre
    for (int i = tid; i < (kHtH + 2) * (kHtW + 2); i += kHtW * kHtH) {
        const int ly = i / (kHtW + 2), lx = i - ly * (kHtW + 2);
        const int cx = refl101(min(max(x0 + lx - 1, -1), w), w), cy = refl101(min(max(y0 + ly - 1, -1), h), h);
        const int xm = refl101(cx - 1, w), xp = refl101(cx + 1, w), ym = refl101(cy - 1, h), yp = refl101(cy + 1, h);
        const uint8_t *r0 = img + (size_t)ym * pitch, *r1 = img + (size_t)cy * pitch, *r2 = img + (size_t)yp * pitch;
        // dx: row filter [-1 0 1] (exact), column filter s*[1 2 1]: (2s)*r(y) + s*(r(y+1) + r(y-1))
        const float d0 = (float)((int)r0[xp] - (int)r0[xm]), d1 = (float)((int)r1[xp] - (int)r1[xm]), d2 = (float)((int)r2[xp] - (int)r2[xm]);
        s_dx[i] = __fadd_rn(__fmul_rn(s2, d1), __fmul_rn(s1, __fadd_rn(d2, d0)));
        // dy: row filter s*[1 2 1]: p(x)*(2s) + (p(x-1) + p(x+1))*s, column filter [-1 0 1] (exact difference of the two rows)
        const float t0 = __fadd_rn(__fmul_rn((float)r0[cx], s2), __fmul_rn((float)((int)r0[xm] + (int)r0[xp]), s1));
        const float t2 = __fadd_rn(__fmul_rn((float)r2[cx], s2), __fmul_rn((float)((int)r2[xm] + (int)r2[xp]), s1));
        s_dy[i] = __fsub_rn(t2, t0);
    }
    __syncthreads();
    const int x = x0 + tx, y = y0 + ty;
    if (x >= w || y >= h) return;
    double a = 0, b = 0, c = 0;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const float dx = s_dx[(ty + j) * (kHtW + 2) + tx + i], dy = s_dy[(ty + j) * (kHtW + 2) + tx + i];
            a += (double)__fmul_rn(dx, dx);
            b += (double)__fmul_rn(dx, dy);
            c += (double)__fmul_rn(dy, dy);
        }
    const float fa = (float)a, fb = (float)b, fc = (float)c;   // the box filter stores floats
    const float r = (float)((double)fa * fc - (double)fb * fb - k * ((double)fa + fc) * ((double)fa + fc));
    resp[(size_t)y * w + x] = r;
    if (r > 0.f) atomicMax(max_bits, __float_as_uint(r));   // positive floats order like their bit patterns
}

// goodFeaturesToTrack: threshold(eig, max * quality, THRESH_TOZERO), dilate 3x3, keep interior pixels with val != 0 && val == dilated.
__global__ void __launch_bounds__(256) harris_candidates_kernel(const float* __restrict__ resp, int w, int h, const unsigned* __restrict__ max_bits, float quality,
                                                                float2* __restrict__ cand, int* __restrict__ n_cand, int cap) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x < 1 || y < 1 || x >= w - 1 || y >= h - 1) return;
    const float thresh = (float)((double)__uint_as_float(*max_bits) * (double)quality);
    const float v = resp[(size_t)y * w + x];
    if (!(v > thresh)) return;
    float m = v;
#pragma unroll
    for (int j = -1; j <= 1; j++)
#pragma unroll
        for (int i = -1; i <= 1; i++) {
            const float n = resp[(size_t)(y + j) * w + x + i];
            if (n > thresh) m = fmaxf(m, n);
        }
    if (v == m) {
        const int pos = atomicAdd(n_cand, 1);
        if (pos < cap) cand[pos] = make_float2(v, __int_as_float(y * w + x));
    }
}

// goodFeaturesToTrack sorts the candidates by (response descending, address descending): both orders are the order of the 64-bit
// key response-bits << 32 | raster index (responses are positive floats). One CTA sorts up to kSortCap keys in shared memory
// (bitonic network); longer lists are sorted on the host.
constexpr int kSortCap = 8192;
// The sorted list and the {max response bits, count} words go straight into the caller's mapped pinned block (`host`: 16 bytes of
// info, then the candidates): no device-to-host copy nodes behind the kernel. Lists the kernel does not sort (n > kSortCap) are fetched
// from `cand` by the host.
// The kernel is the last reader of the two words and clears them for the next call's Harris kernels (no memset node per call).
//
// With sel.enabled the same CTA goes on to goodFeaturesToTrack's minimum-distance pass (imgproc/featureselect.cpp: candidates in sorted
// order; one is accepted unless an ALREADY ACCEPTED one lies closer than minDistance; stop at maxCorners), so that the corners never
// visit the host and cornerSubPix / LK can be enqueued behind this kernel without a synchronisation. The pass is sequential as written;
// here accepted(i) = no accepted j < i within the distance is iterated to its (unique: the dependencies run from earlier to later
// candidates only) fixed point, 1024 candidates at a time in sorted order: everything before the current chunk is final, so a
// candidate is either killed by a final earlier one or depends on the few close candidates of its own chunk (listed once in shared
// memory); in-place updates, a chunk is done when one sweep changes nothing. The chunks stop once maxCorners are accepted.
constexpr int kSelMaxCells = 4800, kSelNb = 12, kSelAccSlots = 4, kSelChunkSlots = 12;
// shared-memory layout of the pass (bytes; it aliases the sorted keys, which are turned into positions first). The per-cell counters are
// bytes, four to a word (atomicAdd on the word; a cell holds at most 64 pixels), which is what lets the chunk lists have 12 slots per cell:
// with 8, strong corners packed along one edge overflowed them on some frames and the spill list below made the pass three times longer.
constexpr int oSelPos = 0;                                             // unsigned[kSortCap]: x | y << 16 of candidate i (sorted order)
constexpr int oSelA = oSelPos + kSortCap * 4;                          // uchar[1024]: accepted flag of the chunk's candidates
constexpr int oSelAccCnt = oSelA + 1024;                               // uchar[cells]: accepted corners per cell (all chunks so far)
constexpr int oSelAccSlot = oSelAccCnt + kSelMaxCells;                 // ushort[cells][kSelAccSlots]: their candidate numbers
constexpr int oSelChCnt = oSelAccSlot + 2 * kSelMaxCells * kSelAccSlots;   // uchar[cells]: live candidates of the current chunk per cell
constexpr int oSelChSlot = oSelChCnt + kSelMaxCells;                   // ushort[cells][kSelChunkSlots]: their thread numbers
constexpr int oSelNb = oSelChSlot + 2 * kSelMaxCells * kSelChunkSlots; // ushort[1024][kSelNb]: close live predecessors of a thread's candidate
constexpr int oSelWarp = oSelNb + 1024 * kSelNb * 2;                   // int[40]: scan scratch, accepted so far, overflow flag, spill count
constexpr int kSelSpill = 256;
constexpr int oSelSpill = oSelWarp + 160;                              // ushort[kSelSpill]: live candidates of the chunk whose cell list was full
constexpr int kSelPool = 4096;
constexpr int oSelPool = oSelSpill + 2 * kSelSpill;                    // ushort[kSelPool]: close live predecessors beyond the kSelNb of a thread's own list
constexpr int kSelSmem = oSelPool + 2 * kSelPool;
static_assert(oSelAccSlot % 8 == 0 && kSelAccSlots == 4 && oSelChSlot % 8 == 0 && oSelNb % 8 == 0 && kSelChunkSlots % 4 == 0 && kSelNb % 4 == 0, "slot rows and neighbour lists are read four entries (one uint2) at a time");
static_assert(oSelAccCnt % 4 == 0 && oSelChCnt % 4 == 0 && kSelMaxCells % 4 == 0, "byte counters are updated through their words");
// byte counter c of a packed array: add one (returns the old count) / read / clear
__device__ __forceinline__ int cnt8_inc(unsigned* w, int c) { const int sh = 8 * (c & 3); return (int)((atomicAdd(&w[c >> 2], 1u << sh) >> sh) & 0xFFu); }
__device__ __forceinline__ int cnt8_get(const unsigned* w, int c) { return (int)((w[c >> 2] >> (8 * (c & 3))) & 0xFFu); }
__device__ __forceinline__ void cnt8_clear(unsigned* w, int c) { atomicAnd(&w[c >> 2], ~(0xFFu << (8 * (c & 3)))); }
static_assert(kSelSmem >= kSortCap * 8 && kSelSmem <= 227 * 1024, "the selection arrays alias the sorted keys and fit one SM");
struct SelectArgs {
    int enabled, w, cell, gw, gh, max_corners, d2max;   // d2max: largest integer squared distance that is < minDistance^2
    float2 *out_dev, *out_host;   // accepted corners in order (device for cornerSubPix, mapped host copy for the caller)
    int *n_dev, *n_host;          // their number; -1: not selected here (too many candidates, or a grid that does not fit): host pass
};
// Bitonic sort (descending) of the n candidate keys (response bits << 32 | pixel index; padded with zeros to m, a power of two) by one CTA of
// 1024 threads, result in s_key[0..m). A thread holds K = m / 1024 consecutive keys in registers: the stages with a stride below K are
// compare-exchanges inside the thread, those below 32 K are warp shuffles, and only the few strides that cross warps go through shared
// memory (15 of the 66 stages of 2048 keys; the all-shared form was bound by its four 8-byte shared-memory accesses per pair and stage).
template <int K>
__device__ __forceinline__ void sort_keys_desc(unsigned long long* __restrict__ s_key, const float2* __restrict__ cand, int n, int m, int tid) {
    unsigned long long k[K];
#pragma unroll
    for (int r = 0; r < K; r++) {
        const int e = tid * K + r;
        unsigned long long v = 0ull;   // padding sorts last
        if (e < n) { const float2 c = cand[e]; v = ((unsigned long long)__float_as_uint(c.x) << 32) | (unsigned)__float_as_int(c.y); }
        k[r] = v;
    }
    for (int size = 2; size <= m; size <<= 1) {
        int stride = size >> 1;
        if (stride >= 32 * K) {   // strides that cross warps: through shared memory
#pragma unroll
            for (int r = 0; r < K; r++) s_key[tid * K + r] = k[r];
            __syncthreads();
            for (; stride >= 32 * K; stride >>= 1) {
                for (int t = tid; t < (m >> 1); t += 1024) {
                    const int i = 2 * t - (t & (stride - 1)), j = i + stride;
                    const bool desc = (i & size) == 0;
                    const unsigned long long a = s_key[i], b = s_key[j];
                    if ((a < b) == desc) { s_key[i] = b; s_key[j] = a; }
                }
                __syncthreads();
            }
#pragma unroll
            for (int r = 0; r < K; r++) k[r] = s_key[tid * K + r];
        }
        for (; stride >= K; stride >>= 1) {   // strides inside the warp: the partner key sits in lane ^ (stride / K), same register
            const int lx = stride / K;
            const bool upper = (tid & lx) != 0;
#pragma unroll
            for (int r = 0; r < K; r++) {
                const bool desc = ((tid * K + r) & size) == 0;
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, k[r], lx);
                k[r] = ((k[r] > o) == (desc != upper)) ? k[r] : o;   // keeps the larger key where the pair's larger one belongs
            }
        }
#pragma unroll
        for (int st = K / 2; st >= 1; st >>= 1) {   // strides inside the thread
            if (st > (size >> 1)) continue;
#pragma unroll
            for (int r = 0; r < K; r++) {
                if (r & st) continue;
                const bool desc = ((tid * K + r) & size) == 0;
                const unsigned long long a = k[r], b = k[r | st];
                if ((a < b) == desc) { k[r] = b; k[r | st] = a; }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < K; r++) s_key[tid * K + r] = k[r];
    __syncthreads();
}

__global__ void __launch_bounds__(1024) sort_candidates_kernel(float2* __restrict__ cand, unsigned* __restrict__ info, char* __restrict__ host, const SelectArgs sel) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned long long* const s_key = reinterpret_cast<unsigned long long*>(s_raw);
    const int tid = threadIdx.x, T = blockDim.x;
    const int n = (int)info[1];
#ifdef COEB_SEL_DEBUG   // development build (tools/build_variant.sh NAME motion.cu -DCOEB_SEL_DEBUG=1): one line per phase and chunk with SM cycles and list statistics
    const long long dbg_t0 = clock64();
#endif
    if (tid < 4) reinterpret_cast<unsigned*>(host)[tid] = info[tid];
    __syncthreads();
    if (tid < 4) info[tid] = 0u;
    float2* const out = reinterpret_cast<float2*>(host + 16);
    const bool select_here = sel.enabled && n <= kSortCap && sel.gw * sel.gh <= kSelMaxCells;
    if (sel.enabled && !select_here && tid == 0) { *sel.n_dev = -8; *sel.n_host = -8; }
    if (n == 0 && select_here && tid == 0) { *sel.n_dev = 0; *sel.n_host = 0; }
    if (n == 1 && tid == 0) {
        out[0] = cand[0];
        if (select_here) {
            const int idx = __float_as_int(cand[0].y), y = idx / sel.w;
            const float2 c = make_float2((float)(idx - y * sel.w), (float)y);
            sel.out_dev[0] = c; sel.out_host[0] = c; *sel.n_dev = 1; *sel.n_host = 1;
        }
    }
    if (n > kSortCap || n < 2) return;
    int m = 2;
    while (m < n) m <<= 1;
    if (m <= 1024) sort_keys_desc<1>(s_key, cand, n, m, tid);
    else if (m == 2048) sort_keys_desc<2>(s_key, cand, n, m, tid);
    else if (m == 4096) sort_keys_desc<4>(s_key, cand, n, m, tid);
    else sort_keys_desc<8>(s_key, cand, n, m, tid);
#ifdef COEB_SEL_DEBUG
    const long long dbg_t1 = clock64();
#endif
    if (!select_here) {
        for (int i = tid; i < n; i += T) {
            const unsigned long long k = s_key[i];
            out[i] = make_float2(__uint_as_float((unsigned)(k >> 32)), __int_as_float((int)(unsigned)k));
        }
        return;
    }
    // ---- minimum-distance pass ----
    unsigned* const pos = reinterpret_cast<unsigned*>(s_raw + oSelPos);
    unsigned char* const acc = s_raw + oSelA;
    unsigned* const acc_cnt = reinterpret_cast<unsigned*>(s_raw + oSelAccCnt);
    unsigned short* const acc_slot = reinterpret_cast<unsigned short*>(s_raw + oSelAccSlot);
    unsigned* const ch_cnt = reinterpret_cast<unsigned*>(s_raw + oSelChCnt);
    unsigned short* const ch_slot = reinterpret_cast<unsigned short*>(s_raw + oSelChSlot);
    unsigned short* const nbl = reinterpret_cast<unsigned short*>(s_raw + oSelNb) + tid * kSelNb;
    int* const s_warp = reinterpret_cast<int*>(s_raw + oSelWarp);   // [0..32] warp totals, [33] chunk total, [34] accepted so far, [35] a list overflowed for good, [36] spilled candidates, [37] pool entries handed out
    unsigned short* const spill = reinterpret_cast<unsigned short*>(s_raw + oSelSpill);
    unsigned short* const pool = reinterpret_cast<unsigned short*>(s_raw + oSelPool);
    const int ncell = sel.gw * sel.gh, cell = sel.cell;
    unsigned mine[kSortCap / 1024];
#pragma unroll
    for (int k = 0; k < kSortCap / 1024; k++) {
        const int i = tid + k * 1024;
        unsigned v = 0u;
        if (i < n) { const int idx = (int)(unsigned)s_key[i], y = idx / sel.w; v = (unsigned)(idx - y * sel.w) | ((unsigned)y << 16); }
        mine[k] = v;
    }
    __syncthreads();   // every key has been read: its storage is reused from here on
#pragma unroll
    for (int k = 0; k < kSortCap / 1024; k++) { const int i = tid + k * 1024; if (i < n) pos[i] = mine[k]; }
    for (int c = tid; c < (ncell + 3) / 4; c += T) { acc_cnt[c] = 0u; ch_cnt[c] = 0u; }
    if (tid < 4) s_warp[34 + tid] = 0;
    __syncthreads();
    const int lane = tid & 31, wid = tid >> 5;
#ifdef COEB_SEL_DEBUG
    if (tid == 0) printf("[sel] cycles: sort %lld, prepare %lld\n", dbg_t1 - dbg_t0, clock64() - dbg_t1);
#endif
    for (int base = 0; base < n; base += T) {
#ifdef COEB_SEL_DEBUG
        const long long dbg_c0 = clock64();
#endif
        if (s_warp[34] >= sel.max_corners || s_warp[35]) break;   // uniform: both written before the last barrier of the previous chunk
        const int i = base + tid;
        const bool valid = i < n;
        // 1. killed by a corner accepted in an earlier chunk? (the reference's own test: a few accepted corners per cell)
        bool live = valid;
        int xi = 0, yi = 0, cx0 = 0, cx1 = -1, cy0 = 0, cy1 = -1, mycell = 0;
        if (valid) {
            const unsigned p = pos[i];
            xi = (int)(p & 0xFFFFu); yi = (int)(p >> 16);
            const int xc = xi / cell, yc = yi / cell;
            mycell = yc * sel.gw + xc;
            cx0 = max(0, xc - 1); cx1 = min(sel.gw - 1, xc + 1); cy0 = max(0, yc - 1); cy1 = min(sel.gh - 1, yc + 1);
            for (int yy = cy0; yy <= cy1 && live; yy++)
                for (int xx = cx0; xx <= cx1 && live; xx++) {
                    const int c = yy * sel.gw + xx;
                    const int cnt = min(cnt8_get(acc_cnt, c), kSelAccSlots);
                    if (cnt == 0) continue;
                    const uint2 w = *reinterpret_cast<const uint2*>(acc_slot + c * kSelAccSlots);   // the cell's four slots in one load
                    const unsigned j[4] = {w.x & 0xFFFFu, w.x >> 16, w.y & 0xFFFFu, w.y >> 16};
                    unsigned q[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) q[u] = pos[u < cnt ? (j[u] & (kSortCap - 1)) : 0];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int dx = xi - (int)(q[u] & 0xFFFFu), dy = yi - (int)(q[u] >> 16);
                        if (u < cnt && dx * dx + dy * dy <= sel.d2max) live = false;
                    }
                }
            // 2. the live candidates of this chunk are listed by cell
            if (live) {
                const int slot = cnt8_inc(ch_cnt, mycell);
                if (slot < kSelChunkSlots) ch_slot[mycell * kSelChunkSlots + slot] = (unsigned short)tid;
                else {   // a crowded cell (many strong local maxima side by side): the candidate goes to a chunk-wide list every live candidate also walks
                    const int k = atomicAdd(&s_warp[36], 1);
                    if (k < kSelSpill) spill[k] = (unsigned short)tid; else atomicOr(&s_warp[35], 1);
                }
            }
        }
        acc[tid] = live ? 1 : 0;
        __syncthreads();
#ifdef COEB_SEL_DEBUG
        const long long dbg_c1 = clock64();
#endif
        // 3. close live predecessors inside the chunk: the first kSelNb in the thread's own list, the rest (dense clusters of strong corners:
        // up to ~50 local maxima lie within minDistance of one) in a slice of a pool shared by the chunk, filled by a second identical walk
        int nnb = 0, pool_off = -1;
        bool near_spill = false;   // one of the nine cells was full: only then can a spilled candidate be close
        if (live) {
            auto walk = [&](auto&& emit) {
                for (int yy = cy0; yy <= cy1; yy++)
                    for (int xx = cx0; xx <= cx1; xx++) {
                        const int c = yy * sel.gw + xx;
                        const int raw = cnt8_get(ch_cnt, c);
                        near_spill |= raw > kSelChunkSlots;
                        const int cnt = min(raw, kSelChunkSlots);
                        const uint2* const row = reinterpret_cast<const uint2*>(ch_slot + c * kSelChunkSlots);
                        for (int k0 = 0; k0 < cnt; k0 += 4) {   // four slots at a time: one load, four independent position reads (in slot order)
                            const uint2 w = row[k0 >> 2];
                            const int tj[4] = {(int)(w.x & 0xFFFFu), (int)(w.x >> 16), (int)(w.y & 0xFFFFu), (int)(w.y >> 16)};
                            bool ok[4];
                            unsigned q[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) { ok[u] = k0 + u < cnt && tj[u] < tid; q[u] = pos[base + (ok[u] ? tj[u] : 0)]; }
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int dx = xi - (int)(q[u] & 0xFFFFu), dy = yi - (int)(q[u] >> 16);
                                if (ok[u] && dx * dx + dy * dy <= sel.d2max) emit(tj[u]);
                            }
                        }
                    }
                const int nsp = near_spill ? min(s_warp[36], kSelSpill) : 0;
                for (int k = 0; k < nsp; k++) {
                    const int tj = spill[k];
                    if (tj >= tid) continue;
                    const unsigned q = pos[base + tj];
                    const int dx = xi - (int)(q & 0xFFFFu), dy = yi - (int)(q >> 16);
                    if (dx * dx + dy * dy <= sel.d2max) emit(tj);
                }
            };
            walk([&](int tj) { if (nnb < kSelNb) nbl[nnb] = (unsigned short)tj; nnb++; });
            if (nnb > kSelNb) {
                const int extra = nnb - kSelNb, off = atomicAdd(&s_warp[37], extra);
                if (off + extra <= kSelPool) {   // (else: the cells are walked again in every sweep)
                    pool_off = off;
                    int idx = 0;
                    walk([&](int tj) { if (idx >= kSelNb) pool[off + idx - kSelNb] = (unsigned short)tj; idx++; });
                }
            }
        }
        // 4. accepted(i) = no accepted close predecessor, iterated in place until a sweep changes nothing
        bool a = live;
#ifdef COEB_SEL_DEBUG
        int dbg_sweeps = 0;
        const long long dbg_c2 = (__syncthreads(), clock64());
        const int dbg_live = __syncthreads_count(live), dbg_over = __syncthreads_count(nnb > kSelNb && pool_off < 0), dbg_nsp = __syncthreads_count(near_spill);
#endif
        for (;;) {
#ifdef COEB_SEL_DEBUG
            dbg_sweeps++;
#endif
            bool changed = false;
            if (live) {
                bool now = true;
                if (nnb <= kSelNb || pool_off >= 0) {
                    const int n1 = min(nnb, kSelNb);
                    const uint2* const nrow = reinterpret_cast<const uint2*>(nbl);
                    unsigned hit = 0u;
                    for (int k0 = 0; k0 < n1; k0 += 4) {   // (entries past n1 are stale: masked into range, their flags ignored)
                        const uint2 w = nrow[k0 >> 2];
                        const unsigned a0 = acc[w.x & 1023u], a1 = acc[(w.x >> 16) & 1023u], a2 = acc[w.y & 1023u], a3 = acc[(w.y >> 16) & 1023u];
                        hit |= a0 | (k0 + 1 < n1 ? a1 : 0u) | (k0 + 2 < n1 ? a2 : 0u) | (k0 + 3 < n1 ? a3 : 0u);
                    }
                    now = hit == 0u;
                    for (int k = 0; k < nnb - kSelNb && now; k++) if (acc[pool[pool_off + k]]) now = false;
                } else {
                    for (int yy = cy0; yy <= cy1 && now; yy++)
                        for (int xx = cx0; xx <= cx1 && now; xx++) {
                            const int c = yy * sel.gw + xx;
                            const int cnt = min(cnt8_get(ch_cnt, c), kSelChunkSlots);
                            for (int k = 0; k < cnt; k++) {
                                const int tj = ch_slot[c * kSelChunkSlots + k];
                                if (tj >= tid || !acc[tj]) continue;
                                const unsigned q = pos[base + tj];
                                const int dx = xi - (int)(q & 0xFFFFu), dy = yi - (int)(q >> 16);
                                if (dx * dx + dy * dy <= sel.d2max) { now = false; break; }
                            }
                        }
                    const int nsp = near_spill ? min(s_warp[36], kSelSpill) : 0;
                    for (int k = 0; k < nsp && now; k++) {
                        const int tj = spill[k];
                        if (tj >= tid || !acc[tj]) continue;
                        const unsigned q = pos[base + tj];
                        const int dx = xi - (int)(q & 0xFFFFu), dy = yi - (int)(q >> 16);
                        if (dx * dx + dy * dy <= sel.d2max) now = false;
                    }
                }
                if (now != a) { a = now; acc[tid] = now ? 1 : 0; changed = true; }
            }
            if (!__syncthreads_or(changed)) break;
        }
#ifdef COEB_SEL_DEBUG
        if (tid == 0) printf("[sel] n %d base %d live %d sweeps %d over-list %d near-spill %d spilled %d accepted-before %d | cycles: accepted grid + listing %lld, walk %lld, sweeps %lld\n", n, base, dbg_live, dbg_sweeps, dbg_over, dbg_nsp, s_warp[36], s_warp[34], dbg_c1 - dbg_c0, dbg_c2 - dbg_c1, clock64() - dbg_c2);
#endif
        // 5. the accepted ones join the accepted grid and are emitted in sorted order
        if (a) {
            const int slot = cnt8_inc(acc_cnt, mycell);
            if (slot < kSelAccSlots) acc_slot[mycell * kSelAccSlots + slot] = (unsigned short)i;
            else atomicOr(&s_warp[35], 4);
        }
        if (live) cnt8_clear(ch_cnt, mycell);   // (every listed candidate clears its cell: ready for the next chunk)
        const unsigned bal = __ballot_sync(0xffffffffu, a);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        if (wid == 0) {
            const int v = s_warp[lane];
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t2 = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t2; }
            s_warp[lane] = inc - v;               // exclusive prefix of the warp totals
            if (lane == 31) s_warp[33] = inc;     // chunk total
        }
        __syncthreads();
        const int before = s_warp[34];
        if (a) {
            const int rank = before + s_warp[wid] + __popc(bal & ((1u << lane) - 1u));
            if (rank < sel.max_corners) { const float2 c = make_float2((float)xi, (float)yi); sel.out_dev[rank] = c; sel.out_host[rank] = c; }
        }
        __syncthreads();
        if (tid == 0) { s_warp[34] = before + s_warp[33]; s_warp[36] = 0; s_warp[37] = 0; }
        __syncthreads();
    }
    if (tid == 0) {
        const int k = s_warp[35] ? -s_warp[35] : min(s_warp[34], sel.max_corners);   // (negative: which list overflowed for good, 1 spill list | 4 accepted cell; 8: not attempted)   // a cell list overflowed (never seen: accepted corners are minDistance apart): host pass
        *sel.n_dev = k; *sel.n_host = k;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// cv::cornerSubPix(src, corners, win (hw, hw), zeroZone (-1,-1), criteria): one warp per corner. Each iteration samples a
// (2hw+3)^2 patch around the current estimate (cv::getRectSubPix: float bilinear weights, replicated border) and solves the
// 2x2 gradient-weighted system in double.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kSpMaxWin = 23;   // 2 * 10 + 3
template <int HW>
__global__ void __launch_bounds__(128) corner_subpix_kernel(const uint8_t* __restrict__ img, int w, int h, int pitch, float2* __restrict__ pts, int n,
                                                            int max_iters, double eps2, const float* __restrict__ mask) {
    constexpr int hw = HW, win = 2 * HW + 1, pw = win + 2;
    constexpr int NP = (pw * pw + 31) / 32, NG = (win * win + 31) / 32;   // patch samples / gradient terms per lane
    __shared__ float s_patch[4][pw * pw];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int p = blockIdx.x * 4 + wid;
    if (p >= n) return;
    float* patch = s_patch[wid];
    // this lane's share of the window, fixed over the iterations: sample offsets, and for the gradient terms the patch index
    // and the Gaussian weight
    int soff[NP], gidx[NG];
    float gm[NG];
#pragma unroll
    for (int k = 0; k < NP; k++) { const int i = lane + 32 * k, py = i / pw; soff[k] = py * pitch + (i - py * pw); }
#pragma unroll
    for (int k = 0; k < NG; k++) {
        const int i = lane + 32 * k, yy = i / win, xx = i - yy * win;
        gidx[k] = (yy + 1) * pw + (xx + 1);
        gm[k] = i < win * win ? mask[i] : 0.f;
    }
    const float2 cT = pts[p];
    float2 cI = cT;
    int iter = 0;
    double err = 0;
    do {
        // getRectSubPix(src, (pw, pw), cI, patch, CV_32F)
        const float cxf = cI.x - (float)(pw - 1) * 0.5f, cyf = cI.y - (float)(pw - 1) * 0.5f;
        const int ipx = (int)floorf(cxf), ipy = (int)floorf(cyf);
        const float a = cxf - (float)ipx, b = cyf - (float)ipy;
        const float a11 = __fmul_rn(1.f - a, 1.f - b), a12 = __fmul_rn(a, 1.f - b), a21 = __fmul_rn(1.f - a, b), a22 = __fmul_rn(a, b);
        if (ipx >= 0 && ipy >= 0 && ipx + pw + 1 <= w && ipy + pw + 1 <= h) {   // the patch lies inside the image
            const uint8_t* base = img + (size_t)ipy * pitch + ipx;
#pragma unroll
            for (int k = 0; k < NP; k++) {
                const int i = lane + 32 * k;
                if (i < pw * pw) {
                    const uint8_t* q = base + soff[k];
                    patch[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn((float)q[0], a11), __fmul_rn((float)q[1], a12)), __fmul_rn((float)q[pitch], a21)), __fmul_rn((float)q[pitch + 1], a22));
                }
            }
        } else {
            for (int i = lane; i < pw * pw; i += 32) {
                const int py = i / pw, px = i - py * pw;
                const int x0 = min(max(ipx + px, 0), w - 1), x1 = min(max(ipx + px + 1, 0), w - 1);
                const int y0 = min(max(ipy + py, 0), h - 1), y1 = min(max(ipy + py + 1, 0), h - 1);
                const float v00 = img[(size_t)y0 * pitch + x0], v01 = img[(size_t)y0 * pitch + x1], v10 = img[(size_t)y1 * pitch + x0], v11 = img[(size_t)y1 * pitch + x1];
                patch[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00, a11), __fmul_rn(v01, a12)), __fmul_rn(v10, a21)), __fmul_rn(v11, a22));
            }
        }
        __syncwarp();
        double sa = 0, sb = 0, sc = 0, sbb1 = 0, sbb2 = 0;
#pragma unroll
        for (int k = 0; k < NG; k++) {
            const int i = lane + 32 * k;
            if (i < win * win) {
                const float* sp = patch + gidx[k];
                const int yy = i / win, xx = i - yy * win;
                const double m = gm[k];
                const double tgx = (double)__fsub_rn(sp[1], sp[-1]), tgy = (double)__fsub_rn(sp[pw], sp[-pw]);
                const double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m;
                const double px = xx - hw, py = yy - hw;
                sa += gxx; sb += gxy; sc += gyy;
                sbb1 += gxx * px + gxy * py;
                sbb2 += gxy * px + gyy * py;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); sc += __shfl_xor_sync(0xffffffffu, sc, o);
            sbb1 += __shfl_xor_sync(0xffffffffu, sbb1, o); sbb2 += __shfl_xor_sync(0xffffffffu, sbb2, o);
        }
        __syncwarp();
        const double det = sa * sc - sb * sb;
        if (fabs(det) <= DBL_EPSILON * DBL_EPSILON) break;
        const double scale = 1.0 / det;
        float2 cI2;
        cI2.x = (float)((double)cI.x + sc * scale * sbb1 - sb * scale * sbb2);
        cI2.y = (float)((double)cI.y - sb * scale * sbb1 + sa * scale * sbb2);
        err = ((double)cI2.x - cI.x) * ((double)cI2.x - cI.x) + ((double)cI2.y - cI.y) * ((double)cI2.y - cI.y);
        cI = cI2;
        if (cI.x < 0 || cI.x >= (float)w || cI.y < 0 || cI.y >= (float)h) break;
    } while (++iter < max_iters && err > eps2);
    // poor convergence: the initial point stays
    if (fabsf(cI.x - cT.x) > (float)hw || fabsf(cI.y - cT.y) > (float)hw) cI = cT;
    if (lane == 0) pts[p] = cI;
}

// The same iteration with one CTA (kPtWarps warps) per corner: a frame has ~600 corners and the device 148 SMs, so one warp per corner left
// four warps per SM, each walking 17 patch samples and 14 gradient terms per lane and iteration with nothing to hide their latency
// behind (57 us for 591 corners). With the window spread over 128 threads an iteration is 5 samples and 4 terms per thread, one block
// barrier for the patch and one for the five sums (per-warp partials in shared memory, added in warp order by every thread, so all threads
// hold the same doubles and take the same branches).
constexpr int kPtWarps = 4;   // measured: 8 warps per point are slower again (subpix + LK 146 us against 122 us, 187 us with one)
template <int HW>
__global__ void __launch_bounds__(32 * kPtWarps) corner_subpix_cta_kernel(const uint8_t* __restrict__ img, int w, int h, int pitch, float2* __restrict__ pts, int n,
                                                                         int max_iters, double eps2, const float* __restrict__ mask, float2* __restrict__ pts_host,
                                                                         const float2* __restrict__ pts_in /* initial corners, if not pts (mapped host memory) */,
                                                                         const int* __restrict__ n_ptr /* the number of corners, if only the device knows it yet */) {
    constexpr int hw = HW, win = 2 * HW + 1, pw = win + 2, T = 32 * kPtWarps;
    if (n_ptr) n = *n_ptr;
    constexpr int NP = (pw * pw + T - 1) / T, NG = (win * win + T - 1) / T;   // patch samples / gradient terms per thread
    __shared__ float patch[pw * pw];
    __shared__ double s_red[2][kPtWarps][5];
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int p = blockIdx.x;
    if (p >= n) return;
    int soff[NP], gidx[NG];
    float gm[NG];
#pragma unroll
    for (int k = 0; k < NP; k++) { const int i = t + T * k, py = i / pw; soff[k] = py * pitch + (i - py * pw); }
#pragma unroll
    for (int k = 0; k < NG; k++) {
        const int i = t + T * k, yy = i / win, xx = i - yy * win;
        gidx[k] = (yy + 1) * pw + (xx + 1);
        gm[k] = i < win * win ? mask[i] : 0.f;
    }
    const float2 cT = pts_in ? pts_in[p] : pts[p];
    float2 cI = cT;
    int iter = 0, par = 0;
    double err = 0;
    do {
        // getRectSubPix(src, (pw, pw), cI, patch, CV_32F)
        const float cxf = cI.x - (float)(pw - 1) * 0.5f, cyf = cI.y - (float)(pw - 1) * 0.5f;
        const int ipx = (int)floorf(cxf), ipy = (int)floorf(cyf);
        const float a = cxf - (float)ipx, b = cyf - (float)ipy;
        const float a11 = __fmul_rn(1.f - a, 1.f - b), a12 = __fmul_rn(a, 1.f - b), a21 = __fmul_rn(1.f - a, b), a22 = __fmul_rn(a, b);
        if (ipx >= 0 && ipy >= 0 && ipx + pw + 1 <= w && ipy + pw + 1 <= h) {   // the patch lies inside the image
            const uint8_t* base = img + (size_t)ipy * pitch + ipx;
#pragma unroll
            for (int k = 0; k < NP; k++) {
                const int i = t + T * k;
                if (i < pw * pw) {
                    const uint8_t* q = base + soff[k];
                    patch[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn((float)q[0], a11), __fmul_rn((float)q[1], a12)), __fmul_rn((float)q[pitch], a21)), __fmul_rn((float)q[pitch + 1], a22));
                }
            }
        } else {
            for (int i = t; i < pw * pw; i += T) {
                const int py = i / pw, px = i - py * pw;
                const int x0 = min(max(ipx + px, 0), w - 1), x1 = min(max(ipx + px + 1, 0), w - 1);
                const int y0 = min(max(ipy + py, 0), h - 1), y1 = min(max(ipy + py + 1, 0), h - 1);
                const float v00 = img[(size_t)y0 * pitch + x0], v01 = img[(size_t)y0 * pitch + x1], v10 = img[(size_t)y1 * pitch + x0], v11 = img[(size_t)y1 * pitch + x1];
                patch[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00, a11), __fmul_rn(v01, a12)), __fmul_rn(v10, a21)), __fmul_rn(v11, a22));
            }
        }
        __syncthreads();
        double sa = 0, sb = 0, sc = 0, sbb1 = 0, sbb2 = 0;
#pragma unroll
        for (int k = 0; k < NG; k++) {
            const int i = t + T * k;
            if (i < win * win) {
                const float* sp = patch + gidx[k];
                const int yy = i / win, xx = i - yy * win;
                const double m = gm[k];
                const double tgx = (double)__fsub_rn(sp[1], sp[-1]), tgy = (double)__fsub_rn(sp[pw], sp[-pw]);
                const double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m;
                const double px = xx - hw, py = yy - hw;
                sa += gxx; sb += gxy; sc += gyy;
                sbb1 += gxx * px + gxy * py;
                sbb2 += gxy * px + gyy * py;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); sc += __shfl_xor_sync(0xffffffffu, sc, o);
            sbb1 += __shfl_xor_sync(0xffffffffu, sbb1, o); sbb2 += __shfl_xor_sync(0xffffffffu, sbb2, o);
        }
        if (lane == 0) { double* r = s_red[par][wid]; r[0] = sa; r[1] = sb; r[2] = sc; r[3] = sbb1; r[4] = sbb2; }
        __syncthreads();   // (also: every thread has read the patch before the next iteration overwrites it)
        sa = sb = sc = sbb1 = sbb2 = 0;
#pragma unroll
        for (int g = 0; g < kPtWarps; g++) { const double* r = s_red[par][g]; sa += r[0]; sb += r[1]; sc += r[2]; sbb1 += r[3]; sbb2 += r[4]; }
        par ^= 1;
        const double det = sa * sc - sb * sb;
        if (fabs(det) <= DBL_EPSILON * DBL_EPSILON) break;
        const double scale = 1.0 / det;
        float2 cI2;
        cI2.x = (float)((double)cI.x + sc * scale * sbb1 - sb * scale * sbb2);
        cI2.y = (float)((double)cI.y - sb * scale * sbb1 + sa * scale * sbb2);
        err = ((double)cI2.x - cI.x) * ((double)cI2.x - cI.x) + ((double)cI2.y - cI.y) * ((double)cI2.y - cI.y);
        cI = cI2;
        if (cI.x < 0 || cI.x >= (float)w || cI.y < 0 || cI.y >= (float)h) break;
    } while (++iter < max_iters && err > eps2);
    // poor convergence: the initial point stays
    if (fabsf(cI.x - cT.x) > (float)hw || fabsf(cI.y - cT.y) > (float)hw) cI = cT;
    if (t == 0) {
        pts[p] = cI;
        if (pts_host) pts_host[p] = cI;   // mapped pinned copy for the caller: no device-to-host copy node
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// cv::pyrDown (8-bit): 5x5 binomial [1 4 6 4 1] x [1 4 6 4 1], (sum + 128) >> 8, BORDER_REFLECT_101; size ((w+1)/2, (h+1)/2).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pyr_down_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch, uint8_t* __restrict__ dst, int dw, int dh, int dpitch) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= dw || y >= dh) return;
    const int wgt[5] = {1, 4, 6, 4, 1};
    int sum = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const uint8_t* row = src + (size_t)refl101(2 * y + j - 2, sh) * spitch;
        int rs = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) rs += wgt[i] * row[refl101(2 * x + i - 2, sw)];
        sum += wgt[j] * rs;
    }
    dst[(size_t)y * dpitch + x] = (uint8_t)((sum + 128) >> 8);
}

// calcScharrDeriv (video/lkpyramid.cpp): Ix = [3 10 3]^T x [-1 0 1], Iy = [-1 0 1]^T x [3 10 3], int16, reflected borders.
__global__ void __launch_bounds__(256) scharr_kernel(const uint8_t* __restrict__ src, int w, int h, int pitch, short2* __restrict__ d) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const int xm = refl101(x - 1, w), xp = refl101(x + 1, w);
    const uint8_t *r0 = src + (size_t)refl101(y - 1, h) * pitch, *r1 = src + (size_t)y * pitch, *r2 = src + (size_t)refl101(y + 1, h) * pitch;
    auto t0 = [&](int xx) { return ((int)r0[xx] + (int)r2[xx]) * 3 + (int)r1[xx] * 10; };   // vertical smoothing
    auto t1 = [&](int xx) { return (int)r2[xx] - (int)r0[xx]; };                             // vertical difference
    const int ix = t0(xp) - t0(xm);
    const int iy = (t1(xp) + t1(xm)) * 3 + t1(x) * 10;
    d[(size_t)y * w + x] = make_short2((short)ix, (short)iy);
}

// ---------------------------------------------------------------------------------------------------------------------
// LKTrackerInvoker (video/lkpyramid.cpp) for every pyramid level, coarsest first, one warp per point. The image windows read
// the level through BORDER_REFLECT_101 (the reference pyramid carries a winSize border), the derivative window reads zeros
// outside the level (BORDER_CONSTANT). Bilinear weights are Q14 integers, the patch and its derivatives are int16, the sums
// are float: exactly the reference's types.
// ---------------------------------------------------------------------------------------------------------------------
struct LkLevels {
    int nlevels;                       // levels actually built (maxLevel + 1)
    int w[kMoMaxLevels], h[kMoMaxLevels], pitch[kMoMaxLevels];
    const uint8_t* prev[kMoMaxLevels];
    const uint8_t* cur[kMoMaxLevels];
    const short2* deriv[kMoMaxLevels];
};
constexpr int kLkMaxWin = 32;
constexpr int kLkWarps = 4;

__global__ void __launch_bounds__(32 * kLkWarps) lk_kernel(const __grid_constant__ LkLevels L, const float2* __restrict__ prev_pts, int n, int win, int max_iters,
                                                           float eps2, float min_eig_thr, float2* __restrict__ next_pts, uint8_t* __restrict__ status,
                                                           int edge, float sad_limit) {
    __shared__ short s_I[kLkWarps][kLkMaxWin * kLkMaxWin], s_Ix[kLkWarps][kLkMaxWin * kLkMaxWin], s_Iy[kLkWarps][kLkMaxWin * kLkMaxWin];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int p = blockIdx.x * kLkWarps + wid;
    if (p >= n) return;
    short *Ib = s_I[wid], *Ixb = s_Ix[wid], *Iyb = s_Iy[wid];
    const float2 pt0 = prev_pts[p];
    const float half = (float)(win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    float2 nextPt = make_float2(0.f, 0.f);
    bool st = true;
    for (int level = L.nlevels - 1; level >= 0; level--) {
        const int w = L.w[level], h = L.h[level], pitch = L.pitch[level];
        const uint8_t *I = L.prev[level], *J = L.cur[level];
        const short2* D = L.deriv[level];
        const float sc = (float)(1. / (double)(1 << level));
        float2 prevPt = make_float2(__fmul_rn(pt0.x, sc), __fmul_rn(pt0.y, sc));
        if (level == L.nlevels - 1) nextPt = prevPt;
        else nextPt = make_float2(__fmul_rn(nextPt.x, 2.f), __fmul_rn(nextPt.y, 2.f));
        prevPt.x = __fsub_rn(prevPt.x, half); prevPt.y = __fsub_rn(prevPt.y, half);
        const int ipx = (int)floorf(prevPt.x), ipy = (int)floorf(prevPt.y);
        if (ipx < -win || ipx >= w || ipy < -win || ipy >= h) {
            if (level == 0) st = false;
            continue;
        }
        float a = __fsub_rn(prevPt.x, (float)ipx), b = __fsub_rn(prevPt.y, (float)ipy);
        int iw00 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, 1.f - b), 16384.f));
        int iw01 = __float2int_rn(__fmul_rn(__fmul_rn(a, 1.f - b), 16384.f));
        int iw10 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, b), 16384.f));
        int iw11 = 16384 - iw00 - iw01 - iw10;
        float A11 = 0, A12 = 0, A22 = 0;
        for (int i = lane; i < win * win; i += 32) {
            const int y = i / win, x = i - y * win;
            const int gx0 = ipx + x, gy0 = ipy + y;
            const int x0 = refl101(gx0, w), x1 = refl101(gx0 + 1, w), y0 = refl101(gy0, h), y1 = refl101(gy0 + 1, h);
            const int ival = ((int)I[(size_t)y0 * pitch + x0] * iw00 + (int)I[(size_t)y0 * pitch + x1] * iw01 + (int)I[(size_t)y1 * pitch + x0] * iw10 +
                              (int)I[(size_t)y1 * pitch + x1] * iw11 + (1 << 8)) >> 9;
            auto dv = [&](int gx, int gy) { return ((unsigned)gx < (unsigned)w && (unsigned)gy < (unsigned)h) ? D[(size_t)gy * w + gx] : make_short2(0, 0); };
            const short2 d00 = dv(gx0, gy0), d01 = dv(gx0 + 1, gy0), d10 = dv(gx0, gy0 + 1), d11 = dv(gx0 + 1, gy0 + 1);
            const int ixval = ((int)d00.x * iw00 + (int)d01.x * iw01 + (int)d10.x * iw10 + (int)d11.x * iw11 + (1 << 13)) >> 14;
            const int iyval = ((int)d00.y * iw00 + (int)d01.y * iw01 + (int)d10.y * iw10 + (int)d11.y * iw11 + (1 << 13)) >> 14;
            Ib[i] = (short)ival; Ixb[i] = (short)ixval; Iyb[i] = (short)iyval;
            A11 += (float)(ixval * ixval); A12 += (float)(ixval * iyval); A22 += (float)(iyval * iyval);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            A11 += __shfl_xor_sync(0xffffffffu, A11, o); A12 += __shfl_xor_sync(0xffffffffu, A12, o); A22 += __shfl_xor_sync(0xffffffffu, A22, o);
        }
        __syncwarp();
        A11 *= FLT_SCALE; A12 *= FLT_SCALE; A22 *= FLT_SCALE;
        float Dt = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA = __fsub_rn(A11, A22);
        const float minEig = (A22 + A11 - sqrtf(__fadd_rn(__fmul_rn(dA, dA), __fmul_rn(4.f, __fmul_rn(A12, A12))))) / (float)(2 * win * win);
        if (minEig < min_eig_thr || Dt < FLT_EPSILON) {
            if (level == 0) st = false;
            continue;
        }
        Dt = 1.f / Dt;
        nextPt.x = __fsub_rn(nextPt.x, half); nextPt.y = __fsub_rn(nextPt.y, half);
        float2 prevDelta = make_float2(0.f, 0.f);
        float2 result = make_float2(__fadd_rn(nextPt.x, half), __fadd_rn(nextPt.y, half));
        for (int j = 0; j < max_iters; j++) {
            const int inx = (int)floorf(nextPt.x), iny = (int)floorf(nextPt.y);
            if (inx < -win || inx >= w || iny < -win || iny >= h) {
                if (level == 0) st = false;
                break;
            }
            a = __fsub_rn(nextPt.x, (float)inx); b = __fsub_rn(nextPt.y, (float)iny);
            iw00 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, 1.f - b), 16384.f));
            iw01 = __float2int_rn(__fmul_rn(__fmul_rn(a, 1.f - b), 16384.f));
            iw10 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, b), 16384.f));
            iw11 = 16384 - iw00 - iw01 - iw10;
            float b1 = 0, b2 = 0;
            for (int i = lane; i < win * win; i += 32) {
                const int y = i / win, x = i - y * win;
                const int x0 = refl101(inx + x, w), x1 = refl101(inx + x + 1, w), y0 = refl101(iny + y, h), y1 = refl101(iny + y + 1, h);
                const int jv = ((int)J[(size_t)y0 * pitch + x0] * iw00 + (int)J[(size_t)y0 * pitch + x1] * iw01 + (int)J[(size_t)y1 * pitch + x0] * iw10 +
                                (int)J[(size_t)y1 * pitch + x1] * iw11 + (1 << 8)) >> 9;
                const int diff = jv - (int)Ib[i];
                b1 += (float)(diff * (int)Ixb[i]);
                b2 += (float)(diff * (int)Iyb[i]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { b1 += __shfl_xor_sync(0xffffffffu, b1, o); b2 += __shfl_xor_sync(0xffffffffu, b2, o); }
            b1 *= FLT_SCALE; b2 *= FLT_SCALE;
            const float2 delta = make_float2(__fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), Dt), __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), Dt));
            nextPt.x = __fadd_rn(nextPt.x, delta.x); nextPt.y = __fadd_rn(nextPt.y, delta.y);
            result = make_float2(__fadd_rn(nextPt.x, half), __fadd_rn(nextPt.y, half));
            if ((double)delta.x * delta.x + (double)delta.y * delta.y <= (double)eps2) break;
            if (j > 0 && fabsf(delta.x + prevDelta.x) < 0.01f && fabsf(delta.y + prevDelta.y) < 0.01f) {
                result.x = __fsub_rn(result.x, __fmul_rn(delta.x, 0.5f)); result.y = __fsub_rn(result.y, __fmul_rn(delta.y, 0.5f));
                break;
            }
            prevDelta = delta;
        }
        nextPt = result;
        __syncwarp();
    }
    if (lane == 0) {
        next_pts[p] = nextPt;
        // src/Frame.cc:336-364: 5-px border test on the truncated coordinates, then the 3x3 sum of absolute differences
        if (st && edge >= 0) {
            const int w = L.w[0], h = L.h[0], pitch = L.pitch[0];
            const int x1 = (int)pt0.x, y1 = (int)pt0.y, x2 = (int)nextPt.x, y2 = (int)nextPt.y;
            if (x1 < edge || x1 >= w - edge || x2 < edge || x2 >= w - edge || y1 < edge || y1 >= h - edge || y2 < edge || y2 >= h - edge) {
                st = false;
            } else {
                int sad = 0;
                for (int j = -1; j <= 1; j++)
                    for (int i = -1; i <= 1; i++) sad += abs((int)L.prev[0][(size_t)(y1 + j) * pitch + x1 + i] - (int)L.cur[0][(size_t)(y2 + j) * pitch + x2 + i]);
                if ((float)sad > sad_limit) st = false;
            }
        }
        status[p] = st ? 1 : 0;
    }
}

// The same tracker for a compile-time window (the reference's 22 x 22): the patch of the first image and its derivatives stay in
// registers (NS = ceil(WIN^2 / 32) pixels per lane), and a window that lies inside the level is read through per-lane offsets
// computed once per level, without the per-pixel reflection and index arithmetic of the general kernel above. Same operations in
// the same order per lane, so both kernels give the same floats.
template <int WIN>
__global__ void __launch_bounds__(32 * kLkWarps) lk_kernel_w(const __grid_constant__ LkLevels L, const float2* __restrict__ prev_pts, int n, int max_iters,
                                                           float eps2, float min_eig_thr, float2* __restrict__ next_pts, uint8_t* __restrict__ status,
                                                           int edge, float sad_limit) {
    constexpr int win = WIN, NS = (WIN * WIN + 31) / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int p = blockIdx.x * kLkWarps + wid;
    if (p >= n) return;
    int Ib[NS], Ixb[NS], Iyb[NS], off[NS];   // this lane's pixels i = lane + 32 k of the window: patch value, derivatives, y * pitch + x
    const float2 pt0 = prev_pts[p];
    const float half = (float)(win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    float2 nextPt = make_float2(0.f, 0.f);
    bool st = true;
    for (int level = L.nlevels - 1; level >= 0; level--) {
        const int w = L.w[level], h = L.h[level], pitch = L.pitch[level];
        const uint8_t *I = L.prev[level], *J = L.cur[level];
        const short2* D = L.deriv[level];
#pragma unroll
        for (int k = 0; k < NS; k++) { const int i = lane + 32 * k, y = i / WIN; off[k] = y * pitch + (i - y * WIN); }
        const float sc = (float)(1. / (double)(1 << level));
        float2 prevPt = make_float2(__fmul_rn(pt0.x, sc), __fmul_rn(pt0.y, sc));
        if (level == L.nlevels - 1) nextPt = prevPt;
        else nextPt = make_float2(__fmul_rn(nextPt.x, 2.f), __fmul_rn(nextPt.y, 2.f));
        prevPt.x = __fsub_rn(prevPt.x, half); prevPt.y = __fsub_rn(prevPt.y, half);
        const int ipx = (int)floorf(prevPt.x), ipy = (int)floorf(prevPt.y);
        if (ipx < -win || ipx >= w || ipy < -win || ipy >= h) {
            if (level == 0) st = false;
            continue;
        }
        float a = __fsub_rn(prevPt.x, (float)ipx), b = __fsub_rn(prevPt.y, (float)ipy);
        int iw00 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, 1.f - b), 16384.f));
        int iw01 = __float2int_rn(__fmul_rn(__fmul_rn(a, 1.f - b), 16384.f));
        int iw10 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, b), 16384.f));
        int iw11 = 16384 - iw00 - iw01 - iw10;
        float A11 = 0, A12 = 0, A22 = 0;
        const bool insideI = ipx >= 0 && ipy >= 0 && ipx + WIN + 1 <= w && ipy + WIN + 1 <= h;
#pragma unroll
        for (int k = 0; k < NS; k++) {
            const int i = lane + 32 * k;
            Ib[k] = Ixb[k] = Iyb[k] = 0;
            if (i < WIN * WIN) {
                int ival, ixval, iyval;
                if (insideI) {
                    const uint8_t* q = I + (size_t)ipy * pitch + ipx + off[k];
                    ival = ((int)q[0] * iw00 + (int)q[1] * iw01 + (int)q[pitch] * iw10 + (int)q[pitch + 1] * iw11 + (1 << 8)) >> 9;
                    const int y = i / WIN, x = i - y * WIN;
                    const short2* dq = D + (size_t)(ipy + y) * w + ipx + x;
                    const short2 d00 = dq[0], d01 = dq[1], d10 = dq[w], d11 = dq[w + 1];
                    ixval = ((int)d00.x * iw00 + (int)d01.x * iw01 + (int)d10.x * iw10 + (int)d11.x * iw11 + (1 << 13)) >> 14;
                    iyval = ((int)d00.y * iw00 + (int)d01.y * iw01 + (int)d10.y * iw10 + (int)d11.y * iw11 + (1 << 13)) >> 14;
                } else {
                    const int y = i / WIN, x = i - y * WIN;
                    const int gx0 = ipx + x, gy0 = ipy + y;
                    const int x0 = refl101(gx0, w), x1 = refl101(gx0 + 1, w), y0 = refl101(gy0, h), y1 = refl101(gy0 + 1, h);
                    ival = ((int)I[(size_t)y0 * pitch + x0] * iw00 + (int)I[(size_t)y0 * pitch + x1] * iw01 + (int)I[(size_t)y1 * pitch + x0] * iw10 +
                            (int)I[(size_t)y1 * pitch + x1] * iw11 + (1 << 8)) >> 9;
                    auto dv = [&](int gx, int gy) { return ((unsigned)gx < (unsigned)w && (unsigned)gy < (unsigned)h) ? D[(size_t)gy * w + gx] : make_short2(0, 0); };
                    const short2 d00 = dv(gx0, gy0), d01 = dv(gx0 + 1, gy0), d10 = dv(gx0, gy0 + 1), d11 = dv(gx0 + 1, gy0 + 1);
                    ixval = ((int)d00.x * iw00 + (int)d01.x * iw01 + (int)d10.x * iw10 + (int)d11.x * iw11 + (1 << 13)) >> 14;
                    iyval = ((int)d00.y * iw00 + (int)d01.y * iw01 + (int)d10.y * iw10 + (int)d11.y * iw11 + (1 << 13)) >> 14;
                }
                ival = (int)(short)ival; ixval = (int)(short)ixval; iyval = (int)(short)iyval;   // the reference stores int16
                Ib[k] = ival; Ixb[k] = ixval; Iyb[k] = iyval;
                A11 += (float)(ixval * ixval); A12 += (float)(ixval * iyval); A22 += (float)(iyval * iyval);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            A11 += __shfl_xor_sync(0xffffffffu, A11, o); A12 += __shfl_xor_sync(0xffffffffu, A12, o); A22 += __shfl_xor_sync(0xffffffffu, A22, o);
        }
        A11 *= FLT_SCALE; A12 *= FLT_SCALE; A22 *= FLT_SCALE;
        float Dt = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA = __fsub_rn(A11, A22);
        const float minEig = (A22 + A11 - sqrtf(__fadd_rn(__fmul_rn(dA, dA), __fmul_rn(4.f, __fmul_rn(A12, A12))))) / (float)(2 * win * win);
        if (minEig < min_eig_thr || Dt < FLT_EPSILON) {
            if (level == 0) st = false;
            continue;
        }
        Dt = 1.f / Dt;
        nextPt.x = __fsub_rn(nextPt.x, half); nextPt.y = __fsub_rn(nextPt.y, half);
        float2 prevDelta = make_float2(0.f, 0.f);
        float2 result = make_float2(__fadd_rn(nextPt.x, half), __fadd_rn(nextPt.y, half));
        for (int j = 0; j < max_iters; j++) {
            const int inx = (int)floorf(nextPt.x), iny = (int)floorf(nextPt.y);
            if (inx < -win || inx >= w || iny < -win || iny >= h) {
                if (level == 0) st = false;
                break;
            }
            a = __fsub_rn(nextPt.x, (float)inx); b = __fsub_rn(nextPt.y, (float)iny);
            iw00 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, 1.f - b), 16384.f));
            iw01 = __float2int_rn(__fmul_rn(__fmul_rn(a, 1.f - b), 16384.f));
            iw10 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, b), 16384.f));
            iw11 = 16384 - iw00 - iw01 - iw10;
            float b1 = 0, b2 = 0;
            const bool insideJ = inx >= 0 && iny >= 0 && inx + WIN + 1 <= w && iny + WIN + 1 <= h;
            if (insideJ) {
                const uint8_t* base = J + (size_t)iny * pitch + inx;
#pragma unroll
                for (int k = 0; k < NS; k++)
                    if (lane + 32 * k < WIN * WIN) {
                        const uint8_t* q = base + off[k];
                        const int jv = ((int)q[0] * iw00 + (int)q[1] * iw01 + (int)q[pitch] * iw10 + (int)q[pitch + 1] * iw11 + (1 << 8)) >> 9;
                        const int diff = jv - Ib[k];
                        b1 += (float)(diff * Ixb[k]);
                        b2 += (float)(diff * Iyb[k]);
                    }
            } else {
#pragma unroll
                for (int k = 0; k < NS; k++) {
                    const int i = lane + 32 * k;
                    if (i < WIN * WIN) {
                        const int y = i / WIN, x = i - y * WIN;
                        const int x0 = refl101(inx + x, w), x1 = refl101(inx + x + 1, w), y0 = refl101(iny + y, h), y1 = refl101(iny + y + 1, h);
                        const int jv = ((int)J[(size_t)y0 * pitch + x0] * iw00 + (int)J[(size_t)y0 * pitch + x1] * iw01 + (int)J[(size_t)y1 * pitch + x0] * iw10 +
                                        (int)J[(size_t)y1 * pitch + x1] * iw11 + (1 << 8)) >> 9;
                        const int diff = jv - Ib[k];
                        b1 += (float)(diff * Ixb[k]);
                        b2 += (float)(diff * Iyb[k]);
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { b1 += __shfl_xor_sync(0xffffffffu, b1, o); b2 += __shfl_xor_sync(0xffffffffu, b2, o); }
            b1 *= FLT_SCALE; b2 *= FLT_SCALE;
            const float2 delta = make_float2(__fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), Dt), __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), Dt));
            nextPt.x = __fadd_rn(nextPt.x, delta.x); nextPt.y = __fadd_rn(nextPt.y, delta.y);
            result = make_float2(__fadd_rn(nextPt.x, half), __fadd_rn(nextPt.y, half));
            if ((double)delta.x * delta.x + (double)delta.y * delta.y <= (double)eps2) break;
            if (j > 0 && fabsf(delta.x + prevDelta.x) < 0.01f && fabsf(delta.y + prevDelta.y) < 0.01f) {
                result.x = __fsub_rn(result.x, __fmul_rn(delta.x, 0.5f)); result.y = __fsub_rn(result.y, __fmul_rn(delta.y, 0.5f));
                break;
            }
            prevDelta = delta;
        }
        nextPt = result;
        __syncwarp();
    }
    if (lane == 0) {
        next_pts[p] = nextPt;
        // src/Frame.cc:336-364: 5-px border test on the truncated coordinates, then the 3x3 sum of absolute differences
        if (st && edge >= 0) {
            const int w = L.w[0], h = L.h[0], pitch = L.pitch[0];
            const int x1 = (int)pt0.x, y1 = (int)pt0.y, x2 = (int)nextPt.x, y2 = (int)nextPt.y;
            if (x1 < edge || x1 >= w - edge || x2 < edge || x2 >= w - edge || y1 < edge || y1 >= h - edge || y2 < edge || y2 >= h - edge) {
                st = false;
            } else {
                int sad = 0;
                for (int j = -1; j <= 1; j++)
                    for (int i = -1; i <= 1; i++) sad += abs((int)L.prev[0][(size_t)(y1 + j) * pitch + x1 + i] - (int)L.cur[0][(size_t)(y2 + j) * pitch + x2 + i]);
                if ((float)sad > sad_limit) st = false;
            }
        }
        status[p] = st ? 1 : 0;
    }
}

// lk_kernel_w with one CTA (kPtWarps warps) per point, for the same reason as corner_subpix_cta_kernel: 591 points as 591 warps were four
// warps per SM, 16 window pixels per lane and iteration (83 us). Here a thread owns 4 pixels of the 22 x 22 window; the sums of a level
// (A11, A12, A22) and of an iteration (b1, b2) are warp-reduced, exchanged through shared memory and added in warp order by every thread,
// so the point's state is identical in all 128 threads and every branch is uniform.
__device__ __forceinline__ void cta_sum3(float& a, float& b, float& c, float (*s_red)[kPtWarps][3], int& par) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s_red[par][wid][0] = a; s_red[par][wid][1] = b; s_red[par][wid][2] = c; }
    __syncthreads();
    a = b = c = 0.f;
#pragma unroll
    for (int g = 0; g < kPtWarps; g++) { a += s_red[par][g][0]; b += s_red[par][g][1]; c += s_red[par][g][2]; }
    par ^= 1;   // the next exchange uses the other buffer: one barrier per exchange is enough
}

template <int WIN>
__global__ void __launch_bounds__(32 * kPtWarps) lk_cta_kernel(const __grid_constant__ LkLevels L, const float2* __restrict__ prev_pts, int n, int max_iters,
                                                              float eps2, float min_eig_thr, float2* __restrict__ next_pts, uint8_t* __restrict__ status,
                                                              int edge, float sad_limit, float2* __restrict__ next_host, uint8_t* __restrict__ status_host,
                                                              const int* __restrict__ n_ptr) {
    constexpr int win = WIN, T = 32 * kPtWarps, NS = (WIN * WIN + T - 1) / T;
    if (n_ptr) n = *n_ptr;
    __shared__ float s_red[2][kPtWarps][3];
    const int t = threadIdx.x;
    const int p = blockIdx.x;
    if (p >= n) return;
    int par = 0;
    int Ib[NS], Ixb[NS], Iyb[NS], off[NS];   // this thread's pixels i = t + T k of the window: patch value, derivatives, y * pitch + x
    const float2 pt0 = prev_pts[p];
    const float half = (float)(win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    float2 nextPt = make_float2(0.f, 0.f);
    bool st = true;
    for (int level = L.nlevels - 1; level >= 0; level--) {
        const int w = L.w[level], h = L.h[level], pitch = L.pitch[level];
        const uint8_t *I = L.prev[level], *J = L.cur[level];
        const short2* D = L.deriv[level];
#pragma unroll
        for (int k = 0; k < NS; k++) { const int i = t + T * k, y = i / WIN; off[k] = y * pitch + (i - y * WIN); }
        const float sc = (float)(1. / (double)(1 << level));
        float2 prevPt = make_float2(__fmul_rn(pt0.x, sc), __fmul_rn(pt0.y, sc));
        if (level == L.nlevels - 1) nextPt = prevPt;
        else nextPt = make_float2(__fmul_rn(nextPt.x, 2.f), __fmul_rn(nextPt.y, 2.f));
        prevPt.x = __fsub_rn(prevPt.x, half); prevPt.y = __fsub_rn(prevPt.y, half);
        const int ipx = (int)floorf(prevPt.x), ipy = (int)floorf(prevPt.y);
        if (ipx < -win || ipx >= w || ipy < -win || ipy >= h) {
            if (level == 0) st = false;
            continue;
        }
        float a = __fsub_rn(prevPt.x, (float)ipx), b = __fsub_rn(prevPt.y, (float)ipy);
        int iw00 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, 1.f - b), 16384.f));
        int iw01 = __float2int_rn(__fmul_rn(__fmul_rn(a, 1.f - b), 16384.f));
        int iw10 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, b), 16384.f));
        int iw11 = 16384 - iw00 - iw01 - iw10;
        float A11 = 0, A12 = 0, A22 = 0;
        const bool insideI = ipx >= 0 && ipy >= 0 && ipx + WIN + 1 <= w && ipy + WIN + 1 <= h;
#pragma unroll
        for (int k = 0; k < NS; k++) {
            const int i = t + T * k;
            Ib[k] = Ixb[k] = Iyb[k] = 0;
            if (i < WIN * WIN) {
                int ival, ixval, iyval;
                const int y = i / WIN, x = i - y * WIN;
                if (insideI) {
                    const uint8_t* q = I + (size_t)ipy * pitch + ipx + off[k];
                    ival = ((int)q[0] * iw00 + (int)q[1] * iw01 + (int)q[pitch] * iw10 + (int)q[pitch + 1] * iw11 + (1 << 8)) >> 9;
                    const short2* dq = D + (size_t)(ipy + y) * w + ipx + x;
                    const short2 d00 = dq[0], d01 = dq[1], d10 = dq[w], d11 = dq[w + 1];
                    ixval = ((int)d00.x * iw00 + (int)d01.x * iw01 + (int)d10.x * iw10 + (int)d11.x * iw11 + (1 << 13)) >> 14;
                    iyval = ((int)d00.y * iw00 + (int)d01.y * iw01 + (int)d10.y * iw10 + (int)d11.y * iw11 + (1 << 13)) >> 14;
                } else {
                    const int gx0 = ipx + x, gy0 = ipy + y;
                    const int x0 = refl101(gx0, w), x1 = refl101(gx0 + 1, w), y0 = refl101(gy0, h), y1 = refl101(gy0 + 1, h);
                    ival = ((int)I[(size_t)y0 * pitch + x0] * iw00 + (int)I[(size_t)y0 * pitch + x1] * iw01 + (int)I[(size_t)y1 * pitch + x0] * iw10 +
                            (int)I[(size_t)y1 * pitch + x1] * iw11 + (1 << 8)) >> 9;
                    auto dv = [&](int gx, int gy) { return ((unsigned)gx < (unsigned)w && (unsigned)gy < (unsigned)h) ? D[(size_t)gy * w + gx] : make_short2(0, 0); };
                    const short2 d00 = dv(gx0, gy0), d01 = dv(gx0 + 1, gy0), d10 = dv(gx0, gy0 + 1), d11 = dv(gx0 + 1, gy0 + 1);
                    ixval = ((int)d00.x * iw00 + (int)d01.x * iw01 + (int)d10.x * iw10 + (int)d11.x * iw11 + (1 << 13)) >> 14;
                    iyval = ((int)d00.y * iw00 + (int)d01.y * iw01 + (int)d10.y * iw10 + (int)d11.y * iw11 + (1 << 13)) >> 14;
                }
                ival = (int)(short)ival; ixval = (int)(short)ixval; iyval = (int)(short)iyval;   // the reference stores int16
                Ib[k] = ival; Ixb[k] = ixval; Iyb[k] = iyval;
                A11 += (float)(ixval * ixval); A12 += (float)(ixval * iyval); A22 += (float)(iyval * iyval);
            }
        }
        cta_sum3(A11, A12, A22, s_red, par);
        A11 *= FLT_SCALE; A12 *= FLT_SCALE; A22 *= FLT_SCALE;
        float Dt = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA = __fsub_rn(A11, A22);
        const float minEig = (A22 + A11 - sqrtf(__fadd_rn(__fmul_rn(dA, dA), __fmul_rn(4.f, __fmul_rn(A12, A12))))) / (float)(2 * win * win);
        if (minEig < min_eig_thr || Dt < FLT_EPSILON) {
            if (level == 0) st = false;
            continue;
        }
        Dt = 1.f / Dt;
        nextPt.x = __fsub_rn(nextPt.x, half); nextPt.y = __fsub_rn(nextPt.y, half);
        float2 prevDelta = make_float2(0.f, 0.f);
        float2 result = make_float2(__fadd_rn(nextPt.x, half), __fadd_rn(nextPt.y, half));
        for (int j = 0; j < max_iters; j++) {
            const int inx = (int)floorf(nextPt.x), iny = (int)floorf(nextPt.y);
            if (inx < -win || inx >= w || iny < -win || iny >= h) {
                if (level == 0) st = false;
                break;
            }
            a = __fsub_rn(nextPt.x, (float)inx); b = __fsub_rn(nextPt.y, (float)iny);
            iw00 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, 1.f - b), 16384.f));
            iw01 = __float2int_rn(__fmul_rn(__fmul_rn(a, 1.f - b), 16384.f));
            iw10 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, b), 16384.f));
            iw11 = 16384 - iw00 - iw01 - iw10;
            float b1 = 0, b2 = 0, unused = 0;
            const bool insideJ = inx >= 0 && iny >= 0 && inx + WIN + 1 <= w && iny + WIN + 1 <= h;
            if (insideJ) {
                const uint8_t* base = J + (size_t)iny * pitch + inx;
#pragma unroll
                for (int k = 0; k < NS; k++)
                    if (t + T * k < WIN * WIN) {
                        const uint8_t* q = base + off[k];
                        const int jv = ((int)q[0] * iw00 + (int)q[1] * iw01 + (int)q[pitch] * iw10 + (int)q[pitch + 1] * iw11 + (1 << 8)) >> 9;
                        const int diff = jv - Ib[k];
                        b1 += (float)(diff * Ixb[k]);
                        b2 += (float)(diff * Iyb[k]);
                    }
            } else {
#pragma unroll
                for (int k = 0; k < NS; k++) {
                    const int i = t + T * k;
                    if (i < WIN * WIN) {
                        const int y = i / WIN, x = i - y * WIN;
                        const int x0 = refl101(inx + x, w), x1 = refl101(inx + x + 1, w), y0 = refl101(iny + y, h), y1 = refl101(iny + y + 1, h);
                        const int jv = ((int)J[(size_t)y0 * pitch + x0] * iw00 + (int)J[(size_t)y0 * pitch + x1] * iw01 + (int)J[(size_t)y1 * pitch + x0] * iw10 +
                                        (int)J[(size_t)y1 * pitch + x1] * iw11 + (1 << 8)) >> 9;
                        const int diff = jv - Ib[k];
                        b1 += (float)(diff * Ixb[k]);
                        b2 += (float)(diff * Iyb[k]);
                    }
                }
            }
            cta_sum3(b1, b2, unused, s_red, par);
            b1 *= FLT_SCALE; b2 *= FLT_SCALE;
            const float2 delta = make_float2(__fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), Dt), __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), Dt));
            nextPt.x = __fadd_rn(nextPt.x, delta.x); nextPt.y = __fadd_rn(nextPt.y, delta.y);
            result = make_float2(__fadd_rn(nextPt.x, half), __fadd_rn(nextPt.y, half));
            if ((double)delta.x * delta.x + (double)delta.y * delta.y <= (double)eps2) break;
            if (j > 0 && fabsf(delta.x + prevDelta.x) < 0.01f && fabsf(delta.y + prevDelta.y) < 0.01f) {
                result.x = __fsub_rn(result.x, __fmul_rn(delta.x, 0.5f)); result.y = __fsub_rn(result.y, __fmul_rn(delta.y, 0.5f));
                break;
            }
            prevDelta = delta;
        }
        nextPt = result;
    }
    if (t == 0) {
        next_pts[p] = nextPt;
        if (next_host) next_host[p] = nextPt;
        // src/Frame.cc:336-364: 5-px border test on the truncated coordinates, then the 3x3 sum of absolute differences
        if (st && edge >= 0) {
            const int w = L.w[0], h = L.h[0], pitch = L.pitch[0];
            const int x1 = (int)pt0.x, y1 = (int)pt0.y, x2 = (int)nextPt.x, y2 = (int)nextPt.y;
            if (x1 < edge || x1 >= w - edge || x2 < edge || x2 >= w - edge || y1 < edge || y1 >= h - edge || y2 < edge || y2 >= h - edge) {
                st = false;
            } else {
                int sad = 0;
                for (int j = -1; j <= 1; j++)
                    for (int i = -1; i <= 1; i++) sad += abs((int)L.prev[0][(size_t)(y1 + j) * pitch + x1 + i] - (int)L.cur[0][(size_t)(y2 + j) * pitch + x2 + i]);
                if ((float)sad > sad_limit) st = false;
            }
        }
        status[p] = st ? 1 : 0;
        if (status_host) status_host[p] = st ? 1 : 0;
    }
}

// src/Frame.cc:372-385: the epipolar distance of every tracked point, in double, with the reference's expression.
__global__ void epipolar_kernel(const float2* __restrict__ pre, const float2* __restrict__ nxt, const uint8_t* __restrict__ status, int n, const double* __restrict__ F,
                                double limit, uint8_t* __restrict__ moving, double* __restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t mv = 0;
    double dd = -1.0;
    if (status[i]) {
        const double px = pre[i].x, py = pre[i].y;
        const double A = __dadd_rn(__dadd_rn(__dmul_rn(F[0], px), __dmul_rn(F[1], py)), F[2]);
        const double B = __dadd_rn(__dadd_rn(__dmul_rn(F[3], px), __dmul_rn(F[4], py)), F[5]);
        const double C = __dadd_rn(__dadd_rn(__dmul_rn(F[6], px), __dmul_rn(F[7], py)), F[8]);
        dd = fabs(__dadd_rn(__dadd_rn(__dmul_rn(A, (double)nxt[i].x), __dmul_rn(B, (double)nxt[i].y)), C)) / sqrt(__dadd_rn(__dmul_rn(A, A), __dmul_rn(B, B)));
        mv = !(dd <= limit);
    }
    moving[i] = mv;
    if (dist) dist[i] = dd;
}

}  // namespace coeb

using namespace coeb;

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
// Mapped pinned block of coeb_process_moving_object (h_out): refined corners | tracked positions | states | selected corners | count
constexpr size_t kOutPre = 0, kOutNext = kOutPre + sizeof(float2) * kMoMaxPts, kOutState = kOutNext + sizeof(float2) * kMoMaxPts,
                 kOutSelected = kOutState + kMoMaxPts, kOutCount = kOutSelected + sizeof(float2) * kMoMaxPts, kOutBytes = kOutCount + 16;
struct coeb_motion {
    int device = 0;
    cudaStream_t stream = nullptr;
    int w = 0, h = 0, pitch = 0, nlevels = 0;
    int lw[kMoMaxLevels] = {0}, lh[kMoMaxLevels] = {0}, lp[kMoMaxLevels] = {0};
    uint8_t* d_pyr[2][kMoMaxLevels] = {{nullptr}};   // [0] previous frame, [1] current frame
    short2* d_deriv[kMoMaxLevels] = {nullptr};
    float* d_resp = nullptr;
    unsigned* d_max = nullptr;      // [0] max response bits, [1] candidate count
    bool info_dirty = true;         // d_max has to be cleared before the next Harris pass (the sort kernel normally does it)
    bool force_host_select = false; // coeb_process_moving_object's second attempt after the device declined the minimum-distance pass
    float2* d_cand = nullptr;
    float2 *d_pre = nullptr, *d_next = nullptr;
    uint8_t *d_status = nullptr, *d_moving = nullptr;
    double *d_F = nullptr, *d_dist = nullptr;
    float* d_mask = nullptr; int mask_hw = 0;
    char* h_pin = nullptr; size_t pin_bytes = 0;
    char* h_out = nullptr;          // mapped pinned block the point kernels mirror their outputs into (coeb_process_moving_object)
    std::vector<float2> cand_host;
    // coeb_process_moving_object: the current frame's upload, both pyramids and the derivative images run on a side stream beside the
    // corner stage (which ends in a host round trip); `side` describes the work good_features() enqueues there before it synchronises
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_prev = nullptr, ev_side = nullptr;
    uint8_t* h_frame[2] = {nullptr, nullptr}; size_t frame_bytes = 0;   // pinned staging of the two (pageable) frames, dense rows
    // the side stream's work as captured graphs (fixed addresses, one launch): [0..1] for a call that uploads both frames, [2..3] for a
    // sequence call (the previous frame is resident, its pyramid built); two of each because the two pyramid buffers swap roles from call
    // to call, keyed by the level-0 address of the "previous" buffer at capture time
    cudaGraphExec_t side_graph[4] = {nullptr, nullptr, nullptr, nullptr};
    const void* side_graph_key[4] = {nullptr, nullptr, nullptr, nullptr};
    // coeb_process_moving_object_next: the current frame of the last call stays in d_pyr[1] with its pyramid
    bool cur_resident = false; int res_w = 0, res_h = 0;
    struct { const uint8_t* cur_gray = nullptr; int stride = 0; bool done = false; bool select_on_device = false; bool prev_resident = false; } side;
};

namespace {

void drop_side_graph(coeb_motion* m);
void motion_free_images(coeb_motion* m) {
    drop_side_graph(m);
    for (int f = 0; f < 2; f++)
        for (int l = 0; l < kMoMaxLevels; l++) { cudaFree(m->d_pyr[f][l]); m->d_pyr[f][l] = nullptr; }
    for (int l = 0; l < kMoMaxLevels; l++) { cudaFree(m->d_deriv[l]); m->d_deriv[l] = nullptr; }
    cudaFree(m->d_resp); m->d_resp = nullptr;
    m->w = m->h = 0;
}

// (Re)allocates the per-size buffers. Pyramid levels as cv::buildOpticalFlowPyramid lays them out: ((w+1)/2, (h+1)/2) per level,
// stopping before a level that is not larger than the tracking window.
int motion_prepare(coeb_motion* m, int w, int h, int win, int max_level) {
    int nl = 1, cw = w, ch = h;
    int lw[kMoMaxLevels], lh[kMoMaxLevels];
    lw[0] = w; lh[0] = h;
    while (nl <= max_level && nl < kMoMaxLevels) {
        const int nw = (cw + 1) / 2, nh = (ch + 1) / 2;
        if (nw <= win || nh <= win) break;
        lw[nl] = nw; lh[nl] = nh; cw = nw; ch = nh; nl++;
    }
    bool same = m->w == w && m->h == h && m->nlevels == nl;
    if (same) return COEB_OK;
    motion_free_images(m);
    m->w = w; m->h = h; m->nlevels = nl; m->pitch = (w + 63) & ~63;
    for (int l = 0; l < nl; l++) {
        m->lw[l] = lw[l]; m->lh[l] = lh[l]; m->lp[l] = (lw[l] + 63) & ~63;
        for (int f = 0; f < 2; f++) CUDA_TRY(cudaMalloc(&m->d_pyr[f][l], (size_t)m->lp[l] * lh[l]));
        CUDA_TRY(cudaMalloc(&m->d_deriv[l], sizeof(short2) * (size_t)lw[l] * lh[l]));
    }
    CUDA_TRY(cudaMalloc(&m->d_resp, sizeof(float) * (size_t)w * h));
    return COEB_OK;
}

int motion_pin(coeb_motion* m, size_t bytes) {
    if (bytes <= m->pin_bytes) return COEB_OK;
    if (m->h_pin) cudaFreeHost(m->h_pin);
    m->h_pin = nullptr; m->pin_bytes = 0;
    CUDA_TRY(cudaHostAlloc((void**)&m->h_pin, bytes, cudaHostAllocMapped));   // (the sort kernel writes into it)
    m->pin_bytes = bytes;
    return COEB_OK;
}

void drop_side_graph(coeb_motion* m) {
    for (int k = 0; k < 4; k++) {
        if (m->side_graph[k]) cudaGraphExecDestroy(m->side_graph[k]);
        m->side_graph[k] = nullptr; m->side_graph_key[k] = nullptr;
    }
    m->cur_resident = false;   // (called whenever the image buffers are laid out again)
}

// The caller's frame is pageable as a rule (a cv::Mat): copied into pinned memory here (~10 us for 640x480) it travels by DMA while
// the host goes on; handed to cudaMemcpy2DAsync directly the driver stages it itself and the call blocks for ~35 us.
int stage_frame(coeb_motion* m, int which, const uint8_t* gray, int stride) {
    const size_t need = (size_t)m->w * m->h;
    if (need > m->frame_bytes) {
        drop_side_graph(m);
        for (int f = 0; f < 2; f++) { if (m->h_frame[f]) cudaFreeHost(m->h_frame[f]); m->h_frame[f] = nullptr; }
        m->frame_bytes = 0;
        for (int f = 0; f < 2; f++) CUDA_TRY(cudaHostAlloc((void**)&m->h_frame[f], need, cudaHostAllocDefault));
        m->frame_bytes = need;
    }
    if (stride == m->w) std::memcpy(m->h_frame[which], gray, need);
    else for (int y = 0; y < m->h; y++) std::memcpy(m->h_frame[which] + (size_t)y * m->w, gray + (size_t)y * stride, m->w);
    return COEB_OK;
}

int upload_level0(coeb_motion* m, int which, const uint8_t* gray, int stride, cudaStream_t stream = nullptr) {
    int st = stage_frame(m, which, gray, stride);
    if (st != COEB_OK) return st;
    CUDA_TRY(cudaMemcpy2DAsync(m->d_pyr[which][0], m->lp[0], m->h_frame[which], m->w, m->w, m->h, cudaMemcpyHostToDevice, stream ? stream : m->stream));
    return COEB_OK;
}

void build_pyramid(coeb_motion* m, int which, cudaStream_t stream = nullptr) {
    if (!stream) stream = m->stream;
    for (int l = 1; l < m->nlevels; l++)
        pyr_down_kernel<<<dim3((m->lw[l] + 31) / 32, (m->lh[l] + 7) / 8), 256, 0, stream>>>(m->d_pyr[which][l - 1], m->lw[l - 1], m->lh[l - 1], m->lp[l - 1],
                                                                                        m->d_pyr[which][l], m->lw[l], m->lh[l], m->lp[l]);
}

void derivative_images(coeb_motion* m, cudaStream_t stream) {
    for (int l = 0; l < m->nlevels; l++)
        scharr_kernel<<<dim3((m->lw[l] + 31) / 32, (m->lh[l] + 7) / 8), 256, 0, stream>>>(m->d_pyr[0][l], m->lw[l], m->lh[l], m->lp[l], m->d_deriv[l]);
}

// Called by good_features() between its last launch and its synchronisation, when coeb_process_moving_object asked for it: everything
// calcOpticalFlowPyrLK needs besides the points (current frame, both pyramids, Scharr images of the previous one) is enqueued on the
// side stream, so the staging of the pageable current frame overlaps the Harris kernels and the kernels overlap the host's
// minimum-distance pass. The LK launch waits for ev_side.
int enqueue_side_work(coeb_motion* m) {
    if (!m->side.cur_gray) return COEB_OK;
    int st = stage_frame(m, 1, m->side.cur_gray, m->side.stride);
    if (st != COEB_OK) return st;
    CUDA_TRY(cudaStreamWaitEvent(m->stream2, m->ev_prev, 0));   // level 0 of the previous frame has landed
    const int base = m->side.prev_resident ? 2 : 0;
    int slot = -1;
    for (int k = base; k < base + 2; k++) if (m->side_graph[k] && m->side_graph_key[k] == m->d_pyr[0][0]) slot = k;
    if (slot < 0) {   // every address is fixed until the buffers are laid out again: captured once per role assignment, replayed as one launch
        slot = m->side_graph[base] ? base + 1 : base;
        if (m->side_graph[slot]) { cudaGraphExecDestroy(m->side_graph[slot]); m->side_graph[slot] = nullptr; }
        cudaGraph_t g = nullptr;
        CUDA_TRY(cudaStreamBeginCapture(m->stream2, cudaStreamCaptureModeThreadLocal));
        cudaMemcpy2DAsync(m->d_pyr[1][0], m->lp[0], m->h_frame[1], m->w, m->w, m->h, cudaMemcpyHostToDevice, m->stream2);
        if (!m->side.prev_resident) build_pyramid(m, 0, m->stream2);   // a resident previous frame was the current one of the last call: its pyramid exists
        derivative_images(m, m->stream2);
        build_pyramid(m, 1, m->stream2);
        cudaError_t e = cudaStreamEndCapture(m->stream2, &g);
        if (e == cudaSuccess) e = cudaGraphInstantiate(&m->side_graph[slot], g, 0);
        if (g) cudaGraphDestroy(g);
        if (e != cudaSuccess) { m->side_graph[slot] = nullptr; return fail(COEB_ERR_CUDA, "side-stream graph: %s", cudaGetErrorString(e)); }
        m->side_graph_key[slot] = m->d_pyr[0][0];
    }
    CUDA_TRY(cudaGraphLaunch(m->side_graph[slot], m->stream2));
    CUDA_TRY(cudaEventRecord(m->ev_side, m->stream2));
    m->side.done = true;
    return COEB_OK;
}

int ensure_mask(coeb_motion* m, int hw) {
    if (m->d_mask && m->mask_hw == hw) return COEB_OK;
    cudaFree(m->d_mask); m->d_mask = nullptr;
    const int win = 2 * hw + 1;
    std::vector<float> mask((size_t)win * win);
    for (int i = 0; i < win; i++) {   // cornersubpix.cpp: float arguments, float exp
        const float y = (float)(i - hw) / hw;
        const float vy = std::exp(-y * y);
        for (int j = 0; j < win; j++) {
            const float x = (float)(j - hw) / hw;
            mask[(size_t)i * win + j] = (float)(vy * std::exp(-x * x));
        }
    }
    CUDA_TRY(cudaMalloc(&m->d_mask, mask.size() * sizeof(float)));
    CUDA_TRY(cudaMemcpy(m->d_mask, mask.data(), mask.size() * sizeof(float), cudaMemcpyHostToDevice));
    m->mask_hw = hw;
    return COEB_OK;
}

// goodFeaturesToTrack after the response: candidates sorted by (value descending, raster position descending -- the pointer
// tie-break of greaterThanPtr), then the minimum-distance pass over a grid of cell size round(minDistance).
int select_corners(std::vector<float2>& cand, bool sorted, int w, int h, int max_corners, double min_distance, float* xy_out, int cap) {
    // positive floats order like their bit patterns: one 64-bit key (value bits, raster position) per candidate, sorted descending
    static_assert(sizeof(float2) == sizeof(unsigned long long), "a candidate is one 64-bit key");
    if (!sorted) {
        unsigned long long* key = reinterpret_cast<unsigned long long*>(cand.data());
        for (size_t i = 0; i < cand.size(); i++) {
            unsigned vb, ib;
            std::memcpy(&vb, &cand[i].x, 4); std::memcpy(&ib, &cand[i].y, 4);
            key[i] = ((unsigned long long)vb << 32) | ib;
        }
        std::sort(key, key + cand.size(), [](unsigned long long a, unsigned long long b) { return a > b; });
        for (size_t i = 0; i < cand.size(); i++) {
            const unsigned vb = (unsigned)(key[i] >> 32), ib = (unsigned)key[i];
            float2 c;
            std::memcpy(&c.x, &vb, 4); std::memcpy(&c.y, &ib, 4);
            cand[i] = c;
        }
    }
    int n = 0;
    if (min_distance >= 1) {
        const int cell = (int)std::lrint(min_distance);
        const int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
        const double md2 = min_distance * min_distance;
        // Flat grid: accepted corners are >= minDistance apart, so a cell of that size holds a handful; kSlots each, counts in one byte
        // array (a vector per cell cost this pass 4800 constructions and a heap allocation per occupied cell: 70 us of the call).
        // Should a cell ever fill up, the pass starts over with the general container below.
        constexpr int kSlots = 6;
        static thread_local std::vector<unsigned char> cnt;
        static thread_local std::vector<short> slot;
        cnt.assign((size_t)gw * gh, 0);
        if (slot.size() < (size_t)gw * gh * kSlots * 2) slot.resize((size_t)gw * gh * kSlots * 2);
        bool overflow = false;
        const unsigned rcp_w = w < 65536 ? (unsigned)(((1ull << 32) + w - 1) / w) : 0;   // idx / w by multiply-shift (exact for idx < 2^32 / w ... checked below)
        for (size_t i = 0; i < cand.size() && !overflow; i++) {
            int idx;
            std::memcpy(&idx, &cand[i].y, 4);
            int y = rcp_w ? (int)(((unsigned long long)(unsigned)idx * rcp_w) >> 32) : idx / w;
            int x = idx - y * w;
            if (x < 0 || x >= w) { y = idx / w; x = idx - y * w; }   // the reciprocal may be one off at the far end: redo exactly
            const int xc = x / cell, yc = y / cell;
            const int x1 = std::max(0, xc - 1), y1 = std::max(0, yc - 1), x2 = std::min(gw - 1, xc + 1), y2 = std::min(gh - 1, yc + 1);
            bool good = true;
            for (int yy = y1; yy <= y2 && good; yy++)
                for (int xx = x1; xx <= x2 && good; xx++) {
                    const size_t c = (size_t)yy * gw + xx;
                    const short* q = &slot[c * kSlots * 2];
                    for (int k = 0; k < cnt[c]; k++) {
                        const float dx = (float)(x - q[2 * k]), dy = (float)(y - q[2 * k + 1]);
                        if (dx * dx + dy * dy < md2) { good = false; break; }
                    }
                }
            if (good) {
                const size_t c = (size_t)yc * gw + xc;
                if (cnt[c] == kSlots) { overflow = true; break; }
                slot[(c * kSlots + cnt[c]) * 2] = (short)x; slot[(c * kSlots + cnt[c]) * 2 + 1] = (short)y;
                cnt[c]++;
                if (n < cap) { xy_out[2 * n] = (float)x; xy_out[2 * n + 1] = (float)y; }
                n++;
                if (max_corners > 0 && n == max_corners) break;
            }
        }
        if (!overflow) return n;
        n = 0;
        std::vector<std::vector<float2> > grid((size_t)gw * gh);
        for (size_t i = 0; i < cand.size(); i++) {
            int idx;
            std::memcpy(&idx, &cand[i].y, 4);
            const int y = idx / w, x = idx - y * w;
            bool good = true;
            const int xc = x / cell, yc = y / cell;
            const int x1 = std::max(0, xc - 1), y1 = std::max(0, yc - 1), x2 = std::min(gw - 1, xc + 1), y2 = std::min(gh - 1, yc + 1);
            for (int yy = y1; yy <= y2 && good; yy++)
                for (int xx = x1; xx <= x2 && good; xx++)
                    for (const float2& q : grid[(size_t)yy * gw + xx]) {
                        const float dx = (float)x - q.x, dy = (float)y - q.y;
                        if (dx * dx + dy * dy < md2) { good = false; break; }
                    }
            if (good) {
                grid[(size_t)yc * gw + xc].push_back(make_float2((float)x, (float)y));
                if (n < cap) { xy_out[2 * n] = (float)x; xy_out[2 * n + 1] = (float)y; }
                n++;
                if (max_corners > 0 && n == max_corners) break;
            }
        }
    } else {
        for (size_t i = 0; i < cand.size(); i++) {
            int idx;
            std::memcpy(&idx, &cand[i].y, 4);
            if (n < cap) { xy_out[2 * n] = (float)(idx % w); xy_out[2 * n + 1] = (float)(idx / w); }
            n++;
            if (max_corners > 0 && n == max_corners) break;
        }
    }
    return n;
}

// ---- fundamental matrix: normalised 8-point inside RANSAC (host, double) -------------------------------------------------
// Smallest-eigenvalue eigenvector of a symmetric n x n matrix by cyclic Jacobi rotations.
void jacobi_eigen(double* A, int n, double* V, double* ev) {
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * n + j] = i == j;
    double diag2 = 0;
    for (int i = 0; i < n; i++) diag2 += A[i * n + i] * A[i * n + i];
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
        if (off < 1e-30 || off < 1e-28 * diag2) break;   // converged to ~1e-14 of the matrix scale (the convergence is quadratic: one more sweep squares it)
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                if (std::fabs(A[p * n + q]) < 1e-300) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2 * A[p * n + q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
                const double c = 1 / std::sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < n; k++) { const double akp = A[k * n + p], akq = A[k * n + q]; A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq; }
                for (int k = 0; k < n; k++) { const double apk = A[p * n + k], aqk = A[q * n + k]; A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk; }
                for (int k = 0; k < n; k++) { const double vkp = V[k * n + p], vkq = V[k * n + q]; V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq; }
            }
    }
    for (int i = 0; i < n; i++) ev[i] = A[i * n + i];
}

// Unit eigenvector of the smallest eigenvalue of a symmetric positive semi-definite 3 x 3 matrix in closed form (trigonometric
// eigenvalues, then the largest cross product of two rows of G - lambda I): the right singular vector the rank-2 projection of the
// minimal-sample hypotheses removes. The Jacobi sweeps this replaces were a third of a hypothesis (0.47 of 1.35 us, ~40 per call).
// false if the matrix is (numerically) a multiple of the identity.
bool smallest_eigvec3(const double* G, double v[3]) {
    const double p1 = G[1] * G[1] + G[2] * G[2] + G[5] * G[5];
    const double q = (G[0] + G[4] + G[8]) / 3.0;
    const double d0 = G[0] - q, d1 = G[4] - q, d2 = G[8] - q;
    const double p2 = d0 * d0 + d1 * d1 + d2 * d2 + 2.0 * p1;
    if (!(p2 > 0.0)) return false;
    const double p = std::sqrt(p2 / 6.0), ip = 1.0 / p;
    const double b00 = d0 * ip, b11 = d1 * ip, b22 = d2 * ip, b01 = G[1] * ip, b02 = G[2] * ip, b12 = G[5] * ip;
    double r = 0.5 * (b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02));
    r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
    const double phi = std::acos(r) / 3.0;
    const double lam = q + 2.0 * p * std::cos(phi + 2.0943951023931953);   // + 2 pi / 3: the smallest of the three
    const double r0[3] = {G[0] - lam, G[1], G[2]}, r1[3] = {G[3], G[4] - lam, G[5]}, r2[3] = {G[6], G[7], G[8] - lam};
    const double c[3][3] = {{r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]},
                            {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]},
                            {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]}};
    int best = 0;
    double nb = -1.0;
    for (int i = 0; i < 3; i++) {
        const double n2 = c[i][0] * c[i][0] + c[i][1] * c[i][1] + c[i][2] * c[i][2];
        if (n2 > nb) { nb = n2; best = i; }
    }
    if (!(nb > 0.0)) return false;
    const double inv = 1.0 / std::sqrt(nb);
    v[0] = c[best][0] * inv; v[1] = c[best][1] * inv; v[2] = c[best][2] * inv;
    return true;
}

// Unit eigenvector of the smallest eigenvalue of a symmetric positive semi-definite n x n matrix (n <= 9) by inverse iteration on
// A + shift I (Cholesky factor, two triangular solves per step): the least-squares null vector of the final 8-point fit, whose smallest
// eigenvalue (the noise of the consensus set) lies orders of magnitude below the next, so a handful of steps reach machine precision
// where the Jacobi sweeps took two thirds of the 27-us fit. false (the caller falls back to the sweeps) if the factorisation breaks
// down or the iteration has not settled after 40 steps (a degenerate configuration: two comparable smallest eigenvalues).
bool smallest_eigvec_inverse_iteration(const double* A, int n, double* v) {
    double L[81], tr = 0;
    for (int i = 0; i < n; i++) tr += A[i * n + i];
    if (!(tr > 0)) return false;
    const double shift = 1e-12 * tr;
    for (int j = 0; j < n; j++) {
        double d = A[j * n + j] + shift;
        for (int k = 0; k < j; k++) d -= L[j * n + k] * L[j * n + k];
        if (!(d > 0)) return false;
        const double ljj = std::sqrt(d), inv = 1.0 / ljj;
        L[j * n + j] = ljj;
        for (int i = j + 1; i < n; i++) {
            double t = A[i * n + j];
            for (int k = 0; k < j; k++) t -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = t * inv;
        }
    }
    double x[9], y[9];
    for (int i = 0; i < n; i++) x[i] = 1.0 / std::sqrt((double)n) * ((i & 1) ? -1.0 : 1.0) * (1.0 + 0.1 * i);   // no special direction
    double nrm = 0;
    for (int i = 0; i < n; i++) nrm += x[i] * x[i];
    nrm = 1.0 / std::sqrt(nrm);
    for (int i = 0; i < n; i++) x[i] *= nrm;
    for (int it = 0; it < 40; it++) {
        for (int i = 0; i < n; i++) {   // L y = x
            double t = x[i];
            for (int k = 0; k < i; k++) t -= L[i * n + k] * y[k];
            y[i] = t / L[i * n + i];
        }
        for (int i = n - 1; i >= 0; i--) {   // L^T z = y (in place)
            double t = y[i];
            for (int k = i + 1; k < n; k++) t -= L[k * n + i] * y[k];
            y[i] = t / L[i * n + i];
        }
        double n2 = 0, dot = 0;
        for (int i = 0; i < n; i++) n2 += y[i] * y[i];
        if (!(n2 > 0) || !std::isfinite(n2)) return false;
        const double inv = 1.0 / std::sqrt(n2);
        for (int i = 0; i < n; i++) { y[i] *= inv; dot += y[i] * x[i]; }
        const bool settled = it > 0 && 1.0 - std::fabs(dot) < 1e-15;
        for (int i = 0; i < n; i++) x[i] = y[i];
        if (settled) { for (int i = 0; i < n; i++) v[i] = x[i]; return true; }
    }
    return false;
}

// The 9 x 9 normal matrix of the final fit, sum over the consensus set of r r^T with r the epipolar row of a normalised correspondence, as whole
// rows: every entry is the same sum of the same products in the same (point) order as the upper-triangle loop in eight_point (a product does not
// depend on the order of its factors; multiplies and adds are separate instructions here: explicit intrinsics, no FMA in the target), so the matrix
// is the same to the last bit, but a row of twelve (nine and three zeros) is three 4-wide multiplies and adds where the triangle was 45 scalar
// pairs. false: this CPU has no AVX2, the caller runs the scalar loop.
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2"))) bool normal_matrix_rows_avx2(const float* p1, const float* p2, const int* idx, int cnt, double m1x, double m1y, double m2x, double m2y, double s1, double s2,
                                                             double A[81]) {
    typedef double v4 __attribute__((vector_size(32)));
    v4 acc[9][3];
    for (int a = 0; a < 9; a++) for (int q = 0; q < 3; q++) acc[a][q] = v4{0, 0, 0, 0};
    for (int k = 0; k < cnt; k++) {
        const int i = idx[k];
        const double x1 = (p1[2 * i] - m1x) * s1, y1 = (p1[2 * i + 1] - m1y) * s1, x2 = (p2[2 * i] - m2x) * s2, y2 = (p2[2 * i + 1] - m2y) * s2;
        const double r[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1};
        const v4 r0 = {r[0], r[1], r[2], r[3]}, r1 = {r[4], r[5], r[6], r[7]}, r2 = {r[8], 0, 0, 0};
#pragma GCC unroll 9
        for (int a = 0; a < 9; a++) {
            const v4 ra = {r[a], r[a], r[a], r[a]};
            const v4 t0 = ra * r0, t1 = ra * r1, t2 = ra * r2;   // (no FMA in this target: the sums below round like the scalar loop's)
            acc[a][0] += t0; acc[a][1] += t1; acc[a][2] += t2;
        }
    }
    for (int a = 0; a < 9; a++) for (int b = 0; b < 9; b++) A[a * 9 + b] = acc[a][b >> 2][b & 3];
    return true;
}
bool normal_matrix_rows(const float* p1, const float* p2, const int* idx, int cnt, double m1x, double m1y, double m2x, double m2y, double s1, double s2, double A[81]) {
    static const bool avx2 = __builtin_cpu_supports("avx2") && getenv("COEB_MOTION_SCALAR_FIT") == nullptr;   // (development switch: the scalar triangle loop)
    return avx2 && normal_matrix_rows_avx2(p1, p2, idx, cnt, m1x, m1y, m2x, m2y, s1, s2, A);
}
#else
bool normal_matrix_rows(const float*, const float*, const int*, int, double, double, double, double, double, double, double*) { return false; }
#endif

// 8-point algorithm on the points idx[0..cnt): Hartley normalisation, least-squares null vector, rank-2 projection. false if degenerate.
bool eight_point(const float* p1, const float* p2, const int* idx, int cnt, double F[9]) {
    double m1x = 0, m1y = 0, m2x = 0, m2y = 0;
    for (int k = 0; k < cnt; k++) { const int i = idx[k]; m1x += p1[2 * i]; m1y += p1[2 * i + 1]; m2x += p2[2 * i]; m2y += p2[2 * i + 1]; }
    m1x /= cnt; m1y /= cnt; m2x /= cnt; m2y /= cnt;
    double s1 = 0, s2 = 0;
    for (int k = 0; k < cnt; k++) {
        const int i = idx[k];
        s1 += std::sqrt((p1[2 * i] - m1x) * (p1[2 * i] - m1x) + (p1[2 * i + 1] - m1y) * (p1[2 * i + 1] - m1y));
        s2 += std::sqrt((p2[2 * i] - m2x) * (p2[2 * i] - m2x) + (p2[2 * i + 1] - m2y) * (p2[2 * i + 1] - m2y));
    }
    if (s1 < 1e-12 || s2 < 1e-12) return false;
    s1 = std::sqrt(2.0) * cnt / s1; s2 = std::sqrt(2.0) * cnt / s2;
    double A[81] = {0};
    if (!normal_matrix_rows(p1, p2, idx, cnt, m1x, m1y, m2x, m2y, s1, s2, A)) {
        for (int k = 0; k < cnt; k++) {
            const int i = idx[k];
            const double x1 = (p1[2 * i] - m1x) * s1, y1 = (p1[2 * i + 1] - m1y) * s1, x2 = (p2[2 * i] - m2x) * s2, y2 = (p2[2 * i + 1] - m2y) * s2;
            const double r[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1};
            for (int a = 0; a < 9; a++) for (int b = a; b < 9; b++) A[a * 9 + b] += r[a] * r[b];   // upper triangle, mirrored below
        }
        for (int a = 0; a < 9; a++) for (int b = 0; b < a; b++) A[a * 9 + b] = A[b * 9 + a];
    }
    double F0[9];
    if (!smallest_eigvec_inverse_iteration(A, 9, F0)) {
        double V[81], ev[9];
        jacobi_eigen(A, 9, V, ev);
        int mi = 0;
        for (int i = 1; i < 9; i++) if (ev[i] < ev[mi]) mi = i;
        for (int i = 0; i < 9; i++) F0[i] = V[i * 9 + mi];
    }
    // rank 2: F0 = U S V^T, zero the smallest singular value. Through the eigen decomposition of F0^T F0.
    double G[9] = {0}, Vg[9], eg[3];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) for (int k = 0; k < 3; k++) G[a * 3 + b] += F0[k * 3 + a] * F0[k * 3 + b];
    jacobi_eigen(G, 3, Vg, eg);
    int sm = 0;
    for (int i = 1; i < 3; i++) if (eg[i] < eg[sm]) sm = i;
    // F2 = F0 (I - v v^T) with v the right singular vector of the smallest singular value
    double v[3] = {Vg[0 * 3 + sm], Vg[1 * 3 + sm], Vg[2 * 3 + sm]}, F2[9];
    for (int a = 0; a < 3; a++) {
        const double d = F0[a * 3] * v[0] + F0[a * 3 + 1] * v[1] + F0[a * 3 + 2] * v[2];
        for (int b = 0; b < 3; b++) F2[a * 3 + b] = F0[a * 3 + b] - d * v[b];
    }
    // denormalise: F = T2^T F2 T1, T = [s 0 -s m; 0 s -s m; 0 0 1]
    const double T1[9] = {s1, 0, -s1 * m1x, 0, s1, -s1 * m1y, 0, 0, 1}, T2[9] = {s2, 0, -s2 * m2x, 0, s2, -s2 * m2y, 0, 0, 1};
    double tmp[9] = {0};
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) for (int k = 0; k < 3; k++) tmp[a * 3 + b] += T2[k * 3 + a] * F2[k * 3 + b];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { F[a * 3 + b] = 0; for (int k = 0; k < 3; k++) F[a * 3 + b] += tmp[a * 3 + k] * T1[k * 3 + b]; }
    if (std::fabs(F[8]) > FLT_EPSILON) { const double inv = 1.0 / F[8]; for (int i = 0; i < 9; i++) F[i] *= inv; }   // cv scales so that F(2,2) = 1
    return true;
}

// The null vector of the 8 x 9 epipolar system of exactly eight correspondences by Gaussian elimination with partial pivoting
// (the minimal-sample hypotheses of RANSAC; the least-squares fit above is kept for the final estimate on the consensus set),
// followed by the same rank-2 projection. Points are Hartley-normalised first. false if the sample is degenerate.
bool eight_point_minimal(const float* p1, const float* p2, const int* idx, double F[9]) {
    double m1x = 0, m1y = 0, m2x = 0, m2y = 0;
    for (int k = 0; k < 8; k++) { const int i = idx[k]; m1x += p1[2 * i]; m1y += p1[2 * i + 1]; m2x += p2[2 * i]; m2y += p2[2 * i + 1]; }
    m1x /= 8; m1y /= 8; m2x /= 8; m2y /= 8;
    double s1 = 0, s2 = 0;
    for (int k = 0; k < 8; k++) {
        const int i = idx[k];
        s1 += std::sqrt((p1[2 * i] - m1x) * (p1[2 * i] - m1x) + (p1[2 * i + 1] - m1y) * (p1[2 * i + 1] - m1y));
        s2 += std::sqrt((p2[2 * i] - m2x) * (p2[2 * i] - m2x) + (p2[2 * i + 1] - m2y) * (p2[2 * i + 1] - m2y));
    }
    if (s1 < 1e-12 || s2 < 1e-12) return false;
    s1 = std::sqrt(2.0) * 8 / s1; s2 = std::sqrt(2.0) * 8 / s2;
    double A[8][9];
    for (int k = 0; k < 8; k++) {
        const int i = idx[k];
        const double x1 = (p1[2 * i] - m1x) * s1, y1 = (p1[2 * i + 1] - m1y) * s1, x2 = (p2[2 * i] - m2x) * s2, y2 = (p2[2 * i + 1] - m2y) * s2;
        const double r[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1};
        for (int a = 0; a < 9; a++) A[k][a] = r[a];
    }
    double F0[9];
    // Fast path: elimination with row pivoting on the first eight columns and the last one free (F0[8] = 1), a third of the operations of
    // the fully pivoted form below, which takes over when a pivot says that this choice of free column is a bad one (F0[8] ~ 0).
    bool solved = true;
    {
        double B[8][9];
        std::memcpy(B, A, sizeof(B));
        for (int c = 0; c < 8 && solved; c++) {
            int pr = c;
            double best = std::fabs(B[c][c]);
            for (int rr = c + 1; rr < 8; rr++) if (std::fabs(B[rr][c]) > best) { best = std::fabs(B[rr][c]); pr = rr; }
            if (best < 1e-7) { solved = false; break; }
            if (pr != c) for (int cc = c; cc < 9; cc++) std::swap(B[c][cc], B[pr][cc]);
            const double inv = 1.0 / B[c][c];
            for (int rr = c + 1; rr < 8; rr++) {
                const double f = B[rr][c] * inv;
                for (int cc = c + 1; cc < 9; cc++) B[rr][cc] -= f * B[c][cc];
            }
        }
        if (solved) {
            F0[8] = 1.0;
            for (int c = 7; c >= 0; c--) {
                double t = B[c][8];
                for (int cc = c + 1; cc < 8; cc++) t += B[c][cc] * F0[cc];
                F0[c] = -t / B[c][c];
            }
        }
    }
    if (!solved) {
    // reduced row echelon form with column pivoting bookkeeping: the one free column gives the null vector
    int piv_col[8];
    bool used[9] = {false};
    for (int r = 0; r < 8; r++) {
        int br = r, bc = -1;
        double best = 0;
        for (int rr = r; rr < 8; rr++)
            for (int c = 0; c < 9; c++)
                if (!used[c] && std::fabs(A[rr][c]) > best) { best = std::fabs(A[rr][c]); br = rr; bc = c; }
        if (bc < 0 || best < 1e-10) return false;
        if (br != r) for (int c = 0; c < 9; c++) std::swap(A[r][c], A[br][c]);
        used[bc] = true; piv_col[r] = bc;
        const double inv = 1.0 / A[r][bc];
        for (int c = 0; c < 9; c++) A[r][c] *= inv;
        for (int rr = 0; rr < 8; rr++)
            if (rr != r && A[rr][bc] != 0.0) {
                const double f = A[rr][bc];
                for (int c = 0; c < 9; c++) A[rr][c] -= f * A[r][c];
            }
    }
    int free_col = 0;
    while (used[free_col]) free_col++;
    F0[free_col] = 1.0;
    for (int r = 0; r < 8; r++) F0[piv_col[r]] = -A[r][free_col];
    }
    double G[9] = {0}, v[3], F2[9];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) for (int k = 0; k < 3; k++) G[a * 3 + b] += F0[k * 3 + a] * F0[k * 3 + b];
    if (!smallest_eigvec3(G, v)) {   // (a multiple of the identity: any direction serves; keep the general routine for it)
        double Vg[9], eg[3];
        jacobi_eigen(G, 3, Vg, eg);
        int sm = 0;
        for (int i = 1; i < 3; i++) if (eg[i] < eg[sm]) sm = i;
        v[0] = Vg[0 * 3 + sm]; v[1] = Vg[1 * 3 + sm]; v[2] = Vg[2 * 3 + sm];
    }
    for (int a = 0; a < 3; a++) {
        const double d = F0[a * 3] * v[0] + F0[a * 3 + 1] * v[1] + F0[a * 3 + 2] * v[2];
        for (int b = 0; b < 3; b++) F2[a * 3 + b] = F0[a * 3 + b] - d * v[b];
    }
    const double T1[9] = {s1, 0, -s1 * m1x, 0, s1, -s1 * m1y, 0, 0, 1}, T2[9] = {s2, 0, -s2 * m2x, 0, s2, -s2 * m2y, 0, 0, 1};
    double tmp[9] = {0};
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) for (int k = 0; k < 3; k++) tmp[a * 3 + b] += T2[k * 3 + a] * F2[k * 3 + b];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { F[a * 3 + b] = 0; for (int k = 0; k < 3; k++) F[a * 3 + b] += tmp[a * 3 + k] * T1[k * 3 + b]; }
    return true;
}

// OpenCV's error of a correspondence under F is the larger of the two squared point-to-epipolar-line distances,
//   max( (x2.F x1)^2 / |(F x1)_xy|^2 , (x1.F^T x2)^2 / |(F^T x2)_xy|^2 ),
// and a correspondence is an inlier when it is <= threshold^2. RANSAC only needs that decision, so the count below compares
// d^2 <= thr2 * (a^2 + b^2) for both lines: no division, no branch, structure-of-arrays doubles -- a loop the host compiler turns
// into packed arithmetic (the AVX2 + FMA clone is picked at run time when the CPU has both). The scoring of ~30 hypotheses over ~600 tracks
// was two thirds of the 190 us this step took.
// (`best`: the count to beat. The points are scored in blocks of 128 and a hypothesis that cannot reach best + 1 any more is dropped;
// its count and mask are then partial, which the caller never uses: it only keeps a hypothesis whose count exceeds `best`.)
#define COEB_FM_COUNT_BODY                                                                                              \
    int cnt = 0;                                                                                                        \
    for (int i0 = 0; i0 < n; i0 += 128) {                                                                               \
        const int i1 = i0 + 128 < n ? i0 + 128 : n;                                                                     \
        int c = 0;                                                                                                      \
        for (int i = i0; i < i1; i++) {                                                                                 \
            const double a = F[0] * x1[i] + F[1] * y1[i] + F[2], b = F[3] * x1[i] + F[4] * y1[i] + F[5], cc = F[6] * x1[i] + F[7] * y1[i] + F[8]; \
            const double d = x2[i] * a + y2[i] * b + cc;   /* x2 . F x1 == x1 . F^T x2: one residual serves both lines */ \
            const double at = F[0] * x2[i] + F[3] * y2[i] + F[6], bt = F[1] * x2[i] + F[4] * y2[i] + F[7];              \
            const double dd = d * d;                                                                                    \
            const unsigned char in = (unsigned char)((dd <= thr2 * (a * a + b * b)) & (dd <= thr2 * (at * at + bt * bt))); \
            mask[i] = in;                                                                                               \
            c += in;                                                                                                    \
        }                                                                                                               \
        cnt += c;                                                                                                       \
        if (cnt + (n - i1) <= best) break;                                                                              \
    }                                                                                                                   \
    return cnt;
int fm_count_inliers_base(const double* F, const double* x1, const double* y1, const double* x2, const double* y2, int n, double thr2, unsigned char* mask, int best) {
    COEB_FM_COUNT_BODY
}
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2,fma"))) int fm_count_inliers_avx2(const double* F, const double* x1, const double* y1, const double* x2, const double* y2, int n,
                                                          double thr2, unsigned char* mask, int best) {
    COEB_FM_COUNT_BODY
}
__attribute__((target("avx512f,avx512vl,avx512bw,avx512dq,fma,prefer-vector-width=512"))) int fm_count_inliers_avx512(const double* F, const double* x1, const double* y1, const double* x2,
                                                                                                               const double* y2, int n, double thr2, unsigned char* mask, int best) {
    COEB_FM_COUNT_BODY
}
int fm_count_inliers(const double* F, const double* x1, const double* y1, const double* x2, const double* y2, int n, double thr2, unsigned char* mask, int best) {
    static const bool avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
    static const bool avx512 = avx2 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512dq") &&
                               getenv("COEB_MOTION_NO_AVX512") == nullptr;   // (development switch: the 256-bit clone)
    if (avx512) return fm_count_inliers_avx512(F, x1, y1, x2, y2, n, thr2, mask, best);
    return avx2 ? fm_count_inliers_avx2(F, x1, y1, x2, y2, n, thr2, mask, best) : fm_count_inliers_base(F, x1, y1, x2, y2, n, thr2, mask, best);
}
#else
int fm_count_inliers(const double* F, const double* x1, const double* y1, const double* x2, const double* y2, int n, double thr2, unsigned char* mask, int best) {
    return fm_count_inliers_base(F, x1, y1, x2, y2, n, thr2, mask, best);
}
#endif

}  // namespace

// Development switch (read once): the one-warp-per-point cornerSubPix / LK kernels instead of the CTA-per-point ones.
static bool motion_warp_per_point() {
    static const bool on = getenv("COEB_MOTION_WARP_PER_POINT") != nullptr;
    return on;
}

extern "C" {

int coeb_motion_create(int device, coeb_motion** out) {
    if (!out) return fail(COEB_ERR_INVALID_ARG, "null argument");
    int st = check_device(device);
    if (st != COEB_OK) return st;
    CUDA_TRY(cudaSetDevice(device));
    coeb_motion* m = new coeb_motion();
    m->device = device;
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) { delete m; return fail(COEB_ERR_CUDA, "cudaStreamCreate failed"); }
    if (cudaStreamCreateWithFlags(&m->stream2, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&m->ev_prev, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_side, cudaEventDisableTiming) != cudaSuccess) {
        coeb_motion_destroy(m);
        return fail(COEB_ERR_CUDA, "side stream / event creation failed");
    }
    cudaError_t e = cudaMalloc(&m->d_max, 16);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_cand, sizeof(float2) * kMoMaxCand);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_pre, sizeof(float2) * kMoMaxPts);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_next, sizeof(float2) * kMoMaxPts);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_status, kMoMaxPts);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_moving, kMoMaxPts);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_F, 9 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&m->d_dist, kMoMaxPts * sizeof(double));
    if (e != cudaSuccess) { coeb_motion_destroy(m); return fail(COEB_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e)); }
    *out = m;
    return COEB_OK;
}

void coeb_motion_destroy(coeb_motion* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    if (m->stream2) { cudaStreamSynchronize(m->stream2); cudaStreamDestroy(m->stream2); }
    if (m->ev_prev) cudaEventDestroy(m->ev_prev);
    if (m->ev_side) cudaEventDestroy(m->ev_side);
    motion_free_images(m);
    cudaFree(m->d_max); cudaFree(m->d_cand); cudaFree(m->d_pre); cudaFree(m->d_next); cudaFree(m->d_status); cudaFree(m->d_moving); cudaFree(m->d_F);
    cudaFree(m->d_dist); cudaFree(m->d_mask);
    if (m->h_pin) cudaFreeHost(m->h_pin);
    if (m->h_out) cudaFreeHost(m->h_out);
    for (int f = 0; f < 2; f++) if (m->h_frame[f]) cudaFreeHost(m->h_frame[f]);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

// COEB_MOTION_TRACE: host time marks inside good_features() (us since its entry): 0 frame staged + upload enqueued, 1 kernels and
// result copies enqueued, 2 side-stream work enqueued, 3 synchronised, 4 minimum-distance pass done
static thread_local double g_gf_mark[5];
int coeb_motion_good_features(coeb_motion* m, const uint8_t* gray, int width, int height, int stride, int max_corners, double quality, double min_distance,
                              double harris_k, float* xy_out, int cap, int* n_out) {
    const auto gf_t0 = std::chrono::steady_clock::now();
    auto gf_mark = [&](int k) { g_gf_mark[k] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - gf_t0).count(); };
    if (!m || !gray || !xy_out || !n_out || width < 8 || height < 8 || stride < width || cap < 0) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    *n_out = 0;
    CUDA_TRY(cudaSetDevice(m->device));
    int st = motion_prepare(m, width, height, 22, 5);
    if (st != COEB_OK) return st;
    if (!m->side.prev_resident && (st = upload_level0(m, 0, gray, stride)) != COEB_OK) return st;
    if (m->side.cur_gray) CUDA_TRY(cudaEventRecord(m->ev_prev, m->stream));
    gf_mark(0);
    if (m->info_dirty) CUDA_TRY(cudaMemsetAsync(m->d_max, 0, 16, m->stream));   // first call, or a call that failed before its sort kernel
    m->info_dirty = true;
    // Sobel scale of cornerHarris for 8-bit input: 1 / (2^(ksize-1) * blockSize * 255); the kernel taps are float(1*scale), float(2*scale)
    const double scale = 1.0 / ((double)(1 << 2) * 3 * 255.0);
    harris_response_kernel<<<dim3((width + kHtW - 1) / kHtW, (height + kHtH - 1) / kHtH), dim3(kHtW, kHtH), 0, m->stream>>>(
        m->d_pyr[0][0], width, height, m->lp[0], harris_k, (float)(1.0 * scale), (float)(2.0 * scale), m->d_resp, m->d_max);
    harris_candidates_kernel<<<dim3((width + 31) / 32, (height + 7) / 8), 256, 0, m->stream>>>(m->d_resp, width, height, m->d_max, (float)quality, m->d_cand,
                                                                                            (int*)(m->d_max + 1), kMoMaxCand);
    // the count and the first kSortCap candidates arrive in the mapped pinned block (the count sits right in front of the list)
    int st2 = motion_pin(m, 16 + sizeof(float2) * kSortCap);
    if (st2 != COEB_OK) return st2;
    SelectArgs sel{};
    const int sel_cell = min_distance >= 1 ? (int)std::lrint(min_distance) : 1;
    const bool sel_fits = ((width + sel_cell - 1) / sel_cell) * ((height + sel_cell - 1) / sel_cell) <= kSelMaxCells;   // else the host pass, decided here: no wasted attempt
    if (m->side.select_on_device && m->h_out && min_distance >= 1 && width < 65536 && height < 65536 && sel_fits) {   // coeb_process_moving_object: the minimum-distance pass stays on the device
        sel.enabled = 1; sel.w = width; sel.cell = (int)std::lrint(min_distance);
        sel.gw = (width + sel.cell - 1) / sel.cell; sel.gh = (height + sel.cell - 1) / sel.cell;
        sel.max_corners = std::min(max_corners > 0 ? max_corners : kMoMaxPts, std::min(cap, kMoMaxPts));
        sel.d2max = (int)std::ceil(min_distance * min_distance) - 1;
        sel.out_dev = m->d_pre; sel.out_host = reinterpret_cast<float2*>(m->h_out + kOutSelected);
        sel.n_dev = reinterpret_cast<int*>(m->d_max + 2); sel.n_host = reinterpret_cast<int*>(m->h_out + kOutCount);
    }
    {
        static bool configured[64] = {};
        if (!configured[m->device & 63]) { cudaFuncSetAttribute(sort_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSelSmem); configured[m->device & 63] = true; }
        sort_candidates_kernel<<<1, 1024, sel.enabled ? kSelSmem : kSortCap * 8, m->stream>>>(m->d_cand, m->d_max, m->h_pin, sel);
    }
    CUDA_TRY(cudaGetLastError());
    m->info_dirty = false;   // the sort kernel leaves the words cleared
    gf_mark(1);
    if ((st2 = enqueue_side_work(m)) != COEB_OK) return st2;
    gf_mark(2);
    if (sel.enabled) { *n_out = -1; return COEB_OK; }   // the caller goes on enqueueing; the count is in the mapped block after its synchronisation
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    gf_mark(3);
    const unsigned* info = reinterpret_cast<const unsigned*>(m->h_pin);
    if ((int)info[1] > kMoMaxCand) return fail(COEB_ERR_CAPACITY, "%u corner candidates (at most %d)", info[1], kMoMaxCand);
    const int nc = (int)info[1];
    m->cand_host.resize(nc);
    const bool sorted = nc <= kSortCap;
    if (sorted) { if (nc) std::memcpy(m->cand_host.data(), m->h_pin + 16, sizeof(float2) * nc); }
    else CUDA_TRY(cudaMemcpy(m->cand_host.data(), m->d_cand, sizeof(float2) * nc, cudaMemcpyDeviceToHost));
    const int n = select_corners(m->cand_host, sorted, width, height, max_corners, min_distance, xy_out, cap);
    gf_mark(4);
    *n_out = n;
    return n > cap ? COEB_ERR_CAPACITY : COEB_OK;
}

int coeb_motion_corner_subpix(coeb_motion* m, const uint8_t* gray, int width, int height, int stride, float* xy_inout, int n, int half_win, int max_iters,
                              double eps) {
    if (!m || !gray || (n > 0 && !xy_inout) || n < 0 || n > kMoMaxPts || half_win < 1 || 2 * half_win + 3 > kSpMaxWin) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (n == 0) return COEB_OK;
    CUDA_TRY(cudaSetDevice(m->device));
    int st = motion_prepare(m, width, height, 22, 5);
    if (st != COEB_OK) return st;
    if ((st = ensure_mask(m, half_win)) != COEB_OK) return st;
    if ((st = upload_level0(m, 0, gray, stride)) != COEB_OK) return st;
    CUDA_TRY(cudaMemcpyAsync(m->d_pre, xy_inout, sizeof(float2) * n, cudaMemcpyHostToDevice, m->stream));
    const double e = std::max(eps, 0.0);
    switch (half_win) {   // the window is a compile-time size: its per-lane offsets and weights live in registers
        case 10:
            if (motion_warp_per_point()) corner_subpix_kernel<10><<<(n + 3) / 4, 128, 0, m->stream>>>(m->d_pyr[0][0], width, height, m->lp[0], m->d_pre, n, std::max(max_iters, 1), e * e, m->d_mask);
            else corner_subpix_cta_kernel<10><<<n, 32 * kPtWarps, 0, m->stream>>>(m->d_pyr[0][0], width, height, m->lp[0], m->d_pre, n, std::max(max_iters, 1), e * e, m->d_mask, nullptr, nullptr, nullptr);
            break;
        case 5: corner_subpix_kernel<5><<<(n + 3) / 4, 128, 0, m->stream>>>(m->d_pyr[0][0], width, height, m->lp[0], m->d_pre, n, std::max(max_iters, 1), e * e, m->d_mask); break;
        case 3: corner_subpix_kernel<3><<<(n + 3) / 4, 128, 0, m->stream>>>(m->d_pyr[0][0], width, height, m->lp[0], m->d_pre, n, std::max(max_iters, 1), e * e, m->d_mask); break;
        default: return fail(COEB_ERR_UNSUPPORTED, "cornerSubPix half window %d (built for 3, 5 and 10)", half_win);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(xy_inout, m->d_pre, sizeof(float2) * n, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return COEB_OK;
}

static int run_lk(coeb_motion* m, int n, int win, int max_iters, double eps, double min_eig, int edge, float sad_limit, bool images_on_side_stream = false,
                  float2* next_host = nullptr, uint8_t* status_host = nullptr, bool* mirrored = nullptr, const int* n_ptr = nullptr) {
    if (mirrored) *mirrored = false;
    if (images_on_side_stream) {
        CUDA_TRY(cudaStreamWaitEvent(m->stream, m->ev_side, 0));
    } else {
        build_pyramid(m, 0);
        build_pyramid(m, 1);
        derivative_images(m, m->stream);
    }
    LkLevels L{};
    L.nlevels = m->nlevels;
    for (int l = 0; l < m->nlevels; l++) {
        L.w[l] = m->lw[l]; L.h[l] = m->lh[l]; L.pitch[l] = m->lp[l];
        L.prev[l] = m->d_pyr[0][l]; L.cur[l] = m->d_pyr[1][l]; L.deriv[l] = m->d_deriv[l];
    }
    const bool warp_per_point = motion_warp_per_point();
    if (win == 22 && !warp_per_point)
    {
        lk_cta_kernel<22><<<n, 32 * kPtWarps, 0, m->stream>>>(L, m->d_pre, n, max_iters, (float)(eps * eps), (float)min_eig, m->d_next, m->d_status, edge, sad_limit, next_host,
                                                              status_host, n_ptr);
        if (mirrored) *mirrored = next_host && status_host;
    }
    else if (win == 22)
        lk_kernel_w<22><<<(n + kLkWarps - 1) / kLkWarps, 32 * kLkWarps, 0, m->stream>>>(L, m->d_pre, n, max_iters, (float)(eps * eps), (float)min_eig, m->d_next, m->d_status, edge,
                                                                                  sad_limit);
    else
        lk_kernel<<<(n + kLkWarps - 1) / kLkWarps, 32 * kLkWarps, 0, m->stream>>>(L, m->d_pre, n, win, max_iters, (float)(eps * eps), (float)min_eig, m->d_next, m->d_status, edge,
                                                                            sad_limit);
    CUDA_TRY(cudaGetLastError());
    return COEB_OK;
}

int coeb_motion_lk(coeb_motion* m, const uint8_t* prev_gray, const uint8_t* cur_gray, int width, int height, int stride, const float* prev_xy, int n, int win,
                   int max_level, int max_iters, double eps, double min_eig_threshold, float* next_xy, uint8_t* status) {
    if (!m || !prev_gray || !cur_gray || n < 0 || n > kMoMaxPts || (n > 0 && (!prev_xy || !next_xy || !status)) || win < 3 || win > kLkMaxWin || max_level < 0)
        return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (n == 0) return COEB_OK;
    CUDA_TRY(cudaSetDevice(m->device));
    m->w = 0;   // the level layout depends on (win, max_level): lay it out again
    int st = motion_prepare(m, width, height, win, max_level);
    if (st != COEB_OK) return st;
    if ((st = upload_level0(m, 0, prev_gray, stride)) != COEB_OK) return st;
    if ((st = upload_level0(m, 1, cur_gray, stride)) != COEB_OK) return st;
    CUDA_TRY(cudaMemcpyAsync(m->d_pre, prev_xy, sizeof(float2) * n, cudaMemcpyHostToDevice, m->stream));
    if ((st = run_lk(m, n, win, std::max(max_iters, 1), eps, min_eig_threshold, -1, 0.f)) != COEB_OK) return st;
    CUDA_TRY(cudaMemcpyAsync(next_xy, m->d_next, sizeof(float2) * n, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaMemcpyAsync(status, m->d_status, n, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return COEB_OK;
}

int coeb_fundamental_ransac(const float* p1_xy, const float* p2_xy, int n, double threshold, double confidence, int max_iters, unsigned seed, double F_out[9],
                            uint8_t* inlier_mask, int* n_inliers) {
    if (!p1_xy || !p2_xy || !F_out || n < 0) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (n_inliers) *n_inliers = 0;
    if (n < 8) return fail(COEB_ERR_INVALID_ARG, "the fundamental matrix needs at least 8 correspondences (got %d)", n);
    const double thr2 = threshold * threshold;
    uint64_t rng = 0x9E3779B97F4A7C15ull ^ ((uint64_t)seed * 0xD1342543DE82EF95ull + 1);
    auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (uint32_t)(rng >> 32); };
    std::vector<uint8_t> best_mask(n, 0), mask(n);
    std::vector<double> soa(4 * (size_t)n);
    double *sx1 = soa.data(), *sy1 = sx1 + n, *sx2 = sy1 + n, *sy2 = sx2 + n;
    for (int i = 0; i < n; i++) { sx1[i] = p1_xy[2 * i]; sy1[i] = p1_xy[2 * i + 1]; sx2[i] = p2_xy[2 * i]; sy2[i] = p2_xy[2 * i + 1]; }
    int best = 0, iters = std::max(max_iters, 1);
    double Fb[9] = {0}, F[9];
    for (int it = 0; it < iters; it++) {
        int idx[8];
        for (int k = 0; k < 8;) {   // 8 distinct indices
            const int c = (int)(next() % (uint32_t)n);
            bool dup = false;
            for (int j = 0; j < k; j++) dup |= idx[j] == c;
            if (!dup) idx[k++] = c;
        }
        if (!eight_point_minimal(p1_xy, p2_xy, idx, F)) continue;
        const int cnt = fm_count_inliers(F, sx1, sy1, sx2, sy2, n, thr2, mask.data(), best);
        if (cnt > best) {
            best = cnt; best_mask.swap(mask); std::memcpy(Fb, F, sizeof(F));
            // cv::RANSACUpdateNumIters: log(1 - confidence) / log(1 - inlier_ratio^8)
            const double ep = 1.0 - (double)cnt / n, num = std::log(std::max(1.0 - confidence, DBL_MIN)), den = std::log(std::max(1.0 - std::pow(1.0 - ep, 8), DBL_MIN));
            if (den < 0 && -num < (double)iters * -den) iters = std::max(it + 1, (int)std::lrint(num / den));
        }
    }
    if (best < 8) return fail(COEB_ERR_UNSUPPORTED, "RANSAC found no model with 8 inliers");
    std::vector<int> in;
    for (int i = 0; i < n; i++) if (best_mask[i]) in.push_back(i);
    if (!eight_point(p1_xy, p2_xy, in.data(), (int)in.size(), F)) std::memcpy(F, Fb, sizeof(F));   // final fit on the consensus set
    std::memcpy(F_out, F, sizeof(F));
    if (inlier_mask) std::memcpy(inlier_mask, best_mask.data(), n);
    if (n_inliers) *n_inliers = best;
    return COEB_OK;
}

int coeb_epipolar_outliers(coeb_motion* m, const float* pre_xy, const float* next_xy, const uint8_t* status, int n, const double F[9], double limit,
                           uint8_t* moving_out, double* dist_out) {
    if (!m || n < 0 || n > kMoMaxPts || (n > 0 && (!pre_xy || !next_xy || !status || !moving_out)) || !F) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (n == 0) return COEB_OK;
    CUDA_TRY(cudaSetDevice(m->device));
    CUDA_TRY(cudaMemcpyAsync(m->d_pre, pre_xy, sizeof(float2) * n, cudaMemcpyHostToDevice, m->stream));
    CUDA_TRY(cudaMemcpyAsync(m->d_next, next_xy, sizeof(float2) * n, cudaMemcpyHostToDevice, m->stream));
    CUDA_TRY(cudaMemcpyAsync(m->d_status, status, n, cudaMemcpyHostToDevice, m->stream));
    CUDA_TRY(cudaMemcpyAsync(m->d_F, F, 9 * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    epipolar_kernel<<<(n + 127) / 128, 128, 0, m->stream>>>(m->d_pre, m->d_next, m->d_status, n, m->d_F, limit, m->d_moving, m->d_dist);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(moving_out, m->d_moving, n, cudaMemcpyDeviceToHost, m->stream));
    if (dist_out) CUDA_TRY(cudaMemcpyAsync(dist_out, m->d_dist, sizeof(double) * n, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return COEB_OK;
}

// prev_resident: the previous frame is the current frame of the last call, still on the device with its pyramid (d_pyr[1]): the two
// pyramid buffers swap roles (unless `swapped` says a first attempt of this call already did) and nothing of it is uploaded or rebuilt.
static int process_impl(coeb_motion* m, const uint8_t* prev_gray, const uint8_t* cur_gray, int width, int height, int stride, float* tm_xy_out, int cap,
                        int* n_tm_out, coeb_motion_trace* trace, bool prev_resident, bool swapped) {
    *n_tm_out = 0;
    if (trace) { trace->n_points = trace->n_tracked = trace->n_inliers = 0; trace->have_F = 0; }
    CUDA_TRY(cudaSetDevice(m->device));
    if (prev_resident && !swapped)
        for (int l = 0; l < kMoMaxLevels; l++) std::swap(m->d_pyr[0][l], m->d_pyr[1][l]);
    m->cur_resident = false;   // until this call has left its current frame and pyramid behind
    if (prev_resident) prev_gray = cur_gray;   // (never read: good_features skips the upload)
    const bool timeline = getenv("COEB_MOTION_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    const auto t0 = now();
    // goodFeaturesToTrack + cornerSubPix on the previous frame (:333-334)
    std::vector<float> pre(2 * 1000);
    int n = 0;
    static const bool one_stream = getenv("COEB_MOTION_ONE_STREAM") != nullptr;   // development switch
    // refined corners, tracked positions and states are written by the kernels into this mapped pinned block as well
    if (!m->h_out) CUDA_TRY(cudaHostAlloc((void**)&m->h_out, kOutBytes, cudaHostAllocMapped));
    float2* const h_pre = reinterpret_cast<float2*>(m->h_out + kOutPre);
    float2* const h_next = reinterpret_cast<float2*>(m->h_out + kOutNext);
    uint8_t* const h_state = reinterpret_cast<uint8_t*>(m->h_out + kOutState);
    float2* const h_in = reinterpret_cast<float2*>(m->h_out + kOutSelected);   // the selected corners
    const bool warp_per_point = motion_warp_per_point();
    // The minimum-distance pass normally runs on the device, behind the sort (SelectArgs): the corners never visit the host and the call
    // synchronises once, after LK. COEB_MOTION_HOST_SELECT (read per call, so that a test can run both) keeps the host pass.
    m->side.select_on_device = !warp_per_point && !m->force_host_select && getenv("COEB_MOTION_HOST_SELECT") == nullptr;
    m->side.cur_gray = one_stream ? nullptr : cur_gray; m->side.stride = stride; m->side.done = false;
    m->side.prev_resident = prev_resident;
    int st = coeb_motion_good_features(m, prev_gray, width, height, stride, 1000, 0.01, 8.0, 0.04, pre.data(), 1000, &n);
    m->side.cur_gray = nullptr;
    m->side.select_on_device = false;
    m->side.prev_resident = false;
    const bool side_done = m->side.done;
    const bool dev_sel = st == COEB_OK && n == -1;   // the corners and their number are on the device (and on their way into h_out)
    if (st != COEB_OK || n == 0) { if (side_done) cudaStreamSynchronize(m->stream2); }   // nothing follows: the side stream must not outlive the call
    if (st != COEB_OK) return st;
    if (n == 0) { m->cur_resident = side_done; m->res_w = width; m->res_h = height; return COEB_OK; }
    const auto t1 = now();
    if ((st = ensure_mask(m, 10)) != COEB_OK) return st;
    bool pre_mirrored = false, lk_mirrored = false;
    const int n_launch = dev_sel ? 1000 : n;                                      // CTAs beyond the device's count leave at once
    const int* const n_ptr = dev_sel ? reinterpret_cast<const int*>(m->d_max + 2) : nullptr;
    if (!dev_sel) {
        if (warp_per_point) CUDA_TRY(cudaMemcpyAsync(m->d_pre, pre.data(), sizeof(float2) * n, cudaMemcpyHostToDevice, m->stream));   // level 0 of the previous frame is resident
        else std::memcpy(h_in, pre.data(), sizeof(float2) * n);   // read by the sub-pixel kernel across PCIe (8 bytes per CTA), no upload node
    }
    if (warp_per_point) corner_subpix_kernel<10><<<(n + 3) / 4, 128, 0, m->stream>>>(m->d_pyr[0][0], width, height, m->lp[0], m->d_pre, n, 20, 0.03 * 0.03, m->d_mask);
    else corner_subpix_cta_kernel<10><<<n_launch, 32 * kPtWarps, 0, m->stream>>>(m->d_pyr[0][0], width, height, m->lp[0], m->d_pre, n_launch, 20, 0.03 * 0.03, m->d_mask, h_pre,
                                                                                 dev_sel ? nullptr : h_in, n_ptr);
    pre_mirrored = !warp_per_point;
    // calcOpticalFlowPyrLK + border / SAD tests (:335-364)
    if (!side_done && (st = upload_level0(m, 1, cur_gray, stride)) != COEB_OK) return st;
    if ((st = run_lk(m, n_launch, 22, 20, 0.01, 1e-4, /*limit_edge_corner*/ 5, /*limit_of_check*/ 2120.f, side_done, h_next, h_state, &lk_mirrored, n_ptr)) != COEB_OK) return st;
    if (!dev_sel) {
        if (!pre_mirrored) CUDA_TRY(cudaMemcpyAsync(pre.data(), m->d_pre, sizeof(float2) * n, cudaMemcpyDeviceToHost, m->stream));
    }
    std::vector<float> nxt;
    std::vector<uint8_t> state;
    if (!lk_mirrored) {
        nxt.resize(2 * (size_t)n); state.resize(n);
        CUDA_TRY(cudaMemcpyAsync(nxt.data(), m->d_next, sizeof(float2) * n, cudaMemcpyDeviceToHost, m->stream));
        CUDA_TRY(cudaMemcpyAsync(state.data(), m->d_status, n, cudaMemcpyDeviceToHost, m->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    m->cur_resident = true; m->res_w = width; m->res_h = height;   // LK has run: the current frame and its pyramid are in d_pyr[1]
    if (dev_sel) {
        n = *reinterpret_cast<const int*>(m->h_out + kOutCount);
        if (n < 0 && timeline) fprintf(stderr, "[coeb motion] device selection declined (code %d): host pass\n", -n);
        if (n < 0) {   // the device left the selection to the host (more candidates than it sorts, or a grid that does not fit): once more, the host way
            m->force_host_select = true;
            st = process_impl(m, prev_gray, cur_gray, width, height, stride, tm_xy_out, cap, n_tm_out, trace, prev_resident, /*swapped*/ true);
            m->force_host_select = false;
            return st;
        }
        if (n == 0) return COEB_OK;
    }
    if (trace) trace->n_points = n;
    if (pre_mirrored) std::memcpy(pre.data(), h_pre, sizeof(float2) * n);
    if (lk_mirrored) { nxt.resize(2 * (size_t)n); state.resize(n); std::memcpy(nxt.data(), h_next, sizeof(float2) * n); std::memcpy(state.data(), h_state, n); }
    const auto t2 = now();
    // findFundamentalMat(F_prepoint, F_nextpoint, FM_RANSAC, 0.1, 0.99) on the surviving pairs (:353-370)
    std::vector<float> f1, f2;
    for (int i = 0; i < n; i++)
        if (state[i]) { f1.push_back(pre[2 * i]); f1.push_back(pre[2 * i + 1]); f2.push_back(nxt[2 * i]); f2.push_back(nxt[2 * i + 1]); }
    const int nf = (int)f1.size() / 2;
    if (trace) {
        trace->n_tracked = nf;
        const int c = std::min(n, COEB_MOTION_TRACE_POINTS);
        std::memcpy(trace->pre_xy, pre.data(), sizeof(float) * 2 * c);
        std::memcpy(trace->next_xy, nxt.data(), sizeof(float) * 2 * c);
        std::memcpy(trace->state, state.data(), c);
    }
    double F[9];
    int ninl = 0;
    if (nf < 8 || coeb_fundamental_ransac(f1.data(), f2.data(), nf, 0.1, 0.99, 1000, 12345u, F, nullptr, &ninl) != COEB_OK) return COEB_OK;   // no model: no T_M
    if (trace) { trace->n_inliers = ninl; trace->have_F = 1; std::memcpy(trace->F, F, sizeof(F)); }
    const auto t3 = now();
    // epipolar distance > 1 -> T_M, in point order (:372-385). The tracks are on the host already (the RANSAC ran on them) and there are
    // a few hundred of them: the test runs here, in the operation order of epipolar_kernel (IEEE double, no contraction: same decisions),
    // instead of costing the call another upload, launch, download and synchronisation (22 us).
    static const bool device_epipolar = getenv("COEB_MOTION_DEVICE_EPIPOLAR") != nullptr;   // development switch
    std::vector<uint8_t> mv(n, 0);
    if (device_epipolar) {
        CUDA_TRY(cudaMemcpyAsync(m->d_F, F, sizeof(F), cudaMemcpyHostToDevice, m->stream));
        epipolar_kernel<<<(n + 127) / 128, 128, 0, m->stream>>>(m->d_pre, m->d_next, m->d_status, n, m->d_F, 1.0, m->d_moving, nullptr);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(mv.data(), m->d_moving, n, cudaMemcpyDeviceToHost, m->stream));
        CUDA_TRY(cudaStreamSynchronize(m->stream));
    } else {
        for (int i = 0; i < n; i++) {
            if (!state[i]) continue;
            const volatile double px = pre[2 * i], py = pre[2 * i + 1], qx = nxt[2 * i], qy = nxt[2 * i + 1];
            const volatile double A = (F[0] * px + F[1] * py) + F[2];
            const volatile double B = (F[3] * px + F[4] * py) + F[5];
            const volatile double C = (F[6] * px + F[7] * py) + F[8];
            const volatile double num = std::fabs((A * qx + B * qy) + C), den = std::sqrt(A * A + B * B);
            const double dd = num / den;
            mv[i] = !(dd <= 1.0);
        }
    }
    int k = 0;
    for (int i = 0; i < n; i++)
        if (mv[i]) {
            if (k < cap) { tm_xy_out[2 * k] = nxt[2 * i]; tm_xy_out[2 * k + 1] = nxt[2 * i + 1]; }
            k++;
        }
    *n_tm_out = k;
    if (timeline)
        fprintf(stderr, "[coeb motion] corners (upload, Harris, candidates, sort, min-distance) %.0f us [staged %.0f, enqueued %.0f, side work %.0f, synchronised %.0f, selected %.0f], "
                        "subpix + pyramids + LK %.0f us, RANSAC %.0f us (%d of %d inliers), epipolar %.0f us\n",
                us(t0, t1), g_gf_mark[0], g_gf_mark[1], g_gf_mark[2], g_gf_mark[3], g_gf_mark[4], us(t1, t2), us(t2, t3), ninl, nf, us(t3, now()));
    return k > cap ? COEB_ERR_CAPACITY : COEB_OK;
}

int coeb_process_moving_object(coeb_motion* m, const uint8_t* prev_gray, const uint8_t* cur_gray, int width, int height, int stride, float* tm_xy_out, int cap,
                               int* n_tm_out, coeb_motion_trace* trace) {
    if (!m || !prev_gray || !cur_gray || !n_tm_out || cap < 0 || (cap > 0 && !tm_xy_out)) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    return process_impl(m, prev_gray, cur_gray, width, height, stride, tm_xy_out, cap, n_tm_out, trace, false, false);
}

int coeb_process_moving_object_next(coeb_motion* m, const uint8_t* cur_gray, int width, int height, int stride, float* tm_xy_out, int cap, int* n_tm_out,
                                    coeb_motion_trace* trace) {
    if (!m || !cur_gray || !n_tm_out || cap < 0 || (cap > 0 && !tm_xy_out) || width < 8 || height < 8 || stride < width) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (m->cur_resident && m->res_w == width && m->res_h == height)
        return process_impl(m, nullptr, cur_gray, width, height, stride, tm_xy_out, cap, n_tm_out, trace, /*prev_resident*/ true, false);
    // first frame of a sequence (`if (imGrayPre.data)` fails, src/Frame.cc:164): nothing to compare with; the frame and its pyramid stay on the device
    *n_tm_out = 0;
    if (trace) { trace->n_points = trace->n_tracked = trace->n_inliers = 0; trace->have_F = 0; }
    CUDA_TRY(cudaSetDevice(m->device));
    m->cur_resident = false;
    int st = motion_prepare(m, width, height, 22, 5);
    if (st != COEB_OK) return st;
    if ((st = upload_level0(m, 1, cur_gray, stride)) != COEB_OK) return st;
    build_pyramid(m, 1);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    m->cur_resident = true; m->res_w = width; m->res_h = height;
    return COEB_OK;
}

}  // extern "C"
