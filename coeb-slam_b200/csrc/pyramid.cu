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
// The per-column (offset, a0|a1<<16) and per-row (row0|row1<<16, b0|b1<<16) tables are built on the host (coeb_api.cu) with the
// exact double/float arithmetic of cv::resize and kept resident.
// Each thread produces 4 adjacent output pixels and stores them as one 32-bit word.
// ------------------------------------------------------------------------------------------------
// kResizeRows = destination rows per thread (rows dy, dy+8, ... of a 128 x 8*kResizeRows tile): 4 for batches, 1 when the
// grid would otherwise be too small to fill the GPU (single-frame latency path).
//
// Fast path (scale factors up to 2, i.e. every ORB pyramid in practice): the two source pixels of an output pixel are
// cut out of two aligned source words with one PRMT whose selector depends only on the column (computed once per
// thread), and h = a0*S[sx] + a1*S[sx+1] is one IDP.2A with the table's packed (a0, a1) halfwords as they are. Output
// pixels 0,1 and 2,3 of a thread each share one pair of source words per source row.
constexpr int kRzGroups = 16, kRzRowsPerBlock = 16;   // block = 16 four-pixel groups x 16 rows
#ifndef COEB_RZ_MINB
#define COEB_RZ_MINB 8   // 32 registers, full occupancy: pyramid 0.236 -> 0.230 ms (the kernel waits on loads)
#endif
template <int kResizeRows>
__global__ void __launch_bounds__(256, COEB_RZ_MINB) resize_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v,
                                                     int level) {
    COEB_TRACE(v, 8);
    const LevelGeom& D = g.lv[level];
    const LevelGeom& S = g.lv[level - 1];
    const int frame = blockIdx.z;
    const int dx0 = (blockIdx.x * kRzGroups + threadIdx.x) * 4;   // a warp covers 2 rows x 64 pixels: less of it hangs over the row's end than with 1 x 128
    if (dx0 >= D.w) return;
    const uint8_t* __restrict__ src = level_ptr(g, v, level - 1, frame);
    const int spitch = level_pitch(g, v, level - 1);
    uint8_t* dst = v.pyr + D.img_base + (unsigned long long)frame * D.img_stride;
    const int2* __restrict__ xt = v.tabs + D.tab_base;
    const int2* __restrict__ yt = xt + D.w;
    // column entries (source offset, packed weights) are the same for every row: fetch them once
    uint32_t wgt[4], sel[4];
    const uint8_t* colp[2];   // source column of each pixel pair's two aligned words (row 0)
#pragma unroll
    for (int u = 0; u < 2; u++) {
        const int2 e0 = __ldg(&xt[min(dx0 + 2 * u, D.w - 1)]), e1 = __ldg(&xt[min(dx0 + 2 * u + 1, D.w - 1)]);   // tail -> row padding
        const int wb = min(e0.x & ~3, spitch - 8);   // two aligned words from here hold both pixels' source pairs
        const int o0 = e0.x - wb, o1 = e1.x - wb;
        sel[2 * u] = (uint32_t)(o0 | (min(o0 + 1, 7) << 4));       // beyond byte 7 only when sx is the last column: weight 0
        sel[2 * u + 1] = (uint32_t)(o1 | (min(o1 + 1, 7) << 4));
        wgt[2 * u] = (uint32_t)e0.y;
        wgt[2 * u + 1] = (uint32_t)e1.y;
        colp[u] = src + wb;
    }
    uint8_t* out = dst + dx0;
#pragma unroll
    for (int j = 0; j < kResizeRows; j++) {
        const int dy = blockIdx.y * (kRzRowsPerBlock * kResizeRows) + threadIdx.y + kRzRowsPerBlock * j;
        if (dy >= D.h) break;
        const int2 ye = __ldg(&yt[dy]);
        const int sy0 = ye.x & 0xFFFF, sy1 = (int)((uint32_t)ye.x >> 16);   // clamped on the host
        const uint32_t b0 = ye.y & 0xFFFF, b1 = (uint32_t)ye.y >> 16;
        uint32_t o[4];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const uint32_t* t = reinterpret_cast<const uint32_t*>(colp[u] + (long long)sy0 * spitch);
            const uint32_t* q = reinterpret_cast<const uint32_t*>(colp[u] + (long long)sy1 * spitch);
            const uint32_t t0 = __ldg(t), t1 = __ldg(t + 1), q0 = __ldg(q), q1 = __ldg(q + 1);
#pragma unroll
            for (int i = 2 * u; i < 2 * u + 2; i++) {
                const uint32_t h0 = __dp2a_lo(wgt[i], __byte_perm(t0, t1, sel[i]), 0u);
                const uint32_t h1 = __dp2a_lo(wgt[i], __byte_perm(q0, q1, sel[i]), 0u);
                o[i] = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2u) >> 2;   // <= 255 by construction (a multiply-high form was measured slower)
            }
        }
        // pitch is a multiple of 64: padding absorbs the tail
        *reinterpret_cast<uint32_t*>(out + (long long)dy * D.pitch) = __byte_perm(__byte_perm(o[0], o[1], 0x0040), __byte_perm(o[2], o[3], 0x0040), 0x5410);
    }
}

// Any scale factor: four byte loads per output pixel.
__global__ void __launch_bounds__(256) resize_generic_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v,
                                                             int level) {
    const LevelGeom& D = g.lv[level];
    const LevelGeom& S = g.lv[level - 1];
    const int frame = blockIdx.z;
    const int dx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int dy = blockIdx.y * 8 + threadIdx.y;
    if (dx0 >= D.w || dy >= D.h) return;
    const uint8_t* __restrict__ src = level_ptr(g, v, level - 1, frame);
    const int spitch = level_pitch(g, v, level - 1);
    uint8_t* dst = v.pyr + D.img_base + (unsigned long long)frame * D.img_stride;
    const int2* __restrict__ xt = v.tabs + D.tab_base;
    const int2 ye = __ldg(&xt[D.w + dy]);
    const int sy0 = ye.x & 0xFFFF, sy1 = (int)((uint32_t)ye.x >> 16);   // clamped on the host
    const int b0 = ye.y & 0xFFFF, b1 = ye.y >> 16;
    const uint8_t* r0 = src + (size_t)sy0 * spitch;
    const uint8_t* r1 = src + (size_t)sy1 * spitch;
    uint32_t packed = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int2 xe = __ldg(&xt[min(dx0 + i, D.w - 1)]);
        const int sx = xe.x, sx1 = min(xe.x + 1, S.w - 1), a0 = xe.y & 0xFFFF, a1 = xe.y >> 16;
        const int h0 = __ldg(r0 + sx) * a0 + __ldg(r0 + sx1) * a1;
        const int h1 = __ldg(r1 + sx) * a0 + __ldg(r1 + sx1) * a1;
        const int o = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        packed |= (uint32_t)min(max(o, 0), 255) << (8 * i);
    }
    *reinterpret_cast<uint32_t*>(dst + (size_t)dy * D.pitch + dx0) = packed;
}

void launch_pyramid(const Geometry& g, const BatchView& v, cudaStream_t stream) {
    for (int l = 1; l < g.nlevels; l++) {
        dim3 block(32, 8);
        const int tiles_x = (g.lv[l].w + 127) / 128;
        const dim3 rz_block(kRzGroups, kRzRowsPerBlock);
        const int rz_x = (g.lv[l].w + 4 * kRzGroups - 1) / (4 * kRzGroups);
        // consecutive output columns are at most 2 source columns apart up to a 2:1 reduction: the paired-word path applies
        const bool paired = 2LL * g.lv[l].w >= g.lv[l - 1].w && g.lv[l - 1].w >= 8;
        if (!paired) {
            resize_generic_kernel<<<dim3(tiles_x, (g.lv[l].h + 7) / 8, v.B), block, 0, stream>>>(g, v, l);
        } else if ((long long)rz_x * ((g.lv[l].h + 4 * kRzRowsPerBlock - 1) / (4 * kRzRowsPerBlock)) * v.B >= 2 * 148) {
            resize_kernel<4><<<dim3(rz_x, (g.lv[l].h + 4 * kRzRowsPerBlock - 1) / (4 * kRzRowsPerBlock), v.B), rz_block, 0, stream>>>(g, v, l);
        } else {
            resize_kernel<1><<<dim3(rz_x, (g.lv[l].h + kRzRowsPerBlock - 1) / kRzRowsPerBlock, v.B), rz_block, 0, stream>>>(g, v, l);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Small batches (the tracking thread's single frame): the chain above is seven dependent launches of ~4 us each, whatever the level
// size, i.e. latency and not work. Here ONE launch builds the whole pyramid: the image is cut into regions, a CTA owns the same
// relative rectangle of every level and computes the chain for it in shared memory, level after level, with the same tables and the
// same arithmetic. A level-l rectangle reads a slightly larger rectangle of level l-1 than the CTA owns there, so the computed
// rectangles carry a halo that grows towards level 1 (about 15 px) and neighbouring CTAs compute those pixels twice (identical
// values; only the owner writes them to global memory). The rectangles come from the host (build_pyramid_regions), which walks the
// very tables the kernel uses, so every source pixel a CTA reads is inside the rectangle it computed one level up.
// ------------------------------------------------------------------------------------------------
#ifndef COEB_PR_THREADS
#define COEB_PR_THREADS 1024
#endif
#ifndef COEB_PR_CELL
#define COEB_PR_CELL 60
#endif
constexpr int kPrThreads = COEB_PR_THREADS;

__global__ void __launch_bounds__(kPrThreads) pyramid_regions_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v) {
    COEB_TRACE(v, 7);
    extern __shared__ __align__(16) uint8_t pr_smem[];
    __shared__ PyrRegionLevel R[COEB_MAX_LEVELS];
    const int frame = blockIdx.y;
    static_assert(sizeof(PyrRegionLevel) == 24, "copied as six words");
    // Global round trips are what this kernel's time is made of, so there are two: the region's rectangles, then -- all loads in
    // flight together -- the table slices of every level it touches and the level-0 rectangle. The level loop below only waits on
    // shared memory. Slices live behind the image rectangles (int2 units from s_tab): per level the columns [cx0, cx1) clamped to
    // the last column, then the rows [cy0, cy1).
    if (threadIdx.x < 6 * g.nlevels) reinterpret_cast<int*>(R)[threadIdx.x] = __ldg(reinterpret_cast<const int*>(v.pyr_regions + (size_t)blockIdx.x * g.nlevels) + threadIdx.x);
    __syncthreads();
    int2* const s_tab = reinterpret_cast<int2*>(pr_smem + R[g.nlevels - 1].soff + ((R[g.nlevels - 1].pitch * (R[g.nlevels - 1].cy1 - R[g.nlevels - 1].cy0) + 15) & ~15));
    {
        int total = 0;
        for (int l = 1; l < g.nlevels; l++) total += (R[l].cx1 - R[l].cx0) + (R[l].cy1 - R[l].cy0);
        for (int i = threadIdx.x; i < total; i += kPrThreads) {
            int l = 1, j = i;
            for (; l < g.nlevels - 1; l++) {
                const int n = (R[l].cx1 - R[l].cx0) + (R[l].cy1 - R[l].cy0);
                if (j < n) break;
                j -= n;
            }
            const LevelGeom& D = g.lv[l];
            const int2* __restrict__ xt = v.tabs + D.tab_base;
            const int cw = R[l].cx1 - R[l].cx0;
            s_tab[i] = j < cw ? __ldg(&xt[min(R[l].cx0 + j, D.w - 1)]) : __ldg(&xt[D.w + R[l].cy0 + (j - cw)]);   // tail -> row padding, like the chain
        }
    }
    {   // level 0: the rectangle the chain reads, as aligned words (base, pitch and cx0 are multiples of 4)
        const PyrRegionLevel r = R[0];
        const int wq = (r.cx1 - r.cx0) >> 2, ch = r.cy1 - r.cy0;
        const int pitch = level_pitch(g, v, 0);
        const uint8_t* __restrict__ src = level_ptr(g, v, 0, frame) + r.cx0;
        const uint32_t rcp = 0xFFFFFFFFu / (uint32_t)wq + 1u;   // i / wq == umulhi(i, rcp) for wq >= 2 and every i here (checked on the host)
#pragma unroll 4
        for (int i = threadIdx.x; i < wq * ch; i += kPrThreads) {
            const int y = wq > 1 ? (int)__umulhi((uint32_t)i, rcp) : i, q = i - y * wq;
            reinterpret_cast<uint32_t*>(pr_smem + r.soff + y * r.pitch)[q] = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)(r.cy0 + y) * pitch) + q);
        }
    }
    int toff = 0;
    for (int l = 1; l < g.nlevels; l++) {
        __syncthreads();
        const PyrRegionLevel d = R[l], s = R[l - 1];
        const LevelGeom& D = g.lv[l];
        const int cw = d.cx1 - d.cx0, ch = d.cy1 - d.cy0, gq = cw >> 2;
        const int2* xs = s_tab + toff;
        const int2* ys = xs + cw;
        toff += cw + ch;
        const int sw1 = g.lv[l - 1].w - 1;
        const uint8_t* S = pr_smem + s.soff - s.cy0 * s.pitch - s.cx0;   // indexed by absolute level coordinates
        uint8_t* O = pr_smem + d.soff;
        uint8_t* dst = v.pyr + D.img_base + (unsigned long long)frame * D.img_stride;
        // (row, 4-pixel group) tasks spread over all threads: a level is a handful of dependent shared-memory round trips
        const uint32_t rcp = 0xFFFFFFFFu / (uint32_t)gq + 1u;   // i / gq == umulhi(i, rcp), as above
        for (int i = threadIdx.x; i < gq * ch; i += kPrThreads) {
            const int y = gq > 1 ? (int)__umulhi((uint32_t)i, rcp) : i, q = i - y * gq;
            const int dx0 = d.cx0 + 4 * q, dy = d.cy0 + y;
            const int2 ye = ys[y];
            const uint8_t* r0 = S + (ye.x & 0xFFFF) * s.pitch;
            const uint8_t* r1 = S + (int)((uint32_t)ye.x >> 16) * s.pitch;
            const uint32_t b0 = (uint32_t)ye.y & 0xFFFFu, b1 = (uint32_t)ye.y >> 16;
            uint32_t packed = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int2 xe = xs[4 * q + k];
                const int sx = xe.x, sx1 = min(xe.x + 1, sw1);
                const uint32_t a0 = (uint32_t)xe.y & 0xFFFFu, a1 = (uint32_t)xe.y >> 16;
                const uint32_t h0 = r0[sx] * a0 + r0[sx1] * a1;
                const uint32_t h1 = r1[sx] * a0 + r1[sx1] * a1;
                packed |= ((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2u) >> 2) << (8 * k);
            }
            *reinterpret_cast<uint32_t*>(O + y * d.pitch + 4 * q) = packed;
            if (dx0 >= d.ox0 && dx0 < d.ox1 && dy >= d.oy0 && dy < d.oy1) *reinterpret_cast<uint32_t*>(dst + (size_t)dy * D.pitch + dx0) = packed;
        }
    }
}

// Region table: regions x levels. x bounds of owned and computed rectangles are multiples of 4 (the last one rounded up into the
// row padding, which the chain's last 4-pixel group writes as well). Returns 0 when the image is too small to cut.
int build_pyramid_regions(const Geometry& g, const int2* tabs, PyrRegionLevel* out, int* smem_bytes) {
    const int nl = g.nlevels;
    *smem_bytes = 0;
    if (nl < 2) return 0;
    const LevelGeom& T = g.lv[nl - 1];
    const int RX = std::max(1, std::min(std::min(16, (g.w0 + COEB_PR_CELL / 2) / COEB_PR_CELL), T.w / 8));
    const int RY = std::max(1, std::min(std::min(16, (g.h0 + COEB_PR_CELL / 2) / COEB_PR_CELL), T.h / 4));
    auto up4 = [](int x) { return (x + 3) & ~3; };
    int worst = 0;
    for (int j = 0; j < RY; j++)
        for (int i = 0; i < RX; i++) {
            PyrRegionLevel r[COEB_MAX_LEVELS];
            for (int l = 0; l < nl; l++) {
                const LevelGeom& L = g.lv[l];
                r[l].ox0 = (short)(((long long)i * L.w / RX) & ~3LL);
                r[l].ox1 = (short)(i + 1 == RX ? up4(L.w) : (((long long)(i + 1) * L.w / RX) & ~3LL));
                r[l].oy0 = (short)((long long)j * L.h / RY);
                r[l].oy1 = (short)((long long)(j + 1) * L.h / RY);
                if (r[l].ox1 <= r[l].ox0 || r[l].oy1 <= r[l].oy0) return 0;
            }
            r[nl - 1].cx0 = r[nl - 1].ox0; r[nl - 1].cx1 = r[nl - 1].ox1; r[nl - 1].cy0 = r[nl - 1].oy0; r[nl - 1].cy1 = r[nl - 1].oy1;
            for (int l = nl - 1; l >= 1; l--) {
                const LevelGeom& D = g.lv[l];
                const LevelGeom& S = g.lv[l - 1];
                const int2* xt = tabs + D.tab_base;
                const int2* yt = xt + D.w;
                const int xa = r[l].cx0, xb = std::min((int)r[l].cx1, D.w) - 1;   // columns whose table entries are read
                int nx0 = xt[xa].x, nx1 = std::min(xt[xb].x + 1, S.w - 1);
                for (int x = xa; x <= xb; x++) { nx0 = std::min(nx0, xt[x].x); nx1 = std::max(nx1, std::min(xt[x].x + 1, S.w - 1)); }
                int ny0 = S.h, ny1 = 0;
                for (int y = r[l].cy0; y < r[l].cy1; y++) {
                    const int s0 = yt[y].x & 0xFFFF, s1 = (int)((unsigned)yt[y].x >> 16);
                    ny0 = std::min(ny0, std::min(s0, s1));
                    ny1 = std::max(ny1, std::max(s0, s1));
                }
                r[l - 1].cx0 = (short)(std::min((int)r[l - 1].ox0, nx0) & ~3);
                r[l - 1].cx1 = (short)up4(std::max((int)r[l - 1].ox1, nx1 + 1));
                r[l - 1].cy0 = (short)std::min((int)r[l - 1].oy0, ny0);
                r[l - 1].cy1 = (short)std::max((int)r[l - 1].oy1, ny1 + 1);
            }
            int off = 0, tab = 0;
            for (int l = 0; l < nl; l++) {
                r[l].pitch = r[l].cx1 - r[l].cx0;
                r[l].soff = off;
                off += (r[l].pitch * (r[l].cy1 - r[l].cy0) + 15) & ~15;
                if (l > 0) tab += r[l].pitch + (r[l].cy1 - r[l].cy0);   // the table slices follow the rectangles
                // the kernel turns a task index into (row, group) with a 32-bit reciprocal and a multiply-high
                const int gq = r[l].pitch >> 2, ch = r[l].cy1 - r[l].cy0;
                if (gq > 1) {
                    const unsigned rcp = 0xFFFFFFFFu / (unsigned)gq + 1u;
                    for (int t = 0; t < gq * ch; t++)
                        if ((int)(((unsigned long long)t * rcp) >> 32) != t / gq) return 0;
                }
            }
            off += tab * (int)sizeof(int2);
            worst = std::max(worst, off);
            if (out) std::copy(r, r + nl, out + (size_t)(j * RX + i) * nl);
        }
    if (worst > 200 * 1024) return 0;
    *smem_bytes = worst;
    return RX * RY;
}

bool launch_pyramid_regions(const Geometry& g, const BatchView& v, cudaStream_t stream) {
    if (!v.pyr_regions || v.n_pyr_regions <= 0 || (((uintptr_t)level_ptr(g, v, 0, 0) | (uintptr_t)level_pitch(g, v, 0) | (uintptr_t)v.l0_stride) & 3)) return false;
    static int configured[64] = {};   // opt-in shared-memory size: a per-device function attribute
    int dev = 0;
    cudaGetDevice(&dev);
    if (v.pyr_regions_smem > configured[dev & 63]) {
        cudaFuncSetAttribute(pyramid_regions_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v.pyr_regions_smem);
        configured[dev & 63] = v.pyr_regions_smem;
    }
    pyramid_regions_kernel<<<dim3(v.n_pyr_regions, v.B), kPrThreads, v.pyr_regions_smem, stream>>>(g, v);
    return true;
}

// ------------------------------------------------------------------------------------------------
// Gaussian 7x7 sigma=2, OpenCV bit-exact fixed point: Q8 kernel {18,34,48,56,48,34,18};
// horizontal pass -> Q8.8 (uint16), vertical pass -> Q16.16, (v + 32768) >> 16. BORDER_REFLECT_101
// on the level itself (the reference blurs a clone of the ROI, so the pyramid border is not seen).
// Both passes are exact integer sums (no intermediate rounding, no saturation: 255*256 < 65536), and both are
// byte / halfword dot products, which is what IDP.4A / IDP.2A compute on the multiply pipe (the integer ALU pipe,
// which bounds FAST, stays free):
//   stage : aligned 16-byte vectors into shared memory (rows reflected by index, the <= 3 reflected columns at the
//           image's left/right edge patched in place)
//   H pass: 4 px x 2 rows per thread. With the three aligned words around the 4 pixels, output x+i is
//           dp4a(wm, Km[i]) + dp4a(w0, K0[i]) + dp4a(wp, Kp[i]) with the taps shifted inside the weight words
//           (10 IDP.4A per 4 px, no byte shuffling). Rows 2m and 2m+1 are stored as one (lo, hi) 16-bit pair.
//   V pass: 4 px x 4 rows per thread: an output row is 4 IDP.2A over four row pairs; even and odd rows use the same
//           pairs with the tap pattern moved by one row.
// One CTA = 64 x 32 output tile of one level of one frame (host-built tile table); all levels in one launch.
// ------------------------------------------------------------------------------------------------
constexpr int kBlurTW = 64, kBlurTH = 32, kBlurThreads = 128;
constexpr int kBlurInWords = 24;                    // 96 staged bytes per row from x = tx0-16: six 16-byte vectors
constexpr int kBlurRows = kBlurTH + 6;              // staged row r <-> y = ty0 - 3 + r
constexpr int kBlurPairRows = kBlurRows / 2;        // 19
constexpr int kBlurVPitch = kBlurTW + 4;            // words per pair row (+4: rows of 16-byte vectors on distinct banks)

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// Taps {18,34,48,56,48,34,18} centred on byte i of w0, laid over the bytes of (wm, w0, wp): weight words per output byte,
// kept in constant memory so that they are instruction operands (as immediates the compiler re-materialises them in
// uniform registers on every trip of the loop).
__constant__ uint32_t c_blur_taps[10] = {0x30221200u, 0x12223038u, 0x22120000u, 0x22303830u, 0x00000012u,
                                         0x12000000u, 0x30383022u, 0x00001222u, 0x38302212u, 0x00122230u};
__constant__ uint32_t c_blur_vtaps[8] = {18u | (34u << 8), 48u | (56u << 8), 48u | (34u << 8), 18u,
                                          18u << 8,         34u | (48u << 8), 56u | (48u << 8), 34u | (18u << 8)};
struct BlurTaps {
    __device__ __forceinline__ uint32_t h0(uint32_t wm, uint32_t w0) const { return __dp4a(wm, c_blur_taps[0], __dp4a(w0, c_blur_taps[1], 0u)); }
    __device__ __forceinline__ uint32_t h1(uint32_t wm, uint32_t w0, uint32_t wp) const {
        return __dp4a(wm, c_blur_taps[2], __dp4a(w0, c_blur_taps[3], __dp4a(wp, c_blur_taps[4], 0u)));
    }
    __device__ __forceinline__ uint32_t h2(uint32_t wm, uint32_t w0, uint32_t wp) const {
        return __dp4a(wm, c_blur_taps[5], __dp4a(w0, c_blur_taps[6], __dp4a(wp, c_blur_taps[7], 0u)));
    }
    __device__ __forceinline__ uint32_t h3(uint32_t w0, uint32_t wp) const { return __dp4a(w0, c_blur_taps[8], __dp4a(wp, c_blur_taps[9], 0u)); }
};

template <bool kTma>
__global__ void __launch_bounds__(kBlurThreads) blur_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v,
                                                            const int4* __restrict__ tiles, const __grid_constant__ TmaMaps maps) {
    __shared__ __align__(128) uint32_t s_in[kBlurRows * kBlurInWords];
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ __align__(16) uint32_t s_v[kBlurPairRows * kBlurVPitch];   // [pair row][x]: H(row 2m) | H(row 2m+1) << 16
    const int frame = blockIdx.y;
    const int4 ti = __ldg(&tiles[blockIdx.x]);   // {level, tx0, ty0, -}
    const int level = ti.x, tx0 = ti.y, ty0 = ti.z;
    const LevelGeom& L = g.lv[level];
    const int tid = threadIdx.x;

    COEB_TRACE(v, 9);
    // stage rows ty0-3 .. ty0+34, 96 bytes from x = tx0-16
    if (kTma) {
        // one TMA box (out-of-image bytes read as 0); the rows above / below the image are then filled by reflection from the
        // staged rows themselves (BORDER_REFLECT_101), before the column patches below see them
        const uint32_t a_mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
        if (tid == 0) tma_issue_box(a_mbar, (uint32_t)__cvta_generic_to_shared(s_in), &maps.m[level], tx0 - 16, ty0 - 3, frame, kBlurRows * kBlurInWords * 4);
        __syncthreads();
        tma_wait(a_mbar);
        if (ty0 < 3 || ty0 + kBlurRows - 3 > L.h) {   // uniform: the tile sees rows outside the image
            // only the three rows just above (y = -3..-1) and just below (y = h..h+2) the image are ever read by an in-image output
            for (int i = tid; i < 6 * kBlurInWords; i += kBlurThreads) {
                const int k = i / kBlurInWords, wv = i - k * kBlurInWords;
                const int y = k < 3 ? k - 3 : L.h + (k - 3);
                const int r = y - ty0 + 3;
                const int yy = y < 0 ? -y : 2 * L.h - 2 - y;
                const int rr = yy - ty0 + 3;
                if (r >= 0 && r < kBlurRows && rr >= 0 && rr < kBlurRows) s_in[r * kBlurInWords + wv] = s_in[rr * kBlurInWords + wv];   // source rows are in-image rows
            }
        }
    } else {   // six 16-byte loads per row, rows reflected by index
        const uint8_t* __restrict__ src = level_ptr(g, v, level, frame);
        const int spitch = level_pitch(g, v, level);
        const int q = tid & 7, r0 = tid >> 3;
        // vectors left of the image or beyond the row pitch (always whole vectors: both are multiples of 16) are clamped: the
        // only bytes of them an in-image output can see are the 3 reflected columns patched below
        const uint8_t* __restrict__ col = src + min(max(tx0 - 16 + 16 * q, 0), spitch - 16);
        if (q < 6) {
#pragma unroll
            for (int ry = r0; ry < kBlurRows; ry += kBlurThreads / 8) {
                int gy = ty0 + ry - 3;
                gy = gy < 0 ? -gy : gy;
                gy = gy >= L.h ? 2 * L.h - 2 - gy : gy;
                gy = min(max(gy, 0), L.h - 1);   // rows far below a short level: never used by an output row
                reinterpret_cast<uint4*>(s_in)[ry * 6 + q] = __ldg(reinterpret_cast<const uint4*>(col + (size_t)gy * spitch));
            }
        }
    }
    __syncthreads();
    constexpr int kRowBytes = kBlurInWords * 4;
    if (tx0 == 0 || tx0 + kBlurTW + 3 > L.w - 1) {   // uniform: tiles touching the left / right image edge
        uint8_t* s_b = reinterpret_cast<uint8_t*>(s_in);   // byte view: column c <-> x = tx0 - 16 + c
        if (tx0 == 0) {   // x = -1,-2,-3  <-  x = 1,2,3
            for (int i = tid; i < kBlurRows * 3; i += kBlurThreads) {
                const int ry = i / 3, k = i - ry * 3 + 1;
                s_b[ry * kRowBytes + 16 - k] = s_b[ry * kRowBytes + 16 + reflect101(k, L.w)];
            }
        }
        if (tx0 + kBlurTW + 3 > L.w - 1) {   // x = w, w+1, w+2  <-  x = w-2, w-3, w-4
            for (int i = tid; i < kBlurRows * 3; i += kBlurThreads) {
                const int ry = i / 3, k = i - ry * 3;
                const int c = L.w + k - tx0 + 16, cs = reflect101(L.w + k, L.w) - tx0 + 16;
                if (c < kRowBytes && cs >= 0) s_b[ry * kRowBytes + c] = s_b[ry * kRowBytes + cs];
            }
        }
        __syncthreads();
    }

    // H pass: (pair row, 4-pixel group) tasks
    const BlurTaps tp;
    for (int i = tid; i < kBlurPairRows * (kBlurTW / 4); i += kBlurThreads) {
        const int pr = i >> 4, gx = i & 15;
        const uint32_t* w = &s_in[(2 * pr) * kBlurInWords + gx + 3];   // w[0]: x-4.., w[1]: the 4 output pixels, w[2]: x+4..
        const uint32_t am = w[0], a0 = w[1], ap = w[2];
        const uint32_t bm = w[kBlurInWords], b0 = w[kBlurInWords + 1], bp = w[kBlurInWords + 2];
        uint4 o;
        o.x = tp.h0(am, a0) | (tp.h0(bm, b0) << 16);
        o.y = tp.h1(am, a0, ap) | (tp.h1(bm, b0, bp) << 16);
        o.z = tp.h2(am, a0, ap) | (tp.h2(bm, b0, bp) << 16);
        o.w = tp.h3(a0, ap) | (tp.h3(b0, bp) << 16);
        *reinterpret_cast<uint4*>(&s_v[pr * kBlurVPitch + 4 * gx]) = o;
    }
    __syncthreads();

    // V pass: 16 column groups x 8 strips of 4 rows; strip s reads pair rows 2s .. 2s+4
    {
        const int gx = tid & 15, strip = tid >> 4;
        const int x = tx0 + 4 * gx;
        if (x < L.w) {
            uint4 p[5];
#pragma unroll
            for (int m = 0; m < 5; m++) p[m] = *reinterpret_cast<const uint4*>(&s_v[(2 * strip + m) * kBlurVPitch + 4 * gx]);
            uint8_t* dst = blur_ptr(g, v, level, frame) + (size_t)(ty0 + 4 * strip) * L.pitch + x;
            // taps over rows r..r+6 as (lo, hi) weights of four row pairs: an even row starts on a pair, an odd row one
            // halfword later
            // (kept in constant memory like the H taps: as immediates they are re-materialised in uniform registers per use)
            const uint32_t E0 = c_blur_vtaps[0], E1 = c_blur_vtaps[1], E2 = c_blur_vtaps[2], E3 = c_blur_vtaps[3];
            const uint32_t O0 = c_blur_vtaps[4], O1 = c_blur_vtaps[5], O2 = c_blur_vtaps[6], O3 = c_blur_vtaps[7];
            const int rows = min(4, L.h - (ty0 + 4 * strip));   // rows of this strip inside the image (may be <= 0)
#pragma unroll
            for (int r = 0; r < 4; r++) {
                if (r < rows) {
                    const int m = r >> 1;
                    const uint32_t k0 = (r & 1) ? O0 : E0, k1 = (r & 1) ? O1 : E1, k2 = (r & 1) ? O2 : E2, k3 = (r & 1) ? O3 : E3;
                    uint32_t acc[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const uint32_t q0 = c == 0 ? p[m].x : c == 1 ? p[m].y : c == 2 ? p[m].z : p[m].w;
                        const uint32_t q1 = c == 0 ? p[m + 1].x : c == 1 ? p[m + 1].y : c == 2 ? p[m + 1].z : p[m + 1].w;
                        const uint32_t q2 = c == 0 ? p[m + 2].x : c == 1 ? p[m + 2].y : c == 2 ? p[m + 2].z : p[m + 2].w;
                        const uint32_t q3 = c == 0 ? p[m + 3].x : c == 1 ? p[m + 3].y : c == 2 ? p[m + 3].z : p[m + 3].w;
                        acc[c] = __dp2a_lo(q0, k0, __dp2a_lo(q1, k1, __dp2a_lo(q2, k2, __dp2a_lo(q3, k3, 32768u))));
                    }
                    // (acc >> 16) is the output byte: gather byte 2 of the four accumulators
                    const uint32_t lo = __byte_perm(acc[0], acc[1], 0x0062), hi = __byte_perm(acc[2], acc[3], 0x0062);
                    *reinterpret_cast<uint32_t*>(dst) = __byte_perm(lo, hi, 0x5410);   // tail lands in row padding
                    dst += L.pitch;
                }
            }
        }
    }
}

int build_blur_tiles(const Geometry& g, int4* out) {
    int total = 0;
    for (int l = 0; l < g.nlevels; l++)
        for (int ty = 0; ty < g.lv[l].h; ty += kBlurTH)
            for (int tx = 0; tx < g.lv[l].w; tx += kBlurTW, total++)
                if (out) out[total] = make_int4(l, tx, ty, 0);
    return total;
}

void launch_blur(const Geometry& g, const BatchView& v, cudaStream_t stream) {
    TmaMaps maps;
    if (tma_enabled() && encode_level_maps(g, v, kBlurInWords * 4, kBlurRows, &maps))
        blur_kernel<true><<<dim3(g.blur_tiles_per_frame, v.B), kBlurThreads, 0, stream>>>(g, v, v.blur_tiles, maps);
    else
        blur_kernel<false><<<dim3(g.blur_tiles_per_frame, v.B), kBlurThreads, 0, stream>>>(g, v, v.blur_tiles, maps);
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
    COEB_TRACE(v, 0);
    const int frame = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (frame >= v.B) return;
    // the FAST stage's counters of this frame start at zero (a memset node costs a small batch 3-4 us of its latency chain)
    for (int i = lane; i < g.nlevels; i += 32) v.lmax_count[frame * g.nlevels + i] = 0;
    for (int i = lane; i < g.cells_per_frame; i += 32) v.cell_count[(size_t)frame * g.cells_per_frame + i] = 0;
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
