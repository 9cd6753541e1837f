// fast.cu -- per-cell FAST-9/16 detection with threshold fallback (E4) and the pre-octree cull (E5).
//
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree (src/ORBextractor.cc:793-850): one
// cv::FAST(roi, iniThFAST, nms=true) per ~30x30 cell (+6 px apron), repeated with minThFAST when the
// cell yields nothing. One CTA per (cell, frame); the ROI is staged in shared memory.
//
// Arithmetic notes (OpenCV FAST_t<16> + cornerScore<16>, pinned in tests against cv2 4.13):
//   A(p)  = max over the 16 circular 9-arcs of min_k (v - ring_k)  (centre brighter than the arc), and
//           of min_k (ring_k - v) (centre darker)   -- the largest threshold t for which p is a corner is A-1
//   corner(p, th) <=> A(p) > th ;  response = A(p) - 1  (independent of th for a corner)
//   3x3 NMS keeps p iff response(p) > response(n) for all 8 neighbours, where non-corners and
//   pixels outside [3,dim-3) of the ROI score 0  <=>  A(p) > A(n) for all neighbours that are corners
//   at the lowest threshold. So NMS is threshold-independent and one pass serves both thresholds:
//   kept(p, th) = localmax(p) && A(p) > th. The cell falls back to minTh iff no pixel is kept at iniTh.
#include "coeb_device.cuh"

namespace coeb {

constexpr int kMaxRoi = 72;  // ROI side bound: wCell + 6 <= 61 + 6 (nCols >= 1), rounded up

__device__ __forceinline__ bool arc9(uint32_t m) {  // 9 contiguous set bits in a circular 16-bit mask
    m |= m << 16;
    uint32_t a = m & (m >> 1);      // 2
    a &= a >> 2;                    // 4
    a &= a >> 4;                    // 8
    a &= m >> 8;                    // 9
    return (a & 0xFFFFu) != 0;
}

__global__ void __launch_bounds__(256) fast_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v) {
    __shared__ uint8_t s_img[kMaxRoi * kMaxRoi];
    __shared__ uint8_t s_A[(kMaxRoi - 4) * (kMaxRoi - 4)];  // detection area + 1 px zero border
    __shared__ int s_cnt[2];
    __shared__ int s_base;

    const int frame = blockIdx.y;
    int level = 0;
    while (level + 1 < g.nlevels && (int)blockIdx.x >= g.lv[level + 1].cell_base) level++;
    const LevelGeom& L = g.lv[level];
    const int cell = blockIdx.x - L.cell_base;
    const int ci = cell / L.nCols, cj = cell - ci * L.nCols;
    // cell ROI (src/ORBextractor.cc:813-828), level coordinates
    const int iniX = kMinBorder + cj * L.wCell, iniY = kMinBorder + ci * L.hCell;
    if (iniY >= L.maxBY - 3 || iniX >= L.maxBX - 6) return;
    const int maxX = min(iniX + L.wCell + 6, L.maxBX), maxY = min(iniY + L.hCell + 6, L.maxBY);
    const int rw = maxX - iniX, rh = maxY - iniY;
    if (rw < 7 || rh < 7) return;
    const int dw = rw - 6, dh = rh - 6;          // detection area: ROI rows/cols [3, dim-3)
    const int aw = dw + 2;                       // s_A row pitch (1 px zero border each side)
    const int tid = threadIdx.x;

    const uint8_t* __restrict__ img = level_ptr(g, v, level, frame) + (size_t)iniY * level_pitch(g, v, level) + iniX;
    const int pitch = level_pitch(g, v, level);
    for (int i = tid; i < rw * rh; i += 256) {
        const int y = i / rw, x = i - y * rw;
        s_img[y * kMaxRoi + x] = __ldg(img + (size_t)y * pitch + x);
    }
    for (int i = tid; i < aw * (dh + 2); i += 256) s_A[i] = 0;
    if (tid < 2) s_cnt[tid] = 0;
    __syncthreads();

    const DynState& dyn = v.dyn[frame];
    const int thIni = dyn.area_flag ? 30 : 20;   // threshold override, src/ORBextractor.cc:775-784
    const int thMin = dyn.area_flag ? 10 : 7;

    // ring offsets in the staged ROI
    const int off[16] = {3 * kMaxRoi,      3 * kMaxRoi + 1,  2 * kMaxRoi + 2,  kMaxRoi + 3,  3,           -kMaxRoi + 3,
                         -2 * kMaxRoi + 2, -3 * kMaxRoi + 1, -3 * kMaxRoi,     -3 * kMaxRoi - 1, -2 * kMaxRoi - 2, -kMaxRoi - 3,
                         -3,               kMaxRoi - 3,      2 * kMaxRoi - 2,  3 * kMaxRoi - 1};
    for (int i = tid; i < dw * dh; i += 256) {
        const int y = i / dw, x = i - y * dw;
        const uint8_t* p = &s_img[(y + 3) * kMaxRoi + (x + 3)];
        const int c = p[0];
        int d[16];
        uint32_t mb = 0, md = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            d[k] = c - (int)p[off[k]];
            mb |= (uint32_t)(d[k] > thMin) << k;     // centre brighter than ring pixel
            md |= (uint32_t)(d[k] < -thMin) << k;    // centre darker
        }
        int A = 0;
        if (arc9(mb) || arc9(md)) {
            // sliding minimum / maximum over 9 circular neighbours by doubling (2,4,8,+1)
            int mn[16], mx[16];
#pragma unroll
            for (int k = 0; k < 16; k++) { mn[k] = min(d[k], d[(k + 1) & 15]); mx[k] = max(d[k], d[(k + 1) & 15]); }
            int mn4[16], mx4[16];
#pragma unroll
            for (int k = 0; k < 16; k++) { mn4[k] = min(mn[k], mn[(k + 2) & 15]); mx4[k] = max(mx[k], mx[(k + 2) & 15]); }
            int best_b = -256, best_d = 256;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int m8 = min(mn4[k], mn4[(k + 4) & 15]);
                const int M8 = max(mx4[k], mx4[(k + 4) & 15]);
                best_b = max(best_b, min(m8, d[(k + 8) & 15]));
                best_d = min(best_d, max(M8, d[(k + 8) & 15]));
            }
            A = max(best_b, -best_d);
        }
        s_A[(y + 1) * aw + (x + 1)] = (uint8_t)A;
    }
    __syncthreads();

    // NMS + per-threshold counts. Each thread keeps its (few) local maxima in a small register list.
    int cnt_ini = 0, cnt_min = 0;
    for (int i = tid; i < dw * dh; i += 256) {
        const int y = i / dw, x = i - y * dw;
        const uint8_t* a = &s_A[(y + 1) * aw + (x + 1)];
        const int A = a[0];
        if (A > thMin) {
            const int nb = max(max(max(a[-1], a[1]), max(a[-aw - 1], a[-aw])), max(max(a[-aw + 1], a[aw - 1]), max(a[aw], a[aw + 1])));
            if (A > nb) {
                cnt_min++;
                cnt_ini += A > thIni;
            }
        }
    }
    if (cnt_min) atomicAdd(&s_cnt[1], cnt_min);
    if (cnt_ini) atomicAdd(&s_cnt[0], cnt_ini);
    __syncthreads();
    const int th = s_cnt[0] > 0 ? thIni : thMin;
    const int total = s_cnt[0] > 0 ? s_cnt[0] : s_cnt[1];
    if (total == 0) return;

    // Emit. Candidates are appended to the (frame, level) list with one global atomic per CTA; the
    // list order is not the reference's vector order, which only matters for response ties inside an
    // octree node -- the select kernel re-derives that order from (x, y), see octree.cu.
    __syncthreads();
    if (tid == 0) s_cnt[0] = 0;
    __syncthreads();
    uint32_t* out = v.cand + (size_t)frame * g.cand_per_frame + L.cand_base;
    int* gcount = v.cand_count + frame * g.nlevels + level;
    // pass 1: count survivors of the optional cull to reserve space
    int mine = 0;
    for (int i = tid; i < dw * dh; i += 256) {
        const int y = i / dw, x = i - y * dw;
        const uint8_t* a = &s_A[(y + 1) * aw + (x + 1)];
        const int A = a[0];
        if (A > th) {
            const int nb = max(max(max(a[-1], a[1]), max(a[-aw - 1], a[-aw])), max(max(a[-aw + 1], a[aw - 1]), max(a[aw], a[aw + 1])));
            if (A > nb) {
                const int px = x + 3 + cj * L.wCell, py = y + 3 + ci * L.hCell;  // minBorder-relative (:844-845)
                // CheckMovingKeyPoints before the octree, only on the area_flag path (:854-858)
                if (dyn.area_flag && is_moving(dyn, (float)px, (float)py, level, L.scale, g.w0, g.h0)) continue;
                mine++;
            }
        }
    }
    const int my_off = mine ? atomicAdd(&s_cnt[0], mine) : 0;
    __syncthreads();
    if (tid == 0) s_base = s_cnt[0] ? atomicAdd(gcount, s_cnt[0]) : 0;
    __syncthreads();
    if (!mine) return;
    int w = s_base + my_off;
    for (int i = tid; i < dw * dh; i += 256) {
        const int y = i / dw, x = i - y * dw;
        const uint8_t* a = &s_A[(y + 1) * aw + (x + 1)];
        const int A = a[0];
        if (A > th) {
            const int nb = max(max(max(a[-1], a[1]), max(a[-aw - 1], a[-aw])), max(max(a[-aw + 1], a[aw - 1]), max(a[aw], a[aw + 1])));
            if (A > nb) {
                const int px = x + 3 + cj * L.wCell, py = y + 3 + ci * L.hCell;
                if (dyn.area_flag && is_moving(dyn, (float)px, (float)py, level, L.scale, g.w0, g.h0)) continue;
                if (w < L.cand_cap) out[w] = (uint32_t)px | ((uint32_t)py << 12) | ((uint32_t)(A - 1) << 24);
                w++;
            }
        }
    }
}

__global__ void zero_counts_kernel(int* a, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = 0;
}

void launch_fast(const Geometry& g, const BatchView& v, cudaStream_t stream) {
    const int n = v.B * g.nlevels;
    zero_counts_kernel<<<(n + 255) / 256, 256, 0, stream>>>(v.cand_count, n);
    fast_kernel<<<dim3(g.cells_per_frame, v.B), 256, 0, stream>>>(g, v);
}

}  // namespace coeb
