// fast.cu -- FAST-9/16 corner strength, cell-aware 3x3 non-max suppression and the per-cell threshold fallback (E4).
//
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree (src/ORBextractor.cc:793-850): one
// cv::FAST(roi, iniThFAST, nms=true) per ~30x30 cell (+6 px apron), repeated with minThFAST when the cell yields
// nothing. The reference makes up to 2 x 815 library calls per 640x480 frame; here one streaming kernel walks
// 64x30 tiles of every level of every frame at iniTh, and a second, almost idle kernel redoes the few cells that
// came out empty at minTh.
//
// Arithmetic (OpenCV FAST_t<16> + cornerScore<16>, pinned against cv2 4.13 in tests/):
//   d_k   = v - ring_k, k = 0..15
//   A(p)  = max( max over the 16 circular 9-arcs of min d_k ,  max over arcs of min (-d_k) )
//   corner(p, th) <=> A(p) > th ;  response = A(p) - 1 for a corner, whatever th was
//   OpenCV's 3x3 NMS keeps p iff response(p) > response(n) for its 8 neighbours, non-corners and pixels outside the
//   cell's detection area [3, dim-3) scoring 0  <=>  A(p) > A(n) over the same-cell neighbours with A(n) > th.
//   A cell uses minTh iff it has no kept pixel at iniTh (:834-838).
//
// Main kernel, per tile:
//   1a  bound: every 9-arc contains two compass points (ring 0,4,8,12) that are neighbours on the compass, so
//       A <= max over k in {0,4,8,12} of min(d_k, d_k+4) (and the mirrored bound for dark arcs). Pixel pairs whose
//       bound is <= iniTh cannot be corners at iniTh and are dropped (about 3 of 4 on textured frames); the others
//       are queued in shared memory, which re-packs the survivors densely over the threads.
//   1b  exact A for the queued pairs, branch-free, with packed 16-bit SIMD (VIMNMX3.S16x2 on sm_100a): the ring bytes
//       come from aligned shared-memory words via funnel shifts, are widened to s16x2 pairs and biased by +256 so a
//       plain 32-bit subtract cannot borrow across lanes; the sliding 9-window min / max is two min3 (max3) stages:
//       m3[k] = min3(d[k], d[k+1], d[k+2]); m9[k] = min3(m3[k], m3[k+3], m3[k+6]).
//   2   cell-aware NMS over A > iniTh, local maxima appended to the level's list, one count per FAST cell.
// Fallback kernel: persistent CTAs scan the cell counters; a cell with count 0 is recomputed alone (plain scalar code,
// it is rare) and its local maxima above minTh are appended to the same list.
#include <type_traits>

#include "coeb_device.cuh"

namespace coeb {

constexpr int kFtW = 64, kFtH = 30;            // output tile
constexpr int kFtThreads = 288;                // 18 four-pixel groups x 32 rows of A = 576 = 2 per thread
constexpr int kImgWords = 28;                  // 24 words (96 px from x = tx0-16, six 16-byte loads) + pad to a 16-byte multiple
constexpr int kImgRows = kFtH + 8;             // 3 (ring) + 1 (NMS halo) each side
constexpr int kAW = 72, kARows = kFtH + 2;     // A tile: x from tx0-4 (18 groups), y from ty0-1
constexpr int kGroups = 18 * kARows;           // 576 four-pixel groups = 1152 pixel pairs

__device__ __forceinline__ uint32_t pair_lo(uint32_t w) { return __byte_perm(w, 0u, 0x4140); }
__device__ __forceinline__ uint32_t pair_hi(uint32_t w) { return __byte_perm(w, 0u, 0x4342); }

// A for the two pixels packed in `c` (centre, s16x2) given their 16 ring pairs r[k] (s16x2), all unbiased bytes.
// kBright / kDark select which arc polarity is evaluated: a pair whose compass bound rules one polarity out (for both
// of its pixels) only needs the other half of the min/max network.
template <bool kBright, bool kDark>
__device__ __forceinline__ uint32_t corner_strength2(uint32_t c, const uint32_t (&r)[16]) {
    const uint32_t cb = c + 0x01000100u;  // +256 per lane: d' = d + 256 in [1, 511], no borrow between lanes
    uint32_t d[16];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = cb - r[k];
    uint32_t ab = 0u, ad = 0u;
    if (kBright) {
        uint32_t mn[16];
#pragma unroll
        for (int k = 0; k < 16; k++) mn[k] = __vimin3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
        uint32_t best_b = 0u;
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
            const uint32_t a0 = __vimin3_s16x2(mn[k], mn[(k + 3) & 15], mn[(k + 6) & 15]);
            const uint32_t a1 = __vimin3_s16x2(mn[k + 1], mn[(k + 4) & 15], mn[(k + 7) & 15]);
            best_b = __vimax3_s16x2(best_b, a0, a1);
        }
        ab = __vsub2(best_b, 0x01000100u);   // best_b - 256
    }
    if (kDark) {
        uint32_t mx[16];
#pragma unroll
        for (int k = 0; k < 16; k++) mx[k] = __vimax3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
        uint32_t best_d = 0x7fff7fffu;
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
            const uint32_t b0 = __vimax3_s16x2(mx[k], mx[(k + 3) & 15], mx[(k + 6) & 15]);
            const uint32_t b1 = __vimax3_s16x2(mx[k + 1], mx[(k + 4) & 15], mx[(k + 7) & 15]);
            best_d = __vimin3_s16x2(best_d, b0, b1);
        }
        ad = __vsub2(0x01000100u, best_d);   // 256 - best_d
    }
    return __vimax3_s16x2(ab, ad, 0u);       // A = max(bright, dark), clamped at 0
}

// Upper bounds of the bright and dark arc strengths of the two pixels packed in `c`, from the four compass ring pairs
// (ring 0, 4, 8, 12). Returns bit 0: some lane's bright bound exceeds th; bit 1: same for dark.
__device__ __forceinline__ int compass_bound2(uint32_t c, uint32_t r0, uint32_t r4, uint32_t r8, uint32_t r12, uint32_t th2) {
    const uint32_t cb = c + 0x01000100u;
    const uint32_t d0 = cb - r0, d4 = cb - r4, d8 = cb - r8, d12 = cb - r12;
    const uint32_t bb = __vimax3_s16x2(__vimax3_s16x2(__vmins2(d0, d4), __vmins2(d4, d8), __vmins2(d8, d12)), __vmins2(d12, d0), 0u);
    const uint32_t bd = __vimin3_s16x2(__vimin3_s16x2(__vmaxs2(d0, d4), __vmaxs2(d4, d8), __vmaxs2(d8, d12)), __vmaxs2(d12, d0), 0x7fff7fffu);
    const uint32_t ub = __vimax3_s16x2(__vsub2(bb, 0x01000100u), th2, 0u), ud = __vimax3_s16x2(__vsub2(0x01000100u, bd), th2, 0u);
    return (int)(ub != th2) | ((int)(ud != th2) << 1);
}

__global__ void __launch_bounds__(kFtThreads, 6) fast_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v,
                                                             const int4* __restrict__ tiles) {
    __shared__ __align__(16) uint32_t s_img[kImgRows * kImgWords];
    __shared__ __align__(4) uint8_t s_A[kARows * kAW];
    __shared__ __align__(8) short s_colcell[kAW];
    __shared__ short s_rowcell[kARows];
    __shared__ uint32_t s_list[kFtW * kFtH / 2];
    __shared__ unsigned short s_queue[3][2 * kGroups];   // pairs needing the bright network, the dark one, both
    __shared__ unsigned short s_cq[kFtW * kFtH];   // interior pixels with A > iniTh, to be checked by the NMS
    __shared__ int s_n, s_base, s_nq[3], s_nc;

    const int frame = blockIdx.y;
    const int4 ti = __ldg(&tiles[blockIdx.x]);   // {level, tx0, ty0, -} built on the host: no per-thread div/mod or level search
    const int level = ti.x, tx0 = ti.y, ty0 = ti.z;
    const LevelGeom& L = g.lv[level];
    const int tid = threadIdx.x;
    const uint8_t* __restrict__ img = level_ptr(g, v, level, frame);
    const int pitch = level_pitch(g, v, level);
    const DynState& dyn = v.dyn[frame];
    const int thIni = dyn.area_flag ? 30 : 20;   // threshold override, src/ORBextractor.cc:775-784

    // ---- stage the tile: rows ty0-4 .. ty0+33, 96 bytes from x = tx0-16 as six 16-byte loads per row (228 loads per
    //      CTA); outside the image -> 0. Bytes between w and the row pitch are padding and never reach an in-domain pixel.
    for (int i = tid; i < kImgRows * 6; i += kFtThreads) {
        const int ry = i / 6, q = i - ry * 6;
        const int gy = ty0 - 4 + ry, gx = tx0 - 16 + 16 * q;
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        if (gy >= 0 && gy < L.h && gx + 16 <= pitch) w = __ldg(reinterpret_cast<const uint4*>(img + (size_t)gy * pitch + gx));
        *reinterpret_cast<uint4*>(&s_img[ry * kImgWords + 4 * q]) = w;
    }
    for (int i = tid; i < kARows * kAW / 4; i += kFtThreads) reinterpret_cast<uint32_t*>(s_A)[i] = 0u;
    // cell index of every column / row of the A tile (minBorder-relative detection coordinates, src/ORBextractor.cc:813-828)
    const int lastJ = max(min(L.nCols - 1, (L.maxBX - 6 - kMinBorder - 1) / L.wCell), 0);
    const int lastI = max(min(L.nRows - 1, (L.maxBY - 3 - kMinBorder - 1) / L.hCell), 0);
    if (tid < kAW) {
        const int x = tx0 - 4 + tid;   // level coordinate
        s_colcell[tid] = (x >= kEdge && x < L.w - kEdge) ? (short)min((x - kEdge) / L.wCell, lastJ) : (short)-1;
    } else if (tid >= 96 && tid < 96 + kARows) {
        const int y = ty0 - 1 + (tid - 96);
        s_rowcell[tid - 96] = (y >= kEdge && y < L.h - kEdge) ? (short)min((y - kEdge) / L.hCell, lastI) : (short)-1;
    }
    if (tid == 0) { s_n = 0; s_nq[0] = s_nq[1] = s_nq[2] = 0; s_nc = 0; }
    __syncthreads();

    // ---- 1a: compass bound for every pixel pair of the tile + 1 px halo; survivors go to the queue ----
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
        const int grp = tid + it * kFtThreads;          // 0..575
        const int ay = grp / 18, gxi = grp - ay * 18;   // A row (y = ty0-1+ay), group (x = tx0-4+4*gxi)
        const uint32_t* row = &s_img[(ay + 3) * kImgWords + gxi + 3];   // word holding the 4 centre pixels
        const uint32_t c = row[0];
        const uint32_t S0 = row[3 * kImgWords], S8 = row[-3 * kImgWords];
        const uint32_t S4 = __funnelshift_r(c, row[1], 24), S12 = __funnelshift_r(row[-1], c, 8);
        const uint32_t th2 = (uint32_t)thIni * 0x00010001u;
        const int f01 = compass_bound2(pair_lo(c), pair_lo(S0), pair_lo(S4), pair_lo(S8), pair_lo(S12), th2);
        const int f23 = compass_bound2(pair_hi(c), pair_hi(S0), pair_hi(S4), pair_hi(S8), pair_hi(S12), th2);
        if (f01) s_queue[f01 - 1][atomicAdd(&s_nq[f01 - 1], 1)] = (unsigned short)(2 * grp);
        if (f23) s_queue[f23 - 1][atomicAdd(&s_nq[f23 - 1], 1)] = (unsigned short)(2 * grp + 1);
    }
    __syncthreads();

    // ---- 1b: exact corner strength of the queued pairs, one dense pass per class (no divergence inside a pass) ----
    auto exact_pass = [&](auto bright, auto dark, const unsigned short* queue, int nq) {
#pragma unroll 1
        for (int qi = tid; qi < nq; qi += kFtThreads) {
            const int e = queue[qi];
            const int grp = e >> 1, half = e & 1;
            const int ay = grp / 18, gxi = grp - ay * 18;
            const uint32_t* row = &s_img[(ay + 3) * kImgWords + gxi + 3];
            uint32_t S[16];
            {   // ring words: 4 consecutive bytes starting at x+dx on row y+dy (FAST circle, OpenCV order)
                const uint32_t *r3 = row + 3 * kImgWords, *rm3 = row - 3 * kImgWords, *r2 = row + 2 * kImgWords, *rm2 = row - 2 * kImgWords,
                               *r1 = row + kImgWords, *rm1 = row - kImgWords;
                S[0] = r3[0];                                   // ( 0, 3)
                S[1] = __funnelshift_r(r3[0], r3[1], 8);        // ( 1, 3)
                S[2] = __funnelshift_r(r2[0], r2[1], 16);       // ( 2, 2)
                S[3] = __funnelshift_r(r1[0], r1[1], 24);       // ( 3, 1)
                S[4] = __funnelshift_r(row[0], row[1], 24);     // ( 3, 0)
                S[5] = __funnelshift_r(rm1[0], rm1[1], 24);     // ( 3,-1)
                S[6] = __funnelshift_r(rm2[0], rm2[1], 16);     // ( 2,-2)
                S[7] = __funnelshift_r(rm3[0], rm3[1], 8);      // ( 1,-3)
                S[8] = rm3[0];                                  // ( 0,-3)
                S[9] = __funnelshift_r(rm3[-1], rm3[0], 24);    // (-1,-3)
                S[10] = __funnelshift_r(rm2[-1], rm2[0], 16);   // (-2,-2)
                S[11] = __funnelshift_r(rm1[-1], rm1[0], 8);    // (-3,-1)
                S[12] = __funnelshift_r(row[-1], row[0], 8);    // (-3, 0)
                S[13] = __funnelshift_r(r1[-1], r1[0], 8);      // (-3, 1)
                S[14] = __funnelshift_r(r2[-1], r2[0], 16);     // (-2, 2)
                S[15] = __funnelshift_r(r3[-1], r3[0], 24);     // (-1, 3)
            }
            const uint32_t sel = half ? 0x4342u : 0x4140u;
            uint32_t r[16];
#pragma unroll
            for (int k = 0; k < 16; k++) r[k] = __byte_perm(S[k], 0u, sel);
            const uint32_t a2 = corner_strength2<decltype(bright)::value, decltype(dark)::value>(__byte_perm(row[0], 0u, sel), r);
            // blank pixels outside the detection domain of the level, store the two bytes (one strength per 16-bit lane)
            const int rc = s_rowcell[ay];
            const int c0 = s_colcell[4 * gxi + 2 * half], c1 = s_colcell[4 * gxi + 2 * half + 1];
            const uint32_t lo = (rc >= 0 && c0 >= 0) ? (a2 & 0xFFu) : 0u, hi = (rc >= 0 && c1 >= 0) ? ((a2 >> 16) & 0xFFu) : 0u;
            *reinterpret_cast<unsigned short*>(&s_A[ay * kAW + 4 * gxi + 2 * half]) = (unsigned short)(lo | (hi << 8));
            // tile-interior pixels above iniTh are the NMS candidates (few per tile): queue them for a dense second pass
            const int py = ay - 1, px = 4 * gxi - 4 + 2 * half;
            if (py >= 0 && py < kFtH) {
                const bool q0 = (int)lo > thIni && px >= 0 && px < kFtW, q1 = (int)hi > thIni && px + 1 >= 0 && px + 1 < kFtW;
                if (q0 || q1) {
                    const int pos = atomicAdd(&s_nc, (int)q0 + (int)q1);
                    if (q0) s_cq[pos] = (unsigned short)(py * kFtW + px);
                    if (q1) s_cq[pos + (int)q0] = (unsigned short)(py * kFtW + px + 1);
                }
            }
        }
    };
    exact_pass(std::true_type{}, std::false_type{}, s_queue[0], s_nq[0]);
    exact_pass(std::false_type{}, std::true_type{}, s_queue[1], s_nq[1]);
    exact_pass(std::true_type{}, std::true_type{}, s_queue[2], s_nq[2]);
    __syncthreads();

    // ---- 2: cell-aware NMS over the tile interior at iniTh; local maxima are emitted, each FAST cell counts its own ----
    int* cellcnt = v.cell_count + (size_t)frame * g.cells_per_frame + L.cell_base;
    const int nc = s_nc;
    for (int i = tid; i < nc; i += kFtThreads) {
        const int e = s_cq[i];
        const int py = e >> 6, px = e & 63;
        const uint8_t* a = &s_A[(py + 1) * kAW + px + 4];
        const int A = a[0];
        const int n_l = a[-1], n_r = a[1], n_ul = a[-kAW - 1], n_u = a[-kAW], n_ur = a[-kAW + 1], n_dl = a[kAW - 1], n_d = a[kAW], n_dr = a[kAW + 1];
        bool keep = A > max(max(max(n_l, n_r), max(n_ul, n_u)), max(max(n_ur, n_dl), max(n_d, n_dr)));
        const int cj = s_colcell[px + 4], ci = s_rowcell[py + 1];
        if (!keep) {
            // a larger neighbour only counts if it belongs to the same cell (each cell is an independent cv::FAST call)
            const bool sl = s_colcell[px + 3] == cj, sr = s_colcell[px + 5] == cj, su = s_rowcell[py] == ci, sd = s_rowcell[py + 2] == ci;
            if (!(sl && sr && su && sd)) {
                int m = 0;
                if (sl) m = max(m, n_l);
                if (sr) m = max(m, n_r);
                if (su) { m = max(m, n_u); if (sl) m = max(m, n_ul); if (sr) m = max(m, n_ur); }
                if (sd) { m = max(m, n_d); if (sl) m = max(m, n_dl); if (sr) m = max(m, n_dr); }
                keep = A > m;
            }
        }
        if (keep) {
            const int x = tx0 + px - kMinBorder, y = ty0 + py - kMinBorder;   // minBorder-relative (:844-845)
            s_list[atomicAdd(&s_n, 1)] = (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)(A - 1) << 24);
            atomicAdd(&cellcnt[ci * L.nCols + cj], 1);
        }
    }
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;
    if (tid == 0) s_base = atomicAdd(v.lmax_count + frame * g.nlevels + level, n);
    __syncthreads();
    uint32_t* out = v.lmax + (size_t)frame * g.cand_per_frame + L.cand_base + s_base;
    const int room = L.cand_cap - s_base;
    for (int i = tid; i < n && i < room; i += kFtThreads) out[i] = s_list[i];
}

// ------------------------------------------------------------------------------------------------------------------
// minTh fallback (src/ORBextractor.cc:834-838): cells without a keypoint at iniTh are redone at minTh. Persistent CTAs
// scan the per-cell counters written by fast_kernel; an empty cell is processed alone, exactly like one cv::FAST call
// on its ROI: scalar sliding-window A, NMS inside the cell, local maxima above minTh appended to the level's list.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kMaxRoi = 72;   // ROI side bound: wCell + 6 <= 66 (checked on the host)

// Lists the FAST cells that exist (src/ORBextractor.cc:816-826) and produced nothing at iniTh. One thread per cell.
__global__ void __launch_bounds__(256) fast_empty_cells_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = v.B * g.cells_per_frame;
    bool empty = false;
    if (c < total && v.cell_count[c] == 0) {
        const int cf = c % g.cells_per_frame;
        int level = 0;
        while (level + 1 < g.nlevels && cf >= g.lv[level + 1].cell_base) level++;
        const LevelGeom& L = g.lv[level];
        const int cell = cf - L.cell_base;
        const int ci = cell / L.nCols, cj = cell - ci * L.nCols;
        const int iniX = kMinBorder + cj * L.wCell, iniY = kMinBorder + ci * L.hCell;
        if (!(iniY >= L.maxBY - 3 || iniX >= L.maxBX - 6)) {
            const int rw = min(iniX + L.wCell + 6, L.maxBX) - iniX, rh = min(iniY + L.hCell + 6, L.maxBY) - iniY;
            empty = rw >= 7 && rh >= 7;
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, empty);
    if (m) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == 0) base = atomicAdd(v.empty_count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (empty) v.empty_cells[base + __popc(m & ((1u << lane) - 1))] = c;
    }
}

__global__ void __launch_bounds__(128) fast_fallback_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v) {
    __shared__ uint8_t s_img[kMaxRoi * kMaxRoi];
    __shared__ uint8_t s_A[(kMaxRoi - 4) * (kMaxRoi - 4)];
    __shared__ uint32_t s_list[(kMaxRoi - 6) * (kMaxRoi - 6) / 2];
    __shared__ int s_n, s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_empty = *v.empty_count;
    for (int e = blockIdx.x; e < n_empty; e += gridDim.x) {
        const int c = v.empty_cells[e];
        const int frame = c / g.cells_per_frame, cf = c - frame * g.cells_per_frame;
        int level = 0;
        while (level + 1 < g.nlevels && cf >= g.lv[level + 1].cell_base) level++;
        const LevelGeom& L = g.lv[level];
        const int cell = cf - L.cell_base;
        const int ci = cell / L.nCols, cj = cell - ci * L.nCols;
        // cell ROI (src/ORBextractor.cc:813-828), level coordinates
        const int iniX = kMinBorder + cj * L.wCell, iniY = kMinBorder + ci * L.hCell;
        const int maxX = min(iniX + L.wCell + 6, L.maxBX), maxY = min(iniY + L.hCell + 6, L.maxBY);
        const int rw = maxX - iniX, rh = maxY - iniY;
        const int dw = rw - 6, dh = rh - 6;          // detection area: ROI rows/cols [3, dim-3)
        const int aw = dw + 2;                       // s_A row pitch (1 px zero border each side)
        const int thMin = v.dyn[frame].area_flag ? 10 : 7;
        const int pitch = level_pitch(g, v, level);
        const uint8_t* __restrict__ img = level_ptr(g, v, level, frame) + (size_t)iniY * pitch + iniX;
        __syncthreads();   // previous cell's shared data fully consumed
        for (int y = warp; y < rh; y += 4)
            for (int x = lane; x < rw; x += 32) s_img[y * kMaxRoi + x] = __ldg(img + (size_t)y * pitch + x);
        for (int i = tid; i < aw * (dh + 2); i += 128) s_A[i] = 0;
        if (tid == 0) s_n = 0;
        __syncthreads();
        const int off[16] = {3 * kMaxRoi,      3 * kMaxRoi + 1,  2 * kMaxRoi + 2,  kMaxRoi + 3,      3,                -kMaxRoi + 3,
                             -2 * kMaxRoi + 2, -3 * kMaxRoi + 1, -3 * kMaxRoi,     -3 * kMaxRoi - 1, -2 * kMaxRoi - 2, -kMaxRoi - 3,
                             -3,               kMaxRoi - 3,      2 * kMaxRoi - 2,  3 * kMaxRoi - 1};
        for (int y = warp; y < dh; y += 4)
            for (int x = lane; x < dw; x += 32) {
                const uint8_t* p = &s_img[(y + 3) * kMaxRoi + (x + 3)];
                const int cc = p[0];
                {   // compass bound (see fast_kernel 1a): in flat cells almost every pixel stops here
                    const int e0 = cc - (int)p[off[0]], e4 = cc - (int)p[off[4]], e8 = cc - (int)p[off[8]], e12 = cc - (int)p[off[12]];
                    const int ub = max(max(min(e0, e4), min(e4, e8)), max(min(e8, e12), min(e12, e0)));
                    const int ud = min(min(max(e0, e4), max(e4, e8)), min(max(e8, e12), max(e12, e0)));
                    if (ub <= thMin && -ud <= thMin) continue;   // s_A stays 0
                }
                int d[16];
#pragma unroll
                for (int k = 0; k < 16; k++) d[k] = cc - (int)p[off[k]];
                int mn[16], mx[16];
#pragma unroll
                for (int k = 0; k < 16; k++) { mn[k] = min(d[k], min(d[(k + 1) & 15], d[(k + 2) & 15])); mx[k] = max(d[k], max(d[(k + 1) & 15], d[(k + 2) & 15])); }
                int best_b = -256, best_d = 256;
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    best_b = max(best_b, min(mn[k], min(mn[(k + 3) & 15], mn[(k + 6) & 15])));
                    best_d = min(best_d, max(mx[k], max(mx[(k + 3) & 15], mx[(k + 6) & 15])));
                }
                s_A[(y + 1) * aw + (x + 1)] = (uint8_t)max(max(best_b, -best_d), 0);
            }
        __syncthreads();
        for (int y = warp; y < dh; y += 4)
            for (int x = lane; x < dw; x += 32) {
                const uint8_t* a = &s_A[(y + 1) * aw + (x + 1)];
                const int A = a[0];
                if (A > thMin) {
                    const int nb = max(max(max(a[-1], a[1]), max(a[-aw - 1], a[-aw])), max(max(a[-aw + 1], a[aw - 1]), max(a[aw], a[aw + 1])));
                    if (A > nb) {
                        const int px = x + 3 + cj * L.wCell, py = y + 3 + ci * L.hCell;  // minBorder-relative (:844-845)
                        s_list[atomicAdd(&s_n, 1)] = (uint32_t)px | ((uint32_t)py << 12) | ((uint32_t)(A - 1) << 24);
                    }
                }
            }
        __syncthreads();
        const int n = s_n;
        if (n > 0) {
            if (tid == 0) s_base = atomicAdd(v.lmax_count + frame * g.nlevels + level, n);
            __syncthreads();
            uint32_t* out = v.lmax + (size_t)frame * g.cand_per_frame + L.cand_base + s_base;
            const int room = L.cand_cap - s_base;
            for (int i = tid; i < n && i < room; i += 128) out[i] = s_list[i];
        }
    }
}

// Tiles cover the detection domain x in [19, w-19), y in [19, h-19) of every level, starting at the 4-aligned x = 16.
int build_fast_tiles(const Geometry& g, int4* out) {
    int total = 0;
    for (int l = 0; l < g.nlevels; l++) {
        const int tiles_x = std::max(1, (g.lv[l].w - kEdge - kMinBorder + kFtW - 1) / kFtW);
        const int tiles_y = std::max(1, (g.lv[l].h - kEdge - kMinBorder + kFtH - 1) / kFtH);
        for (int ty = 0; ty < tiles_y; ty++)
            for (int tx = 0; tx < tiles_x; tx++, total++)
                if (out) out[total] = make_int4(l, kMinBorder + tx * kFtW, kMinBorder + ty * kFtH, 0);
    }
    return total;
}

void launch_fast(const Geometry& g, const BatchView& v, cudaStream_t stream) {
    cudaMemsetAsync(v.lmax_count, 0, sizeof(int) * (size_t)v.B * g.nlevels, stream);
    cudaMemsetAsync(v.cell_count, 0, sizeof(int) * (size_t)v.B * g.cells_per_frame, stream);
    fast_kernel<<<dim3(g.fast_tiles_per_frame, v.B), kFtThreads, 0, stream>>>(g, v, v.fast_tiles);
    const int cells = v.B * g.cells_per_frame;
    cudaMemsetAsync(v.empty_count, 0, sizeof(int), stream);
    fast_empty_cells_kernel<<<(cells + 255) / 256, 256, 0, stream>>>(g, v);
    fast_fallback_kernel<<<std::min(cells, 148 * 8), 128, 0, stream>>>(g, v);
}

}  // namespace coeb
