// fast.cu -- FAST-9/16 corner strength, cell-aware 3x3 non-max suppression and the per-cell threshold vote (E4).
//
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree (src/ORBextractor.cc:793-850): one
// cv::FAST(roi, iniThFAST, nms=true) per ~30x30 cell (+6 px apron), repeated with minThFAST when the cell yields
// nothing. The reference makes up to 2 x 815 library calls per 640x480 frame; here ONE streaming kernel walks
// 64x30 tiles of every level of every frame.
//
// Arithmetic (OpenCV FAST_t<16> + cornerScore<16>, pinned against cv2 4.13 in tests/):
//   d_k   = v - ring_k, k = 0..15
//   A(p)  = max( max over the 16 circular 9-arcs of min d_k ,  max over arcs of min (-d_k) )
//   corner(p, th) <=> A(p) > th ;  response = A(p) - 1 for a corner, whatever th was
//   OpenCV's 3x3 NMS keeps p iff response(p) > response(n) for its 8 neighbours, non-corners and pixels outside the
//   cell's detection area [3, dim-3) scoring 0  <=>  A(p) > A(n) over the neighbours that lie in the SAME cell.
//   So NMS does not depend on the threshold: kept(p, th) = localmax(p) && A(p) > th, and the cell uses minTh iff it
//   has no local maximum above iniTh.
// The kernel therefore emits every cell-local maximum with A > minTh once, and counts per cell those above iniTh;
// the consumer (octree.cu, gather phase) applies the cell's threshold. No per-cell kernel, no second FAST pass.
//
// A is computed branch-free for 4 pixels per thread with packed 16-bit SIMD (VIMNMX3.S16x2 on sm_100a): the 16 ring
// bytes of 4 neighbouring pixels come from 21 aligned shared-memory words via funnel shifts, are widened to s16x2
// pairs, biased by +256 so a plain 32-bit subtract cannot borrow across lanes, and the sliding 9-window min / max is
// two min3 (max3) stages: m3[k] = min3(d[k], d[k+1], d[k+2]); m9[k] = min3(m3[k], m3[k+3], m3[k+6]).
#include "coeb_device.cuh"

namespace coeb {

constexpr int kFtW = 64, kFtH = 30;            // output tile
constexpr int kFtThreads = 288;                // 18 four-pixel groups x 32 rows of A = 576 = 2 per thread
constexpr int kImgWords = 28;                  // 24 words (96 px from x = tx0-16, six 16-byte loads) + pad to a 16-byte multiple
constexpr int kImgRows = kFtH + 8;             // 3 (ring) + 1 (NMS halo) each side
constexpr int kAW = 72, kARows = kFtH + 2;     // A tile: x from tx0-4 (18 groups), y from ty0-1

__device__ __forceinline__ uint32_t pair_lo(uint32_t w) { return __byte_perm(w, 0u, 0x4140); }
__device__ __forceinline__ uint32_t pair_hi(uint32_t w) { return __byte_perm(w, 0u, 0x4342); }

// A for the two pixels packed in `c` (centre, s16x2) given their 16 ring pairs r[k] (s16x2), all unbiased bytes.
__device__ __forceinline__ uint32_t corner_strength2(uint32_t c, const uint32_t (&r)[16]) {
    const uint32_t cb = c + 0x01000100u;  // +256 per lane: d' = d + 256 in [1, 511], no borrow between lanes
    uint32_t d[16];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = cb - r[k];
    uint32_t mn[16], mx[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn[k] = __vimin3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
        mx[k] = __vimax3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    }
    uint32_t best_b = 0u, best_d = 0x7fff7fffu;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const uint32_t a0 = __vimin3_s16x2(mn[k], mn[(k + 3) & 15], mn[(k + 6) & 15]);
        const uint32_t a1 = __vimin3_s16x2(mn[k + 1], mn[(k + 4) & 15], mn[(k + 7) & 15]);
        best_b = __vimax3_s16x2(best_b, a0, a1);
        const uint32_t b0 = __vimax3_s16x2(mx[k], mx[(k + 3) & 15], mx[(k + 6) & 15]);
        const uint32_t b1 = __vimax3_s16x2(mx[k + 1], mx[(k + 4) & 15], mx[(k + 7) & 15]);
        best_d = __vimin3_s16x2(best_d, b0, b1);
    }
    // A = max(best_b - 256, 256 - best_d), clamped at 0
    const uint32_t ab = __vsub2(best_b, 0x01000100u);
    const uint32_t ad = __vsub2(0x01000100u, best_d);
    return __vimax3_s16x2(ab, ad, 0u);
}

__global__ void __launch_bounds__(kFtThreads, 4) fast_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v,
                                                          const __grid_constant__ TileMap tm) {
    __shared__ __align__(16) uint32_t s_img[kImgRows * kImgWords];
    __shared__ __align__(4) uint8_t s_A[kARows * kAW];
    __shared__ __align__(8) short s_colcell[kAW];
    __shared__ short s_rowcell[kARows];
    __shared__ uint32_t s_list[kFtW * kFtH / 2];
    __shared__ int s_n, s_base;

    const int frame = blockIdx.y;
    int level = 0;
    while (level + 1 < g.nlevels && (int)blockIdx.x >= tm.tile_base[level + 1]) level++;
    const LevelGeom& L = g.lv[level];
    const int t = blockIdx.x - tm.tile_base[level];
    const int tx0 = kMinBorder + (t % tm.tiles_x[level]) * kFtW, ty0 = kMinBorder + (t / tm.tiles_x[level]) * kFtH;
    const int tid = threadIdx.x;
    const uint8_t* __restrict__ img = level_ptr(g, v, level, frame);
    const int pitch = level_pitch(g, v, level);

    // ---- stage the tile: rows ty0-4 .. ty0+33, 96 bytes from x = tx0-16 as six 16-byte loads per row (228 loads per
    //      CTA); outside the image -> 0. Bytes between w and the row pitch are padding and never reach an in-domain pixel.
    for (int i = tid; i < kImgRows * 6; i += kFtThreads) {
        const int ry = i / 6, q = i - ry * 6;
        const int gy = ty0 - 4 + ry, gx = tx0 - 16 + 16 * q;
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        if (gy >= 0 && gy < L.h && gx + 16 <= pitch) w = __ldg(reinterpret_cast<const uint4*>(img + (size_t)gy * pitch + gx));
        *reinterpret_cast<uint4*>(&s_img[ry * kImgWords + 4 * q]) = w;
    }
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
    if (tid == 0) s_n = 0;
    __syncthreads();

    // ---- corner strength A for the tile + 1 px halo (rounded to 4-px groups) ----
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
        const int grp = tid + it * kFtThreads;          // 0..575
        const int ay = grp / 18, gxi = grp - ay * 18;   // A row (y = ty0-1+ay), group (x = tx0-4+4*gxi)
        const uint32_t* row = &s_img[(ay + 3) * kImgWords + gxi + 3];   // word holding the 4 centre pixels
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
        const uint32_t c = row[0];
        uint32_t r[16];
#pragma unroll
        for (int k = 0; k < 16; k++) r[k] = pair_lo(S[k]);
        const uint32_t a01 = corner_strength2(pair_lo(c), r);
#pragma unroll
        for (int k = 0; k < 16; k++) r[k] = pair_hi(S[k]);
        const uint32_t a23 = corner_strength2(pair_hi(c), r);
        // pack the four strengths (each 0..255) and blank pixels outside the detection domain of the level
        uint32_t a4 = __byte_perm(a01, a23, 0x6420);
        const int rc = s_rowcell[ay];
        const uint2 cc = *reinterpret_cast<const uint2*>(&s_colcell[4 * gxi]);  // four shorts
        uint32_t keep = 0;
        if (rc >= 0) {
            keep = ((short)(cc.x & 0xFFFF) >= 0 ? 0x000000FFu : 0u) | ((short)(cc.x >> 16) >= 0 ? 0x0000FF00u : 0u) |
                   ((short)(cc.y & 0xFFFF) >= 0 ? 0x00FF0000u : 0u) | ((short)(cc.y >> 16) >= 0 ? 0xFF000000u : 0u);
        }
        a4 &= keep;
        *reinterpret_cast<uint32_t*>(&s_A[ay * kAW + 4 * gxi]) = a4;
    }
    __syncthreads();

    // ---- cell-aware NMS over the tile interior; emit local maxima above minTh ----
    const DynState& dyn = v.dyn[frame];
    const int thIni = dyn.area_flag ? 30 : 20;   // threshold override, src/ORBextractor.cc:775-784
    const int thMin = dyn.area_flag ? 10 : 7;
    int* cellcnt = v.cell_count + (size_t)frame * g.cells_per_frame + L.cell_base;
    for (int i = tid; i < (kFtW / 4) * kFtH; i += kFtThreads) {
        const int py = i >> 4, px0 = (i & 15) * 4;
        const uint32_t a4 = *reinterpret_cast<const uint32_t*>(&s_A[(py + 1) * kAW + px0 + 4]);
        if ((a4 & 0xF8F8F8F8u) == 0u) continue;   // all four strengths < 8 <= minTh + 1: nothing to do (minTh is 7 or 10)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int A = (a4 >> (8 * k)) & 0xFF;
            if (A <= thMin) continue;
            const int px = px0 + k;
            const uint8_t* a = &s_A[(py + 1) * kAW + px + 4];
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
                if (A > thIni) atomicAdd(&cellcnt[ci * L.nCols + cj], 1);
            }
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

void launch_fast(const Geometry& g, const BatchView& v, cudaStream_t stream) {
    TileMap tm;
    int total = 0;
    for (int l = 0; l < g.nlevels; l++) {
        tm.tile_base[l] = total;
        // tiles cover the detection domain x in [19, w-19), y in [19, h-19), starting at the 4-aligned x = 16
        tm.tiles_x[l] = std::max(1, (g.lv[l].w - kEdge - kMinBorder + kFtW - 1) / kFtW);
        const int tiles_y = std::max(1, (g.lv[l].h - kEdge - kMinBorder + kFtH - 1) / kFtH);
        total += tm.tiles_x[l] * tiles_y;
    }
    tm.tile_base[g.nlevels] = total;
    cudaMemsetAsync(v.lmax_count, 0, sizeof(int) * (size_t)v.B * g.nlevels, stream);
    cudaMemsetAsync(v.cell_count, 0, sizeof(int) * (size_t)v.B * g.cells_per_frame, stream);
    fast_kernel<<<dim3(total, v.B), kFtThreads, 0, stream>>>(g, v, tm);
}

}  // namespace coeb
