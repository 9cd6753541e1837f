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
// Fallback kernel: the cells whose counter stayed 0 are listed and recomputed alone at minTh, one warp per cell, in the same
// packed two-pixel form; their local maxima are appended to the same list.
#include <cstdlib>

#include "coeb_device.cuh"

namespace coeb {

#ifndef COEB_FT_H
#define COEB_FT_H 64
#endif
#ifndef COEB_FT_MINB
#define COEB_FT_MINB 6
#endif
constexpr int kFtW = 64, kFtH = COEB_FT_H;     // output tile: 16 four-pixel groups x kFtH rows
constexpr int kFtThreads = 256;
constexpr int kImgWords = 24;                  // 96 staged bytes per row from x = tx0-16: six 16-byte loads
constexpr int kImgPitch = 4 * kImgWords;
constexpr int kImgRows = kFtH + 8;             // 3 (ring) + 1 (NMS halo) each side
constexpr int kAW = 72, kARows = kFtH + 2;     // A tile: x from tx0-4 (alignment), y from ty0-1
constexpr int kPairIters = kFtH / 16;          // 1a iterations: a warp covers 2 rows x 16 groups per iteration
constexpr int kRingPairs = 2 * (kFtW / 2 + 2) + 2 * kFtH;   // pixel pairs holding the 1 px ring around the tile
constexpr int kPairCap = kFtW * kFtH / 2 + kRingPairs;
static_assert(kFtH % 16 == 0 && kFtH <= 64 && kRingPairs <= kFtThreads, "tile shape");
// one shared arena (a single base register for every access), byte offsets
constexpr int oImg = 0;
constexpr int oA = oImg + kImgRows * kImgPitch;
constexpr int oQueue = oA + kARows * kAW;                 // ushort[kPairCap]     pixel pairs that pass the compass bound
constexpr int oCq = (oQueue + 2 * kPairCap + 15) & ~15;   // ushort[kFtW*kFtH]    pixels with A > iniTh (NMS candidates)
// Local maxima a tile can emit: a cell keeps at most one pixel of every 2x2 block (strict 3x3 maxima inside the cell), and 64 pixels
// touch at most four cells (cells are >= 30 px), so at most (32 + 2)^2 = 1156 per tile.
constexpr int kListCap = 1280;
static_assert(kFtW == 64 && kFtH <= 64 && kListCap >= (kFtW / 2 + 2) * (kFtH / 2 + 2), "local-maximum bound of a tile");
constexpr int oList = oCq + 2 * kFtW * kFtH;              // uint32[kListCap]  emitted local maxima
constexpr int oCtr = oList + 4 * kListCap;                // int[8]: nq, nc, n, base
constexpr int kFastSmem = oCtr + 32;
static_assert(oA % 16 == 0 && oQueue % 16 == 0 && (kARows * kAW) % 16 == 0, "16-byte stores");

__device__ __forceinline__ uint32_t pair_lo(uint32_t w) { return __byte_perm(w, 0u, 0x4140); }
__device__ __forceinline__ uint32_t pair_hi(uint32_t w) { return __byte_perm(w, 0u, 0x4342); }
__device__ __forceinline__ int smem_add(uint32_t addr, int val) {
    int old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(val) : "memory");
    return old;
}

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
    // A = max(bright, dark), clamped at 0
    return __vimax3_s16x2(__vsub2(best_b, 0x01000100u), __vsub2(0x01000100u, best_d), 0u);
}

// Can either pixel packed in `c` be a corner at the threshold? Every 9-arc holds two neighbouring compass points
// (ring 0, 4, 8, 12), so a bright corner needs d_k > th on two neighbouring compass points (mirrored for dark arcs).
// "Some neighbouring pair passes" is the complement of "{0, 8} both fail or {4, 12} both fail" (the vertex covers of the
// 4-cycle), so with one fail bit per compass point the test is two AND/OR steps per polarity. The fail bit is bit 14 of
// r_k + (0x4000 + th - c) per 16-bit lane (set iff r_k >= c - th, i.e. d_k <= th; lanes stay in (0x3f00, 0x4200), no carries),
// and of (0x4000 + th + c) - r_k for dark arcs: the eight additions run on the multiply pipe (IMAD), six logic operations and
// one compare on the ALU pipe FAST is bound by (the min/max form of the same bound took fifteen). tb = 0x40004000 + th * 0x10001.
__device__ __forceinline__ bool compass_bound2(uint32_t c, uint32_t r0, uint32_t r4, uint32_t r8, uint32_t r12, uint32_t tb) {
    const uint32_t kb = tb - c, kd = tb + c;
    const uint32_t fb = ((r0 + kb) & (r8 + kb)) | ((r4 + kb) & (r12 + kb));   // bit 14: no bright pair of neighbours
    const uint32_t fd = ((kd - r0) & (kd - r8)) | ((kd - r4) & (kd - r12));   // bit 14: no dark pair of neighbours
    return (~(fb & fd) & 0x40004000u) != 0u;
}

// FAST cell of a level coordinate (minBorder-relative detection coordinates, src/ORBextractor.cc:813-828); -1 outside the
// detection domain. rcp = ceil(2^20 / cell size): exact quotient for coordinates below 4128 (checked on the host).
__device__ __forceinline__ int cell_of(int x, int size, int rcp, int last) {
    return (x >= kEdge && x < size - kEdge) ? min((int)(((unsigned)(x - kEdge) * (unsigned)rcp) >> 20), last) : -1;
}

template <bool kTma>
__global__ void __launch_bounds__(kFtThreads, COEB_FT_MINB) fast_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v,
                                                             const int4* __restrict__ tiles, const __grid_constant__ TmaMaps maps) {
    __shared__ __align__(128) uint8_t smem[kFastSmem];
    __shared__ __align__(8) unsigned long long s_mbar;
    uint32_t* const s_img = reinterpret_cast<uint32_t*>(smem + oImg);
    uint8_t* const s_A = smem + oA;
    unsigned short* const s_queue = reinterpret_cast<unsigned short*>(smem + oQueue);
    unsigned short* const s_cq = reinterpret_cast<unsigned short*>(smem + oCq);
    uint32_t* const s_list = reinterpret_cast<uint32_t*>(smem + oList);
    int* const s_ctr = reinterpret_cast<int*>(smem + oCtr);   // 0 nq, 1 nc, 2 n, 3 base, 4 n2
    const uint32_t a_ctr = (uint32_t)__cvta_generic_to_shared(s_ctr);

    const int frame = blockIdx.y;
    const int4 ti = __ldg(&tiles[kFastTileInt4 * blockIdx.x]);   // {level, tx0, ty0, -} built on the host: no per-thread div/mod or level search
    const int level = ti.x, tx0 = ti.y, ty0 = ti.z;
    const LevelGeom& L = g.lv[level];
    const int tid = threadIdx.x, lane = tid & 31;
    COEB_TRACE(v, tiles[0].x == 0 ? 1 : 2);
    const int thIni = v.dyn[frame].area_flag ? 30 : 20;   // threshold override, src/ORBextractor.cc:775-784

    // ---- stage the tile: rows ty0-4 .. ty0+kFtH+3, 96 bytes from x = tx0-16 as six 16-byte loads per row. Rows below the
    //      image and columns beyond the row pitch are clamped: only pixels outside the detection domain (whose strength is
    //      never used) can see them, as can the padding bytes between w and the pitch.
    const uint32_t a_mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
    if (kTma) {
        // one thread hands the whole 96 x 72 byte box to the TMA engine (bytes outside the level read as 0, which only
        // out-of-domain pixels can see); the CTA zeroes its other arrays meanwhile and waits on the mbarrier below
        if (tid == 0) tma_issue_box(a_mbar, (uint32_t)__cvta_generic_to_shared(s_img), &maps.m[level], tx0 - 16, ty0 - 4, frame, kImgRows * kImgPitch);
    } else {
        const int pitch = level_pitch(g, v, level);
        const int q = tid & 7, r0 = tid >> 3;
        const uint8_t* __restrict__ src = level_ptr(g, v, level, frame) + min(tx0 - 16 + 16 * q, pitch - 16);
        if (q < 6) {
#pragma unroll
            for (int ry = r0; ry < kImgRows; ry += kFtThreads / 8)
                reinterpret_cast<uint4*>(s_img)[ry * 6 + q] = __ldg(reinterpret_cast<const uint4*>(src + (long long)min(ty0 - 4 + ry, L.h - 1) * pitch));
        }
    }
    for (int i = tid; i < kARows * kAW / 16; i += kFtThreads) reinterpret_cast<uint4*>(s_A)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid < 5) s_ctr[tid] = 0;
    __syncthreads();
    if (kTma) tma_wait(a_mbar);

    // detection domain [19, w-19) x [19, h-19) in tile coordinates, cut to the A tile [-2, kFtW+2) x [-1, kFtH+1)
    const int xlo = max(kEdge - tx0, -2), xspan = max(min(L.w - kEdge - tx0, kFtW + 2) - xlo, 0);
    const int ylo = max(kEdge - ty0, -1), yspan = max(min(L.h - kEdge - ty0, kFtH + 1) - ylo, 0);

    // ---- 1a: compass bound of every pixel pair; survivors are queued, which re-packs them densely over the threads.
    //      A warp takes 16 groups of rows y and y+2 (24-word row pitch: the two half-warps hit disjoint banks); the 1 px
    //      ring the NMS needs around the tile is one more pass of single pairs.
    {
        const uint32_t tb = 0x40004000u + (uint32_t)thIni * 0x00010001u;
        const int gxi = lane & 15, px = 4 * gxi;
        const int ay0 = 4 * (tid >> 6) + ((tid >> 5) & 1) + ((lane >> 4) << 1);   // + 16 per iteration
        const bool x01 = (unsigned)(px + 1 - xlo) < (unsigned)(xspan + 1), x23 = (unsigned)(px + 3 - xlo) < (unsigned)(xspan + 1);
        const uint32_t* row = &s_img[(ay0 + 4) * kImgWords + gxi + 4];   // word holding the 4 centre pixels
        unsigned flags = 0u;
#pragma unroll
        for (int it = 0; it < kPairIters; it++, row += 16 * kImgWords) {
            const uint32_t c = row[0];
            const uint32_t S0 = row[3 * kImgWords], S8 = row[-3 * kImgWords];
            const uint32_t S4 = __funnelshift_r(c, row[1], 24), S12 = __funnelshift_r(row[-1], c, 8);
            const bool yok = (unsigned)(ay0 + 16 * it - ylo) < (unsigned)yspan;
            const bool f01 = compass_bound2(pair_lo(c), pair_lo(S0), pair_lo(S4), pair_lo(S8), pair_lo(S12), tb) & yok & x01;
            const bool f23 = compass_bound2(pair_hi(c), pair_hi(S0), pair_hi(S4), pair_hi(S8), pair_hi(S12), tb) & yok & x23;
            flags |= ((unsigned)f01 << (2 * it)) | ((unsigned)f23 << (2 * it + 1));
        }
        // ring pairs: rows -1 and kFtH over x = -2 .. kFtW+1, columns (-2,-1) and (kFtW, kFtW+1) over the tile rows
        int re = 0;
        if (tid < kRingPairs) {   // whole warps except the last one
            int ray, rpi;   // A row, pair index (first pixel x = 2 * rpi)
            if (tid < kFtW / 2 + 2) { ray = -1; rpi = tid - 1; }
            else if (tid < kFtW + 4) { ray = kFtH; rpi = tid - (kFtW / 2 + 2) - 1; }
            else if (tid < kFtW + 4 + kFtH) { ray = tid - (kFtW + 4); rpi = -1; }
            else { ray = tid - (kFtW + 4 + kFtH); rpi = kFtW / 2; }
            const uint32_t* rr = &s_img[(ray + 4) * kImgWords + 4 + (rpi >> 1)];
            const uint32_t sel = (rpi & 1) ? 0x4342u : 0x4140u;
            const uint32_t c = rr[0];
            const uint32_t S4 = __funnelshift_r(c, rr[1], 24), S12 = __funnelshift_r(rr[-1], c, 8);
            const bool ok = (unsigned)(ray - ylo) < (unsigned)yspan && (unsigned)(2 * rpi + 1 - xlo) < (unsigned)(xspan + 1);
            const bool f = compass_bound2(__byte_perm(c, 0u, sel), __byte_perm(rr[3 * kImgWords], 0u, sel), __byte_perm(S4, 0u, sel),
                                          __byte_perm(rr[-3 * kImgWords], 0u, sel), __byte_perm(S12, 0u, sel), tb) & ok;
            re = ((ray + 1) << 6) | (rpi + 1);
            flags |= (unsigned)f << (2 * kPairIters);
        }
        // one queue reservation per warp: ballots give every set flag its slot
        unsigned m[2 * kPairIters + 1];
        int n = 0;
#pragma unroll
        for (int j = 0; j <= 2 * kPairIters; j++) { m[j] = __ballot_sync(0xffffffffu, (flags >> j) & 1u); n += __popc(m[j]); }
        if (n > 0) {
            int base = 0;
            if (lane == 0) base = smem_add(a_ctr, n);
            base = __shfl_sync(0xffffffffu, base, 0);
            const unsigned lt = (1u << lane) - 1u;
            const int e0 = ((ay0 + 1) << 6) | (2 * gxi + 1);   // A row + 1, pair index + 1
#pragma unroll
            for (int j = 0; j < 2 * kPairIters; j++) {
                if ((flags >> j) & 1u) s_queue[base + __popc(m[j] & lt)] = (unsigned short)(e0 + ((j >> 1) << 10) + (j & 1));
                base += __popc(m[j]);
            }
            if ((flags >> (2 * kPairIters)) & 1u) s_queue[base + __popc(m[2 * kPairIters] & lt)] = (unsigned short)re;
        }
    }
    __syncthreads();

    // ---- 1b: exact corner strength of the queued pairs, branch-free ----
    {
        const int nq = s_ctr[0];
#pragma unroll 1
        for (int qi = tid; qi < nq; qi += kFtThreads) {
            const int e = s_queue[qi];
            const int er = e >> 6, ep = e & 63;                 // A row + 1, pair index + 1
            const int px = 2 * ep - 2;                          // first pixel of the pair, tile coordinate
            const uint32_t* row = &s_img[(er + 3) * kImgWords + 4 + ((ep - 1) >> 1)];
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
            const uint32_t sel = (ep & 1) ? 0x4140u : 0x4342u;   // odd ep = even pair index = low half of the word
            uint32_t r[16];
#pragma unroll
            for (int k = 0; k < 16; k++) r[k] = __byte_perm(S[k], 0u, sel);
            const uint32_t a2 = corner_strength2(__byte_perm(row[0], 0u, sel), r);
            // blank the pixel of a straddling pair that lies outside the detection domain; one strength per byte
            const uint32_t lo = (unsigned)(px - xlo) < (unsigned)xspan ? (a2 & 0xFFu) : 0u;
            const uint32_t hi = (unsigned)(px + 1 - xlo) < (unsigned)xspan ? (a2 >> 16) : 0u;
            *reinterpret_cast<unsigned short*>(&s_A[er * kAW + px + 4]) = (unsigned short)(lo | (hi << 8));
            // tile pixels above iniTh are the NMS candidates: queue them for a dense second pass
            const bool in = (unsigned)(er - 1) < (unsigned)kFtH && (unsigned)px < (unsigned)kFtW;
            const bool q0 = in && (int)lo > thIni, q1 = in && (int)hi > thIni;
            if (q0 || q1) {
                const int id = ((er - 1) << 6) | px;
                const int pos = smem_add(a_ctr + 4, (int)q0 + (int)q1);
                if (q0) s_cq[pos] = (unsigned short)id;
                if (q1) s_cq[pos + (int)q0] = (unsigned short)(id + 1);
            }
        }
    }
    __syncthreads();

    // ---- 2: cell-aware NMS at iniTh. 2a tests every candidate against its 8 neighbours on full warps; the few that are kept, and
    //      the suppressed ones that touch a cell boundary (a larger neighbour only counts inside the same cell: each cell is an
    //      independent cv::FAST call), go to a second queue, so that the boundary test and the emission (2b) run on full warps too.
    //      The queue lives in the staged tile, which is dead after 1b.
    const int nc = s_ctr[1];
    // cell-boundary masks of the tile (host-built, build_fast_tiles): bit i of eL / eR = the left / right neighbour of column tx0+i lies
    // in another FAST cell (or outside the detection domain) and so cannot suppress it; eU / eD the same for the rows above / below
    const int4 ex_ = __ldg(&tiles[kFastTileInt4 * blockIdx.x + 1]), ey_ = __ldg(&tiles[kFastTileInt4 * blockIdx.x + 2]);
    const unsigned long long eL = (unsigned)ex_.x | ((unsigned long long)(unsigned)ex_.y << 32), eR = (unsigned)ex_.z | ((unsigned long long)(unsigned)ex_.w << 32);
    const unsigned long long eU = (unsigned)ey_.x | ((unsigned long long)(unsigned)ey_.y << 32), eD = (unsigned)ey_.z | ((unsigned long long)(unsigned)ey_.w << 32);
    int* cellcnt = v.cell_count + (size_t)frame * g.cells_per_frame + L.cell_base;
    unsigned short* const s_q2 = reinterpret_cast<unsigned short*>(smem + oImg);
    constexpr int kQ2Cap = kImgRows * kImgPitch / 2;
    auto emit = [&](int px, int py, int A) {   // level coordinates; emitted minBorder-relative (:844-845)
        s_list[smem_add(a_ctr + 8, 1)] = (uint32_t)(tx0 + px - kMinBorder) | ((uint32_t)(ty0 + py - kMinBorder) << 12) | ((uint32_t)(A - 1) << 24);
    };
    auto keep_in_cell = [&](int px, int py) {   // the NMS of a pixel that touches a cell boundary: neighbours of other cells do not count
        const uint8_t* a = &s_A[(py + 1) * kAW + px + 4];
        const bool sl = !((eL >> px) & 1ull), sr = !((eR >> px) & 1ull), su = !((eU >> py) & 1ull), sd = !((eD >> py) & 1ull);
        int m = 0;
        if (sl) m = max(m, (int)a[-1]);
        if (sr) m = max(m, (int)a[1]);
        if (su) { m = max(m, (int)a[-kAW]); if (sl) m = max(m, (int)a[-kAW - 1]); if (sr) m = max(m, (int)a[-kAW + 1]); }
        if (sd) { m = max(m, (int)a[kAW]); if (sl) m = max(m, (int)a[kAW - 1]); if (sr) m = max(m, (int)a[kAW + 1]); }
        return (int)a[0] > m;
    };
    for (int i = tid; i < nc; i += kFtThreads) {
        const int e = s_cq[i];
        const int py = e >> 6, px = e & 63;
        const uint8_t* a = &s_A[(py + 1) * kAW + px + 4];
        const int A = a[0];
        const int n_l = a[-1], n_r = a[1], n_ul = a[-kAW - 1], n_u = a[-kAW], n_ur = a[-kAW + 1], n_dl = a[kAW - 1], n_d = a[kAW], n_dr = a[kAW + 1];
        const bool keep = A > max(max(max(n_l, n_r), max(n_ul, n_u)), max(max(n_ur, n_dl), max(n_d, n_dr)));
        const bool edge = !keep && ((((eL | eR) >> px) | ((eU | eD) >> py)) & 1ull);
        if (keep || edge) {
            const int pos = smem_add(a_ctr + 16, 1);
            if (pos < kQ2Cap) s_q2[pos] = (unsigned short)(e | (edge ? 0x8000 : 0));
            else if (keep || keep_in_cell(px, py)) emit(px, py, A);   // more entries than a tile can plausibly produce: handled in place
        }
    }
    __syncthreads();
    {
        const int n2 = min(s_ctr[4], kQ2Cap);
        for (int i = tid; i < n2; i += kFtThreads) {
            const int e = s_q2[i];
            const int py = (e >> 6) & 63, px = e & 63;
            if (!(e & 0x8000) || keep_in_cell(px, py)) emit(px, py, s_A[(py + 1) * kAW + px + 4]);
        }
    }
    __syncthreads();
    const int n = s_ctr[2];
    if (n == 0) return;
    if (tid == 0) s_ctr[3] = atomicAdd(v.lmax_count + frame * g.nlevels + level, n);
    __syncthreads();
    const int base = s_ctr[3];
    uint32_t* out = v.lmax + (size_t)frame * g.cand_per_frame + L.cand_base + base;
    const int room = L.cand_cap - base;
    for (int i = tid; i < n; i += kFtThreads) {
        const uint32_t key = s_list[i];
        const int x = (int)(key & 0xFFFu) + kMinBorder, y = (int)((key >> 12) & 0xFFFu) + kMinBorder;
        atomicAdd(&cellcnt[cell_of(y, L.h, L.rcpH, L.lastI) * L.nCols + cell_of(x, L.w, L.rcpW, L.lastJ)], 1);
        if (i < room) out[i] = key;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// minTh fallback (src/ORBextractor.cc:834-838): cells without a keypoint at iniTh are redone at minTh. Persistent CTAs
// scan the per-cell counters written by fast_kernel; an empty cell is processed alone, exactly like one cv::FAST call
// on its ROI: scalar sliding-window A, NMS inside the cell, local maxima above minTh appended to the level's list.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kMaxRoi = 72;   // ROI side bound: wCell + 6 <= 72 (checked on the host)
constexpr int kRoiPitch = kMaxRoi + 4;   // staged with aligned word loads: up to 3 bytes of lead-in per row

// Lists the FAST cells that exist (src/ORBextractor.cc:816-826) and produced nothing at iniTh. One thread per cell.
__global__ void __launch_bounds__(256) fast_empty_cells_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v) {
    COEB_TRACE(v, 11);
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

// One WARP per empty cell, no block barrier: the cell's ROI is staged as aligned words, the compass bound runs in the packed
// two-pixel form of the main kernel on four pixels per lane and step, the few survivors are scored exactly (same code path as
// 1b above) and the pixels above minTh are remembered, so that the cell-local 3x3 NMS only visits those. The per-warp shared
// region is sized on the host from the largest cell of the geometry (a 640x480 pyramid needs 4.6 KB per warp, so that an SM
// holds 40 such warps; the compile-time maximum would be 12 KB): the kernel is bound by load and shared-memory latency.
#ifndef COEB_FB_WARPS
#define COEB_FB_WARPS 8
#endif
constexpr int kFbWarps = COEB_FB_WARPS;
constexpr int kFbQueue = 192;             // survivor queue entries per warp, flushed before a step could overflow it
constexpr int kFbList = 96;               // staged local maxima per warp, flushed to the level's list when nearly full
#ifndef COEB_FB_STRIPS
#define COEB_FB_STRIPS 2
#endif
constexpr int kFbStrips = COEB_FB_STRIPS;   // small batches: horizontal strips per cell, one warp each
constexpr int kFbCand = 160;              // remembered pixels above minTh per warp; more than that: the NMS scans the whole map
struct FbLayout {
    int pw;        // staged row pitch in words: 1 pad + data words (3 lead-in + widest ROI + 3) + 1 pad
    int rows;      // tallest ROI
    int aw;        // strength map row pitch in bytes (multiple of 16)
    int arows;     // detection rows + a zero row above and below
    int img_bytes; // staged ROI, rounded up to 16
    int bytes;     // per-warp region
};
__host__ __device__ inline FbLayout fb_layout(const Geometry& g) {
    int wc = 0, hc = 0;
    for (int l = 0; l < g.nlevels; l++) { wc = max(wc, g.lv[l].wCell); hc = max(hc, g.lv[l].hCell); }
    FbLayout f;
    f.pw = ((3 + wc + 6 + 3) >> 2) + 2;
    f.rows = hc + 6;
    f.aw = (wc + 8 + 15) & ~15;
    f.arows = hc + 2;
    f.img_bytes = (f.rows * f.pw * 4 + 15) & ~15;
    f.bytes = (f.img_bytes + f.arows * f.aw + 2 * kFbQueue + 4 * kFbList + 2 * kFbCand + 15) & ~15;
    return f;
}

// kDirect (small batches): no list of empty cells; the warps cover every cell of levels [cell_lo, cell_hi) of every frame and each
// one tests its own cell's counter (and the geometry conditions of fast_empty_cells_kernel), which takes the list kernel and one
// dependent global load off the latency path.
template <bool kDirect>
__global__ void __launch_bounds__(32 * kFbWarps) fast_fallback_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v,
                                                                      const __grid_constant__ FbLayout F, const int cell_lo, const int cell_hi) {
    extern __shared__ __align__(16) uint8_t fb_smem[];
    COEB_TRACE(v, kDirect && cell_lo > 0 ? 4 : 3);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* const region = fb_smem + (size_t)warp * F.bytes;
    uint32_t* const s_img = reinterpret_cast<uint32_t*>(region);
    uint8_t* const s_A = region + F.img_bytes;
    unsigned short* const s_queue = reinterpret_cast<unsigned short*>(s_A + F.arows * F.aw);
    uint32_t* const s_list = reinterpret_cast<uint32_t*>(s_queue + kFbQueue);
    unsigned short* const s_cand = reinterpret_cast<unsigned short*>(s_list + kFbList);
    const int PW = F.pw, AW = F.aw;
    const unsigned lt = (1u << lane) - 1u;
    // kDirect also cuts a cell into kFbStrips horizontal strips, one warp each: a single frame has a few hundred empty cells and
    // 148 SMs, and one warp's compass pass and exact scores of a whole 30 x 30 cell at minTh were 6.5 of the kernel's 10 us. A strip
    // computes one more row above and below (the 3x3 NMS of its own rows reads them) and emits its own rows only; the warps never meet.
    constexpr int kStrips = kDirect ? kFbStrips : 1;
    const int n_empty = kDirect ? v.B * (cell_hi - cell_lo) * kStrips : *v.empty_count;
    for (int e = blockIdx.x * kFbWarps + warp; e < n_empty; e += gridDim.x * kFbWarps) {
        int c, strip = 0;
        if (kDirect) {
            const int per_frame = (cell_hi - cell_lo) * kStrips;
            const int f = e / per_frame, rem = e - f * per_frame;
            strip = rem % kStrips;
            c = f * g.cells_per_frame + cell_lo + rem / kStrips;
            if (v.cell_count[c] != 0) continue;
        } else {
            c = v.empty_cells[e];
        }
        const int frame = c / g.cells_per_frame, cf = c - frame * g.cells_per_frame;
        int level = 0;
        while (level + 1 < g.nlevels && cf >= g.lv[level + 1].cell_base) level++;
        const LevelGeom& L = g.lv[level];
        const int cell = cf - L.cell_base;
        const int ci = cell / L.nCols, cj = cell - ci * L.nCols;
        // cell ROI (src/ORBextractor.cc:813-828), level coordinates
        const int iniX = kMinBorder + cj * L.wCell;
        int iniY = kMinBorder + ci * L.hCell;
        if (kDirect && (iniY >= L.maxBY - 3 || iniX >= L.maxBX - 6)) continue;   // cells the reference skips (:816-826)
        const int rw = min(iniX + L.wCell + 6, L.maxBX) - iniX;
        int rh = min(iniY + L.hCell + 6, L.maxBY) - iniY;
        if (kDirect && (rw < 7 || rh < 7)) continue;
        // rows of the cell's detection area this warp emits [y0, y1) and computes [cy0, cy0 + dh): the whole cell unless kDirect
        int y0 = 0, y1 = rh - 6, cy0 = 0;
        if (kStrips > 1) {
            const int dh_cell = rh - 6;
            y0 = strip * dh_cell / kStrips; y1 = (strip + 1) * dh_cell / kStrips;
            if (y1 <= y0) continue;
            cy0 = max(0, y0 - 1);
            const int cy1 = min(dh_cell, y1 + 1);
            iniY += cy0;               // the strip is staged and scored like a short cell of its own
            rh = cy1 - cy0 + 6;
        }
        const int dw = rw - 6, dh = rh - 6;          // detection area: ROI rows/cols [3, dim-3)
        const int thMin = v.dyn[frame].area_flag ? 10 : 7;
        const int pitch = level_pitch(g, v, level);
        const int ax = iniX & 3, nw = (ax + rw + 3) >> 2;   // aligned words per ROI row (the ROI ends 16 px before the row does)
        const uint8_t* __restrict__ img = level_ptr(g, v, level, frame) + (size_t)iniY * pitch + (iniX - ax);
        const int aoff = 4 + ((ax + 1) & 1);        // strength map column of detection x = 0: keeps a pixel pair 2-byte aligned
        __syncwarp();   // the previous cell's shared data is fully consumed
        {
            const uint32_t rcp = ((1u << 16) + nw - 1) / nw;   // i / nw == (i * rcp) >> 16 for i < 72 * 20
            for (int i = lane; i < rh * nw; i += 32) {
                const int y = (int)(((uint32_t)i * rcp) >> 16), wi = i - y * nw;
                s_img[y * PW + 1 + wi] = __ldg(reinterpret_cast<const uint32_t*>(img + (size_t)y * pitch) + wi);
            }
            for (int i = lane; i < (dh + 2) * AW / 16; i += 32) reinterpret_cast<uint4*>(s_A)[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncwarp();
        int n_out = 0, n_cand = 0;   // staged local maxima, remembered pixels above minTh (warp-uniform)
        auto flush_list = [&]() {
            int base = 0;
            if (lane == 0) base = atomicAdd(v.lmax_count + frame * g.nlevels + level, n_out);
            base = __shfl_sync(0xffffffffu, base, 0);
            uint32_t* out = v.lmax + (size_t)frame * g.cand_per_frame + L.cand_base;
            for (int i = lane; i < n_out; i += 32)
                if (base + i < L.cand_cap) out[base + i] = s_list[i];
            n_out = 0;
            __syncwarp();
        };
        // exact strength of the queued pairs -> strength map (pixels outside the detection columns are blanked)
        auto score_queue = [&](int nq) {
            for (int q0 = 0; q0 < nq; q0 += 32) {
                const int q = q0 + lane;
                bool c0 = false, c1 = false;
                int id = 0;
                if (q < nq) {
                    const int ent = s_queue[q];
                    const int y = ent >> 6, p = ent & 63;            // staged row, pair index (first pixel = staged column 2p)
                    const uint32_t* row = &s_img[y * PW + 1 + (p >> 1)];
                    const uint32_t *r3 = row + 3 * PW, *rm3 = row - 3 * PW, *r2 = row + 2 * PW, *rm2 = row - 2 * PW, *r1 = row + PW, *rm1 = row - PW;
                    uint32_t S[16];
                    S[0] = r3[0];                                   S[1] = __funnelshift_r(r3[0], r3[1], 8);
                    S[2] = __funnelshift_r(r2[0], r2[1], 16);       S[3] = __funnelshift_r(r1[0], r1[1], 24);
                    S[4] = __funnelshift_r(row[0], row[1], 24);     S[5] = __funnelshift_r(rm1[0], rm1[1], 24);
                    S[6] = __funnelshift_r(rm2[0], rm2[1], 16);     S[7] = __funnelshift_r(rm3[0], rm3[1], 8);
                    S[8] = rm3[0];                                  S[9] = __funnelshift_r(rm3[-1], rm3[0], 24);
                    S[10] = __funnelshift_r(rm2[-1], rm2[0], 16);   S[11] = __funnelshift_r(rm1[-1], rm1[0], 8);
                    S[12] = __funnelshift_r(row[-1], row[0], 8);    S[13] = __funnelshift_r(r1[-1], r1[0], 8);
                    S[14] = __funnelshift_r(r2[-1], r2[0], 16);     S[15] = __funnelshift_r(r3[-1], r3[0], 24);
                    const uint32_t sel = (p & 1) ? 0x4342u : 0x4140u;
                    uint32_t r[16];
#pragma unroll
                    for (int k = 0; k < 16; k++) r[k] = __byte_perm(S[k], 0u, sel);
                    const uint32_t a2 = corner_strength2(__byte_perm(row[0], 0u, sel), r);
                    const int dx = 2 * p - ax - 3;                   // detection column of the pair's first pixel
                    const uint32_t lo = (unsigned)dx < (unsigned)dw ? (a2 & 0xFFu) : 0u;
                    const uint32_t hi = (unsigned)(dx + 1) < (unsigned)dw ? (a2 >> 16) : 0u;
                    id = (y - 3) * AW + dx + aoff;                   // byte offset inside the map, without the zero row
                    *reinterpret_cast<unsigned short*>(&s_A[id + AW]) = (unsigned short)(lo | (hi << 8));
                    c0 = (int)lo > thMin;
                    c1 = (int)hi > thMin;
                }
                const unsigned m0 = __ballot_sync(0xffffffffu, c0), m1 = __ballot_sync(0xffffffffu, c1);
                const int add = __popc(m0) + __popc(m1);
                if (n_cand + add <= kFbCand) {
                    if (c0) s_cand[n_cand + __popc(m0 & lt)] = (unsigned short)id;
                    if (c1) s_cand[n_cand + __popc(m0) + __popc(m1 & lt)] = (unsigned short)(id + 1);
                    n_cand += add;
                } else {
                    n_cand = kFbCand + 1;   // too many to remember: the NMS scans the map
                }
            }
            __syncwarp();
        };
        // pass A: compass bound at minTh on every pixel pair of the detection area
        {
            const uint32_t tb = 0x40004000u + (uint32_t)thMin * 0x00010001u;
            const int g0 = (ax + 3) >> 2, ng = ((ax + 3 + dw - 1) >> 2) - g0 + 1;   // aligned 4-pixel groups that touch the detection columns
            const uint32_t rcpg = ((1u << 16) + ng - 1) / ng;
            const int total = dh * ng;
            int nq = 0;   // warp-uniform
            for (int i0 = 0; i0 < total; i0 += 32) {
                const int i = i0 + lane;
                bool f01 = false, f23 = false;
                int ent = 0;
                if (i < total) {
                    const int yy = (int)(((uint32_t)i * rcpg) >> 16), gi = g0 + (i - yy * ng);
                    const int y = yy + 3;
                    const uint32_t* row = &s_img[y * PW + 1 + gi];
                    const uint32_t cw = row[0];
                    const uint32_t S0 = row[3 * PW], S8 = row[-3 * PW];
                    const uint32_t S4 = __funnelshift_r(cw, row[1], 24), S12 = __funnelshift_r(row[-1], cw, 8);
                    const int dx = 4 * gi - ax - 3;   // detection column of the group's first pixel
                    const bool x01 = (unsigned)(dx + 1) < (unsigned)(dw + 1), x23 = (unsigned)(dx + 3) < (unsigned)(dw + 1);   // either pixel of the pair inside
                    f01 = x01 && compass_bound2(pair_lo(cw), pair_lo(S0), pair_lo(S4), pair_lo(S8), pair_lo(S12), tb);
                    f23 = x23 && compass_bound2(pair_hi(cw), pair_hi(S0), pair_hi(S4), pair_hi(S8), pair_hi(S12), tb);
                    ent = (y << 6) | (2 * gi);
                }
                const unsigned m0 = __ballot_sync(0xffffffffu, f01), m1 = __ballot_sync(0xffffffffu, f23);
                if (f01) s_queue[nq + __popc(m0 & lt)] = (unsigned short)ent;
                nq += __popc(m0);
                if (f23) s_queue[nq + __popc(m1 & lt)] = (unsigned short)(ent + 1);
                nq += __popc(m1);
                if (nq > kFbQueue - 64) {   // the next step may add 64 entries
                    __syncwarp();
                    score_queue(nq);
                    nq = 0;
                }
            }
            __syncwarp();
            score_queue(nq);
        }
        // cell-local NMS (each cell is an independent cv::FAST call) of the pixels above minTh
        auto nms_emit = [&](bool have, int id) {   // id: byte offset of the pixel inside the map without the zero row
            bool keep = false;
            int A = 0;
            if (have) {
                const uint8_t* a = &s_A[id + AW];
                A = a[0];
                if (A > thMin) {
                    const int nb = max(max(max(a[-1], a[1]), max(a[-AW - 1], a[-AW])), max(max(a[-AW + 1], a[AW - 1]), max(a[AW], a[AW + 1])));
                    keep = A > nb;
                }
            }
            if (kStrips > 1 && keep) { const int row = id / AW + cy0; keep = row >= y0 && row < y1; }   // halo rows belong to the neighbouring strip
            const unsigned mk = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int dy = id / AW, col = id - dy * AW;
                const int px = (col - aoff) + 3 + cj * L.wCell, py = dy + cy0 + 3 + ci * L.hCell;   // minBorder-relative (:844-845)
                s_list[n_out + __popc(mk & lt)] = (uint32_t)px | ((uint32_t)py << 12) | ((uint32_t)(A - 1) << 24);
            }
            n_out += __popc(mk);
            if (n_out > kFbList - 32) { __syncwarp(); flush_list(); }
        };
        if (n_cand <= kFbCand) {
            for (int i0 = 0; i0 < n_cand; i0 += 32) nms_emit(i0 + lane < n_cand, i0 + lane < n_cand ? (int)s_cand[i0 + lane] : 0);
        } else {
            const int total = dh * AW;
            for (int i0 = 0; i0 < total; i0 += 32) nms_emit(i0 + lane < total, i0 + lane);
        }
        __syncwarp();
        if (n_out > 0) flush_list();
    }
}

// Tiles cover the detection domain x in [19, w-19), y in [19, h-19) of every level, starting at the 4-aligned x = 16.
static int cell_of_host(int x, int size, int rcp, int last) {   // cell_of() above
    return (x >= kEdge && x < size - kEdge) ? std::min((int)(((unsigned)(x - kEdge) * (unsigned)rcp) >> 20), last) : -1;
}

// kFastTileInt4 entries per tile: {level, tx0, ty0, 0}, then the tile's cell-boundary masks {eL, eR}, {eU, eD} as 64-bit pairs (the
// kernel used to derive them with 128 threads, three divisions each and four ballots per tile: ~4 % of its instructions).
int build_fast_tiles(const Geometry& g, int4* out) {
    int total = 0;
    for (int l = 0; l < g.nlevels; l++) {
        const LevelGeom& L = g.lv[l];
        const int tiles_x = std::max(1, (L.w - kEdge - kMinBorder + kFtW - 1) / kFtW);
        const int tiles_y = std::max(1, (L.h - kEdge - kMinBorder + kFtH - 1) / kFtH);
        for (int ty = 0; ty < tiles_y; ty++)
            for (int tx = 0; tx < tiles_x; tx++, total++) {
                if (!out) continue;
                const int tx0 = kMinBorder + tx * kFtW, ty0 = kMinBorder + ty * kFtH;
                unsigned long long e[4] = {0, 0, 0, 0};   // L, R, U, D
                for (int i = 0; i < 64; i++) {
                    const int cx = cell_of_host(tx0 + i, L.w, L.rcpW, L.lastJ), cy = cell_of_host(ty0 + i, L.h, L.rcpH, L.lastI);
                    if (cell_of_host(tx0 + i - 1, L.w, L.rcpW, L.lastJ) != cx) e[0] |= 1ull << i;
                    if (cell_of_host(tx0 + i + 1, L.w, L.rcpW, L.lastJ) != cx) e[1] |= 1ull << i;
                    if (cell_of_host(ty0 + i - 1, L.h, L.rcpH, L.lastI) != cy) e[2] |= 1ull << i;
                    if (cell_of_host(ty0 + i + 1, L.h, L.rcpH, L.lastI) != cy) e[3] |= 1ull << i;
                }
                int4* t = out + (size_t)kFastTileInt4 * total;
                t[0] = make_int4(l, tx0, ty0, 0);
                t[1] = make_int4((int)(unsigned)e[0], (int)(unsigned)(e[0] >> 32), (int)(unsigned)e[1], (int)(unsigned)(e[1] >> 32));
                t[2] = make_int4((int)(unsigned)e[2], (int)(unsigned)(e[2] >> 32), (int)(unsigned)e[3], (int)(unsigned)(e[3] >> 32));
            }
    }
    return total;
}

bool tma_enabled() {
    static const bool on = [] { const char* e = getenv("COEB_TMA"); return !e || atoi(e) != 0; }();
    return on;
}

bool encode_level_maps(const Geometry& g, const BatchView& v, int box_w, int box_h, TmaMaps* out, bool blurred) {
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiled encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
        return (EncodeTiled)fn;
    }();
    if (!encode) return false;
    for (int l = 0; l < g.nlevels; l++) {
        const cuuint64_t pitch = blurred ? (cuuint64_t)g.lv[l].pitch : (cuuint64_t)level_pitch(g, v, l);
        const cuuint64_t fstride = (l == 0 && !blurred) ? v.l0_stride : g.lv[l].img_stride;
        const cuuint64_t dims[3] = {pitch, (cuuint64_t)g.lv[l].h, (cuuint64_t)v.B};
        const cuuint64_t strides[2] = {pitch, fstride};
        const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u}, estr[3] = {1u, 1u, 1u};
        void* base = blurred ? (void*)blur_ptr(g, v, l, 0) : (void*)const_cast<uint8_t*>(level_ptr(g, v, l, 0));
        if (((uintptr_t)base | pitch | fstride) & 15) return false;
        if (encode(&out->m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
}

// The FAST stage in three pieces, so that a small batch can run the level-0 tiles (which need no resize) beside the pyramid chain:
//   (classify_kernel, which always runs first, zeroes the per-level and per-cell counters)
//   launch_fast_tiles  runs the main kernel on tiles [first, first + count) of the level-major tile table;
//   launch_fast_tail   lists the cells that stayed empty and redoes them at minTh.
int fast_tiles_of_level0(const Geometry& g) {
    const int tiles_x = std::max(1, (g.lv[0].w - kEdge - kMinBorder + kFtW - 1) / kFtW);
    const int tiles_y = std::max(1, (g.lv[0].h - kEdge - kMinBorder + kFtH - 1) / kFtH);
    return tiles_x * tiles_y;
}

void launch_fast_tiles(const Geometry& g, const BatchView& v, cudaStream_t stream, int first, int count) {
    if (count <= 0) return;
    TmaMaps maps;
    if (tma_enabled() && encode_level_maps(g, v, kImgPitch, kImgRows, &maps)) fast_kernel<true><<<dim3(count, v.B), kFtThreads, 0, stream>>>(g, v, v.fast_tiles + (size_t)kFastTileInt4 * first, maps);
    else fast_kernel<false><<<dim3(count, v.B), kFtThreads, 0, stream>>>(g, v, v.fast_tiles + (size_t)kFastTileInt4 * first, maps);
}

static size_t fallback_smem(const FbLayout& F) {
    const size_t smem = (size_t)F.bytes * kFbWarps;
    static size_t configured[64] = {};   // opt-in shared-memory size: a per-device function attribute
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > configured[dev & 63]) {
        cudaFuncSetAttribute(fast_fallback_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(fast_fallback_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63] = smem;
    }
    return smem;
}

void launch_fast_tail(const Geometry& g, const BatchView& v, cudaStream_t stream) {
    const int cells = v.B * g.cells_per_frame;
    cudaMemsetAsync(v.empty_count, 0, sizeof(int), stream);
    fast_empty_cells_kernel<<<(cells + 255) / 256, 256, 0, stream>>>(g, v);
    const FbLayout F = fb_layout(g);
    const size_t smem = fallback_smem(F);
    const int ctas_per_sm = std::max(1, std::min(8, (int)((200u << 10) / smem)));
    fast_fallback_kernel<false><<<148 * ctas_per_sm, 32 * kFbWarps, smem, stream>>>(g, v, F, 0, 0);
}

void launch_fast_tail_levels(const Geometry& g, const BatchView& v, cudaStream_t stream, int level_lo, int level_hi) {
    if (level_hi <= level_lo) return;
    const int cell_lo = g.lv[level_lo].cell_base, cell_hi = level_hi < g.nlevels ? g.lv[level_hi].cell_base : g.cells_per_frame;
    const FbLayout F = fb_layout(g);
    const size_t smem = fallback_smem(F);
    const int warps = v.B * (cell_hi - cell_lo) * kFbStrips;
    fast_fallback_kernel<true><<<(warps + kFbWarps - 1) / kFbWarps, 32 * kFbWarps, smem, stream>>>(g, v, F, cell_lo, cell_hi);
}

void launch_fast(const Geometry& g, const BatchView& v, cudaStream_t stream) {
    launch_fast_tiles(g, v, stream, 0, g.fast_tiles_per_frame);
    launch_fast_tail(g, v, stream);
}

}  // namespace coeb
