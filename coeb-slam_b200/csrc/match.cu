// match.cu -- Frame keypoint grid (F2, F3), Hamming matchers (M1-M6), stereo matcher (F4) and the
// brute-force k=2 matcher (config 5) behind the C ABI of include/coeb_frontend.h.
//
// Reference: Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea (src/Frame.cc:396-411, 503-568),
// ORBmatcher (src/ORBmatcher.cc:37-137, 405-520, 1329-1471, 1602-1664), Frame::ComputeStereoMatches
// (src/Frame.cc:644-818).
//
// Loop-carried state. The reference matchers are sequential: a query skips keypoints that EARLIER
// queries of the same call already claimed (:87-89 with :123; :1404-1406 with :1429), and
// SearchForInitialization carries vMatchedDistance / vnMatches21 (:444, :463-471). The result of
// query i is a function F(i, results of queries j < i). The kernels evaluate all queries in
// parallel and iterate R <- F(R) to a fixed point; any fixed point equals the sequential result
// (induction on i), and the iteration reaches it after at most n rounds (round t fixes queries < t);
// in practice 2-4 rounds, since conflicts are rare and local. The iteration runs in ONE CTA; m2_resolve_cached_kernel is its
// form cut for the latency of a single call (records written by the collect kernels, rotating claim tables, a programmatic
// dependent launch; SearchByProjection, SearchLocalPoints, the frame-to-frame search and its relocalisation overload), the
// m2 / m3 / m4 resolve kernels the general ones (large maps, SearchForInitialization, the window-walking fallback).
//
// Tie-breaks follow the grid traversal order of GetFeaturesInArea (ix outer, iy inner, insertion
// order inside a cell): the device grid stores cells ix-major, items in ascending keypoint index, so
// one query walks, per ix, one contiguous item range, and "first candidate wins" (strict '<') is the
// natural outcome of the per-thread sequential scan.
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstring>
#include <mutex>
#include <set>
#include <vector>

#include "../../include/coeb_frontend.h"
#include "coeb_device.cuh"
#include "coeb_host.hpp"

namespace coeb {

constexpr int kGridCells = COEB_GRID_COLS * COEB_GRID_ROWS;
constexpr int kInf = 0x7fffffff;

struct FrameDev {
    int n;
    const float* x; const float* y; const float* angle; const int* octave;
    const uint32_t* desc;       // n x 8 words
    const float* uright;        // n floats, or nullptr (all -1)
    const int* cell_start;      // [kGridCells + 1], cell = ix * ROWS + iy
    const int* cell_items;      // [n_in_grid] keypoint indices, ascending inside a cell
    float min_x, min_y, max_x, max_y, gw_inv, gh_inv;
    float fx, fy, cx, cy, bf, b;
    float scale[COEB_MAX_LEVELS];
};

__device__ __forceinline__ int hamming256(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b) {
    const uint4 a0 = *reinterpret_cast<const uint4*>(a), a1 = *reinterpret_cast<const uint4*>(a + 4);
    const uint4 b0 = *reinterpret_cast<const uint4*>(b), b1 = *reinterpret_cast<const uint4*>(b + 4);
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
           __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// Frame::GetFeaturesInArea (src/Frame.cc:503-556): calls fn(idx) for every keypoint of the window in the
// reference's order.
template <class F>
__device__ __forceinline__ void for_each_in_area(const FrameDev& f, float x, float y, float r, int minLevel, int maxLevel, F&& fn) {
    const int nMinCellX = max(0, (int)floorf((x - f.min_x - r) * f.gw_inv));
    if (nMinCellX >= COEB_GRID_COLS) return;
    const int nMaxCellX = min(COEB_GRID_COLS - 1, (int)ceilf((x - f.min_x + r) * f.gw_inv));
    if (nMaxCellX < 0) return;
    const int nMinCellY = max(0, (int)floorf((y - f.min_y - r) * f.gh_inv));
    if (nMinCellY >= COEB_GRID_ROWS) return;
    const int nMaxCellY = min(COEB_GRID_ROWS - 1, (int)ceilf((y - f.min_y + r) * f.gh_inv));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++) {
        const int lo = f.cell_start[ix * COEB_GRID_ROWS + nMinCellY], hi = f.cell_start[ix * COEB_GRID_ROWS + nMaxCellY + 1];
        for (int p = lo; p < hi; p++) {
            const int idx = f.cell_items[p];
            if (bCheckLevels) {
                const int o = f.octave[idx];
                if (o < minLevel) continue;
                if (maxLevel >= 0 && o > maxLevel) continue;
            }
            const float dx = f.x[idx] - x, dy = f.y[idx] - y;
            if (fabsf(dx) < r && fabsf(dy) < r) fn(idx);
        }
    }
}

// Warp-cooperative form of the same walk. The window's columns are contiguous item ranges (cells are stored ix-major), so
// the raw items of the window are numbered column by column with a warp scan, 32 of them are tested per trip (level and
// distance filters of GetFeaturesInArea, then accept(idx) for the caller's own static filters) and the survivors are
// ranked with a ballot: emit(pos, idx) sees pos = rank of idx among the survivors in the reference's traversal order.
// All 32 lanes must call with the same arguments; returns the number of survivors.
template <class Accept, class Emit>
__device__ __forceinline__ int warp_for_each_in_area(const FrameDev& f, float x, float y, float r, int minLevel, int maxLevel, Accept&& accept,
                                                     Emit&& emit) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int nMinCellX = max(0, (int)floorf((x - f.min_x - r) * f.gw_inv));
    if (nMinCellX >= COEB_GRID_COLS) return 0;
    const int nMaxCellX = min(COEB_GRID_COLS - 1, (int)ceilf((x - f.min_x + r) * f.gw_inv));
    if (nMaxCellX < 0) return 0;
    const int nMinCellY = max(0, (int)floorf((y - f.min_y - r) * f.gh_inv));
    if (nMinCellY >= COEB_GRID_ROWS) return 0;
    const int nMaxCellY = min(COEB_GRID_ROWS - 1, (int)ceilf((y - f.min_y + r) * f.gh_inv));
    if (nMaxCellY < 0) return 0;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    const unsigned lt = (1u << lane) - 1u;
    int cnt = 0;
    for (int c0 = nMinCellX; c0 <= nMaxCellX; c0 += 32) {   // one trip unless the window is wider than 32 columns
        const int ix = c0 + lane;
        int lo = 0, hi = 0;
        if (ix <= nMaxCellX) { lo = f.cell_start[ix * COEB_GRID_ROWS + nMinCellY]; hi = f.cell_start[ix * COEB_GRID_ROWS + nMaxCellY + 1]; }
        const int n = hi - lo;
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(full, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(full, incl, 31);
        const int ncol = min(32, nMaxCellX - c0 + 1);
        for (int base = 0; base < total; base += 32) {
            const int t = base + lane;
            int col = 0;   // column of raw item t = number of columns that end at or before it
            for (int c = 0; c < ncol; c++) col += t >= __shfl_sync(full, incl, c);
            col = min(col, 31);
            const int ex = __shfl_sync(full, incl - n, col), l0 = __shfl_sync(full, lo, col);
            bool ok = t < total;
            int idx = -1;
            if (ok) {
                idx = f.cell_items[l0 + (t - ex)];
                if (bCheckLevels) {
                    const int o = f.octave[idx];
                    if (o < minLevel || (maxLevel >= 0 && o > maxLevel)) ok = false;
                }
                if (ok) {
                    const float dx = f.x[idx] - x, dy = f.y[idx] - y;
                    ok = fabsf(dx) < r && fabsf(dy) < r && accept(idx);
                }
            }
            const unsigned m = __ballot_sync(full, ok);
            if (ok) emit(cnt + __popc(m & lt), idx);
            cnt += __popc(m);
        }
    }
    return cnt;
}

// glibc 2.39 logf (sysdeps/ieee754/flt-32/e_logf.c: 16-entry table, degree-3 polynomial in double), restated so that
// MapPoint::PredictScale's `ceil(log(ratio) / mfLogScaleFactor)` (src/MapPoint.cc:402-417, float overloads) is the same
// float on the device as on the reference's host; the oracle carries the same restatement and pins it against libm.
__constant__ double kLogfTab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};

__device__ __forceinline__ float glibc_logf(float x) {   // normal positive x only (the caller guarantees it)
    const uint32_t ix = __float_as_uint(x);
    if (ix == 0x3f800000u) return 0.f;
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (tmp >> 19) & 15;
    const int k = (int)tmp >> 23;
    const double z = (double)__uint_as_float(ix - (tmp & 0xff800000u));
    const double r = z * kLogfTab[i][0] - 1, y0 = kLogfTab[i][1] + (double)k * 0x1.62e42fefa39efp-1, r2 = r * r;
    double y = 0x1.5575b0be00b6ap-2 * r + -0x1.ffffef20a4123p-2;
    y = -0x1.00ea348b88334p-2 * r2 + y;
    y = y * r2 + (y0 + r);
    return (float)y;
}

#ifdef COEB_KERNEL_TRACE
// Development timeline (tools/build_trace.sh builds): first-CTA start / last-CTA end of the matcher kernels of one call, %globaltimer ns.
// ids: 0 frame tail, 1 grid build, 2 frustum + collect, 3 m2 collect, 4 m2 resolve (marks 8..15 inside it), 5 m3 collect, 6 m3 resolve
__device__ unsigned long long g_mtrace[32];
struct MatchTrace {
    int id; unsigned long long t0;
    __device__ __forceinline__ static unsigned long long now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
    __device__ __forceinline__ MatchTrace(int i) : id(i), t0(0) { if (threadIdx.x == 0 && threadIdx.y == 0) t0 = now(); }
    __device__ __forceinline__ ~MatchTrace() { if (threadIdx.x == 0 && threadIdx.y == 0) { atomicMin(&g_mtrace[2 * id], t0); atomicMax(&g_mtrace[2 * id + 1], now()); } }
};
#define COEB_MTRACE(id) MatchTrace coeb_match_trace_(id)
#define COEB_MMARK(slot) do { if (threadIdx.x == 0) g_mtrace[16 + (slot)] = MatchTrace::now(); } while (0)
__device__ int g_mstat[16];   // resolve kernel: [0] active queries, [1 + round] queries whose deciding entries were claimed (slow path)
#define COEB_MSTAT(slot, val) do { if (threadIdx.x == 0 && (slot) < 16) g_mstat[slot] = (val); } while (0)
#else
#define COEB_MTRACE(id)
#define COEB_MMARK(slot)
#define COEB_MSTAT(slot, val)
#endif

// ---- grid build (Frame::AssignFeaturesToGrid, src/Frame.cc:396-411) -------------------------------------
// One CTA. Cells ix-major; items ascending by keypoint index (the reference pushes in index order).
__global__ void __launch_bounds__(1024) grid_build_kernel(FrameDev f, int* cell_start, int* cell_items, int* kp_cell) {
    COEB_MTRACE(1);
    __shared__ int s_cnt[kGridCells + 1];
    __shared__ int s_warp[33];
    const int tid = threadIdx.x, T = blockDim.x;
    for (int i = tid; i <= kGridCells; i += T) s_cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < f.n; i += T) {
        // PosInGrid: round() half away from zero (src/Frame.cc:560-561)
        const int px = (int)roundf((f.x[i] - f.min_x) * f.gw_inv), py = (int)roundf((f.y[i] - f.min_y) * f.gh_inv);
        int c = -1;
        if (!(px < 0 || px >= COEB_GRID_COLS || py < 0 || py >= COEB_GRID_ROWS)) {
            c = px * COEB_GRID_ROWS + py;
            atomicAdd(&s_cnt[c], 1);
        }
        kp_cell[i] = c;
    }
    __syncthreads();
    const int total = block_exclusive_scan(s_cnt, kGridCells, s_warp);
    if (tid == 0) s_cnt[kGridCells] = total;
    __syncthreads();
    for (int i = tid; i <= kGridCells; i += T) cell_start[i] = s_cnt[i];
    __syncthreads();
    // fill with atomic slots, then sort each (tiny) cell: items end up ascending by keypoint index,
    // the order in which the reference pushes them.
    for (int i = tid; i < f.n; i += T) {
        const int c = kp_cell[i];
        if (c >= 0) cell_items[atomicAdd(&s_cnt[c], 1)] = i;
    }
    __syncthreads();
    for (int c = tid; c < kGridCells; c += T) {
        const int a = cell_start[c], b = cell_start[c + 1];
        for (int i = a + 1; i < b; i++) {  // insertion sort, ascending keypoint index
            const int v = cell_items[i];
            int j = i - 1;
            while (j >= a && cell_items[j] > v) { cell_items[j + 1] = cell_items[j]; j--; }
            cell_items[j + 1] = v;
        }
    }
}

// The same grid for frames of up to kGridSmemMax keypoints with the cell table, the fill cursors and the item list in shared memory
// (one pass over the coordinates, no global round trips between the steps; see frame_tail_grid_kernel for the RGB-D constructor's form).
constexpr int kGridSmemMax = 2048;
__global__ void __launch_bounds__(1024, 1) grid_build_smem_kernel(FrameDev f, int* __restrict__ cell_start, int* __restrict__ cell_items) {
    __shared__ int s_start[kGridCells + 1], s_fill[kGridCells];
    __shared__ short s_cell[kGridSmemMax], s_items[kGridSmemMax];
    __shared__ int s_warp[33];
    const int tid = threadIdx.x, T = blockDim.x, n = f.n;
    for (int i = tid; i <= kGridCells; i += T) s_start[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += T) {
        // PosInGrid: round() half away from zero (src/Frame.cc:560-561)
        const int px = (int)roundf((f.x[i] - f.min_x) * f.gw_inv), py = (int)roundf((f.y[i] - f.min_y) * f.gh_inv);
        int c = -1;
        if (!(px < 0 || px >= COEB_GRID_COLS || py < 0 || py >= COEB_GRID_ROWS)) {
            c = px * COEB_GRID_ROWS + py;
            atomicAdd(&s_start[c], 1);
        }
        s_cell[i] = (short)c;
    }
    __syncthreads();
    const int total = block_exclusive_scan(s_start, kGridCells, s_warp);
    if (tid == 0) s_start[kGridCells] = total;
    __syncthreads();
    for (int i = tid; i <= kGridCells; i += T) {
        const int v = s_start[i];
        cell_start[i] = v;
        if (i < kGridCells) s_fill[i] = v;
    }
    __syncthreads();
    for (int i = tid; i < n; i += T) {
        const int c = s_cell[i];
        if (c >= 0) s_items[atomicAdd(&s_fill[c], 1)] = (short)i;
    }
    __syncthreads();
    for (int c = tid; c < kGridCells; c += T) {   // items of a cell in ascending keypoint index, the order in which the reference pushes them
        const int lo = s_start[c], hi = s_start[c + 1];
        for (int i = lo + 1; i < hi; i++) {
            const short v = s_items[i];
            int j = i - 1;
            while (j >= lo && s_items[j] > v) { s_items[j + 1] = s_items[j]; j--; }
            s_items[j + 1] = v;
        }
    }
    __syncthreads();
    for (int i = tid; i < total; i += T) cell_items[i] = s_items[i];
}

__global__ void features_in_area_kernel(FrameDev f, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap, int* n_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int n = 0;
    for_each_in_area(f, x, y, r, minLevel, maxLevel, [&](int idx) { if (n < cap) out[n] = idx; n++; });
    *n_out = n;
}

__global__ void hamming_pairs_kernel(const uint32_t* a, const uint32_t* b, int n, int* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = hamming256(a + 8 * (size_t)i, b + 8 * (size_t)i);
}

// =====================================================================================================================
// Windowed matchers. Each has
//   * a static visitor  xx_visit(i, fn): walks query i's grid window in the reference's traversal order, applies the
//     filters that do not depend on other queries of the call, computes the Hamming distance and calls
//     fn(idx, dist, octave) per surviving candidate;
//   * a collect kernel (one thread per query, many CTAs) that stores those candidates as per-query lists;
//   * a resolve kernel (one CTA) that iterates R <- F(R) to the fixed point over the stored lists -- or, when
//     kLists == false, by re-walking the windows (fallback used only if some list overflowed its capacity).
// =====================================================================================================================
struct CandLists {
    int2* items;   // [n][cap]: x = idx | octave << 24, y = distance
    int* count;    // [n] candidates found (may exceed cap: overflow)
    int cap;
    int* active;   // [n] queries with at least one candidate, in no particular order
    int* meta;     // [0] number of active queries, [1] some list overflowed its capacity. Zero before the collect kernel: a fixed
                   // per-matcher word pair that the (single-CTA) resolve kernel clears again once it has read it -- no memset node per call
    int* top;      // [n] list positions of the two smallest (distance, position) entries: best | second << 16 (0xFFFF = none); M2 only
    int4* rec;     // [n] M2 only, indexed like `active`: {query | kNoObsBit, best entry .x, second entry .x or -1, best distance | second
                   // distance << 9 | min(count, 255) << 18}: everything the resolve kernel caches per query, as one coalesced 16-byte load
                   // (fetched through active -> top / count / has_obs -> items, ~5 scattered sectors per query, one SM needed 4 us per 1000 queries)
};
constexpr int kNoObsBit = 0x40000000;   // query ids in the resolve kernels: Observations() == 0, its claims do not block (:87-89)

// Lane 0 of the collecting warp records the query's candidate count.
__device__ __forceinline__ void cand_finish(const CandLists& C, int i, int cnt) {
    C.count[i] = cnt;
    if (cnt > 0) C.active[atomicAdd(&C.meta[0], 1)] = i;
    if (cnt > C.cap) C.meta[1] = 1;
}

// Sequential walk of one candidate list by one thread, eight entries in flight: the entries are independent loads, but a loop
// that tests each one before fetching the next pays one memory round trip per entry (a 32-entry list cost ~10 us per round
// of the resolve kernels, which was most of their time).
template <class Fn>
__device__ __forceinline__ void walk_list(const int2* __restrict__ it, int n, Fn&& fn) {
    for (int c0 = 0; c0 < n; c0 += 8) {
        int2 e[8];
#pragma unroll
        for (int u = 0; u < 8; u++) e[u] = it[min(c0 + u, n - 1)];
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (c0 + u < n) fn(e[u]);
    }
}

// Programmatic dependent launch (sm_90+): a collect kernel lets the dependent resolve kernel be scheduled as soon as every CTA of
// the collect grid has started; the resolve kernel waits (for completion and visibility of the whole collect grid) before it reads.
// Both are no-ops for launches without the programmatic-serialisation attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- M2: SearchByProjection(Frame&, vector<MapPoint*>&, th) (src/ORBmatcher.cc:45-129) -------------------
struct MapDev {
    int n;
    const uint8_t *track_in_view, *bad, *has_obs;
    const float *proj_x, *proj_y, *proj_xr, *view_cos;
    const int* level;
    const uint32_t* desc;
};

template <class Fn>
__device__ __forceinline__ void m2_visit(const FrameDev& F, const MapDev& M, float th, const int* __restrict__ kp_state, int i, Fn&& fn) {
    if (!M.track_in_view[i] || M.bad[i]) return;
    const int lvl = M.level[i];
    float r = ((double)M.view_cos[i] > 0.998) ? 2.5f : 4.0f;  // RadiusByViewingCos (:131-137)
    if (th != 1.0f) r *= th;                                   // bFactor (:49, :65-66)
    const float rs = r * F.scale[lvl];
    const uint32_t* d = M.desc + 8 * (size_t)i;
    const float pxr = M.proj_xr[i];
    for_each_in_area(F, M.proj_x[i], M.proj_y[i], rs, lvl - 1, lvl, [&](int idx) {
        if (kp_state[idx] == -2) return;                       // already holds a MapPoint with observations (:87-89)
        if (F.uright) {
            const float ur = F.uright[idx];
            if (ur > 0) { const float er = fabsf(pxr - ur); if (er > rs) return; }   // :91-96
        }
        fn(idx, hamming256(d, F.desc + 8 * (size_t)idx), F.octave[idx]);
    });
}

// Static part of query i of SearchByProjection(F, MapPoints): window and filters that do not depend on other queries.
struct M2Query { float x, y, rs, pxr; int lvl; const uint32_t* d; };
__device__ __forceinline__ bool m2_query(const FrameDev& F, const MapDev& M, float th, int i, M2Query& q) {
    if (!M.track_in_view[i] || M.bad[i]) return false;
    q.lvl = M.level[i];
    float r = ((double)M.view_cos[i] > 0.998) ? 2.5f : 4.0f;  // RadiusByViewingCos (:131-137)
    if (th != 1.0f) r *= th;                                   // bFactor (:49, :65-66)
    q.rs = r * F.scale[q.lvl];
    q.d = M.desc + 8 * (size_t)i;
    q.pxr = M.proj_xr[i];
    q.x = M.proj_x[i]; q.y = M.proj_y[i];
    return true;
}
// One warp per map point: the window's keypoints are tested 32 at a time, the survivors written in traversal order.
// `valid` / `q`: the query's static part (m2_query), `obs`: Observations() > 0.
__device__ __forceinline__ void m2_collect_query(const FrameDev& F, const int* __restrict__ kp_state, const CandLists& C, int i, bool valid, const M2Query& q,
                                                 bool obs) {
    int cnt = 0;
    unsigned k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;   // this lane's two smallest distance << 16 | list position
    int k1x = -1, k2x = -1;                        // ... and their entries' .x (keypoint | octave << 24)
    if (valid) {
        int2* out = C.items + (size_t)i * C.cap;
        cnt = warp_for_each_in_area(
            F, q.x, q.y, q.rs, q.lvl - 1, q.lvl,
            [&](int idx) {
                if (kp_state[idx] == -2) return false;                 // already holds a MapPoint with observations (:87-89)
                if (F.uright) {
                    const float ur = F.uright[idx];
                    if (ur > 0 && fabsf(q.pxr - ur) > q.rs) return false;   // :91-96
                }
                return true;
            },
            [&](int pos, int idx) {
                if (pos < C.cap) {
                    const int dist = hamming256(q.d, F.desc + 8 * (size_t)idx);
                    const int ex = idx | (F.octave[idx] << 24);
                    out[pos] = make_int2(ex, dist);
                    const unsigned key = ((unsigned)dist << 16) | (unsigned)pos;
                    if (key < k1) { k2 = k1; k2x = k1x; k1 = key; k1x = ex; } else if (key < k2) { k2 = key; k2x = ex; }
                }
            });
    }
    // The sequential best / second-best bookkeeping of :100-113 ends with the two smallest entries in (distance, traversal position)
    // order, whatever the order of arrival: the resolve kernel starts from them and only walks a list when a claim removes one.
    unsigned b1 = k1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b1 = min(b1, __shfl_xor_sync(0xffffffffu, b1, o));
    const unsigned c2 = (k1 == b1) ? k2 : k1;
    const int c2x = (k1 == b1) ? k2x : k1x;
    unsigned b2 = c2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b2 = min(b2, __shfl_xor_sync(0xffffffffu, b2, o));
    // keys carry their list position, so each winner sits in exactly one lane
    const int x1 = __shfl_sync(0xffffffffu, k1x, __ffs(__ballot_sync(0xffffffffu, k1 == b1)) - 1);
    const int x2 = __shfl_sync(0xffffffffu, c2x, __ffs(__ballot_sync(0xffffffffu, c2 == b2)) - 1);
    if ((threadIdx.x & 31) == 0) {
        if (C.top) C.top[i] = (int)((b1 & 0xFFFFu) | ((b2 == 0xFFFFFFFFu ? 0xFFFFu : (b2 & 0xFFFFu)) << 16));
        C.count[i] = cnt;
        // A query whose smallest distance exceeds TH_HIGH can never match, whatever the other queries claim (exclusions only raise
        // its best distance, :115), and so never claims either: it is not handed to the resolve kernel at all. (On the benchmark's
        // map four of five queries with candidates are of this kind, and the resolve kernel's warps ran 5.6 live lanes of 32.)
        if (cnt > 0 && (b1 >> 16) <= (unsigned)COEB_TH_HIGH) {
            const int a = atomicAdd(&C.meta[0], 1);
            C.active[a] = i;
            if (C.rec) {
                const int d1 = (int)min(b1 >> 16, 256u), d2 = (int)min(b2 >> 16, 256u);   // (no key: 0xFFFF -> 256, "none")
                C.rec[a] = make_int4(obs ? i : (i | kNoObsBit), x1, d2 < 256 ? x2 : -1, d1 | (d2 << 9) | (min(cnt, 255) << 18));
            }
        }
        if (cnt > C.cap) C.meta[1] = 1;
    }
}
__device__ __forceinline__ void m2_collect_warp(const FrameDev& F, const MapDev& M, float th, const int* __restrict__ kp_state, const CandLists& C, int i) {
    M2Query q;
    const bool obs = M.has_obs[i] != 0;
    const bool valid = m2_query(F, M, th, i, q);
    m2_collect_query(F, kp_state, C, i, valid, q, obs);
}

__global__ void __launch_bounds__(256) m2_collect_kernel(FrameDev F, MapDev M, float th, const int* kp_state, CandLists C) {
    COEB_MTRACE(3);
    pdl_launch_dependents();   // the resolve kernel may take its SM and run its prologue now; it waits for this grid before reading
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= M.n) return;
    m2_collect_warp(F, M, th, kp_state, C, i);
}

// One CTA iterates to the fixed point. res[i] = keypoint claimed by map point i or -1.
// claim_min[idx] = smallest i (with Observations()>0) that currently claims idx.
// The claim table (one int per keypoint of the frame) lives in shared memory when it fits: dynamic shared memory = F.n ints, else 0
// and the global scratch array is used.
extern __shared__ __align__(16) int s_claim_dyn[];
// kSmem (the claim table's address space) is a template argument: behind a run-time choice of pointer every access to the table is a
// GENERIC load / store / atomic (ATOM.E.MIN on a shared-memory address), several times the cost of LDS / ATOMS.
template <bool kLists, bool kSmem>
__global__ void __launch_bounds__(1024, 1) m2_resolve_kernel(FrameDev F, MapDev M, float th, float nnratio, const int* kp_state, CandLists C,
                                                          int* kp_match /*out: every entry written*/, int* res, int* claim_glob, int* out_info,
                                                          int cache_cap /*active queries the dynamic shared memory can hold*/) {
    constexpr bool claim_in_smem = kSmem;
    COEB_MTRACE(4);
    __shared__ int s_changed, s_count, s_nslow;
    constexpr int kSlowCap = 1024;
    __shared__ int s_slow[kSlowCap];
    const int tid = threadIdx.x, T = blockDim.x;
    int* const claim_min = kSmem ? s_claim_dyn : claim_glob;
    const int na = kLists ? C.meta[0] : M.n;
    const bool overflow = kLists && C.meta[1];
    if (kLists) {   // every thread has read the two words: cleared for the next call's collect kernel
        __syncthreads();
        if (tid == 0) { C.meta[0] = 0; C.meta[1] = 0; }
    }
    if (overflow) {   // a list overflowed: report and let the host rerun the window-walking variant
        if (tid == 0) { out_info[0] = 0; out_info[1] = 0; out_info[2] = 1; }
        return;
    }
    // Only queries with candidates can match; with lists they were compacted by the collect kernel (in no particular
    // order: the evaluation of a query depends on the others through claim_min only).
    auto query = [&](int a) { return kLists ? C.active[a] : a; };
    if (kLists && na <= cache_cap) {
        // Cached path: the active queries and, for each, the two list entries that decide the outcome when no claim touches them
        // (C.top) are copied to shared memory once, so that a round of the fixed-point iteration reads shared memory only and the loops
        // stay rolled (the kernel runs on one SM: an unrolled register version was bound by instruction fetch). A query one of whose
        // two entries is claimed by an earlier map point is queued and its list re-examined by a whole warp.
        int* const qi = s_claim_dyn + (claim_in_smem ? F.n : 0);
        int* const qres = qi + cache_cap;
        int2* const q1 = reinterpret_cast<int2*>(qres + cache_cap + (((claim_in_smem ? F.n : 0) + 2 * cache_cap) & 1));   // 8-byte aligned
        int2* const q2 = q1 + cache_cap;
        unsigned char* const qcnt = reinterpret_cast<unsigned char*>(q2 + cache_cap);   // list lengths (at most C.cap <= 64)
        const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
        for (int a = tid; a < na; a += T) {
            const int i = C.active[a];
            const int2* it = C.items + (size_t)i * C.cap;
            const unsigned tp = (unsigned)C.top[i];
            const int2 e1 = it[tp & 0xFFFFu];
            int2 e2 = make_int2(-1, 256);
            if ((tp >> 16) != 0xFFFFu) { e2 = it[tp >> 16]; if (e2.y >= 256) e2 = make_int2(-1, 256); }   // (a distance of 256 never enters, :100)
            qi[a] = M.has_obs[i] ? i : (i | (int)0x40000000);   // bit 30: Observations() == 0, its claims do not block (:87-89)
            qres[a] = -1;
            q1[a] = e1;
            q2[a] = e2;
            qcnt[a] = (unsigned char)min(C.count[i], 255);
        }
        auto decide = [&](int bestDist, int bestLevel, int bestIdx, int bestDist2, int bestLevel2) {
            return (bestDist <= COEB_TH_HIGH && !(bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2)) ? bestIdx : -1;
        };
        COEB_MMARK(0);   // active queries cached
        COEB_MSTAT(0, na);
        for (int round = 0; round <= M.n; round++) {
            if (round < 6) COEB_MMARK(1 + round);
            for (int k = tid; k < F.n; k += T) claim_min[k] = kInf;
            if (tid == 0) { s_changed = 0; s_nslow = 0; }
            __syncthreads();
            for (int a = tid; a < na; a += T) {
                const int r = qres[a], iq = qi[a];
                if (r >= 0 && !(iq & 0x40000000)) atomicMin(&claim_min[r], iq);
            }
            __syncthreads();
            bool changed_here = false;
            for (int a = tid; a < na; a += T) {
                const int i = qi[a] & 0x3FFFFFFF;
                const int2 e1 = q1[a], e2 = q2[a];
                int out;
                if (e1.y > COEB_TH_HIGH) {
                    out = -1;   // exclusions can only raise the best distance
                } else if (claim_min[e1.x & 0xFFFFFF] >= i && (e2.x < 0 || claim_min[e2.x & 0xFFFFFF] >= i)) {
                    out = decide(e1.y, e1.x >> 24, e1.x & 0xFFFFFF, e2.y, e2.x < 0 ? -1 : (e2.x >> 24));   // every other exclusion leaves these two in place
                } else {
                    const int k = atomicAdd(&s_nslow, 1);
                    if (k < kSlowCap) { s_slow[k] = a; continue; }
                    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;   // queue full: walked in place
                    walk_list(C.items + (size_t)i * C.cap, C.count[i], [&](const int2 e) {
                        const int idx = e.x & 0xFFFFFF, dist = e.y, oct = e.x >> 24;
                        if (claim_min[idx] < i) return;   // claimed earlier in this call by a MapPoint with observations (:87-89, :123)
                        if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = oct; bestIdx = idx; }
                        else if (dist < bestDist2) { bestLevel2 = oct; bestDist2 = dist; }
                    });
                    out = decide(bestDist, bestLevel, bestIdx, bestDist2, bestLevel2);
                }
                if (out != qres[a]) { qres[a] = out; changed_here = true; }
            }
            __syncthreads();
            {   // one warp per queued query: its entries are tested side by side and the two smallest (distance, position) keys among
                // those no earlier map point claims are found with two warp reductions. A lane holds entries lane and lane + 32 (lists have
                // at most 64 entries); the lists of a warp's next query are already in flight while the current one is reduced.
                const int ns = min(s_nslow, kSlowCap);
                COEB_MSTAT(1 + round, s_nslow);
                auto fetch = [&](int k, int& i, int& n, int2& ea, int2& eb) {
                    i = -1; n = 0; ea = eb = make_int2(0, 256);
                    if (k < ns) {
                        const int a = s_slow[k];
                        i = qi[a] & 0x3FFFFFFF;
                        n = qcnt[a];
                        const int2* it = C.items + (size_t)i * C.cap;
                        if (lane < n) ea = it[lane];
                        if (lane + 32 < n) eb = it[lane + 32];
                    }
                };
                int i, n, i_nx, n_nx;
                int2 ea, eb, ea_nx, eb_nx;
                fetch(warp, i, n, ea, eb);
                for (int k = warp; k < ns; k += nwarps) {
                    fetch(k + nwarps, i_nx, n_nx, ea_nx, eb_nx);
                    const bool oka = lane < n && ea.y < 256 && claim_min[ea.x & 0xFFFFFF] >= i;
                    const bool okb = lane + 32 < n && eb.y < 256 && claim_min[eb.x & 0xFFFFFF] >= i;
                    const unsigned ka = oka ? (((unsigned)ea.y << 16) | (unsigned)lane) : 0xFFFFFFFFu;
                    const unsigned kb = okb ? (((unsigned)eb.y << 16) | (unsigned)(lane + 32)) : 0xFFFFFFFFu;
                    const unsigned k1 = min(ka, kb), k2 = max(ka, kb);
                    unsigned b1 = k1;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) b1 = min(b1, __shfl_xor_sync(0xffffffffu, b1, o));
                    unsigned b2 = (k1 == b1) ? k2 : k1;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) b2 = min(b2, __shfl_xor_sync(0xffffffffu, b2, o));
                    // the winners' entries come from the lanes that hold them
                    const int l1 = (int)(b1 & 31u), l2 = (int)(b2 & 31u);
                    const int x1a = __shfl_sync(0xffffffffu, ea.x, l1), x1b = __shfl_sync(0xffffffffu, eb.x, l1);
                    const int x2a = __shfl_sync(0xffffffffu, ea.x, l2), x2b = __shfl_sync(0xffffffffu, eb.x, l2);
                    if (lane == 0) {
                        int out = -1;
                        if (b1 != 0xFFFFFFFFu) {
                            const int x1 = (b1 & 32u) ? x1b : x1a;
                            int d2 = 256, lv2 = -1;
                            if (b2 != 0xFFFFFFFFu) { d2 = (int)(b2 >> 16); lv2 = ((b2 & 32u) ? x2b : x2a) >> 24; }
                            out = decide((int)(b1 >> 16), x1 >> 24, x1 & 0xFFFFFF, d2, lv2);
                        }
                        const int a = s_slow[k];
                        if (out != qres[a]) { qres[a] = out; s_changed = 1; }
                    }
                    i = i_nx; n = n_nx; ea = ea_nx; eb = eb_nx;
                }
            }
            if (changed_here) s_changed = 1;
            __syncthreads();
            const int changed = s_changed;
            __syncthreads();
            if (!changed) { if (tid == 0) out_info[1] = round + 1; break; }
        }
        // F.mvpMapPoints[bestIdx] = pMP in query order: the last writer wins; every success counts (:123-124)
        if (tid == 0) s_count = 0;
        for (int k = tid; k < F.n; k += T) claim_min[k] = -1;  // reuse as "last claimant"
        __syncthreads();
        int mine = 0;
        for (int a = tid; a < na; a += T)
            if (qres[a] >= 0) { atomicMax(&claim_min[qres[a]], qi[a] & 0x3FFFFFFF); mine++; }
        if (mine) atomicAdd(&s_count, mine);
        __syncthreads();
        for (int k = tid; k < F.n; k += T)
            kp_match[k] = claim_min[k] >= 0 ? claim_min[k] : kp_state[k];   // untouched entries keep the caller's state
        if (tid == 0) { out_info[0] = s_count; out_info[2] = 0; }
        return;
    }
    for (int a = tid; a < na; a += T) res[query(a)] = -1;
    for (int round = 0; round <= M.n; round++) {
        for (int k = tid; k < F.n; k += T) claim_min[k] = kInf;
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int a = tid; a < na; a += T) {
            const int i = query(a), k = res[i];
            if (k >= 0 && M.has_obs[i]) atomicMin(&claim_min[k], i);
        }
        __syncthreads();
        for (int a = tid; a < na; a += T) {
            const int i = query(a);
            int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
            auto consider = [&](int idx, int dist, int oct) {
                if (claim_min[idx] < i) return;   // claimed earlier in this call by a MapPoint with observations (:87-89, :123)
                if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = oct; bestIdx = idx; }
                else if (dist < bestDist2) { bestLevel2 = oct; bestDist2 = dist; }
            };
            if (kLists) {
                walk_list(C.items + (size_t)i * C.cap, C.count[i], [&](const int2 e) { consider(e.x & 0xFFFFFF, e.y, e.x >> 24); });
            } else {
                m2_visit(F, M, th, kp_state, i, consider);
            }
            int out = -1;
            if (bestDist <= COEB_TH_HIGH && !(bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2)) out = bestIdx;
            if (out != res[i]) { res[i] = out; s_changed = 1; }
        }
        __syncthreads();
        const int changed = s_changed;
        __syncthreads();
        if (!changed) { if (tid == 0) out_info[1] = round + 1; break; }
    }
    // F.mvpMapPoints[bestIdx] = pMP in query order: the last writer wins; every success counts (:123-124)
    if (tid == 0) s_count = 0;
    for (int k = tid; k < F.n; k += T) claim_min[k] = -1;  // reuse as "last claimant"
    __syncthreads();
    int mine = 0;
    for (int a = tid; a < na; a += T) {
        const int i = query(a);
        if (res[i] >= 0) { atomicMax(&claim_min[res[i]], i); mine++; }
    }
    if (mine) atomicAdd(&s_count, mine);
    __syncthreads();
    for (int k = tid; k < F.n; k += T)
        kp_match[k] = claim_min[k] >= 0 ? claim_min[k] : kp_state[k];   // untouched entries keep the caller's state
    if (tid == 0) { out_info[0] = s_count; out_info[2] = 0; }
}

// Cached form of the fixed-point iteration above for maps whose active queries all fit in shared memory (the host checks that
// M.n <= cache_cap before choosing it), re-cut for the latency of ONE call (SearchLocalPoints runs once per tracked frame). The
// kernel is one CTA on one SM, and what bounds it there is that SM's issue rate (32 warps, 4 instructions per clock), not latency:
//   * launched with programmatic stream serialisation: the CTA is resident and has initialised its tables while the collect
//     kernel still runs (griddepcontrol.wait orders every read of the collect kernel's output);
//   * round 0 (no claims yet: every query decides from its two smallest entries) is fused into the caching pass;
//   * three claim tables rotate (read X = claims of R_t, write Y = claims of R_t+1 as each query decides, reset Z), so a round
//     is two barriers instead of five: decide | claimed queries | __syncthreads_or(changed);
//   * a query whose deciding entries were claimed once stays queued (its evaluation there is exact whether or not the claim
//     persists), so its queue position is stable: a list of up to kSl entries (nearly all of them: a window holds a handful of
//     keypoints) is copied to shared memory once and walked by ONE thread per round in the reference's order -- a warp per
//     query spent ~100 instructions per query and round on reductions over mostly empty lanes, 3 of a round's 4.3 us.
//     Longer lists keep the warp-per-query form.
constexpr int kSl = 8;   // entries of a short list
// kMode 2: SearchByProjection(F, MapPoints) -- best and second best with the level / ratio test (:115-121).
// kMode 3: SearchByProjection(cur, last) and its relocalisation overload -- the best unclaimed candidate within `accept` (:1421-1429),
//          then the rotation-consistency histogram over the accepted matches (:1431-1468); records carry no second entry.
__device__ __forceinline__ int rot_bin(float a1, float a2);
__device__ void three_maxima(const int* histo, int L, int& ind1, int& ind2, int& ind3);
struct RotArgs { int check_ori; const float* q_angle; const float* k_angle; };   // kMode 3: angles of the queries / of the frame's keypoints
template <bool kSmem, int kMode>
__global__ void __launch_bounds__(1024, 1) m2_resolve_cached_kernel(int Fn, int Mn, float nnratio, int accept, const int* __restrict__ kp_state, CandLists C,
                                                                 int* kp_match, int* claim_glob, int* out_info, int cache_cap, int slow_cap,
                                                                 const uint4* __restrict__ flags_dev, uint4* flags_host, int flags_vec, RotArgs rot) {
    constexpr bool claim_in_smem = kSmem;
    COEB_MTRACE(4);
    constexpr int kLongCap = 256, kSlowBit = 0x20000000, kIdMask = 0x1FFFFFFF;
    __shared__ int s_nslow, s_nlong, s_count;
    __shared__ int s_slow[1024], s_long[kLongCap];
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    int* const tab = kSmem ? s_claim_dyn : claim_glob;   // three claim tables of Fn ints
    const int tab_ints = claim_in_smem ? 3 * Fn : 0;
    int* const qi = s_claim_dyn + tab_ints;
    int* const qres = qi + cache_cap;
    int2* const q1 = reinterpret_cast<int2*>(qres + cache_cap + ((tab_ints + 2 * cache_cap) & 1));   // 8-byte aligned
    int2* const q2 = q1 + cache_cap;
    unsigned char* const qcnt = reinterpret_cast<unsigned char*>(q2 + cache_cap);   // list lengths (at most C.cap <= 64)
    int2* const sl = reinterpret_cast<int2*>(qcnt + ((cache_cap + 7) & ~7));        // [kSl][slow_cap]: entry u of queued query k
    slow_cap = min(slow_cap, 1024);
    for (int k = tid; k < 3 * Fn; k += T) tab[k] = kInf;
    if (tid == 0) { s_nslow = 0; s_nlong = 0; }
    COEB_MMARK(12);
    pdl_wait();   // the collect kernel has finished and its lists are visible
    COEB_MMARK(13);
    // SearchLocalPoints: the visibility flags the collect kernel left in device memory travel to the caller's (mapped) array from
    // here, as 16-byte posted writes that overlap the rounds below. Written by the collect kernel itself they were 5000 one-byte
    // writes across PCIe that its completion, and so this kernel's start, waited 3-4 us for.
    for (int k = tid; k < flags_vec; k += T) flags_host[k] = flags_dev[k];
    const int na = C.meta[0];
    const bool overflow = C.meta[1] != 0 || na > cache_cap;
    __syncthreads();   // every thread has read the two words: cleared for the next call's collect kernel
    COEB_MMARK(14);
    if (tid == 0) { C.meta[0] = 0; C.meta[1] = 0; }
    if (overflow) {   // a list overflowed: report and let the host rerun the window-walking variant
        if (tid == 0) { out_info[0] = 0; out_info[1] = 0; out_info[2] = 1; }
        return;
    }
    auto decide = [&](int bestDist, int bestLevel, int bestIdx, int bestDist2, int bestLevel2) {
        if (kMode == 3) return bestDist <= accept ? bestIdx : -1;
        return (bestDist <= accept && !(bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2)) ? bestIdx : -1;
    };
    // caching pass + round 0: R_1 = F(nothing claimed); its claims go to table 0
    for (int a = tid; a < na; a += T) {
        const int4 r = C.rec[a];
        const int iq = r.x, d2 = (r.w >> 9) & 511;
        const int2 e1 = make_int2(r.y, r.w & 511);
        const int2 e2 = (r.z < 0 || d2 >= 256) ? make_int2(-1, 256) : make_int2(r.z, d2);   // (a distance of 256 never enters, :100)
        const int out = e1.y <= accept ? decide(e1.y, e1.x >> 24, e1.x & 0xFFFFFF, e2.y, e2.x < 0 ? -1 : (e2.x >> 24)) : -1;
        qi[a] = iq;
        qres[a] = out;
        q1[a] = e1;
        q2[a] = e2;
        qcnt[a] = (unsigned char)(r.w >> 18);
        if (out >= 0 && !(iq & kNoObsBit)) atomicMin(&tab[out], iq & kIdMask);
    }
    COEB_MMARK(0);   // active queries cached, round 0 done
    COEB_MSTAT(0, na);
#ifdef COEB_KERNEL_TRACE
    const long long trace_c0 = clock64();
#endif
    __syncthreads();
    int ns_loaded = 0, round = 1;
    for (;; round++) {
        if (round < 7) COEB_MMARK(round);
        const int* const X = tab + ((round - 1) % 3) * Fn;
        int* const Y = tab + (round % 3) * Fn;
        int* const Z = tab + ((round + 1) % 3) * Fn;
        bool changed_here = false;
        auto settle = [&](int a, int iq, int out) {   // R_t+1 of query a, and its claim on behalf of the next round
            if (out != qres[a]) { qres[a] = out; changed_here = true; }
            if (out >= 0 && !(iq & kNoObsBit)) atomicMin(&Y[out], iq & kIdMask);
        };
        for (int k = tid; k < Fn; k += T) Z[k] = kInf;   // read last in the previous round, written next in the following one
        if (round == 2) COEB_MMARK(8);
        for (int a = tid; a < na; a += T) {
            const int iq = qi[a];
            if (iq & kSlowBit) continue;
            const int2 e1 = q1[a];
            if (e1.y > accept) continue;   // exclusions can only raise the best distance: stays -1
            const int i = iq & kIdMask;
            const int2 e2 = q2[a];
            int out;
            const int c1 = X[e1.x & 0xFFFFFF], c2 = X[max(e2.x, 0) & 0xFFFFFF];   // both look-ups in flight
            if (c1 >= i && (e2.x < 0 || c2 >= i)) {
                out = decide(e1.y, e1.x >> 24, e1.x & 0xFFFFFF, e2.y, e2.x < 0 ? -1 : (e2.x >> 24));   // every other exclusion leaves these two in place
            } else {
                if ((int)qcnt[a] <= kSl) {
                    const int k = atomicAdd(&s_nslow, 1);
                    if (k < slow_cap) { s_slow[k] = a; qi[a] = iq | kSlowBit; continue; }
                } else {
                    const int k = atomicAdd(&s_nlong, 1);
                    if (k < kLongCap) { s_long[k] = a; qi[a] = iq | kSlowBit; continue; }
                }
                int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;   // queue full: walked in place
                walk_list(C.items + (size_t)i * C.cap, C.count[i], [&](const int2 e) {
                    const int idx = e.x & 0xFFFFFF, dist = e.y, oct = e.x >> 24;
                    if (X[idx] < i) return;   // claimed earlier in this call by a MapPoint with observations (:87-89, :123)
                    if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = oct; bestIdx = idx; }
                    else if (dist < bestDist2) { bestLevel2 = oct; bestDist2 = dist; }
                });
                out = decide(bestDist, bestLevel, bestIdx, bestDist2, bestLevel2);
            }
            settle(a, iq, out);
        }
        if (round == 2) COEB_MMARK(9);
        __syncthreads();
        if (round == 2) COEB_MMARK(10);
        const int ns = min(s_nslow, slow_cap), nl = min(s_nlong, kLongCap);
        COEB_MSTAT(round, s_nslow + s_nlong);
        // short lists: one thread per queued query walks its entries in list order, exactly the loop of :84-113
        for (int k = tid; k < ns; k += T) {
            const int a = s_slow[k];
            const int iq = qi[a], i = iq & kIdMask, n = qcnt[a];
            int2 e[kSl];
            if (k >= ns_loaded) {   // queued in this round: fetch the list (all entries in flight) and keep it
                const int2* it = C.items + (size_t)i * C.cap;
#pragma unroll
                for (int u = 0; u < kSl; u++) e[u] = it[min(u, n - 1)];
#pragma unroll
                for (int u = 0; u < kSl; u++) sl[u * slow_cap + k] = e[u];
            } else {
#pragma unroll
                for (int u = 0; u < kSl; u++) e[u] = sl[u * slow_cap + k];
            }
            // the walk of :84-113 without branches (the claim look-ups are independent loads, the bookkeeping a chain of selects; the
            // branchy form cost a dependent look-up and two branch resolutions per entry, and 24 of the 32 warps wait for this loop)
            int claim[kSl];
#pragma unroll
            for (int u = 0; u < kSl; u++) claim[u] = X[e[u].x & 0xFFFFFF];
            int bestDist = 256, bestX = -1, bestDist2 = 256, bestLevel2 = -1;   // bestX: keypoint | level << 24 of the best entry
#pragma unroll
            for (int u = 0; u < kSl; u++) {
                const int dist = (u < n && claim[u] >= i) ? e[u].y : 256;   // an entry that is skipped behaves like distance 256: never enters
                const bool lt1 = dist < bestDist, lt2 = dist < bestDist2;
                bestLevel2 = lt1 ? (bestX >> 24) : (lt2 ? (e[u].x >> 24) : bestLevel2);
                bestDist2 = lt1 ? bestDist : (lt2 ? dist : bestDist2);
                bestX = lt1 ? e[u].x : bestX;
                bestDist = lt1 ? dist : bestDist;
            }
            settle(a, iq, decide(bestDist, bestX >> 24, bestX < 0 ? -1 : (bestX & 0xFFFFFF), bestDist2, bestLevel2));
        }
        ns_loaded = ns;
        if (round == 2) COEB_MMARK(11);
        if (nl > 0) {   // long lists: one warp per queued query; its entries are tested side by side and the two smallest (distance, position)
                        // keys among those no earlier map point claims are found with two warp reductions. A lane holds entries lane and lane + 32.
            auto fetch = [&](int k, int& a, int2& ea, int2& eb) {   // the next query's list is in flight while the current one is reduced
                a = 0; ea = eb = make_int2(0, 256);
                if (k < nl) {
                    a = s_long[k];
                    const int n = qcnt[a];
                    const int2* it = C.items + (size_t)(qi[a] & kIdMask) * C.cap;
                    if (lane < n) ea = it[lane];
                    if (lane + 32 < n) eb = it[lane + 32];
                }
            };
            int a, a_nx;
            int2 ea, eb, ea_nx, eb_nx;
            fetch(warp, a, ea, eb);
            for (int k = warp; k < nl; k += nwarps) {
                fetch(k + nwarps, a_nx, ea_nx, eb_nx);
                const int iq = qi[a], i = iq & kIdMask, n = qcnt[a];
                const bool oka = lane < n && ea.y < 256 && X[ea.x & 0xFFFFFF] >= i;
                const bool okb = lane + 32 < n && eb.y < 256 && X[eb.x & 0xFFFFFF] >= i;
                const unsigned ka = oka ? (((unsigned)ea.y << 16) | (unsigned)lane) : 0xFFFFFFFFu;
                const unsigned kb = okb ? (((unsigned)eb.y << 16) | (unsigned)(lane + 32)) : 0xFFFFFFFFu;
                const unsigned k1 = min(ka, kb), k2 = max(ka, kb);
                const unsigned b1 = __reduce_min_sync(0xffffffffu, k1);
                const unsigned b2 = __reduce_min_sync(0xffffffffu, (k1 == b1) ? k2 : k1);
                const int l1 = (int)(b1 & 31u), l2 = (int)(b2 & 31u);   // the winners' entries come from the lanes that hold them
                const int x1a = __shfl_sync(0xffffffffu, ea.x, l1), x1b = __shfl_sync(0xffffffffu, eb.x, l1);
                const int x2a = __shfl_sync(0xffffffffu, ea.x, l2), x2b = __shfl_sync(0xffffffffu, eb.x, l2);
                if (lane == 0) {
                    int out = -1;
                    if (b1 != 0xFFFFFFFFu) {
                        const int x1 = (b1 & 32u) ? x1b : x1a;
                        int d2 = 256, lv2 = -1;
                        if (b2 != 0xFFFFFFFFu) { d2 = (int)(b2 >> 16); lv2 = ((b2 & 32u) ? x2b : x2a) >> 24; }
                        out = decide((int)(b1 >> 16), x1 >> 24, x1 & 0xFFFFFF, d2, lv2);
                    }
                    settle(a, iq, out);
                }
                a = a_nx; ea = ea_nx; eb = eb_nx;
            }
        }
        if (round == 2) COEB_MMARK(15);
        if (!__syncthreads_or(changed_here) || round > Mn) break;
    }
#ifdef COEB_KERNEL_TRACE
    COEB_MMARK(7);   // rounds done
    COEB_MSTAT(12, (int)(clock64() - trace_c0));   // SM cycles between mark 0 and mark 7
#endif
    // F.mvpMapPoints[bestIdx] = pMP in query order: the last writer wins; every success counts (:123-124, :1429-1430)
    __shared__ int s_hist[COEB_HISTO_LENGTH];
    __shared__ int s_ind[3];
    if (tid == 0) s_count = 0;
    if (kMode == 3) for (int k = tid; k < COEB_HISTO_LENGTH; k += T) s_hist[k] = 0;
    for (int k = tid; k < Fn; k += T) tab[k] = -1;   // "last claimant"
    __syncthreads();
    int mine = 0;
    for (int a = tid; a < na; a += T)
        if (qres[a] >= 0) {
            const int i = qi[a] & kIdMask;
            atomicMax(&tab[qres[a]], i);
            mine++;
            if (kMode == 3 && rot.check_ori) atomicAdd(&s_hist[rot_bin(rot.q_angle[i], rot.k_angle[qres[a]])], 1);
        }
    if (mine) atomicAdd(&s_count, mine);
    __syncthreads();
    for (int k = tid; k < Fn; k += T)
        kp_match[k] = tab[k] >= 0 ? tab[k] : kp_state[k];   // untouched entries keep the caller's state
    if (kMode == 3 && rot.check_ori) {   // rotation consistency (:1449-1468): entries of the losing bins are cleared, each decrements
        if (tid == 0) {
            int a, b, c;
            three_maxima(s_hist, COEB_HISTO_LENGTH, a, b, c);
            s_ind[0] = a; s_ind[1] = b; s_ind[2] = c;
        }
        __syncthreads();   // (also orders the kp_match writes above before the clearing writes below)
        int removed = 0;
        for (int a = tid; a < na; a += T)
            if (qres[a] >= 0) {
                const int bin = rot_bin(rot.q_angle[qi[a] & kIdMask], rot.k_angle[qres[a]]);
                if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) { kp_match[qres[a]] = -1; removed++; }
            }
        if (removed) atomicSub(&s_count, removed);
        __syncthreads();
    }
    if (tid == 0) { out_info[0] = s_count; out_info[1] = round + 1; out_info[2] = 0; }
}

// ---- M5: ComputeThreeMaxima (src/ORBmatcher.cc:1602-1643) ------------------------------------------------------
__device__ void three_maxima(const int* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = histo[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
}

__device__ __forceinline__ int rot_bin(float a1, float a2) {  // :1434-1439
    const float factor = 1.0f / COEB_HISTO_LENGTH;
    float rot = a1 - a2;
    if (rot < 0.0f) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == COEB_HISTO_LENGTH) bin = 0;
    return bin;
}

// ---- M3: SearchByProjection(Frame& cur, const Frame& last, th, bMono) (src/ORBmatcher.cc:1329-1471) ------------
struct LastDev {
    int n;
    const uint8_t *valid, *has_obs;
    const float* xyz;
    const int* octave;
    const float* angle;
    const uint32_t* desc;
    float T[12];       // Tcw of the current frame, 3x4 row-major
    int forward, backward;
    int max_accept;    // TH_HIGH for the frame-to-frame overload, ORBdist for the relocalisation overload
    // relocalisation overload (src/ORBmatcher.cc:1473-1600): the level comes from MapPoint::PredictScale, points behind the camera are
    // not rejected, every claim blocks, no stereo test
    int reloc, nlevels;
    const float *min_dist, *max_dist;
    float Ow[3];
};

// Static part of query i of SearchByProjection(cur, last) (:1356-1391) and of the relocalisation overload (:1490-1530):
// projection, window and level range. fp32 left to right (no FMA: the file is built with -fmad=false).
struct M3Query { float u, v, radius, ur_proj; int minL, maxL; const uint32_t* d; };
__device__ __forceinline__ bool m3_query(const FrameDev& C, const LastDev& L, float th, int i, M3Query& q) {
    if (!L.valid[i]) return false;
    const float X = L.xyz[3 * i], Y = L.xyz[3 * i + 1], Z = L.xyz[3 * i + 2];
    const float xc = L.T[0] * X + L.T[1] * Y + L.T[2] * Z + L.T[3];
    const float yc = L.T[4] * X + L.T[5] * Y + L.T[6] * Z + L.T[7];
    const float zc = L.T[8] * X + L.T[9] * Y + L.T[10] * Z + L.T[11];
    const float invzc = (float)(1.0 / (double)zc);
    if (!L.reloc && invzc < 0) return false;   // :1368-1369; the relocalisation overload has no such test
    q.u = C.fx * xc * invzc + C.cx;
    q.v = C.fy * yc * invzc + C.cy;
    if (q.u < C.min_x || q.u > C.max_x) return false;
    if (q.v < C.min_y || q.v > C.max_y) return false;
    int oct;
    if (L.reloc) {
        const float ox = X - L.Ow[0], oy = Y - L.Ow[1], oz = Z - L.Ow[2];
        const double s2 = __dadd_rn(__dadd_rn(__dmul_rn((double)ox, (double)ox), __dmul_rn((double)oy, (double)oy)), __dmul_rn((double)oz, (double)oz));
        const float dist3D = (float)sqrt(s2);   // cv::norm: double accumulation
        if (dist3D < 0.8f * L.min_dist[i] || dist3D > 1.2f * L.max_dist[i]) return false;   // :1514-1519
        const float ratio = L.max_dist[i] / dist3D;   // MapPoint::PredictScale (src/MapPoint.cc:402-417)
        int nScale = 0;
        if (ratio >= 1.17549435e-38f && ratio <= 3.402823466e+38f) nScale = (int)ceilf(glibc_logf(ratio) / (L.nlevels > 1 ? glibc_logf(C.scale[1]) : 1.f));
        oct = nScale < 0 ? 0 : (nScale >= L.nlevels ? L.nlevels - 1 : nScale);
    } else {
        oct = L.octave[i];
    }
    q.radius = th * C.scale[oct];
    if (L.forward) { q.minL = oct; q.maxL = -1; }           // :1386-1391
    else if (L.backward) { q.minL = 0; q.maxL = oct; }
    else { q.minL = oct - 1; q.maxL = oct + 1; }
    q.d = L.desc + 8 * (size_t)i;
    q.ur_proj = q.u - C.bf * invzc;
    return true;
}
// Per-candidate static filter of both overloads: a keypoint that already holds a MapPoint (with observations for the
// frame-to-frame overload: kp_state -2; the caller of the relocalisation overload encodes every non-null entry as -2), and the
// stereo consistency test (:1408-1414, frame-to-frame only).
__device__ __forceinline__ bool m3_accept(const FrameDev& C, const LastDev& L, const M3Query& q, const int* __restrict__ kp_state, int i2) {
    if (kp_state[i2] == -2) return false;
    if (!L.reloc && C.uright) {
        const float ur = C.uright[i2];
        if (ur > 0 && fabsf(q.ur_proj - ur) > q.radius) return false;
    }
    return true;
}

template <class Fn>
__device__ __forceinline__ void m3_visit(const FrameDev& C, const LastDev& L, float th, const int* __restrict__ kp_state, int i, Fn&& fn) {
    M3Query q;
    if (!m3_query(C, L, th, i, q)) return;
    for_each_in_area(C, q.u, q.v, q.radius, q.minL, q.maxL, [&](int i2) {
        if (m3_accept(C, L, q, kp_state, i2)) fn(i2, hamming256(q.d, C.desc + 8 * (size_t)i2), 0);
    });
}

__global__ void __launch_bounds__(256) m3_collect_kernel(FrameDev Cf, LastDev L, float th, const int* kp_state, CandLists C) {
    COEB_MTRACE(5);
    pdl_launch_dependents();   // see m2_collect_kernel
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per last-frame point
    if (i >= L.n) return;
    M3Query q;
    int cnt = 0;
    unsigned k1 = 0xFFFFFFFFu;   // this lane's smallest distance << 16 | list position
    int k1x = -1;                // ... and its keypoint
    const bool obs = L.has_obs[i] != 0;
    if (m3_query(Cf, L, th, i, q)) {
        int2* out = C.items + (size_t)i * C.cap;
        cnt = warp_for_each_in_area(
            Cf, q.u, q.v, q.radius, q.minL, q.maxL,
            [&](int i2) { return m3_accept(Cf, L, q, kp_state, i2); },
            [&](int pos, int i2) {
                if (pos < C.cap) {
                    const int dist = hamming256(q.d, Cf.desc + 8 * (size_t)i2);
                    out[pos] = make_int2(i2, dist);
                    const unsigned key = ((unsigned)dist << 16) | (unsigned)pos;
                    if (key < k1) { k1 = key; k1x = i2; }
                }
            });
    }
    unsigned b1 = k1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b1 = min(b1, __shfl_xor_sync(0xffffffffu, b1, o));   // strict '<' of :1421-1425: first of the smallest
    const int x1 = __shfl_sync(0xffffffffu, k1x, __ffs(__ballot_sync(0xffffffffu, k1 == b1)) - 1);   // (the key carries its position: one holder)
    if ((threadIdx.x & 31) == 0) {
        if (C.top) C.top[i] = (int)(b1 & 0xFFFFu);
        C.count[i] = cnt;
        // a query whose smallest distance exceeds the acceptance bound can never match or claim (exclusions only raise it): not handed on
        if (cnt > 0 && (b1 >> 16) <= (unsigned)L.max_accept) {
            const int a = atomicAdd(&C.meta[0], 1);
            C.active[a] = i;
            if (C.rec) C.rec[a] = make_int4(obs ? i : (i | kNoObsBit), x1, -1, (int)(b1 >> 16) | (256 << 9) | (min(cnt, 255) << 18));
        }
        if (cnt > C.cap) C.meta[1] = 1;
    }
}

template <bool kLists, bool kSmem>
__global__ void __launch_bounds__(512, 1) m3_resolve_kernel(FrameDev C, LastDev L, float th, int check_ori, const int* kp_state, CandLists Cl,
                                                          int* kp_match, int* res, int* claim_glob, int* out_info) {
    COEB_MTRACE(6);
    constexpr bool claim_in_smem = kSmem;
    __shared__ int s_changed, s_count;
    __shared__ int s_hist[COEB_HISTO_LENGTH];
    __shared__ int s_ind[3];
    const int tid = threadIdx.x, T = blockDim.x;
    int* const claim_min = kSmem ? s_claim_dyn : claim_glob;
    const int na = kLists ? Cl.meta[0] : L.n;   // queries with candidates (compacted by the collect kernel, any order)
    const bool overflow = kLists && Cl.meta[1];
    if (kLists) {   // cleared for the next call's collect kernel (see CandLists::meta)
        __syncthreads();
        if (tid == 0) { Cl.meta[0] = 0; Cl.meta[1] = 0; }
    }
    if (overflow) {
        if (tid == 0) { out_info[0] = 0; out_info[1] = 0; out_info[2] = 1; }
        return;
    }
    auto query = [&](int a) { return kLists ? Cl.active[a] : a; };
    constexpr int Q = 10;
    if (kLists && na <= Q * T) {
        // Register path (see m2_resolve_kernel): a thread owns up to Q active queries and keeps the list entry that wins when no claim
        // touches it, so a round of the fixed-point iteration reads shared memory only.
        int qi[Q], qres[Q];
        int2 q1[Q];
        bool qobs[Q];
#pragma unroll
        for (int j = 0; j < Q; j++) {
            const int a = tid + j * T;
            qi[j] = -1; qres[j] = -1; qobs[j] = false; q1[j] = make_int2(0, 256);
            if (a < na) {
                const int i = Cl.active[a];
                qi[j] = i;
                q1[j] = Cl.items[(size_t)i * Cl.cap + Cl.top[i]];
                qobs[j] = L.has_obs[i] != 0;
            }
        }
        for (int round = 0; round <= L.n; round++) {
            for (int k = tid; k < C.n; k += T) claim_min[k] = kInf;
            if (tid == 0) s_changed = 0;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < Q; j++)
                if (qres[j] >= 0 && qobs[j]) atomicMin(&claim_min[qres[j]], qi[j]);
            __syncthreads();
            bool changed_here = false;
#pragma unroll
            for (int j = 0; j < Q; j++) {
                const int i = qi[j];
                if (i < 0) continue;
                int bestDist = 256, bestIdx2 = -1;
                if (claim_min[q1[j].x] >= i) {
                    if (q1[j].y < 256) { bestDist = q1[j].y; bestIdx2 = q1[j].x; }
                } else {
                    walk_list(Cl.items + (size_t)i * Cl.cap, Cl.count[i], [&](const int2 e) {
                        if (claim_min[e.x] < i) return;                          // :1404-1406 with :1429
                        if (e.y < bestDist) { bestDist = e.y; bestIdx2 = e.x; }
                    });
                }
                const int out = bestDist <= L.max_accept ? bestIdx2 : -1;
                if (out != qres[j]) { qres[j] = out; changed_here = true; }
            }
            if (changed_here) s_changed = 1;
            __syncthreads();
            const int changed = s_changed;
            __syncthreads();
            if (!changed) { if (tid == 0) out_info[1] = round + 1; break; }
        }
        if (tid == 0) s_count = 0;
        for (int k = tid; k < COEB_HISTO_LENGTH; k += T) s_hist[k] = 0;
        for (int k = tid; k < C.n; k += T) claim_min[k] = -1;
        __syncthreads();
        int mine = 0;
        int qbin[Q];
#pragma unroll
        for (int j = 0; j < Q; j++) {
            qbin[j] = -1;
            if (qres[j] >= 0) {
                atomicMax(&claim_min[qres[j]], qi[j]);
                mine++;
                if (check_ori) { qbin[j] = rot_bin(L.angle[qi[j]], C.angle[qres[j]]); atomicAdd(&s_hist[qbin[j]], 1); }
            }
        }
        if (mine) atomicAdd(&s_count, mine);
        __syncthreads();
        for (int k = tid; k < C.n; k += T)
            kp_match[k] = claim_min[k] >= 0 ? claim_min[k] : kp_state[k];   // untouched entries keep the caller's state
        if (tid == 0 && check_ori) {
            int a, b, c;
            three_maxima(s_hist, COEB_HISTO_LENGTH, a, b, c);
            s_ind[0] = a; s_ind[1] = b; s_ind[2] = c;
        }
        __syncthreads();
        if (check_ori) {  // rotation consistency (:1449-1468): entries of the losing bins are cleared, each decrements
            int removed = 0;
#pragma unroll
            for (int j = 0; j < Q; j++)
                if (qres[j] >= 0 && qbin[j] != s_ind[0] && qbin[j] != s_ind[1] && qbin[j] != s_ind[2]) { kp_match[qres[j]] = -1; removed++; }
            if (removed) atomicSub(&s_count, removed);
        }
        __syncthreads();
        if (tid == 0) { out_info[0] = s_count; out_info[2] = 0; }
        return;
    }
    for (int a = tid; a < na; a += T) res[query(a)] = -1;
    for (int round = 0; round <= L.n; round++) {
        for (int k = tid; k < C.n; k += T) claim_min[k] = kInf;
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int a = tid; a < na; a += T) {
            const int i = query(a), k = res[i];
            if (k >= 0 && L.has_obs[i]) atomicMin(&claim_min[k], i);
        }
        __syncthreads();
        for (int a = tid; a < na; a += T) {
            const int i = query(a);
            int bestDist = 256, bestIdx2 = -1;
            auto consider = [&](int i2, int dist, int) {
                if (claim_min[i2] < i) return;                           // :1404-1406 with :1429
                if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
            };
            if (kLists) {
                walk_list(Cl.items + (size_t)i * Cl.cap, Cl.count[i], [&](const int2 e) { consider(e.x, e.y, 0); });
            } else {
                m3_visit(C, L, th, kp_state, i, consider);
            }
            const int out = bestDist <= L.max_accept ? bestIdx2 : -1;
            if (out != res[i]) { res[i] = out; s_changed = 1; }
        }
        __syncthreads();
        const int changed = s_changed;
        __syncthreads();
        if (!changed) { if (tid == 0) out_info[1] = round + 1; break; }
    }
    if (tid == 0) s_count = 0;
    for (int k = tid; k < COEB_HISTO_LENGTH; k += T) s_hist[k] = 0;
    for (int k = tid; k < C.n; k += T) claim_min[k] = -1;
    __syncthreads();
    int mine = 0;
    for (int a = tid; a < na; a += T) {
        const int i = query(a);
        if (res[i] >= 0) {
            atomicMax(&claim_min[res[i]], i);
            mine++;
            if (check_ori) atomicAdd(&s_hist[rot_bin(L.angle[i], C.angle[res[i]])], 1);
        }
    }
    if (mine) atomicAdd(&s_count, mine);
    __syncthreads();
    for (int k = tid; k < C.n; k += T)
        kp_match[k] = claim_min[k] >= 0 ? claim_min[k] : kp_state[k];   // untouched entries keep the caller's state
    if (tid == 0 && check_ori) {
        int a, b, c;
        three_maxima(s_hist, COEB_HISTO_LENGTH, a, b, c);
        s_ind[0] = a; s_ind[1] = b; s_ind[2] = c;
    }
    __syncthreads();
    if (check_ori) {  // rotation consistency (:1449-1468): entries of the losing bins are cleared, each decrements
        int removed = 0;
        for (int a = tid; a < na; a += T) {
            const int i = query(a);
            if (res[i] >= 0) {
                const int bin = rot_bin(L.angle[i], C.angle[res[i]]);
                if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) { kp_match[res[i]] = -1; removed++; }
            }
        }
        if (removed) atomicSub(&s_count, removed);
    }
    __syncthreads();
    if (tid == 0) { out_info[0] = s_count; out_info[2] = 0; }
}

// ---- M4: SearchForInitialization (src/ORBmatcher.cc:405-520) -----------------------------------------------------
template <class Fn>
__device__ __forceinline__ void m4_visit(const FrameDev& F1, const FrameDev& F2, const float* __restrict__ prev_in, float window, int i1, Fn&& fn) {
    if (F1.octave[i1] > 0) return;   // level1 > 0 -> continue (:421-423)
    const uint32_t* d1 = F1.desc + 8 * (size_t)i1;
    for_each_in_area(F2, prev_in[2 * i1], prev_in[2 * i1 + 1], window, 0, 0,
                     [&](int i2) { fn(i2, hamming256(d1, F2.desc + 8 * (size_t)i2), 0); });
}

__global__ void __launch_bounds__(256) m4_collect_kernel(FrameDev F1, FrameDev F2, const float* prev_in, float window, CandLists C) {
    COEB_MTRACE(5);
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per keypoint of F1
    if (i >= F1.n) return;
    int cnt = 0;
    if (F1.octave[i] <= 0) {   // level1 > 0 -> continue (:421-423)
        const uint32_t* d1 = F1.desc + 8 * (size_t)i;
        int2* out = C.items + (size_t)i * C.cap;
        cnt = warp_for_each_in_area(
            F2, prev_in[2 * i], prev_in[2 * i + 1], window, 0, 0, [](int) { return true; },
            [&](int pos, int i2) {
                if (pos < C.cap) out[pos] = make_int2(i2, hamming256(d1, F2.desc + 8 * (size_t)i2));
            });
    }
    if ((threadIdx.x & 31) == 0) cand_finish(C, i, cnt);
}

// res[i1] = claimed F2 keypoint or -1, rdist[i1] = its distance. vMatchedDistance seen by query i1 at keypoint k is
// min{rdist[j] : j < i1, res[j] == k}; claimants of every k are gathered into CSR lists each round.
template <bool kLists>
__global__ void __launch_bounds__(1024) m4_resolve_kernel(FrameDev F1, FrameDev F2, const float* prev_in, float window, float nnratio,
                                                          int check_ori, CandLists C, int* res, int* rdist, int* cl_start, int* cl_fill,
                                                          int2* cl_items, int* matches12, float* prev_out, int* out_info, int* m12_host) {
    // (prev_out and out_info may be mapped pinned memory: written once; matches12 is read back and stays on the device, m12_host is its copy)
    COEB_MTRACE(6);
    __shared__ int s_changed, s_count;
    __shared__ int s_hist[COEB_HISTO_LENGTH];
    __shared__ int s_ind[3];
    __shared__ int s_warp[33];
    const int tid = threadIdx.x, T = blockDim.x;
    const bool overflow = kLists && C.meta[1];
    if (kLists) {   // cleared for the next call's collect kernel (see CandLists::meta)
        __syncthreads();
        if (tid == 0) { C.meta[0] = 0; C.meta[1] = 0; }
    }
    if (overflow) {
        if (tid == 0) { out_info[0] = 0; out_info[1] = 0; out_info[2] = 1; }
        return;
    }
    for (int i = tid; i < F1.n; i += T) { res[i] = -1; rdist[i] = kInf; }
    for (int round = 0; round <= F1.n; round++) {
        for (int k = tid; k <= F2.n; k += T) cl_start[k] = 0;
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int i = tid; i < F1.n; i += T)
            if (res[i] >= 0) atomicAdd(&cl_start[res[i]], 1);
        __syncthreads();
        block_exclusive_scan(cl_start, F2.n + 1, s_warp);
        for (int k = tid; k < F2.n; k += T) cl_fill[k] = cl_start[k];
        __syncthreads();
        for (int i = tid; i < F1.n; i += T)
            if (res[i] >= 0) cl_items[atomicAdd(&cl_fill[res[i]], 1)] = make_int2(i, rdist[i]);
        __syncthreads();
        for (int i1 = tid; i1 < F1.n; i1 += T) {
            int bestDist = kInf, bestDist2 = kInf, bestIdx2 = -1;
            auto consider = [&](int i2, int dist, int) {
                int md = kInf;
                for (int p = cl_start[i2]; p < cl_start[i2 + 1]; p++) {
                    const int2 c = cl_items[p];
                    if (c.x < i1) md = min(md, c.y);
                }
                if (md <= dist) return;   // vMatchedDistance[i2] <= dist (:444)
                if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
                else if (dist < bestDist2) bestDist2 = dist;
            };
            if (kLists) {
                walk_list(C.items + (size_t)i1 * C.cap, C.count[i1], [&](const int2 e) { consider(e.x, e.y, 0); });
            } else {
                m4_visit(F1, F2, prev_in, window, i1, consider);
            }
            int out = -1, outd = kInf;
            if (bestDist <= COEB_TH_LOW && (float)bestDist < (float)bestDist2 * nnratio) { out = bestIdx2; outd = bestDist; }
            if (out != res[i1] || outd != rdist[i1]) { res[i1] = out; rdist[i1] = outd; s_changed = 1; }
        }
        __syncthreads();
        const int changed = s_changed;
        __syncthreads();
        if (!changed) { if (tid == 0) out_info[1] = round + 1; break; }
    }
    // vnMatches12[i1] survives iff i1 is the last claimant of its keypoint (steal-back, :463-469);
    // nmatches = keypoints with at least one claimant.
    for (int k = tid; k < F2.n; k += T) cl_fill[k] = -1;
    for (int k = tid; k < COEB_HISTO_LENGTH; k += T) s_hist[k] = 0;
    if (tid == 0) s_count = 0;
    __syncthreads();
    for (int i = tid; i < F1.n; i += T)
        if (res[i] >= 0) {
            atomicMax(&cl_fill[res[i]], i);
            if (check_ori) atomicAdd(&s_hist[rot_bin(F1.angle[i], F2.angle[res[i]])], 1);  // every success is binned (:482)
        }
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < F1.n; i += T) {
        const int k = res[i];
        const bool keep = k >= 0 && cl_fill[k] == i;
        matches12[i] = keep ? k : -1;
        mine += keep;
    }
    if (mine) atomicAdd(&s_count, mine);
    if (tid == 0 && check_ori) {
        int a, b, c;
        three_maxima(s_hist, COEB_HISTO_LENGTH, a, b, c);
        s_ind[0] = a; s_ind[1] = b; s_ind[2] = c;
    }
    __syncthreads();
    if (check_ori) {
        int removed = 0;
        for (int i = tid; i < F1.n; i += T)
            if (res[i] >= 0) {
                const int bin = rot_bin(F1.angle[i], F2.angle[res[i]]);
                if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2] && matches12[i] >= 0) { matches12[i] = -1; removed++; }
            }
        if (removed) atomicSub(&s_count, removed);
    }
    __syncthreads();
    for (int i = tid; i < F1.n; i += T) {  // update prev matched (:515-517)
        const int k = matches12[i];
        m12_host[i] = k;
        prev_out[2 * i] = k >= 0 ? F2.x[k] : prev_in[2 * i];
        prev_out[2 * i + 1] = k >= 0 ? F2.y[k] : prev_in[2 * i + 1];
    }
    if (tid == 0) { out_info[0] = s_count; out_info[2] = 0; }
}

// ---- F4: Frame::ComputeStereoMatches (src/Frame.cc:644-818) -------------------------------------------------------
struct StereoDev {
    int N, Nr;
    const float *xl, *yl; const int* octl; const uint32_t* descl;
    const float *xr, *yr; const int* octr; const uint32_t* descr;
    float scale[COEB_MAX_LEVELS], inv_scale[COEB_MAX_LEVELS];
    const uint8_t* pyrL[COEB_MAX_LEVELS]; const uint8_t* pyrR[COEB_MAX_LEVELS];
    int pitchL[COEB_MAX_LEVELS], pitchR[COEB_MAX_LEVELS], lw[COEB_MAX_LEVELS], lh[COEB_MAX_LEVELS];
    int nRows;
    float bf, b;
};

// One warp per left keypoint. The reference's row table (vRowIndices) lists, for image row (int)vL, the right
// keypoints iR (ascending) with floor(yR - r) <= row <= ceil(yR + r), r = 2*scale[octR]; membership is tested directly.
// Best Hamming: strict '<' from TH_HIGH, first iR wins -> argmin over (dist, iR). Then the 11x11 SAD search over 11
// shifts at the left keypoint's level, parabola refinement and the disparity test.
__global__ void __launch_bounds__(256) stereo_match_kernel(StereoDev S, float* uright, float* depth, int* sad_out) {
    const int lane = threadIdx.x & 31;
    const int iL = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (iL >= S.N) return;
    float out_u = -1.f, out_d = -1.f;
    int out_sad = -1;
    const float uL = S.xl[iL], vL = S.yl[iL];
    const int levelL = S.octl[iL];
    const int row = (int)vL;
    const float minZ = S.b, minD = 0.f, maxD = S.bf / minZ;
    const float minU = uL - maxD, maxU = uL - minD;
    unsigned long long best = ((unsigned long long)COEB_TH_HIGH << 32) | 0xFFFFFFFFull;
    bool any = false;
    if (row >= 0 && row < S.nRows && !(maxU < 0)) {
        const uint32_t* dL = S.descl + 8 * (size_t)iL;
        for (int iR = lane; iR < S.Nr; iR += 32) {
            const float kpY = S.yr[iR];
            const int oR = S.octr[iR];
            const float r = 2.0f * S.scale[oR];
            const int maxr = (int)ceilf(kpY + r), minr = (int)floorf(kpY - r);
            if (row < minr || row > maxr) continue;
            any = true;
            if (oR < levelL - 1 || oR > levelL + 1) continue;
            const float uR = S.xr[iR];
            if (uR >= minU && uR <= maxU) {
                const int dist = hamming256(dL, S.descr + 8 * (size_t)iR);
                if (dist < COEB_TH_HIGH) {
                    const unsigned long long key = ((unsigned long long)dist << 32) | (unsigned)iR;
                    best = key < best ? key : best;
                }
            }
        }
    }
    any = __any_sync(0xffffffffu, any);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    const int bestDist = (int)(best >> 32);
    const int thOrbDist = (COEB_TH_HIGH + COEB_TH_LOW) / 2;
    if (any && bestDist < thOrbDist) {
        const int bestIdxR = (int)(best & 0xFFFFFFFFu);
        const float uR0 = S.xr[bestIdxR];
        const float sf = S.inv_scale[levelL];
        const float scaleduL = roundf(uL * sf), scaledvL = roundf(vL * sf), scaleduR0 = roundf(uR0 * sf);
        const int w = 5, L = 5;
        const int yl = (int)(scaledvL - w), xl = (int)(scaleduL - w);
        const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
        const int W = S.lw[levelL], H = S.lh[levelL];
        bool ok = !(iniu < 0 || endu >= (float)W);
        // the reference's rowRange/colRange would throw outside the image; such points are dropped
        if (yl < 0 || yl + 2 * w + 1 > H || xl < 0 || xl + 2 * w + 1 > W) ok = false;
        if (ok) {
            const uint8_t* IL = S.pyrL[levelL];
            const uint8_t* IR = S.pyrR[levelL];
            const int pL = S.pitchL[levelL], pR = S.pitchR[levelL];
            const int cL = IL[(size_t)(yl + w) * pL + xl + w];
            // each lane owns up to 4 of the 121 patch pixels
            int lv[4], py[4], px[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int p = lane + 32 * k;
                py[k] = p / 11; px[k] = p - py[k] * 11;
                lv[k] = p < 121 ? (int)IL[(size_t)(yl + py[k]) * pL + xl + px[k]] - cL : 0;
            }
            int bestSad = 0x7fffffff, bestinc = 0;
            float vd[11];
#pragma unroll
            for (int inc = -5; inc <= 5; inc++) {
                const int xr = (int)(scaleduR0 + (float)inc - w);
                const int cR = IR[(size_t)(yl + w) * pR + xr + w];
                int acc = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int p = lane + 32 * k;
                    if (p < 121) acc += abs(lv[k] - ((int)IR[(size_t)(yl + py[k]) * pR + xr + px[k]] - cR));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                vd[inc + 5] = (float)acc;   // exact: the L1 norm of integer differences
                if (acc < bestSad) { bestSad = acc; bestinc = inc; }
            }
            if (!(bestinc == -L || bestinc == L)) {
                float d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                for (int k = 1; k < 10; k++)
                    if (k == bestinc + 5) { d1 = vd[k - 1]; d2 = vd[k]; d3 = vd[k + 1]; }
                const float deltaR = (d1 - d3) / (2.0f * (d1 + d3 - 2.0f * d2));
                if (!(deltaR < -1 || deltaR > 1)) {
                    float bestuR = S.scale[levelL] * (scaleduR0 + (float)bestinc + deltaR);
                    float disparity = uL - bestuR;
                    if (disparity >= minD && disparity < maxD) {
                        if (disparity <= 0) { disparity = 0.01f; bestuR = (float)((double)uL - 0.01); }
                        out_d = S.bf / disparity;
                        out_u = bestuR;
                        out_sad = bestSad;
                    }
                }
            }
        }
    }
    if (lane == 0) { uright[iL] = out_u; depth[iL] = out_d; sad_out[iL] = out_sad; }
}

// Outlier cut (:804-817): median = SAD at sorted position size/2; entries with SAD >= 1.5*1.4*median are removed.
// The final mvuRight / mvDepth and the count also go into mapped pinned memory (`*_host`): no device-to-host copy node behind the kernel.
__global__ void __launch_bounds__(1024) stereo_outlier_kernel(int N, float* uright, float* depth, const int* sad, int* out_info, float* uright_host, float* depth_host,
                                                              int* info_host) {
    __shared__ int s_cnt, s_hi, s_sel[2];
    __shared__ int s_hist[256];
    const int tid = threadIdx.x, T = blockDim.x;
    if (tid == 0) s_cnt = 0;
    if (tid < 256) s_hist[tid] = 0;
    __syncthreads();
    // the median SAD by a two-level histogram (SADs are below 121 * 510 < 2^16: high byte, then low byte inside the bin that holds the
    // rank): three passes over the values instead of the 16 of a bisection on the value range
    int mine = 0;
    for (int i = tid; i < N; i += T) {
        const int v = sad[i];
        if (v >= 0) { mine++; atomicAdd(&s_hist[v >> 8], 1); }
    }
    if (mine) atomicAdd(&s_cnt, mine);
    __syncthreads();
    const int M = s_cnt;
    if (M == 0) {
        for (int i = tid; i < N; i += T) { uright_host[i] = uright[i]; depth_host[i] = depth[i]; }
        if (tid == 0) { out_info[0] = 0; info_host[0] = 0; }
        return;
    }
    // bin that holds ascending rank r (0-based) of the histogram, and the rank inside it: warp 0, eight bins per lane
    auto locate = [&](int r) {
        if (tid < 32) {
            int h[8], sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { h[k] = s_hist[8 * tid + k]; sum += h[k]; }
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t2 = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += t2; }
            int c = inc - sum;
            if (r >= c && r < inc) {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (r < c + h[k]) { s_sel[0] = 8 * tid + k; s_sel[1] = r - c; break; }
                    c += h[k];
                }
            }
        }
    };
    locate(M / 2);   // sorted position size / 2 (:806)
    __syncthreads();
    const int hi_bin = s_sel[0], r2 = s_sel[1];
    __syncthreads();
    if (tid < 256) s_hist[tid] = 0;
    if (tid == 0) s_hi = 0;
    __syncthreads();
    for (int i = tid; i < N; i += T) {
        const int v = sad[i];
        if (v >= 0 && (v >> 8) == hi_bin) atomicAdd(&s_hist[v & 255], 1);
    }
    __syncthreads();
    locate(r2);
    __syncthreads();
    const float median = (float)((hi_bin << 8) | s_sel[0]);
    const float thDist = 1.5f * 1.4f * median;
    int kept = 0;
    for (int i = tid; i < N; i += T) {
        float u = uright[i], d = depth[i];
        if (sad[i] >= 0) {
            if ((float)sad[i] < thDist) kept++;
            else { u = -1.f; d = -1.f; uright[i] = u; depth[i] = d; }
        }
        uright_host[i] = u; depth_host[i] = d;
    }
    if (kept) atomicAdd(&s_hi, kept);
    __syncthreads();
    if (tid == 0) { out_info[0] = s_hi; info_host[0] = s_hi; }
}

// ---- config 5: brute-force k=2 Hamming search + ratio test -----------------------------------------------------------
// Integer-pipe bound: 8 XOR + 8 POPC + adds per pair. Each thread keeps kKnnQ query descriptors in registers and
// walks a train chunk staged in shared memory (all lanes read the same train word: broadcast). Partial (d1, idx, d2)
// per (query, chunk) are merged by a second kernel; the strict-'<' / first-index-wins rule is order independent once
// ties are broken on the index, and d2 is the second smallest distance with multiplicity.
constexpr int kKnnQ = 2;         // queries per thread
constexpr int kKnnThreads = 128;
constexpr int kKnnTile = 256;    // train descriptors per shared-memory stage

// Hamming distance of two 256-bit descriptors with five population counts instead of eight. POPC issues on the XU pipe at a
// quarter of the integer-ALU rate (ncu: the plain 8-POPC loop ran the XU pipe at 91.6 % with the ALU pipe half idle), so three
// carry-save adders (two LOP3 each: parity 0x96, majority 0xE8) fold seven of the eight difference words into two "ones" words and
// three "twos" words first: distance = popc(s3) + popc(w7) + 2 * (popc(c1) + popc(c2) + popc(c3)). Exact, like the plain sum.
__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ int hamming256_csa(const uint4& qa, const uint4& qb, const uint4& ta, const uint4& tb) {
    const uint32_t w0 = qa.x ^ ta.x, w1 = qa.y ^ ta.y, w2 = qa.z ^ ta.z, w3 = qa.w ^ ta.w;
    const uint32_t w4 = qb.x ^ tb.x, w5 = qb.y ^ tb.y, w6 = qb.z ^ tb.z, w7 = qb.w ^ tb.w;
    const uint32_t s1 = lop3_xor3(w0, w1, w2), c1 = lop3_maj(w0, w1, w2);
    const uint32_t s2 = lop3_xor3(w3, w4, w5), c2 = lop3_maj(w3, w4, w5);
    const uint32_t s3 = lop3_xor3(s1, s2, w6), c3 = lop3_maj(s1, s2, w6);
    // (folding further, down to four population counts, makes the ALU pipe the bound: 0.70 ms against 0.61 ms for this form)
    return (__popc(s3) + __popc(w7)) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
}

__global__ void __launch_bounds__(kKnnThreads) knn2_partial_kernel(const uint32_t* __restrict__ q, int nq, const uint32_t* __restrict__ t, int nt,
                                                                    int chunk, int* p_d1, int* p_idx, int* p_d2) {
    __shared__ uint4 s_t[kKnnTile * 2];
    const int q0 = (blockIdx.x * kKnnThreads + threadIdx.x) * kKnnQ;
    const int c = blockIdx.y;
    const int t_lo = c * chunk, t_hi = min(t_lo + chunk, nt);
    uint4 qa[kKnnQ], qb[kKnnQ];
    int d1[kKnnQ], d2[kKnnQ], bi[kKnnQ];
#pragma unroll
    for (int k = 0; k < kKnnQ; k++) {
        const int qi = min(q0 + k, nq - 1);
        qa[k] = *reinterpret_cast<const uint4*>(q + 8 * (size_t)qi);
        qb[k] = *reinterpret_cast<const uint4*>(q + 8 * (size_t)qi + 4);
        d1[k] = 256; d2[k] = 256; bi[k] = -1;
    }
    for (int base = t_lo; base < t_hi; base += kKnnTile) {
        const int m = min(kKnnTile, t_hi - base);
        __syncthreads();
        for (int i = threadIdx.x; i < m * 2; i += kKnnThreads) s_t[i] = *reinterpret_cast<const uint4*>(t + 8 * (size_t)base + 4 * (size_t)i);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < m; j++) {
            const uint4 ta = s_t[2 * j], tb = s_t[2 * j + 1];
#pragma unroll
            for (int k = 0; k < kKnnQ; k++) {
                const int dist = hamming256_csa(qa[k], qb[k], ta, tb);
                if (dist < d1[k]) { d2[k] = d1[k]; d1[k] = dist; bi[k] = base + j; }
                else if (dist < d2[k]) d2[k] = dist;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kKnnQ; k++)
        if (q0 + k < nq) {
            const size_t o = (size_t)c * nq + q0 + k;
            p_d1[o] = d1[k]; p_idx[o] = bi[k]; p_d2[o] = d2[k];
        }
}

__global__ void knn2_merge_kernel(int nq, int nchunks, const int* p_d1, const int* p_idx, const int* p_d2, float nnratio, int* best_idx,
                                  int* out_d1, int* out_d2, int* accepted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    int d1 = 256, d2 = 256, idx = -1;
    for (int c = 0; c < nchunks; c++) {  // chunks in ascending train order, so an equal distance never replaces the earlier index
        const size_t o = (size_t)c * nq + i;
        const int a = p_d1[o], b = p_d2[o];
        if (a < d1) { d2 = min(d1, b); d1 = a; idx = p_idx[o]; }
        else { d2 = min(d2, a); }
    }
    const bool ok = d1 <= COEB_TH_LOW && (float)d1 < nnratio * (float)d2;
    best_idx[i] = ok ? idx : -1;
    if (out_d1) out_d1[i] = d1;
    if (out_d2) out_d2[i] = d2;
    if (ok && accepted) atomicAdd(accepted, 1);
}


// =====================================================================================================================
// Either side of the extractor and the matchers ("next" rows): the tail of the RGB-D Frame constructor and the
// visibility test of Tracking::SearchLocalPoints, so that keypoints, descriptors and projected map points never make a
// host round trip between the extractor and the matcher.
// =====================================================================================================================

// cv::undistortPoints(src, dst, K, D, noArray(), K) for one point (Frame::UndistortKeyPoints, src/Frame.cc:579-609):
// double precision, 5 fixed-point iterations, no FMA (-fmad=false). dist = {k1, k2, p1, p2, k3}.
__device__ __forceinline__ void undistort_point(float u, float v, float fx, float fy, float cx, float cy, const float* dist, float& xo, float& yo) {
    const double dfx = fx, dfy = fy, dcx = cx, dcy = cy;
    const double k0 = dist[0], k1 = dist[1], k2 = dist[2], k3 = dist[3], k4 = dist[4];
    const double ifx = 1. / dfx, ify = 1. / dfy;
    double x = ((double)u - dcx) * ifx, y = ((double)v - dcy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = 1. / (1 + ((k4 * r2 + k1) * r2 + k0) * r2);
        if (icdist < 0) { x = x0; y = y0; break; }
        const double deltaX = 2 * k2 * x * y + k3 * (r2 + 2 * x * x);
        const double deltaY = k2 * (r2 + 2 * y * y) + 2 * k3 * x * y;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    xo = (float)(dfx * x + dcx);
    yo = (float)(dfy * y + dcy);
}

struct TailArgs {
    const coeb_keypoint* kps;    // mvKeys (extractor output, device)
    const uint32_t* desc;        // mDescriptors, n x 8 words
    int n;
    float fx, fy, cx, cy, bf;
    float dist[5];
    int undistort;
    const void* depth; int depth_kind, depth_stride, depth_w, depth_h; float depth_factor;   // kind 3 / 4: one float per keypoint, gathered on the host (4: raw uint16 value, factor still to apply)
    // frame block
    float *x, *y, *angle; int* octave; float* uright; uint32_t* desc_out;
    coeb_keypoint* keys_un; float* depth_out; float* uright_dl;   // download block: mvKeysUn | mvuRight | mvDepth, contiguous
};

// UndistortKeyPoints + ComputeStereoFromRGBD (src/Frame.cc:579-609, 820-842): one thread per keypoint writes the SoA
// the matchers read and the AoS the host downloads; the descriptor rows are copied as words.
__device__ __forceinline__ void frame_tail_point(const TailArgs& a, int i, float& xu, float& yu) {
    coeb_keypoint kp = a.kps[i];
    const float u = kp.x, v = kp.y;
    if (a.undistort) undistort_point(u, v, a.fx, a.fy, a.cx, a.cy, a.dist, kp.x, kp.y);
    float ur = -1.f, dp = -1.f;
    if (a.depth_kind) {
        const int row = (int)v, col = (int)u;   // imDepth.at<float>(v, u): float -> int truncation
        float d = 0.f;
        if (a.depth_kind >= 3) {
            const float g = ((const float*)a.depth)[i];
            d = a.depth_kind == 3 ? g : __fmul_rn(g, a.depth_factor);
        } else if ((unsigned)row < (unsigned)a.depth_h && (unsigned)col < (unsigned)a.depth_w) {
            const char* p = (const char*)a.depth + (size_t)row * a.depth_stride;
            d = a.depth_kind == 1 ? ((const float*)p)[col] : __fmul_rn((float)((const unsigned short*)p)[col], a.depth_factor);
        }
        if (d > 0) { dp = d; ur = __fsub_rn(kp.x, __fdiv_rn(a.bf, d)); }
    }
    a.x[i] = kp.x; a.y[i] = kp.y; a.angle[i] = kp.angle; a.octave[i] = kp.octave; a.uright[i] = ur;
    a.keys_un[i] = kp; a.depth_out[i] = dp; a.uright_dl[i] = ur;
    xu = kp.x; yu = kp.y;
}
__global__ void __launch_bounds__(128) frame_tail_kernel(TailArgs a) {
    COEB_MTRACE(0);
    const int n = a.n;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (int w = i; w < n * 8; w += gridDim.x * blockDim.x) a.desc_out[w] = a.desc[w];
    if (i >= n) return;
    float xu, yu;
    frame_tail_point(a, i, xu, yu);
}

// Both steps in ONE single-CTA launch for frames of up to kTailGridMax keypoints (a tracked frame has 1000-2000): the per-keypoint
// tail, then Frame::AssignFeaturesToGrid from the undistorted coordinates still in registers, with the cell table, the fill cursors
// and the item list in shared memory. As two launches the grid kernel started 7 us after the 3-us tail kernel (whose writes into the
// caller's mapped arrays have to reach the host first) and then spent 9 us on global round trips.
constexpr int kTailGridMax = 2048;
__global__ void __launch_bounds__(1024, 1) frame_tail_grid_kernel(TailArgs a, FrameDev f, int* __restrict__ cell_start, int* __restrict__ cell_items) {
    COEB_MTRACE(0);
    __shared__ int s_start[kGridCells + 1], s_fill[kGridCells];
    __shared__ short s_cell[kTailGridMax], s_items[kTailGridMax];
    __shared__ int s_warp[33];
    const int tid = threadIdx.x, T = blockDim.x, n = a.n;
    for (int i = tid; i <= kGridCells; i += T) s_start[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += T) {
        float xu, yu;
        frame_tail_point(a, i, xu, yu);
        // PosInGrid: round() half away from zero (src/Frame.cc:560-561)
        const int px = (int)roundf((xu - f.min_x) * f.gw_inv), py = (int)roundf((yu - f.min_y) * f.gh_inv);
        int c = -1;
        if (!(px < 0 || px >= COEB_GRID_COLS || py < 0 || py >= COEB_GRID_ROWS)) {
            c = px * COEB_GRID_ROWS + py;
            atomicAdd(&s_start[c], 1);
        }
        s_cell[i] = (short)c;
    }
    for (int w = tid; w < n * 8; w += T) a.desc_out[w] = a.desc[w];
    __syncthreads();
    const int total = block_exclusive_scan(s_start, kGridCells, s_warp);
    if (tid == 0) s_start[kGridCells] = total;
    __syncthreads();
    for (int i = tid; i <= kGridCells; i += T) {
        const int v = s_start[i];
        cell_start[i] = v;
        if (i < kGridCells) s_fill[i] = v;
    }
    __syncthreads();
    // fill with atomic slots, then sort each (tiny) cell: items end up ascending by keypoint index, the order in which the reference pushes them
    for (int i = tid; i < n; i += T) {
        const int c = s_cell[i];
        if (c >= 0) s_items[atomicAdd(&s_fill[c], 1)] = (short)i;
    }
    __syncthreads();
    for (int c = tid; c < kGridCells; c += T) {
        const int lo = s_start[c], hi = s_start[c + 1];
        for (int i = lo + 1; i < hi; i++) {
            const short v = s_items[i];
            int j = i - 1;
            while (j >= lo && s_items[j] > v) { s_items[j + 1] = s_items[j]; j--; }
            s_items[j + 1] = v;
        }
    }
    __syncthreads();
    for (int i = tid; i < total; i += T) cell_items[i] = s_items[i];
}

struct LocalMapDev {
    int n;
    const float *xyz, *normal, *min_dist, *max_dist;
    const uint32_t* desc;
};
struct PoseArgs { float T[12]; float Ow[3]; float cos_limit; int nlevels; };
struct MapFields {   // the MapPoint members isInFrustum writes, as device arrays (they are MapDev's inputs)
    uint8_t* track_in_view; float *proj_x, *proj_y, *proj_xr, *view_cos; int* level;
    uint8_t* track_in_view_host;   // second copy of the visibility flags in mapped host memory (the caller's mbTrackInView), may be null
};

// Frame::isInFrustum (src/Frame.cc:445-501) + the candidate collection of SearchByProjection for the same map point.
// fp32 projection left to right without FMA (cv::gemm on 3x3 . 3x1, pinned in tests/test_oracle_vs_cv2.py); cv::norm and
// Mat::dot accumulate in double.
__global__ void __launch_bounds__(256) frustum_collect_kernel(FrameDev F, LocalMapDev LM, PoseArgs P, const uint8_t* __restrict__ skip,
                                                              MapFields out, MapDev M, float th, const int* kp_state, CandLists C, float* proj_out) {
    COEB_MTRACE(2);
    pdl_launch_dependents();   // see m2_collect_kernel
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per map point: every lane evaluates the (cheap) frustum test
    if (i >= LM.n) return;
    const bool writer = (threadIdx.x & 31) == 0;
    bool in = false;
    float u = 0.f, v = 0.f, ur = 0.f, viewCos = 0.f;
    int lvl = 0;
    const float log_scale = P.nlevels > 1 ? glibc_logf(F.scale[1]) : 1.f;   // mfLogScaleFactor = log(mfScaleFactor) (src/Frame.cc:151)
    // every field of the map point is fetched up front (one round trip to L2 instead of one per nested test of isInFrustum: the kernel
    // is a chain of such round trips under 5000 resident warps)
    const bool skipped = skip[i] != 0, obs = M.has_obs[i] != 0;
    const float X = LM.xyz[3 * i], Y = LM.xyz[3 * i + 1], Z = LM.xyz[3 * i + 2];
    const float max_dist = LM.max_dist[i], min_dist = LM.min_dist[i];
    const float nx = LM.normal[3 * i], ny = LM.normal[3 * i + 1], nz = LM.normal[3 * i + 2];
    if (!skipped) {
        const float* T = P.T;
        const float PcX = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], X), __fmul_rn(T[1], Y)), __fmul_rn(T[2], Z)), T[3]);
        const float PcY = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[4], X), __fmul_rn(T[5], Y)), __fmul_rn(T[6], Z)), T[7]);
        const float PcZ = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[8], X), __fmul_rn(T[9], Y)), __fmul_rn(T[10], Z)), T[11]);
        if (!(PcZ < 0.0f)) {
            const float invz = __fdiv_rn(1.0f, PcZ);
            u = __fadd_rn(__fmul_rn(__fmul_rn(F.fx, PcX), invz), F.cx);
            v = __fadd_rn(__fmul_rn(__fmul_rn(F.fy, PcY), invz), F.cy);
            if (!(u < F.min_x || u > F.max_x) && !(v < F.min_y || v > F.max_y)) {
                const float maxD = __fmul_rn(1.2f, max_dist), minD = __fmul_rn(0.8f, min_dist);
                const float ox = __fsub_rn(X, P.Ow[0]), oy = __fsub_rn(Y, P.Ow[1]), oz = __fsub_rn(Z, P.Ow[2]);
                const double s = __dadd_rn(__dadd_rn(__dmul_rn((double)ox, (double)ox), __dmul_rn((double)oy, (double)oy)), __dmul_rn((double)oz, (double)oz));
                const float dist = (float)sqrt(s);
                if (!(dist < minD || dist > maxD)) {
                    const double dot = __dadd_rn(__dadd_rn(__dmul_rn((double)ox, (double)nx), __dmul_rn((double)oy, (double)ny)), __dmul_rn((double)oz, (double)nz));
                    viewCos = (float)(dot / (double)dist);
                    if (!(viewCos < P.cos_limit)) {
                        const float ratio = __fdiv_rn(max_dist, dist);
                        int nScale = 0;
                        if (ratio >= 1.17549435e-38f && ratio <= 3.402823466e+38f) nScale = (int)ceilf(__fdiv_rn(glibc_logf(ratio), log_scale));
                        lvl = nScale < 0 ? 0 : (nScale >= P.nlevels ? P.nlevels - 1 : nScale);
                        ur = __fsub_rn(u, __fmul_rn(F.bf, invz));
                        in = true;
                    }
                }
            }
        }
    }
    if (!in) { u = v = ur = viewCos = 0.f; lvl = 0; }
    if (writer) {
        out.track_in_view[i] = in ? 1 : 0;
        if (out.track_in_view_host) out.track_in_view_host[i] = in ? 1 : 0;
        out.proj_x[i] = u; out.proj_y[i] = v; out.proj_xr[i] = ur; out.view_cos[i] = viewCos; out.level[i] = lvl;
        if (proj_out) { float* q = proj_out + 5 * (size_t)i; q[0] = u; q[1] = v; q[2] = ur; q[3] = viewCos; q[4] = (float)lvl; }
    }
    // the collection of SearchByProjection for this map point, from the registers (m2_query's arithmetic on the values just written;
    // isBad() points arrive as `skip`)
    M2Query q;
    q.lvl = lvl;
    float r = ((double)viewCos > 0.998) ? 2.5f : 4.0f;   // RadiusByViewingCos (:131-137)
    if (th != 1.0f) r *= th;                              // bFactor (:49, :65-66)
    q.rs = r * F.scale[lvl];
    q.d = LM.desc + 8 * (size_t)i;
    q.pxr = ur; q.x = u; q.y = v;
    m2_collect_query(F, kp_state, C, i, in, q, obs);
}

// ---- search half of ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th) (src/ORBmatcher.cc:826-961) ---------------------------------
// One warp per map point: projection and the visibility tests (every lane, the arithmetic is a few dozen operations), then the
// window walk with the level and chi-square filters, Hamming distances lane-parallel and an argmin over (distance, traversal
// position). No query reads what another one writes: the reference's Replace / AddObservation side effects are the caller's.
__global__ void __launch_bounds__(256) fuse_search_kernel(FrameDev F, LocalMapDev LM, PoseArgs P, const uint8_t* __restrict__ valid, float th, int chi2_tests,
                                                          int* best_out) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= LM.n) return;
    const int lane = threadIdx.x & 31;
    int result = -1;
    if (valid[i]) {
        const float X0 = LM.xyz[3 * i], Y0 = LM.xyz[3 * i + 1], Z0 = LM.xyz[3 * i + 2];
        const float* T = P.T;
        const float X = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], X0), __fmul_rn(T[1], Y0)), __fmul_rn(T[2], Z0)), T[3]);
        const float Y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[4], X0), __fmul_rn(T[5], Y0)), __fmul_rn(T[6], Z0)), T[7]);
        const float Z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[8], X0), __fmul_rn(T[9], Y0)), __fmul_rn(T[10], Z0)), T[11]);
        bool ok = !(Z < 0.0f);
        const float invz = __fdiv_rn(1.0f, Z);
        const float u = __fadd_rn(__fmul_rn(F.fx, __fmul_rn(X, invz)), F.cx), v = __fadd_rn(__fmul_rn(F.fy, __fmul_rn(Y, invz)), F.cy);
        ok = ok && (u >= F.min_x && u < F.max_x && v >= F.min_y && v < F.max_y);   // KeyFrame::IsInImage
        const float ur = __fsub_rn(u, __fmul_rn(F.bf, invz));
        const float ox = __fsub_rn(X0, P.Ow[0]), oy = __fsub_rn(Y0, P.Ow[1]), oz = __fsub_rn(Z0, P.Ow[2]);
        const double s2 = __dadd_rn(__dadd_rn(__dmul_rn((double)ox, (double)ox), __dmul_rn((double)oy, (double)oy)), __dmul_rn((double)oz, (double)oz));
        const float dist3D = (float)sqrt(s2);
        ok = ok && !(dist3D < __fmul_rn(0.8f, LM.min_dist[i]) || dist3D > __fmul_rn(1.2f, LM.max_dist[i]));
        const double dot = __dadd_rn(__dadd_rn(__dmul_rn((double)ox, (double)LM.normal[3 * i]), __dmul_rn((double)oy, (double)LM.normal[3 * i + 1])),
                                     __dmul_rn((double)oz, (double)LM.normal[3 * i + 2]));
        ok = ok && !(dot < __dmul_rn(0.5, (double)dist3D));
        if (ok) {
            const float ratio = __fdiv_rn(LM.max_dist[i], dist3D);
            int lvl = 0;
            if (ratio >= 1.17549435e-38f && ratio <= 3.402823466e+38f) lvl = (int)ceilf(__fdiv_rn(glibc_logf(ratio), P.nlevels > 1 ? glibc_logf(F.scale[1]) : 1.f));
            lvl = lvl < 0 ? 0 : (lvl >= P.nlevels ? P.nlevels - 1 : lvl);
            const float radius = __fmul_rn(th, F.scale[lvl]);
            const uint32_t* d = LM.desc + 8 * (size_t)i;
            int key = kInf, myidx = -1;   // lane-local best: (distance << 16 | traversal position)
            warp_for_each_in_area(
                F, u, v, radius, -1, -1,
                [&](int idx) {
                    const int kl = F.octave[idx];
                    if (kl < lvl - 1 || kl > lvl) return false;
                    if (!chi2_tests) return true;   // Fuse(pKF, Scw, ...) of loop closing has no reprojection-error test
                    const float sc = F.scale[kl];
                    const float inv = __fdiv_rn(1.0f, __fmul_rn(sc, sc));
                    const float ex = __fsub_rn(u, F.x[idx]), ey = __fsub_rn(v, F.y[idx]);
                    const float kr = F.uright ? F.uright[idx] : -1.f;
                    if (kr >= 0) {
                        const float er = __fsub_rn(ur, kr);
                        const float e2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(er, er));
                        return !((double)__fmul_rn(e2, inv) > 7.8);
                    }
                    const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                    return !((double)__fmul_rn(e2, inv) > 5.99);
                },
                [&](int pos, int idx) {
                    const int dist = hamming256(d, F.desc + 8 * (size_t)idx);
                    const int k = (dist << 16) | min(pos, 0xFFFF);
                    if (dist < 256 && k < key) { key = k; myidx = idx; }
                });
            const int best = __reduce_min_sync(0xffffffffu, key);
            const unsigned who = __ballot_sync(0xffffffffu, key == best && key != kInf);
            if (who) {
                const int idx = __shfl_sync(0xffffffffu, myidx, __ffs(who) - 1);
                if ((best >> 16) <= COEB_TH_LOW) result = idx;
            }
        }
    }
    if (lane == 0) best_out[i] = result;
}

// ---- SearchByBoW (src/ORBmatcher.cc:158-288 and :522-655): matching restricted to features of the same vocabulary node --
// The node lists (DBoW2::FeatureVector) arrive as CSR; the host intersects the two sorted node-id arrays and hands over
// one {start1, end1, start2, end2} record per common node. A feature belongs to exactly one node, so the loop-carried
// state of the reference (a feature of F2 matched by an earlier query is skipped) never crosses a node: one warp owns a
// node and walks its queries in list order, exactly like the reference; the 32 lanes share the candidate loop.
struct BowArgs {
    int n_common;
    const int4* nodes;             // {start1, end1, start2, end2} into items1 / items2
    const int *items1, *items2;
    const uint8_t *valid1, *valid2;   // valid2 may be nullptr (Frame overload)
    float nnratio;
    int strict_low;
};

__global__ void __launch_bounds__(128) bow_node_kernel(FrameDev F1, FrameDev F2, BowArgs a, int* match12, int* matched2) {
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= a.n_common) return;
    const int4 nd = a.nodes[w];
    for (int p1 = nd.x; p1 < nd.y; p1++) {
        const int idx1 = a.items1[p1];
        if (!a.valid1[idx1]) continue;   // warp-uniform
        const uint32_t* d1 = F1.desc + 8 * (size_t)idx1;
        // lane-local scan of candidates lane, lane+32, ... in list order: strict '<' keeps the earliest position
        int b1 = 256, pos1 = 0x7fffffff, i2best = -1, b2 = 256;
        for (int p2 = nd.z + lane; p2 < nd.w; p2 += 32) {
            const int idx2 = a.items2[p2];
            if (matched2[idx2] || (a.valid2 && !a.valid2[idx2])) continue;
            const int dist = hamming256(d1, F2.desc + 8 * (size_t)idx2);
            if (dist < b1) { b2 = b1; b1 = dist; pos1 = p2; i2best = idx2; }
            else if (dist < b2) b2 = dist;
        }
        // warp merge: best = lexicographic min of (dist, list position); second = smallest of the remaining values
        int gb = b1, gp = pos1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, gb, o), op = __shfl_xor_sync(0xffffffffu, gp, o);
            if (ob < gb || (ob == gb && op < gp)) { gb = ob; gp = op; }
        }
        const bool winner = (pos1 == gp) && (gp != 0x7fffffff);
        int second = winner ? b2 : b1;   // the winner lane contributes its own runner-up, every other lane its best
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) second = min(second, __shfl_xor_sync(0xffffffffu, second, o));
        const unsigned wm = __ballot_sync(0xffffffffu, winner);
        const int gi = __shfl_sync(0xffffffffu, i2best, wm ? __ffs(wm) - 1 : 0);
        const bool low = a.strict_low ? gb < COEB_TH_LOW : gb <= COEB_TH_LOW;
        if (wm && low && (float)gb < a.nnratio * (float)second) {
            if (lane == 0) { match12[idx1] = gi; matched2[gi] = 1; }
            __syncwarp();   // the claim is visible to the next query's candidate loop (same warp)
        }
    }
}

// ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:657-824): same node walk without loop-carried state (the
// `vbMatched2[bestIdx2] = true` of the reference is commented out, :765). `dist > bestDist` skips, so among the
// candidates that pass the geometric tests the smallest distance wins and, on a tie, the LAST one in list order.
struct TriArgs {
    float F12[9];
    float ex, ey;
    int only_stereo;
};

__device__ __forceinline__ bool check_dist_epipolar_line(float x1, float y1, float x2, float y2, const float* F12, float sigma2) {   // :140-156
    const float a = __fadd_rn(__fadd_rn(__fmul_rn(x1, F12[0]), __fmul_rn(y1, F12[3])), F12[6]);
    const float b = __fadd_rn(__fadd_rn(__fmul_rn(x1, F12[1]), __fmul_rn(y1, F12[4])), F12[7]);
    const float c = __fadd_rn(__fadd_rn(__fmul_rn(x1, F12[2]), __fmul_rn(y1, F12[5])), F12[8]);
    const float num = __fadd_rn(__fadd_rn(__fmul_rn(a, x2), __fmul_rn(b, y2)), c);
    const float den = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
    if (den == 0.f) return false;
    const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
    return (double)dsqr < __dmul_rn(3.84, (double)sigma2);
}

__global__ void __launch_bounds__(128) triangulation_node_kernel(FrameDev F1, FrameDev F2, BowArgs a, TriArgs t, int* match12) {
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= a.n_common) return;
    const int4 nd = a.nodes[w];
    for (int p1 = nd.x; p1 < nd.y; p1++) {
        const int idx1 = a.items1[p1];
        if (!a.valid1[idx1]) continue;   // pMP1 exists (:697-701)
        const bool stereo1 = F1.uright && F1.uright[idx1] >= 0.f;
        if (t.only_stereo && !stereo1) continue;
        const uint32_t* d1 = F1.desc + 8 * (size_t)idx1;
        const float x1 = F1.x[idx1], y1 = F1.y[idx1];
        int best = COEB_TH_LOW, pos = -1, i2best = -1;
        for (int p2 = nd.z + lane; p2 < nd.w; p2 += 32) {
            const int idx2 = a.items2[p2];
            if (!a.valid2[idx2]) continue;
            const bool stereo2 = F2.uright && F2.uright[idx2] >= 0.f;
            if (t.only_stereo && !stereo2) continue;
            const int dist = hamming256(d1, F2.desc + 8 * (size_t)idx2);
            if (dist > COEB_TH_LOW || dist > best) continue;
            const float x2 = F2.x[idx2], y2 = F2.y[idx2], sc = F2.scale[F2.octave[idx2]];
            if (!stereo1 && !stereo2) {
                const float dxe = __fsub_rn(t.ex, x2), dye = __fsub_rn(t.ey, y2);
                if (__fadd_rn(__fmul_rn(dxe, dxe), __fmul_rn(dye, dye)) < __fmul_rn(100.f, sc)) continue;
            }
            if (check_dist_epipolar_line(x1, y1, x2, y2, t.F12, __fmul_rn(sc, sc))) { best = dist; pos = p2; i2best = idx2; }
        }
        // warp merge: smallest distance, then the largest list position
        int gb = pos >= 0 ? best : 0x7fffffff, gp = pos;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, gb, o), op = __shfl_xor_sync(0xffffffffu, gp, o);
            if (ob < gb || (ob == gb && op > gp)) { gb = ob; gp = op; }
        }
        if (gp >= 0 && pos == gp) match12[idx1] = i2best;   // exactly one lane
    }
}

// Rotation-consistency check of SearchByBoW (:275-293, :632-652) + the match count. One CTA.
__global__ void __launch_bounds__(1024) bow_finish_kernel(FrameDev F1, FrameDev F2, int check_ori, int* match12, int* out_info) {
    __shared__ int s_hist[COEB_HISTO_LENGTH];
    __shared__ int s_keep[3], s_count;
    const int tid = threadIdx.x, T = blockDim.x;
    if (tid < COEB_HISTO_LENGTH) s_hist[tid] = 0;
    if (tid == 0) s_count = 0;
    __syncthreads();
    if (check_ori) {
        for (int i = tid; i < F1.n; i += T) {
            const int j = match12[i];
            if (j >= 0) atomicAdd(&s_hist[rot_bin(F1.angle[i], F2.angle[j])], 1);
        }
        __syncthreads();
        if (tid == 0) three_maxima(s_hist, COEB_HISTO_LENGTH, s_keep[0], s_keep[1], s_keep[2]);
        __syncthreads();
    }
    int mine = 0;
    for (int i = tid; i < F1.n; i += T) {
        const int j = match12[i];
        if (j < 0) continue;
        if (check_ori) {
            const int b = rot_bin(F1.angle[i], F2.angle[j]);
            if (b != s_keep[0] && b != s_keep[1] && b != s_keep[2]) { match12[i] = -1; continue; }
        }
        mine++;
    }
    if (mine) atomicAdd(&s_count, mine);
    __syncthreads();
    if (tid == 0) out_info[0] = s_count;
}

}  // namespace coeb

// =====================================================================================================================
// C ABI (host side)
//
// Every call packs its inputs into ONE pinned staging block, issues one host->device copy, the kernel(s) and one
// device->host copy of the packed results: the tracking thread calls these once per frame, so the fixed cost per
// call (copies, launches, one synchronisation) is what matters, not bandwidth.
// =====================================================================================================================
using namespace coeb;

namespace {

// Pinned host block with a device mirror, grown geometrically.
struct Stage {
    char* h = nullptr;
    char* d = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return COEB_OK;
        size_t want = std::max<size_t>(bytes + bytes / 2, 1 << 16);
        if (h) cudaFreeHost(h);
        if (d) cudaFree(d);
        h = d = nullptr;
        cap = 0;
        CUDA_TRY(cudaHostAlloc((void**)&h, want, cudaHostAllocDefault));
        CUDA_TRY(cudaMalloc((void**)&d, want));
        cap = want;
        return COEB_OK;
    }
    // Results the kernels write exactly once (match tables, counts) go straight into mapped pinned host memory (d == h): the
    // posted writes cost the call nothing, a device-to-host copy behind the last kernel costs it 8-10 us.
    int reserve_mapped(size_t bytes) {
        if (bytes <= cap) return COEB_OK;
        size_t want = std::max<size_t>(bytes + bytes / 2, 1 << 16);
        if (h) cudaFreeHost(h);
        h = d = nullptr;
        cap = 0;
        CUDA_TRY(cudaHostAlloc((void**)&h, want, cudaHostAllocMapped));
        d = h;
        cap = want;
        return COEB_OK;
    }
    void release() {
        if (h) cudaFreeHost(h);
        if (d && d != h) cudaFree(d);
        h = d = nullptr;
        cap = 0;
    }
};

inline size_t al(size_t b) { return (b + 255) & ~(size_t)255; }

// Lays arrays out in a Stage: place() copies host data in and returns the device address it will have.
struct Packer {
    Stage& st;
    size_t off = 0;
    explicit Packer(Stage& s) : st(s) {}
    template <typename T> T* place(const T* src, size_t n) {
        T* dev = reinterpret_cast<T*>(st.d + off);
        if (n && src) std::memcpy(st.h + off, src, n * sizeof(T));
        off += al(n * sizeof(T));
        return dev;
    }
    template <typename T> T* host_at(const T* dev) { return reinterpret_cast<T*>(st.h + (reinterpret_cast<const char*>(dev) - st.d)); }
    template <typename T> T* room(size_t n) {   // uninitialised device space (outputs / scratch)
        T* dev = reinterpret_cast<T*>(st.d + off);
        off += al(n * sizeof(T));
        return dev;
    }
};

}  // namespace

struct coeb_matcher {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    Stage in, out;                                   // inputs (H2D) and results (D2H)
    Stage outm;                                      // results written once by the last kernel: mapped pinned memory, no copy
    Stage inm;                                       // small inputs read once by the first kernel: mapped pinned memory, no copy
    int* d_meta = nullptr;                           // CandLists::meta: zero between calls (the resolve kernels clear it again)
    uint8_t* d_zero = nullptr; size_t zero_bytes = 0;   // all-zero flags (isBad() of a resident local map)
    void* d_scratch = nullptr; size_t scratch_bytes = 0;
    void* d_in = nullptr; size_t in_bytes = 0;        // kNN partials
    void* d_depth = nullptr; size_t depth_bytes = 0;  // uploaded depth map of coeb_frame_from_extractor
    cudaEvent_t ev_ex = nullptr;                       // orders the matcher's stream after an extractor's
};

struct coeb_frame {
    coeb_matcher* m = nullptr;   // the matcher that built the frame; only dereferenced while it is registered as alive
    int device = 0;
    int n = 0, nlevels = 0;
    void* block = nullptr;
    size_t block_bytes = 0;
    FrameDev dev{};
};

struct coeb_local_map {
    coeb_matcher* m = nullptr;
    int device = 0;
    int n = 0;
    void* block = nullptr;
    coeb::LocalMapDev dev{};
};

// The extractor handle is opaque here; coeb_api.cu exports the pyramid accessor used by the stereo matcher.
extern "C" int coeb_pyramid_level(coeb_extractor* ex, int frame, int level, int blurred, const uint8_t** dev_ptr, int* width,
                                  int* height, int* pitch);
extern "C" int coeb_extractor_tables(const coeb_extractor* ex, int* nlevels, float* scale, float* inv_scale, float* sigma2,
                                     float* inv_sigma2, int* features_per_level);
extern "C" int coeb_extractor_device_stream(coeb_extractor* ex, int* device, void** stream);

namespace {

// Frames outlive the call that made them and, in ORB-SLAM, the thread: a Frame / KeyFrame that owns a device twin is handed from
// Tracking to LocalMapping and LoopClosing and destroyed there, while the matcher is thread-local to Tracking. So the pool of
// recycled frame blocks and the set of live matchers are process-wide and locked; coeb_frame_destroy never touches the matcher.
struct FrameBlocks {
    std::mutex mu;
    std::vector<std::pair<size_t, void*>> pool[16];   // per device: blocks of destroyed frames, reused by the next ones
    std::set<const coeb_matcher*> alive;
    int matchers_on[16] = {0};
    void* take(int device, size_t need, size_t* bytes) {
        std::lock_guard<std::mutex> g(mu);
        auto& v = pool[device & 15];
        for (size_t i = 0; i < v.size(); i++)
            if (v[i].first >= need) {
                void* p = v[i].second;
                *bytes = v[i].first;
                v.erase(v.begin() + i);
                return p;
            }
        return nullptr;
    }
    void give(int device, size_t bytes, void* p) {
        {
            std::lock_guard<std::mutex> g(mu);
            auto& v = pool[device & 15];
            if (matchers_on[device & 15] > 0 && v.size() < 8) { v.push_back({bytes, p}); return; }
        }
        int cur = -1;
        cudaGetDevice(&cur);
        cudaSetDevice(device);
        cudaFree(p);
        if (cur >= 0) cudaSetDevice(cur);
    }
    void matcher_created(const coeb_matcher* m, int device) {
        std::lock_guard<std::mutex> g(mu);
        alive.insert(m);
        matchers_on[device & 15]++;
    }
    void matcher_destroyed(const coeb_matcher* m, int device) {   // the caller has made `device` current
        std::vector<std::pair<size_t, void*>> drop;
        {
            std::lock_guard<std::mutex> g(mu);
            alive.erase(m);
            if (--matchers_on[device & 15] <= 0) drop.swap(pool[device & 15]);
        }
        for (auto& b : drop) cudaFree(b.second);
    }
    bool is_alive(const coeb_matcher* m) {
        std::lock_guard<std::mutex> g(mu);
        return alive.count(m) != 0;
    }
};
FrameBlocks g_frames;

int grow(void** p, size_t* cap, size_t bytes) {
    if (bytes <= *cap && *p) return COEB_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    CUDA_TRY(cudaMalloc(p, std::max<size_t>(bytes + bytes / 2, 1 << 16)));
    *cap = std::max<size_t>(bytes + bytes / 2, 1 << 16);
    return COEB_OK;
}

// Dynamic shared memory of a resolve kernel: the claim table (when it fits) followed by the cached active queries
// (bytes_per_query each; see the cached path of m2_resolve_kernel). Returns the byte count, sets *cache_cap.
size_t resolve_smem(size_t claim_bytes, int n_queries, int bytes_per_query, int* cache_cap) {
    const size_t budget = 200 * 1024, fixed = claim_bytes + 16;
    const size_t room = budget > fixed ? budget - fixed : 0;
    *cache_cap = (int)std::min<size_t>((size_t)std::max(n_queries, 0), room / bytes_per_query) & ~1;
    return fixed + (size_t)*cache_cap * bytes_per_query;
}

int push_inputs(coeb_matcher* m, const Packer& p) {
    if (p.off) CUDA_TRY(cudaMemcpyAsync(m->in.d, m->in.h, p.off, cudaMemcpyHostToDevice, m->stream));
    return COEB_OK;
}
// Candidate lists of n queries inside a scratch block: counts | active | meta | items. Returns the bytes used.
size_t lists_bytes(size_t n, int cap) { return 3 * al(n * 4) + al(8) + al(n * 16) + al(n * cap * 8); }
CandLists carve_lists(char* sc, size_t n, int cap, int* meta) {
    CandLists C{};
    C.count = (int*)sc; sc += al(n * 4);
    C.active = (int*)sc; sc += al(n * 4);
    C.meta = meta; sc += al(8);
    C.top = (int*)sc; sc += al(n * 4);
    C.rec = (int4*)sc; sc += al(n * 16);
    C.items = (int2*)sc;
    C.cap = cap;
    return C;
}

// The fixed-point resolve of SearchByProjection(F, MapPoints) behind a collect kernel already enqueued on m->stream. Maps whose
// queries all fit in shared memory (about 7000 map points) take m2_resolve_cached_kernel, launched programmatically dependent on the
// collect kernel; larger ones the general kernel. d_claim holds 3 * K ints.
bool m2_cached_fits(const coeb_frame* F, int n_map) {
    static const bool general = getenv("COEB_M2_GENERAL") != nullptr;   // development switch: always the general kernel
    const size_t claim3 = (size_t)F->n * 12 <= 48 * 1024 ? (size_t)F->n * 12 : 0;
    int cc = 0;
    resolve_smem(claim3, n_map + 1, 25, &cc);
    return cc >= n_map && !general;
}
// flags_dev / flags_host (optional, 16-byte aligned, flags_bytes rounded up to 16): copied device -> mapped host memory by the cached
// kernel; only passed when m2_cached_fits() says that kernel runs.
// Launches the cached resolve kernel (kMode 2 or 3) programmatically dependent on the collect kernel enqueued before it.
// The caller has checked m2_cached_fits(F, n_queries).
template <int kMode>
int launch_cached_resolve(coeb_matcher* m, const coeb_frame* F, int n_queries, float nnratio, int accept, const int* d_state, const CandLists& C, int* d_kpm,
                          int* d_claim, int* d_info, const uint8_t* flags_dev, uint8_t* flags_host, size_t flags_bytes, const RotArgs& rot) {
    static bool big_smem[64] = {};   // opt-in above 48 KB of dynamic shared memory: a per-device function attribute, set once
    const int dv = m->device & 63;
    const size_t claim3 = (size_t)F->n * 12 <= 48 * 1024 ? (size_t)F->n * 12 : 0;
    int cc = 0;
    const size_t sm3 = resolve_smem(claim3, n_queries + 1, 25, &cc);
    const uint4* fd = reinterpret_cast<const uint4*>(flags_dev);
    uint4* fh = reinterpret_cast<uint4*>(flags_host);
    const int fv = flags_dev && flags_host ? (int)((flags_bytes + 15) / 16) : 0;
    if (!big_smem[dv]) {
        CUDA_TRY(cudaFuncSetAttribute(m2_resolve_cached_kernel<true, kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(m2_resolve_cached_kernel<false, kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
        big_smem[dv] = true;
    }
    cudaLaunchConfig_t cfg = {};
    // the shared-memory copies of the short lists of queued queries (kSl entries each) take what is left of the budget
    const size_t lists_at = (sm3 + 7 + 8) & ~(size_t)7, budget = 204 * 1024;
    const int slow_cap = (int)std::min<size_t>(1024, budget > lists_at ? (budget - lists_at) / (kSl * 8) : 0) & ~31;
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = lists_at + (size_t)slow_cap * kSl * 8; cfg.stream = m->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (claim3) CUDA_TRY(cudaLaunchKernelEx(&cfg, m2_resolve_cached_kernel<true, kMode>, F->n, n_queries, nnratio, accept, d_state, C, d_kpm, d_claim, d_info, cc, slow_cap, fd, fh, fv, rot));
    else CUDA_TRY(cudaLaunchKernelEx(&cfg, m2_resolve_cached_kernel<false, kMode>, F->n, n_queries, nnratio, accept, d_state, C, d_kpm, d_claim, d_info, cc, slow_cap, fd, fh, fv, rot));
    return COEB_OK;
}

int launch_m2_resolve(coeb_matcher* m, const coeb_frame* F, const MapDev& M, float th, float nnratio, const int* d_state, const CandLists& C, int* d_kpm,
                      int* d_res, int* d_claim, int* d_info, const uint8_t* flags_dev = nullptr, uint8_t* flags_host = nullptr, size_t flags_bytes = 0) {
    if (m2_cached_fits(F, M.n)) return launch_cached_resolve<2>(m, F, M.n, nnratio, COEB_TH_HIGH, d_state, C, d_kpm, d_claim, d_info, flags_dev, flags_host, flags_bytes, RotArgs{});
    static bool big_smem[64] = {};   // opt-in above 48 KB of dynamic shared memory: a per-device function attribute, set once
    const int dv = m->device & 63;
    int cc = 0;
    const size_t claim_smem = (size_t)F->n * 4 <= 32 * 1024 ? (size_t)F->n * 4 : 0;   // claim table in shared memory when it fits
    const size_t sm = resolve_smem(claim_smem, M.n + 1, 25, &cc);
    if (!big_smem[dv]) {
        CUDA_TRY(cudaFuncSetAttribute(m2_resolve_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(m2_resolve_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
        big_smem[dv] = true;
    }
    if (claim_smem) m2_resolve_kernel<true, true><<<1, 1024, sm, m->stream>>>(F->dev, M, th, nnratio, d_state, C, d_kpm, d_res, d_claim, d_info, cc);
    else m2_resolve_kernel<true, false><<<1, 1024, sm, m->stream>>>(F->dev, M, th, nnratio, d_state, C, d_kpm, d_res, d_claim, d_info, cc);
    CUDA_TRY(cudaGetLastError());
    return COEB_OK;
}

void print_mtrace(const char* what);
int sync_outputs(coeb_matcher* m) {   // results in m->outm: the kernels wrote them into host memory themselves
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return COEB_OK;
}

int pull_outputs(coeb_matcher* m, size_t bytes) {
    if (bytes) CUDA_TRY(cudaMemcpyAsync(m->out.h, m->out.d, bytes, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return COEB_OK;
}

}  // namespace

extern "C" {

int coeb_matcher_create(int device, coeb_matcher** out) {
    if (!out) return fail(COEB_ERR_INVALID_ARG, "null argument");
    int st = check_device(device);
    if (st != COEB_OK) return st;
    CUDA_TRY(cudaSetDevice(device));
    coeb_matcher* m = new coeb_matcher();
    m->device = device;
    if (cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete m; return fail(COEB_ERR_CUDA, "cudaStreamCreate failed"); }
    m->stream = m->own_stream;
    if (cudaMalloc((void**)&m->d_meta, 256) != cudaSuccess || cudaMemset(m->d_meta, 0, 256) != cudaSuccess) {
        cudaStreamDestroy(m->own_stream);
        delete m;
        return fail(COEB_ERR_CUDA, "matcher scratch allocation failed");
    }
    g_frames.matcher_created(m, device);
    *out = m;
    return COEB_OK;
}

void coeb_matcher_destroy(coeb_matcher* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaStreamSynchronize(m->stream);
    m->in.release();
    m->out.release();
    m->outm.release();
    m->inm.release();
    cudaFree(m->d_meta);
    cudaFree(m->d_zero);
    cudaFree(m->d_scratch);
    cudaFree(m->d_in);
    cudaFree(m->d_depth);
    if (m->ev_ex) cudaEventDestroy(m->ev_ex);
    g_frames.matcher_destroyed(m, m->device);
    cudaStreamDestroy(m->own_stream);
    delete m;
}

int coeb_matcher_set_stream(coeb_matcher* m, void* cuda_stream) {
    if (!m) return fail(COEB_ERR_INVALID_ARG, "null matcher");
    m->stream = cuda_stream ? (cudaStream_t)cuda_stream : m->own_stream;
    return COEB_OK;
}

int coeb_frame_create(coeb_matcher* m, const coeb_keypoint* kps, const uint8_t* desc, int n, const float* uright,
                      const coeb_camera* cam, const float* scale_factors, int nlevels, coeb_frame** out) {
    if (!m || !out || !cam || !scale_factors || n < 0 || (n > 0 && (!kps || !desc))) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (nlevels < 1 || nlevels > COEB_MAX_LEVELS) return fail(COEB_ERR_INVALID_ARG, "nlevels %d", nlevels);
    CUDA_TRY(cudaSetDevice(m->device));
    const size_t nn = std::max(n, 1);
    // one device block per frame: x, y, angle, octave, uright, desc, cell_start, cell_items, kp_cell
    const size_t need = 5 * al(nn * 4) + al(nn * 32) + al((kGridCells + 1) * 4) + 2 * al(nn * 4);
    coeb_frame* f = new coeb_frame();
    f->m = m; f->device = m->device; f->n = n; f->nlevels = nlevels;
    f->block = g_frames.take(m->device, need, &f->block_bytes);
    if (!f->block) {
        const size_t want = std::max<size_t>(need + need / 4, 1 << 17);
        if (cudaMalloc(&f->block, want) != cudaSuccess) { delete f; return fail(COEB_ERR_CUDA, "cudaMalloc(%zu) failed", want); }
        f->block_bytes = want;
    }
    int st = m->in.reserve(4 * al(nn * 4) + al(nn * 4) + al(nn * 32));
    if (st != COEB_OK) { coeb_frame_destroy(f); return st; }
    // pack SoA in pinned memory in the same order as the device block, then one copy
    char* base = (char*)f->block;
    size_t off = 0;
    auto carve = [&](size_t bytes) { char* p = base + off; off += al(bytes); return p; };
    float* d_x = (float*)carve(nn * 4); float* d_y = (float*)carve(nn * 4); float* d_angle = (float*)carve(nn * 4);
    int* d_octave = (int*)carve(nn * 4); float* d_uright = (float*)carve(nn * 4); uint32_t* d_desc = (uint32_t*)carve(nn * 32);
    const size_t upload_bytes = off;
    int* d_cell_start = (int*)carve((kGridCells + 1) * 4); int* d_cell_items = (int*)carve(nn * 4); int* d_kp_cell = (int*)carve(nn * 4);
    {
        char* h = m->in.h;
        float* hx = (float*)h; float* hy = (float*)(h + al(nn * 4)); float* ha = (float*)(h + 2 * al(nn * 4));
        int* ho = (int*)(h + 3 * al(nn * 4)); float* hu = (float*)(h + 4 * al(nn * 4)); uint8_t* hd = (uint8_t*)(h + 5 * al(nn * 4));
        for (int i = 0; i < n; i++) { hx[i] = kps[i].x; hy[i] = kps[i].y; ha[i] = kps[i].angle; ho[i] = kps[i].octave; hu[i] = uright ? uright[i] : -1.f; }
        if (n) std::memcpy(hd, desc, (size_t)n * 32);
    }
    cudaStream_t s = m->stream;
    cudaError_t e = cudaMemcpyAsync(base, m->in.h, upload_bytes, cudaMemcpyHostToDevice, s);
    FrameDev& d = f->dev;
    d.n = n; d.x = d_x; d.y = d_y; d.angle = d_angle; d.octave = d_octave; d.desc = d_desc; d.uright = uright ? d_uright : nullptr;
    d.cell_start = d_cell_start; d.cell_items = d_cell_items;
    d.min_x = cam->min_x; d.min_y = cam->min_y; d.max_x = cam->max_x; d.max_y = cam->max_y;
    d.gw_inv = (float)COEB_GRID_COLS / (cam->max_x - cam->min_x);   // mfGridElementWidthInv (src/Frame.cc:233)
    d.gh_inv = (float)COEB_GRID_ROWS / (cam->max_y - cam->min_y);
    d.fx = cam->fx; d.fy = cam->fy; d.cx = cam->cx; d.cy = cam->cy; d.bf = cam->bf; d.b = cam->b;
    for (int i = 0; i < COEB_MAX_LEVELS; i++) d.scale[i] = i < nlevels ? scale_factors[i] : 0.f;
    if (n <= kGridSmemMax) grid_build_smem_kernel<<<1, 1024, 0, s>>>(d, d_cell_start, d_cell_items);
    else grid_build_kernel<<<1, 1024, 0, s>>>(d, d_cell_start, d_cell_items, d_kp_cell);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);  // the pinned staging block is reused by the next call
    if (e != cudaSuccess) { coeb_frame_destroy(f); return fail(COEB_ERR_CUDA, "frame upload / grid build failed: %s", cudaGetErrorString(e)); }
    *out = f;
    return COEB_OK;
}

void coeb_frame_destroy(coeb_frame* f) {
    if (!f) return;
    if (f->block) g_frames.give(f->device, f->block_bytes, f->block);   // any thread; the owning matcher may be gone
    delete f;
}

int coeb_frame_features_in_area(coeb_frame* f, float x, float y, float r, int min_level, int max_level, int* idx_out, int cap,
                                int* n_out) {
    if (!f || !n_out || cap < 0) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (!g_frames.is_alive(f->m)) return fail(COEB_ERR_INVALID_ARG, "the matcher that built this frame has been destroyed");
    coeb_matcher* m = f->m;
    CUDA_TRY(cudaSetDevice(m->device));
    int st = m->out.reserve((size_t)(cap + 1) * 4 + 256);
    if (st != COEB_OK) return st;
    int* d_n = (int*)m->out.d;
    int* d_out = d_n + 1;
    features_in_area_kernel<<<1, 32, 0, m->stream>>>(f->dev, x, y, r, min_level, max_level, d_out, cap, d_n);
    CUDA_TRY(cudaGetLastError());
    if ((st = pull_outputs(m, (size_t)(cap + 1) * 4)) != COEB_OK) return st;
    const int* h = (const int*)m->out.h;
    *n_out = h[0];
    const int n = std::min(*n_out, cap);
    if (n > 0 && idx_out) std::memcpy(idx_out, h + 1, (size_t)n * 4);
    return *n_out > cap ? fail(COEB_ERR_CAPACITY, "need %d entries", *n_out) : COEB_OK;
}

int coeb_hamming256(const void* a, const void* b) {
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t x, y;
        std::memcpy(&x, (const char*)a + 4 * i, 4);
        std::memcpy(&y, (const char*)b + 4 * i, 4);
        dist += __builtin_popcount(x ^ y);
    }
    return dist;
}

int coeb_hamming256_batch(coeb_matcher* m, const uint8_t* a, const uint8_t* b, int n, int* dist_out) {
    if (!m || !a || !b || !dist_out || n < 0) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (n == 0) return COEB_OK;
    CUDA_TRY(cudaSetDevice(m->device));
    int st;
    if ((st = m->in.reserve(2 * al((size_t)n * 32))) != COEB_OK || (st = m->out.reserve((size_t)n * 4)) != COEB_OK) return st;
    Packer p(m->in);
    const uint32_t* da = (const uint32_t*)p.place(a, (size_t)n * 32);
    const uint32_t* db = (const uint32_t*)p.place(b, (size_t)n * 32);
    if ((st = push_inputs(m, p)) != COEB_OK) return st;
    hamming_pairs_kernel<<<(n + 255) / 256, 256, 0, m->stream>>>(da, db, n, (int*)m->out.d);
    CUDA_TRY(cudaGetLastError());
    if ((st = pull_outputs(m, (size_t)n * 4)) != COEB_OK) return st;
    std::memcpy(dist_out, m->out.h, (size_t)n * 4);
    return COEB_OK;
}

int coeb_match_projection(coeb_matcher* m, coeb_frame* F, int n, const uint8_t* track_in_view, const uint8_t* bad,
                          const uint8_t* has_obs, const float* proj_x, const float* proj_y, const float* proj_xr,
                          const int* level, const float* view_cos, const uint8_t* desc, float th, float nnratio,
                          int* kp_match, int* nmatches_out) {
    if (!m || !F || n < 0 || !kp_match) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (nmatches_out) *nmatches_out = 0;
    if (n == 0 || F->n == 0) return COEB_OK;
    if (!track_in_view || !bad || !has_obs || !proj_x || !proj_y || !proj_xr || !level || !view_cos || !desc)
        return fail(COEB_ERR_INVALID_ARG, "null map-point array");
    {   // (one branch-free pass the compiler vectorises; the offending point is looked up only if there is one)
        const unsigned nl = (unsigned)F->nlevels;
        unsigned any = 0;
        for (int i = 0; i < n; i++) any |= (unsigned)(track_in_view[i] != 0) & (unsigned)(bad[i] == 0) & (unsigned)((unsigned)level[i] >= nl);
        if (any)
            for (int i = 0; i < n; i++)
                if (track_in_view[i] && !bad[i] && (level[i] < 0 || level[i] >= F->nlevels)) return fail(COEB_ERR_INVALID_ARG, "map point %d: level %d", i, level[i]);
    }
    CUDA_TRY(cudaSetDevice(m->device));
    const bool trace = getenv("COEB_MATCH_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    const size_t N = n, K = F->n;
    int st;
    if ((st = m->in.reserve(3 * al(N) + 5 * al(N * 4) + al(N * 32) + al(K * 4))) != COEB_OK) return st;
    if ((st = m->outm.reserve_mapped(al(K * 4) + 256)) != COEB_OK) return st;
    const int cap = 32;   // candidates kept per map point; a fuller window falls back to the window-walking kernel
    const size_t claim_smem = (size_t)F->n * 4 <= 32 * 1024 ? (size_t)F->n * 4 : 0;   // claim table in shared memory when it fits
    if ((st = grow(&m->d_scratch, &m->scratch_bytes, al(N * 4) + al(K * 12) + lists_bytes(N, cap))) != COEB_OK) return st;
    Packer p(m->in);
    MapDev M{};
    M.n = n;
    M.track_in_view = p.place(track_in_view, N); M.bad = p.place(bad, N); M.has_obs = p.place(has_obs, N);
    M.proj_x = p.place(proj_x, N); M.proj_y = p.place(proj_y, N); M.proj_xr = p.place(proj_xr, N); M.view_cos = p.place(view_cos, N);
    M.level = p.place(level, N);
    M.desc = (const uint32_t*)p.place(desc, N * 32);
    const int* d_state = p.place(kp_match, K);
    const auto t_packed = std::chrono::steady_clock::now();
    if ((st = push_inputs(m, p)) != COEB_OK) return st;
    int* d_kpm = (int*)m->outm.d;                       // [K] then info[2]
    int* d_info = (int*)(m->outm.d + al(K * 4));
    int* d_res = (int*)m->d_scratch;
    int* d_claim = (int*)((char*)m->d_scratch + al(N * 4));
    const CandLists C = carve_lists((char*)m->d_scratch + al(N * 4) + al(K * 12), N, cap, m->d_meta);
    m2_collect_kernel<<<(n + 7) / 8, 256, 0, m->stream>>>(F->dev, M, th, d_state, C);
    CUDA_TRY(cudaGetLastError());
    if ((st = launch_m2_resolve(m, F, M, th, nnratio, d_state, C, d_kpm, d_res, d_claim, d_info)) != COEB_OK) return st;
    const auto t_queued = std::chrono::steady_clock::now();
    if ((st = sync_outputs(m)) != COEB_OK) return st;
    if (trace) {
        const auto t_done = std::chrono::steady_clock::now();
        auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
        fprintf(stderr, "[coeb match] SearchByProjection(map) host timeline: validate + pack %.1f us, enqueue %.1f us, wait %.1f us\n", us(t_begin, t_packed), us(t_packed, t_queued), us(t_queued, t_done));
    }
    if (((const int*)(m->outm.h + al(K * 4)))[2]) {   // a candidate list overflowed: exact window-walking variant
        if (claim_smem) m2_resolve_kernel<false, true><<<1, 1024, claim_smem + 16, m->stream>>>(F->dev, M, th, nnratio, d_state, C, d_kpm, d_res, d_claim, d_info, 0);
        else m2_resolve_kernel<false, false><<<1, 1024, 16, m->stream>>>(F->dev, M, th, nnratio, d_state, C, d_kpm, d_res, d_claim, d_info, 0);
        CUDA_TRY(cudaGetLastError());
        if ((st = sync_outputs(m)) != COEB_OK) return st;
    }
    std::memcpy(kp_match, m->outm.h, K * 4);
    if (nmatches_out) *nmatches_out = ((const int*)(m->outm.h + al(K * 4)))[0];
    print_mtrace("SearchByProjection(map)");
    if (getenv("COEB_MATCH_TRACE")) fprintf(stderr, "[coeb match] SearchByProjection(map): %d map points, %d keypoints, %d fixed-point rounds\n", n, F->n, ((const int*)(m->outm.h + al(K * 4)))[1]);
    return COEB_OK;
}

int coeb_match_lastframe(coeb_matcher* m, coeb_frame* cur, int n, const uint8_t* valid, const uint8_t* has_obs,
                         const float* xyz, const int* octave, const float* angle, const uint8_t* desc,
                         const float* Tcw_cur, const float* Tcw_last, float th, int mono, int check_ori,
                         int* kp_match, int* nmatches_out) {
    if (!m || !cur || n < 0 || !kp_match || !Tcw_cur || !Tcw_last) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (nmatches_out) *nmatches_out = 0;
    if (n == 0 || cur->n == 0) return COEB_OK;
    if (!valid || !has_obs || !xyz || !octave || !angle || !desc) return fail(COEB_ERR_INVALID_ARG, "null last-frame array");
    for (int i = 0; i < n; i++)
        if (valid[i] && (octave[i] < 0 || octave[i] >= cur->nlevels)) return fail(COEB_ERR_INVALID_ARG, "last-frame point %d: octave %d", i, octave[i]);
    CUDA_TRY(cudaSetDevice(m->device));
    const size_t N = n, K = cur->n;
    int st;
    if ((st = m->in.reserve(2 * al(N) + al(N * 12) + 2 * al(N * 4) + al(N * 32) + al(K * 4))) != COEB_OK) return st;
    if ((st = m->outm.reserve_mapped(al(K * 4) + 256)) != COEB_OK) return st;
    const int cap = 64;
    const size_t claim_smem = K * 4 <= 32 * 1024 ? K * 4 : 0;   // claim table in shared memory when it fits
    if ((st = grow(&m->d_scratch, &m->scratch_bytes, al(N * 4) + al(K * 12) + lists_bytes(N, cap))) != COEB_OK) return st;
    Packer p(m->in);
    LastDev L{};
    L.n = n;
    L.valid = p.place(valid, N); L.has_obs = p.place(has_obs, N); L.xyz = p.place(xyz, N * 3); L.octave = p.place(octave, N);
    L.angle = p.place(angle, N); L.desc = (const uint32_t*)p.place(desc, N * 32);
    const int* d_state = p.place(kp_match, K);
    if ((st = push_inputs(m, p)) != COEB_OK) return st;
    for (int i = 0; i < 12; i++) L.T[i] = Tcw_cur[i];
    // tlc = Rlw * twc + tlw with twc = -Rcw^T tcw (src/ORBmatcher.cc:1339-1350); only its z component is used. fp32, fixed order.
    float twc[3];
    for (int k = 0; k < 3; k++) twc[k] = -(Tcw_cur[0 * 4 + k] * Tcw_cur[3] + Tcw_cur[1 * 4 + k] * Tcw_cur[7] + Tcw_cur[2 * 4 + k] * Tcw_cur[11]);
    const float tz = Tcw_last[8] * twc[0] + Tcw_last[9] * twc[1] + Tcw_last[10] * twc[2] + Tcw_last[11];
    L.forward = (tz > cur->dev.b && !mono) ? 1 : 0;
    L.backward = (-tz > cur->dev.b && !mono) ? 1 : 0;
    L.max_accept = COEB_TH_HIGH;   // :1425
    int* d_kpm = (int*)m->outm.d;
    int* d_info = (int*)(m->outm.d + al(K * 4));
    int* d_res = (int*)m->d_scratch;
    int* d_claim = (int*)((char*)m->d_scratch + al(N * 4));
    const CandLists C = carve_lists((char*)m->d_scratch + al(N * 4) + al(K * 12), N, cap, m->d_meta);
    m3_collect_kernel<<<(n + 7) / 8, 256, 0, m->stream>>>(cur->dev, L, th, d_state, C);
    CUDA_TRY(cudaGetLastError());
    if (m2_cached_fits(cur, n)) {   // the latency form of the fixed-point iteration (m2_resolve_cached_kernel, kMode 3)
        RotArgs rot{check_ori, L.angle, cur->dev.angle};
        if ((st = launch_cached_resolve<3>(m, cur, n, 0.f, L.max_accept, d_state, C, d_kpm, d_claim, d_info, nullptr, nullptr, 0, rot)) != COEB_OK) return st;
    }
    else if (claim_smem) m3_resolve_kernel<true, true><<<1, 512, claim_smem, m->stream>>>(cur->dev, L, th, check_ori, d_state, C, d_kpm, d_res, d_claim, d_info);
    else m3_resolve_kernel<true, false><<<1, 512, 0, m->stream>>>(cur->dev, L, th, check_ori, d_state, C, d_kpm, d_res, d_claim, d_info);
    CUDA_TRY(cudaGetLastError());
    if ((st = sync_outputs(m)) != COEB_OK) return st;
    if (((const int*)(m->outm.h + al(K * 4)))[2]) {
        if (claim_smem) m3_resolve_kernel<false, true><<<1, 512, claim_smem, m->stream>>>(cur->dev, L, th, check_ori, d_state, C, d_kpm, d_res, d_claim, d_info);
    else m3_resolve_kernel<false, false><<<1, 512, 0, m->stream>>>(cur->dev, L, th, check_ori, d_state, C, d_kpm, d_res, d_claim, d_info);
        CUDA_TRY(cudaGetLastError());
        if ((st = sync_outputs(m)) != COEB_OK) return st;
    }
    std::memcpy(kp_match, m->outm.h, K * 4);
    if (nmatches_out) *nmatches_out = ((const int*)(m->outm.h + al(K * 4)))[0];
    print_mtrace("SearchByProjection(last frame)");
    return COEB_OK;
}

int coeb_match_init(coeb_matcher* m, coeb_frame* f1, coeb_frame* f2, float* prev_matched, int* matches12, int window_size,
                    float nnratio, int check_ori, int* nmatches_out) {
    if (!m || !f1 || !f2 || !prev_matched || !matches12) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (nmatches_out) *nmatches_out = 0;
    if (f1->n == 0) return COEB_OK;
    if (f2->n == 0) { for (int i = 0; i < f1->n; i++) matches12[i] = -1; return COEB_OK; }
    CUDA_TRY(cudaSetDevice(m->device));
    const size_t N1 = f1->n, N2 = f2->n;
    int st;
    if ((st = m->in.reserve(al(N1 * 8))) != COEB_OK) return st;
    if ((st = m->out.reserve(al(N1 * 4))) != COEB_OK) return st;
    if ((st = m->outm.reserve_mapped(al(N1 * 4) + al(N1 * 8) + 256)) != COEB_OK) return st;
    const int cap = 256;
    if ((st = grow(&m->d_scratch, &m->scratch_bytes, 2 * al(N1 * 4) + 2 * al((N2 + 1) * 4) + al(N1 * 8) + lists_bytes(N1, cap))) != COEB_OK) return st;
    Packer p(m->in);
    const float* d_prev = p.place(prev_matched, N1 * 2);
    if ((st = push_inputs(m, p)) != COEB_OK) return st;
    char* sc = (char*)m->d_scratch;
    int* d_res = (int*)sc; sc += al(N1 * 4);
    int* d_rdist = (int*)sc; sc += al(N1 * 4);
    int* d_cls = (int*)sc; sc += al((N2 + 1) * 4);
    int* d_clf = (int*)sc; sc += al((N2 + 1) * 4);
    int2* d_items = (int2*)sc; sc += al(N1 * 8);
    const CandLists C = carve_lists(sc, N1, cap, m->d_meta);
    int* d_m12 = (int*)m->out.d;                                  // read back by the kernel: device; its final state is mirrored below
    int* h_m12 = (int*)m->outm.d;                                 // results, written once: mapped pinned memory, no copy node
    float* d_prev_out = (float*)(m->outm.d + al(N1 * 4));
    int* d_info = (int*)(m->outm.d + al(N1 * 4) + al(N1 * 8));
    m4_collect_kernel<<<(f1->n + 7) / 8, 256, 0, m->stream>>>(f1->dev, f2->dev, d_prev, (float)window_size, C);
    m4_resolve_kernel<true><<<1, 1024, 0, m->stream>>>(f1->dev, f2->dev, d_prev, (float)window_size, nnratio, check_ori, C, d_res, d_rdist, d_cls,
                                                       d_clf, d_items, d_m12, d_prev_out, d_info, h_m12);
    CUDA_TRY(cudaGetLastError());
    if ((st = sync_outputs(m)) != COEB_OK) return st;
    if (((const int*)(m->outm.h + al(N1 * 4) + al(N1 * 8)))[2]) {
        m4_resolve_kernel<false><<<1, 1024, 0, m->stream>>>(f1->dev, f2->dev, d_prev, (float)window_size, nnratio, check_ori, C, d_res, d_rdist,
                                                            d_cls, d_clf, d_items, d_m12, d_prev_out, d_info, h_m12);
        CUDA_TRY(cudaGetLastError());
        if ((st = sync_outputs(m)) != COEB_OK) return st;
    }
    std::memcpy(matches12, m->outm.h, N1 * 4);
    std::memcpy(prev_matched, m->outm.h + al(N1 * 4), N1 * 8);
    if (nmatches_out) *nmatches_out = ((const int*)(m->outm.h + al(N1 * 4) + al(N1 * 8)))[0];
    print_mtrace("SearchForInitialization");
    return COEB_OK;
}

int coeb_stereo_match(coeb_matcher* m, coeb_extractor* left, coeb_extractor* right, int N, const coeb_keypoint* keys_left,
                      const uint8_t* desc_left, int Nr, const coeb_keypoint* keys_right, const uint8_t* desc_right, float bf,
                      float b, float* uright_out, float* depth_out, int* nmatched_out) {
    return coeb_stereo_match_frames(m, left, 0, right, 0, N, keys_left, desc_left, Nr, keys_right, desc_right, bf, b, uright_out, depth_out, nmatched_out);
}

int coeb_stereo_match_frames(coeb_matcher* m, coeb_extractor* left, int frame_left, coeb_extractor* right, int frame_right, int N,
                             const coeb_keypoint* keys_left, const uint8_t* desc_left, int Nr, const coeb_keypoint* keys_right,
                             const uint8_t* desc_right, float bf, float b, float* uright_out, float* depth_out, int* nmatched_out) {
    if (!m || !left || !right || N < 0 || Nr < 0 || !uright_out || !depth_out || frame_left < 0 || frame_right < 0) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (nmatched_out) *nmatched_out = 0;
    for (int i = 0; i < N; i++) { uright_out[i] = -1.f; depth_out[i] = -1.f; }
    if (N == 0 || Nr == 0) return COEB_OK;
    CUDA_TRY(cudaSetDevice(m->device));
    StereoDev S{};
    int nl = 0;
    int st = coeb_extractor_tables(left, &nl, S.scale, S.inv_scale, nullptr, nullptr, nullptr);
    if (st != COEB_OK) return st;
    for (int l = 0; l < nl; l++) {
        int w2 = 0, h2 = 0;
        if ((st = coeb_pyramid_level(left, frame_left, l, 0, &S.pyrL[l], &S.lw[l], &S.lh[l], &S.pitchL[l])) != COEB_OK) return st;
        if ((st = coeb_pyramid_level(right, frame_right, l, 0, &S.pyrR[l], &w2, &h2, &S.pitchR[l])) != COEB_OK) return st;
        if (w2 != S.lw[l] || h2 != S.lh[l]) return fail(COEB_ERR_INVALID_ARG, "left/right pyramids differ in size at level %d", l);
    }
    for (int i = 0; i < N; i++) if (keys_left[i].octave < 0 || keys_left[i].octave >= nl) return fail(COEB_ERR_INVALID_ARG, "left keypoint %d: octave", i);
    for (int i = 0; i < Nr; i++) if (keys_right[i].octave < 0 || keys_right[i].octave >= nl) return fail(COEB_ERR_INVALID_ARG, "right keypoint %d: octave", i);
    S.nRows = S.lh[0];
    S.N = N; S.Nr = Nr; S.bf = bf; S.b = b;
    const size_t NL = N, NR = Nr;
    if ((st = m->in.reserve(3 * al(NL * 4) + al(NL * 32) + 3 * al(NR * 4) + al(NR * 32))) != COEB_OK) return st;
    if ((st = m->out.reserve(2 * al(NL * 4) + 256)) != COEB_OK) return st;
    if ((st = grow(&m->d_scratch, &m->scratch_bytes, al(NL * 4))) != COEB_OK) return st;
    Packer p(m->in);
    float *dxl = p.room<float>(NL), *dyl = p.room<float>(NL); int* dol = p.room<int>(NL);
    const uint32_t* ddl = (const uint32_t*)p.place(desc_left, NL * 32);
    float *dxr = p.room<float>(NR), *dyr = p.room<float>(NR); int* dor_ = p.room<int>(NR);
    const uint32_t* ddr = (const uint32_t*)p.place(desc_right, NR * 32);
    {
        float *hx = p.host_at(dxl), *hy = p.host_at(dyl); int* ho = p.host_at(dol);
        for (int i = 0; i < N; i++) { hx[i] = keys_left[i].x; hy[i] = keys_left[i].y; ho[i] = keys_left[i].octave; }
        float *rx = p.host_at(dxr), *ry = p.host_at(dyr); int* ro = p.host_at(dor_);
        for (int i = 0; i < Nr; i++) { rx[i] = keys_right[i].x; ry[i] = keys_right[i].y; ro[i] = keys_right[i].octave; }
    }
    if ((st = push_inputs(m, p)) != COEB_OK) return st;
    S.xl = dxl; S.yl = dyl; S.octl = dol; S.descl = ddl; S.xr = dxr; S.yr = dyr; S.octr = dor_; S.descr = ddr;
    float* dur = (float*)m->out.d;
    float* ddp = (float*)(m->out.d + al(NL * 4));
    int* dinfo = (int*)(m->out.d + 2 * al(NL * 4));
    int* dsad = (int*)m->d_scratch;
    stereo_match_kernel<<<(N + 7) / 8, 256, 0, m->stream>>>(S, dur, ddp, dsad);
    if ((st = m->outm.reserve_mapped(2 * al(NL * 4) + 256)) != COEB_OK) return st;
    stereo_outlier_kernel<<<1, 1024, 0, m->stream>>>(N, dur, ddp, dsad, dinfo, (float*)m->outm.d, (float*)(m->outm.d + al(NL * 4)), (int*)(m->outm.d + 2 * al(NL * 4)));
    CUDA_TRY(cudaGetLastError());
    if ((st = sync_outputs(m)) != COEB_OK) return st;
    std::memcpy(uright_out, m->outm.h, NL * 4);
    std::memcpy(depth_out, m->outm.h + al(NL * 4), NL * 4);
    if (nmatched_out) *nmatched_out = ((const int*)(m->outm.h + 2 * al(NL * 4)))[0];
    return COEB_OK;
}

int coeb_knn2_device(coeb_matcher* m, const uint8_t* d_query, int nq, const uint8_t* d_train, int nt, float nnratio,
                     int* d_best_idx, int* d_d1, int* d_d2) {
    if (!m || !d_query || !d_train || nq < 1 || nt < 1 || !d_best_idx) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(m->device));
    const int qblocks = (nq + kKnnThreads * kKnnQ - 1) / (kKnnThreads * kKnnQ);
    int nchunks = std::max(1, (148 * 4 + qblocks - 1) / qblocks);          // ~4 CTAs per SM overall
    nchunks = std::min(nchunks, std::max(1, nt / 64));
    int chunk = (nt + nchunks - 1) / nchunks;
    nchunks = (nt + chunk - 1) / chunk;
    int st = grow(&m->d_in, &m->in_bytes, 3 * al((size_t)nchunks * nq * 4));
    if (st != COEB_OK) return st;
    int* p1 = (int*)m->d_in;
    int* pi = (int*)((char*)m->d_in + al((size_t)nchunks * nq * 4));
    int* p2 = (int*)((char*)m->d_in + 2 * al((size_t)nchunks * nq * 4));
    knn2_partial_kernel<<<dim3(qblocks, nchunks), kKnnThreads, 0, m->stream>>>((const uint32_t*)d_query, nq, (const uint32_t*)d_train, nt, chunk,
                                                                              p1, pi, p2);
    knn2_merge_kernel<<<(nq + 255) / 256, 256, 0, m->stream>>>(nq, nchunks, p1, pi, p2, nnratio, d_best_idx, d_d1, d_d2, nullptr);
    CUDA_TRY(cudaGetLastError());
    return COEB_OK;
}

int coeb_knn2(coeb_matcher* m, const uint8_t* query, int nq, const uint8_t* train, int nt, float nnratio, int* best_idx,
              int* d1, int* d2, int* naccepted_out) {
    if (!m || !query || !train || nq < 0 || nt < 0 || !best_idx) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (naccepted_out) *naccepted_out = 0;
    if (nq == 0) return COEB_OK;
    if (nt == 0) { for (int i = 0; i < nq; i++) { best_idx[i] = -1; if (d1) d1[i] = 256; if (d2) d2[i] = 256; } return COEB_OK; }
    CUDA_TRY(cudaSetDevice(m->device));
    const size_t NQ = nq, NT = nt;
    int st;
    if ((st = m->in.reserve(al(NQ * 32) + al(NT * 32))) != COEB_OK || (st = m->out.reserve(3 * al(NQ * 4))) != COEB_OK) return st;
    Packer p(m->in);
    const uint8_t* dq = p.place(query, NQ * 32);
    const uint8_t* dt = p.place(train, NT * 32);
    if ((st = push_inputs(m, p)) != COEB_OK) return st;
    int* di = (int*)m->out.d;
    int* dd1 = (int*)(m->out.d + al(NQ * 4));
    int* dd2 = (int*)(m->out.d + 2 * al(NQ * 4));
    st = coeb_knn2_device(m, dq, nq, dt, nt, nnratio, di, dd1, dd2);
    if (st != COEB_OK) return st;
    if ((st = pull_outputs(m, 3 * al(NQ * 4))) != COEB_OK) return st;
    std::memcpy(best_idx, m->out.h, NQ * 4);
    if (d1) std::memcpy(d1, m->out.h + al(NQ * 4), NQ * 4);
    if (d2) std::memcpy(d2, m->out.h + 2 * al(NQ * 4), NQ * 4);
    if (naccepted_out) { int a = 0; for (int i = 0; i < nq; i++) a += best_idx[i] >= 0; *naccepted_out = a; }
    return COEB_OK;
}

// ---- Frame constructor tail from the extractor's device output ------------------------------------------------------------
int coeb_frame_from_extractor(coeb_matcher* m, coeb_extractor* ex, int frame_index, int n, const coeb_camera* cam,
                              const float* dist_coef5, const coeb_depth_image* depth, coeb_keypoint* keys_un_out,
                              float* uright_out, float* depth_out, int* n_out, coeb_frame** out) {
    if (!m || !ex || !cam || !out) return fail(COEB_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (n_out) *n_out = 0;
    int ex_dev = 0;
    void* ex_stream = nullptr;
    int st = coeb_extractor_device_stream(ex, &ex_dev, &ex_stream);
    if (st != COEB_OK) return st;
    if (ex_dev != m->device) return fail(COEB_ERR_INVALID_ARG, "extractor on device %d, matcher on device %d", ex_dev, m->device);
    const coeb_keypoint* d_kps = nullptr;
    const uint8_t* d_desc = nullptr;
    const int* d_count = nullptr;
    int cap = 0;
    if ((st = coeb_extractor_device_outputs(ex, frame_index, &d_kps, &d_desc, &d_count, &cap)) != COEB_OK) return st;
    CUDA_TRY(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    if (n < 0) {   // count unknown to the caller: one 4-byte read behind the extractor's stream
        CUDA_TRY(cudaStreamSynchronize((cudaStream_t)ex_stream));
        CUDA_TRY(cudaMemcpy(&n, d_count, sizeof(int), cudaMemcpyDeviceToHost));
    } else if ((cudaStream_t)ex_stream != s) {
        if (!m->ev_ex) CUDA_TRY(cudaEventCreateWithFlags(&m->ev_ex, cudaEventDisableTiming));
        CUDA_TRY(cudaEventRecord(m->ev_ex, (cudaStream_t)ex_stream));
        CUDA_TRY(cudaStreamWaitEvent(s, m->ev_ex, 0));
    }
    if (n < 0 || n > cap) return fail(COEB_ERR_INVALID_ARG, "n = %d outside the extractor's capacity %d", n, cap);
    int nlevels = 0;
    float scale[COEB_MAX_LEVELS] = {0};
    if ((st = coeb_extractor_tables(ex, &nlevels, scale, nullptr, nullptr, nullptr, nullptr)) != COEB_OK) return st;
    const int kind = depth ? depth->kind : 0;
    if (kind < 0 || kind > 2) return fail(COEB_ERR_INVALID_ARG, "depth kind %d", kind);
    if (kind && (!depth->data || depth->width < 1 || depth->height < 1 || depth->stride_bytes < depth->width * (kind == 1 ? 4 : 2)))
        return fail(COEB_ERR_INVALID_ARG, "bad depth image");

    const size_t nn = std::max(n, 1);
    // device block: x, y, angle, octave, uright, desc | cell_start, cell_items, kp_cell | keys_un (AoS), depth: the last two
    // are contiguous so that the host copy is one transfer
    const size_t need = 5 * al(nn * 4) + al(nn * 32) + al((kGridCells + 1) * 4) + 2 * al(nn * 4) + al(nn * sizeof(coeb_keypoint)) + 2 * al(nn * 4);
    coeb_frame* f = new coeb_frame();
    f->m = m; f->device = m->device; f->n = n; f->nlevels = nlevels;
    f->block = g_frames.take(m->device, need, &f->block_bytes);
    if (!f->block) {
        const size_t want = std::max<size_t>(need + need / 4, 1 << 17);
        if (cudaMalloc(&f->block, want) != cudaSuccess) { delete f; return fail(COEB_ERR_CUDA, "cudaMalloc(%zu) failed", want); }
        f->block_bytes = want;
    }
    char* base = (char*)f->block;
    size_t off = 0;
    auto carve = [&](size_t bytes) { char* p = base + off; off += al(bytes); return p; };
    TailArgs a{};
    a.kps = d_kps; a.desc = (const uint32_t*)d_desc; a.n = n;
    a.fx = cam->fx; a.fy = cam->fy; a.cx = cam->cx; a.cy = cam->cy; a.bf = cam->bf;
    a.undistort = (dist_coef5 && dist_coef5[0] != 0.0f) ? 1 : 0;   // `if (mDistCoef.at<float>(0) == 0.0)` (src/Frame.cc:581)
    for (int i = 0; i < 5; i++) a.dist[i] = dist_coef5 ? dist_coef5[i] : 0.f;
    a.x = (float*)carve(nn * 4); a.y = (float*)carve(nn * 4); a.angle = (float*)carve(nn * 4); a.octave = (int*)carve(nn * 4);
    a.uright = (float*)carve(nn * 4); a.desc_out = (uint32_t*)carve(nn * 32);
    int* d_cell_start = (int*)carve((kGridCells + 1) * 4); int* d_cell_items = (int*)carve(nn * 4); int* d_kp_cell = (int*)carve(nn * 4);
    a.keys_un = (coeb_keypoint*)carve(nn * sizeof(coeb_keypoint));
    float* d_uright_copy = (float*)carve(nn * 4);
    a.uright_dl = d_uright_copy;
    a.depth_out = (float*)carve(nn * 4);
    const size_t dl_bytes = (size_t)((char*)a.depth_out - (char*)a.keys_un) + nn * 4;

    auto bail = [&](int code) { coeb_frame_destroy(f); return code; };
    a.depth_kind = kind;
    if (kind) {
        a.depth_w = depth->width; a.depth_h = depth->height; a.depth_factor = depth->factor;
        const coeb_keypoint* h_kps = nullptr;
        int h_n = 0;
        if (!depth->on_device && n > 0 && !getenv("COEB_NO_DEPTH_GATHER")) coeb_extractor_host_outputs(ex, frame_index, &h_kps, &h_n);
        if (depth->on_device) { a.depth = depth->data; a.depth_stride = depth->stride_bytes; }
        else if (h_kps && h_n == n) {
            // The extractor left its keypoints in host memory (single-frame call): the ~1000 depth values ComputeStereoFromRGBD reads
            // are gathered here and uploaded as 4 KB instead of the whole depth map (0.6 MB: 11 us from pinned memory, ~40 us from a
            // pageable cv::Mat). Same truncation and the same bounds test as the kernel's own gather.
            // The values sit in mapped pinned memory and the kernel reads them across PCIe itself (32 coalesced 128-byte reads): a
            // host-to-device copy node in front of the kernel costs the call more than the transfer.
            if ((st = m->inm.reserve_mapped(al((size_t)n * 4))) != COEB_OK) return bail(st);
            float* g = (float*)m->inm.h;
            const char* img = (const char*)depth->data;
            for (int i = 0; i < n; i++) {
                const int row = (int)h_kps[i].y, col = (int)h_kps[i].x;
                float val = 0.f;
                if ((unsigned)row < (unsigned)depth->height && (unsigned)col < (unsigned)depth->width) {
                    const char* p = img + (size_t)row * depth->stride_bytes;
                    val = kind == 1 ? ((const float*)p)[col] : (float)((const unsigned short*)p)[col];
                }
                g[i] = val;
            }
            a.depth = m->inm.d; a.depth_stride = 0;
            a.depth_kind = kind == 1 ? 3 : 4;
        }
        else {
            const size_t row = (size_t)depth->width * (kind == 1 ? 4 : 2);
            const size_t pitch = (row + 15) & ~(size_t)15;
            if ((st = grow(&m->d_depth, &m->depth_bytes, pitch * depth->height)) != COEB_OK) return bail(st);
            cudaError_t e = (size_t)depth->stride_bytes == row && pitch == row
                                ? cudaMemcpyAsync(m->d_depth, depth->data, row * depth->height, cudaMemcpyHostToDevice, s)
                                : cudaMemcpy2DAsync(m->d_depth, pitch, depth->data, depth->stride_bytes, row, depth->height, cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) return bail(fail(COEB_ERR_CUDA, "depth upload failed: %s", cudaGetErrorString(e)));
            a.depth = m->d_depth; a.depth_stride = (int)pitch;
        }
    }
    FrameDev& d = f->dev;
    d.n = n; d.x = a.x; d.y = a.y; d.angle = a.angle; d.octave = a.octave; d.desc = a.desc_out; d.uright = kind ? a.uright : nullptr;
    d.cell_start = d_cell_start; d.cell_items = d_cell_items;
    d.min_x = cam->min_x; d.min_y = cam->min_y; d.max_x = cam->max_x; d.max_y = cam->max_y;
    d.gw_inv = (float)COEB_GRID_COLS / (cam->max_x - cam->min_x);
    d.gh_inv = (float)COEB_GRID_ROWS / (cam->max_y - cam->min_y);
    d.fx = cam->fx; d.fy = cam->fy; d.cx = cam->cx; d.cy = cam->cy; d.bf = cam->bf; d.b = cam->b;
    for (int i = 0; i < COEB_MAX_LEVELS; i++) d.scale[i] = i < nlevels ? scale[i] : 0.f;
    // the AoS copies the caller asked for (mvKeysUn, mvuRight, mvDepth) are written by the kernel straight into mapped pinned host
    // memory, in the layout of the frame block's own copies: no device-to-host copy behind the kernels
    const bool want = n > 0 && (keys_un_out || uright_out || depth_out);
    const size_t off_ur = (size_t)((char*)d_uright_copy - (char*)a.keys_un), off_dp = (size_t)((char*)a.depth_out - (char*)a.keys_un);
    if (want) {
        if ((st = m->outm.reserve_mapped(dl_bytes)) != COEB_OK) return bail(st);
        a.keys_un = (coeb_keypoint*)m->outm.d;
        a.uright_dl = (float*)(m->outm.d + off_ur);
        a.depth_out = (float*)(m->outm.d + off_dp);
    }
    static const bool two_launches = getenv("COEB_TAIL_TWO_LAUNCHES") != nullptr;   // development switch
    if (n <= kTailGridMax && !two_launches) {
        frame_tail_grid_kernel<<<1, 1024, 0, s>>>(a, d, d_cell_start, d_cell_items);
    } else {
        if (n > 0) frame_tail_kernel<<<(n + 127) / 128, 128, 0, s>>>(a);
        grid_build_kernel<<<1, 1024, 0, s>>>(d, d_cell_start, d_cell_items, d_kp_cell);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bail(fail(COEB_ERR_CUDA, "frame tail launch failed: %s", cudaGetErrorString(e)));
    // blocking like every Frame-building call: the caller may reuse its depth buffer and the extractor
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return bail(fail(COEB_ERR_CUDA, "frame tail failed: %s", cudaGetErrorString(e)));
    if (want) {
        const char* h = m->outm.h;
        if (keys_un_out) std::memcpy(keys_un_out, h, (size_t)n * sizeof(coeb_keypoint));
        if (uright_out) std::memcpy(uright_out, h + off_ur, (size_t)n * 4);
        if (depth_out) std::memcpy(depth_out, h + off_dp, (size_t)n * 4);
    }
    if (n_out) *n_out = n;
    *out = f;
    return COEB_OK;
}

// ---- resident local map + SearchLocalPoints -------------------------------------------------------------------------------
int coeb_local_map_create(coeb_matcher* m, int n, const float* xyz, const float* normal, const float* min_dist, const float* max_dist,
                          const uint8_t* desc, coeb_local_map** out) {
    if (!m || !out || n < 0 || (n > 0 && (!xyz || !normal || !min_dist || !max_dist || !desc))) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(m->device));
    const size_t N = std::max(n, 1);
    const size_t bytes = 2 * al(N * 12) + 2 * al(N * 4) + al(N * 32);
    int st = m->in.reserve(bytes);
    if (st != COEB_OK) return st;
    coeb_local_map* lm = new coeb_local_map();
    lm->m = m; lm->device = m->device; lm->n = n;
    if (cudaMalloc(&lm->block, bytes) != cudaSuccess) { delete lm; return fail(COEB_ERR_CUDA, "cudaMalloc(%zu) failed", bytes); }
    size_t off = 0;
    auto put = [&](const void* src, size_t elem_bytes) {   // stages one array, returns its device address
        char* d = (char*)lm->block + off;
        if (n) std::memcpy(m->in.h + off, src, (size_t)n * elem_bytes);
        off += al(N * elem_bytes);
        return d;
    };
    lm->dev.n = n;
    lm->dev.xyz = (const float*)put(xyz, 12);
    lm->dev.normal = (const float*)put(normal, 12);
    lm->dev.min_dist = (const float*)put(min_dist, 4);
    lm->dev.max_dist = (const float*)put(max_dist, 4);
    lm->dev.desc = (const uint32_t*)put(desc, 32);
    cudaError_t e = cudaMemcpyAsync(lm->block, m->in.h, bytes, cudaMemcpyHostToDevice, m->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
    if (e != cudaSuccess) { cudaFree(lm->block); delete lm; return fail(COEB_ERR_CUDA, "local map upload failed: %s", cudaGetErrorString(e)); }
    *out = lm;
    return COEB_OK;
}

void coeb_local_map_destroy(coeb_local_map* lm) {
    if (!lm) return;
    cudaSetDevice(lm->device);
    cudaFree(lm->block);
    delete lm;
}

#ifdef COEB_KERNEL_TRACE
extern "C++" {
namespace {
void print_mtrace(const char* what) {
    if (!getenv("COEB_KERNEL_TRACE")) return;
    static const char* names[8] = {"frame tail", "grid build", "frustum+collect", "m2 collect", "m2 resolve", "m3/m4 collect", "m3/m4 resolve", ""};
    unsigned long long t[32];
    cudaMemcpyFromSymbol(t, g_mtrace, sizeof(t));
    unsigned long long t0 = ~0ull;
    for (int i = 0; i < 8; i++) if (t[2 * i + 1]) t0 = std::min(t0, t[2 * i]);
    fprintf(stderr, "[coeb match kernels] %s:", what);
    for (int i = 0; i < 8; i++) if (t[2 * i + 1]) fprintf(stderr, " %s %.1f-%.1f |", names[i], (double)(t[2 * i] - t0) * 1e-3, (double)(t[2 * i + 1] - t0) * 1e-3);
    fprintf(stderr, " marks:");
    for (int i = 16; i < 32; i++) if (t[i]) fprintf(stderr, " %.1f", (double)(t[i] - t0) * 1e-3);
    int ms[16];
    cudaMemcpyFromSymbol(ms, g_mstat, sizeof(ms));
    fprintf(stderr, " | active %d, slow per round:", ms[0]);
    for (int i = 1; i < 12 && ms[i] >= 0; i++) fprintf(stderr, " %d", ms[i]);
    if (ms[12] > 0 && t[16] && t[23]) fprintf(stderr, " | rounds: %d SM cycles in %.1f us = %.0f MHz", ms[12], (double)(t[23] - t[16]) * 1e-3, ms[12] / ((double)(t[23] - t[16]) * 1e-3));
    fprintf(stderr, "\n");
    for (int i = 0; i < 16; i++) ms[i] = -1;
    cudaMemcpyToSymbol(g_mstat, ms, sizeof(ms));
    unsigned long long init[32];
    for (int i = 0; i < 32; i++) init[i] = (i < 16 && !(i & 1)) ? ~0ull : 0ull;
    cudaMemcpyToSymbol(g_mtrace, init, sizeof(init));
}
}  // namespace
}  // extern "C++"
#else
extern "C++" { namespace { void print_mtrace(const char*) {} } }
#endif

int coeb_search_local_points(coeb_matcher* m, coeb_frame* F, coeb_local_map* lm, const uint8_t* skip, const uint8_t* has_obs,
                             const float* Tcw, const float* Ow, float viewing_cos_limit, float th, float nnratio, int* kp_match,
                             uint8_t* in_view_out, float* proj_out, int* nmatches_out) {
    if (!m || !F || !lm || !Tcw || !Ow || !kp_match) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (nmatches_out) *nmatches_out = 0;
    const int n = lm->n;
    if (n == 0) return COEB_OK;
    if (!skip || !has_obs) return fail(COEB_ERR_INVALID_ARG, "null map-point flags");
    // an empty frame still gets its visibility flags: every window is empty, nothing matches
    CUDA_TRY(cudaSetDevice(m->device));
    const size_t N = n, K = std::max(F->n, 1);
    int st;
    if ((st = m->in.reserve(2 * al(N) + al(K * 4))) != COEB_OK) return st;
    const size_t out_bytes = al(K * 4) + 256 + al(N) + (proj_out ? al(N * 20) : 0);
    if ((st = m->outm.reserve_mapped(out_bytes)) != COEB_OK) return st;
    const int cap = 32;
    const size_t claim_smem = (size_t)F->n * 4 <= 32 * 1024 ? (size_t)F->n * 4 : 0;
    // scratch: res, claim, list counts, lists | MapPoint fields written by the frustum pass
    const size_t sc_bytes = al(N * 4) + al(K * 12) + lists_bytes(N, cap) + 5 * al(N * 4) + 2 * al(N);
    if (N > m->zero_bytes) {   // all-zero isBad() flags: written once, when the buffer grows
        cudaFree(m->d_zero);
        m->d_zero = nullptr; m->zero_bytes = 0;
        CUDA_TRY(cudaMalloc((void**)&m->d_zero, al(N + N / 2)));
        CUDA_TRY(cudaMemsetAsync(m->d_zero, 0, al(N + N / 2), m->stream));
        m->zero_bytes = al(N + N / 2);
    }
    if ((st = grow(&m->d_scratch, &m->scratch_bytes, sc_bytes)) != COEB_OK) return st;
    Packer p(m->in);
    const uint8_t* d_skip = p.place(skip, N);
    const uint8_t* d_obs = p.place(has_obs, N);
    const int* d_state = p.place(kp_match, (size_t)F->n);
    if ((st = push_inputs(m, p)) != COEB_OK) return st;
    char* sc = (char*)m->d_scratch;
    int* d_res = (int*)sc; sc += al(N * 4);
    int* d_claim = (int*)sc; sc += al(K * 12);
    const CandLists C = carve_lists(sc, N, cap, m->d_meta); sc += lists_bytes(N, cap);
    MapFields mf{};
    mf.proj_x = (float*)sc; sc += al(N * 4);
    mf.proj_y = (float*)sc; sc += al(N * 4);
    mf.proj_xr = (float*)sc; sc += al(N * 4);
    mf.view_cos = (float*)sc; sc += al(N * 4);
    mf.level = (int*)sc; sc += al(N * 4);
    uint8_t* d_zero = m->d_zero;                   // isBad(): bad points arrive as skip
    mf.track_in_view = (uint8_t*)sc; sc += al(N);  // device copy: the collection reads it back
    int* d_kpm = (int*)m->outm.d;
    int* d_info = (int*)(m->outm.d + al(K * 4));
    uint8_t* const flags_host = (uint8_t*)(m->outm.d + al(K * 4) + 256);
    const bool cached = m2_cached_fits(F, n);      // the cached resolve kernel forwards the flags to the host (see there)
    mf.track_in_view_host = cached ? nullptr : flags_host;
    float* d_proj = proj_out ? (float*)(m->outm.d + al(K * 4) + 256 + al(N)) : nullptr;
    MapDev M{};
    M.n = n; M.track_in_view = mf.track_in_view; M.bad = d_zero; M.has_obs = d_obs;
    M.proj_x = mf.proj_x; M.proj_y = mf.proj_y; M.proj_xr = mf.proj_xr; M.view_cos = mf.view_cos; M.level = mf.level; M.desc = lm->dev.desc;
    PoseArgs P{};
    for (int i = 0; i < 12; i++) P.T[i] = Tcw[i];
    for (int i = 0; i < 3; i++) P.Ow[i] = Ow[i];
    P.cos_limit = viewing_cos_limit;
    P.nlevels = F->nlevels;
    frustum_collect_kernel<<<(n + 7) / 8, 256, 0, m->stream>>>(F->dev, lm->dev, P, d_skip, mf, M, th, d_state, C, d_proj);
    CUDA_TRY(cudaGetLastError());
    if ((st = launch_m2_resolve(m, F, M, th, nnratio, d_state, C, d_kpm, d_res, d_claim, d_info, cached ? mf.track_in_view : nullptr, flags_host, N)) != COEB_OK) return st;
    if ((st = sync_outputs(m)) != COEB_OK) return st;
    if (((const int*)(m->outm.h + al(K * 4)))[2]) {   // a candidate list overflowed: exact window-walking variant
        if (claim_smem) m2_resolve_kernel<false, true><<<1, 1024, claim_smem + 16, m->stream>>>(F->dev, M, th, nnratio, d_state, C, d_kpm, d_res, d_claim, d_info, 0);
        else m2_resolve_kernel<false, false><<<1, 1024, 16, m->stream>>>(F->dev, M, th, nnratio, d_state, C, d_kpm, d_res, d_claim, d_info, 0);
        CUDA_TRY(cudaGetLastError());
        if ((st = sync_outputs(m)) != COEB_OK) return st;
    }
    if (F->n) std::memcpy(kp_match, m->outm.h, (size_t)F->n * 4);
    if (in_view_out) std::memcpy(in_view_out, m->outm.h + al(K * 4) + 256, N);
    if (proj_out) std::memcpy(proj_out, m->outm.h + al(K * 4) + 256 + al(N), N * 20);
    if (nmatches_out) *nmatches_out = ((const int*)(m->outm.h + al(K * 4)))[0];
    print_mtrace("SearchLocalPoints");
    return COEB_OK;
}

// ---- SearchByBoW ------------------------------------------------------------------------------------------------------------
namespace {
// A DBoW2::FeatureVector as CSR must have ascending unique node ids, monotone starts and every feature index at most once.
int check_featvec(const char* which, int n_feat, int nn, const int* node, const int* start, const int* items, std::vector<uint8_t>& seen) {
    if (nn < 0 || (nn > 0 && (!node || !start || !items))) return fail(COEB_ERR_INVALID_ARG, "%s: null feature vector", which);
    if (nn == 0) return COEB_OK;
    if (start[0] != 0) return fail(COEB_ERR_INVALID_ARG, "%s: start[0] != 0", which);
    seen.assign((size_t)std::max(n_feat, 1), 0);
    for (int k = 0; k < nn; k++) {
        if (k && node[k] <= node[k - 1]) return fail(COEB_ERR_INVALID_ARG, "%s: node ids not ascending at %d", which, k);
        if (start[k + 1] < start[k]) return fail(COEB_ERR_INVALID_ARG, "%s: starts not monotone at %d", which, k);
        for (int p = start[k]; p < start[k + 1]; p++) {
            const int i = items[p];
            if (i < 0 || i >= n_feat) return fail(COEB_ERR_INVALID_ARG, "%s: feature index %d out of range", which, i);
            if (seen[i]) return fail(COEB_ERR_INVALID_ARG, "%s: feature %d listed under two nodes", which, i);
            seen[i] = 1;
        }
    }
    return COEB_OK;
}
}  // namespace

// Validates the two feature vectors, intersects their node lists (the `while (KFit != KFend && Fit != Fend)` merge with
// lower_bound, :183-262) and stages nodes, item lists and the per-feature flags. Returns COEB_OK with a.n_common == 0 when
// nothing can match.
static int stage_bow(coeb_matcher* m, coeb_frame* f1, coeb_frame* f2, const uint8_t* valid1, const uint8_t* valid2, int nn1,
                     const int* node1, const int* start1, const int* items1, int nn2, const int* node2, const int* start2,
                     const int* items2, BowArgs* out) {
    BowArgs a{};
    *out = a;
    std::vector<uint8_t> seen;
    int st;
    if ((st = check_featvec("feature vector 1", f1->n, nn1, node1, start1, items1, seen)) != COEB_OK) return st;
    if ((st = check_featvec("feature vector 2", f2->n, nn2, node2, start2, items2, seen)) != COEB_OK) return st;
    std::vector<int4> common;
    for (int i = 0, j = 0; i < nn1 && j < nn2;) {
        if (node1[i] == node2[j]) {
            if (start1[i + 1] > start1[i] && start2[j + 1] > start2[j]) common.push_back(make_int4(start1[i], start1[i + 1], start2[j], start2[j + 1]));
            i++; j++;
        } else if (node1[i] < node2[j]) i = (int)(std::lower_bound(node1 + i, node1 + nn1, node2[j]) - node1);
        else j = (int)(std::lower_bound(node2 + j, node2 + nn2, node1[i]) - node2);
    }
    if (common.empty()) return COEB_OK;
    CUDA_TRY(cudaSetDevice(m->device));
    const size_t N1 = f1->n, N2 = f2->n, NC = common.size(), I1 = start1[nn1], I2 = start2[nn2];
    if ((st = m->in.reserve(al(NC * 16) + al(I1 * 4) + al(I2 * 4) + al(N1) + al(N2))) != COEB_OK) return st;
    if ((st = m->out.reserve(al(N1 * 4) + 256)) != COEB_OK) return st;
    if ((st = grow(&m->d_scratch, &m->scratch_bytes, al(N2 * 4))) != COEB_OK) return st;
    Packer p(m->in);
    a.n_common = (int)NC;
    a.nodes = p.place(common.data(), NC);
    a.items1 = p.place(items1, I1);
    a.items2 = p.place(items2, I2);
    a.valid1 = p.place(valid1, N1);
    a.valid2 = valid2 ? p.place(valid2, N2) : nullptr;
    if ((st = push_inputs(m, p)) != COEB_OK) return st;
    *out = a;
    return COEB_OK;
}

int coeb_match_bow(coeb_matcher* m, coeb_frame* f1, coeb_frame* f2, const uint8_t* valid1, const uint8_t* valid2, int nn1,
                   const int* node1, const int* start1, const int* items1, int nn2, const int* node2, const int* start2,
                   const int* items2, float nnratio, int check_ori, int strict_low, int* match12, int* nmatches_out) {
    if (!m || !f1 || !f2 || !match12) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (nmatches_out) *nmatches_out = 0;
    for (int i = 0; i < f1->n; i++) match12[i] = -1;
    if (f1->n == 0 || f2->n == 0) return COEB_OK;
    if (!valid1) return fail(COEB_ERR_INVALID_ARG, "null valid1");
    BowArgs a{};
    int st = stage_bow(m, f1, f2, valid1, valid2, nn1, node1, start1, items1, nn2, node2, start2, items2, &a);
    if (st != COEB_OK || a.n_common == 0) return st;
    a.nnratio = nnratio;
    a.strict_low = strict_low;
    const size_t N1 = f1->n, N2 = f2->n;
    int* d_m12 = (int*)m->out.d;
    int* d_info = (int*)(m->out.d + al(N1 * 4));
    int* d_matched2 = (int*)m->d_scratch;
    CUDA_TRY(cudaMemsetAsync(d_m12, 0xFF, N1 * 4, m->stream));
    CUDA_TRY(cudaMemsetAsync(d_matched2, 0, N2 * 4, m->stream));
    bow_node_kernel<<<(unsigned)(((size_t)a.n_common * 32 + 127) / 128), 128, 0, m->stream>>>(f1->dev, f2->dev, a, d_m12, d_matched2);
    bow_finish_kernel<<<1, 1024, 0, m->stream>>>(f1->dev, f2->dev, check_ori, d_m12, d_info);
    CUDA_TRY(cudaGetLastError());
    if ((st = pull_outputs(m, al(N1 * 4) + 4)) != COEB_OK) return st;
    std::memcpy(match12, m->out.h, N1 * 4);
    if (nmatches_out) *nmatches_out = ((const int*)(m->out.h + al(N1 * 4)))[0];
    return COEB_OK;
}

int coeb_match_triangulation(coeb_matcher* m, coeb_frame* f1, coeb_frame* f2, const uint8_t* free1, const uint8_t* free2, int nn1,
                             const int* node1, const int* start1, const int* items1, int nn2, const int* node2, const int* start2,
                             const int* items2, const float* F12, const float* epipole_xy, int only_stereo, int check_ori,
                             int* match12, int* nmatches_out) {
    if (!m || !f1 || !f2 || !match12 || !F12 || !epipole_xy) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (nmatches_out) *nmatches_out = 0;
    for (int i = 0; i < f1->n; i++) match12[i] = -1;
    if (f1->n == 0 || f2->n == 0) return COEB_OK;
    if (!free1 || !free2) return fail(COEB_ERR_INVALID_ARG, "null free1 / free2");
    BowArgs a{};
    int st = stage_bow(m, f1, f2, free1, free2, nn1, node1, start1, items1, nn2, node2, start2, items2, &a);
    if (st != COEB_OK || a.n_common == 0) return st;
    TriArgs t{};
    for (int i = 0; i < 9; i++) t.F12[i] = F12[i];
    t.ex = epipole_xy[0]; t.ey = epipole_xy[1];
    t.only_stereo = only_stereo;
    const size_t N1 = f1->n;
    int* d_m12 = (int*)m->out.d;
    int* d_info = (int*)(m->out.d + al(N1 * 4));
    CUDA_TRY(cudaMemsetAsync(d_m12, 0xFF, N1 * 4, m->stream));
    triangulation_node_kernel<<<(unsigned)(((size_t)a.n_common * 32 + 127) / 128), 128, 0, m->stream>>>(f1->dev, f2->dev, a, t, d_m12);
    bow_finish_kernel<<<1, 1024, 0, m->stream>>>(f1->dev, f2->dev, check_ori, d_m12, d_info);
    CUDA_TRY(cudaGetLastError());
    if ((st = pull_outputs(m, al(N1 * 4) + 4)) != COEB_OK) return st;
    std::memcpy(match12, m->out.h, N1 * 4);
    if (nmatches_out) *nmatches_out = ((const int*)(m->out.h + al(N1 * 4)))[0];
    return COEB_OK;
}

// ---- relocalisation overload of SearchByProjection ---------------------------------------------------------------------------
int coeb_match_reloc(coeb_matcher* m, coeb_frame* cur, int n, const uint8_t* valid, const float* xyz, const float* min_dist,
                     const float* max_dist, const float* angle, const uint8_t* desc, const float* Tcw, const float* Ow, float th,
                     int orb_dist, int check_ori, int* kp_match, int* nmatches_out) {
    if (!m || !cur || n < 0 || !kp_match || !Tcw || !Ow) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (nmatches_out) *nmatches_out = 0;
    if (n == 0 || cur->n == 0) return COEB_OK;
    if (!valid || !xyz || !min_dist || !max_dist || !angle || !desc) return fail(COEB_ERR_INVALID_ARG, "null keyframe array");
    CUDA_TRY(cudaSetDevice(m->device));
    const size_t N = n, K = cur->n;
    int st;
    if ((st = m->in.reserve(2 * al(N) + al(N * 12) + 3 * al(N * 4) + al(N * 32) + al(K * 4))) != COEB_OK) return st;
    if ((st = m->outm.reserve_mapped(al(K * 4) + 256)) != COEB_OK) return st;
    const int cap = 64;
    const size_t claim_smem = K * 4 <= 32 * 1024 ? K * 4 : 0;   // claim table in shared memory when it fits
    if ((st = grow(&m->d_scratch, &m->scratch_bytes, al(N * 4) + al(K * 12) + lists_bytes(N, cap))) != COEB_OK) return st;
    Packer p(m->in);
    LastDev L{};
    L.n = n;
    L.valid = p.place(valid, N);
    uint8_t* d_ones = p.room<uint8_t>(N);   // every successful query claims its keypoint (:1546-1547): has_obs = 1
    std::memset(p.host_at(d_ones), 1, N);
    L.has_obs = d_ones;
    L.xyz = p.place(xyz, N * 3);
    L.min_dist = p.place(min_dist, N); L.max_dist = p.place(max_dist, N);
    L.angle = p.place(angle, N); L.desc = (const uint32_t*)p.place(desc, N * 32);
    L.octave = nullptr;
    // every non-null entry of CurrentFrame.mvpMapPoints blocks: -3 ("holds a MapPoint without observations") counts like -2
    int* h_state = p.host_at(p.room<int>(K));
    const int* d_state = (const int*)((char*)m->in.d + ((char*)h_state - m->in.h));
    for (size_t k = 0; k < K; k++) h_state[k] = kp_match[k] == -1 ? -1 : -2;
    if ((st = push_inputs(m, p)) != COEB_OK) return st;
    for (int i = 0; i < 12; i++) L.T[i] = Tcw[i];
    for (int i = 0; i < 3; i++) L.Ow[i] = Ow[i];
    L.forward = L.backward = 0;   // levels nPredictedLevel - 1 .. + 1 (:1528)
    L.reloc = 1; L.nlevels = cur->nlevels; L.max_accept = orb_dist;
    int* d_kpm = (int*)m->outm.d;
    int* d_info = (int*)(m->outm.d + al(K * 4));
    int* d_res = (int*)m->d_scratch;
    int* d_claim = (int*)((char*)m->d_scratch + al(N * 4));
    const CandLists C = carve_lists((char*)m->d_scratch + al(N * 4) + al(K * 12), N, cap, m->d_meta);
    m3_collect_kernel<<<(n + 7) / 8, 256, 0, m->stream>>>(cur->dev, L, th, d_state, C);
    CUDA_TRY(cudaGetLastError());
    if (m2_cached_fits(cur, n)) {   // the latency form of the fixed-point iteration (m2_resolve_cached_kernel, kMode 3)
        RotArgs rot{check_ori, L.angle, cur->dev.angle};
        if ((st = launch_cached_resolve<3>(m, cur, n, 0.f, L.max_accept, d_state, C, d_kpm, d_claim, d_info, nullptr, nullptr, 0, rot)) != COEB_OK) return st;
    }
    else if (claim_smem) m3_resolve_kernel<true, true><<<1, 512, claim_smem, m->stream>>>(cur->dev, L, th, check_ori, d_state, C, d_kpm, d_res, d_claim, d_info);
    else m3_resolve_kernel<true, false><<<1, 512, 0, m->stream>>>(cur->dev, L, th, check_ori, d_state, C, d_kpm, d_res, d_claim, d_info);
    CUDA_TRY(cudaGetLastError());
    if ((st = sync_outputs(m)) != COEB_OK) return st;
    if (((const int*)(m->outm.h + al(K * 4)))[2]) {
        if (claim_smem) m3_resolve_kernel<false, true><<<1, 512, claim_smem, m->stream>>>(cur->dev, L, th, check_ori, d_state, C, d_kpm, d_res, d_claim, d_info);
    else m3_resolve_kernel<false, false><<<1, 512, 0, m->stream>>>(cur->dev, L, th, check_ori, d_state, C, d_kpm, d_res, d_claim, d_info);
        CUDA_TRY(cudaGetLastError());
        if ((st = sync_outputs(m)) != COEB_OK) return st;
    }
    // entries that were occupied keep the caller's own encoding; free ones take the result (>= 0 assigned, -1 still free)
    const int* h = (const int*)m->outm.h;
    for (size_t k = 0; k < K; k++)
        if (kp_match[k] == -1) kp_match[k] = h[k];
    if (nmatches_out) *nmatches_out = ((const int*)(m->outm.h + al(K * 4)))[0];
    return COEB_OK;
}

// ---- search half of Fuse ------------------------------------------------------------------------------------------------------
int coeb_fuse_search(coeb_matcher* m, coeb_frame* kf, coeb_local_map* lm, const uint8_t* valid, const float* Tcw, const float* Ow, float th,
                     int chi2_tests, int* best_idx, int* nfused_out) {
    if (!m || !kf || !lm || !Tcw || !Ow || !best_idx) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    if (nfused_out) *nfused_out = 0;
    const int n = lm->n;
    if (n == 0) return COEB_OK;
    if (!valid) return fail(COEB_ERR_INVALID_ARG, "null valid");
    if (kf->n == 0) { for (int i = 0; i < n; i++) best_idx[i] = -1; return COEB_OK; }
    CUDA_TRY(cudaSetDevice(m->device));
    const size_t N = n;
    int st;
    if ((st = m->in.reserve(al(N))) != COEB_OK || (st = m->out.reserve(al(N * 4))) != COEB_OK) return st;
    Packer p(m->in);
    const uint8_t* d_valid = p.place(valid, N);
    if ((st = push_inputs(m, p)) != COEB_OK) return st;
    PoseArgs P{};
    for (int i = 0; i < 12; i++) P.T[i] = Tcw[i];
    for (int i = 0; i < 3; i++) P.Ow[i] = Ow[i];
    P.nlevels = kf->nlevels;
    fuse_search_kernel<<<(n + 7) / 8, 256, 0, m->stream>>>(kf->dev, lm->dev, P, d_valid, th, chi2_tests, (int*)m->out.d);
    CUDA_TRY(cudaGetLastError());
    if ((st = pull_outputs(m, N * 4)) != COEB_OK) return st;
    std::memcpy(best_idx, m->out.h, N * 4);
    if (nfused_out) { int c = 0; for (int i = 0; i < n; i++) c += best_idx[i] >= 0; *nfused_out = c; }
    return COEB_OK;
}

}  // extern "C"
