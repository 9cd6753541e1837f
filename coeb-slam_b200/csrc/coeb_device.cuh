// coeb_device.cuh -- shared host/device declarations of the B200 front end (internal; the public
// boundary is include/coeb_frontend.h).
//
// HBM layout:
//   pyramid arena : for each level l a [B][h_l][pitch_l] uint8 block (pitch_l = w_l rounded up to 64 B),
//                   levels back to back; level 0 may alias the caller's device input instead.
//   blurred arena : same shape, the 7x7 sigma=2 fixed-point Gaussian of every level.
//   candidates    : per (frame, level) a uint32 array (x:12 | y:12 | response:8, minBorder-relative)
//                   sized to the 3x3-NMS upper bound, plus one counter per (frame, level).
//   level keys    : per (frame, level) up to key_cap selected keypoints {x,y (level coords), response,
//                   angle}, list order, after culling; plus one counter per (frame, level).
#pragma once
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "../../include/coeb_types.h"

namespace coeb {

constexpr int kEdge = 19;        // EDGE_THRESHOLD, reference src/ORBextractor.cc:76
constexpr int kHalfPatch = 15;   // HALF_PATCH_SIZE, :75
constexpr int kPatch = 31;       // PATCH_SIZE, :74
constexpr int kMinBorder = 16;   // EDGE_THRESHOLD - 3, :795
constexpr int kCellW = 30;       // W, :789
constexpr int kFastTileInt4 = 3; // int4 entries per tile of the FAST tile table (fast.cu build_fast_tiles)

struct LevelGeom {
    int w, h, pitch;             // level image size and row pitch (bytes)
    int maxBX, maxBY;            // w - 16, h - 16
    int nCols, nRows, wCell, hCell;
    int lastJ, lastI;            // last cell column / row that exists (:816-826); detection pixels beyond it belong to it
    int rcpW, rcpH;              // ceil(2^20 / wCell), ceil(2^20 / hCell): (n * rcp) >> 20 == n / cell for n < 4128
    int cell_base;               // index of this level's first cell in the per-frame cell list
    int n_target;                // mnFeaturesPerLevel[l]
    int n_ini;                   // octree root count, round(W/H)
    int cand_cap, cand_base;     // per-frame candidate capacity / offset (uint32 units)
    int key_cap, key_base;       // per-frame selected-key capacity / offset
    int scaled_patch;            // (int)(31 * scale)
    float hX;                    // (float)(maxX-minX)/nIni
    float scale;                 // mvScaleFactor[l]
    unsigned long long img_base;   // byte offset of this level's [B] block inside the pyramid arena
    unsigned long long img_stride; // bytes between consecutive frames of this level
    unsigned long long tab_base;   // offset (int2 units) of this level's resize tables: dw x-entries then dh y-entries
};

struct Geometry {
    int nlevels;
    int w0, h0;
    int cells_per_frame;
    int cand_per_frame;          // sum of cand_cap
    int keys_per_frame;          // sum of key_cap
    int out_cap;                 // per-frame capacity of the caller's keypoint/descriptor arrays
    int max_nodes;               // octree node-table capacity (per generation)
    int fast_tiles_per_frame;    // entries of the FAST tile table
    int blur_tiles_per_frame;    // entries of the blur tile table
    LevelGeom lv[COEB_MAX_LEVELS];
    int umax[16];
};

// One selected keypoint of a level before the final rescale (16 bytes).
struct LevelKey {
    float x, y;      // level coordinates including the +16 border offset
    float response;
    float angle;
};

// Per-frame dynamic-object decision kept on the device (mirrors coeb_dyn_info).
struct DynState {
    int area_flag;
    int n_dynamic;
    int rect[COEB_MAX_BOXES][4];
    float area;
    int bad_box;     // a box lay outside the image: the frame is reported as COEB_ERR_BAD_BOX
};

// First tile index of every level inside one frame's tile list (kernels that walk all levels in one launch).
struct TileMap {
    int tile_base[COEB_MAX_LEVELS + 1];
    int tiles_x[COEB_MAX_LEVELS];
};

// Small-batch pyramid (pyramid.cu, pyramid_regions_kernel): one CTA owns a rectangle of EVERY level and computes the whole chain
// for it in shared memory. Per (region, level): the rectangle it computes (owned + the halo the next level's rectangle reads, x
// bounds multiples of 4), the rectangle it owns (written to global memory), and where the computed rectangle lives in shared memory.
struct PyrRegionLevel {
    short cx0, cy0, cx1, cy1;
    short ox0, oy0, ox1, oy1;
    int soff, pitch;
};

// Device pointers of one batch launch.
struct BatchView {
    int B;
    const uint8_t* l0;                 // level-0 images
    int l0_pitch;
    unsigned long long l0_stride;      // bytes between level-0 frames
    uint8_t* pyr;                      // pyramid arena (levels >= 1; level 0 too unless aliased)
    uint8_t* blur;                     // blurred arena
    const int2* tabs;                  // resize tables
    const int4* fast_tiles;            // FAST tile table, kFastTileInt4 entries per tile of one frame: {level, tx0, ty0, 0} and the tile's cell-boundary masks
    const int4* blur_tiles;            // blur tile table, same layout
    const uint32_t* ic_mask;           // [4 alignments][16 |v|][9 words]: 0xFF per patch byte inside the circular IC_Angle patch
    const PyrRegionLevel* pyr_regions; // [n_pyr_regions][nlevels] small-batch pyramid regions (null: resize chain only)
    int n_pyr_regions, pyr_regions_smem;
    uint32_t* cand;                    // [B][cand_per_frame]
    int* cand_count;                   // [B][nlevels]
    LevelKey* keys;                    // [B][keys_per_frame]
    int* key_count;                    // [B][nlevels]
    DynState* dyn;                     // [B]
    uint16_t* knode;                   // [B][cand_per_frame] octree scratch: node id per candidate
    uint32_t* lmax;                    // [B][cand_per_frame] FAST cell-local maxima above minTh (same packing as cand)
    int* lmax_count;                   // [B][nlevels]
    int* cell_count;                   // [B][cells_per_frame] local maxima above iniTh per FAST cell
    int* empty_cells;                  // [B * cells_per_frame] cells that need the minTh fallback (frame * cells_per_frame + cell)
    int* empty_count;                  // [1] (one per batch view)
    // dynamic-object inputs
    const float* boxes; const int* nbox; int max_box;
    const float* tm; const int* ntm; int max_tm;
    const int* blur_flag;
    // outputs
    coeb_keypoint* out_kps;            // [B][out_cap]
    uint8_t* out_desc;                 // [B][out_cap][32]
    int* out_count;                    // [B]
    int* status;                       // [B] per-frame coeb_status
    // single-frame host calls: the descriptor stage also writes its results straight into mapped pinned host memory (a device-to-
    // host copy node costs ~10 us of latency, the posted writes of 60 KB cost none); null otherwise. hdr = {count, final status}
    int* mirror_hdr; coeb_keypoint* mirror_kps; uint8_t* mirror_desc;
    unsigned long long* trace;         // development (-DCOEB_KERNEL_TRACE builds, COEB_KERNEL_TRACE=1): [16][2] first start / last end per kernel, %globaltimer ns
};

__host__ __device__ inline const uint8_t* level_ptr(const Geometry& g, const BatchView& v, int level, int frame) {
    if (level == 0) return v.l0 + (unsigned long long)frame * v.l0_stride;
    return v.pyr + g.lv[level].img_base + (unsigned long long)frame * g.lv[level].img_stride;
}
__host__ __device__ inline int level_pitch(const Geometry& g, const BatchView& v, int level) {
    return level == 0 ? v.l0_pitch : g.lv[level].pitch;
}
__host__ __device__ inline uint8_t* blur_ptr(const Geometry& g, const BatchView& v, int level, int frame) {
    return v.blur + g.lv[level].img_base + (unsigned long long)frame * g.lv[level].img_stride;
}

// `*mask.ptr<uchar>(y, x) == 0` of the reference (src/ORBextractor.cc:1397,1440): the mask is the union
// of the zero-filled dynamic rectangles, so it is evaluated from the rectangle list.
__device__ inline bool mask_is_zero(const DynState& d, int x, int y) {
    for (int i = 0; i < d.n_dynamic; i++) {
        if (x >= d.rect[i][0] && x < d.rect[i][2] && y >= d.rect[i][1] && y < d.rect[i][3]) return true;
    }
    return false;
}

// CheckMovingKeyPoints / CheckMovingKeyPoints_finall predicate (src/ORBextractor.cc:1391-1397, 1426-1440).
__device__ inline bool is_moving(const DynState& d, float ptx, float pty, int level, float scale_l, int w0, int h0) {
    const float s = level != 0 ? scale_l : 1.f;
    float sx = __fmul_rn(ptx, s), sy = __fmul_rn(pty, s);
    if (sx >= (float)(w0 - 1)) sx = (float)(w0 - 1);
    if (sy >= (float)(h0 - 1)) sy = (float)(h0 - 1);
    return mask_is_zero(d, (int)sx, (int)sy);
}

#ifdef __CUDACC__
// Development timeline of one call as the GPU ran it (graph replay included), for the latency path where ncu's serialised,
// cold-cache launch list misleads: every kernel notes the start of its first CTA and the end of its last one in v.trace.
// Compiled in only with -DCOEB_KERNEL_TRACE (tools/build_trace.sh); ids: 0 classify, 1/2 FAST level 0 / levels 1.., 3/4 minTh
// fallback, 5/6 octree, 7 pyramid (regions), 8 resize chain, 9 blur, 10 describe, 11 empty-cell list.
#ifdef COEB_KERNEL_TRACE
struct KernelTrace {
    unsigned long long* slot;
    unsigned long long t0;
    __device__ __forceinline__ static unsigned long long now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
    __device__ __forceinline__ KernelTrace(unsigned long long* trace, int id) : slot(trace ? trace + 2 * id : nullptr), t0(0) {
        if (slot && threadIdx.x == 0 && threadIdx.y == 0) t0 = now();
    }
    __device__ __forceinline__ ~KernelTrace() {
        if (slot && threadIdx.x == 0 && threadIdx.y == 0) { atomicMin(slot, t0); atomicMax(slot + 1, now()); }
    }
};
#define COEB_TRACE(v, id) KernelTrace coeb_kernel_trace_((v).trace, (id))
// phase marks inside one chosen CTA: slots 24..31 hold absolute stamps (0: not reached)
#define COEB_TRACE_MARK(v, cond, slot) do { if ((v).trace && (cond) && threadIdx.x == 0) (v).trace[24 + (slot)] = KernelTrace::now(); } while (0)
#else
#define COEB_TRACE(v, id)
#define COEB_TRACE_MARK(v, cond, slot)
#endif

// In-place exclusive prefix sum of data[0..n) (shared or global memory) by the whole CTA; returns the total.
// All threads must call; s_warp is a 33-int shared scratch array. kT = the CTA size when it is known at compile time (the
// chunk size then is a multiply-shift instead of a division by blockDim.x).
template <int kT = 0>
__device__ inline int block_exclusive_scan(int* data, int n, int* s_warp) {
    const int T = kT ? kT : (int)blockDim.x, tid = threadIdx.x;
    const int chunk = (n + T - 1) / T;
    const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += data[i];
    // block scan of per-thread sums
    const int lane = tid & 31, wid = tid >> 5;
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int wv = lane < (T >> 5) ? s_warp[lane] : 0;
        int winc = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        s_warp[lane] = winc - wv;       // exclusive warp offsets
        if (lane == 31) s_warp[32] = winc;  // total
    }
    __syncthreads();
    int run = s_warp[wid] + inc - sum;
    for (int i = lo; i < hi; i++) {
        const int t = data[i];
        data[i] = run;
        run += t;
    }
    const int total = s_warp[32];
    __syncthreads();
    return total;
}

#endif

// One TMA descriptor per pyramid level: the level's [B][h][pitch] block as a 3-D byte tensor whose box is one staged tile
// (box_w x box_h bytes of one frame). Bytes outside the tensor read as 0.
struct TmaMaps { CUtensorMap m[COEB_MAX_LEVELS]; };
// false if the driver entry point is missing, an address / stride is not 16-byte aligned or an encode fails: the kernels
// then stage with plain vector loads.
bool encode_level_maps(const Geometry& g, const BatchView& v, int box_w, int box_h, TmaMaps* out, bool blurred = false);
bool tma_enabled();   // COEB_TMA=0 switches the TMA staging off (development)

#ifdef __CUDACC__
// Single-use mbarrier + one TMA box load, issued by one thread; every thread of the CTA then calls tma_wait() after a
// __syncthreads() that follows the issue (which makes the initialised barrier visible).
__device__ __forceinline__ void tma_issue_box(uint32_t mbar, uint32_t dst, const CUtensorMap* map, int x, int y, int z, int bytes) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(mbar)
                 : "memory");
}
// Re-usable form: initialise once (then __syncthreads / __syncwarp), arm + load per use, wait on the use's parity (0, 1, 0, ...).
__device__ __forceinline__ void tma_bar_init(uint32_t mbar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_box(uint32_t mbar, uint32_t dst, const CUtensorMap* map, int x, int y, int z, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void tma_wait_parity(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TMA_WAITP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TMA_DONEP;\n"
        "bra TMA_WAITP;\n"
        "TMA_DONEP:\n"
        "}\n" ::"r"(mbar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_wait(uint32_t mbar) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TMA_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra TMA_DONE;\n"
        "bra TMA_WAIT;\n"
        "TMA_DONE:\n"
        "}\n" ::"r"(mbar)
        : "memory");
}
#endif

// kernel launchers (each enqueues on `stream`, no synchronisation)
void launch_classify(const Geometry& g, const BatchView& v, cudaStream_t stream);
void launch_pyramid(const Geometry& g, const BatchView& v, cudaStream_t stream);
// Region table of the small-batch pyramid from the host copy of the resize tables; returns the region count (0: not applicable)
// and the dynamic shared memory one CTA needs.
int build_pyramid_regions(const Geometry& g, const int2* h_tabs, PyrRegionLevel* out_or_null, int* smem_bytes);
bool launch_pyramid_regions(const Geometry& g, const BatchView& v, cudaStream_t stream);   // false: no region table, nothing launched
void launch_blur(const Geometry& g, const BatchView& v, cudaStream_t stream);
void launch_fast(const Geometry& g, const BatchView& v, cudaStream_t stream);
int fast_tiles_of_level0(const Geometry& g);
void launch_fast_tiles(const Geometry& g, const BatchView& v, cudaStream_t stream, int first, int count);
void launch_fast_tail(const Geometry& g, const BatchView& v, cudaStream_t stream);
// Small batches: the minTh fallback of the cells of levels [level_lo, level_hi), one warp per cell, each warp testing its own
// cell's counter (no list kernel in between).
void launch_fast_tail_levels(const Geometry& g, const BatchView& v, cudaStream_t stream, int level_lo, int level_hi);
// Host-side FAST tile table of one frame (level, tx0, ty0 per 64x30 tile), uploaded once per geometry.
int build_fast_tiles(const Geometry& g, int4* out_or_null);
int build_blur_tiles(const Geometry& g, int4* out_or_null);
constexpr int kIcMaskWords = 4 * 16 * 9;
void build_ic_masks(const Geometry& g, uint32_t* out);
void launch_select(const Geometry& g, const BatchView& v, cudaStream_t stream, int level_lo = 0, int level_hi = -1);   // levels [lo, hi), default all
void launch_describe(const Geometry& g, const BatchView& v, cudaStream_t stream);

}  // namespace coeb
