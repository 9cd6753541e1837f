// coeb_api.cu -- C ABI of the extractor (include/coeb_frontend.h): handle, geometry, device arenas,
// launch sequence. Host side of ORBextractor (reference src/ORBextractor.cc:418-477, 1088-1367).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/coeb_frontend.h"
#include "coeb_device.cuh"
#include "coeb_host.hpp"

namespace coeb {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

int check_device(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(COEB_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(COEB_ERR_INVALID_ARG, "device %d out of range (%d devices)", device, n);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(COEB_ERR_NO_DEVICE, "device %d is sm_%d%d; the kernels are built for sm_100a only", device, prop.major,
                    prop.minor);
    return COEB_OK;
}

static inline int round_half_even(float v) { return (int)lrintf(v); }  // cvRound(float)

}  // namespace coeb

using namespace coeb;

struct coeb_extractor {
    coeb_orb_params params;
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    // tables of the reference constructor
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> per_level;
    int umax[16];
    // geometry for the current image size
    Geometry geom{};
    bool geom_valid = false;
    std::vector<int2> h_tabs;
    int2* d_tabs = nullptr;
    int4* d_fast_tiles = nullptr;
    PyrRegionLevel* d_pyr_regions = nullptr; int n_pyr_regions = 0, pyr_regions_smem = 0;   // small-batch pyramid (one launch)
    size_t pyr_bytes_per_frame = 0;
    // arenas
    int cap_B = 0;
    int last_B = 0;
    uint8_t *d_pyr = nullptr, *d_blur = nullptr;
    uint32_t* d_cand = nullptr;
    uint16_t* d_knode = nullptr;
    uint32_t* d_lmax = nullptr;
    int *d_lmax_count = nullptr, *d_cell_count = nullptr, *d_empty_cells = nullptr, *d_empty_count = nullptr;
    int *d_cand_count = nullptr, *d_key_count = nullptr;
    LevelKey* d_keys = nullptr;
    DynState* d_dyn = nullptr;
    // staging for the host entry points
    uint8_t* d_in_gray = nullptr; size_t in_gray_bytes = 0; int in_pitch = 0;
    uint8_t* d_in_linear = nullptr; size_t in_linear_bytes = 0;   // landing buffer for tightly packed frames whose width is not the arena pitch
    char *h_dynin = nullptr, *d_dynin = nullptr; size_t dyn_cap = 0;   // boxes | nbox | blur flags | T_M | ntm of a call: one pinned block, one copy
    char* d_out_block = nullptr;                                    // counts | status | keypoints | descriptors: one allocation
    coeb_keypoint* d_out_kps = nullptr; uint8_t* d_out_desc = nullptr; int *d_out_count = nullptr, *d_out_status = nullptr;
    size_t out_cap_elems = 0; int out_cap_B = 0;
    char* h_out1 = nullptr; size_t h_out1_cap = 0;                  // mapped pinned block of single-frame calls: the descriptor stage writes the results into it
    uint8_t* h_in1 = nullptr; size_t h_in1_cap = 0;                 // pinned staging of a single pageable frame (pitch layout)
    BatchView last_view{};
    // copy/compute pipelining of the host entry point
    cudaStream_t pipe_stream[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t pipe_done[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t pipe_ready = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;     // dedicated H2D / D2H streams of the pipelined host entry point
    cudaEvent_t copy_out_done = nullptr;
    std::vector<cudaEvent_t> chunk_in, chunk_done;           // per sub-batch: input resident / kernels finished
    // blur runs beside FAST + octree on a side stream (both only need the pyramid); one lane per launching stream
    struct Lane { cudaStream_t main; cudaStream_t aux; cudaEvent_t fork, join, fork0, cls; cudaStream_t aux2; cudaEvent_t sel0; };
    std::vector<Lane> lanes;
    // CUDA graphs of the kernel sequence for small non-pipelined host calls (single-frame latency path)
    struct GraphEntry { BatchView view; int w, h, cap, chunk; cudaGraphExec_t exec; };
    std::vector<GraphEntry> graphs;
    int graph_warm = 0;
    int last_passes = 1;   // sub-batches the last batch call was split into (launch accounting)
    int last_pass_launches = 0;   // kernels one pass launched (set by enqueue)
    unsigned long long* d_trace = nullptr;   // COEB_KERNEL_TRACE=1 with a -DCOEB_KERNEL_TRACE build: per-kernel start / end stamps
    // optional per-stage CUDA events (benchmark accounting)
    bool profiling = false;
    cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

namespace {

constexpr int kPipeStreams = 3;
constexpr int kPipeChunk = 32;
constexpr int kDevChunk = 128; // sub-batch of the device-resident batch call: batches of 2 * kDevChunk and more are split (0 = never)
constexpr int kGraphMaxBatch = 8;

// ORBextractor::ORBextractor tables (src/ORBextractor.cc:418-477)
void build_tables(coeb_extractor* ex) {
    const int nl = ex->params.nlevels;
    const double scaleFactor = (double)ex->params.scale_factor;  // double member initialised from a float
    ex->scale.assign(nl, 1.f);
    ex->sigma2.assign(nl, 1.f);
    for (int i = 1; i < nl; i++) {
        ex->scale[i] = (float)(ex->scale[i - 1] * scaleFactor);
        ex->sigma2[i] = ex->scale[i] * ex->scale[i];
    }
    ex->inv_scale.resize(nl);
    ex->inv_sigma2.resize(nl);
    for (int i = 0; i < nl; i++) {
        ex->inv_scale[i] = 1.0f / ex->scale[i];
        ex->inv_sigma2[i] = 1.0f / ex->sigma2[i];
    }
    ex->per_level.assign(nl, 0);
    const float factor = (float)(1.0f / scaleFactor);
    float desired = ex->params.nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) {
        ex->per_level[l] = round_half_even(desired);
        sum += ex->per_level[l];
        desired *= factor;
    }
    ex->per_level[nl - 1] = std::max(ex->params.nfeatures - sum, 0);
    // circular patch row ends (:461-476)
    const int hp = kHalfPatch;
    int vmax = (int)std::floor(hp * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(hp * std::sqrt(2.f) / 2);
    for (int v = 0; v <= hp; v++) ex->umax[v] = 0;
    for (int v = 0; v <= vmax; v++) ex->umax[v] = (int)lrint(std::sqrt((double)hp * hp - (double)v * v));
    for (int v = hp, v0 = 0; v >= vmin; --v) {
        while (ex->umax[v0] == ex->umax[v0 + 1]) ++v0;
        ex->umax[v] = v0;
        ++v0;
    }
}

// cv::resize INTER_LINEAR coefficient tables (imgproc/resize.cpp), one (offset, a0 | a1<<16) per
// destination column and row.
void build_resize_tables(int sw, int sh, int dw, int dh, std::vector<int2>& out) {
    const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)std::floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        const int a0 = (int)lrintf((1.f - fx) * 2048.f), a1 = (int)lrintf(fx * 2048.f);
        out.push_back(make_int2(sx, (a0 & 0xFFFF) | (a1 << 16)));
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)std::floor(fy);
        fy -= sy;
        const int b0 = (int)lrintf((1.f - fy) * 2048.f), b1 = (int)lrintf(fy * 2048.f);
        const int sy0 = std::min(std::max(sy, 0), sh - 1), sy1 = std::min(std::max(sy + 1, 0), sh - 1);   // rows clamped here, once
        out.push_back(make_int2(sy0 | (sy1 << 16), (b0 & 0xFFFF) | (b1 << 16)));
    }
}

int build_geometry(coeb_extractor* ex, int w, int h) {
    Geometry& g = ex->geom;
    if (ex->geom_valid && g.w0 == w && g.h0 == h) return COEB_OK;
    const int nl = ex->params.nlevels;
    std::memset(&g, 0, sizeof(g));
    g.nlevels = nl;
    g.w0 = w;
    g.h0 = h;
    for (int i = 0; i < 16; i++) g.umax[i] = ex->umax[i];
    ex->h_tabs.clear();
    unsigned long long img_off = 0;
    int cells = 0, cand = 0, keys = 0, max_nodes = 8;
    for (int l = 0; l < nl; l++) {
        LevelGeom& L = g.lv[l];
        // level size (src/ORBextractor.cc:1348-1349): both from the level-0 size
        L.w = round_half_even((float)w * ex->inv_scale[l]);
        L.h = round_half_even((float)h * ex->inv_scale[l]);
        if (L.w > 4095 + 32 || L.h > 4095 + 32) return fail(COEB_ERR_UNSUPPORTED, "level %d is %dx%d; at most 4127 px per side", l, L.w, L.h);
        L.pitch = (L.w + 63) & ~63;
        L.maxBX = L.w - kEdge + 3;
        L.maxBY = L.h - kEdge + 3;
        const float width = (float)(L.maxBX - kMinBorder), height = (float)(L.maxBY - kMinBorder);
        L.nCols = (int)(width / (float)kCellW);
        L.nRows = (int)(height / (float)kCellW);
        if (L.nCols < 1 || L.nRows < 1)
            return fail(COEB_ERR_UNSUPPORTED, "level %d (%dx%d) is smaller than one 30-px FAST cell", l, L.w, L.h);
        L.wCell = (int)std::ceil(width / L.nCols);
        L.hCell = (int)std::ceil(height / L.nRows);
        if (L.wCell + 6 > 72 || L.hCell + 6 > 72) return fail(COEB_ERR_UNSUPPORTED, "cell larger than the staged ROI");
        if (L.nCols > 4095 || L.nRows > 4095) return fail(COEB_ERR_UNSUPPORTED, "too many cells");
        L.lastJ = std::max(std::min(L.nCols - 1, (L.maxBX - 6 - kMinBorder - 1) / L.wCell), 0);
        L.lastI = std::max(std::min(L.nRows - 1, (L.maxBY - 3 - kMinBorder - 1) / L.hCell), 0);
        L.rcpW = ((1 << 20) + L.wCell - 1) / L.wCell;
        L.rcpH = ((1 << 20) + L.hCell - 1) / L.hCell;
        for (int n = 0; n < 4128; n++)
            if ((int)(((unsigned)n * (unsigned)L.rcpW) >> 20) != n / L.wCell || (int)(((unsigned)n * (unsigned)L.rcpH) >> 20) != n / L.hCell)
                return fail(COEB_ERR_UNSUPPORTED, "internal: reciprocal cell division is inexact at level %d", l);
        L.cell_base = cells;
        cells += L.nCols * L.nRows;
        L.n_target = ex->per_level[l];
        L.n_ini = (int)std::round((float)(L.maxBX - kMinBorder) / (float)(L.maxBY - kMinBorder));  // :550
        if (L.n_ini < 1) return fail(COEB_ERR_UNSUPPORTED, "image taller than wide by more than 2:1 (no octree root)");
        L.hX = (float)(L.maxBX - kMinBorder) / L.n_ini;  // :552
        // candidate capacity = 3x3-NMS upper bound summed over the cells of this level
        int cap = 0;
        for (int i = 0; i < L.nRows; i++) {
            const int iniY = kMinBorder + i * L.hCell;
            if (iniY >= L.maxBY - 3) continue;
            const int rh = std::min(iniY + L.hCell + 6, L.maxBY) - iniY;
            for (int j = 0; j < L.nCols; j++) {
                const int iniX = kMinBorder + j * L.wCell;
                if (iniX >= L.maxBX - 6) continue;
                const int rw = std::min(iniX + L.wCell + 6, L.maxBX) - iniX;
                if (rw < 7 || rh < 7) continue;
                cap += ((rw - 6 + 1) / 2) * ((rh - 6 + 1) / 2);
            }
        }
        L.cand_cap = (cap + 31) & ~31;
        L.cand_base = cand;
        cand += L.cand_cap;
        L.key_cap = std::max(L.n_target + 3, 4 * L.n_ini) + 1;
        L.key_base = keys;
        keys += L.key_cap;
        max_nodes = std::max(max_nodes, L.key_cap + 3);
        L.scale = ex->scale[l];
        L.scaled_patch = (int)(kPatch * ex->scale[l]);  // :877
        L.img_base = img_off;   // frame-major arena: every frame holds its whole pyramid, levels back to back
        L.tab_base = ex->h_tabs.size();
        if (l > 0) build_resize_tables(g.lv[l - 1].w, g.lv[l - 1].h, L.w, L.h, ex->h_tabs);
        img_off += ((unsigned long long)L.pitch * L.h + 255) & ~255ull;
    }
    for (int l = 0; l < nl; l++) g.lv[l].img_stride = img_off;   // bytes between consecutive frames
    g.cells_per_frame = cells;
    g.cand_per_frame = cand;
    g.keys_per_frame = keys;
    g.max_nodes = max_nodes;
    ex->pyr_bytes_per_frame = img_off;
    if (select_smem_bytes(max_nodes) > 200 * 1024)
        return fail(COEB_ERR_UNSUPPORTED, "nfeatures too large for the shared-memory octree (%d nodes)", max_nodes);
    if (ex->d_tabs) cudaFree(ex->d_tabs);
    ex->d_tabs = nullptr;
    if (ex->d_fast_tiles) cudaFree(ex->d_fast_tiles);
    ex->d_fast_tiles = nullptr;
    {
        g.fast_tiles_per_frame = build_fast_tiles(g, nullptr);
        g.blur_tiles_per_frame = build_blur_tiles(g, nullptr);
        // FAST tiles, then blur tiles, then the IC_Angle patch masks
        const size_t fast_int4 = (size_t)kFastTileInt4 * g.fast_tiles_per_frame;
        std::vector<int4> tiles(fast_int4 + g.blur_tiles_per_frame + kIcMaskWords / 4);
        build_fast_tiles(g, tiles.data());
        build_blur_tiles(g, tiles.data() + fast_int4);
        build_ic_masks(g, reinterpret_cast<uint32_t*>(tiles.data() + fast_int4 + g.blur_tiles_per_frame));
        CUDA_TRY(cudaMalloc(&ex->d_fast_tiles, tiles.size() * sizeof(int4)));
        CUDA_TRY(cudaMemcpy(ex->d_fast_tiles, tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice));
    }
    if (!ex->h_tabs.empty()) {
        CUDA_TRY(cudaMalloc(&ex->d_tabs, ex->h_tabs.size() * sizeof(int2)));
        CUDA_TRY(cudaMemcpy(ex->d_tabs, ex->h_tabs.data(), ex->h_tabs.size() * sizeof(int2), cudaMemcpyHostToDevice));
    }
    if (ex->d_pyr_regions) cudaFree(ex->d_pyr_regions);
    ex->d_pyr_regions = nullptr;
    ex->n_pyr_regions = build_pyramid_regions(g, ex->h_tabs.data(), nullptr, &ex->pyr_regions_smem);
    if (ex->n_pyr_regions > 0) {
        std::vector<PyrRegionLevel> regions((size_t)ex->n_pyr_regions * nl);
        build_pyramid_regions(g, ex->h_tabs.data(), regions.data(), &ex->pyr_regions_smem);
        CUDA_TRY(cudaMalloc(&ex->d_pyr_regions, regions.size() * sizeof(PyrRegionLevel)));
        CUDA_TRY(cudaMemcpy(ex->d_pyr_regions, regions.data(), regions.size() * sizeof(PyrRegionLevel), cudaMemcpyHostToDevice));
    }
    ex->geom_valid = true;
    ex->cap_B = 0;  // arenas must be re-laid out for the new geometry
    for (auto& e : ex->graphs) cudaGraphExecDestroy(e.exec);
    ex->graphs.clear();
    return COEB_OK;
}

// A captured graph names device pointers (arenas, staging, output block): every (re)allocation of one of them drops the cache,
// so that a stale graph can never replay writes into freed memory.
static void drop_graphs(coeb_extractor* ex) {
    for (auto& e : ex->graphs) cudaGraphExecDestroy(e.exec);
    ex->graphs.clear();
}

void free_arenas(coeb_extractor* ex) {
    drop_graphs(ex);
    cudaFree(ex->d_lmax); cudaFree(ex->d_lmax_count); cudaFree(ex->d_cell_count); cudaFree(ex->d_empty_cells); cudaFree(ex->d_empty_count);
    ex->d_lmax = nullptr; ex->d_lmax_count = ex->d_cell_count = ex->d_empty_cells = ex->d_empty_count = nullptr;
    cudaFree(ex->d_pyr); cudaFree(ex->d_blur); cudaFree(ex->d_cand); cudaFree(ex->d_knode); cudaFree(ex->d_cand_count);
    cudaFree(ex->d_key_count); cudaFree(ex->d_keys); cudaFree(ex->d_dyn);
    ex->d_pyr = ex->d_blur = nullptr; ex->d_cand = nullptr; ex->d_knode = nullptr; ex->d_cand_count = ex->d_key_count = nullptr;
    ex->d_keys = nullptr; ex->d_dyn = nullptr;
    ex->cap_B = 0;
}

// Arena layout: frame-major ([B] frames, each a whole pyramid), so a sub-batch is a pointer offset.
int ensure_arenas(coeb_extractor* ex, int B) {
    Geometry& g = ex->geom;
    if (B <= ex->cap_B) return COEB_OK;
    free_arenas(ex);
    const unsigned long long off = (unsigned long long)ex->pyr_bytes_per_frame * B;
    CUDA_TRY(cudaMalloc(&ex->d_pyr, off));
    CUDA_TRY(cudaMalloc(&ex->d_blur, off));
    CUDA_TRY(cudaMalloc(&ex->d_cand, (size_t)B * g.cand_per_frame * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&ex->d_knode, (size_t)B * g.cand_per_frame * sizeof(uint16_t)));
    CUDA_TRY(cudaMalloc(&ex->d_lmax, (size_t)B * g.cand_per_frame * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&ex->d_lmax_count, (size_t)B * g.nlevels * sizeof(int)));
    CUDA_TRY(cudaMalloc(&ex->d_cell_count, (size_t)B * g.cells_per_frame * sizeof(int)));
    CUDA_TRY(cudaMalloc(&ex->d_empty_cells, (size_t)B * g.cells_per_frame * sizeof(int)));
    CUDA_TRY(cudaMalloc(&ex->d_empty_count, (size_t)B * sizeof(int)));   // one counter per possible sub-batch start
    CUDA_TRY(cudaMalloc(&ex->d_cand_count, (size_t)B * g.nlevels * sizeof(int)));
    CUDA_TRY(cudaMalloc(&ex->d_key_count, (size_t)B * g.nlevels * sizeof(int)));
    CUDA_TRY(cudaMalloc(&ex->d_keys, (size_t)B * g.keys_per_frame * sizeof(LevelKey)));
    CUDA_TRY(cudaMalloc(&ex->d_dyn, (size_t)B * sizeof(DynState)));
    ex->cap_B = B;
    return COEB_OK;
}

template <typename T>
int ensure_buf(T** p, size_t* cap, size_t n) {
    if (n <= *cap && *p) return COEB_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    CUDA_TRY(cudaMalloc(p, std::max<size_t>(n, 1) * sizeof(T)));
    *cap = n;
    return COEB_OK;
}

coeb_extractor::Lane* lane_for(coeb_extractor* ex, cudaStream_t s) {
    for (auto& l : ex->lanes)
        if (l.main == s) return &l;
    coeb_extractor::Lane l{s, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (cudaStreamCreateWithFlags(&l.aux, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithFlags(&l.aux2, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&l.sel0, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&l.fork0, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&l.cls, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    ex->lanes.push_back(l);
    return &ex->lanes.back();
}

int enqueue(coeb_extractor* ex, const BatchView& v, cudaStream_t s, bool prof) {
    const Geometry& g = ex->geom;
    coeb_extractor::Lane* lane = prof ? nullptr : lane_for(ex, s);
    if (lane && v.B <= 4 && !getenv("COEB_NO_L0_OVERLAP")) {
        // Small batches (the tracking thread's single frame) are latency, not work: every kernel is a few microseconds whatever its
        // size, so the call is as long as its longest chain of dependent launches. Level 0 needs no resize and owns the longest
        // octree, so it runs end to end on a side stream beside everything else:
        //   aux : classify, counters | FAST level 0 -> minTh fallback level 0 -> octree level 0
        //   main: pyramid (one launch: pyramid_regions_kernel) -> FAST levels 1.. -> fallback -> octree levels 1..
        //   aux2: blur (once the pyramid is there)
        //   join -> describe
        static const bool no_regions = getenv("COEB_NO_PYR_REGIONS") != nullptr;
        const int n0 = fast_tiles_of_level0(g);
        CUDA_TRY(cudaEventRecord(lane->fork0, s));
        CUDA_TRY(cudaStreamWaitEvent(lane->aux, lane->fork0, 0));
        launch_classify(g, v, lane->aux);
        CUDA_TRY(cudaEventRecord(lane->cls, lane->aux));     // classification and zeroed counters are in place
        launch_fast_tiles(g, v, lane->aux, 0, n0);
        launch_fast_tail_levels(g, v, lane->aux, 0, 1);
        launch_select(g, v, lane->aux, 0, 1);
        CUDA_TRY(cudaEventRecord(lane->sel0, lane->aux));
        ex->last_pass_launches = 1 + 1 + 1 + 2 + 2 + 2 + 1;
        if (no_regions || !launch_pyramid_regions(g, v, s)) { launch_pyramid(g, v, s); ex->last_pass_launches += g.nlevels - 2; }
        CUDA_TRY(cudaStreamWaitEvent(s, lane->cls, 0));
        launch_fast_tiles(g, v, s, n0, g.fast_tiles_per_frame - n0);
        CUDA_TRY(cudaEventRecord(lane->fork, s));
        CUDA_TRY(cudaStreamWaitEvent(lane->aux2, lane->fork, 0));
        launch_blur(g, v, lane->aux2);   // after the FAST tiles: measured, a blur that runs beside them more than doubles their time
        CUDA_TRY(cudaEventRecord(lane->join, lane->aux2));
        launch_fast_tail_levels(g, v, s, 1, g.nlevels);
        launch_select(g, v, s, 1, g.nlevels);
        CUDA_TRY(cudaStreamWaitEvent(s, lane->sel0, 0));
        CUDA_TRY(cudaStreamWaitEvent(s, lane->join, 0));
        launch_describe(g, v, s);
        CUDA_TRY(cudaGetLastError());
        return COEB_OK;
    }
    ex->last_pass_launches = 1 + (g.nlevels - 1) + 1 + 3 + 1 + 1;
    if (lane) {
        // side stream: classify beside the pyramid, then blur beside FAST + octree | main: pyramid, FAST, octree | join | describe
        CUDA_TRY(cudaEventRecord(lane->fork0, s));
        CUDA_TRY(cudaStreamWaitEvent(lane->aux, lane->fork0, 0));
        launch_classify(g, v, lane->aux);
        CUDA_TRY(cudaEventRecord(lane->cls, lane->aux));
        launch_pyramid(g, v, s);
        CUDA_TRY(cudaEventRecord(lane->fork, s));
        CUDA_TRY(cudaStreamWaitEvent(lane->aux, lane->fork, 0));
        launch_blur(g, v, lane->aux);
        CUDA_TRY(cudaEventRecord(lane->join, lane->aux));
        CUDA_TRY(cudaStreamWaitEvent(s, lane->cls, 0));   // FAST reads the per-frame threshold decision
        launch_fast(g, v, s);
        launch_select(g, v, s);
        CUDA_TRY(cudaStreamWaitEvent(s, lane->join, 0));
        launch_describe(g, v, s);
        CUDA_TRY(cudaGetLastError());
        return COEB_OK;
    }
    if (prof) cudaEventRecord(ex->ev[0], s);
    launch_classify(g, v, s);
    if (prof) cudaEventRecord(ex->ev[1], s);
    launch_pyramid(g, v, s);
    if (prof) cudaEventRecord(ex->ev[2], s);
    launch_blur(g, v, s);
    if (prof) cudaEventRecord(ex->ev[3], s);
    launch_fast(g, v, s);
    if (prof) cudaEventRecord(ex->ev[4], s);
    launch_select(g, v, s);
    if (prof) cudaEventRecord(ex->ev[5], s);
    launch_describe(g, v, s);
    if (prof) cudaEventRecord(ex->ev[6], s);
    CUDA_TRY(cudaGetLastError());
    return COEB_OK;
}

// View of frames [f0, f0 + n) of a full-batch view: every per-frame array is [B][...], so this is pointer arithmetic.
BatchView sub_view(const Geometry& g, const BatchView& v, int f0, int n, size_t pyr_frame_bytes) {
    BatchView s = v;
    s.B = n;
    s.l0 = v.l0 + (size_t)f0 * v.l0_stride;
    s.pyr = v.pyr + (size_t)f0 * pyr_frame_bytes;
    s.blur = v.blur + (size_t)f0 * pyr_frame_bytes;
    s.cand = v.cand + (size_t)f0 * g.cand_per_frame;
    s.knode = v.knode + (size_t)f0 * g.cand_per_frame;
    s.lmax = v.lmax + (size_t)f0 * g.cand_per_frame;
    s.cand_count = v.cand_count + (size_t)f0 * g.nlevels;
    s.lmax_count = v.lmax_count + (size_t)f0 * g.nlevels;
    s.key_count = v.key_count + (size_t)f0 * g.nlevels;
    s.cell_count = v.cell_count + (size_t)f0 * g.cells_per_frame;
    s.empty_cells = v.empty_cells + (size_t)f0 * g.cells_per_frame;
    s.empty_count = v.empty_count + f0;
    s.keys = v.keys + (size_t)f0 * g.keys_per_frame;
    s.dyn = v.dyn + f0;
    if (v.boxes) s.boxes = v.boxes + (size_t)f0 * v.max_box * 4;
    if (v.nbox) s.nbox = v.nbox + f0;
    if (v.tm) s.tm = v.tm + (size_t)f0 * v.max_tm * 2;
    if (v.ntm) s.ntm = v.ntm + f0;
    if (v.blur_flag) s.blur_flag = v.blur_flag + (size_t)f0 * v.max_box;
    s.out_kps = v.out_kps + (size_t)f0 * g.out_cap;
    s.out_desc = v.out_desc + (size_t)f0 * g.out_cap * 32;
    s.out_count = v.out_count + f0;
    s.status = v.status + f0;
    return s;
}

int ensure_pipe_streams(coeb_extractor* ex) {
    if (ex->pipe_stream[0]) return COEB_OK;
    for (int i = 0; i < kPipeStreams; i++) {
        CUDA_TRY(cudaStreamCreateWithFlags(&ex->pipe_stream[i], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&ex->pipe_done[i], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&ex->pipe_ready, cudaEventDisableTiming));
    CUDA_TRY(cudaStreamCreateWithFlags(&ex->copy_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&ex->copy_out, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&ex->copy_out_done, cudaEventDisableTiming));
    return COEB_OK;
}

int ensure_chunk_events(coeb_extractor* ex, int n) {
    while ((int)ex->chunk_in.size() < n) {
        cudaEvent_t a = nullptr, b = nullptr;
        CUDA_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        ex->chunk_in.push_back(a);
        ex->chunk_done.push_back(b);
    }
    return COEB_OK;
}

int validate_common(coeb_extractor* ex, int B, const uint8_t* gray, int w, int h, int stride, int cap) {
    if (!ex) return fail(COEB_ERR_INVALID_ARG, "null extractor");
    if (B < 1) return fail(COEB_ERR_INVALID_ARG, "batch size %d", B);
    if (!gray || w <= 0 || h <= 0 || stride < w) return fail(COEB_ERR_INVALID_ARG, "bad image arguments (w=%d h=%d stride=%d)", w, h, stride);
    if (cap < 1) return fail(COEB_ERR_INVALID_ARG, "output capacity %d", cap);
    return COEB_OK;
}

}  // namespace

extern "C" {

const char* coeb_last_error(void) { return g_last_error.c_str(); }
const char* coeb_version(void) { return "coeb-frontend-b200 0.1 (sm_100a)"; }

int coeb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; i++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major >= 10) ok++;
    }
    return ok;
}

int coeb_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(COEB_ERR_INVALID_ARG, "null pointer");
    CUDA_TRY(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return COEB_OK;
}
int coeb_host_free(void* ptr) {
    CUDA_TRY(cudaFreeHost(ptr));
    return COEB_OK;
}

int coeb_extractor_create(const coeb_orb_params* params, int device, coeb_extractor** out) {
    if (!params || !out) return fail(COEB_ERR_INVALID_ARG, "null argument");
    if (params->nlevels < 1 || params->nlevels > COEB_MAX_LEVELS) return fail(COEB_ERR_INVALID_ARG, "nlevels %d not in [1,%d]", params->nlevels, COEB_MAX_LEVELS);
    if (params->nfeatures < 1 || !(params->scale_factor > 1.0f)) return fail(COEB_ERR_INVALID_ARG, "nfeatures=%d scale_factor=%f", params->nfeatures, params->scale_factor);
    int st = check_device(device);
    if (st != COEB_OK) return st;
    CUDA_TRY(cudaSetDevice(device));
    coeb_extractor* ex = new coeb_extractor();
    ex->params = *params;
    ex->device = device;
    build_tables(ex);
    if (getenv("COEB_KERNEL_TRACE")) cudaMalloc(&ex->d_trace, 32 * sizeof(unsigned long long));
    if (cudaStreamCreateWithFlags(&ex->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ex;
        return fail(COEB_ERR_CUDA, "cudaStreamCreate failed");
    }
    ex->stream = ex->own_stream;
    *out = ex;
    return COEB_OK;
}

void coeb_extractor_destroy(coeb_extractor* ex) {
    if (!ex) return;
    cudaSetDevice(ex->device);
    cudaStreamSynchronize(ex->stream);
    free_arenas(ex);
    cudaFree(ex->d_tabs);
    cudaFree(ex->d_fast_tiles);
    cudaFree(ex->d_pyr_regions);
    cudaFree(ex->d_trace);
    cudaFree(ex->d_in_gray); cudaFree(ex->d_in_linear); cudaFree(ex->d_dynin); cudaFree(ex->d_out_block);
    if (ex->h_dynin) cudaFreeHost(ex->h_dynin);
    if (ex->h_out1) cudaFreeHost(ex->h_out1);
    if (ex->h_in1) cudaFreeHost(ex->h_in1);
    for (auto& e : ex->graphs) cudaGraphExecDestroy(e.exec);
    for (auto& l : ex->lanes) { cudaStreamDestroy(l.aux); cudaStreamDestroy(l.aux2); cudaEventDestroy(l.sel0); cudaEventDestroy(l.fork); cudaEventDestroy(l.join); cudaEventDestroy(l.fork0); cudaEventDestroy(l.cls); }
    for (int i = 0; i < 7; i++) if (ex->ev[i]) cudaEventDestroy(ex->ev[i]);
    for (int i = 0; i < 3; i++) {
        if (ex->pipe_stream[i]) cudaStreamDestroy(ex->pipe_stream[i]);
        if (ex->pipe_done[i]) cudaEventDestroy(ex->pipe_done[i]);
    }
    if (ex->pipe_ready) cudaEventDestroy(ex->pipe_ready);
    if (ex->copy_in) cudaStreamDestroy(ex->copy_in);
    if (ex->copy_out) cudaStreamDestroy(ex->copy_out);
    if (ex->copy_out_done) cudaEventDestroy(ex->copy_out_done);
    for (auto e : ex->chunk_in) cudaEventDestroy(e);
    for (auto e : ex->chunk_done) cudaEventDestroy(e);
    cudaStreamDestroy(ex->own_stream);
    delete ex;
}

int coeb_extractor_set_stream(coeb_extractor* ex, void* cuda_stream) {
    if (!ex) return fail(COEB_ERR_INVALID_ARG, "null extractor");
    ex->stream = cuda_stream ? (cudaStream_t)cuda_stream : ex->own_stream;
    return COEB_OK;
}

int coeb_extractor_reserve(coeb_extractor* ex, int width, int height, int max_batch) {
    if (!ex || max_batch < 1) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(ex->device));
    int st = build_geometry(ex, width, height);
    if (st != COEB_OK) return st;
    return ensure_arenas(ex, max_batch);
}

int coeb_extractor_tables(const coeb_extractor* ex, int* nlevels, float* scale, float* inv_scale, float* sigma2,
                          float* inv_sigma2, int* features_per_level) {
    if (!ex) return fail(COEB_ERR_INVALID_ARG, "null extractor");
    const int nl = ex->params.nlevels;
    if (nlevels) *nlevels = nl;
    for (int i = 0; i < nl; i++) {
        if (scale) scale[i] = ex->scale[i];
        if (inv_scale) inv_scale[i] = ex->inv_scale[i];
        if (sigma2) sigma2[i] = ex->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = ex->inv_sigma2[i];
        if (features_per_level) features_per_level[i] = ex->per_level[i];
    }
    return COEB_OK;
}

// Internal: device ordinal and launch stream of a handle, for the translation units that hold other kernels.
int coeb_extractor_device_stream(coeb_extractor* ex, int* device, void** stream) {
    if (!ex || !device || !stream) return fail(COEB_ERR_INVALID_ARG, "null argument");
    *device = ex->device;
    *stream = (void*)ex->stream;
    return COEB_OK;
}

/* Device-resident results of the last coeb_extract* call (the handle's own output arrays for the host entry points, the
 * caller's for coeb_extract_batch_device): keypoints [cap], descriptors [cap][32] and the count of frame `frame`. */
int coeb_extractor_device_outputs(coeb_extractor* ex, int frame, const coeb_keypoint** d_kps, const uint8_t** d_desc,
                                  const int** d_count, int* cap) {
    if (!ex || !ex->geom_valid || frame < 0 || frame >= ex->last_B) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    const BatchView& v = ex->last_view;
    const int c = ex->geom.out_cap;
    if (d_kps) *d_kps = v.out_kps + (size_t)frame * c;
    if (d_desc) *d_desc = v.out_desc + (size_t)frame * c * 32;
    if (d_count) *d_count = v.out_count + frame;
    if (cap) *cap = c;
    return COEB_OK;
}

int coeb_extractor_host_outputs(coeb_extractor* ex, int frame, const coeb_keypoint** h_kps, int* count) {
    if (!ex || !h_kps) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    *h_kps = nullptr;
    if (count) *count = 0;
    if (!ex->geom_valid || frame != 0 || ex->last_B != 1 || !ex->last_view.mirror_hdr || !ex->h_out1) return COEB_OK;
    const int* hdr = (const int*)ex->h_out1;
    if (hdr[1] != COEB_OK) return COEB_OK;
    *h_kps = ex->last_view.mirror_kps;
    if (count) *count = hdr[0];
    return COEB_OK;
}

int coeb_extractor_launches_per_call(const coeb_extractor* ex) {
    if (!ex) return 0;
    // batches: classify + (nlevels-1) resizes + blur + FAST + empty-cell list + FAST fallback + select + describe;
    // small batches: classify + pyramid (1 launch, or the chain) + blur + 2 FAST + 2 fallback + 2 select + describe
    const int per_pass = ex->last_pass_launches ? ex->last_pass_launches : 1 + (ex->params.nlevels - 1) + 1 + 3 + 1 + 1;
    return per_pass * ex->last_passes;
}

int coeb_extractor_set_profiling(coeb_extractor* ex, int on) {
    if (!ex) return fail(COEB_ERR_INVALID_ARG, "null extractor");
    CUDA_TRY(cudaSetDevice(ex->device));
    if (on && !ex->ev[0])
        for (int i = 0; i < 7; i++) CUDA_TRY(cudaEventCreate(&ex->ev[i]));
    ex->profiling = on != 0;
    return COEB_OK;
}

int coeb_extractor_stage_ms(coeb_extractor* ex, float* ms6) {
    if (!ex || !ms6 || !ex->ev[0]) return fail(COEB_ERR_INVALID_ARG, "profiling not enabled");
    CUDA_TRY(cudaSetDevice(ex->device));
    CUDA_TRY(cudaEventSynchronize(ex->ev[6]));
    for (int i = 0; i < 6; i++) CUDA_TRY(cudaEventElapsedTime(&ms6[i], ex->ev[i], ex->ev[i + 1]));
    return COEB_OK;
}

// Replays the kernel sequence of `v` as a CUDA graph (captured on first use for this exact view). All pointers in the
// view are the handle's own staging buffers and arenas, so the same graph serves every later call of the same shape.
static bool same_view(const BatchView& a, const BatchView& b) {
    return a.B == b.B && a.l0 == b.l0 && a.l0_pitch == b.l0_pitch && a.l0_stride == b.l0_stride && a.pyr == b.pyr && a.blur == b.blur &&
           a.tabs == b.tabs && a.cand == b.cand && a.keys == b.keys && a.dyn == b.dyn && a.boxes == b.boxes && a.nbox == b.nbox &&
           a.max_box == b.max_box && a.tm == b.tm && a.ntm == b.ntm && a.max_tm == b.max_tm && a.blur_flag == b.blur_flag &&
           a.out_kps == b.out_kps && a.out_desc == b.out_desc && a.out_count == b.out_count && a.status == b.status &&
           a.mirror_hdr == b.mirror_hdr && a.mirror_kps == b.mirror_kps && a.mirror_desc == b.mirror_desc && a.pyr_regions == b.pyr_regions;
}

// Large resident batches run as sub-batches round-robin over the pipeline streams: kernel tails of one sub-batch overlap
// the next one's kernels, and a sub-batch's pyramid is more likely still in L2 when the later stages read it.
static int enqueue_chunked(coeb_extractor* ex, const BatchView& v, cudaStream_t s, int chunk) {
    if (chunk <= 0 || v.B < 2 * chunk) return enqueue(ex, v, s, false);
    int st = ensure_pipe_streams(ex);
    if (st != COEB_OK) return st;
    CUDA_TRY(cudaEventRecord(ex->pipe_ready, s));
    for (int i = 0; i < kPipeStreams; i++) CUDA_TRY(cudaStreamWaitEvent(ex->pipe_stream[i], ex->pipe_ready, 0));
    for (int f0 = 0, c = 0; f0 < v.B; f0 += chunk, c++) {
        st = enqueue(ex, sub_view(ex->geom, v, f0, std::min(chunk, v.B - f0), ex->pyr_bytes_per_frame), ex->pipe_stream[c % kPipeStreams], false);
        if (st != COEB_OK) return st;
    }
    for (int i = 0; i < kPipeStreams; i++) {
        CUDA_TRY(cudaEventRecord(ex->pipe_done[i], ex->pipe_stream[i]));
        CUDA_TRY(cudaStreamWaitEvent(s, ex->pipe_done[i], 0));
    }
    return COEB_OK;
}

static int launch_graphed(coeb_extractor* ex, const BatchView& v, cudaStream_t s, int chunk = 0) {
    const Geometry& g = ex->geom;
    for (auto& e : ex->graphs)
        if (e.w == g.w0 && e.h == g.h0 && e.cap == g.out_cap && e.chunk == chunk && same_view(e.view, v)) {
            CUDA_TRY(cudaGraphLaunch(e.exec, s));
            return COEB_OK;
        }
    if (ex->graph_warm < 2) {   // first calls run eagerly (lazy module loading, function attributes) before anything is captured
        ex->graph_warm++;
        return enqueue_chunked(ex, v, s, chunk);
    }
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    int st = enqueue_chunked(ex, v, s, chunk);
    cudaError_t e = cudaStreamEndCapture(s, &graph);
    if (st != COEB_OK || e != cudaSuccess || !graph) {
        if (getenv("COEB_DEBUG_GRAPH")) fprintf(stderr, "[coeb] graph capture failed: st=%d, %s\n", st, cudaGetErrorString(e));
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return enqueue_chunked(ex, v, s, chunk);   // capture not possible here: plain launches
    }
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        if (getenv("COEB_DEBUG_GRAPH")) fprintf(stderr, "[coeb] graph instantiate failed: %s\n", cudaGetErrorString(e));
        cudaGetLastError();
        return enqueue_chunked(ex, v, s, chunk);
    }
    if (ex->graphs.size() >= 32) { cudaGraphExecDestroy(ex->graphs.front().exec); ex->graphs.erase(ex->graphs.begin()); }
    coeb_extractor::GraphEntry ge;
    ge.view = v; ge.w = g.w0; ge.h = g.h0; ge.cap = g.out_cap; ge.chunk = chunk; ge.exec = exec;
    ex->graphs.push_back(ge);
    CUDA_TRY(cudaGraphLaunch(exec, s));
    return COEB_OK;
}

// Validates, sizes the arenas and fills the full-batch view (no launch).
static int prepare_view(coeb_extractor* ex, int B, const uint8_t* gray, int width, int height, int stride, size_t frame_stride,
                        const float* boxes, const int* nbox, int max_box, const float* tm, const int* ntm, int max_tm, const int* blur_flag,
                        coeb_keypoint* kps_out, uint8_t* desc_out, int* counts_out, int* status_out, int cap, BatchView* out) {
    int st = validate_common(ex, B, gray, width, height, stride, cap);
    if (st != COEB_OK) return st;
    if (!kps_out || !desc_out || !counts_out || !status_out) return fail(COEB_ERR_INVALID_ARG, "null output pointer");
    if (nbox && (!boxes || !blur_flag || max_box < 1)) return fail(COEB_ERR_INVALID_ARG, "nbox given without boxes/blur_flag");
    if (ntm && (!tm || max_tm < 1)) return fail(COEB_ERR_INVALID_ARG, "ntm given without tm");
    CUDA_TRY(cudaSetDevice(ex->device));
    st = build_geometry(ex, width, height);
    if (st != COEB_OK) return st;
    st = ensure_arenas(ex, B);
    if (st != COEB_OK) return st;
    ex->geom.out_cap = cap;
    BatchView v{};
    v.B = B;
    v.l0 = gray; v.l0_pitch = stride; v.l0_stride = frame_stride;
    v.pyr = ex->d_pyr; v.blur = ex->d_blur; v.tabs = ex->d_tabs; v.fast_tiles = ex->d_fast_tiles; v.blur_tiles = ex->d_fast_tiles + (size_t)kFastTileInt4 * ex->geom.fast_tiles_per_frame;
    v.ic_mask = reinterpret_cast<const uint32_t*>(v.blur_tiles + ex->geom.blur_tiles_per_frame);
    v.pyr_regions = ex->d_pyr_regions; v.n_pyr_regions = ex->n_pyr_regions; v.pyr_regions_smem = ex->pyr_regions_smem;
    v.cand = ex->d_cand; v.cand_count = ex->d_cand_count; v.keys = ex->d_keys; v.key_count = ex->d_key_count;
    v.dyn = ex->d_dyn; v.knode = ex->d_knode;
    v.lmax = ex->d_lmax; v.lmax_count = ex->d_lmax_count; v.cell_count = ex->d_cell_count;
    v.empty_cells = ex->d_empty_cells; v.empty_count = ex->d_empty_count;
    v.boxes = boxes; v.nbox = nbox; v.max_box = max_box; v.tm = tm; v.ntm = ntm; v.max_tm = max_tm; v.blur_flag = blur_flag;
    v.out_kps = kps_out; v.out_desc = desc_out; v.out_count = counts_out; v.status = status_out;
    v.trace = ex->d_trace;
    *out = v;
    return COEB_OK;
}

int coeb_extract_batch_device(coeb_extractor* ex, int B, const uint8_t* gray, int width, int height, int stride,
                              size_t frame_stride, const float* boxes, const int* nbox, int max_box, const float* tm,
                              const int* ntm, int max_tm, const int* blur_flag, coeb_keypoint* kps_out,
                              uint8_t* desc_out, int* counts_out, int* status_out, int cap) {
    BatchView v;
    int st = prepare_view(ex, B, gray, width, height, stride, frame_stride, boxes, nbox, max_box, tm, ntm, max_tm, blur_flag, kps_out,
                          desc_out, counts_out, status_out, cap, &v);
    if (st != COEB_OK) return st;
    if (((uintptr_t)gray | (uintptr_t)stride | (uintptr_t)frame_stride) & 15) {
        // The tile loaders read aligned 16-byte vectors. A caller buffer that is not 16-byte aligned in base, row pitch and
        // frame stride (e.g. tightly packed 1241-px rows) is first copied into the arena's pitch-aligned level-0 block.
        const LevelGeom& L0 = ex->geom.lv[0];
        uint8_t* dst = ex->d_pyr + L0.img_base;
        for (int i = 0; i < B; i++)
            CUDA_TRY(cudaMemcpy2DAsync(dst + (size_t)i * L0.img_stride, L0.pitch, gray + (size_t)i * frame_stride, stride, width, height,
                                       cudaMemcpyDeviceToDevice, ex->stream));
        v.l0 = dst; v.l0_pitch = L0.pitch; v.l0_stride = L0.img_stride;
    }
    ex->last_view = v;
    ex->last_B = B;
    ex->last_passes = 1;
    if (ex->profiling) return enqueue(ex, v, ex->stream, true);
    static const int dev_chunk = [] { const char* e = getenv("COEB_DEV_CHUNK"); return e ? atoi(e) : kDevChunk; }();
    if (dev_chunk > 0 && B >= 2 * dev_chunk) ex->last_passes = (B + dev_chunk - 1) / dev_chunk;
    return launch_graphed(ex, v, ex->stream, dev_chunk);
}

int coeb_extract_batch_host(coeb_extractor* ex, int B, const uint8_t* gray, int width, int height, int stride,
                            size_t frame_stride, const float* boxes, const int* nbox, int max_box, const float* tm,
                            const int* ntm, int max_tm, const int* blur_flag, coeb_keypoint* kps_out,
                            uint8_t* desc_out, int* counts_out, int* status_out, int cap) {
    static const bool host_trace = getenv("COEB_HOST_TRACE") != nullptr;   // development: host-side timeline of the call on stderr
    double ht[8] = {0};
    auto stamp = [&](int i) { if (host_trace) ht[i] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    stamp(0);
    int st = validate_common(ex, B, gray, width, height, stride, cap);
    if (st != COEB_OK) return st;
    if (!counts_out) return fail(COEB_ERR_INVALID_ARG, "null counts_out");
    if (nbox && (!boxes || !blur_flag || max_box < 1)) return fail(COEB_ERR_INVALID_ARG, "nbox given without boxes/blur_flag");
    if (ntm && (!tm || max_tm < 1)) return fail(COEB_ERR_INVALID_ARG, "ntm given without tm");
    if (nbox) for (int i = 0; i < B; i++) if (nbox[i] < 0 || nbox[i] > max_box || nbox[i] > COEB_MAX_BOXES)
        return fail(COEB_ERR_INVALID_ARG, "frame %d: nbox=%d (max_box=%d, at most %d boxes per frame)", i, nbox[i], max_box, COEB_MAX_BOXES);
    if (ntm) for (int i = 0; i < B; i++) if (ntm[i] < 0 || ntm[i] > max_tm) return fail(COEB_ERR_INVALID_ARG, "frame %d: ntm=%d > max_tm=%d", i, ntm[i], max_tm);
    CUDA_TRY(cudaSetDevice(ex->device));
    cudaStream_t s = ex->stream;
    // device staging: level 0 goes into a pitch-aligned buffer; outputs into [B][cap] arrays
    const int pitch = (width + 63) & ~63;
    const size_t fstride = (size_t)pitch * height;
    if (fstride * B > ex->in_gray_bytes) drop_graphs(ex);
    st = ensure_buf(&ex->d_in_gray, &ex->in_gray_bytes, fstride * B);
    if (st != COEB_OK) return st;
    if (frame_stride == (size_t)stride * height && !(stride == pitch && stride == width)) {
        if ((st = ensure_buf(&ex->d_in_linear, &ex->in_linear_bytes, frame_stride * B)) != COEB_OK) return st;
    }
    const float *dboxes = nullptr, *dtm = nullptr;
    const int *dnbox = nullptr, *dntm = nullptr, *dblur = nullptr;
    if (nbox || ntm) {   // the per-frame box / T_M arrays are tiny: packed into one pinned block, one copy up front
        auto a256 = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t o_boxes = 0, o_nbox = o_boxes + a256(nbox ? sizeof(float) * B * max_box * 4 : 0), o_blur = o_nbox + a256(nbox ? sizeof(int) * B : 0);
        const size_t o_tm = o_blur + a256(nbox ? sizeof(int) * B * max_box : 0), o_ntm = o_tm + a256(ntm ? sizeof(float) * B * max_tm * 2 : 0);
        const size_t total = o_ntm + a256(ntm ? sizeof(int) * B : 0);
        if (total > ex->dyn_cap) {
            drop_graphs(ex);
            if (ex->h_dynin) cudaFreeHost(ex->h_dynin);
            cudaFree(ex->d_dynin);
            ex->h_dynin = ex->d_dynin = nullptr; ex->dyn_cap = 0;
            CUDA_TRY(cudaHostAlloc((void**)&ex->h_dynin, total, cudaHostAllocDefault));
            CUDA_TRY(cudaMalloc((void**)&ex->d_dynin, total));
            ex->dyn_cap = total;
        }
        if (nbox) {
            std::memcpy(ex->h_dynin + o_boxes, boxes, sizeof(float) * B * max_box * 4);
            std::memcpy(ex->h_dynin + o_nbox, nbox, sizeof(int) * B);
            std::memcpy(ex->h_dynin + o_blur, blur_flag, sizeof(int) * B * max_box);
            dboxes = (const float*)(ex->d_dynin + o_boxes); dnbox = (const int*)(ex->d_dynin + o_nbox); dblur = (const int*)(ex->d_dynin + o_blur);
        }
        if (ntm) {
            std::memcpy(ex->h_dynin + o_tm, tm, sizeof(float) * B * max_tm * 2);
            std::memcpy(ex->h_dynin + o_ntm, ntm, sizeof(int) * B);
            dtm = (const float*)(ex->d_dynin + o_tm); dntm = (const int*)(ex->d_dynin + o_ntm);
        }
        CUDA_TRY(cudaMemcpyAsync(ex->d_dynin, ex->h_dynin, total, cudaMemcpyHostToDevice, s));
    }
    // a single-frame call always comes back as one block through pinned memory (only the rows that exist are then copied on to
    // the caller), so its block is laid out for exactly (1, cap)
    if ((size_t)B * cap > ex->out_cap_elems || B > ex->out_cap_B || (B == 1 && (ex->out_cap_B != 1 || ex->out_cap_elems != (size_t)cap))) {
        drop_graphs(ex);
        cudaFree(ex->d_out_block);
        ex->d_out_block = nullptr; ex->d_out_kps = nullptr; ex->d_out_desc = nullptr; ex->d_out_count = ex->d_out_status = nullptr;
        // counts | status | keypoints | descriptors in one block: a single-frame call downloads it with one copy
        const size_t o_kps = ((size_t)2 * B * sizeof(int) + 255) & ~(size_t)255;
        const size_t o_desc = o_kps + (((size_t)B * cap * sizeof(coeb_keypoint) + 255) & ~(size_t)255);
        CUDA_TRY(cudaMalloc((void**)&ex->d_out_block, o_desc + (size_t)B * cap * 32));
        ex->d_out_count = (int*)ex->d_out_block;
        ex->d_out_status = ex->d_out_count + B;
        ex->d_out_kps = (coeb_keypoint*)(ex->d_out_block + o_kps);
        ex->d_out_desc = (uint8_t*)(ex->d_out_block + o_desc);
        ex->out_cap_elems = (size_t)B * cap;
        ex->out_cap_B = B;
    }
    BatchView v;
    st = prepare_view(ex, B, ex->d_in_gray, width, height, pitch, fstride, dboxes, dnbox, max_box, dtm, dntm, max_tm, dblur, ex->d_out_kps,
                      ex->d_out_desc, ex->d_out_count, ex->d_out_status, cap, &v);
    if (st != COEB_OK) return st;
    // single frame: the results land in mapped pinned memory straight from the descriptor stage, no copy node
    const bool mirrored = B == 1 && ex->out_cap_B == 1 && ex->out_cap_elems == (size_t)cap && !getenv("COEB_NO_MIRROR");
    if (mirrored) {
        const size_t bytes = (size_t)((char*)ex->d_out_desc - ex->d_out_block) + (size_t)cap * 32;
        if (bytes > ex->h_out1_cap) {
            drop_graphs(ex);
            if (ex->h_out1) cudaFreeHost(ex->h_out1);
            ex->h_out1 = nullptr; ex->h_out1_cap = 0;
            CUDA_TRY(cudaHostAlloc((void**)&ex->h_out1, bytes, cudaHostAllocMapped));
            ex->h_out1_cap = bytes;
        }
        v.mirror_hdr = (int*)ex->h_out1;
        v.mirror_kps = (coeb_keypoint*)(ex->h_out1 + ((char*)ex->d_out_kps - ex->d_out_block));
        v.mirror_desc = (uint8_t*)(ex->h_out1 + ((char*)ex->d_out_desc - ex->d_out_block));
    }
    ex->last_view = v;
    ex->last_B = B;
    std::vector<int> status_local;
    int* hstatus = status_out;
    if (!hstatus) { status_local.resize(B); hstatus = status_local.data(); }

    // Sub-batches of kPipeChunk frames: all H2D copies go back to back on one copy stream, the kernels of sub-batch c run on
    // compute stream c % kPipeStreams once its input has landed, and the D2H copies follow on a second copy stream. The H2D
    // engine never waits for a kernel (the arenas are frame-major, so a sub-batch is a pointer offset and no buffer is reused
    // inside a call).
    // 64-frame sub-batches from 256 frames on (a 64-frame launch fills the GPU better than a 32-frame one: 134k against 127k
    // frames/s device-resident), 32-frame ones below; COEB_PIPE_CHUNK overrides (development).
    static const int pipe_chunk_env = [] { const char* e = getenv("COEB_PIPE_CHUNK"); int v = e ? atoi(e) : 0; return v > 0 ? v : 0; }();
    const int pipe_chunk = pipe_chunk_env ? pipe_chunk_env : (B >= 256 ? 2 * kPipeChunk : kPipeChunk);
    const int chunk = B <= 2 * kPipeChunk || B <= pipe_chunk ? B : pipe_chunk;
    const int nchunks = (B + chunk - 1) / chunk;
    const bool piped = nchunks > 1;
    ex->last_passes = nchunks;
    if (piped) {
        if ((st = ensure_pipe_streams(ex)) != COEB_OK) return st;
        if ((st = ensure_chunk_events(ex, nchunks)) != COEB_OK) return st;
    }
    static const bool trace = getenv("COEB_PIPE_TRACE") != nullptr;   // development: per-chunk device timeline on stderr
    std::vector<cudaEvent_t> tev;
    if (trace) {
        tev.resize(1 + 3 * (size_t)nchunks);
        for (auto& e : tev) cudaEventCreate(&e);
        cudaEventRecord(tev[0], s);
    }
    auto upload = [&](int f0, int n, cudaStream_t cs) -> cudaError_t {
        if (B == 1 && fstride <= (size_t)(8u << 20) && !getenv("COEB_NO_STAGE")) {
            // A single frame in pageable memory: cudaMemcpyAsync would stage it inside the driver and block for the whole copy
            // (~25 us for 300 KB). Staging it here, in two halves, lets the second half's memcpy and the graph launch overlap the DMA,
            // and lays rows of any stride out in the arena pitch on the way (no device-side re-pitch).
            cudaPointerAttributes pa;
            if (cudaPointerGetAttributes(&pa, gray) != cudaSuccess) { cudaGetLastError(); pa.type = cudaMemoryTypeUnregistered; }
            if (pa.type == cudaMemoryTypeUnregistered) {
                if (fstride > ex->h_in1_cap) {
                    if (ex->h_in1) cudaFreeHost(ex->h_in1);
                    ex->h_in1 = nullptr; ex->h_in1_cap = 0;
                    if (cudaHostAlloc((void**)&ex->h_in1, fstride, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); ex->h_in1 = nullptr; }
                    else ex->h_in1_cap = fstride;
                }
                if (ex->h_in1) {
                    const int half = (height + 1) / 2;
                    for (int part = 0; part < 2; part++) {
                        const int y0 = part * half, y1 = std::min(height, y0 + half);
                        if (y1 <= y0) break;
                        // (the caller's last row may end after `width` bytes: its padding is not read)
                        if (stride == pitch) std::memcpy(ex->h_in1 + (size_t)y0 * pitch, gray + (size_t)y0 * stride, (size_t)(y1 - y0 - 1) * pitch + width);
                        else for (int y = y0; y < y1; y++) std::memcpy(ex->h_in1 + (size_t)y * pitch, gray + (size_t)y * stride, width);
                        cudaError_t e = cudaMemcpyAsync(ex->d_in_gray + (size_t)y0 * pitch, ex->h_in1 + (size_t)y0 * pitch, (size_t)(y1 - y0) * pitch, cudaMemcpyHostToDevice, cs);
                        if (e != cudaSuccess) return e;
                    }
                    return cudaSuccess;
                }
            }
        }
        if (frame_stride == (size_t)stride * height && stride == pitch && stride == width)
            // tightly packed frames whose width is already the arena pitch: one linear DMA (2D copies go row by row)
            return cudaMemcpyAsync(ex->d_in_gray + fstride * f0, gray + frame_stride * f0, fstride * n, cudaMemcpyHostToDevice, cs);
        if (frame_stride == (size_t)stride * height && ex->d_in_linear) {
            // contiguous rows of another pitch (e.g. 1241-pixel KITTI rows): one linear DMA into a landing buffer, then a device-side
            // 2-D copy into the pitch-aligned arena (a host-to-device 2-D copy goes row by row and is several times slower)
            cudaError_t e = cudaMemcpyAsync(ex->d_in_linear + frame_stride * f0, gray + frame_stride * f0, frame_stride * n, cudaMemcpyHostToDevice, cs);
            if (e != cudaSuccess) return e;
            return cudaMemcpy2DAsync(ex->d_in_gray + fstride * f0, pitch, ex->d_in_linear + frame_stride * f0, stride, width, (size_t)height * n,
                                     cudaMemcpyDeviceToDevice, cs);
        }
        if (frame_stride == (size_t)stride * height)
            return cudaMemcpy2DAsync(ex->d_in_gray + fstride * f0, pitch, gray + frame_stride * f0, stride, width, (size_t)height * n,
                                     cudaMemcpyHostToDevice, cs);
        for (int i = f0; i < f0 + n; i++) {
            cudaError_t e = cudaMemcpy2DAsync(ex->d_in_gray + fstride * i, pitch, gray + frame_stride * i, stride, width, height,
                                              cudaMemcpyHostToDevice, cs);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    auto download = [&](int f0, int n, cudaStream_t cs) -> cudaError_t {
        cudaError_t e = cudaMemcpyAsync(counts_out + f0, ex->d_out_count + f0, sizeof(int) * n, cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hstatus + f0, ex->d_out_status + f0, sizeof(int) * n, cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess && kps_out)
            e = cudaMemcpyAsync(kps_out + (size_t)f0 * cap, ex->d_out_kps + (size_t)f0 * cap, sizeof(coeb_keypoint) * n * cap, cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess && desc_out)
            e = cudaMemcpyAsync(desc_out + (size_t)f0 * cap * 32, ex->d_out_desc + (size_t)f0 * cap * 32, (size_t)32 * n * cap, cudaMemcpyDeviceToHost, cs);
        return e;
    };
    bool landed = false;   // single frame: the descriptor stage wrote the whole output block into mapped pinned memory
    if (!piped) {
        stamp(1);
        CUDA_TRY(upload(0, B, s));
        stamp(2);
        if (trace) cudaEventRecord(tev[1], s);
        if (ex->d_trace) {   // development: stamps reset before the kernels (min slots all ones, max slots zero)
            unsigned long long init[32];
            for (int i = 0; i < 32; i++) init[i] = (i & 1) || i >= 24 ? 0ull : ~0ull;
            CUDA_TRY(cudaMemcpyAsync(ex->d_trace, init, sizeof(init), cudaMemcpyHostToDevice, s));
        }
        st = ex->profiling ? enqueue(ex, v, s, true) : launch_graphed(ex, v, s);   // one graph launch instead of ~20 stream operations
        if (st != COEB_OK) return st;
        if (trace) cudaEventRecord(tev[2], s);
        stamp(3);
        if (mirrored) landed = true;
        else CUDA_TRY(download(0, B, s));
        if (trace) cudaEventRecord(tev[3], s);
    } else {
        CUDA_TRY(cudaEventRecord(ex->pipe_ready, s));   // the box / T_M copies above, and whatever the caller queued before
        CUDA_TRY(cudaStreamWaitEvent(ex->copy_in, ex->pipe_ready, 0));
        for (int c = 0; c < nchunks; c++) {
            const int f0 = c * chunk, n = std::min(chunk, B - f0);
            CUDA_TRY(upload(f0, n, ex->copy_in));
            CUDA_TRY(cudaEventRecord(ex->chunk_in[c], ex->copy_in));
            if (trace) cudaEventRecord(tev[1 + 3 * c], ex->copy_in);
        }
        for (int c = 0; c < nchunks; c++) {
            cudaStream_t ps = ex->pipe_stream[c % kPipeStreams];
            const int f0 = c * chunk, n = std::min(chunk, B - f0);
            CUDA_TRY(cudaStreamWaitEvent(ps, ex->chunk_in[c], 0));
            const BatchView sv = sub_view(ex->geom, v, f0, n, ex->pyr_bytes_per_frame);
            // the staging buffers and arenas are the handle's own, so a sub-view recurs call after call: graph replay
            // (a call with more sub-batches than the cache holds would re-capture every one of them each time: plain launches)
            st = (ex->profiling || nchunks > 24) ? enqueue(ex, sv, ps, false) : launch_graphed(ex, sv, ps);
            if (st != COEB_OK) return st;
            CUDA_TRY(cudaEventRecord(ex->chunk_done[c], ps));
            if (trace) cudaEventRecord(tev[2 + 3 * c], ps);
            CUDA_TRY(cudaStreamWaitEvent(ex->copy_out, ex->chunk_done[c], 0));
            CUDA_TRY(download(f0, n, ex->copy_out));
            if (trace) cudaEventRecord(tev[3 + 3 * c], ex->copy_out);
        }
        CUDA_TRY(cudaEventRecord(ex->copy_out_done, ex->copy_out));
        CUDA_TRY(cudaStreamWaitEvent(s, ex->copy_out_done, 0));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    stamp(4);
    if (trace) {
        for (int c = 0; c < nchunks; c++) {
            float a = 0, b = 0, d = 0;
            cudaEventElapsedTime(&a, tev[0], tev[1 + 3 * c]);
            cudaEventElapsedTime(&b, tev[0], tev[2 + 3 * c]);
            cudaEventElapsedTime(&d, tev[0], tev[3 + 3 * c]);
            fprintf(stderr, "[coeb pipe] chunk %d: h2d done %.3f ms, kernels done %.3f ms, d2h done %.3f ms\n", c, a, b, d);
        }
        for (auto& e : tev) cudaEventDestroy(e);
    }
    if (ex->d_trace && !piped) {
        static const char* names[16] = {"classify", "fast L0", "fast L1..", "fallback L0", "fallback L1..", "octree L0", "octree L1..", "pyramid regions",
                                        "resize chain", "blur", "describe", "empty-cell list", "", "", "", ""};
        unsigned long long t[32];
        cudaMemcpy(t, ex->d_trace, sizeof(t), cudaMemcpyDeviceToHost);
        unsigned long long t0 = ~0ull;
        for (int i = 0; i < 12; i++) if (t[2 * i + 1]) t0 = std::min(t0, t[2 * i]);
        fprintf(stderr, "[coeb kernels]");
        for (int i = 0; i < 12; i++)
            if (t[2 * i + 1]) fprintf(stderr, " %s %.1f-%.1f |", names[i], (double)(t[2 * i] - t0) * 1e-3, (double)(t[2 * i + 1] - t0) * 1e-3);
        fprintf(stderr, " marks (octree L0: gathered, roots, rounds[%d], best):", (int)t[31]);
        for (int i = 24; i < 31; i++) if (t[i]) fprintf(stderr, " %.1f", (double)(t[i] - t0) * 1e-3);
        fprintf(stderr, "\n");
    }
    if (landed) {   // only the keypoints that exist are copied on to the caller's arrays
        const int n = ((const int*)ex->h_out1)[0];
        counts_out[0] = n;
        hstatus[0] = ((const int*)ex->h_out1)[1];
        // on COEB_ERR_CAPACITY the device wrote no keypoint (only the required count): nothing to copy
        const int m = hstatus[0] == COEB_OK ? std::min(std::max(n, 0), cap) : 0;
        if (kps_out && m) std::memcpy(kps_out, ex->h_out1 + ((char*)ex->d_out_kps - ex->d_out_block), sizeof(coeb_keypoint) * m);
        if (desc_out && m) std::memcpy(desc_out, ex->h_out1 + ((char*)ex->d_out_desc - ex->d_out_block), (size_t)32 * m);
    }
    stamp(5);
    if (host_trace && !piped)
        fprintf(stderr, "[coeb host] set-up %.1f us | upload enqueued %.1f | kernels enqueued %.1f | synchronised %.1f | results copied %.1f\n", ht[1] - ht[0], ht[2] - ht[0],
                ht[3] - ht[0], ht[4] - ht[0], ht[5] - ht[0]);
    int worst = COEB_OK;
    for (int i = 0; i < B; i++)
        if (hstatus[i] != COEB_OK && worst == COEB_OK) {
            worst = hstatus[i];
            fail(worst, "frame %d: %s", i,
                 worst == COEB_ERR_BAD_BOX ? "a box lies outside the image (the reference's cv::Mat ROI would throw)"
                 : worst == COEB_ERR_CAPACITY ? "output capacity too small" : "device-side failure");
        }
    return worst;
}

int coeb_extractor_max_keypoints(const coeb_extractor* ex, int width, int height, int* max_out) {
    if (!ex || !max_out || width < 1 || height < 1) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    int total = 0;
    for (int l = 0; l < ex->params.nlevels; l++) {   // same level geometry as build_geometry()
        const int lw = round_half_even((float)width * ex->inv_scale[l]), lh = round_half_even((float)height * ex->inv_scale[l]);
        const int bw = lw - kEdge + 3 - kMinBorder, bh = lh - kEdge + 3 - kMinBorder;
        if (bw < 1 || bh < 1) return fail(COEB_ERR_UNSUPPORTED, "level %d (%dx%d) is smaller than the FAST border", l, lw, lh);
        const int n_ini = (int)std::round((float)bw / (float)bh);
        total += std::max(ex->per_level[l] + 3, 4 * n_ini);
    }
    *max_out = total;
    return COEB_OK;
}

int coeb_extract(coeb_extractor* ex, const uint8_t* gray, int width, int height, int stride, const float* boxes_xyxy,
                 int nbox, const float* tm_xy, int ntm, const int* blur_flag, int nblur, coeb_keypoint* kps_out,
                 uint8_t* desc_out, int cap, int* n_out) {
    if (n_out) *n_out = 0;
    if (!ex) return fail(COEB_ERR_INVALID_ARG, "null extractor");
    if (!gray || width <= 0 || height <= 0) return COEB_OK;  // `if (_image.empty()) return;` (src/ORBextractor.cc:1096)
    if (nbox < 0 || ntm < 0 || nblur < 0) return fail(COEB_ERR_INVALID_ARG, "negative count");
    if (nbox > COEB_MAX_BOXES) return fail(COEB_ERR_INVALID_ARG, "at most %d boxes per frame", COEB_MAX_BOXES);
    if ((nbox > 0 && !boxes_xyxy) || (ntm > 0 && !tm_xy) || (nblur > 0 && !blur_flag)) return fail(COEB_ERR_INVALID_ARG, "a count is positive but its array is null");
    // Fixed-shape staging (COEB_MAX_BOXES boxes, T_M capacity rounded up to 64) so that every call of this handle has the
    // same device view and replays the same CUDA graph. blur_flag is indexed by box id in the reference (:1168);
    // missing entries count as 0.
    float boxes[COEB_MAX_BOXES * 4] = {0};
    int blur[COEB_MAX_BOXES] = {0};
    for (int i = 0; i < nbox * 4; i++) boxes[i] = boxes_xyxy[i];
    for (int i = 0; i < nbox && i < nblur; i++) blur[i] = blur_flag[i];
    const int max_tm = std::max(64, (ntm + 63) & ~63);
    std::vector<float> tm((size_t)max_tm * 2, 0.f);
    if (ntm) std::memcpy(tm.data(), tm_xy, sizeof(float) * 2 * ntm);
    int count = 0, status = 0;
    const int nb = nbox, nt = ntm;
    int st = coeb_extract_batch_host(ex, 1, gray, width, height, stride, (size_t)stride * height, boxes, &nb, COEB_MAX_BOXES, tm.data(), &nt,
                                     max_tm, blur, kps_out, desc_out, &count, &status, cap);
    if (n_out) *n_out = count;
    return st;
}

int coeb_extractor_dyn_info(coeb_extractor* ex, int frame, coeb_dyn_info* out) {
    if (!ex || !out || frame < 0 || frame >= ex->last_B) return fail(COEB_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(ex->device));
    CUDA_TRY(cudaStreamSynchronize(ex->stream));
    DynState d;
    CUDA_TRY(cudaMemcpy(&d, ex->d_dyn + frame, sizeof(d), cudaMemcpyDeviceToHost));
    out->area_flag = d.area_flag;
    out->n_dynamic = d.n_dynamic;
    std::memcpy(out->rect, d.rect, sizeof(d.rect));
    out->area = d.area;
    return COEB_OK;
}

int coeb_pyramid_level(coeb_extractor* ex, int frame, int level, int blurred, const uint8_t** dev_ptr, int* width,
                       int* height, int* pitch) {
    if (!ex || !ex->geom_valid || frame < 0 || frame >= ex->last_B || level < 0 || level >= ex->geom.nlevels)
        return fail(COEB_ERR_INVALID_ARG, "bad argument");
    const Geometry& g = ex->geom;
    const BatchView& v = ex->last_view;
    if (dev_ptr) *dev_ptr = blurred ? blur_ptr(g, v, level, frame) : level_ptr(g, v, level, frame);
    if (width) *width = g.lv[level].w;
    if (height) *height = g.lv[level].h;
    if (pitch) *pitch = blurred ? g.lv[level].pitch : level_pitch(g, v, level);
    return COEB_OK;
}

int coeb_pyramid_level_copy(coeb_extractor* ex, int frame, int level, int blurred, uint8_t* host_dst) {
    const uint8_t* p = nullptr;
    int w = 0, h = 0, pitch = 0;
    int st = coeb_pyramid_level(ex, frame, level, blurred, &p, &w, &h, &pitch);
    if (st != COEB_OK) return st;
    CUDA_TRY(cudaSetDevice(ex->device));
    CUDA_TRY(cudaStreamSynchronize(ex->stream));
    CUDA_TRY(cudaMemcpy2D(host_dst, w, p, pitch, w, h, cudaMemcpyDeviceToHost));
    return COEB_OK;
}

int coeb_debug_candidates(coeb_extractor* ex, int frame, int level, uint32_t* host_out, int cap, int* n_out) {
    if (!ex || !ex->geom_valid || frame < 0 || frame >= ex->last_B || level < 0 || level >= ex->geom.nlevels || !n_out)
        return fail(COEB_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(ex->device));
    CUDA_TRY(cudaStreamSynchronize(ex->stream));
    const Geometry& g = ex->geom;
    int n = 0;
    CUDA_TRY(cudaMemcpy(&n, ex->d_cand_count + frame * g.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    *n_out = n;
    if (n > cap) return fail(COEB_ERR_CAPACITY, "need %d entries", n);
    if (n > 0 && host_out)
        CUDA_TRY(cudaMemcpy(host_out, ex->d_cand + (size_t)frame * g.cand_per_frame + g.lv[level].cand_base, sizeof(uint32_t) * n,
                            cudaMemcpyDeviceToHost));
    return COEB_OK;
}

int coeb_debug_level_keys(coeb_extractor* ex, int frame, int level, float* host_out, int cap, int* n_out) {
    if (!ex || !ex->geom_valid || frame < 0 || frame >= ex->last_B || level < 0 || level >= ex->geom.nlevels || !n_out)
        return fail(COEB_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(ex->device));
    CUDA_TRY(cudaStreamSynchronize(ex->stream));
    const Geometry& g = ex->geom;
    int n = 0;
    CUDA_TRY(cudaMemcpy(&n, ex->d_key_count + frame * g.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    *n_out = n;
    if (n > cap) return fail(COEB_ERR_CAPACITY, "need %d entries", n);
    if (n > 0 && host_out)
        CUDA_TRY(cudaMemcpy(host_out, ex->d_keys + (size_t)frame * g.keys_per_frame + g.lv[level].key_base, sizeof(LevelKey) * n,
                            cudaMemcpyDeviceToHost));
    return COEB_OK;
}

}  // extern "C"
