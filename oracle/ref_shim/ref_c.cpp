// ref_c.cpp -- C entry points around the reference's OWN ORBextractor, compiled unchanged from
// /root/reference/src/ORBextractor.cc into oracle/_ref/ (see oracle/Makefile, target _ref).
//
// TEST INFRASTRUCTURE ONLY: loaded by tests/ (oracle B == _ref), by bench.py's reference arm / cpu_baseline leg and by
// nothing else. The reference class is driven exactly the way src/Frame.cc:413-419 drives it:
//   (*extractor)(im, cv::Mat(), img, imD, mvKeys, mDescriptors, box, T_M, mask_frame, blur_flag)
//
// What this wrapper adds around the unchanged code, and why:
//   * the reference hard-codes a 480x640 mask (src/ORBextractor.cc:1103,1123) and never checks a box: frames that
//     carry boxes must be 640x480 and the boxes inside the image, else the call is refused here (the real cv::Mat ROI
//     would throw / the fill loops would write out of bounds);
//   * frames without boxes may have any size for which the reference's own mask lookups stay inside its 480x640
//     buffer ((rows-1)*640 + cols-1 < 307200): true for 1241x376, false for 1920x1080 (the reference reads out of
//     bounds there), so 1080p frames are refused;
//   * blur_flag is indexed by box id (:1168) although Frame.cc:205-209 may pass fewer entries: padded with zeros.
//
// -DCOEB_REF_MONOTONIC_HEAP builds the "monotonic heap" variant: operator new inside this library hands out strictly
// increasing addresses during one extract call and never reuses one. DistributeOctTree sorts
// pair<int, ExtractorNode*> (src/ORBextractor.cc:691), so nodes of equal size are expanded in heap-address order; with
// glibc's malloc that order depends on which freed chunks get recycled. Under the monotonic heap "higher address" is
// exactly "created later", the documented rule of oracle B and of the CUDA path, so this variant isolates every OTHER
// possible difference: oracle B must equal it on every frame. The plain build shows what glibc's recycling does on top.
#include <atomic>
#include <chrono>
#include <new>
#include <thread>
#include <sys/mman.h>

#include "cvshim.hpp"
#include "ORBextractor.h"
#include "../../include/coeb_types.h"
#include "ref_handles.hpp"

#ifdef COEB_REF_MONOTONIC_HEAP
namespace {
struct Arena {
    char* base = nullptr;
    size_t cap = 0, off = 0;
    bool active = false;
    long fallbacks = 0;
};
thread_local Arena g_arena;
const size_t kArenaBytes = (size_t)2 << 30;   // virtual; pages are touched on demand and reused after every reset

void arena_begin() {
    Arena& a = g_arena;
    if (!a.base) {
        void* p = mmap(nullptr, kArenaBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p != MAP_FAILED) { a.base = (char*)p; a.cap = kArenaBytes; }
    }
    a.off = 0;
    a.active = a.base != nullptr;
}
void arena_end() { g_arena.active = false; }
inline void* arena_alloc(size_t n) {
    Arena& a = g_arena;
    if (a.active) {
        const size_t need = (n + 15) & ~(size_t)15;
        if (a.off + need <= a.cap) { void* p = a.base + a.off; a.off += need; return p; }
        a.fallbacks++;
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
inline void arena_free(void* p) {
    if (!p) return;
    const Arena& a = g_arena;
    if (a.base && (char*)p >= a.base && (char*)p < a.base + a.cap) return;   // arena memory is never recycled
    std::free(p);
}
}  // namespace
void* operator new(size_t n) { return arena_alloc(n); }
void* operator new[](size_t n) { return arena_alloc(n); }
void operator delete(void* p) noexcept { arena_free(p); }
void operator delete[](void* p) noexcept { arena_free(p); }
void operator delete(void* p, size_t) noexcept { arena_free(p); }
void operator delete[](void* p, size_t) noexcept { arena_free(p); }
#else
static void arena_begin() {}
static void arena_end() {}
#endif

using ORB_SLAM2::ORBextractor;


static int check_inputs(int w, int h, const float* boxes, int nbox) {
    if (w < 64 || h < 64) return COEB_ERR_INVALID_ARG;
    if ((size_t)(h - 1) * 640 + (size_t)(w - 1) >= (size_t)480 * 640) return COEB_ERR_UNSUPPORTED;
    if (nbox > 0) {
        if (w != 640 || h != 480) return COEB_ERR_UNSUPPORTED;
        for (int b = 0; b < nbox; b++) {
            const float xmin = boxes[4 * b], ymin = boxes[4 * b + 1], xmax = boxes[4 * b + 2], ymax = boxes[4 * b + 3];
            const int rx = (int)xmin, ry = (int)ymin, rw = (int)(xmax - xmin), rh = (int)(ymax - ymin);
            if (rx < 0 || ry < 0 || rw < 0 || rh < 0 || rx + rw > w || ry + rh > h) return COEB_ERR_BAD_BOX;
            if ((int)xmax > w || (int)ymax > h) return COEB_ERR_BAD_BOX;
        }
    }
    return COEB_OK;
}

// One call of the reference operator() (src/ORBextractor.cc:1088) with the containers src/Frame.cc hands it.
static int run_reference(ORBextractor& ex, const uint8_t* gray, int w, int h, int stride, const float* boxes, int nbox,
                         const float* tm, int ntm, const int* blur_flag, int nblur, std::vector<cv::KeyPoint>& keys,
                         cv::Mat& desc) {
    cv::Mat im(h, w, CV_8UC1, (void*)gray, (size_t)stride);
    cv::Mat img, imD, mask_frame;   // colour image (debug drawing only), depth (unused), mask_result (never written)
    std::vector<std::vector<float>> box((size_t)nbox, std::vector<float>(4));
    for (int b = 0; b < nbox; b++)
        for (int k = 0; k < 4; k++) box[b][k] = boxes[4 * b + k];
    std::vector<cv::Point2f> T_M((size_t)ntm);
    for (int t = 0; t < ntm; t++) T_M[t] = cv::Point2f(tm[2 * t], tm[2 * t + 1]);
    std::vector<int> bf((size_t)std::max(nbox, nblur), 0);
    for (int b = 0; b < nblur; b++) bf[b] = blur_flag[b];
    ex(im, cv::Mat(), img, imD, keys, desc, box, T_M, mask_frame, bf);
    return COEB_OK;
}

extern "C" {

const char* ref_build_info() {
#ifdef COEB_REF_MONOTONIC_HEAP
    return "reference src/ORBextractor.cc unchanged; OpenCV shim; monotonic heap";
#else
    return "reference src/ORBextractor.cc unchanged; OpenCV shim; glibc heap";
#endif
}

ref_extractor* ref_extractor_create(const coeb_orb_params* p) {
    ref_extractor* e = new ref_extractor;
    e->ex = new ORBextractor(p->nfeatures, p->scale_factor, p->nlevels, p->ini_th_fast, p->min_th_fast);
    e->nlevels = p->nlevels;
    return e;
}
void ref_extractor_destroy(ref_extractor* e) {
    if (!e) return;
    delete e->ex;
    delete e;
}

void ref_extractor_tables(ref_extractor* e, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2) {
    const std::vector<float> s = e->ex->GetScaleFactors(), is = e->ex->GetInverseScaleFactors(), s2 = e->ex->GetScaleSigmaSquares(),
                             is2 = e->ex->GetInverseScaleSigmaSquares();
    for (int i = 0; i < e->nlevels; i++) {
        if (scale) scale[i] = s[i];
        if (inv_scale) inv_scale[i] = is[i];
        if (sigma2) sigma2[i] = s2[i];
        if (inv_sigma2) inv_sigma2[i] = is2[i];
    }
}

int ref_extract(ref_extractor* e, const uint8_t* gray, int w, int h, int stride, const float* boxes, int nbox, const float* tm,
                int ntm, const int* blur_flag, int nblur, coeb_keypoint* kps_out, uint8_t* desc_out, int cap, int* n_out) {
    if (n_out) *n_out = 0;
    const int st = check_inputs(w, h, boxes, nbox);
    if (st != COEB_OK) return st;
    e->w0 = w; e->h0 = h;
    int n = 0;
    arena_begin();
    {
        std::vector<cv::KeyPoint> keys;
        cv::Mat desc;
        run_reference(*e->ex, gray, w, h, stride, boxes, nbox, tm, ntm, blur_flag, nblur, keys, desc);
        n = (int)keys.size();
        if (n <= cap) {
            for (int i = 0; i < n; i++) {
                const cv::KeyPoint& k = keys[i];
                if (kps_out) kps_out[i] = coeb_keypoint{k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id};
                if (desc_out) std::memcpy(desc_out + (size_t)32 * i, desc.ptr(i), 32);
            }
        }
    }
    arena_end();
    if (n_out) *n_out = n;
    return n > cap ? COEB_ERR_CAPACITY : COEB_OK;
}

// The reference's public mvImagePyramid (include/ORBextractor.h:99) after the last call.
int ref_level_size(ref_extractor* e, int level, int* w, int* h) {
    if (level < 0 || level >= e->nlevels || e->ex->mvImagePyramid[level].empty()) return -1;
    *w = e->ex->mvImagePyramid[level].cols; *h = e->ex->mvImagePyramid[level].rows;
    return 0;
}
// border = 0: the level itself; border = 19: with the EDGE_THRESHOLD frame the reference builds around it (:1347-1364)
int ref_level_image(ref_extractor* e, int level, int border, uint8_t* dst) {
    const cv::Mat& m = e->ex->mvImagePyramid[level];
    if (m.empty() || (border != 0 && border != 19)) return 0;
    const int w = m.cols + 2 * border, h = m.rows + 2 * border;
    for (int y = 0; y < h; y++) std::memcpy(dst + (size_t)y * w, m.data + (ptrdiff_t)(y - border) * (ptrdiff_t)m.step - border, (size_t)w);
    return w * h;
}

long ref_heap_fallbacks() {
#ifdef COEB_REF_MONOTONIC_HEAP
    return g_arena.fallbacks;
#else
    return -1;
#endif
}

int ref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

// Multi-threaded batch for the CPU arm of bench.py: one reference extractor per thread (the class is not re-entrant),
// frames handed out dynamically. Same packing as orc_extract_batch_mt. Returns wall seconds.
double ref_extract_batch_mt(const coeb_orb_params* p, int B, const uint8_t* gray, int w, int h, const float* boxes, const int* nbox,
                            int max_box, const float* tm, const int* ntm, int max_tm, const int* blur, int nthreads, int cap,
                            int* counts, coeb_keypoint* kps, uint8_t* desc) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next(0);
    std::vector<ref_extractor*> exs((size_t)nthreads);
    for (auto& e : exs) e = ref_extractor_create(p);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&, t]() {
            for (;;) {
                const int f = next.fetch_add(1);
                if (f >= B) break;
                int n = 0;
                ref_extract(exs[t], gray + (size_t)f * w * h, w, h, w, boxes + (size_t)f * max_box * 4, nbox[f],
                            tm + (size_t)f * max_tm * 2, ntm[f], blur + (size_t)f * max_box, nbox[f],
                            kps ? kps + (size_t)f * cap : nullptr, desc ? desc + (size_t)f * cap * 32 : nullptr, cap, &n);
                counts[f] = n;
            }
        });
    for (auto& x : th) x.join();
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (auto& e : exs) ref_extractor_destroy(e);
    return s;
}

}  // extern "C"
