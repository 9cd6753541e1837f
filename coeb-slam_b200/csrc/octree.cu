// octree.cu -- DistributeOctTree (E6), border/scale fix-up (E7) and the post-octree cull (E9), one CTA per
// (level, frame). IC_Angle (E8) lives in describe.cu.
//
// Reference: ExtractorNode::DivideNode + ORBextractor::DistributeOctTree (src/ORBextractor.cc:489-769),
// CheckMovingKeyPoints_finall (:1371-1408).
//
// The reference grows a std::list of nodes by repeatedly splitting every multi-key node ("full
// pass"), and, once one more pass could overshoot N, by splitting nodes in descending size order
// until the list holds N nodes ("careful phase"). Its result depends on (a) which keys share a node
// and (b) the list order, which fixes the output order and therefore descriptor row indices.
// Both are reproduced exactly, but with data-parallel steps:
//   * a node is a rectangle; a key's quadrant under a node depends only on the key and the node, so
//     every key finds its child independently and children are counted with shared-memory atomics;
//   * node ids ARE list positions. A round (full or careful) turns the list L into
//         reverse(children created this round, in creation order) ++ (L minus the split nodes)
//     (children are push_front'ed as created, :630-667 / :698-733), which is prefix COUNTS of flags: taken with warp ballots,
//     the three counts of a round packed into one word per warp, one barrier per scan (block_flag_scan and the round body);
//   * per round a key makes two table look-ups: node id -> split lines (s_split, or "not split") gives its quadrant, then
//     (node id, quadrant) -> position in the new list (s_newpos). Both tables are indexed by node id, written by the few threads
//     that own nodes; the passes over the keys carry no slot indirection and no branch;
//   * the careful phase's sort of (size, pointer) pairs (:691) uses the creation index as the second
//     key (documented tie-break, identical to the CPU oracle) and is done by rank counting;
//   * the per-node "best response, first key wins" (:748-766) is a 64-bit atomicMax on
//     (response, reversed candidate order), the candidate order being re-derived from (x, y).
#include "coeb_device.cuh"

namespace coeb {

// CTA size / resident CTAs per SM. Batches: 256 x 4 (measured after the ballot scans and the node-indexed tables: stage 0.103 ms
// per 256 frames against 0.129 with 384 x 3, 0.107 with 192 x 5, 0.109 with 128 x 6, 0.170 with 512 x 2, and the step follows:
// 1.355 against 1.381 ms). A single frame is one CTA per level and pure latency: 384 threads give 105.6 us per call, 256 give 110.3.
#ifndef COEB_SEL_THREADS
#define COEB_SEL_THREADS 256
#endif
#ifndef COEB_SEL_MINB
#define COEB_SEL_MINB 4
#endif
#ifndef COEB_SEL_THREADS_SMALL
#define COEB_SEL_THREADS_SMALL 384
#endif
#ifndef COEB_SEL_GU
#define COEB_SEL_GU 4   // candidates per thread in flight in the gather phase
#endif
#ifndef COEB_SEL_KU
#define COEB_SEL_KU 2   // keys per thread in flight in the two key passes of a round
#endif
constexpr int kSelThreads = COEB_SEL_THREADS;
constexpr int kKeyCache = 4096;   // candidates per (level, frame) kept in shared memory (4 + 2 bytes each)
constexpr unsigned long long kOrdMask = 0xFFFFFFFFFFFFull;  // 48-bit candidate-order field
constexpr uint32_t kNoSplit = 0xFFFFFFFFu;                  // s_split entry of a node that is not split this round (split lines are < 0xFFFF)

struct __align__(16) QNode {
    unsigned short x0, x1, y0, y1;
    unsigned short xm, ym;   // split lines: x0 + ceil((x1 - x0) / 2), likewise y (src/ORBextractor.cc:491-492)
    int count;
};
__device__ __forceinline__ void set_split(QNode& n) {
    n.xm = (unsigned short)(n.x0 + ((n.x1 - n.x0 + 1) >> 1));
    n.ym = (unsigned short)(n.y0 + ((n.y1 - n.y0 + 1) >> 1));
}

// Candidate order of the reference's vToDistributeKeys: cells row-major, raster inside a cell
// (src/ORBextractor.cc:811-848). A cell detects x in [j*wCell+3, (j+1)*wCell+3) (minBorder-relative),
// the last cell of a row/column absorbing the clamped remainder, so (x, y) identifies the cell.
__device__ __forceinline__ unsigned long long order_key(const LevelGeom& L, int x, int y, int lastI, int lastJ) {
    const int j = min((int)(((unsigned)(x - 3) * (unsigned)L.rcpW) >> 20), lastJ), i = min((int)(((unsigned)(y - 3) * (unsigned)L.rcpH) >> 20), lastI);
    const unsigned long long ord = ((((unsigned long long)i << 12 | (unsigned long long)j) << 12 | (unsigned long long)y) << 12) |
                                   (unsigned long long)x;
    return ord;  // < 2^48
}

// Exclusive prefix COUNT of a predicate over i in [0, n): one element per thread and trip, ONE barrier per trip and no shuffle chain
// (the rounds below are chains of barriers and dependent shared-memory round trips, not work): every warp ballots, lane 0 publishes
// the warp's count, and after the barrier every thread adds up the counts of the warps before its own. emit(i, pos, flag) runs for
// every i < n. s_cnt is a double buffer ([2][T / 32]) whose parity the caller carries from call to call. Returns the total.
template <int T, class F, class E>
__device__ __forceinline__ int block_flag_scan(int n, F flag, E emit, int (*s_cnt)[T / 32], int& par) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int run = 0;
    for (int i0 = 0; i0 < n; i0 += T) {
        const int i = i0 + tid;
        const bool f = i < n && flag(i);
        const unsigned b = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_cnt[par][wid] = __popc(b);
        __syncthreads();
        int pre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < T / 32; w++) {
            const int c = s_cnt[par][w];
            tot += c;
            if (w < wid) pre += c;
        }
        par ^= 1;
        if (i < n) emit(i, run + pre + __popc(b & ((1u << lane) - 1u)), f);
        run += tot;
    }
    return run;
}

extern __shared__ __align__(16) unsigned char s_dyn_raw[];

// T = CTA size (compile-time: the scans' warp loops unroll); see COEB_SEL_THREADS above for the two sizes in use.
template <int T, int kMinBlocks>
__global__ void __launch_bounds__(T, kMinBlocks) select_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v, const int level_lo) {
    const int level = level_lo + blockIdx.y, frame = blockIdx.x;   // level-major launch order: the long CTAs (level 0) start first, the short ones fill the tail
    COEB_TRACE(v, level_lo > 0 ? 6 : 5);
    const LevelGeom& L = g.lv[level];
    const DynState& dyn = v.dyn[frame];
    const int tid = threadIdx.x;
    const int LC = g.max_nodes;

    // ---- shared memory carve-up -----------------------------------------------------------------
    unsigned char* sp = s_dyn_raw;
    unsigned long long* s_best = reinterpret_cast<unsigned long long*>(sp); sp += (sizeof(unsigned long long) * LC + 15) & ~(size_t)15;   // QNode is 16-byte aligned
    QNode* s_nodeA = reinterpret_cast<QNode*>(sp); sp += sizeof(QNode) * LC;
    QNode* s_nodeB = reinterpret_cast<QNode*>(sp); sp += sizeof(QNode) * LC;
    int* s_cc = reinterpret_cast<int*>(sp); sp += sizeof(int) * 4 * LC;      // tentative child counts [node id][4]
    int* s_scanA = reinterpret_cast<int*>(sp); sp += sizeof(int) * 4 * LC;   // scratch for scans (up to 4*LC entries)
    unsigned short* s_newpos = reinterpret_cast<unsigned short*>(sp); sp += sizeof(unsigned short) * 4 * LC;   // [node id][quadrant] -> list position after the round (8-byte aligned rows)
    int* s_scanB = reinterpret_cast<int*>(sp); sp += sizeof(int) * LC;
    uint32_t* s_split = reinterpret_cast<uint32_t*>(sp); sp += sizeof(uint32_t) * LC;   // node id -> xm | ym << 16 if the node is split this round, else kNoSplit
    unsigned short* s_slot = reinterpret_cast<unsigned short*>(sp); sp += sizeof(unsigned short) * LC;
    unsigned short* s_P = reinterpret_cast<unsigned short*>(sp); sp += sizeof(unsigned short) * LC;
    unsigned short* s_E = reinterpret_cast<unsigned short*>(sp); sp += sizeof(unsigned short) * LC;
    unsigned short* s_E2 = reinterpret_cast<unsigned short*>(sp); sp += sizeof(unsigned short) * LC;
    sp = s_dyn_raw + ((sp - s_dyn_raw + 15) & ~(size_t)15);
    uint32_t* s_keys = reinterpret_cast<uint32_t*>(sp); sp += sizeof(uint32_t) * kKeyCache;          // candidate cache
    unsigned short* s_knode = reinterpret_cast<unsigned short*>(sp); sp += sizeof(unsigned short) * kKeyCache;
    __shared__ int s_warp[33];
    __shared__ int s_cnt[2][T / 32];   // block_flag_scan's double buffer
    int par = 0;
    __shared__ int s_misc[8];

    int* key_count = v.key_count + frame * g.nlevels + level;
    const int nl = v.lmax_count[frame * g.nlevels + level];
    if (nl > L.cand_cap) {  // cannot happen (capacity is the NMS bound); fail loudly rather than truncate
        if (tid == 0) { *key_count = 0; v.cand_count[frame * g.nlevels + level] = 0; atomicMin(&v.status[frame], (int)COEB_ERR_CAPACITY); }
        return;
    }
    // ---- gather: apply each FAST cell's threshold (iniTh, or minTh if the cell has nothing above iniTh,
    //      src/ORBextractor.cc:831-838) and, on the area_flag path, CheckMovingKeyPoints (:854-858) ----
    // Candidates and their node ids live in shared memory when the level's local maxima fit the cache (the usual
    // case); otherwise the passes below run on the global arrays. The candidate list is always written to global too.
    uint32_t* gkeys = v.cand + (size_t)frame * g.cand_per_frame + L.cand_base;
    const bool cached = nl <= kKeyCache;
    uint32_t* keys = cached ? s_keys : gkeys;
    unsigned short* knode = cached ? s_knode : v.knode + (size_t)frame * g.cand_per_frame + L.cand_base;
    const int lastJ = L.lastJ, lastI = L.lastI;
    {
        const uint32_t* __restrict__ lm = v.lmax + (size_t)frame * g.cand_per_frame + L.cand_base;
        const int* __restrict__ cellcnt = v.cell_count + (size_t)frame * g.cells_per_frame + L.cell_base;
        const int thIni = dyn.area_flag ? 30 : 20, thMin = dyn.area_flag ? 10 : 7;   // :775-784
        if (tid == 0) s_misc[1] = 0;
        __syncthreads();
        // Two dependent global loads per candidate (the candidate, then its cell's counter): eight candidates per thread are in
        // flight together, so a level costs a few round trips instead of one per candidate and thread.
        constexpr int kU = COEB_SEL_GU;
        for (int k0 = tid; k0 < nl; k0 += kU * T) {
            uint32_t key[kU];
            int cnt[kU];
#pragma unroll
            for (int u = 0; u < kU; u++) key[u] = k0 + u * T < nl ? __ldg(&lm[k0 + u * T]) : 0u;
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const int x = key[u] & 0xFFF, y = (key[u] >> 12) & 0xFFF;
                // cell of the candidate, as in fast.cu cell_of (exact multiply-shift quotient, checked on the host)
                const int j = min((int)(((unsigned)(x - 3) * (unsigned)L.rcpW) >> 20), lastJ), i = min((int)(((unsigned)(y - 3) * (unsigned)L.rcpH) >> 20), lastI);
                cnt[u] = k0 + u * T < nl ? cellcnt[i * L.nCols + j] : 0;
            }
#pragma unroll
            for (int u = 0; u < kU; u++) {
                if (k0 + u * T >= nl) break;
                const int x = key[u] & 0xFFF, y = (key[u] >> 12) & 0xFFF, A = (int)(key[u] >> 24) + 1;
                const int th = cnt[u] > 0 ? thIni : thMin;
                if (A > th && !(dyn.area_flag && is_moving(dyn, (float)x, (float)y, level, L.scale, g.w0, g.h0))) {
                    const int pos = atomicAdd(&s_misc[1], 1);
                    gkeys[pos] = key[u];
                    if (cached) s_keys[pos] = key[u];
                }
            }
        }
        __syncthreads();
    }
    const int nkeys = s_misc[1];
    COEB_TRACE_MARK(v, level == 0 && frame == 0, 0);   // gathered
    if (tid == 0) v.cand_count[frame * g.nlevels + level] = nkeys;
    if (nkeys == 0 || dyn.bad_box) {
        if (tid == 0) *key_count = 0;
        return;
    }

    int N = L.n_target;
    if (dyn.area_flag) N = (int)((double)N * 0.7);  // (int)(mnFeaturesPerLevel[level])*0.7 -> const int& (:869)
    const int H = L.maxBY - kMinBorder;
    const int nIni = L.n_ini;
    const float hX = L.hX;

    // ---- roots (:550-592) ------------------------------------------------------------------------
    QNode* cur = s_nodeA;
    QNode* nxt = s_nodeB;
    for (int i = tid; i < nIni; i += T) {
        QNode n;
        n.x0 = (unsigned short)(int)__fmul_rn(hX, (float)i);
        n.x1 = (unsigned short)(int)__fmul_rn(hX, (float)(i + 1));
        n.y0 = 0;
        n.y1 = (unsigned short)H;
        set_split(n);
        n.count = 0;
        nxt[i] = n;
    }
    __syncthreads();
    if (nIni == 1) {   // one root (every image that is not wider than 3:2): all keys are its keys, no counting
        for (int k = tid; k < nkeys; k += T) knode[k] = 0;
        if (tid == 0) nxt[0].count = nkeys;
    } else {
        for (int k = tid; k < nkeys; k += T) {
            const int x = keys[k] & 0xFFF;
            int r = (int)__fdiv_rn((float)x, hX);
            r = min(max(r, 0), nIni - 1);
            knode[k] = (unsigned short)r;
            atomicAdd(&nxt[r].count, 1);
        }
    }
    __syncthreads();
    for (int i = tid; i < nIni; i += T) s_scanA[i] = nxt[i].count > 0;
    __syncthreads();
    int nList = block_exclusive_scan<T>(s_scanA, nIni, s_warp);
    for (int i = tid; i < nIni; i += T)
        if (nxt[i].count > 0) cur[s_scanA[i]] = nxt[i];
    __syncthreads();
    if (nList != nIni) {
        for (int k = tid; k < nkeys; k += T) knode[k] = (unsigned short)s_scanA[knode[k]];
    }
    __syncthreads();

    COEB_TRACE_MARK(v, level == 0 && frame == 0, 1);   // roots
    // ---- rounds (:601-745) -------------------------------------------------------------------------
    bool careful = false;
    int nE = 0;  // expandable nodes created in the previous round, creation order, in s_E
    for (int round = 0; round < 64; round++) {
        const int prevSize = nList;
        int nP;
        if (!careful) {
            // full pass: every multi-key node, in list order (:613-672)
            nP = block_flag_scan<T>(
                nList, [&](int i) { return cur[i].count > 1; },
                [&](int i, int pos, bool f) {
                    if (f) {
                        s_P[pos] = (unsigned short)i; s_slot[i] = (unsigned short)pos;
                        s_split[i] = reinterpret_cast<const uint32_t*>(&cur[i])[2];   // xm | ym << 16
                        *reinterpret_cast<int4*>(&s_cc[4 * i]) = make_int4(0, 0, 0, 0);   // tentative child counts, zeroed in the same phase
                    } else {
                        s_slot[i] = 0xFFFF; s_split[i] = kNoSplit;
                    }
                },
                s_cnt, par);
        } else {
            // careful phase: previous round's multi-key children, largest first, later-created first on ties (:688-692)
            nP = nE;
            for (int i = tid; i < nList; i += T) { s_slot[i] = 0xFFFF; s_split[i] = kNoSplit; }
            __syncthreads();
            for (int e = tid; e < nE; e += T) {
                const int nd = s_E[e];
                const int ce = cur[nd].count;
                int rank = 0;
                for (int f = 0; f < nE; f++) {
                    const int cf = cur[s_E[f]].count;
                    rank += (cf > ce) || (cf == ce && f > e);
                }
                s_P[rank] = (unsigned short)nd;
                s_slot[nd] = (unsigned short)rank;
                s_split[nd] = reinterpret_cast<const uint32_t*>(&cur[nd])[2];
                *reinterpret_cast<int4*>(&s_cc[4 * nd]) = make_int4(0, 0, 0, 0);
            }
        }
        __syncthreads();
        if (nP == 0) break;  // nothing left to split: list size cannot change (:676)

        // tentative children of every node in P: node id -> split lines (or "not split") -> quadrant; a few keys per thread are in
        // flight together so that the chains of dependent shared-memory loads overlap
        constexpr int kKU = COEB_SEL_KU;
        for (int k0 = tid; k0 < nkeys; k0 += kKU * T) {
            int nd[kKU];
            uint32_t key[kKU], xy[kKU];
#pragma unroll
            for (int u = 0; u < kKU; u++) nd[u] = k0 + u * T < nkeys ? (int)knode[k0 + u * T] : -1;
#pragma unroll
            for (int u = 0; u < kKU; u++) {
                xy[u] = nd[u] >= 0 ? s_split[nd[u]] : kNoSplit;
                key[u] = nd[u] >= 0 ? keys[k0 + u * T] : 0u;
            }
#pragma unroll
            for (int u = 0; u < kKU; u++) {
                if (xy[u] != kNoSplit) {
                    const int x = key[u] & 0xFFF, y = (key[u] >> 12) & 0xFFF;
                    const int q = (x < (int)(xy[u] & 0xFFFFu) ? 0 : 1) + (y < (int)(xy[u] >> 16) ? 0 : 2);
                    atomicAdd(&s_cc[4 * nd[u] + q], 1);
                    knode[k0 + u * T] = (unsigned short)(nd[u] | (q << 14));   // the quadrant rides in the two top bits until the keys move below
                }
            }
        }
        __syncthreads();

        // how many nodes get split this round
        int nProc = nP;
        if (careful) {
            for (int p = tid; p < nP; p += T) {
                const int* c = &s_cc[4 * s_P[p]];
                s_scanA[p] = (c[0] > 0) + (c[1] > 0) + (c[2] > 0) + (c[3] > 0) - 1;
            }
            __syncthreads();
            block_exclusive_scan<T>(s_scanA, nP, s_warp);
            if (tid == 0) s_misc[0] = nP;
            __syncthreads();
            for (int p = tid; p < nP; p += T) {
                const int* c = &s_cc[4 * s_P[p]];
                const int after = nList + s_scanA[p] + (c[0] > 0) + (c[1] > 0) + (c[2] > 0) + (c[3] > 0) - 1;
                if (after >= N) atomicMin(&s_misc[0], p + 1);  // break after the first split reaching N (:737-738)
            }
            __syncthreads();
            nProc = s_misc[0];
            __syncthreads();
        }

        // Creation sequence: for p in processing order, children q = 0..3 that hold keys; the multi-key ones among them are next
        // round's expandable list, in creation order; the surviving old nodes keep their relative order behind the new ones. All
        // three are prefix counts of flags, taken together with ballots (thread t of a trip owns processed node t and list entry t)
        // and one barrier per trip: child (p, q)'s creation index is the children of the nodes before p plus p's own earlier ones.
        const int lane = tid & 31, wid = tid >> 5;
        const unsigned lt = (1u << lane) - 1u;
        int totalNew = 0, nE2 = 0, nKeep = 0;
        for (int i0 = 0; i0 < max(nProc, nList); i0 += T) {
            const int p = i0 + tid;
            int4 cc = make_int4(0, 0, 0, 0);
            if (p < nProc) cc = *reinterpret_cast<const int4*>(&s_cc[4 * s_P[p]]);
            bool keep = false;
            if (p < nList) { const int sl = s_slot[p]; keep = !(sl != 0xFFFF && sl < nProc); }
            const unsigned b0 = __ballot_sync(0xffffffffu, cc.x > 0), b1 = __ballot_sync(0xffffffffu, cc.y > 0);
            const unsigned b2 = __ballot_sync(0xffffffffu, cc.z > 0), b3 = __ballot_sync(0xffffffffu, cc.w > 0);
            const unsigned e0 = __ballot_sync(0xffffffffu, cc.x > 1), e1 = __ballot_sync(0xffffffffu, cc.y > 1);
            const unsigned e2 = __ballot_sync(0xffffffffu, cc.z > 1), e3 = __ballot_sync(0xffffffffu, cc.w > 1);
            const unsigned bk = __ballot_sync(0xffffffffu, keep);
            // children | expandable << 11 | kept << 22: the CTA-wide sums stay inside their fields (<= 4 T, <= 4 T, <= T with T <= 512),
            // so the packed words are added as they are
            static_assert(T <= 512, "field widths of the packed prefix counts");
            if (lane == 0)
                s_cnt[par][wid] = (__popc(b0) + __popc(b1) + __popc(b2) + __popc(b3)) | ((__popc(e0) + __popc(e1) + __popc(e2) + __popc(e3)) << 11) | (__popc(bk) << 22);
            __syncthreads();
            unsigned pre = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < T / 32; w++) {
                const unsigned c = (unsigned)s_cnt[par][w];
                tot += c;
                if (w < wid) pre += c;
            }
            const int preC = pre & 0x7FF, preE = (pre >> 11) & 0x7FF, preK = pre >> 22;
            const int totC = tot & 0x7FF, totE = (tot >> 11) & 0x7FF, totK = tot >> 22;
            par ^= 1;
            if (p < nProc)
                s_scanA[p] = (totalNew + preC + __popc(b0 & lt) + __popc(b1 & lt) + __popc(b2 & lt) + __popc(b3 & lt)) |
                             ((nE2 + preE + __popc(e0 & lt) + __popc(e1 & lt) + __popc(e2 & lt) + __popc(e3 & lt)) << 16);
            if (p < nList && keep) s_scanB[p] = nKeep + preK + __popc(bk & lt);
            totalNew += totC; nE2 += totE; nKeep += totK;
        }
        const int newSize = totalNew + nKeep;
        if (newSize > LC) {  // cannot happen for max_nodes >= max(N + 3, 4 * nIni); fail loudly
            if (tid == 0) { *key_count = 0; atomicMin(&v.status[frame], (int)COEB_ERR_CAPACITY); }
            return;
        }
        // same thread <-> same p as above: what it reads from s_scanA / s_scanB it wrote itself
        for (int i0 = 0; i0 < max(nProc, nList); i0 += T) {
            const int p = i0 + tid;
            if (p < nList) {
                const int sl = s_slot[p];
                if (!(sl != 0xFFFF && sl < nProc)) {
                    const uint32_t pos = (uint32_t)(totalNew + s_scanB[p]);
                    // all four quadrants: the keys of a node that was a candidate of the careful phase but is not split after all
                    // already carry a quadrant
                    *reinterpret_cast<uint2*>(&s_newpos[4 * p]) = make_uint2(pos | (pos << 16), pos | (pos << 16));
                    nxt[pos] = cur[p];
                }
            }
            if (p < nProc) {
                const int ndp = s_P[p];
                const int4 cc = *reinterpret_cast<const int4*>(&s_cc[4 * ndp]);
                const int cnts[4] = {cc.x, cc.y, cc.z, cc.w};
                int ci = s_scanA[p] & 0xFFFF, ei = s_scanA[p] >> 16;
                const QNode n = cur[ndp];
                const int xm = n.xm, ym = n.ym;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int cnt = cnts[q];
                    if (cnt > 0) {
                        QNode c;
                        c.x0 = (q & 1) ? xm : n.x0;
                        c.x1 = (q & 1) ? n.x1 : xm;
                        c.y0 = (q & 2) ? ym : n.y0;
                        c.y1 = (q & 2) ? n.y1 : ym;
                        set_split(c);
                        c.count = cnt;
                        const int pos = totalNew - 1 - ci;  // push_front in creation order
                        nxt[pos] = c;
                        s_newpos[4 * ndp + q] = (unsigned short)pos;
                        if (cnt > 1) s_E2[ei++] = (unsigned short)pos;   // next round's expandable list
                        ci++;
                    }
                }
            }
        }
        __syncthreads();

        // move the keys: (node id, quadrant) -> position in the new list, one table look-up
        for (int k0 = tid; k0 < nkeys; k0 += kKU * T) {
            int kn[kKU], to[kKU];
#pragma unroll
            for (int u = 0; u < kKU; u++) kn[u] = k0 + u * T < nkeys ? (int)knode[k0 + u * T] : -1;
#pragma unroll
            for (int u = 0; u < kKU; u++) to[u] = kn[u] >= 0 ? (int)s_newpos[((kn[u] & 0x3FFF) << 2) | (kn[u] >> 14)] : 0;
#pragma unroll
            for (int u = 0; u < kKU; u++)
                if (kn[u] >= 0) knode[k0 + u * T] = (unsigned short)to[u];
        }
        __syncthreads();
        { QNode* t = cur; cur = nxt; nxt = t; }
        { unsigned short* t = s_E; s_E = s_E2; s_E2 = t; }
        nE = nE2;
        nList = newSize;
#ifdef COEB_KERNEL_TRACE
        if (v.trace && level == 0 && frame == 0 && tid == 0) v.trace[31] = (unsigned long long)(round + 1);   // rounds done
        if (round < 3) COEB_TRACE_MARK(v, level == 0 && frame == 0, 4 + round);
#endif
        if (nList >= N || nList == prevSize) break;               // :676, :741
        if (!careful && nList + 3 * nE > N) careful = true;        // :680
    }

    COEB_TRACE_MARK(v, level == 0 && frame == 0, 2);   // rounds
    // ---- best key per node (:748-766) -------------------------------------------------------------
    for (int i = tid; i < nList; i += T) s_best[i] = 0ull;
    __syncthreads();
    {
        for (int k = tid; k < nkeys; k += T) {
            const uint32_t key = keys[k];
            const int x = key & 0xFFF, y = (key >> 12) & 0xFFF;
            const unsigned long long ord = order_key(L, x, y, lastI, lastJ);
            const unsigned long long pri = ((unsigned long long)(key >> 24) << 48) | (kOrdMask - ord);
            atomicMax(&s_best[knode[k]], pri);
        }
    }
    __syncthreads();

    COEB_TRACE_MARK(v, level == 0 && frame == 0, 3);   // best key
    // ---- fix-up (:877-890), orientation (:902-903), post cull (:1204-1207), ordered compaction -----
    // s_scanA: keep flag / position; keys are written level-relative + 16.
    // orientation (:902-903) is computed by describe_kernel, which has the keypoint-parallel shape for it
    LevelKey* out = v.keys + (size_t)frame * g.keys_per_frame + L.key_base;
    const int nOut = block_flag_scan<T>(
        nList,
        [&](int i) {
            const unsigned long long ord = kOrdMask - (s_best[i] & kOrdMask);
            const int x = (int)(ord & 0xFFF), y = (int)((ord >> 12) & 0xFFF);
            return !(!dyn.area_flag && level < 8 && is_moving(dyn, (float)(x + kMinBorder), (float)(y + kMinBorder), level, L.scale, g.w0, g.h0));
        },
        [&](int i, int pos, bool keep) {
            if (!keep || pos >= L.key_cap) return;   // beyond the capacity: reported below, nothing is written out of bounds
            const unsigned long long best = s_best[i];
            const unsigned long long ord = kOrdMask - (best & kOrdMask);
            LevelKey k;
            k.x = (float)((int)(ord & 0xFFF) + kMinBorder);
            k.y = (float)((int)((ord >> 12) & 0xFFF) + kMinBorder);
            k.response = (float)(int)(best >> 48);
            k.angle = 0.f;
            out[pos] = k;
        },
        s_cnt, par);
    if (nOut > L.key_cap) {
        if (tid == 0) { *key_count = 0; atomicMin(&v.status[frame], (int)COEB_ERR_CAPACITY); }
        return;
    }
    if (tid == 0) *key_count = nOut;
}

size_t select_smem_bytes(int LC) {
    return (size_t)LC * (sizeof(unsigned long long) + 2 * sizeof(QNode) + 4 * 4 + 4 * 4 + 4 * 2 + 4 + 4 + 2 * 4) + 48 + (size_t)kKeyCache * 6;
}

void launch_select(const Geometry& g, const BatchView& v, cudaStream_t stream, int level_lo, int level_hi) {
    if (level_hi < 0) level_hi = g.nlevels;
    if (level_hi <= level_lo) return;
    const size_t smem = select_smem_bytes(g.max_nodes);
    // the opt-in shared-memory size is a per-device function attribute: one handle per GPU may live in the same process
    static size_t configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& done = configured[dev & 63];
    if (smem > done) {
        cudaFuncSetAttribute(select_kernel<kSelThreads, COEB_SEL_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(select_kernel<COEB_SEL_THREADS_SMALL, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        done = smem;
    }
    if (v.B <= 4) select_kernel<COEB_SEL_THREADS_SMALL, 3><<<dim3(v.B, level_hi - level_lo), COEB_SEL_THREADS_SMALL, smem, stream>>>(g, v, level_lo);
    else select_kernel<kSelThreads, COEB_SEL_MINB><<<dim3(v.B, level_hi - level_lo), kSelThreads, smem, stream>>>(g, v, level_lo);
}

}  // namespace coeb
