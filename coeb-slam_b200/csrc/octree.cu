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
//     (children are push_front'ed as created, :630-667 / :698-733), which is two prefix sums;
//   * the careful phase's sort of (size, pointer) pairs (:691) uses the creation index as the second
//     key (documented tie-break, identical to the CPU oracle) and is done by rank counting;
//   * the per-node "best response, first key wins" (:748-766) is a 64-bit atomicMax on
//     (response, reversed candidate order), the candidate order being re-derived from (x, y).
#include "coeb_device.cuh"

namespace coeb {

#ifndef COEB_SEL_THREADS
#define COEB_SEL_THREADS 384
#endif
#ifndef COEB_SEL_MINB
#define COEB_SEL_MINB 3
#endif
constexpr int kSelThreads = COEB_SEL_THREADS;
constexpr int kKeyCache = 4096;   // candidates per (level, frame) kept in shared memory (4 + 2 bytes each)
constexpr unsigned long long kOrdMask = 0xFFFFFFFFFFFFull;  // 48-bit candidate-order field

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

// Two independent exclusive prefix sums sharing their three barriers: a[i] = sum of fa(j), j < i, for i in [0, na), likewise b / fb.
// The inputs are computed on the fly (fa / fb read data that is stable since the last barrier), so no flag pass and no barrier
// precede the scan. Returns a's total, b's in *tb.
template <int T, class FA, class FB>
__device__ __forceinline__ int block_exclusive_scan2(int* a, int na, FA fa, int* b, int nb, FB fb, int* s_wa, int* s_wb, int* tb) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int ca = (na + T - 1) / T, cb = (nb + T - 1) / T;
    const int loa = min(tid * ca, na), hia = min(loa + ca, na), lob = min(tid * cb, nb), hib = min(lob + cb, nb);
    int sa = 0, sb = 0;
    for (int i = loa; i < hia; i++) { const int t = fa(i); a[i] = t; sa += t; }
    for (int i = lob; i < hib; i++) { const int t = fb(i); b[i] = t; sb += t; }
    int ia = sa, ib = sb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int ta = __shfl_up_sync(0xffffffffu, ia, o), t2 = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += ta; ib += t2; }
    }
    if (lane == 31) { s_wa[wid] = ia; s_wb[wid] = ib; }
    __syncthreads();
    if (wid == 0) {
        const int va = lane < (T >> 5) ? s_wa[lane] : 0, vb = lane < (T >> 5) ? s_wb[lane] : 0;
        int wa = va, wb = vb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int ta = __shfl_up_sync(0xffffffffu, wa, o), t2 = __shfl_up_sync(0xffffffffu, wb, o);
            if (lane >= o) { wa += ta; wb += t2; }
        }
        s_wa[lane] = wa - va; s_wb[lane] = wb - vb;
        if (lane == 31) { s_wa[32] = wa; s_wb[32] = wb; }
    }
    __syncthreads();
    int ra = s_wa[wid] + ia - sa, rb = s_wb[wid] + ib - sb;
    for (int i = loa; i < hia; i++) { const int t = a[i]; a[i] = ra; ra += t; }
    for (int i = lob; i < hib; i++) { const int t = b[i]; b[i] = rb; rb += t; }
    const int total = s_wa[32];
    *tb = s_wb[32];
    __syncthreads();
    return total;
}

extern __shared__ __align__(16) unsigned char s_dyn_raw[];

// T = CTA size (compile-time, so that the scans' chunking is a multiply-shift). For a single frame a level is one CTA and the longest
// one (level 0) sets the latency: 384, 512, 768 and 1024 threads were measured there and make no difference (151.5-152.3 us per
// call): the passes are chains of barriers and shared-memory round trips, not thread-count bound.
template <int T, int kMinBlocks>
__global__ void __launch_bounds__(T, kMinBlocks) select_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v) {
    const int level = blockIdx.y, frame = blockIdx.x;   // level-major launch order: the long CTAs (level 0) start first, the short ones fill the tail
    const LevelGeom& L = g.lv[level];
    const DynState& dyn = v.dyn[frame];
    const int tid = threadIdx.x;
    const int LC = g.max_nodes;

    // ---- shared memory carve-up -----------------------------------------------------------------
    unsigned char* sp = s_dyn_raw;
    unsigned long long* s_best = reinterpret_cast<unsigned long long*>(sp); sp += (sizeof(unsigned long long) * LC + 15) & ~(size_t)15;   // QNode is 16-byte aligned
    QNode* s_nodeA = reinterpret_cast<QNode*>(sp); sp += sizeof(QNode) * LC;
    QNode* s_nodeB = reinterpret_cast<QNode*>(sp); sp += sizeof(QNode) * LC;
    int* s_cc = reinterpret_cast<int*>(sp); sp += sizeof(int) * 4 * LC;      // tentative child counts [slot][4]
    int* s_scanA = reinterpret_cast<int*>(sp); sp += sizeof(int) * 4 * LC;   // scratch for scans (up to 4*LC entries)
    int* s_scanB = reinterpret_cast<int*>(sp); sp += sizeof(int) * LC;
    unsigned short* s_childpos = reinterpret_cast<unsigned short*>(sp); sp += sizeof(unsigned short) * 4 * LC;
    unsigned short* s_oldpos = reinterpret_cast<unsigned short*>(sp); sp += sizeof(unsigned short) * LC;
    unsigned short* s_slot = reinterpret_cast<unsigned short*>(sp); sp += sizeof(unsigned short) * LC;
    unsigned short* s_P = reinterpret_cast<unsigned short*>(sp); sp += sizeof(unsigned short) * LC;
    unsigned short* s_E = reinterpret_cast<unsigned short*>(sp); sp += sizeof(unsigned short) * LC;
    unsigned short* s_E2 = reinterpret_cast<unsigned short*>(sp); sp += sizeof(unsigned short) * LC;
    sp = s_dyn_raw + ((sp - s_dyn_raw + 15) & ~(size_t)15);
    uint32_t* s_keys = reinterpret_cast<uint32_t*>(sp); sp += sizeof(uint32_t) * kKeyCache;          // candidate cache
    unsigned short* s_knode = reinterpret_cast<unsigned short*>(sp); sp += sizeof(unsigned short) * kKeyCache;
    __shared__ int s_warp[33], s_warp2[33];
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
        for (int k = tid; k < nl; k += T) {
            const uint32_t key = lm[k];
            const int x = key & 0xFFF, y = (key >> 12) & 0xFFF, A = (int)(key >> 24) + 1;
            // cell of the candidate, as in fast.cu cell_of (exact multiply-shift quotient, checked on the host)
            const int j = min((int)(((unsigned)(x - 3) * (unsigned)L.rcpW) >> 20), lastJ), i = min((int)(((unsigned)(y - 3) * (unsigned)L.rcpH) >> 20), lastI);
            const int th = cellcnt[i * L.nCols + j] > 0 ? thIni : thMin;
            if (A > th && !(dyn.area_flag && is_moving(dyn, (float)x, (float)y, level, L.scale, g.w0, g.h0)))
            {
                const int pos = atomicAdd(&s_misc[1], 1);
                gkeys[pos] = key;
                if (cached) s_keys[pos] = key;
            }
        }
        __syncthreads();
    }
    const int nkeys = s_misc[1];
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
    for (int k = tid; k < nkeys; k += T) {
        const int x = keys[k] & 0xFFF;
        int r = (int)__fdiv_rn((float)x, hX);
        r = min(max(r, 0), nIni - 1);
        knode[k] = (unsigned short)r;
        atomicAdd(&nxt[r].count, 1);
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

    // ---- rounds (:601-745) -------------------------------------------------------------------------
    bool careful = false;
    int nE = 0;  // expandable nodes created in the previous round, creation order, in s_E
    for (int round = 0; round < 64; round++) {
        const int prevSize = nList;
        int nP;
        if (!careful) {
            // full pass: every multi-key node, in list order (:613-672)
            int unused;
            nP = block_exclusive_scan2<T>(s_scanA, nList, [&](int i) { return (int)(cur[i].count > 1); }, s_scanB, 0, [](int) { return 0; }, s_warp, s_warp2, &unused);
            for (int i = tid; i < nList; i += T) {
                if (cur[i].count > 1) { s_P[s_scanA[i]] = (unsigned short)i; s_slot[i] = (unsigned short)s_scanA[i]; }
                else s_slot[i] = 0xFFFF;
            }
            for (int i = tid; i < 4 * nP; i += T) s_cc[i] = 0;   // tentative child counts, zeroed in the same phase
        } else {
            // careful phase: previous round's multi-key children, largest first, later-created first on ties (:688-692)
            nP = nE;
            for (int i = tid; i < nList; i += T) s_slot[i] = 0xFFFF;
            __syncthreads();
            for (int e = tid; e < nE; e += T) {
                const int ce = cur[s_E[e]].count;
                int rank = 0;
                for (int f = 0; f < nE; f++) {
                    const int cf = cur[s_E[f]].count;
                    rank += (cf > ce) || (cf == ce && f > e);
                }
                s_P[rank] = s_E[e];
                s_slot[s_E[e]] = (unsigned short)rank;
            }
            for (int i = tid; i < 4 * nP; i += T) s_cc[i] = 0;
        }
        __syncthreads();
        if (nP == 0) break;  // nothing left to split: list size cannot change (:676)

        // tentative children of every node in P
        for (int k = tid; k < nkeys; k += T) {
            const int nd = knode[k];
            const int p = s_slot[nd];
            if (p != 0xFFFF) {
                const QNode n = cur[nd];
                const int xm = n.xm, ym = n.ym;
                const uint32_t key = keys[k];
                const int x = key & 0xFFF, y = (key >> 12) & 0xFFF;
                const int q = (x < xm ? 0 : 1) + (y < ym ? 0 : 2);
                atomicAdd(&s_cc[4 * p + q], 1);
                knode[k] = (unsigned short)(nd | (q << 14));   // the quadrant rides in the two top bits until the keys move below
            }
        }
        __syncthreads();

        // how many nodes get split this round
        int nProc = nP;
        if (careful) {
            for (int p = tid; p < nP; p += T) {
                const int* c = &s_cc[4 * p];
                s_scanA[p] = (c[0] > 0) + (c[1] > 0) + (c[2] > 0) + (c[3] > 0) - 1;
            }
            __syncthreads();
            block_exclusive_scan<T>(s_scanA, nP, s_warp);
            if (tid == 0) s_misc[0] = nP;
            __syncthreads();
            for (int p = tid; p < nP; p += T) {
                const int* c = &s_cc[4 * p];
                const int after = nList + s_scanA[p] + (c[0] > 0) + (c[1] > 0) + (c[2] > 0) + (c[3] > 0) - 1;
                if (after >= N) atomicMin(&s_misc[0], p + 1);  // break after the first split reaching N (:737-738)
            }
            __syncthreads();
            nProc = s_misc[0];
            __syncthreads();
        }

        // creation sequence: for p in processing order, children q = 0..3 that hold keys. One packed scan gives both the creation
        // index of every child (low half) and its rank among the multi-key children (high half): the latter are next round's
        // expandable list, in creation order.
        // (second scan) surviving old nodes keep their relative order behind the new ones
        int nKeep;
        const int totals = block_exclusive_scan2<T>(
            s_scanA, 4 * nProc, [&](int i) { const int c = s_cc[i]; return (c > 0) | ((c > 1) << 16); },
            s_scanB, nList, [&](int i) { const int p = s_slot[i]; return (int)!(p != 0xFFFF && p < nProc); }, s_warp, s_warp2, &nKeep);
        const int totalNew = totals & 0xFFFF, nE2 = totals >> 16;
        const int newSize = totalNew + nKeep;
        if (newSize > LC) {  // cannot happen for max_nodes >= max(N + 3, 4 * nIni); fail loudly
            if (tid == 0) { *key_count = 0; atomicMin(&v.status[frame], (int)COEB_ERR_CAPACITY); }
            return;
        }
        for (int i = tid; i < nList; i += T) {
            const int p = s_slot[i];
            if (!(p != 0xFFFF && p < nProc)) {
                const int pos = totalNew + s_scanB[i];
                s_oldpos[i] = (unsigned short)pos;
                nxt[pos] = cur[i];
            }
        }
        for (int i = tid; i < 4 * nProc; i += T) {
            const int cnt = s_cc[i];
            if (cnt > 0) {
                const int p = i >> 2, q = i & 3;
                const QNode n = cur[s_P[p]];
                const int xm = n.xm, ym = n.ym;
                QNode c;
                c.x0 = (q & 1) ? xm : n.x0;
                c.x1 = (q & 1) ? n.x1 : xm;
                c.y0 = (q & 2) ? ym : n.y0;
                c.y1 = (q & 2) ? n.y1 : ym;
                set_split(c);
                c.count = cnt;
                const int pos = totalNew - 1 - (s_scanA[i] & 0xFFFF);  // push_front in creation order
                nxt[pos] = c;
                s_childpos[i] = (unsigned short)pos;
                if (cnt > 1) s_E2[s_scanA[i] >> 16] = (unsigned short)pos;   // next round's expandable list
            }
        }
        __syncthreads();

        // move the keys
        for (int k = tid; k < nkeys; k += T) {
            const int kn = knode[k];
            const int nd = kn & 0x3FFF, q = kn >> 14;
            const int p = s_slot[nd];
            knode[k] = (p != 0xFFFF && p < nProc) ? s_childpos[4 * p + q] : s_oldpos[nd];
        }
        __syncthreads();
        { QNode* t = cur; cur = nxt; nxt = t; }
        { unsigned short* t = s_E; s_E = s_E2; s_E2 = t; }
        nE = nE2;
        nList = newSize;
        if (nList >= N || nList == prevSize) break;               // :676, :741
        if (!careful && nList + 3 * nE > N) careful = true;        // :680
    }

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

    // ---- fix-up (:877-890), orientation (:902-903), post cull (:1204-1207), ordered compaction -----
    // s_scanA: keep flag / position; keys are written level-relative + 16.
    for (int i = tid; i < nList; i += T) {
        const unsigned long long ord = kOrdMask - (s_best[i] & kOrdMask);
        const int x = (int)(ord & 0xFFF), y = (int)((ord >> 12) & 0xFFF);
        const float fx = (float)(x + kMinBorder), fy = (float)(y + kMinBorder);
        int keep = 1;
        if (!dyn.area_flag && level < 8 && is_moving(dyn, fx, fy, level, L.scale, g.w0, g.h0)) keep = 0;
        s_scanA[i] = keep;
    }
    __syncthreads();
    const int nOut = block_exclusive_scan<T>(s_scanA, nList, s_warp);
    LevelKey* out = v.keys + (size_t)frame * g.keys_per_frame + L.key_base;
    if (nOut > L.key_cap) {
        if (tid == 0) { *key_count = 0; atomicMin(&v.status[frame], (int)COEB_ERR_CAPACITY); }
        return;
    }
    if (tid == 0) *key_count = nOut;

    // orientation (:902-903) is computed by describe_kernel, which has the keypoint-parallel shape for it
    for (int i = tid; i < nList; i += T) {
        const bool keep = (i + 1 < nList ? s_scanA[i + 1] : nOut) != s_scanA[i];
        if (!keep) continue;
        const unsigned long long best = s_best[i];
        const unsigned long long ord = kOrdMask - (best & kOrdMask);
        LevelKey k;
        k.x = (float)((int)(ord & 0xFFF) + kMinBorder);
        k.y = (float)((int)((ord >> 12) & 0xFFF) + kMinBorder);
        k.response = (float)(int)(best >> 48);
        k.angle = 0.f;
        out[s_scanA[i]] = k;
    }
}

size_t select_smem_bytes(int LC) {
    return (size_t)LC * (sizeof(unsigned long long) + 2 * sizeof(QNode) + 4 * 4 + 4 * 4 + 4 + 4 * 2 + 2 * 5) + 32 + (size_t)kKeyCache * 6;
}

void launch_select(const Geometry& g, const BatchView& v, cudaStream_t stream) {
    const size_t smem = select_smem_bytes(g.max_nodes);
    // the opt-in shared-memory size is a per-device function attribute: one handle per GPU may live in the same process
    static size_t configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& done = configured[dev & 63];
    if (smem > done) {
        cudaFuncSetAttribute(select_kernel<kSelThreads, COEB_SEL_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        done = smem;
    }
    select_kernel<kSelThreads, COEB_SEL_MINB><<<dim3(v.B, g.nlevels), kSelThreads, smem, stream>>>(g, v);
}

}  // namespace coeb
