// describe.cu -- IC_Angle orientation (E8), steered rBRIEF descriptors (E11) and output assembly (E12).
//
// Reference: IC_Angle / computeOrientation (src/ORBextractor.cc:80-107, 902-903), computeOrbDescriptor /
// computeDescriptors (:109-156, 1078-1085) and the assembly loop of ORBextractor::operator() (:1291-1337).
// One CTA per (level, frame, 64-keypoint chunk); one warp per keypoint.
//   orientation: lane v+15 owns patch row v. The row's 31 pixels are 9 aligned words funnel-shifted to start at the
//                patch's first column; m10 and the row sum are byte dot products (IDP.4A) with per-lane constant
//                weight words (u inside the circular patch, 0 outside), m01 = v * row sum; a warp reduction and
//                cv::fastAtan2 finish it.
//   descriptor : lane i produces descriptor byte i from pattern pairs 8i..8i+7 on the blurred level.
#include "coeb_device.cuh"

namespace coeb {

__constant__ __align__(4) signed char c_pattern[1024] = {
#include "../../include/coeb_orb_pattern.inc"
};

// cv::fastAtan2 (degrees), scalar fp32 path; explicit rn intrinsics keep the compiler from fusing
// multiply-adds (the CPU build is -ffp-contract=off).
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, scale);
    const float p3 = __fmul_rn(-0.3258083974640975f, scale);
    const float p5 = __fmul_rn(0.1555786518463281f, scale);
    const float p7 = __fmul_rn(-0.04432655554792128f, scale);
    const float ax = fabsf(x), ay = fabsf(y);
    const float eps = (float)2.2204460492503131e-16;
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// cvRound for |x| < 2^22: adding 1.5 * 2^23 rounds to the nearest integer, ties to even, and leaves that integer in the low
// mantissa bits; no F2I (a quarter-rate conversion) on the descriptor's 512 coordinates per keypoint.
__device__ __forceinline__ int round_even(float x) { return __float_as_int(__fadd_rn(x, 12582912.f)) - 0x4B400000; }

#ifndef COEB_DESC_MINB
#define COEB_DESC_MINB 4
#endif
#ifndef COEB_DESC_CHUNK
#define COEB_DESC_CHUNK 64
#endif
constexpr int kDescChunk = COEB_DESC_CHUNK;   // keypoints per CTA: several CTAs per level keep a single frame's latency low

__global__ void __launch_bounds__(256, COEB_DESC_MINB) describe_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v) {
    __shared__ float s_pat[1024];         // pattern as floats, transposed: [4*bit + component][lane], conflict-free per-lane reads
    __shared__ float2 s_cs[256];          // (cos, sin) of the keypoints of the current chunk
    __shared__ float s_angle[256];
    __shared__ int s_ioff[279], s_ipk[279];
    __shared__ uint32_t s_imask[kIcMaskWords];
    const int level = blockIdx.x, frame = blockIdx.y;
    const LevelGeom& L = g.lv[level];
    const int tid = threadIdx.x;

    // row offset of this level in the frame's output = keypoints of the lower levels (:1307-1324)
    const int* kc = v.key_count + frame * g.nlevels;
    int offset = 0, total = 0;
    for (int l = 0; l < g.nlevels; l++) {
        const int c = kc[l];
        if (l < level) offset += c;
        total += c;
    }
    const int n = kc[level];
    const int status = v.status[frame];
    const int chunk0 = blockIdx.z * kDescChunk;   // this CTA describes keypoints [chunk0, chunk0 + kDescChunk) of the level
    if (level == 0 && blockIdx.z == 0 && tid == 0) {
        int cnt = total;
        if (status != COEB_OK) cnt = 0;
        else if (total > g.out_cap) { v.status[frame] = COEB_ERR_CAPACITY; }
        v.out_count[frame] = cnt;
        if (v.mirror_hdr) { v.mirror_hdr[0] = cnt; v.mirror_hdr[1] = status != COEB_OK ? status : (total > g.out_cap ? (int)COEB_ERR_CAPACITY : (int)COEB_OK); }
    }
    if (chunk0 >= n || status != COEB_OK || total > g.out_cap) return;   // uniform; most chunks of the coarse levels are empty
    for (int i = tid; i < 1024; i += 256) s_pat[(i & 31) * 32 + (i >> 5)] = (float)c_pattern[i];   // byte `i>>5` uses ints [32*(i>>5), +32)

    const uint8_t* __restrict__ img = blur_ptr(g, v, level, frame);
    const int pitch = L.pitch;
    const uint8_t* __restrict__ raw = level_ptr(g, v, level, frame);   // IC_Angle runs on the unblurred level
    const int rpitch = level_pitch(g, v, level);
    LevelKey* keys = v.keys + (size_t)frame * g.keys_per_frame + L.key_base;
    coeb_keypoint* okp = v.out_kps + (size_t)frame * g.out_cap + offset;
    uint8_t* odesc = v.out_desc + ((size_t)frame * g.out_cap + offset) * 32;
    const int lane = tid & 31, wid = tid >> 5;
    const float factorPI = (float)(3.14159265358979323846 / 180.0);  // (float)(CV_PI/180.f), :109
    const int nend = min(n, chunk0 + kDescChunk);
    // IC_Angle tables: the patch rows are fetched as aligned words, (row r, word q) -> task r*9 + q; a warp takes the 279
    // tasks of a keypoint in 9 coalesced steps. Per task: byte offset from the patch's aligned origin and
    // v | (4q - 15 + 32) << 8 | mask index << 16.
    for (int i = tid; i < 279; i += 256) {
        const int r = i / 9, q = i - 9 * r, vv = r - kHalfPatch;
        s_ioff[i] = r * rpitch + 4 * q;
        s_ipk[i] = (vv & 0xFF) | ((4 * q - kHalfPatch + 32) << 8) | ((abs(vv) * 9 + q) << 16);   // column weights are biased by +32
    }
    for (int i = tid; i < kIcMaskWords; i += 256) s_imask[i] = __ldg(v.ic_mask + i);
    for (int base = chunk0; base < nend; base += 256) {
        __syncthreads();
        const int m = min(256, nend - base);
        for (int j = wid; j < m; j += 8) {   // IC_Angle: m10 = sum u*I, m01 = sum v*I over the circular patch
            const LevelKey k = keys[base + j];
            const int x0 = (int)k.x - kHalfPatch, y0 = (int)k.y - kHalfPatch;   // coordinates are integers here
            const int al = x0 & 3;
            const uint8_t* org = raw + (size_t)y0 * rpitch + (x0 - al);
            uint32_t w[9];
            int pk[9];
#pragma unroll
            for (int t = 0; t < 9; t++) {
                const int i = lane + 32 * t;
                const bool on = t < 8 || i < 279;
                pk[t] = on ? s_ipk[i] : 0;
                w[t] = on ? __ldg(reinterpret_cast<const uint32_t*>(org + s_ioff[on ? i : 0])) : 0u;
            }
            int m10 = 0, m01 = 0;
            {
                uint32_t m10b = 0u, sum = 0u;   // sum of (u + 32) * I and of I over this lane's words
#pragma unroll
                for (int t = 0; t < 9; t++) {
                    const uint32_t mk = s_imask[al * 144 + (pk[t] >> 16)];
                    const uint32_t u0 = ((uint32_t)(pk[t] >> 8) & 0xFFu) - (uint32_t)al;      // biased column of the word's first byte: 14..49
                    const uint32_t u4 = u0 * 0x01010101u + 0x03020100u;                       // u0 .. u0+3, one per byte, no carries
                    m10b = __dp4a(w[t], u4 & mk, m10b);
                    const uint32_t rs = __dp4a(w[t], mk & 0x01010101u, 0u);
                    sum += rs;
                    m01 += (int)(signed char)pk[t] * (int)rs;
                }
                m10 = (int)m10b - 32 * (int)sum;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                m10 += __shfl_xor_sync(0xffffffffu, m10, o);
                m01 += __shfl_xor_sync(0xffffffffu, m01, o);
            }
            if (lane == 0) {
                const float ang = fast_atan2_deg((float)m01, (float)m10);
                s_angle[j] = ang;
                keys[base + j].angle = ang;
            }
        }
        __syncthreads();
        if (base + tid < nend) {
            // (float)cos(angle), (float)sin(angle) with the float argument promoted to double (:114-115); once per keypoint
            const float angle = __fmul_rn(s_angle[tid], factorPI);
            s_cs[tid] = make_float2((float)cos((double)angle), (float)sin((double)angle));
        }
        __syncthreads();
        for (int j = wid; j < m; j += 8) {
            const int i = base + j;
            const LevelKey k = keys[i];
            const float a = s_cs[j].x, b = s_cs[j].y;
            const int cx = __float2int_rn(k.x), cy = __float2int_rn(k.y);
            const uint8_t* center = img + (size_t)cy * pitch + cx;
            int val = 0;
#pragma unroll
            for (int bit = 0; bit < 8; bit++) {
                const float x0 = s_pat[(4 * bit) * 32 + lane], y0 = s_pat[(4 * bit + 1) * 32 + lane];
                const float x1 = s_pat[(4 * bit + 2) * 32 + lane], y1 = s_pat[(4 * bit + 3) * 32 + lane];
                // cvRound(x*b + y*a) rows, cvRound(x*a - y*b) cols: separate roundings, half-to-even (:121-122)
                const int r0 = round_even(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
                const int c0 = round_even(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
                const int r1 = round_even(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
                const int c1 = round_even(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
                const int t0 = __ldg(center + (ptrdiff_t)r0 * pitch + c0);
                const int t1 = __ldg(center + (ptrdiff_t)r1 * pitch + c1);
                val |= (t0 < t1) << bit;
            }
            odesc[(size_t)i * 32 + lane] = (uint8_t)val;
            if (v.mirror_desc) v.mirror_desc[(size_t)(offset + i) * 32 + lane] = (uint8_t)val;
            if (lane == 0) {
                coeb_keypoint o;
                o.x = level != 0 ? __fmul_rn(k.x, L.scale) : k.x;   // keypoint->pt *= scale (:1327-1334)
                o.y = level != 0 ? __fmul_rn(k.y, L.scale) : k.y;
                o.size = (float)L.scaled_patch;
                o.angle = s_angle[j];
                o.response = k.response;
                o.octave = level;
                o.class_id = -1;
                okp[i] = o;
                if (v.mirror_kps) v.mirror_kps[offset + i] = o;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// TMA variant: same grid, same three phases, but a keypoint's pixels arrive as TMA boxes in shared memory instead of
// per-lane global gathers. A box must start on a 16-byte boundary of its row (measured: any other start raises an illegal
// instruction), so it begins at the patch's first column rounded down to 16 and is 16 bytes wider than the patch:
//   orientation: box 48 x 31 bytes around (x-15, y-15) of the level; lane r owns patch row r: nine shared words from the
//                patch's first word, funnel-shifted to the first column, and 16 IDP.4A against per-lane constant weight
//                words (computed once per CTA from umax);
//   descriptor : box 64 x 37 bytes around (x-18, y-18) of the blurred level (the steered pattern stays within 18 px of the
//                centre: |(x, y)| <= 13 * sqrt(2)); the 512 samples are byte loads from shared memory.
// Every warp double-buffers its boxes: the next keypoint's box is in flight while the current one is reduced.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kRawBoxW = 48, kRawBoxH = kPatch;            // 1488 bytes
constexpr int kBlurBoxW = 64, kBlurBoxH = 37, kBlurR = 18;  // 2368 bytes
constexpr int kBoxSlot = 2432;                              // bytes per buffer slot (128-byte multiple, holds either box)
static_assert(kDescChunk <= 256, "one thread per keypoint computes cos / sin");

// kMirror: single-frame host calls, results also go to mapped pinned host memory (BatchView::mirror_*)
template <bool kMirror>
__global__ void __launch_bounds__(256, COEB_DESC_MINB) describe_tma_kernel(const __grid_constant__ TmaMaps raw_maps, const __grid_constant__ TmaMaps blur_maps,
                                                                          const __grid_constant__ Geometry g, const __grid_constant__ BatchView v,
                                                                          const int chunk /* keypoints per CTA, <= kDescChunk */) {
    COEB_TRACE(v, 10);
    __shared__ __align__(128) uint8_t s_box[8][2][kBoxSlot];
    __shared__ __align__(8) unsigned long long s_bar[8][2];
    __shared__ float2 s_cs[kDescChunk];
    __shared__ float s_angle[kDescChunk];
    __shared__ LevelKey s_key[kDescChunk];   // the chunk's keypoints, fetched once (every warp reads each of its keypoints three times)
    // chunk index fastest: the CTAs that read one level of one frame run together, so overlapping boxes hit in L2
    const int level = blockIdx.y, frame = blockIdx.z;
    const LevelGeom& L = g.lv[level];
    const int tid = threadIdx.x;

    const int* kc = v.key_count + frame * g.nlevels;
    int offset = 0, total = 0;
    for (int l = 0; l < g.nlevels; l++) {
        const int c = kc[l];
        if (l < level) offset += c;
        total += c;
    }
    const int n = kc[level];
    const int status = v.status[frame];
    const int chunk0 = blockIdx.x * chunk;
    if (level == 0 && blockIdx.x == 0 && tid == 0) {
        int cnt = total;
        if (status != COEB_OK) cnt = 0;
        else if (total > g.out_cap) { v.status[frame] = COEB_ERR_CAPACITY; }
        v.out_count[frame] = cnt;
        if (kMirror && v.mirror_hdr) { v.mirror_hdr[0] = cnt; v.mirror_hdr[1] = status != COEB_OK ? status : (total > g.out_cap ? (int)COEB_ERR_CAPACITY : (int)COEB_OK); }
    }
    if (chunk0 >= n || status != COEB_OK || total > g.out_cap) return;   // uniform
    const int lane = tid & 31, wid = tid >> 5;
    const uint32_t a_bar0 = (uint32_t)__cvta_generic_to_shared(&s_bar[wid][0]), a_bar1 = (uint32_t)__cvta_generic_to_shared(&s_bar[wid][1]);
    const uint32_t a_box0 = (uint32_t)__cvta_generic_to_shared(&s_box[wid][0][0]), a_box1 = (uint32_t)__cvta_generic_to_shared(&s_box[wid][1][0]);
    if (lane == 0) { tma_bar_init(a_bar0); tma_bar_init(a_bar1); }

    LevelKey* keys = v.keys + (size_t)frame * g.keys_per_frame + L.key_base;
    coeb_keypoint* okp = v.out_kps + (size_t)frame * g.out_cap + offset;
    uint8_t* odesc = v.out_desc + ((size_t)frame * g.out_cap + offset) * 32;
    const float factorPI = (float)(3.14159265358979323846 / 180.0);
    const int nend = min(n, chunk0 + chunk);
    const int m = nend - chunk0;   // <= kDescChunk keypoints, warp w takes j = w, w + 8, ...

    // per-lane constants of patch row r = lane (v = r - 15): weight words (u + 32 inside the circular patch, else 0) and
    // 0/1 mask words for the row sum; byte b of the row is column u = b - 15, byte 31 is padding
    uint32_t wq[8], mq[8];
    {
        const int vv = lane - kHalfPatch;
        const int um = lane < kPatch ? g.umax[abs(vv)] : -1;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            uint32_t wgt = 0u, msk = 0u;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int u = 4 * q + b - kHalfPatch;
                if (abs(u) <= um && 4 * q + b < kPatch) { wgt |= (uint32_t)(u + 32) << (8 * b); msk |= 1u << (8 * b); }
            }
            wq[q] = wgt; mq[q] = msk;
        }
    }
    if (tid < m) s_key[tid] = keys[chunk0 + tid];
    // the lane's 32 pattern coordinates (descriptor byte `lane`: pairs 8*lane .. 8*lane+7) as 8 words of biased bytes
    // (x0, y0, x1, y1) + 128: a byte becomes its float by one PRMT into the mantissa of 2^23 and one subtraction, instead of a
    // shared-memory load per coordinate (the stage is bound by shared-memory traffic, the ALU / FMA pipes are idle)
    uint32_t pw[8];
#pragma unroll
    for (int bit = 0; bit < 8; bit++) pw[bit] = *reinterpret_cast<const uint32_t*>(&c_pattern[lane * 32 + 4 * bit]) ^ 0x80808080u;
    __syncthreads();   // barriers initialised, keypoints staged
    uint32_t par0 = 0u, par1 = 0u;

    // ---- phase 1: IC_Angle ----
    {
        const CUtensorMap* map = &raw_maps.m[level];
        if (lane == 0 && wid < m) {
            const LevelKey k = s_key[wid];
            tma_load_box(a_bar0, a_box0, map, ((int)k.x - kHalfPatch) & ~15, (int)k.y - kHalfPatch, frame, kRawBoxW * kRawBoxH);
        }
        int it = 0;
        for (int j = wid; j < m; j += 8, it++) {
            const int slot = it & 1;
            if (lane == 0 && j + 8 < m) {
                const LevelKey k = s_key[j + 8];
                tma_load_box(slot ? a_bar0 : a_bar1, slot ? a_box0 : a_box1, map, ((int)k.x - kHalfPatch) & ~15, (int)k.y - kHalfPatch, frame, kRawBoxW * kRawBoxH);
            }
            if (slot) { tma_wait_parity(a_bar1, par1); par1 ^= 1u; } else { tma_wait_parity(a_bar0, par0); par0 ^= 1u; }
            const int ax = ((int)s_key[j].x - kHalfPatch) & 15;   // patch's first column inside the box
            const uint32_t* row = reinterpret_cast<const uint32_t*>(&s_box[wid][slot][(lane < kPatch ? lane : 0) * kRawBoxW]) + (ax >> 2);
            const int sh = 8 * (ax & 3);
            uint32_t m10b = 0u, sum = 0u;
            uint32_t prev = row[0];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint32_t next = row[q + 1];
                const uint32_t w = __funnelshift_r(prev, next, sh);   // bytes ax + 4q .. ax + 4q + 3 of the row
                m10b = __dp4a(w, wq[q], m10b);
                sum = __dp4a(w, mq[q], sum);
                prev = next;
            }
            int m10 = (int)m10b - 32 * (int)sum;
            int m01 = (lane - kHalfPatch) * (int)sum;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                m10 += __shfl_xor_sync(0xffffffffu, m10, o);
                m01 += __shfl_xor_sync(0xffffffffu, m01, o);
            }
            if (lane == 0) {
                const float ang = fast_atan2_deg((float)m01, (float)m10);
                s_angle[j] = ang;
                keys[chunk0 + j].angle = ang;
            }
            __syncwarp();   // every lane is done with this slot before it is refilled
        }
    }
    __syncthreads();
    if (tid < m) {
        const float angle = __fmul_rn(s_angle[tid], factorPI);
        s_cs[tid] = make_float2((float)cos((double)angle), (float)sin((double)angle));
    }
    __syncthreads();

    // ---- phase 2: steered rBRIEF on the blurred level ----
    {
        const CUtensorMap* map = &blur_maps.m[level];
        int it = 0;
        // the slots continue to alternate from 0: both barriers have completed an even or odd number of phases, tracked in par0 / par1
        if (lane == 0 && wid < m) {
            const LevelKey k = s_key[wid];
            tma_load_box(a_bar0, a_box0, map, (__float2int_rn(k.x) - kBlurR) & ~15, __float2int_rn(k.y) - kBlurR, frame, kBlurBoxW * kBlurBoxH);
        }
        for (int j = wid; j < m; j += 8, it++) {
            const int slot = it & 1;
            const int i = chunk0 + j;
            const LevelKey k = s_key[j];
            if (lane == 0 && j + 8 < m) {
                const LevelKey kn = s_key[j + 8];
                tma_load_box(slot ? a_bar0 : a_bar1, slot ? a_box0 : a_box1, map, (__float2int_rn(kn.x) - kBlurR) & ~15, __float2int_rn(kn.y) - kBlurR, frame,
                             kBlurBoxW * kBlurBoxH);
            }
            const float a = s_cs[j].x, b = s_cs[j].y;
            if (slot) { tma_wait_parity(a_bar1, par1); par1 ^= 1u; } else { tma_wait_parity(a_bar0, par0); par0 ^= 1u; }
            const uint8_t* center = &s_box[wid][slot][kBlurR * kBlurBoxW + kBlurR + ((__float2int_rn(k.x) - kBlurR) & 15)];
            int val = 0;
#pragma unroll
            for (int bit = 0; bit < 8; bit++) {
                const float kBias = 8388736.f;   // 2^23 + 128
                const float x0 = __fsub_rn(__uint_as_float(__byte_perm(pw[bit], 0x4B000000u, 0x7650)), kBias);
                const float y0 = __fsub_rn(__uint_as_float(__byte_perm(pw[bit], 0x4B000000u, 0x7651)), kBias);
                const float x1 = __fsub_rn(__uint_as_float(__byte_perm(pw[bit], 0x4B000000u, 0x7652)), kBias);
                const float y1 = __fsub_rn(__uint_as_float(__byte_perm(pw[bit], 0x4B000000u, 0x7653)), kBias);
                const int r0 = round_even(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
                const int c0 = round_even(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
                const int r1 = round_even(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
                const int c1 = round_even(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
                const int t0 = center[r0 * kBlurBoxW + c0];
                const int t1 = center[r1 * kBlurBoxW + c1];
                val |= (t0 < t1) << bit;
            }
            odesc[(size_t)i * 32 + lane] = (uint8_t)val;
            if (kMirror && v.mirror_desc) v.mirror_desc[(size_t)(offset + i) * 32 + lane] = (uint8_t)val;
            if (lane == 0) {
                coeb_keypoint o;
                o.x = level != 0 ? __fmul_rn(k.x, L.scale) : k.x;
                o.y = level != 0 ? __fmul_rn(k.y, L.scale) : k.y;
                o.size = (float)L.scaled_patch;
                o.angle = s_angle[j];
                o.response = k.response;
                o.octave = level;
                o.class_id = -1;
                okp[i] = o;
                if (kMirror && v.mirror_kps) v.mirror_kps[offset + i] = o;
            }
            __syncwarp();
        }
    }
}

// 0xFF for every byte of aligned word q of a patch row |v| that lies inside the circular patch, for the four alignments
// of the patch's first column (byte j of word q is column u = 4q + j - al - 15).
void build_ic_masks(const Geometry& g, uint32_t* out) {
    for (int al = 0; al < 4; al++)
        for (int av = 0; av < 16; av++)
            for (int q = 0; q < 9; q++) {
                uint32_t m = 0;
                for (int j = 0; j < 4; j++) {
                    const int u = 4 * q + j - al - kHalfPatch;
                    if (std::abs(u) <= g.umax[av]) m |= 0xFFu << (8 * j);
                }
                out[(al * 16 + av) * 9 + q] = m;
            }
}

void launch_describe(const Geometry& g, const BatchView& v, cudaStream_t stream) {
    int max_keys = 1;
    for (int l = 0; l < g.nlevels; l++) max_keys = std::max(max_keys, g.lv[l].key_cap);
    const dim3 grid(g.nlevels, v.B, (max_keys + kDescChunk - 1) / kDescChunk);
    TmaMaps raw_maps, blur_maps;
    if (tma_enabled() && encode_level_maps(g, v, kRawBoxW, kRawBoxH, &raw_maps) && encode_level_maps(g, v, kBlurBoxW, kBlurBoxH, &blur_maps, true)) {
        // a few frames (the tracking thread's single-frame call): 16 keypoints per CTA, i.e. 2 per warp, so that the level's
        // keypoints spread over four times as many SMs; batches keep 64 per CTA
        const int chunk = v.B <= 4 ? 16 : kDescChunk;
        if (v.mirror_hdr) describe_tma_kernel<true><<<dim3((max_keys + chunk - 1) / chunk, g.nlevels, v.B), 256, 0, stream>>>(raw_maps, blur_maps, g, v, chunk);
        else describe_tma_kernel<false><<<dim3((max_keys + chunk - 1) / chunk, g.nlevels, v.B), 256, 0, stream>>>(raw_maps, blur_maps, g, v, chunk);
    }
    else {
        describe_kernel<<<grid, 256, 0, stream>>>(g, v);
    }
}

}  // namespace coeb
