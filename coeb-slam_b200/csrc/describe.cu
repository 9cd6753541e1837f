// describe.cu -- steered rBRIEF descriptors (E11) and output assembly (E12).
//
// Reference: computeOrbDescriptor / computeDescriptors (src/ORBextractor.cc:109-156, 1078-1085) and the
// assembly loop of ORBextractor::operator() (:1291-1337). One CTA per (level, frame); one warp per
// keypoint, lane i producing descriptor byte i from pattern pairs 8i..8i+7 on the blurred level.
#include "coeb_device.cuh"

namespace coeb {

__constant__ signed char c_pattern[1024] = {
#include "../../include/coeb_orb_pattern.inc"
};

constexpr int kDescChunk = 64;   // keypoints per CTA: several CTAs per level keep a single frame's latency low

__global__ void __launch_bounds__(256) describe_kernel(const __grid_constant__ Geometry g, const __grid_constant__ BatchView v) {
    __shared__ signed char s_pat[1024];   // transposed: [4*bit + component][lane], conflict-free per-lane reads
    __shared__ float2 s_cs[256];          // (cos, sin) of the keypoints of the current chunk
    const int level = blockIdx.x, frame = blockIdx.y;
    const LevelGeom& L = g.lv[level];
    const int tid = threadIdx.x;
    for (int i = tid; i < 1024; i += 256) s_pat[(i & 31) * 32 + (i >> 5)] = c_pattern[i];   // byte `i>>5` uses ints [32*(i>>5), +32)

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
    }
    __syncthreads();
    if (chunk0 >= n || status != COEB_OK || total > g.out_cap) return;

    const uint8_t* __restrict__ img = blur_ptr(g, v, level, frame);
    const int pitch = L.pitch;
    const LevelKey* keys = v.keys + (size_t)frame * g.keys_per_frame + L.key_base;
    coeb_keypoint* okp = v.out_kps + (size_t)frame * g.out_cap + offset;
    uint8_t* odesc = v.out_desc + ((size_t)frame * g.out_cap + offset) * 32;
    const int lane = tid & 31, wid = tid >> 5;
    const float factorPI = (float)(3.14159265358979323846 / 180.0);  // (float)(CV_PI/180.f), :109
    const int nend = min(n, chunk0 + kDescChunk);
    for (int base = chunk0; base < nend; base += 256) {
        __syncthreads();
        if (base + tid < nend) {
            // (float)cos(angle), (float)sin(angle) with the float argument promoted to double (:114-115); once per keypoint
            const float angle = __fmul_rn(keys[base + tid].angle, factorPI);
            s_cs[tid] = make_float2((float)cos((double)angle), (float)sin((double)angle));
        }
        __syncthreads();
        const int m = min(256, nend - base);
        for (int j = wid; j < m; j += 8) {
            const int i = base + j;
            const LevelKey k = keys[i];
            const float a = s_cs[j].x, b = s_cs[j].y;
            const int cx = __float2int_rn(k.x), cy = __float2int_rn(k.y);
            const uint8_t* center = img + (size_t)cy * pitch + cx;
            int val = 0;
#pragma unroll
            for (int bit = 0; bit < 8; bit++) {
                const float x0 = (float)s_pat[(4 * bit) * 32 + lane], y0 = (float)s_pat[(4 * bit + 1) * 32 + lane];
                const float x1 = (float)s_pat[(4 * bit + 2) * 32 + lane], y1 = (float)s_pat[(4 * bit + 3) * 32 + lane];
                // cvRound(x*b + y*a) rows, cvRound(x*a - y*b) cols: separate roundings, half-to-even (:121-122)
                const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
                const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
                const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
                const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
                const int t0 = __ldg(center + (ptrdiff_t)r0 * pitch + c0);
                const int t1 = __ldg(center + (ptrdiff_t)r1 * pitch + c1);
                val |= (t0 < t1) << bit;
            }
            odesc[(size_t)i * 32 + lane] = (uint8_t)val;
            if (lane == 0) {
                coeb_keypoint o;
                o.x = level != 0 ? __fmul_rn(k.x, L.scale) : k.x;   // keypoint->pt *= scale (:1327-1334)
                o.y = level != 0 ? __fmul_rn(k.y, L.scale) : k.y;
                o.size = (float)L.scaled_patch;
                o.angle = k.angle;
                o.response = k.response;
                o.octave = level;
                o.class_id = -1;
                okp[i] = o;
            }
        }
    }
}

void launch_describe(const Geometry& g, const BatchView& v, cudaStream_t stream) {
    int max_keys = 1;
    for (int l = 0; l < g.nlevels; l++) max_keys = std::max(max_keys, g.lv[l].key_cap);
    describe_kernel<<<dim3(g.nlevels, v.B, (max_keys + kDescChunk - 1) / kDescChunk), 256, 0, stream>>>(g, v);
}

}  // namespace coeb
