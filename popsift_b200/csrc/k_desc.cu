// Stage 3b -- 128-D descriptors + normalisation + Feature records, hand-written for sm_100a.
//
// Replaces ext_desc_loop / ext_desc_loop_sub, normalize_histogram<RootSift|L2> and prep_features
// (reference src/popsift/s_desc_loop.cu:11-158, s_desc_norm_rs.h:41-77, s_desc_norm_l2.h:46-135,
// s_desc_normalize.h:14-33, sift_pyramid.cu:250-280).  Structure:
//   * one launch for all octaves; a fixed grid of warps pulls descriptors from a device-side work
//     counter (the reference reads the count back to the host, sift_desc.cu:55-110);
//   * ONE small CTA (4 warps) per descriptor visits every pixel of the rotated 5x5-SBP support once: gradient,
//     Gaussian weight and orientation split are computed once per pixel and scattered to the (at
//     most) 2x2 cells whose bilinear windows cover it.  The reference runs 16 warps per descriptor,
//     each scanning its own cell's bounding box: ~2.6x redundant atan2f/hypotf and mostly-idle
//     lanes; this form issues ~9x fewer instructions;
//   * the 128 bins live in shared memory (one 512-byte histogram per warp); normalisation is fused
//     and the descriptor leaves with one float4 per lane -- one launch and one HBM round trip less,
//     and without the reference's last-descriptor race.
// Per-sample math keeps the reference's intrinsics (__sincosf, __expf, __fmul_ru, hypotf, atan2f);
// the accumulation differs (order-independent fixed-point shared-memory adds instead of per-lane
// round-up float sums and a shuffle tree), which moves normalised descriptors by ~1e-6 (tolerance of
// the task: 1e-3).
#include "ps_internal.h"

namespace psb {

namespace {

__device__ const float dPi2 = 2.0f * 3.14159265358979323846f;
__device__ const float d4RPi = 4.0f / 3.14159265358979323846f;

__device__ __forceinline__ float plane_at(const float* pl, int w, int h, int pitch, int x, int y)
{
    x = min(max(x, 0), w - 1);
    y = min(max(y, 0), h - 1);
    return __ldg(pl + (size_t)y * pitch + x);
}

__device__ __forceinline__ float tree_down(float v)
{   // lane 0 ends with the reference's shuffle-down tree sum
    v += __shfl_down_sync(0xffffffffu, v, 16);
    v += __shfl_down_sync(0xffffffffu, v, 8);
    v += __shfl_down_sync(0xffffffffu, v, 4);
    v += __shfl_down_sync(0xffffffffu, v, 2);
    v += __shfl_down_sync(0xffffffffu, v, 1);
    return v;
}

constexpr int DWARPS = 4;     // warps that share one descriptor (one CTA per descriptor in flight)
// a bin can reach ~8e4 (232 fully weighted pixels of gradient magnitude 360): 17 integer bits
__device__ const float kFix = 32768.0f;
__device__ const float kUnfix = 1.0f / 32768.0f;

__global__ void __launch_bounds__(DWARPS * 32)
descriptor_kernel(PyramidView pyr, Consts k, const ps_extremum* __restrict__ ext,
                  const int* __restrict__ feat_to_ext, ps_descriptor* __restrict__ desc, Counters* ct)
{
    // 128 bins per warp in 32-bit fixed point (15 fractional bits): integer shared-memory atomics are
    // native (ATOMS.ADD) whereas float ones are compare-and-swap loops, and integer sums do not depend
    // on the order of the adds -> run-to-run deterministic descriptors
    __shared__ __align__(16) unsigned H[128];
    __shared__ int next_d;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int total = ct->ori_total;

    for (;;) {
        // dynamic distribution: descriptor sizes vary by two orders of magnitude with scale
        if (threadIdx.x == 0) next_d = atomicAdd(&ct->work_desc, 1);
        __syncthreads();
        const int d = next_d;
        if (d >= total) break;

        const int ei = feat_to_ext[d];
        const ps_extremum e = ext[ei];
        const float ang = e.orientation[min(max(d - e.idx_ori, 0), PS_MAX_ORI - 1)];
        const OctaveView& ov = pyr.oct[e.octave];
        const int width = ov.w, height = ov.h, pitch = ov.pitch;
        const int lvl = min(max(e.lpos, 0), pyr.levels + 2);
        const float* pl = ov.gauss + (size_t)lvl * ov.plane;

        if (warp == 0) *reinterpret_cast<uint4*>(H + 4 * lane) = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();

        const float x = e.xpos, y = e.ypos;
        const float SBP = fabsf(__fmul_rn(3.0f, e.sigma));
        if (SBP != 0.0f) {
            float sin_t, cos_t;
            __sincosf(ang, &sin_t, &cos_t);
            const float csbp = __fmul_rn(cos_t, SBP), ssbp = __fmul_rn(sin_t, SBP);
            const float crsbp = __fdiv_rn(cos_t, SBP), srsbp = __fdiv_rn(sin_t, SBP);
            // union of the 16 cell windows: |r| < 2.5 in units of SBP, r = R(-ang) (p - keypoint) / SBP
            const float half = __fmul_rn(2.5f, __fadd_rn(fabsf(csbp), fabsf(ssbp)));
            const int xmin = max(1, (int)floorf(x - half) - 1);
            const int ymin = max(1, (int)floorf(y - half) - 1);
            const int xmax = min(width - 2, (int)floorf(x + half) + 1);
            const int ymax = min(height - 2, (int)floorf(y + half) + 1);
            const int wx = xmax - xmin + 1;
            const int hy = ymax - ymin + 1;
            const int loops = (wx > 0 && hy > 0) ? wx * hy : 0;
            const float inv_wx = 1.0f / (float)max(wx, 1);

            for (int i = threadIdx.x; i < loops; i += DWARPS * 32) {
                // i / wx without an integer division (exact for the sizes at hand: fix up by one)
                int q = (int)((float)i * inv_wx);
                if (q * wx > i) --q;
                if ((q + 1) * wx <= i) ++q;
                const int ii = q + ymin;
                const int jj = i - q * wx + xmin;
                const float ddx = (float)jj - x, ddy = (float)ii - y;
                const float rx = __fmaf_rn(crsbp, ddx, __fmul_rn(srsbp, ddy));
                const float ry = __fmaf_rn(crsbp, ddy, __fmul_rn(-srsbp, ddx));
                if (!(fabsf(rx) < 2.5f && fabsf(ry) < 2.5f)) continue;
                const float* p = pl + (size_t)ii * pitch + jj;       // interior pixel: neighbours exist
                const float gdx = __fsub_rn(__ldg(p + 1), __ldg(p - 1));
                const float gdy = __fsub_rn(__ldg(p + pitch), __ldg(p - pitch));
                const float mod = hypotf(gdx, gdy);
                float th = atan2f(gdy, gdx);
                const float ww = __expf(-__fmul_rn(__fmaf_rn(ry, ry, __fmul_rn(rx, rx)), 0.125f));
                th = __fsub_rn(th, ang);
                th = __fadd_rn(th, th < 0.0f ? dPi2 : 0.0f);
                th = __fsub_rn(th, th >= dPi2 ? dPi2 : 0.0f);
                const float tth = __fmul_ru(th, d4RPi);
                const int fo0 = (int)floorf(tth);
                const float do0 = __fsub_rn(tth, (float)fo0);
                const int b0 = fo0 & 7, b1 = (fo0 + 1) & 7;
                const float wm = __fmul_rn(ww, mod);
                const float w0 = __fmul_rn(__fsub_rn(1.0f, do0), wm), w1 = __fmul_rn(do0, wm);
                // the (at most) 2x2 cells whose bilinear window covers this pixel
                const float fx = __fadd_rn(rx, 1.5f), fy = __fadd_rn(ry, 1.5f);
                const float flx = floorf(fx), fly = floorf(fy);
                const int cx0 = (int)flx, cy0 = (int)fly;
                const float ax = __fsub_rn(fx, flx), ay = __fsub_rn(fy, fly);     // in [0,1)
#pragma unroll
                for (int sy = 0; sy < 2; ++sy) {
                    const int cy = cy0 + sy;
                    const float wy = sy ? ay : __fsub_rn(1.0f, ay);
                    if (cy < 0 || cy > 3 || wy <= 0.0f) continue;
#pragma unroll
                    for (int sx = 0; sx < 2; ++sx) {
                        const int cx = cx0 + sx;
                        const float wxx = sx ? ax : __fsub_rn(1.0f, ax);
                        if (cx < 0 || cx > 3 || wxx <= 0.0f) continue;
                        const float wc = __fmul_rn(wxx, wy);
                        unsigned* hb = H + ((cy << 2) + cx) * 8;
                        atomicAdd(hb + b0, __float2uint_rn(__fmul_rn(__fmul_rn(w0, wc), kFix)));
                        atomicAdd(hb + b1, __float2uint_rn(__fmul_rn(__fmul_rn(w1, wc), kFix)));
                    }
                }
            }
        }
        __syncthreads();
        if (warp == 0) {
            const uint4 hv = *reinterpret_cast<const uint4*>(H + 4 * lane);
            float4 v = make_float4((float)hv.x * kUnfix, (float)hv.y * kUnfix, (float)hv.z * kUnfix, (float)hv.w * kUnfix);
            if (k.norm_mode == PS_NORM_ROOTSIFT) {
                float sum = __fadd_rn(__fadd_rn(__fadd_rn(v.x, v.y), v.z), v.w);
                sum = __shfl_sync(0xffffffffu, tree_down(sum), 0);
                v.x = scalbnf(__fsqrt_rn(__fdividef(v.x, sum)), k.norm_multi);
                v.y = scalbnf(__fsqrt_rn(__fdividef(v.y, sum)), k.norm_multi);
                v.z = scalbnf(__fsqrt_rn(__fdividef(v.z, sum)), k.norm_multi);
                v.w = scalbnf(__fsqrt_rn(__fdividef(v.w, sum)), k.norm_multi);
            } else {
                float n = __fmaf_rn(v.w, v.w, __fmaf_rn(v.z, v.z, __fmaf_rn(v.y, v.y, __fmul_rn(v.x, v.x))));
                n = tree_down(n);
                if (lane == 0) n = __fsqrt_rn(n);
                n = __shfl_sync(0xffffffffu, n, 0);
                const float lim = __fmul_rn(0.2f, n);
                v.x = fminf(v.x, lim); v.y = fminf(v.y, lim); v.z = fminf(v.z, lim); v.w = fminf(v.w, lim);
                n = __fmaf_rn(v.w, v.w, __fmaf_rn(v.z, v.z, __fmaf_rn(v.y, v.y, __fmul_rn(v.x, v.x))));
                n = tree_down(n);
                if (lane == 0) n = scalbnf(__frsqrt_rn(n), k.norm_multi);
                n = __shfl_sync(0xffffffffu, n, 0);
                v.x = __fmul_rn(v.x, n); v.y = __fmul_rn(v.y, n); v.z = __fmul_rn(v.z, n); v.w = __fmul_rn(v.w, n);
            }
            reinterpret_cast<float4*>(desc[d].features)[lane] = v;
        }
        // the next iteration's first __syncthreads orders this read of H before the next reset
    }
}

// Extremum -> user-facing Feature (octave-local -> input-image coordinates); desc[] pointers are
// filled in on the host by ps_download.
__global__ void prep_features_kernel(Consts k, const ps_extremum* __restrict__ ext, ps_feature* __restrict__ feat,
                                     const Counters* ct)
{
    const int total = ct->ext_total;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const ps_extremum e = ext[i];
        const float s = powf(2.0f, (float)(e.octave - k.up_fac));
        ps_feature f;
        f.debug_octave = e.octave;
        f.xpos = __fmul_rn(e.xpos, s);
        f.ypos = __fmul_rn(e.ypos, s);
        f.sigma = __fmul_rn(e.sigma, s);
        f.num_ori = e.num_ori;
        f.pad_ = e.idx_ori;          // first descriptor index; ps_download turns it into pointers
#pragma unroll
        for (int r = 0; r < PS_MAX_ORI; ++r) {
            f.orientation[r] = r < e.num_ori ? e.orientation[r] : 0.0f;
            f.desc[r] = nullptr;
        }
        feat[i] = f;
    }
}

} // namespace

int launch_descriptors(const PyramidView& pyr, const Consts& k, const ps_extremum* ext, const int* feat_to_ext,
                       ps_descriptor* desc, Counters* ct, cudaStream_t st)
{
    descriptor_kernel<<<148 * 16, DWARPS * 32, 0, st>>>(pyr, k, ext, feat_to_ext, desc, ct);
    return 1;
}

int launch_prep_features(const Consts& k, const ps_extremum* ext, ps_feature* feat, const Counters* ct, cudaStream_t st)
{
    prep_features_kernel<<<148, 256, 0, st>>>(k, ext, feat, ct);
    return 1;
}

} // namespace psb
