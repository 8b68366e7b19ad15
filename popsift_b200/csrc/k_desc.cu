// Stage 3b -- 128-D descriptors + normalisation + Feature records, hand-written for sm_100a.
//
// Replaces ext_desc_loop / ext_desc_loop_sub, normalize_histogram<RootSift|L2> and prep_features
// (reference src/popsift/s_desc_loop.cu:11-158, s_desc_norm_rs.h:41-77, s_desc_norm_l2.h:46-135,
// s_desc_normalize.h:14-33, sift_pyramid.cu:250-280).  Structure:
//   * one launch for all octaves; a fixed grid of 512-thread CTAs walks the device-side descriptor
//     count (the reference reads the count back to the host, sift_desc.cu:55-110);
//   * normalisation is fused: the 16 cell-warps deposit their 8 bins in shared memory and warp 0
//     normalises and stores the 512-byte descriptor with one float4 per lane -- one launch and one
//     HBM round trip less, and without the reference's last-descriptor race.
// Per-sample math follows the reference line by line (same intrinsics: __sincosf, __expf,
// __fmul_ru, __fmaf_ru; per-lane partial sums; shuffle-down tree 16,8,4,2,1).
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

__global__ void __launch_bounds__(512)
descriptor_kernel(PyramidView pyr, Consts k, const ps_extremum* __restrict__ ext,
                  const int* __restrict__ feat_to_ext, ps_descriptor* __restrict__ desc, const Counters* ct)
{
    __shared__ __align__(16) float feat[128];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;       // cell index: ix = warp & 3, iy = warp >> 2
    const int ix = warp & 3, iy = warp >> 2;
    const int total = ct->ori_total;

    for (int d = blockIdx.x; d < total; d += gridDim.x) {
        const int ei = feat_to_ext[d];
        const ps_extremum e = ext[ei];
        const float ang = e.orientation[min(max(d - e.idx_ori, 0), PS_MAX_ORI - 1)];
        const OctaveView& ov = pyr.oct[e.octave];
        const int width = ov.w, height = ov.h;
        const int lvl = min(max(e.lpos, 0), pyr.levels + 2);
        const float* pl = ov.gauss + (size_t)lvl * ov.plane;

        const float x = e.xpos, y = e.ypos;
        const float SBP = fabsf(__fmul_rn(3.0f, e.sigma));
        float dpt[9];
#pragma unroll
        for (int b = 0; b < 9; ++b) dpt[b] = 0.0f;

        if (SBP != 0.0f) {
            float sin_t, cos_t;
            __sincosf(ang, &sin_t, &cos_t);
            const float csbp = __fmul_rn(cos_t, SBP), ssbp = __fmul_rn(sin_t, SBP);
            const float crsbp = __fdiv_rn(cos_t, SBP), srsbp = __fdiv_rn(sin_t, SBP);
            const float ox = (float)ix - 1.5f, oy = (float)iy - 1.5f;
            const float ptx = __fmaf_rn(csbp, ox, __fmaf_rn(-ssbp, oy, x));
            const float pty = __fmaf_rn(csbp, oy, __fmaf_rn(ssbp, ox, y));
            const float bsz = __fadd_rn(fabsf(csbp), fabsf(ssbp));
            const int xmin = max(1, (int)floorf(__fsub_rn(ptx, bsz)));
            const int ymin = max(1, (int)floorf(__fsub_rn(pty, bsz)));
            const int xmax = min(width - 2, (int)floorf(__fadd_rn(ptx, bsz)));
            const int ymax = min(height - 2, (int)floorf(__fadd_rn(pty, bsz)));
            const int wx = xmax - xmin + 1;
            const int hy = ymax - ymin + 1;
            const int loops = (wx > 0 && hy > 0) ? wx * hy : 0;

            for (int i = lane; i < loops; i += 32) {
                const int q = i / wx;
                const int ii = q + ymin;
                const int jj = i - q * wx + xmin;
                const float ddx = __fsub_rn((float)jj, ptx), ddy = __fsub_rn((float)ii, pty);
                const float nx = __fmaf_rn(crsbp, ddx, __fmul_rn(srsbp, ddy));
                const float ny = __fmaf_rn(crsbp, ddy, __fmul_rn(-srsbp, ddx));
                const float nnx = fabsf(nx), nny = fabsf(ny);
                if (nnx < 1.0f && nny < 1.0f) {
                    const float gdx = __fsub_rn(plane_at(pl, width, height, ov.pitch, jj + 1, ii),
                                                plane_at(pl, width, height, ov.pitch, jj - 1, ii));
                    const float gdy = __fsub_rn(plane_at(pl, width, height, ov.pitch, jj, ii + 1),
                                                plane_at(pl, width, height, ov.pitch, jj, ii - 1));
                    const float mod = hypotf(gdx, gdy);
                    float th = atan2f(gdy, gdx);
                    const float dnx = __fadd_rn(nx, ox), dny = __fadd_rn(ny, oy);
                    const float ww = __expf(-__fmul_rn(__fmaf_rn(dny, dny, __fmul_rn(dnx, dnx)), 0.125f));
                    const float wgt = __fmul_rn(__fmul_rn(__fmul_rn(ww, __fsub_rn(1.0f, nnx)), __fsub_rn(1.0f, nny)), mod);
                    th = __fsub_rn(th, ang);
                    th = __fadd_rn(th, th < 0.0f ? dPi2 : 0.0f);
                    th = __fsub_rn(th, th >= dPi2 ? dPi2 : 0.0f);
                    const float tth = __fmul_ru(th, d4RPi);
                    const int fo0 = (int)floorf(tth);
                    const float do0 = __fsub_rn(tth, (float)fo0);
                    const float wgt1 = __fsub_rn(1.0f, do0), wgt2 = do0;
                    const int fo = fo0 % 8;
                    // dpt[fo] / dpt[fo+1] with a compile-time-indexed register array
#pragma unroll
                    for (int b = 0; b < 8; ++b)
                        if (b == fo) { dpt[b] = __fmaf_ru(wgt1, wgt, dpt[b]); dpt[b + 1] = __fmaf_ru(wgt2, wgt, dpt[b + 1]); }
                }
            }
        }
        dpt[0] += dpt[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) dpt[b] = tree_down(dpt[b]);
        if (lane == 0) {
            float* f = feat + ((iy << 2) + ix) * 8;
#pragma unroll
            for (int b = 0; b < 8; ++b) f[b] = dpt[b];
        }
        __syncthreads();
        if (warp == 0) {
            float4 v = reinterpret_cast<const float4*>(feat)[lane];
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
        __syncthreads();
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
    descriptor_kernel<<<148 * 4, 512, 0, st>>>(pyr, k, ext, feat_to_ext, desc, ct);
    return 1;
}

int launch_prep_features(const Consts& k, const ps_extremum* ext, ps_feature* feat, const Counters* ct, cudaStream_t st)
{
    prep_features_kernel<<<148, 256, 0, st>>>(k, ext, feat, ct);
    return 1;
}

} // namespace psb
