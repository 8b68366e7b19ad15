// Descriptor normalisation shared by the descriptor kernels (reference normalize_histogram<T>, s_desc_normalize.h:14-33;
// NormalizeRootSift, s_desc_norm_rs.h:41-77; NormalizeL2, s_desc_norm_l2.h:46-135).  One warp per descriptor, lane k holds
// floats 4k..4k+3, sums through the reference's shuffle-down tree 16, 8, 4, 2, 1.
#pragma once
#include "ps_internal.h"

namespace psb {

__device__ __forceinline__ float tree_down(float v)
{   // lane 0 ends with the reference's shuffle-down tree sum
    v += __shfl_down_sync(0xffffffffu, v, 16);
    v += __shfl_down_sync(0xffffffffu, v, 8);
    v += __shfl_down_sync(0xffffffffu, v, 4);
    v += __shfl_down_sync(0xffffffffu, v, 2);
    v += __shfl_down_sync(0xffffffffu, v, 1);
    return v;
}

__device__ __forceinline__ float4 normalize_descriptor(float4 v, int lane, int norm_mode, int norm_multi)
{
    if (norm_mode == PS_NORM_ROOTSIFT) {
        float sum = __fadd_rn(__fadd_rn(__fadd_rn(v.x, v.y), v.z), v.w);
        sum = __shfl_sync(0xffffffffu, tree_down(sum), 0);
        v.x = scalbnf(__fsqrt_rn(__fdividef(v.x, sum)), norm_multi);
        v.y = scalbnf(__fsqrt_rn(__fdividef(v.y, sum)), norm_multi);
        v.z = scalbnf(__fsqrt_rn(__fdividef(v.z, sum)), norm_multi);
        v.w = scalbnf(__fsqrt_rn(__fdividef(v.w, sum)), norm_multi);
    } else {
        float n = __fmaf_rn(v.w, v.w, __fmaf_rn(v.z, v.z, __fmaf_rn(v.y, v.y, __fmul_rn(v.x, v.x))));
        n = tree_down(n);
        if (lane == 0) n = __fsqrt_rn(n);
        n = __shfl_sync(0xffffffffu, n, 0);
        const float lim = __fmul_rn(0.2f, n);
        v.x = fminf(v.x, lim); v.y = fminf(v.y, lim); v.z = fminf(v.z, lim); v.w = fminf(v.w, lim);
        n = __fmaf_rn(v.w, v.w, __fmaf_rn(v.z, v.z, __fmaf_rn(v.y, v.y, __fmul_rn(v.x, v.x))));
        n = tree_down(n);
        if (lane == 0) n = scalbnf(__frsqrt_rn(n), norm_multi);
        n = __shfl_sync(0xffffffffu, n, 0);
        v.x = __fmul_rn(v.x, n); v.y = __fmul_rn(v.y, n); v.z = __fmul_rn(v.z, n); v.w = __fmul_rn(v.w, n);
    }
    return v;
}

} // namespace psb
