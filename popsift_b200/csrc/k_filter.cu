// Grid filter -- at most ~filter_max_extrema extrema per image, spread evenly over a g x g grid of cells.
//
// Replaces Pyramid::extrema_filter_grid (reference src/popsift/s_filtergrid.cu:112-325: a Thrust pipeline of fills,
// sorts by (cell, scale), reduce_by_key, host-side limit computation, for_each and copy_if with four host
// synchronisations) by three small kernels on the slot's stream, no host round trip:
//   plan     one block: per-cell counts in shared memory, the trigger test of s_orientation.cu:380-383
//            (int(1.1 * max) < total), and the per-cell limit exactly as the reference's host code computes it;
//   select   one block per over-full cell: 4-pass radix select of the limit-th scale (float bits are monotonic for
//            positive floats; inverted for largest-first), then the marking pass; ties at the cut are resolved in
//            arrival order like the reference's unstable sort resolves them arbitrarily;
//   compact  survivors are appended per octave to a second InitialExtremum array; the orientation stage reads that
//            array and the filtered counts when the filter fired.
// tests/filter_oracle.py restates the reference's algorithm; the kept sets are identical for the `up` / `down`
// orders, the per-cell counts for `random`.
#include "ps_internal.h"

namespace psb {

namespace {

constexpr int kPlanThreads = 1024;
constexpr int kMaxCells = PS_MAX_FILTER_GRID * PS_MAX_FILTER_GRID;     // 1024

// grid cell of an extremum (reference s_extrema.cu:499; w_grid_divider = float(w) / g, sift_octave.cu:40-41)
__device__ __forceinline__ int cell_of(const InitialExtremum& e, int w, int h, int g)
{
    const float wd = __fdiv_rn((float)w, (float)g), hd = __fdiv_rn((float)h, (float)g);
    const float c = __fadd_rn(__fmul_rn(floorf(__fdiv_rn(e.ypos, hd)), (float)g), floorf(__fdiv_rn(e.xpos, wd)));
    return min(max((int)c, 0), g * g - 1);
}

// scale key: sigma * 2^octave as ordered unsigned bits (all scales are positive)
__device__ __forceinline__ unsigned scale_key(const InitialExtremum& e, int octave, bool largest_first)
{
    const unsigned b = __float_as_uint(__fmul_rn(e.sigma, exp2f((float)octave)));
    return largest_first ? ~b : b;
}

struct Items {            // the initial extrema of all octaves as one index space
    const InitialExtremum* iext;
    int max_extrema;
    int first[kMaxOctaves + 1];
    int num_octaves;
    __device__ __forceinline__ int octave_of(int i) const
    {
        int o = 0;
        while (o + 1 < num_octaves && i >= first[o + 1]) ++o;
        return o;
    }
    __device__ __forceinline__ const InitialExtremum& at(int i, int o) const { return iext[(size_t)o * max_extrema + (i - first[o])]; }
};

__device__ __forceinline__ Items make_items(const InitialExtremum* iext, const Counters* ct, const Consts& k, int num_octaves)
{
    Items it;
    it.iext = iext; it.max_extrema = k.max_extrema; it.num_octaves = num_octaves;
    int total = 0;
    for (int o = 0; o < num_octaves; ++o) { it.first[o] = total; total += min(ct->ext_ct[o], k.max_extrema); }
    for (int o = num_octaves; o <= kMaxOctaves; ++o) it.first[o] = total;
    return it;
}

__global__ void __launch_bounds__(kPlanThreads)
filter_plan_kernel(PyramidView pyr, Consts k, FilterCfg fc, const InitialExtremum* __restrict__ iext, Counters* ct, FilterPlan* plan)
{
    __shared__ int cnt[kMaxCells];
    __shared__ int sorted[kMaxCells];
    const int ncell = fc.grid * fc.grid;
    for (int c = threadIdx.x; c < kMaxCells; c += kPlanThreads) cnt[c] = 0;
    __syncthreads();
    const Items it = make_items(iext, ct, k, pyr.num_octaves);
    const int total = it.first[kMaxOctaves];
    // the filter only runs when it has something to remove, with 10 % slack (reference s_orientation.cu:380-383)
    const bool active = fc.max_extrema > 0 && (int)((float)fc.max_extrema * 1.1f) < total;
    if (!active) {
        if (threadIdx.x == 0) { plan->active = 0; plan->limit = 0; ct->filtered = 0; }
        return;
    }
    for (int i = threadIdx.x; i < total; i += kPlanThreads) {
        const int o = it.octave_of(i);
        atomicAdd(&cnt[cell_of(it.at(i, o), pyr.oct[o].w, pyr.oct[o].h, fc.grid)], 1);
    }
    __syncthreads();
    // ascending bitonic sort of the ncell counts (padded with INT_MAX to 1024)
    for (int c = threadIdx.x; c < kMaxCells; c += kPlanThreads) {
        sorted[c] = c < ncell ? cnt[c] : 0x7fffffff;
        plan->cell_count[c] = c < ncell ? cnt[c] : 0;
    }
    __syncthreads();
    for (int size = 2; size <= kMaxCells; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const int i = threadIdx.x;
            const int j = i ^ stride;
            if (j > i) {
                const bool up = (i & size) == 0;
                const int a = sorted[i], b = sorted[j];
                if ((a > b) == up) { sorted[i] = b; sorted[j] = a; }
            }
            __syncthreads();
        }
    if (threadIdx.x == 0) {
        // reference s_filtergrid.cu:225-262 (host code there): sumup[i] = prefix[i] + c[i] * (n-1-i)
        int ctn = 0;
        long long prefix = 0;
        for (int i = 0; i < ncell; ++i) {
            prefix += sorted[i];
            if (prefix + (long long)sorted[i] * (ncell - 1 - i) > fc.max_extrema) ++ctn;
        }
        int limit = 0x7fffffff;
        if (ctn > 0) {
            long long tail = 0;
            for (int i = ncell - ctn; i < ncell; ++i) tail += sorted[i];
            const float tailaverage = __fdiv_rn((float)tail, (float)ctn);
            limit = (int)ceilf(__fsub_rn(tailaverage, (float)((total - fc.max_extrema) / ctn)));
        }
        plan->active = 1;
        plan->limit = limit;
        plan->total = total;
        ct->filtered = 1;
    }
}

__global__ void __launch_bounds__(256)
filter_select_kernel(PyramidView pyr, Consts k, FilterCfg fc, const InitialExtremum* __restrict__ iext, const Counters* ct,
                     const FilterPlan* __restrict__ plan, unsigned char* __restrict__ keep)
{
    if (!plan->active) return;
    const int cell = blockIdx.x;
    const int count = plan->cell_count[cell];
    const int limit = plan->limit;
    if (count <= limit) return;                                // every extremum of this cell stays (keep[] is preset to 1)
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_mask;
    __shared__ int s_k, s_ties;
    const Items it = make_items(iext, ct, k, pyr.num_octaves);
    const int total = it.first[kMaxOctaves];
    const bool largest_first = fc.sort == PS_FILTER_LARGEST_FIRST;
    if (fc.sort == PS_FILTER_RANDOM) {
        // no order: the first `limit` arrivals stay
        if (threadIdx.x == 0) s_ties = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int o = it.octave_of(i);
            if (cell_of(it.at(i, o), pyr.oct[o].w, pyr.oct[o].h, fc.grid) != cell) continue;
            if (atomicAdd(&s_ties, 1) >= limit) keep[i] = 0;
        }
        return;
    }
    // radix select of the limit-th smallest key (0-based rank limit - 1) among the cell's extrema
    if (threadIdx.x == 0) { s_prefix = 0u; s_mask = 0u; s_k = limit - 1; }
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[threadIdx.x] = 0u;
        __syncthreads();
        const unsigned prefix = s_prefix, mask = s_mask;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int o = it.octave_of(i);
            const InitialExtremum& e = it.at(i, o);
            if (cell_of(e, pyr.oct[o].w, pyr.oct[o].h, fc.grid) != cell) continue;
            const unsigned key = scale_key(e, o, largest_first);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int kk = s_k;
            unsigned d = 0;
            for (; d < 256u; ++d) {
                if (kk < (int)hist[d]) break;
                kk -= (int)hist[d];
            }
            s_k = kk;
            s_prefix = prefix | (d << shift);
            s_mask = mask | (255u << shift);
        }
        __syncthreads();
    }
    // s_prefix is the key of rank limit-1; s_k = how many extrema with exactly that key come before it
    const unsigned cut = s_prefix;
    const int ties_kept = s_k + 1;
    if (threadIdx.x == 0) s_ties = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int o = it.octave_of(i);
        const InitialExtremum& e = it.at(i, o);
        if (cell_of(e, pyr.oct[o].w, pyr.oct[o].h, fc.grid) != cell) continue;
        const unsigned key = scale_key(e, o, largest_first);
        if (key > cut) keep[i] = 0;
        else if (key == cut && atomicAdd(&s_ties, 1) >= ties_kept) keep[i] = 0;
    }
}

__global__ void __launch_bounds__(256)
filter_compact_kernel(PyramidView pyr, Consts k, const InitialExtremum* __restrict__ iext, InitialExtremum* __restrict__ iext_f,
                      Counters* ct, const FilterPlan* __restrict__ plan, const unsigned char* __restrict__ keep)
{
    if (!plan->active) return;
    const Items it = make_items(iext, ct, k, pyr.num_octaves);
    const int total = it.first[kMaxOctaves];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        if (!keep[i]) continue;
        const int o = it.octave_of(i);
        const int idx = atomicAdd(&ct->ext_ct_f[o], 1);
        iext_f[(size_t)o * k.max_extrema + idx] = it.at(i, o);
    }
}

} // namespace

int launch_grid_filter(const PyramidView& pyr, const Consts& k, const FilterCfg& fc, const InitialExtremum* iext,
                       InitialExtremum* iext_f, unsigned char* keep, size_t keep_bytes, FilterPlan* plan, Counters* ct,
                       cudaStream_t st)
{
    if (fc.max_extrema <= 0) return 0;
    cudaMemsetAsync(keep, 1, keep_bytes, st);
    filter_plan_kernel<<<1, kPlanThreads, 0, st>>>(pyr, k, fc, iext, ct, plan);
    filter_select_kernel<<<fc.grid * fc.grid, 256, 0, st>>>(pyr, k, fc, iext, ct, plan, keep);
    filter_compact_kernel<<<sm_count(), 256, 0, st>>>(pyr, k, iext, iext_f, ct, plan, keep);
    return 3;
}

} // namespace psb
