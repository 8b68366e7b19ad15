// Internal (device + host) structures shared by the kernels and the C-ABI layer.
#pragma once
#include "popsift_b200.h"

#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>

namespace psb {

constexpr int kMaxOctaves = PS_MAX_OCTAVES;
constexpr int kMaxPlanes  = PS_GAUSS_LEVELS + 3;   // L+3 <= 15
constexpr int kOriBins    = 36;                    // reference sift_constants.h:41
constexpr int kMaxLevels  = PS_GAUSS_LEVELS;

// A candidate from the DoG scan (reference InitialExtremum, sift_extremum.h:25-39, without
// the grid-filter fields).
struct InitialExtremum {
    float xpos, ypos;
    int   lpos;
    float sigma;
};

// Per-image device counters (reference ExtremaCounters, sift_pyramid.h:21-35).
struct Counters {
    int ext_ct[kMaxOctaves];   // extrema found per octave (may exceed max_extrema; clamp on use)
    int ext_total;             // after clamping, written by the orientation-prefix kernel
    int ori_total;
    int overflow;              // bit 0: extrema capacity hit, bit 1: descriptor capacity hit
    int work_ori;              // work-stealing cursors
    int work_desc;
    int filtered;              // the grid filter fired: the orientation stage reads ext_ct_f / the filtered array
    int ext_ct_f[kMaxOctaves]; // extrema per octave that survived the grid filter
    int ori_needed;            // descriptors the image has (may exceed desc_capacity: the host then grows the buffers)
    int pad_[1];
};

// grid filter (k_filter.cu; reference s_filtergrid.cu:112-325)
struct FilterCfg { int max_extrema, grid, sort; };
struct FilterPlan {
    int active, limit, total, pad_;
    int cell_count[PS_MAX_FILTER_GRID * PS_MAX_FILTER_GRID];
};

// One octave's planes in HBM: linear float32, row pitch a multiple of 32 floats (128 B).
struct OctaveView {
    float* gauss;       // (levels+3) planes, plane stride = plane
    float* dog;         // (levels+2) planes
    // candidate lists of the `levels` scanned DoG planes (see CandSink): level q (DoG plane q+1), block b of
    // the level kernel owns entries [(q*cand_blocks + b) * cand_region, +cand_cnt[q*cand_blocks + b])
    uint32_t* cand;
    int*   cand_cnt;
    int    cand_blocks, cand_region;
    int    w, h;
    int    pitch;       // floats per row
    size_t plane;       // floats per plane = pitch * h
};

// Kernel-visible description of one slot's pyramid.
struct PyramidView {
    OctaveView oct[kMaxOctaves];
    int        num_octaves;
    int        levels;      // L
    int        cands_filled;   // the pyramid kernels appended the candidate lists of this image
    // the block counts of all octaves (oct[o].cand_cnt, contiguous in octave order), their number and the
    // array their exclusive prefix sum is written to (cand_regions + 1 entries)
    int*       cand_cnt_all;
    int*       cand_prefix;
    int        cand_regions;
};

// Where a pyramid kernel reports the DoG pixel pairs (x even) in which a sample passes the peak
// threshold while it writes the DoG plane: the extrema stage then only visits those (a few per thousand
// pixels) instead of scanning every plane again.  Each block of the kernel owns a private region of the
// list and a count (zeroed per image), so the kernel needs a shared-memory counter but no global atomic.
struct CandSink {
    uint32_t* list;     // blocks x region entries: x | y << 16 of the pair's left pixel
    int*      counts;   // blocks
    float     thr;
    int       region;   // entries per block
};

// Small constants consumed by the extrema / orientation / descriptor kernels
// (reference ConstInfo, sift_constants.h:56-67).
struct Consts {
    float sigma0;
    float sigma_k;
    float edge_limit;
    float threshold;        // peak threshold
    int   max_extrema;      // per octave
    int   ext_capacity;     // total Extremum records per slot
    int   desc_capacity;    // total descriptors per slot
    int   norm_multi;
    int   norm_mode;
    int   sift_mode;
    int   up_fac;           // int(upscale), reference sift_pyramid.cu:297
};

// |DoG| below this can never be an extremum (reference s_extrema.cu:425-440: the three modes' pre-test)
#ifdef __CUDACC__
__host__ __device__
#endif
inline float extrema_threshold(const Consts& k)
{
    // float(0.8) * 2 == float(1.6): the VLFeat and PopSift expressions agree bit for bit
    return k.sift_mode == PS_MODE_OPENCV ? floorf(k.threshold) : 1.6f * k.threshold;
}

// multiprocessors of the CURRENT device (cudaDeviceProp::multiProcessorCount, cached per device; 148 on B200):
// every fixed grid and the one-wave partition of the pyramid kernels are sized from it
int sm_count();

// ---- launchers (defined in the k_*.cu files); all asynchronous on `st`, return #kernels launched

struct GaussRow { float tap[PS_GAUSS_ALIGN]; int span; };

int level0_plan_for(int w, int h, int W, int H, float shift, int R);   // k_pyramid.cu: LEVEL0_* choice for one geometry
// octave 0, level 0 from the 8-bit or float input image
// (octave > 0: Config::ScaleDirect, the same pass onto a smaller octave with that octave's dd row and shift 0.5)
int launch_level0_u8(const uint8_t* img, size_t img_pitch, int w, int h, float upscale, int sift_mode,
                     const OctaveView& o0, const GaussRow& dd, const GaussRow& inc0, cudaStream_t st, int octave = 0);
int launch_level0_f32(const float* img, size_t img_pitch_floats, int w, int h, float upscale, int sift_mode,
                      const OctaveView& o0, const GaussRow& dd, const GaussRow& inc0, cudaStream_t st, int octave = 0);
// --gauss-mode vlfeat-direct (VLFeat_Relative_All): one level of octave 0 straight from the input image; the DoG planes and
// the next octave's level 0 then come from their own small kernels
int launch_level0_abs_u8(const uint8_t* img, size_t img_pitch, int w, int h, float upscale, int sift_mode,
                         const OctaveView& o0, int level, const GaussRow& taps, cudaStream_t st);
int launch_level0_abs_f32(const float* img, size_t img_pitch_floats, int w, int h, float upscale, int sift_mode,
                          const OctaveView& o0, int level, const GaussRow& taps, cudaStream_t st);
int launch_dog_planes(const OctaveView& o, int nplanes, cudaStream_t st);
// --gauss-mode relative (VLFeat_Relative): the first horizontal pass alone, and one interpolated pass over a plane
int launch_level0_rows_u8(const uint8_t* img, size_t img_pitch, int w, int h, float upscale, int sift_mode,
                          const OctaveView& o0, float* dst, const GaussRow& dd, cudaStream_t st, int octave = 0);
int launch_level0_rows_f32(const float* img, size_t img_pitch_floats, int w, int h, float upscale, int sift_mode,
                           const OctaveView& o0, float* dst, const GaussRow& dd, cudaStream_t st, int octave = 0);
// --gauss-mode fixed9 / fixed15 (s_pyramid_fixed.cu): one level of octave 0 from the input image / of an octave >= 1 from its level 0
int launch_fixed_level0_u8(const uint8_t* img, size_t img_pitch, int w, int h, float upscale, const OctaveView& o0, int level,
                           const GaussRow& taps, int S, float* scratch, cudaStream_t st);
int launch_fixed_level0_f32(const float* img, size_t img_pitch_floats, int w, int h, float upscale, const OctaveView& o0, int level,
                            const GaussRow& taps, int S, float* scratch, cudaStream_t st);
int launch_fixed_levelN(const OctaveView& o, int level, const GaussRow& taps, int S, float* scratch, cudaStream_t st);
int launch_interp_pass(const float* src, float* dst, int W, int H, int pitch, const GaussRow& f, int ispan, int along_x, cudaStream_t st);
int launch_decimate(const OctaveView& prev, int level, const OctaveView& next, cudaStream_t st);
// level l >= 1 of one octave: blur level l-1 -> level l, DoG[l-1] = G[l]-G[l-1]; if next0 != nullptr
// also writes every second pixel into the next octave's level 0.
// `sink` (optional) receives the threshold-passing samples of the DoG plane; only honoured when
// blur_level_collects(g) is true.
int launch_blur_level(const OctaveView& o, int level, const GaussRow& g, const OctaveView* next, const CandSink* sink,
                      cudaStream_t st);
bool blur_level_collects(const GaussRow& g);
// layout of the candidate lists of one w x h plane (k_partition.h): blocks, entries per block, and a
// monotonic upper bound of blocks * entries (memory budget)
int cand_blocks_for(int w, int h);
int cand_region_for(int w, int h);
long long cand_entry_bound_for(int w, int h);

int launch_find_extrema(const PyramidView& pyr, const Consts& k, InitialExtremum* iext, Counters* ct, cudaStream_t st);
// grid filter between the extrema and the orientation stage; no-op (returns 0) when fc.max_extrema <= 0.
// keep: one byte per initial extremum (num_octaves * max_extrema); iext_f: a second InitialExtremum array
int launch_grid_filter(const PyramidView& pyr, const Consts& k, const FilterCfg& fc, const InitialExtremum* iext,
                       InitialExtremum* iext_f, unsigned char* keep, size_t keep_bytes, FilterPlan* plan, Counters* ct,
                       cudaStream_t st);
// slice_sum: ext_capacity / PS_ORI_SLICE + 1 ints, zeroed per image
#define PS_ORI_SLICE 256
// iext_f: the grid-filtered extrema (read instead of iext when ct->filtered); may be nullptr when the filter is off
int launch_orientation(const PyramidView& pyr, const Consts& k, const InitialExtremum* iext, const InitialExtremum* iext_f, ps_extremum* ext,
                       int* feat_to_ext, int* slice_sum, Counters* ct, cudaStream_t st);
int launch_descriptors(const PyramidView& pyr, const Consts& k, const ps_extremum* ext, const int* feat_to_ext,
                       ps_descriptor* desc, Counters* ct, cudaStream_t st);
// --desc-mode grid / igrid / iloop / notile (k_desc_modes.cu; reference s_desc_grid.cu, s_desc_igrid.cu, s_desc_iloop.cu,
// s_desc_notile.cu); PS_DESC_LOOP goes through launch_descriptors
int launch_descriptors_mode(int mode, const PyramidView& pyr, const Consts& k, const ps_extremum* ext, const int* feat_to_ext,
                            ps_descriptor* desc, Counters* ct, cudaStream_t st);
int launch_prep_features(const Consts& k, const ps_extremum* ext, ps_feature* feat, const Counters* ct, cudaStream_t st);
// device copy of the Feature records: desc[] become pointers into `desc` (first index is in pad_)
int launch_fix_feature_pointers(ps_feature* feat, ps_descriptor* desc, int n, cudaStream_t st);

// brute-force 2-NN matcher (k_match.cu; reference features.cu:165-304): out = n_left x (best, second, accept), device
// memory; asynchronous on `st`; returns the number of kernels launched or -1 (*err set)
int run_match(const ps_descriptor* l, int nl, const ps_descriptor* r, int nr, int32_t* out, int flags, cudaStream_t st, const char** err);

} // namespace psb
