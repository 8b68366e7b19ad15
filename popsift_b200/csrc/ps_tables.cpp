// Host-side tables and geometry of the SIFT pipeline (no CUDA needed).
//
// Restates, with bit-identical float/double mixing, what the reference computes in
//   init_filter / GaussTable::computeBlurTable   (reference src/popsift/gauss_filter.cu:127-257,341-371)
//   init_constants                               (reference src/popsift/sift_constants.cu:22-53)
//   Config::getPeakThreshold                     (reference src/popsift/sift_conf.cu:276-279)
//   PopSift::private_apply_scale_factor          (reference src/popsift/popsift.cpp:109-126)
//   Pyramid ctor octave sizes                    (reference src/popsift/sift_pyramid.cu:129-134)
#include "popsift_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

namespace {

// Half-width (including the centre tap) of the kernel for one sigma: VLFeat's rule (reference gauss_filter.cu:301-307)
// or OpenCV's (--gauss-mode opencv, gauss_filter.cu:320-327).  Everything else about the two modes is identical.
// mode: 0 = VLFeat rule, 1 = OpenCV rule, 2 = VLFeat rule rounded up to the next odd span (--gauss-mode relative /
// vlfeat-hw-interpolated: the taps are consumed in pairs, gauss_filter.cu:309-319)
int kernel_span(float sigma, int mode)
{
    if (mode == 3) return 5;     // --gauss-mode fixed9  (gauss_filter.cu:289-290)
    if (mode == 4) return 8;     // --gauss-mode fixed15 (gauss_filter.cu:291-292)
    const bool opencv = mode == 1;
    if (opencv) {
        int span = static_cast<int>(std::roundf(2.0f * 4.0f * sigma + 1.0f)) | 1;
        span >>= 1;
        span += 1;
        return std::min(span, PS_GAUSS_ALIGN - 1);
    }
    int s = std::min(static_cast<int>(std::ceil(4.0f * sigma) + 1.0f), PS_GAUSS_ALIGN - 1);
    if (mode == 2 && (s & 1) == 0) s += 1;
    return std::min(s, PS_GAUSS_ALIGN - 1);
}

// One normalised half-kernel.  The taps are evaluated in double, stored as float; the
// normaliser is a double that accumulates twice the *stored float* tap.
void fill_kernel(float sigma, float* taps, int32_t* span_out, int mode)
{
    const int span = kernel_span(sigma, mode);
    std::fill(taps, taps + PS_GAUSS_ALIGN, 0.0f);
    taps[0] = 1.0f;
    double norm = 1.0;
    for (int k = 1; k < span; ++k) {
        const double q = static_cast<double>(k) / static_cast<double>(sigma);
        const float t = static_cast<float>(std::exp(-0.5 * std::pow(q, 2.0)));
        taps[k] = t;
        norm += static_cast<double>(2.0f * t);
    }
    for (int k = 0; k < span; ++k) taps[k] = static_cast<float>(static_cast<double>(taps[k]) / norm);
    *span_out = span;
}

} // namespace

extern "C" int ps_abi_version(void) { return PS_ABI_VERSION; }

extern "C" void ps_config_default(ps_config* c)
{
    if (!c) return;
    c->octaves = -1;
    c->levels = 3;
    c->sigma = 1.6f;
    c->edge_limit = 10.0f;
    c->threshold = 0.04f;
    c->upscale = 1.0f;
    c->initial_blur = 0.5f;
    c->has_initial_blur = 1;
    c->sift_mode = PS_MODE_POPSIFT;
    c->gauss_mode = PS_GAUSS_VLFEAT_COMPUTE;
    c->desc_mode = PS_DESC_LOOP;
    c->norm_mode = PS_NORM_ROOTSIFT;
    c->norm_multi = 0;
    c->max_extrema = 100000;
    c->scaling_mode = PS_SCALE_DEFAULT;
    c->filter_max_extrema = -1;
    c->filter_grid_size = 2;
    c->filter_sort = PS_FILTER_RANDOM;
}

extern "C" int ps_gauss_tables_compute(const ps_config* cfg, ps_gauss_tables* out)
{
    if (!cfg || !out) return PS_ERR_ARG;
    std::memset(out, 0, sizeof(*out));
    const int levels = std::max(2, cfg->levels);
    const float s0 = cfg->sigma;
    if (s0 > 2.0f) return PS_ERR_ARG;                 // reference gauss_filter.cu:131-137
    if (levels > PS_GAUSS_LEVELS) return PS_ERR_ARG;  // reference gauss_filter.cu:138-144
    if (cfg->gauss_mode != PS_GAUSS_VLFEAT_COMPUTE && cfg->gauss_mode != PS_GAUSS_OPENCV_COMPUTE &&
        cfg->gauss_mode != PS_GAUSS_VLFEAT_RELATIVE_ALL && cfg->gauss_mode != PS_GAUSS_VLFEAT_RELATIVE &&
        cfg->gauss_mode != PS_GAUSS_FIXED9 && cfg->gauss_mode != PS_GAUSS_FIXED15) return PS_ERR_ARG;
    // span rule (gauss_filter.cu:274-296): RELATIVE_ALL uses the vlfeat rule, RELATIVE the odd one, FIXED9 / FIXED15 constants
    const int ocv = cfg->gauss_mode == PS_GAUSS_OPENCV_COMPUTE ? 1 : cfg->gauss_mode == PS_GAUSS_VLFEAT_RELATIVE ? 2 :
                    cfg->gauss_mode == PS_GAUSS_FIXED9 ? 3 : cfg->gauss_mode == PS_GAUSS_FIXED15 ? 4 : 0;
    const int planes = levels + 3;
    const float blur_in = cfg->has_initial_blur ? cfg->initial_blur * std::pow(2.0f, cfg->upscale) : 0.0f;

    // incremental sigmas: level 0 closes the gap from the assumed input blur to sigma0,
    // level l >= 1 goes from sigma0*2^((l-1)/L) to sigma0*2^(l/L); all in float.
    out->inc_sigma[0] = cfg->has_initial_blur ? std::sqrt(std::fabs(s0 * s0 - blur_in * blur_in)) : s0;
    for (int l = 1; l < planes; ++l) {
        const float lo = s0 * std::pow(2.0f, static_cast<float>(l - 1) / static_cast<float>(levels));
        const float hi = s0 * std::pow(2.0f, static_cast<float>(l) / static_cast<float>(levels));
        out->inc_sigma[l] = std::sqrt(hi * hi - lo * lo);
    }
    for (int l = 0; l < PS_GAUSS_LEVELS; ++l)
        fill_kernel(out->inc_sigma[l], &out->inc_filter[l * PS_GAUSS_ALIGN], &out->inc_span[l], ocv);

    // first horizontal pass over the input image (octave 0): sqrt(|sigma0^2 - blur_in^2|)
    {
        const float so = std::scalbn(s0, 0);
        const float gap = std::sqrt(std::fabs(so * so - blur_in * blur_in));
        out->dd_sigma0 = std::scalbn(gap, 0);
        fill_kernel(out->dd_sigma0, out->dd_filter0, &out->dd_span0, ocv);
    }
    // ... and of every octave, for Config::ScaleDirect (gauss_filter.cu:227-238): sqrt(|(sigma0 2^o)^2 - blur_in^2|) / 2^o
    for (int o = 0; o < PS_MAX_OCTAVES; ++o) {
        const float so = std::scalbn(s0, o);
        const float gap = std::sqrt(std::fabs(so * so - blur_in * blur_in));
        out->dd_sigma[o] = std::scalbn(gap, -o);
        fill_kernel(out->dd_sigma[o], &out->dd_filter[o * PS_GAUSS_ALIGN], &out->dd_span[o], ocv);
    }
    // octave 0 straight from the input image, every level (gauss_filter.cu:190-199): sqrt(|(sigma0 2^(l/L))^2 - blur_in^2|)
    for (int l = 0; l < planes; ++l) {
        const float ss = s0 * std::pow(2.0f, static_cast<float>(l) / static_cast<float>(levels));
        out->abs_sigma[l] = std::sqrt(std::fabs(ss * ss - blur_in * blur_in));
    }
    for (int l = 0; l < PS_GAUSS_LEVELS; ++l)
        fill_kernel(out->abs_sigma[l], &out->abs_filter[l * PS_GAUSS_ALIGN], &out->abs_span[l], ocv);
    // levels 1.. of any octave straight from its level 0 (gauss_filter.cu:208-214): sqrt((sigma0 2^(l/L))^2 - sigma0^2); used by
    // the fixed-span modes for octaves >= 1
    for (int l = 1; l < planes; ++l) {
        const float ss = s0 * std::pow(2.0f, static_cast<float>(l) / static_cast<float>(levels));
        out->absn_sigma[l] = std::sqrt(ss * ss - s0 * s0);
    }
    for (int l = 0; l < PS_GAUSS_LEVELS; ++l)
        fill_kernel(out->absn_sigma[l], &out->absn_filter[l * PS_GAUSS_ALIGN], &out->absn_span[l], ocv);
    // pairs of taps merged into one linearly interpolated fetch (GaussTable::transformBlurTable, gauss_filter.cu:372-405):
    // odd entries = the fraction u = a / (a + b), even entries = the weight v = a + b of the pair
    for (int l = 0; l < PS_GAUSS_LEVELS; ++l) {
        int spn = out->inc_span[l];
        if (!(spn & 1)) spn += 1;
        out->inc_ispan[l] = spn;
        float* f = &out->inc_filter[l * PS_GAUSS_ALIGN];
        float* g = &out->inc_ifilter[l * PS_GAUSS_ALIGN];
        g[0] = f[0];
        for (int x = 1; x < spn && x + 1 < PS_GAUSS_ALIGN; x += 2) {
            const float a = f[x], b = f[x + 1];
            g[x] = a / (a + b);
            g[x + 1] = a + b;
        }
    }
    out->peak_threshold = cfg->threshold * 0.5f * 255.0f / static_cast<float>(levels);
    out->sigma_k = std::pow(2.0f, 1.0f / static_cast<float>(levels));
    return PS_OK;
}

extern "C" int ps_geometry(const ps_config* cfg, int w, int h, int32_t* W, int32_t* H)
{
    if (!cfg || w <= 0 || h <= 0) return PS_ERR_ARG;
    const float scale = 1.0f / std::pow(2.0f, -cfg->upscale);
    int n = cfg->octaves;
    if (n < 0) {
        const float lg = std::floor(std::log(static_cast<float>(std::min(w, h))) / std::log(2.0f));
        n = std::max(static_cast<int>(lg - 3.0f + scale), 1);
    }
    n = std::min(n, PS_MAX_OCTAVES);
    if (n < 1) return PS_ERR_ARG;
    int ow = static_cast<int>(std::ceil(w * scale));
    int oh = static_cast<int>(std::ceil(h * scale));
    for (int o = 0; o < n; ++o) {
        if (W) W[o] = ow;
        if (H) H[o] = oh;
        ow = static_cast<int>(std::ceil(ow / 2.0f));
        oh = static_cast<int>(std::ceil(oh / 2.0f));
    }
    return n;
}

// The text the reference prints under Config::setPrintGaussTables() / --print-gauss-tables (init_filter's header,
// gauss_filter.cu:146-161, and print_gauss_filter_symbol(10), gauss_filter.cu:24-121), from the tables of this library.
// Returns the length of the text (without the terminating 0); writes at most cap - 1 characters + 0 into buf.
extern "C" int ps_format_gauss_tables(const ps_config* cfg, char* buf, size_t cap)
{
    ps_gauss_tables t;
    if (!cfg || ps_gauss_tables_compute(cfg, &t) != PS_OK) return PS_ERR_ARG;
    std::string out;
    auto add = [&](const char* fmt, ...) {
        char tmp[256];
        va_list ap;
        va_start(ap, fmt);
        std::vsnprintf(tmp, sizeof(tmp), fmt, ap);
        va_end(ap);
        out += tmp;
    };
    const int stages = std::max(2, cfg->levels) + 3, columns = 10;
    add("\nUpscaling factor: %f (i.e. original image is scaled by a factor of %f)\n\nSigma computations\n"
        "    Initial sigma is %f\n    Input blurriness is assumed to be %f (scaled to %f)\n",
        cfg->upscale, std::pow(2.0f, cfg->upscale), cfg->sigma, cfg->initial_blur, cfg->initial_blur * std::pow(2.0f, cfg->upscale));
    auto table = [&](const float* filter, const float* sigma, const int32_t* span, int rows, bool split_sigma) {
        for (int l = 0; l < rows; ++l) {
            if (split_sigma) { add("      %d %d ", l, 2 * span[l] - 1); add("%2.6f: ", sigma[l]); }
            else add("      %d %d %2.6f: ", l, 2 * span[l] - 1, sigma[l]);
            const int m = std::min(span[l], columns);
            for (int x = 0; x < m; ++x) add("%0.8f ", filter[l * PS_GAUSS_ALIGN + x]);
            add(m < span[l] ? "...\n" : "\n");
        }
    };
    add("\nGauss tables\n      level span sigma : center value -> edge value\n    relative sigma\n");
    table(t.inc_filter, t.inc_sigma, t.inc_span, stages, true);
    add("\n\nGauss tables for hardware interpolation\n"
        "      level span sigma : center value -> ( interpolation value, multiplier ) [one edge value] \n");
    table(t.inc_ifilter, t.inc_sigma, t.inc_ispan, stages, true);
    add("\n\nGauss tables\n      level span sigma : center value -> edge value\n"
        "      absolute filters octave 0 (compute level 0, all other levels directly from level 0)\n");
    table(t.abs_filter, t.abs_sigma, t.abs_span, stages, false);
    add("\n      absolute filters other octaves\n      (level 0 via downscaling, all other levels directly from level 0)\n");
    table(t.absn_filter, t.absn_sigma, t.absn_span, stages, false);
    add("\n    level 0-filters for direct downscaling\n");
    table(t.dd_filter, t.dd_sigma, t.dd_span, PS_MAX_OCTAVES, false);
    add("\n");
    if (buf && cap > 0) {
        const size_t n = std::min(out.size(), cap - 1);
        std::memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return (int)out.size();
}
