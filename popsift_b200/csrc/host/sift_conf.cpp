// popsift::Config -- behaviour per reference src/popsift/sift_conf.cu:18-307 (defaults, string
// parsers, error texts, equality), written from scratch and free of any CUDA call.
#include "popsift/sift_conf.h"
#include "popsift_b200.h"

#include <map>
#include <stdexcept>

namespace popsift {

namespace {
[[noreturn]] void fatal(const std::string& s) { throw std::runtime_error(s); }
}

Config::Config()
    : octaves(-1), levels(3), sigma(1.6f), _edge_limit(10.0f), verbose(false)
    , _threshold(0.04f), _upscale_factor(1.0f), _log_mode(None), _scaling_mode(ScaleDefault)
    , _desc_mode(Loop), _grid_filter_mode(RandomScale), _max_extrema(100000), _filter_max_extrema(-1)
    , _filter_grid_size(2), _gauss_mode(getGaussModeDefault()), _sift_mode(PopSift)
    , _assume_initial_blur(true), _initial_blur(0.5f), _normalization_mode(getNormModeDefault())
    , _normalization_multiplier(0), _print_gauss_tables(false)
{
}

void Config::setMode(SiftMode m) { _sift_mode = m; }
void Config::setGaussMode(GaussMode m) { _gauss_mode = m; }
void Config::setDescMode(DescMode m) { _desc_mode = m; }
void Config::setLogMode(LogMode m) { _log_mode = m; }
void Config::setScalingMode(ScalingMode m) { _scaling_mode = m; }
void Config::setVerbose(bool on) { verbose = on; }
void Config::setFilterSorting(GridFilterMode m) { _grid_filter_mode = m; }
void Config::setNormMode(NormMode m) { _normalization_mode = m; }
void Config::setUseRootSift(bool on) { _normalization_mode = on ? RootSift : Classic; }
void Config::setNormalizationMultiplier(int mul) { _normalization_multiplier = mul; }
void Config::setDownsampling(float v) { _upscale_factor = -v; }
void Config::setOctaves(int v) { octaves = v; }
void Config::setLevels(int v) { levels = v; }
void Config::setSigma(float v) { sigma = v; }
void Config::setEdgeLimit(float v) { _edge_limit = v; }
void Config::setThreshold(float v) { _threshold = v; }
void Config::setPrintGaussTables() { _print_gauss_tables = true; }
void Config::setFilterMaxExtrema(int e) { _filter_max_extrema = e; }
void Config::setFilterGridSize(int s) { _filter_grid_size = s; }

void Config::setInitialBlur(float blur)
{
    _assume_initial_blur = (blur != 0.0f);
    _initial_blur = blur;
}

void Config::setDescMode(const std::string& name)
{
    static const std::map<std::string, DescMode> t = {
        {"loop", Loop}, {"iloop", ILoop}, {"grid", Grid}, {"igrid", IGrid}, {"notile", NoTile}};
    auto it = t.find(name);
    if (it == t.end()) fatal("specified descriptor extraction mode must be one of loop, grid or igrid");
    _desc_mode = it->second;
}

void Config::setGaussMode(const std::string& name)
{
    static const std::map<std::string, GaussMode> t = {
        {"vlfeat", VLFeat_Compute}, {"vlfeat-hw-interpolated", VLFeat_Relative}, {"relative", VLFeat_Relative},
        {"vlfeat-direct", VLFeat_Relative_All}, {"opencv", OpenCV_Compute}, {"fixed9", Fixed9}, {"fixed15", Fixed15}};
    auto it = t.find(name);
    if (it == t.end()) fatal(std::string("Bad Gauss mode.\n") + getGaussModeUsage());
    _gauss_mode = it->second;
}

void Config::setFilterSorting(const std::string& dir)
{
    if (dir == "up") _grid_filter_mode = SmallestScaleFirst;
    else if (dir == "down") _grid_filter_mode = LargestScaleFirst;
    else if (dir == "random") _grid_filter_mode = RandomScale;
    else fatal("filter sorting mode must be one of up, down or random");
}

void Config::setNormMode(const std::string& m)
{
    if (m == "RootSift") _normalization_mode = RootSift;
    else if (m == "classic") _normalization_mode = Classic;
    else fatal(std::string("Bad Normalization mode.\n") + getGaussModeUsage());
}

Config::GaussMode Config::getGaussModeDefault() { return VLFeat_Compute; }
Config::NormMode  Config::getNormModeDefault() { return RootSift; }

const char* Config::getGaussModeUsage()
{
    return "Choice of Gauss filter method. Options are: vlfeat (default), vlfeat-hw-interpolated, "
           "vlfeat-direct, opencv, fixed9, fixed15, relative (synonym for vlfeat-hw-interpolated)";
}

const char* Config::getNormModeUsage()
{
    return "Choice of descriptor normalization modes. Options are: RootSift (L1-like, default), Classic (L2-like)";
}

bool  Config::getCanFilterExtrema() const { return true; }    // reference: !POPSIFT_DISABLE_GRID_FILTER (sift_conf.cu:257-264)
bool  Config::hasInitialBlur() const { return _assume_initial_blur; }
float Config::getInitialBlur() const { return _initial_blur; }
float Config::getPeakThreshold() const { return _threshold * 0.5f * 255.0f / levels; }
bool  Config::ifPrintGaussTables() const { return _print_gauss_tables; }
Config::GaussMode Config::getGaussMode() const { return _gauss_mode; }
Config::SiftMode  Config::getSiftMode() const { return _sift_mode; }
Config::LogMode   Config::getLogMode() const { return _log_mode; }
bool  Config::getUseRootSift() const { return _normalization_mode == RootSift; }
Config::NormMode  Config::getNormMode(NormMode) const { return _normalization_mode; }
int   Config::getNormalizationMultiplier() const { return _normalization_multiplier; }

bool Config::equal(const Config& o) const
{
    // the reference compares exactly these fields (sift_conf.cu:286-304): desc mode, filter settings,
    // log mode and verbosity do not take part
    return octaves == o.octaves && levels == o.levels && sigma == o.sigma && _edge_limit == o._edge_limit &&
           _threshold == o._threshold && _upscale_factor == o._upscale_factor && _scaling_mode == o._scaling_mode &&
           _max_extrema == o._max_extrema && _gauss_mode == o._gauss_mode && _sift_mode == o._sift_mode &&
           _assume_initial_blur == o._assume_initial_blur && _initial_blur == o._initial_blur &&
           _normalization_mode == o._normalization_mode && _normalization_multiplier == o._normalization_multiplier;
}

void Config::toC(ps_config& c) const
{
    c.octaves = octaves;
    c.levels = levels;
    c.sigma = sigma;
    c.edge_limit = _edge_limit;
    c.threshold = _threshold;
    c.upscale = _upscale_factor;
    c.initial_blur = _initial_blur;
    c.has_initial_blur = _assume_initial_blur ? 1 : 0;
    c.sift_mode = (int)_sift_mode;
    c.gauss_mode = (int)_gauss_mode;
    c.desc_mode = (int)_desc_mode;
    c.norm_mode = (int)_normalization_mode;
    c.norm_multi = _normalization_multiplier;
    c.max_extrema = _max_extrema;
    c.scaling_mode = _scaling_mode == ScaleDirect ? PS_SCALE_DIRECT : PS_SCALE_DEFAULT;
    c.filter_max_extrema = _filter_max_extrema;
    c.filter_grid_size = _filter_grid_size;
    c.filter_sort = _grid_filter_mode == LargestScaleFirst ? PS_FILTER_LARGEST_FIRST
                  : _grid_filter_mode == SmallestScaleFirst ? PS_FILTER_SMALLEST_FIRST : PS_FILTER_RANDOM;
}

} // namespace popsift
