// popsift::Feature / FeaturesHost -- containers and the text format of the reference
// (reference src/popsift/features.cu:25-110,310-338): one line per (feature, orientation):
//   x y 1/sigma^2 0 1/sigma^2 d0 .. d127
#include "popsift/features.h"
#include "pinned_pool.h"
#include "popsift_b200.h"

#include <cmath>
#include <cstdlib>
#include <iomanip>
#include <map>
#include <mutex>
#include <new>
#include <stdexcept>

namespace popsift {

FeaturesBase::FeaturesBase() : _num_ext(0), _num_ori(0) {}
FeaturesBase::~FeaturesBase() = default;

FeaturesHost::FeaturesHost() : _ext(nullptr), _ori(nullptr) {}
FeaturesHost::FeaturesHost(int num_ext, int num_ori) : _ext(nullptr), _ori(nullptr) { reset(num_ext, num_ori); }

FeaturesHost::~FeaturesHost()
{
    popsift::detail::pinned_pool().put(_ext);
    popsift::detail::pinned_pool().put(_ori);
}

void FeaturesHost::reset(int num_ext, int num_ori)
{
    popsift::detail::pinned_pool().put(_ext); _ext = nullptr;
    popsift::detail::pinned_pool().put(_ori); _ori = nullptr;
    _ext = static_cast<Feature*>(popsift::detail::pinned_pool().get((size_t)(num_ext > 0 ? num_ext : 1) * sizeof(Feature)));
    _ori = static_cast<Descriptor*>(popsift::detail::pinned_pool().get((size_t)(num_ori > 0 ? num_ori : 1) * sizeof(Descriptor)));
    if (!_ext || !_ori) throw std::runtime_error("Runtime error:\n    Failed to (re)allocate memory for downloading features");
    setFeatureCount(num_ext);
    setDescriptorCount(num_ori);
}

// ---- FeaturesDev (reference features.cu:127-163): device arrays through the C ABI's allocator

FeaturesDev::FeaturesDev() : _ext(nullptr), _ori(nullptr), _rev(nullptr) {}
FeaturesDev::FeaturesDev(int num_ext, int num_ori) : _ext(nullptr), _ori(nullptr), _rev(nullptr) { reset(num_ext, num_ori); }

FeaturesDev::~FeaturesDev()
{
    ps_dev_free(_ext);
    ps_dev_free(_ori);
    ps_dev_free(_rev);
}

void FeaturesDev::reset(int num_ext, int num_ori)
{
    ps_dev_free(_ext); _ext = nullptr;
    ps_dev_free(_ori); _ori = nullptr;
    ps_dev_free(_rev); _rev = nullptr;
    _ext = static_cast<Feature*>(ps_dev_alloc((size_t)(num_ext > 0 ? num_ext : 1) * sizeof(Feature)));
    _ori = static_cast<Descriptor*>(ps_dev_alloc((size_t)(num_ori > 0 ? num_ori : 1) * sizeof(Descriptor)));
    _rev = static_cast<int*>(ps_dev_alloc((size_t)(num_ori > 0 ? num_ori : 1) * sizeof(int)));
    if (!_ext || !_ori || !_rev) throw std::runtime_error("Runtime error:\n    Failed to allocate device memory for features");
    setFeatureCount(num_ext);
    setDescriptorCount(num_ori);
}

void FeaturesDev::match(FeaturesDev* /*other*/)
{
    throw std::runtime_error("popsift_b200: FeaturesDev::match (brute-force matcher) is not implemented");
}

void FeaturesHost::pin() {}
void FeaturesHost::unpin() {}

void FeaturesHost::print(std::ostream& ostr, bool write_as_uchar) const
{
    for (int i = 0; i < size(); i++) _ext[i].print(ostr, write_as_uchar);
}

std::ostream& operator<<(std::ostream& ostr, const FeaturesHost& f)
{
    f.print(ostr, false);
    return ostr;
}

void Feature::print(std::ostream& ostr, bool write_as_uchar) const
{
    const float sigval = 1.0f / (sigma * sigma);
    for (int o = 0; o < num_ori; o++) {
        ostr << xpos << " " << ypos << " " << sigval << " 0 " << sigval << " ";
        if (write_as_uchar) {
            for (int i = 0; i < 128; i++) ostr << roundf(desc[o]->features[i]) << " ";
        } else {
            ostr << std::setprecision(3);
            for (int i = 0; i < 128; i++) ostr << desc[o]->features[i] << " ";
            ostr << std::setprecision(6);
        }
        ostr << std::endl;
    }
}

std::ostream& operator<<(std::ostream& ostr, const Feature& f)
{
    f.print(ostr, false);
    return ostr;
}

} // namespace popsift
