// popsift::Feature / FeaturesHost -- containers and the text format of the reference
// (reference src/popsift/features.cu:25-110,310-338): one line per (feature, orientation):
//   x y 1/sigma^2 0 1/sigma^2 d0 .. d127
#include "popsift/features.h"

#include <cmath>
#include <cstdlib>
#include <iomanip>
#include <new>
#include <stdexcept>

namespace popsift {

FeaturesBase::FeaturesBase() : _num_ext(0), _num_ori(0) {}
FeaturesBase::~FeaturesBase() = default;

FeaturesHost::FeaturesHost() : _ext(nullptr), _ori(nullptr) {}
FeaturesHost::FeaturesHost(int num_ext, int num_ori) : _ext(nullptr), _ori(nullptr) { reset(num_ext, num_ori); }

FeaturesHost::~FeaturesHost()
{
    std::free(_ext);
    std::free(_ori);
}

void FeaturesHost::reset(int num_ext, int num_ori)
{
    std::free(_ext); _ext = nullptr;
    std::free(_ori); _ori = nullptr;
    // page-aligned like the reference's arrays (features.cu:63-84)
    const size_t fb = ((size_t)(num_ext > 0 ? num_ext : 1) * sizeof(Feature) + 4095) / 4096 * 4096;
    const size_t db = ((size_t)(num_ori > 0 ? num_ori : 1) * sizeof(Descriptor) + 4095) / 4096 * 4096;
    _ext = static_cast<Feature*>(std::aligned_alloc(4096, fb));
    _ori = static_cast<Descriptor*>(std::aligned_alloc(4096, db));
    if (!_ext || !_ori) throw std::runtime_error("Runtime error:\n    Failed to (re)allocate memory for downloading features");
    setFeatureCount(num_ext);
    setDescriptorCount(num_ori);
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
