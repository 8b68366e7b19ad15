// popsift::Feature / FeaturesHost -- containers and the text format of the reference
// (reference src/popsift/features.cu:25-110,310-338): one line per (feature, orientation):
//   x y 1/sigma^2 0 1/sigma^2 d0 .. d127
#include "popsift/features.h"
#include "pinned_pool.h"
#include "popsift_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>
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

namespace {

// squared distance in the reference's float32 evaluation order (features.cu:165-185): lane k of a warp holds floats
// 4k..4k+3, computes x*x + y*y + z*z + w*w (SASS: FMUL, FFMA, FFMA, FFMA), then a shuffle-down tree 16, 8, 4, 2, 1
float warp_order_sq_dist(const Descriptor& l, const Descriptor& r)
{
    float lane[32];
    for (int k = 0; k < 32; ++k) {
        const float dx = l.features[4 * k] - r.features[4 * k], dy = l.features[4 * k + 1] - r.features[4 * k + 1];
        const float dz = l.features[4 * k + 2] - r.features[4 * k + 2], dw = l.features[4 * k + 3] - r.features[4 * k + 3];
        float v = dx * dx;
        v = std::fmaf(dy, dy, v);
        v = std::fmaf(dz, dz, v);
        v = std::fmaf(dw, dw, v);
        lane[k] = v;
    }
    for (int s = 16; s > 0; s >>= 1)
        for (int k = 0; k < s; ++k) lane[k] = lane[k] + lane[k + s];
    return lane[0];
}

} // namespace

// reference features.cu:282-304: 2-nearest-neighbour search of every descriptor of *this among `other`'s, then one
// line per descriptor on stdout (show_distance, features.cu:229-277: a device printf there, a host printf here).
// The search runs on the device (ps_match: tcgen05 tensor-core pass + exact re-rank, or the CUDA-core kernel for small sets).
void FeaturesDev::match(FeaturesDev* other)
{
    if (!other) throw std::runtime_error("FeaturesDev::match: null argument");
    const int l_len = getDescriptorCount(), r_len = other->getDescriptorCount();
    if (l_len <= 0 || r_len <= 0) return;
    const std::vector<int> m = matchIndices(other);
    std::vector<Descriptor> lo((size_t)l_len), ro((size_t)r_len);
    std::vector<int> lrev((size_t)l_len), rrev((size_t)r_len);
    if (ps_dev_to_host(lo.data(), _ori, lo.size() * sizeof(Descriptor)) != PS_OK ||
        ps_dev_to_host(ro.data(), other->_ori, ro.size() * sizeof(Descriptor)) != PS_OK ||
        ps_dev_to_host(lrev.data(), _rev, lrev.size() * sizeof(int)) != PS_OK ||
        ps_dev_to_host(rrev.data(), other->_rev, rrev.size() * sizeof(int)) != PS_OK)
        throw std::runtime_error("FeaturesDev::match: device -> host copy failed");
    for (int i = 0; i < l_len; ++i) {
        const int b1 = m[3 * (size_t)i], b2 = m[3 * (size_t)i + 1];
        const float d1 = warp_order_sq_dist(lo[i], ro[b1]), d2 = warp_order_sq_dist(lo[i], ro[b2]);
        std::printf("%s feat %4d [%4d] matches feat %4d [%4d] ( 2nd feat %4d [%4d] ) dist %.3f vs %.3f\n",
                    m[3 * (size_t)i + 2] ? "accept" : "reject", lrev[i], i, rrev[b1], b1, rrev[b2], b2, d1, d2);
    }
    std::fflush(stdout);
}

// the match matrix of the reference's compute_distance as a host vector: 3 ints (best, second, accept) per descriptor
std::vector<int> FeaturesDev::matchIndices(FeaturesDev* other, int flags)
{
    if (!other) throw std::runtime_error("FeaturesDev::matchIndices: null argument");
    const int l_len = getDescriptorCount(), r_len = other->getDescriptorCount();
    std::vector<int> m((size_t)std::max(l_len, 0) * 3, 0);
    if (l_len <= 0) return m;
    int32_t* d_out = static_cast<int32_t*>(ps_dev_alloc(m.size() * sizeof(int32_t)));
    if (!d_out) throw std::runtime_error("FeaturesDev::match: failed to allocate device memory");
    const int dev = std::max(0, ps_pointer_device(_ori));
    int rc = ps_match(dev, reinterpret_cast<const ps_descriptor*>(_ori), l_len, reinterpret_cast<const ps_descriptor*>(other->_ori),
                      r_len, d_out, flags);
    if (rc == PS_OK) rc = ps_dev_to_host(m.data(), d_out, m.size() * sizeof(int32_t));
    ps_dev_free(d_out);
    if (rc != PS_OK) throw std::runtime_error(std::string("FeaturesDev::match failed: ") + ps_last_error(nullptr));
    return m;
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
