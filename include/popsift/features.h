// Result containers of the drop-in (reference src/popsift/features.h:23-122,
// src/popsift/sift_extremum.h:69-72).  Byte layouts are the reference's: Descriptor = 128 floats,
// Feature = 72 bytes with host pointers into the owning FeaturesHost's descriptor array.
#pragma once

#include <iostream>
#include <vector>

#define ORIENTATION_MAX_COUNT 4

namespace popsift {

struct Descriptor
{
    float features[128];   // index (iy*4+ix)*8 + bin
};

struct Feature
{
    int         debug_octave;
    float       xpos;
    float       ypos;
    float       sigma;
    int         num_ori;
    float       orientation[ORIENTATION_MAX_COUNT];
    Descriptor* desc[ORIENTATION_MAX_COUNT];

    void print(std::ostream& ostr, bool write_as_uchar) const;
};

std::ostream& operator<<(std::ostream& ostr, const Feature& feature);

class FeaturesBase
{
    int _num_ext;
    int _num_ori;

public:
    FeaturesBase();
    virtual ~FeaturesBase();

    int  size() const { return _num_ext; }
    int  getFeatureCount() const { return _num_ext; }
    int  getDescriptorCount() const { return _num_ori; }
    void setFeatureCount(int n) { _num_ext = n; }
    void setDescriptorCount(int n) { _num_ori = n; }
};

class FeaturesHost : public FeaturesBase
{
    Feature*    _ext;
    Descriptor* _ori;

public:
    FeaturesHost();
    FeaturesHost(int num_ext, int num_ori);
    ~FeaturesHost() override;

    typedef Feature*       F_iterator;
    typedef const Feature* F_const_iterator;

    F_iterator       begin() { return _ext; }
    F_const_iterator begin() const { return _ext; }
    F_iterator       end() { return _ext + size(); }
    F_const_iterator end() const { return _ext + size(); }

    void reset(int num_ext, int num_ori);
    void pin();     // no-op here: results arrive through the library's own pinned staging
    void unpin();

    Feature*    getFeatures() { return _ext; }
    Descriptor* getDescriptors() { return _ori; }

    void print(std::ostream& ostr, bool write_as_uchar) const;
};

using Features = FeaturesHost;

std::ostream& operator<<(std::ostream& ostr, const FeaturesHost& feature);

/// Device-resident results (Config::MatchingMode; reference features.h:104-122): Feature records,
/// descriptors and the descriptor -> feature reverse map live in device memory owned by this object;
/// Feature::desc[] are device pointers into getDescriptors().  match() is the reference's brute-force
/// 2-nearest-neighbour matcher (features.cu:282-304), computed on the tensor cores (C ABI ps_match).
class FeaturesDev : public FeaturesBase
{
    Feature*    _ext;
    Descriptor* _ori;
    int*        _rev;   // the reverse map from descriptors to extrema

public:
    FeaturesDev();
    FeaturesDev(int num_ext, int num_ori);
    ~FeaturesDev() override;

    void reset(int num_ext, int num_ori);

    /// reference behaviour: prints one "accept/reject feat ..." line per descriptor of *this to stdout
    void match(FeaturesDev* other);
    /// the same search without the printing: 3 ints (best index, second index, accept) per descriptor of *this
    /// (an addition to the reference's interface; flags as for ps_match)
    std::vector<int> matchIndices(FeaturesDev* other, int flags = 0);

    Feature*    getFeatures() { return _ext; }
    Descriptor* getDescriptors() { return _ori; }
    int*        getReverseMap() { return _rev; }
};

} // namespace popsift
