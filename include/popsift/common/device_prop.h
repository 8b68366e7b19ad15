// popsift::cuda::device_prop_t -- device enumeration helper kept for source compatibility with
// the reference's popsift-demo (reference src/popsift/common/device_prop.h:24-108).
// This implementation stores pyramids in linear HBM, so the texture/surface limit checks of the
// reference always pass; only set() / print() do real work.
#pragma once
#include <vector>

namespace popsift { namespace cuda {

class device_prop_t
{
    int _num_devices;
public:
    enum { do_warn = true, dont_warn = false };
    device_prop_t();
    ~device_prop_t();
    void print();
    void set(int n, bool print_choice = false);
    bool checkLimit_2DtexLinear(int& width, int& height, bool printWarn) const;
    bool checkLimit_2DtexArray(int& width, int& height, bool printWarn) const;
    bool checkLimit_2DtexLayered(int& width, int& height, int& layers, bool printWarn) const;
    bool checkLimit_2DsurfLayered(int& width, int& height, int& layers, bool printWarn) const;
};

}} // namespace popsift::cuda
