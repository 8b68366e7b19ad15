// popsift::cuda::device_prop_t -- minimal device enumeration for popsift-demo's --print-dev-info.
#include "popsift/common/device_prop.h"

#include <cuda_runtime.h>
#include <iostream>
#include <stdexcept>

namespace popsift { namespace cuda {

device_prop_t::device_prop_t() : _num_devices(0)
{
    if (cudaGetDeviceCount(&_num_devices) != cudaSuccess) _num_devices = 0;
}
device_prop_t::~device_prop_t() = default;

void device_prop_t::print()
{
    for (int n = 0; n < _num_devices; n++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, n) != cudaSuccess) continue;
        std::cout << "Device information for device " << n << std::endl
                  << "    Name: " << p.name << std::endl
                  << "    Compute Capability:    " << p.major << "." << p.minor << std::endl
                  << "    Total device mem:      " << p.totalGlobalMem << " B" << std::endl
                  << "    Multiprocessors:       " << p.multiProcessorCount << std::endl
                  << "    Shared mem per block:  " << p.sharedMemPerBlockOptin << " B (opt-in)" << std::endl
                  << "    L2 cache:              " << p.l2CacheSize << " B" << std::endl;
    }
}

void device_prop_t::set(int n, bool print_choice)
{
    if (n >= _num_devices) throw std::runtime_error("Error: Choosing a CUDA device that does not exist");
    if (cudaSetDevice(n) != cudaSuccess) throw std::runtime_error("Error: Cannot set CUDA device");
    if (print_choice) std::cout << "Choosing device " << n << std::endl;
}

// pyramids live in linear HBM here: none of the reference's texture / surface limits apply
bool device_prop_t::checkLimit_2DtexLinear(int&, int&, bool) const { return true; }
bool device_prop_t::checkLimit_2DtexArray(int&, int&, bool) const { return true; }
bool device_prop_t::checkLimit_2DtexLayered(int&, int&, int&, bool) const { return true; }
bool device_prop_t::checkLimit_2DsurfLayered(int&, int&, int&, bool) const { return true; }

}} // namespace popsift::cuda
