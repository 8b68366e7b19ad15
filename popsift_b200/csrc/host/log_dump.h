// Config::LogMode::All: the reference's debugging files (see log_dump.cpp).
#pragma once
#include "popsift/features.h"
#include "popsift_b200.h"

namespace popsift { namespace detail {

// every Gaussian and DoG plane of the slot's last image, into the current directory (reference sift_octave.cu:111-188)
bool dump_slot_planes(ps_ctx* ctx, int slot, int levels, const char* basename);
// dir-desc/ and dir-fpt/ (reference sift_pyramid.cu:88-106,401-444)
void dump_descriptors(const FeaturesHost& f, float up_fac, const char* basename);

}} // namespace popsift::detail
