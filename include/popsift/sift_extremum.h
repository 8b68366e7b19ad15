// popsift::Descriptor lives in features.h in this implementation; this header keeps the
// reference's include path (<popsift/sift_extremum.h>) working.
#pragma once
#include "features.h"
