// Build-time feature macros of the popsift_b200 drop-in (the reference generates this file with
// CMake from cmake/sift_config.h.in; the macro names are part of its installed interface).
#pragma once
#define POPSIFT_IS_DEFINED(F) F() == 1
#define POPSIFT_HAVE_SHFL_DOWN_SYNC() 1
#define POPSIFT_HAVE_NORMF()          0
#define POPSIFT_DISABLE_GRID_FILTER() 1
#define POPSIFT_USE_NVTX()            0
