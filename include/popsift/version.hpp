// Version of the popsift_b200 drop-in library (same macro names as the reference's version.hpp).
#pragma once
#define POPSIFT_VERSION_MAJOR 1
#define POPSIFT_VERSION_MINOR 0
#define POPSIFT_VERSION_PATCH 0
#define POPSIFT_VERSION_STRING "1.0.0-b200"
