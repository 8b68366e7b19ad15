#pragma once
#include <string>
/// Reads P2/P3/P5/P6 netpbm files into an 8-bit gray buffer (new[]-allocated, caller deletes).
/// Same conversions as the reference reader (reference src/application/pgmread.cpp:38-258):
/// values scaled by 255/maxval, colour -> gray with 0.298912 R + 0.586611 G + 0.114478 B.
unsigned char* readPGMfile(const std::string& filename, int& w, int& h);
