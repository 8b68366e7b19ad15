#pragma once
#include <string>
/// Reads P2/P3/P5/P6 netpbm files into an 8-bit gray buffer (new[]-allocated, caller deletes).
/// Same conversions as the reference reader (reference src/application/pgmread.cpp:38-258):
/// ASCII and 16-bit gray values scaled by 255/maxval, colour -> gray with the integer weights
/// (4899 R + 9617 G + 1868 B) >> 14 the reference is compiled with.
unsigned char* readPGMfile(const std::string& filename, int& w, int& h);
