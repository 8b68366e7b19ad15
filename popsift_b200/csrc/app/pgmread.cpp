#include "pgmread.h"

#include <cctype>
#include <fstream>
#include <iostream>
#include <iterator>
#include <vector>

namespace {

// next whitespace-delimited header token, skipping '#' comments
bool next_token(const std::vector<char>& d, size_t& pos, std::string& tok)
{
    for (;;) {
        while (pos < d.size() && std::isspace((unsigned char)d[pos])) ++pos;
        if (pos < d.size() && d[pos] == '#') { while (pos < d.size() && d[pos] != '\n') ++pos; continue; }
        break;
    }
    if (pos >= d.size()) return false;
    size_t e = pos;
    while (e < d.size() && !std::isspace((unsigned char)d[e])) ++e;
    tok.assign(d.begin() + pos, d.begin() + e);
    pos = e;
    return true;
}

// colour -> gray exactly as the reference is compiled (RGB2GRAY_IN_INT, reference
// src/application/pgmread.cpp:17-28,165-169): OpenCV's 14-bit integer weights, truncating cast
inline unsigned char to_gray(unsigned r, unsigned g, unsigned b)
{
    return (unsigned char)((4899u * r + 9617u * g + 1868u * b) >> 14);
}

} // namespace

unsigned char* readPGMfile(const std::string& filename, int& w, int& h)
{
    std::ifstream in(filename.c_str(), std::ios::binary);
    if (!in.is_open()) { std::cerr << "File " << filename << " could not be opened for reading" << std::endl; return nullptr; }
    std::vector<char> d((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    size_t pos = 0;
    std::string magic, tw, th, tm;
    if (!next_token(d, pos, magic) || magic.size() != 2 || magic[0] != 'P' || std::string("2356").find(magic[1]) == std::string::npos) {
        std::cerr << "File " << filename << " can only contain P2, P3, P5 or P6 PGM images" << std::endl;
        return nullptr;
    }
    const int type = magic[1] - '0';
    if (!next_token(d, pos, tw) || !next_token(d, pos, th) || !next_token(d, pos, tm)) {
        std::cerr << "File " << filename << " is too short" << std::endl; return nullptr;
    }
    w = std::atoi(tw.c_str()); h = std::atoi(th.c_str());
    const int maxval = std::atoi(tm.c_str());
    if (w <= 0 || h <= 0 || maxval <= 0) { std::cerr << "File " << filename << " has meaningless image size" << std::endl; return nullptr; }
    const int chans = (type == 3 || type == 6) ? 3 : 1;
    const size_t n = (size_t)w * h * chans;
    std::vector<int> v(n);
    if (type == 2 || type == 3) {
        std::string t;
        for (size_t i = 0; i < n; i++) {
            if (!next_token(d, pos, t)) { std::cerr << "File " << filename << " file too short" << std::endl; return nullptr; }
            v[i] = std::atoi(t.c_str());
        }
    } else {
        pos += 1;   // single whitespace after maxval
        const size_t bps = maxval < 256 ? 1 : 2;
        if (pos + n * bps > d.size()) { std::cerr << "File " << filename << " file too short" << std::endl; return nullptr; }
        for (size_t i = 0; i < n; i++)
            v[i] = bps == 1 ? (unsigned char)d[pos + i]
                            : (int)((unsigned char)d[pos + 2 * i]) | ((int)((unsigned char)d[pos + 2 * i + 1]) << 8); // host order, as the reference reads it
    }
    auto scale = [&](int x) -> unsigned char { return maxval == 255 ? (unsigned char)x : (unsigned char)(x * 255.0 / maxval); };
    unsigned char* out = new unsigned char[(size_t)w * h];
    if (chans == 1) {
        for (size_t i = 0; i < (size_t)w * h; i++) out[i] = (type == 5 && maxval < 256) ? (unsigned char)v[i] : scale(v[i]);
    } else {
        // P3 values are scaled to 8 bits first (reference pgmread.cpp:146-155); P6 samples go into the
        // weights as stored, 8- or 16-bit, whatever maxval says (:198-246)
        const bool raw = (type == 6);
        for (size_t i = 0; i < (size_t)w * h; i++) {
            const unsigned r = raw ? (unsigned)v[3 * i] : (unsigned)scale(v[3 * i]);
            const unsigned g = raw ? (unsigned)v[3 * i + 1] : (unsigned)scale(v[3 * i + 1]);
            const unsigned b = raw ? (unsigned)v[3 * i + 2] : (unsigned)scale(v[3 * i + 2]);
            out[i] = to_gray(r, g, b);
        }
    }
    return out;
}
