// Config::LogMode::All ("--log"): the reference's debugging files, written from the slot's planes and results.
// Reference: PopSift::extractDownloadLoop (popsift.cpp:330-337) -> Pyramid::download_and_save_array
// (sift_pyramid.cu:79-86, sift_octave.cu:111-188) and Pyramid::save_descriptors (sift_pyramid.cu:88-106,401-444):
//   dir-octave/pyramid-o-O-l-L.pgm        Gaussian planes, ASCII PGM of the truncated values (write_plane_2d.cu:116-142)
//   dir-octave-dump/pyramid-o-O-l-L.dump  "floats\n<cols> <rows>\n" + raw float32 (write_plane_2d.cu:160-178)
//   dir-dog/d-pyramid-o-O-l-L.pgm         DoG planes scaled to [0, 255] (write_plane_2d.cu:50-96)
//   dir-dog-txt/d-pyramid-o-O-l-L.txt     DoG planes, truncated values + 127
//   dir-dog-dump/d-pyramid-o-O-l-L.dump   raw float32
//   dir-desc/desc-pyramid.txt, dir-fpt/desc-pyramid.txt   one line per (feature, orientation), setprecision(5)
// The raw dumps are the stage-level parity tap: tests read them with the same reader as the reference's.
#include "log_dump.h"

#include <cmath>
#include <fstream>
#include <iomanip>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include <sys/stat.h>

namespace popsift { namespace detail {

namespace {

void ensure_dir(const char* d)
{
    struct stat st;
    if (stat(d, &st) == -1) mkdir(d, 0700);
}

void write_unscaled(const std::string& fn, const std::vector<float>& p, int cols, int rows, int offset)
{
    std::ofstream of(fn, std::ios::binary);
    of << "P2" << std::endl << cols << " " << rows << std::endl << "255" << std::endl;
    for (int r = 0; r < rows; ++r) {
        for (int c = 0; c < cols; ++c) of << (int)p[(size_t)r * cols + c] + offset << " ";
        of << std::endl;
    }
}

void write_scaled(const std::string& fn, const std::vector<float>& p, int cols, int rows)
{
    float minval = std::numeric_limits<float>::max(), maxval = std::numeric_limits<float>::min();
    for (float v : p) { minval = std::min(minval, v); maxval = std::max(maxval, v); }
    const float f = 255.0f / (maxval - minval);
    std::ofstream of(fn, std::ios::binary);
    of << "P2" << std::endl << cols << " " << rows << std::endl << "255" << std::endl;
    for (int r = 0; r < rows; ++r) {
        for (int c = 0; c < cols; ++c) of << (int)(unsigned char)((p[(size_t)r * cols + c] - minval) * f) << " ";
        of << std::endl;
    }
}

void write_dump(const std::string& fn, const std::vector<float>& p, int cols, int rows)
{
    std::ofstream of(fn, std::ios::binary);
    of << "floats" << std::endl << cols << " " << rows << std::endl;
    of.write(reinterpret_cast<const char*>(p.data()), (std::streamsize)p.size() * sizeof(float));
}

void write_descriptors(std::ostream& ostr, const FeaturesHost& f, float up_fac, bool with_orientation)
{
    const float pi2 = 2.0f * 3.14159265358979323846f;
    const Feature* feat = const_cast<FeaturesHost&>(f).getFeatures();
    for (int i = 0; i < f.getFeatureCount(); ++i) {
        const Feature& e = feat[i];
        // the reference scales the (already image-space) coordinates once more (sift_pyramid.cu:409-412); kept as is
        const float s = std::pow(2.0f, (float)e.debug_octave - up_fac);
        const float xpos = e.xpos * s, ypos = e.ypos * s, sigma = e.sigma * s;
        for (int o = 0; o < e.num_ori; ++o) {
            float dom = e.orientation[o] / pi2 * 360;
            if (dom < 0) dom += 360;
            if (with_orientation) ostr << std::setprecision(5) << xpos << " " << ypos << " " << sigma << " " << dom << " ";
            else ostr << std::setprecision(5) << xpos << " " << ypos << " " << 1.0f / (sigma * sigma) << " 0 " << 1.0f / (sigma * sigma) << " ";
            for (float v : e.desc[o]->features) ostr << v << " ";
            ostr << std::endl;
        }
    }
}

} // namespace

bool dump_slot_planes(ps_ctx* ctx, int slot, int levels, const char* basename)
{
    int32_t n = 0, W[PS_MAX_OCTAVES], H[PS_MAX_OCTAVES];
    if (ps_slot_geometry(ctx, slot, &n, W, H) != PS_OK) return false;
    for (const char* d : {"dir-octave", "dir-octave-dump", "dir-dog", "dir-dog-txt", "dir-dog-dump"}) ensure_dir(d);
    std::vector<float> p;
    for (int o = 0; o < n; ++o) {
        p.resize((size_t)W[o] * H[o]);
        for (int l = 0; l < levels + 3; ++l) {
            if (ps_debug_plane(ctx, slot, o, l, PS_PLANE_GAUSS, p.data()) != PS_OK) return false;
            std::ostringstream a, b;
            a << "dir-octave/" << basename << "-o-" << o << "-l-" << l << ".pgm";
            b << "dir-octave-dump/" << basename << "-o-" << o << "-l-" << l << ".dump";
            write_unscaled(a.str(), p, W[o], H[o], 0);
            write_dump(b.str(), p, W[o], H[o]);
        }
        for (int l = 0; l < levels + 2; ++l) {
            if (ps_debug_plane(ctx, slot, o, l, PS_PLANE_DOG, p.data()) != PS_OK) return false;
            std::ostringstream a, b, c;
            a << "dir-dog/d-" << basename << "-o-" << o << "-l-" << l << ".pgm";
            b << "dir-dog-txt/d-" << basename << "-o-" << o << "-l-" << l << ".txt";
            c << "dir-dog-dump/d-" << basename << "-o-" << o << "-l-" << l << ".dump";
            write_scaled(a.str(), p, W[o], H[o]);
            write_unscaled(b.str(), p, W[o], H[o], 127);
            write_dump(c.str(), p, W[o], H[o]);
        }
    }
    return true;
}

void dump_descriptors(const FeaturesHost& f, float up_fac, const char* basename)
{
    if (f.getFeatureCount() == 0) return;
    ensure_dir("dir-desc");
    ensure_dir("dir-fpt");
    std::ofstream a(std::string("dir-desc/desc-") + basename + ".txt"), b(std::string("dir-fpt/desc-") + basename + ".txt");
    write_descriptors(a, f, up_fac, true);
    write_descriptors(b, f, up_fac, false);
}

}} // namespace popsift::detail
